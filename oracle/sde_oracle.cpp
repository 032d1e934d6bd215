// sde_oracle.cpp — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
//
// A plain C++17 restatement of the sde-sim-rs hot path (reference crate v0.5.1), used
// only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs as the checker and the reported CPU baseline.  Nothing under sde-sim-rs_b200/
// includes, links or calls this file.
//
// PARITY STATUS: *** parity unpinned ***.  The reference has no tests, no golden vectors
// and no seed parameter (src/sim/mod.rs:28-29), and neither Rust nor the crates it calls
// (sobol 1.0.2, rand 0.9.2, rand_chacha 0.9.0, fasteval 0.2.4) exist in this image, so
// this restatement cannot be checked against the Rust build.  What pins it instead:
//   * Sobol: scipy's copy of new-joe-kuo-6.21201 + scipy.stats.qmc.Sobol(bits=64)   (tests/test_oracle_sobol.py)
//   * ChaCha: `cryptography` ChaCha20 keystream (r=20) + rand_chacha's own KAT          (tests/test_oracle_chacha.py)
//   * icdf / Poisson: known answers computed with glibc (same libm Rust links to)        (tests/test_oracle_icdf.py)
//   * schemes: an independent pure-Python restatement (oracle/py_restatement.py)         (tests/test_oracle_schemes.py)
//   * committed fixtures: the independent pins above + frozen outputs per BASELINE config (tests/golden/, tests/test_golden.py)
// Two declared deviations from the reference: an explicit `seed` replaces
// rand::rng().random() (src/sim/mod.rs:28-29) and the Sobol point of scenario s is
// n = s + 5 (the single-thread order) instead of the Mutex race (src/sim/mod.rs:36-64).
//
// Every function cites the reference lines it follows.

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace orc {

// ---------------------------------------------------------------------------------------
// Sobol (third-party crate `sobol` 1.0.2, Cargo.toml:29; call sites src/rng/sobol.rs:15-25)
// Published algorithm: Joe & Kuo direction numbers, Antonov–Saleev Gray-code iteration,
// 64-bit integers, render x / 2^64.
// ---------------------------------------------------------------------------------------

// v[d][i] (i = 0-based bit, i.e. direction number i+1) = m_{i+1} << (64 - (i+1)).
// Dim 0 is van der Corput (all m = 1); dims >= 1 use (poly, m_init) from Joe–Kuo.
// poly packs the full primitive polynomial: s = bit_length(poly) - 1, a_k = bit (s-k).
static void direction_numbers(const uint32_t* poly, const uint32_t* minit, int minit_stride,
                              int dims, uint64_t* V /* [dims][64] */) {
    for (int d = 0; d < dims; ++d) {
        uint64_t m[64];
        if (d == 0) {
            for (int i = 0; i < 64; ++i) m[i] = 1;
        } else {
            uint32_t p = poly[d];
            int s = 0;
            while ((p >> (s + 1)) != 0) ++s;  // degree
            for (int i = 0; i < s && i < 64; ++i) m[i] = minit[(size_t)d * minit_stride + i];
            for (int i = s; i < 64; ++i) {
                uint64_t nv = m[i - s] ^ (m[i - s] << s);
                for (int k = 1; k < s; ++k) {
                    if ((p >> (s - k)) & 1u) nv ^= m[i - k] << k;
                }
                m[i] = nv;
            }
        }
        for (int i = 0; i < 64; ++i) V[(size_t)d * 64 + i] = m[i] << (63 - i);
    }
}

// Iterator restatement: point 0 is the origin, x_{n+1} = x_n ^ v[ctz(~n)].
static void points_sequential(const uint64_t* V, int dims, uint64_t first, uint64_t count,
                              uint64_t* out /* [count][dims] */) {
    std::vector<uint64_t> x(dims, 0);
    uint64_t n = 0;
    auto advance = [&]() {
        int c = 0;
        uint64_t t = n;
        while (t & 1) { t >>= 1; ++c; }
        for (int d = 0; d < dims; ++d) x[d] ^= V[(size_t)d * 64 + c];
        ++n;
    };
    while (n < first) advance();
    for (uint64_t i = 0; i < count; ++i) {
        std::memcpy(out + i * dims, x.data(), sizeof(uint64_t) * dims);
        advance();
    }
}

// Closed form of the same sequence: x_n = XOR_{b in bits(n ^ (n>>1))} v[b].
static inline uint64_t point_direct(const uint64_t* Vd, uint64_t n) {
    uint64_t g = n ^ (n >> 1), x = 0;
    for (int b = 0; g; ++b, g >>= 1)
        if (g & 1) x ^= Vd[b];
    return x;
}

// ---------------------------------------------------------------------------------------
// ChaCha (third-party crates rand_chacha 0.9.0 / rand 0.9.2, Cargo.toml:25-26; call sites
// src/rng/pseudo.rs:18,25 and src/rng/sobol.rs:68-69)
// ---------------------------------------------------------------------------------------
static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

static void chacha_block(const uint32_t key[8], uint64_t counter, uint64_t stream, int rounds,
                         uint32_t out[16]) {
    uint32_t s[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u,
                      key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7],
                      (uint32_t)counter, (uint32_t)(counter >> 32),
                      (uint32_t)stream, (uint32_t)(stream >> 32)};
    uint32_t w[16];
    std::memcpy(w, s, sizeof w);
#define ORC_QR(a, b, c, d)                                                     \
    w[a] += w[b]; w[d] ^= w[a]; w[d] = rotl32(w[d], 16);                       \
    w[c] += w[d]; w[b] ^= w[c]; w[b] = rotl32(w[b], 12);                       \
    w[a] += w[b]; w[d] ^= w[a]; w[d] = rotl32(w[d], 8);                        \
    w[c] += w[d]; w[b] ^= w[c]; w[b] = rotl32(w[b], 7);
    for (int r = 0; r < rounds; r += 2) {
        ORC_QR(0, 4, 8, 12) ORC_QR(1, 5, 9, 13) ORC_QR(2, 6, 10, 14) ORC_QR(3, 7, 11, 15)
        ORC_QR(0, 5, 10, 15) ORC_QR(1, 6, 11, 12) ORC_QR(2, 7, 8, 13) ORC_QR(3, 4, 9, 14)
    }
#undef ORC_QR
    for (int i = 0; i < 16; ++i) out[i] = w[i] + s[i];
}

// rand_core SeedableRng::seed_from_u64: PCG32 output fills the 32-byte seed, 4 bytes at a time.
static void seed_from_u64(uint64_t state, uint32_t key[8]) {
    const uint64_t MUL = 6364136223846793005ull, INC = 11634580027462260723ull;
    for (int i = 0; i < 8; ++i) {
        state = state * MUL + INC;
        uint32_t xorshifted = (uint32_t)(((state >> 18) ^ state) >> 27);
        uint32_t rot = (uint32_t)(state >> 59);
        key[i] = (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));
    }
}

// ChaCha8Rng::seed_from_u64(seed) viewed as a stream of `random::<f64>()`:
// next_u64 = w[2i] | w[2i+1] << 32 ; f64 = (u64 >> 11) * 2^-53.
struct ChaCha8F64 {
    uint32_t key[8];
    uint32_t buf[16];
    uint64_t block = 0;
    int pos = 16;
    explicit ChaCha8F64(uint64_t seed) { seed_from_u64(seed, key); }
    uint64_t next_u64() {
        if (pos >= 16) { chacha_block(key, block++, 0, 8, buf); pos = 0; }
        uint64_t v = (uint64_t)buf[pos] | ((uint64_t)buf[pos + 1] << 32);
        pos += 2;
        return v;
    }
    double next_f64() { return (double)(next_u64() >> 11) * (1.0 / 9007199254740992.0); }
};

// ---------------------------------------------------------------------------------------
// Inverse CDFs — src/proc/increment.rs:160-200 (in-tree, exact restatement)
// ---------------------------------------------------------------------------------------
static inline double icdf_normal(double p) {  // increment.rs:161-179
    double t = (p < 0.5) ? std::sqrt(-2.0 * std::log(p)) : std::sqrt(-2.0 * std::log(1.0 - p));
    const double c0 = 2.515517, c1 = 0.802853, c2 = 0.010328;
    const double d1 = 1.432788, d2 = 0.189269, d3 = 0.001308;
    double x = t - ((c2 * t + c1) * t + c0) / (((d3 * t + d2) * t + d1) * t + 1.0);
    return (p < 0.5) ? -x : x;
}

static inline uint64_t icdf_poisson(double u, double lambda) {  // increment.rs:182-200
    if (lambda <= 0.0) return 0;
    double p = std::exp(-lambda), f = p;
    uint64_t k = 0;
    while (u > f && k < 200) {
        k += 1;
        p *= lambda / (double)k;
        f += p;
    }
    return k;
}

// ---------------------------------------------------------------------------------------
// Expressions — restatement of the fasteval 0.2.4 subset the reference can reach through
// src/func.rs:18-42.  [3P-unverified]: precedence (lowest→highest) or, and, comparisons,
// +, -, *, /, %, ^ (right-assoc); unary -, +, ! bind to the following value; builtins as
// listed in eval_func.  Evaluation uses plain IEEE ops (fasteval's internal a*(1/b) and
// constant folding can differ by an ulp; that is inside the 1e-12 budget and documented).
// ---------------------------------------------------------------------------------------
enum NodeKind : int {
    N_CONST, N_VAR_T, N_VAR_P, N_NEG, N_NOT, N_ADD, N_SUB, N_MUL, N_DIV, N_MOD, N_POW,
    N_LT, N_GT, N_LE, N_GE, N_EQ, N_NE, N_AND, N_OR, N_FUNC
};
enum FuncId : int {
    F_INT, F_CEIL, F_FLOOR, F_ABS, F_SIGN, F_LOG, F_ROUND, F_MIN, F_MAX, F_E, F_PI,
    F_SIN, F_COS, F_TAN, F_ASIN, F_ACOS, F_ATAN, F_SINH, F_COSH, F_TANH, F_ASINH, F_ACOSH, F_ATANH
};
struct Node {
    int kind = N_CONST;
    double value = 0.0;
    int slot = -1;        // N_VAR_P: process index
    int func = -1;        // N_FUNC
    std::vector<int> kids;
};
struct Expr {
    std::vector<Node> nodes;
    int root = -1;
    std::string src;
};

struct ParseError { std::string msg; };

class ExprParser {
  public:
    ExprParser(const std::string& s, const std::unordered_map<std::string, int>* vars, Expr* out)
        : s_(s), vars_(vars), e_(out) {}
    void run() {
        e_->src = s_;
        e_->root = parse_expression();
        skip_ws();
        if (i_ != s_.size()) throw ParseError{"unparsed tokens remaining: '" + s_.substr(i_) + "'"};
    }
    // names referenced that were not resolvable (reported by caller)
    std::vector<std::string> unresolved;

  private:
    std::string s_;   // by value: callers pass temporaries
    const std::unordered_map<std::string, int>* vars_;
    Expr* e_;
    size_t i_ = 0;

    void skip_ws() { while (i_ < s_.size() && std::isspace((unsigned char)s_[i_])) ++i_; }
    int add(Node n) { e_->nodes.push_back(std::move(n)); return (int)e_->nodes.size() - 1; }

    // binary operator levels; each fasteval op is its own level except comparisons.
    static int level(int kind) {
        switch (kind) {
            case N_OR: return 1; case N_AND: return 2;
            case N_NE: case N_EQ: case N_GE: case N_LE: case N_GT: case N_LT: return 3;
            case N_ADD: return 4; case N_SUB: return 5; case N_MUL: return 6;
            case N_DIV: return 7; case N_MOD: return 8; case N_POW: return 9;
        }
        return 100;
    }
    bool read_binop(int* kind) {
        skip_ws();
        if (i_ >= s_.size()) return false;
        auto starts = [&](const char* t) { return s_.compare(i_, std::strlen(t), t) == 0; };
        auto word = [&](const char* t) {
            size_t n = std::strlen(t);
            if (s_.compare(i_, n, t) != 0) return false;
            if (i_ + n < s_.size() && (std::isalnum((unsigned char)s_[i_ + n]) || s_[i_ + n] == '_')) return false;
            return true;
        };
        if (starts("||")) { *kind = N_OR; i_ += 2; return true; }
        if (starts("&&")) { *kind = N_AND; i_ += 2; return true; }
        if (word("or")) { *kind = N_OR; i_ += 2; return true; }
        if (word("and")) { *kind = N_AND; i_ += 3; return true; }
        if (starts("!=")) { *kind = N_NE; i_ += 2; return true; }
        if (starts("==")) { *kind = N_EQ; i_ += 2; return true; }
        if (starts(">=")) { *kind = N_GE; i_ += 2; return true; }
        if (starts("<=")) { *kind = N_LE; i_ += 2; return true; }
        char c = s_[i_];
        int k = -1;
        switch (c) {
            case '>': k = N_GT; break; case '<': k = N_LT; break;
            case '+': k = N_ADD; break; case '-': k = N_SUB; break;
            case '*': k = N_MUL; break; case '/': k = N_DIV; break;
            case '%': k = N_MOD; break; case '^': k = N_POW; break;
        }
        if (k < 0) return false;
        *kind = k; ++i_;
        return true;
    }

    int parse_expression() {
        std::vector<int> vals;
        std::vector<int> ops;
        vals.push_back(parse_value());
        int k;
        while (read_binop(&k)) {
            ops.push_back(k);
            vals.push_back(parse_value());
        }
        return build(vals, ops, 0, (int)vals.size());
    }

    // Build values[lo,hi) joined by ops[lo,hi-1) : split at the lowest-precedence level.
    int build(const std::vector<int>& vals, const std::vector<int>& ops, int lo, int hi) {
        if (hi - lo == 1) return vals[lo];
        int lowest = 1000;
        for (int j = lo; j < hi - 1; ++j) lowest = std::min(lowest, level(ops[j]));
        std::vector<int> cut;  // op positions at the lowest level
        for (int j = lo; j < hi - 1; ++j) if (level(ops[j]) == lowest) cut.push_back(j);
        if (lowest == 9) {  // '^' : right-to-left
            int rhs = build(vals, ops, cut.back() + 1, hi);
            for (int c = (int)cut.size() - 1; c >= 0; --c) {
                int seg_lo = (c == 0) ? lo : cut[c - 1] + 1;
                int lhs = build(vals, ops, seg_lo, cut[c] + 1);
                Node n; n.kind = N_POW; n.kids = {lhs, rhs};
                rhs = add(n);
            }
            return rhs;
        }
        int acc = build(vals, ops, lo, cut[0] + 1);
        for (size_t c = 0; c < cut.size(); ++c) {
            int seg_hi = (c + 1 < cut.size()) ? cut[c + 1] + 1 : hi;
            int rhs = build(vals, ops, cut[c] + 1, seg_hi);
            Node n; n.kind = ops[cut[c]]; n.kids = {acc, rhs};
            acc = add(n);
        }
        return acc;
    }

    int parse_value() {
        skip_ws();
        if (i_ >= s_.size()) throw ParseError{"unexpected end of expression"};
        char c = s_[i_];
        if (c == '(') {
            ++i_;
            int v = parse_expression();
            skip_ws();
            if (i_ >= s_.size() || s_[i_] != ')') throw ParseError{"missing ')'"};
            ++i_;
            return v;
        }
        if (c == '-') { ++i_; Node n; n.kind = N_NEG; n.kids = {parse_value()}; return add(n); }
        if (c == '+') { ++i_; return parse_value(); }
        if (c == '!') { ++i_; Node n; n.kind = N_NOT; n.kids = {parse_value()}; return add(n); }
        if (std::isdigit((unsigned char)c) || c == '.') return parse_number();
        if (std::isalpha((unsigned char)c) || c == '_') return parse_ident();
        throw ParseError{std::string("unexpected character '") + c + "'"};
    }

    int parse_number() {
        size_t st = i_;
        while (i_ < s_.size() && (std::isdigit((unsigned char)s_[i_]) || s_[i_] == '.')) ++i_;
        if (i_ < s_.size() && (s_[i_] == 'e' || s_[i_] == 'E')) {
            size_t j = i_ + 1;
            if (j < s_.size() && (s_[j] == '+' || s_[j] == '-')) ++j;
            if (j < s_.size() && std::isdigit((unsigned char)s_[j])) {
                while (j < s_.size() && std::isdigit((unsigned char)s_[j])) ++j;
                i_ = j;
            }
        }
        std::string tok = s_.substr(st, i_ - st);
        char* endp = nullptr;
        double v = std::strtod(tok.c_str(), &endp);
        if (endp == tok.c_str() || *endp != '\0') throw ParseError{"bad number '" + tok + "'"};
        // SI-style suffixes
        if (i_ < s_.size()) {
            double mul = 0.0;
            size_t adv = 1;
            switch (s_[i_]) {
                case 'k': case 'K': mul = 1e3; break; case 'M': mul = 1e6; break;
                case 'G': mul = 1e9; break; case 'T': mul = 1e12; break;
                case 'm': mul = 1e-3; break; case 'u': mul = 1e-6; break;
                case 'n': mul = 1e-9; break; case 'p': mul = 1e-12; break;
                default: break;
            }
            if (mul == 0.0 && s_.compare(i_, 2, "\xC2\xB5") == 0) { mul = 1e-6; adv = 2; }
            if (mul != 0.0) {
                size_t j = i_ + adv;
                bool ident_follows = j < s_.size() && (std::isalnum((unsigned char)s_[j]) || s_[j] == '_');
                if (!ident_follows) { v *= mul; i_ = j; }
            }
        }
        Node n; n.kind = N_CONST; n.value = v;
        return add(n);
    }

    int parse_ident() {
        size_t st = i_;
        while (i_ < s_.size() && (std::isalnum((unsigned char)s_[i_]) || s_[i_] == '_')) ++i_;
        std::string name = s_.substr(st, i_ - st);
        size_t save = i_;
        skip_ws();
        if (i_ < s_.size() && s_[i_] == '(') {
            static const std::map<std::string, int> funcs = {
                {"int", F_INT}, {"ceil", F_CEIL}, {"floor", F_FLOOR}, {"abs", F_ABS}, {"sign", F_SIGN},
                {"log", F_LOG}, {"round", F_ROUND}, {"min", F_MIN}, {"max", F_MAX}, {"e", F_E}, {"pi", F_PI},
                {"sin", F_SIN}, {"cos", F_COS}, {"tan", F_TAN}, {"asin", F_ASIN}, {"acos", F_ACOS},
                {"atan", F_ATAN}, {"sinh", F_SINH}, {"cosh", F_COSH}, {"tanh", F_TANH},
                {"asinh", F_ASINH}, {"acosh", F_ACOSH}, {"atanh", F_ATANH}};
            auto it = funcs.find(name);
            if (it == funcs.end()) throw ParseError{"unsupported function '" + name + "'"};
            ++i_;
            Node n; n.kind = N_FUNC; n.func = it->second;
            skip_ws();
            if (i_ < s_.size() && s_[i_] == ')') { ++i_; }
            else {
                for (;;) {
                    n.kids.push_back(parse_expression());
                    skip_ws();
                    if (i_ < s_.size() && s_[i_] == ',') { ++i_; continue; }
                    if (i_ < s_.size() && s_[i_] == ')') { ++i_; break; }
                    throw ParseError{"missing ')' in call to " + name};
                }
            }
            size_t argc = n.kids.size();
            auto need = [&](size_t lo, size_t hi) {
                if (argc < lo || argc > hi) throw ParseError{"wrong number of arguments to " + name};
            };
            switch (n.func) {
                case F_E: case F_PI: need(0, 0); break;
                case F_LOG: case F_ROUND: need(1, 2); break;
                case F_MIN: case F_MAX: need(1, 1000); break;
                default: need(1, 1);
            }
            return add(n);
        }
        i_ = save;
        Node n;
        auto it = vars_->find(name);
        if (it != vars_->end()) { n.kind = N_VAR_P; n.slot = it->second; }   // process names shadow "t" (filtration.rs:72-78)
        else if (name == "t") { n.kind = N_VAR_T; }
        else { unresolved.push_back(name); n.kind = N_CONST; n.value = std::numeric_limits<double>::quiet_NaN(); }
        return add(n);
    }
};

static double eval_node(const Expr& e, int idx, double t, const double* vals) {
    const Node& n = e.nodes[idx];
    auto K = [&](int k) { return eval_node(e, n.kids[k], t, vals); };
    switch (n.kind) {
        case N_CONST: return n.value;
        case N_VAR_T: return t;
        case N_VAR_P: return vals[n.slot];
        case N_NEG: return -K(0);
        case N_NOT: return (std::fabs(K(0)) <= 8.0 * std::numeric_limits<double>::epsilon()) ? 1.0 : 0.0;
        case N_ADD: return K(0) + K(1);
        case N_SUB: return K(0) - K(1);
        case N_MUL: return K(0) * K(1);
        case N_DIV: return K(0) / K(1);
        case N_MOD: return std::fmod(K(0), K(1));
        case N_POW: return std::pow(K(0), K(1));
        case N_LT: return K(0) < K(1) ? 1.0 : 0.0;
        case N_GT: return K(0) > K(1) ? 1.0 : 0.0;
        case N_LE: return K(0) <= K(1) ? 1.0 : 0.0;
        case N_GE: return K(0) >= K(1) ? 1.0 : 0.0;
        case N_EQ: return (std::fabs(K(0) - K(1)) <= 8.0 * std::numeric_limits<double>::epsilon()) ? 1.0 : 0.0;
        case N_NE: return (std::fabs(K(0) - K(1)) <= 8.0 * std::numeric_limits<double>::epsilon()) ? 0.0 : 1.0;
        case N_AND: { double l = K(0); if (std::fabs(l) <= 8.0 * std::numeric_limits<double>::epsilon()) return l; return K(1); }
        case N_OR: { double l = K(0); if (!(std::fabs(l) <= 8.0 * std::numeric_limits<double>::epsilon())) return l; return K(1); }
        case N_FUNC: {
            switch (n.func) {
                case F_INT: return std::trunc(K(0));
                case F_CEIL: return std::ceil(K(0));
                case F_FLOOR: return std::floor(K(0));
                case F_ABS: return std::fabs(K(0));
                case F_SIGN: { double x = K(0); if (std::isnan(x)) return x; return std::signbit(x) ? -1.0 : 1.0; }
                case F_LOG: if (n.kids.size() == 1) return std::log10(K(0)); else { double b = K(0), x = K(1); return std::log(x) / std::log(b); }
                case F_ROUND: if (n.kids.size() == 1) return std::round(K(0)); else { double m = K(0), x = K(1); return std::round(x / m) * m; }
                case F_MIN: case F_MAX: {
                    double acc = K(0); bool nan = std::isnan(acc);
                    for (size_t k = 1; k < n.kids.size(); ++k) {
                        double x = K((int)k);
                        nan = nan || std::isnan(x);
                        if (n.func == F_MIN ? (x < acc) : (x > acc)) acc = x;
                    }
                    return nan ? std::numeric_limits<double>::quiet_NaN() : acc;
                }
                case F_E: return 2.718281828459045;
                case F_PI: return 3.141592653589793;
                case F_SIN: return std::sin(K(0)); case F_COS: return std::cos(K(0)); case F_TAN: return std::tan(K(0));
                case F_ASIN: return std::asin(K(0)); case F_ACOS: return std::acos(K(0)); case F_ATAN: return std::atan(K(0));
                case F_SINH: return std::sinh(K(0)); case F_COSH: return std::cosh(K(0)); case F_TANH: return std::tanh(K(0));
                case F_ASINH: return std::asinh(K(0)); case F_ACOSH: return std::acosh(K(0)); case F_ATANH: return std::atanh(K(0));
            }
        }
    }
    return std::numeric_limits<double>::quiet_NaN();
}

// ---------------------------------------------------------------------------------------
// Model — src/proc/mod.rs:7-90, src/proc/util.rs:16-166
// ---------------------------------------------------------------------------------------
enum IncKind { INC_DT, INC_DW, INC_DN };
struct Term {
    std::string coeff_src;
    Expr coeff;
    IncKind kind = INC_DT;
    int idx = -1;               // stochastic registry index (dW / dN)
    std::string lambda_src;     // dN only
    Expr lambda;
};
struct Process {
    std::string name;
    bool levy = false;
    std::vector<Term> terms;    // levy
    std::string alg_src;        // algebraic
    Expr alg;
};
struct Universe {
    std::vector<Process> procs;
    std::unordered_map<std::string, int> process_registry;  // later duplicates win (HashMap::insert)
    std::vector<std::string> stochastic_names;               // registry in first-appearance order
    std::vector<int> levy_idx, alg_idx;
    std::vector<double> times;
};

// util.rs:16-38 + the `delimited(char('('), balanced_parens, char(')'))` wrapper used at
// :24,:46,:89,:102.  `s` must start with '('.  Returns false when unbalanced.
static bool delimited_balanced(const std::string& s, size_t start, size_t* end_after, std::string* inside) {
    if (start >= s.size() || s[start] != '(') return false;
    int depth = 0;
    for (size_t j = start; j < s.size(); ++j) {
        if (s[j] == '(') ++depth;
        else if (s[j] == ')') {
            --depth;
            if (depth == 0) {
                *inside = s.substr(start + 1, j - start - 1);
                *end_after = j + 1;
                return true;
            }
        }
    }
    return false;
}
static std::string trim(const std::string& s) {
    size_t a = 0, b = s.size();
    while (a < b && std::isspace((unsigned char)s[a])) ++a;
    while (b > a && std::isspace((unsigned char)s[b - 1])) --b;
    return s.substr(a, b - a);
}
static std::string trim_start(const std::string& s) {
    size_t a = 0;
    while (a < s.size() && std::isspace((unsigned char)s[a])) ++a;
    return s.substr(a);
}

struct PendingExpr { Expr* dst; std::string src; std::string what; };

static void parse_single_equation(const std::string& equation, Universe& u,
                                  std::unordered_map<std::string, int>& stoch_reg,
                                  std::vector<std::pair<std::pair<int,int>, int>>& /*unused*/) {
    // util.rs:73-76
    std::vector<std::string> parts;
    {
        size_t st = 0;
        for (;;) {
            size_t p = equation.find('=', st);
            if (p == std::string::npos) { parts.push_back(equation.substr(st)); break; }
            parts.push_back(equation.substr(st, p - st));
            st = p + 1;
        }
    }
    if (parts.size() != 2) throw ParseError{"Missing '='"};
    std::string lhs = trim(parts[0]), rhs = trim(parts[1]);
    Process pr;
    bool is_sde = !lhs.empty() && lhs[0] == 'd';          // util.rs:80-82
    pr.name = is_sde ? lhs.substr(1) : lhs;
    pr.levy = is_sde;
    if (is_sde) {
        std::string cur = rhs;
        for (;;) {                                         // util.rs:87-123
            size_t start = cur.find('(');
            if (start == std::string::npos) break;
            size_t after = 0; std::string coeff;
            if (!delimited_balanced(cur, start, &after, &coeff)) throw ParseError{"Unbalanced parentheses in coefficient"};
            std::string trimmed_after = trim_start(cur.substr(after));
            if (trimmed_after.empty() || trimmed_after[0] != '*') break;
            std::string after_star = trim_start(trimmed_after.substr(1));
            std::string inc, remaining;
            if (after_star.compare(0, 2, "dN") == 0) {
                size_t dstart = after_star.find('(');
                if (dstart == std::string::npos) throw ParseError{"dN missing opening bracket"};
                size_t dend = 0; std::string inside;
                if (!delimited_balanced(after_star, dstart, &dend, &inside)) throw ParseError{"Unbalanced parentheses in dN intensity"};
                inc = after_star.substr(0, dend);
                remaining = after_star.substr(dend);
            } else {
                size_t end = after_star.find(' ');
                if (end == std::string::npos) end = after_star.size();
                inc = after_star.substr(0, end);
                remaining = after_star.substr(end);
            }
            Term t;
            t.coeff_src = trim(coeff);
            // build_incrementor, util.rs:136-166
            if (inc == "dt") { t.kind = INC_DT; }
            else {
                int next = (int)stoch_reg.size();
                auto ins = stoch_reg.emplace(inc, next);
                if (ins.second) u.stochastic_names.push_back(inc);
                t.idx = ins.first->second;
                if (inc.compare(0, 2, "dW") == 0) t.kind = INC_DW;
                else if (inc.compare(0, 2, "dN") == 0) {
                    t.kind = INC_DN;
                    size_t b = inc.find('(');
                    if (b == std::string::npos) throw ParseError{"Missing '(' in dN incrementor"};
                    size_t e2 = 0; std::string content;
                    if (!delimited_balanced(inc, b, &e2, &content)) throw ParseError{"Unbalanced parentheses in jump term: " + inc};
                    t.lambda_src = trim(content);
                } else throw ParseError{"Unknown incrementor type: " + inc};
            }
            pr.terms.push_back(std::move(t));
            cur = remaining;
        }
    } else {
        pr.alg_src = rhs;                                   // util.rs:127-133
    }
    u.procs.push_back(std::move(pr));
}

static void compile_expr(const std::string& src, const Universe& u, Expr* dst, const std::string& what) {
    ExprParser p(src, &u.process_registry, dst);
    try { p.run(); } catch (ParseError& e) { throw ParseError{what + ": Parse Error: " + e.msg + " in '" + src + "'"}; }
    if (!p.unresolved.empty())
        // reference: panics at the first eval (euler.rs:22 `.unwrap()`); the oracle reports it at parse time.
        throw ParseError{what + ": undefined variable '" + p.unresolved[0] + "' in '" + src + "'"};
}

static Universe* parse_equations(const std::vector<std::string>& eqs, const std::vector<double>& times) {
    auto u = std::make_unique<Universe>();
    u->times = times;
    std::unordered_map<std::string, int> stoch_reg;
    std::vector<std::pair<std::pair<int,int>, int>> dummy;
    for (auto& eq : eqs) parse_single_equation(eq, *u, stoch_reg, dummy);
    // ProcessUniverse::new, mod.rs:71-89
    for (size_t i = 0; i < u->procs.size(); ++i) {
        u->process_registry[u->procs[i].name] = (int)i;
        (u->procs[i].levy ? u->levy_idx : u->alg_idx).push_back((int)i);
    }
    // Expressions are resolved against the final registry (the reference resolves names
    // lazily at eval time against the cache map, filtration.rs:72-78 — same result).
    for (auto& p : u->procs) {
        if (p.levy) {
            for (auto& t : p.terms) {
                compile_expr(t.coeff_src, *u, &t.coeff, "Math error in coefficient");
                if (t.kind == INC_DN) compile_expr(t.lambda_src, *u, &t.lambda, "Math error in jump lambda '" + t.lambda_src + "'");
            }
        } else {
            compile_expr(p.alg_src, *u, &p.alg, "Parse Error");
        }
    }
    return u.release();
}

// ---------------------------------------------------------------------------------------
// Filtration — src/filtration.rs:12-79 (dense rows + the time-keyed evaluation cache)
// ---------------------------------------------------------------------------------------
struct Filtration {
    const Universe* u;
    int P, T;
    std::vector<double> raw;          // [T][P], zero-initialised (filtration.rs:28)
    double cache_time;
    std::vector<double> cache;        // one slot per process *name* = per registry entry; indexed by process idx
    const std::unordered_map<uint64_t, int>* time_registry;

    double get(int t, int p) const { return raw[(size_t)t * P + p]; }
    void set(int t, int p, double v) { raw[(size_t)t * P + p] = v; }
    void refresh_cache(double time) {                     // filtration.rs:70-79
        cache_time = time;
        uint64_t bits; std::memcpy(&bits, &time, 8);
        auto it = time_registry->find(bits);
        int t_idx = (it == time_registry->end()) ? 0 : it->second;
        for (auto& kv : u->process_registry) cache[kv.second] = get(t_idx, kv.second);
    }
    double eval(const Expr& f, double time) {             // func.rs:32-42
        if (time != cache_time) refresh_cache(time);
        return eval_node(f, f.root, cache_time, cache.data());   // "t" is the cache's time entry (filtration.rs:72)
    }
};

// RNG trait — src/rng/mod.rs:5-7
struct Rng {
    virtual ~Rng() {}
    virtual double sample(int time_idx, int inc_idx) = 0;
};
struct PseudoRng : Rng {                                   // src/rng/pseudo.rs:7-60
    ChaCha8F64 rng; int K; int last_t = -1; std::vector<double> vals;
    PseudoRng(uint64_t seed, int K_) : rng(seed), K(K_) {}
    double sample(int t, int k) override {
        if (last_t != t) { vals.resize(K); for (int i = 0; i < K; ++i) vals[i] = rng.next_f64(); last_t = t; }
        if (k >= K) { std::fprintf(stderr, "RNG Index %d out of bounds (max %d)\n", k, K); std::abort(); }
        return vals[k];
    }
};
struct TableRng : Rng {                                    // SobolRng::sample, src/rng/sobol.rs:56-60
    std::vector<double> values; int K;
    double sample(int t, int k) override { return values[(size_t)t * K + k]; }
};
// Test hook: injected per-(t,k) values; for Wiener factors the value is the normal z itself.
struct InjectRng : Rng {
    const double* v; int K1;   // row = K entries + u0 for RK's sk
    double sample(int t, int k) override { return v[(size_t)t * K1 + k]; }
};

struct SimOptions {
    uint64_t seed = 0;
    uint64_t scenario_offset = 0;   // global scenario index of local scenario 0 (multi-GPU shard parity)
    int rng_mode = 0;               // 0 pseudo (ChaCha8), 1 sobol + per-path CP shift (reference), 2 sobol + XOR digital shift, 3 sobol unscrambled, 4 injected
    int scheme = 0;                 // 0 euler, 1 runge-kutta (reference semantics), 2 runge-kutta textbook (k1 at settled state)
    const uint64_t* sobol_V = nullptr;  // [dims][64]
    const double* inject = nullptr;     // [N][S][K+1]
    int nthreads = 0;
};

// XOR digital shift (this build's own RQMC mode; README.md:13 describes it, the reference code does not
// have it): the 32-bit Sobol integer (top half of the u64; exact for n < 2^32) is XORed with a per-dimension
// 32-bit mask = top half of u64 #d of ChaCha8Rng::seed_from_u64(seed); u = (k + 1/2) * 2^-32.
static inline double xor_uniform(uint64_t x, uint64_t mask) {
    uint64_t k = (x ^ mask) >> 32;
    return ((double)k + 0.5) * (1.0 / 4294967296.0);
}

static inline double wiener_sample(Rng& rng, int t, int idx, double sqrt_dt, bool injected) {  // increment.rs:89-97
    double q = rng.sample(t, idx);
    return sqrt_dt * (injected ? q : icdf_normal(q));
}

static void euler_iteration(Filtration& F, const Universe& U, int t, Rng& rng, const std::vector<double>& dts,
                            const std::vector<double>& sqrt_dts, bool injected) {  // src/sim/euler.rs:5-37
    double cur = U.times[t], nxt = U.times[t + 1];
    for (int p : U.levy_idx) {
        const Process& pr = U.procs[p];
        double val = F.get(t, p);
        for (const Term& tm : pr.terms) {
            double c = F.eval(tm.coeff, cur);
            double x;
            if (tm.kind == INC_DT) x = dts[t];
            else if (tm.kind == INC_DW) x = wiener_sample(rng, t, tm.idx, sqrt_dts[t], injected);
            else {                                          // increment.rs:137-148
                double uu = rng.sample(t, tm.idx);
                double lam = F.eval(tm.lambda, U.times[t]) * dts[t];
                x = (double)icdf_poisson(uu, lam);
            }
            val += c * x;
        }
        F.set(t + 1, p, val);
    }
    for (int p : U.alg_idx) F.set(t + 1, p, F.eval(U.procs[p].alg, nxt));
}

static void rk_iteration(Filtration& F, const Universe& U, int t, Rng& rng, const std::vector<double>& dts,
                         const std::vector<double>& sqrt_dts, bool injected, bool textbook, int K) {  // src/sim/runge_kutta.rs:5-107
    int P = F.P;
    double cur = U.times[t], nxt = U.times[t + 1];
    double dt = nxt - cur;
    double sqrt_dt = std::sqrt(dt);
    double u0 = injected ? rng.sample(t, K) : rng.sample(t, 0);   // :18 (injected rows carry u0 in slot K)
    double sk = (u0 > 0.5) ? 1.0 : -1.0;
    if (textbook) F.refresh_cache(cur);                   // "textbook" variant: k1 at the settled row
    std::vector<std::vector<double>> inc(P);
    for (int p = 0; p < P; ++p) {                          // :26-35
        const Process& pr = U.procs[p];
        if (!pr.levy) continue;
        for (const Term& tm : pr.terms) {
            double x;
            if (tm.kind == INC_DT) x = dts[t];
            else if (tm.kind == INC_DW) x = wiener_sample(rng, t, tm.idx, sqrt_dts[t], injected);
            else {
                double uu = rng.sample(t, tm.idx);
                double lam = F.eval(tm.lambda, U.times[t]) * dts[t];
                x = (double)icdf_poisson(uu, lam);
            }
            inc[p].push_back(x);
        }
    }
    std::vector<double> x_t(P), k1(P, 0.0), k2(P, 0.0);
    for (int p = 0; p < P; ++p) x_t[p] = F.get(t, p);     // :38-42
    for (int p = 0; p < P; ++p) {                          // :45-55
        const Process& pr = U.procs[p];
        if (!pr.levy) continue;
        for (size_t j = 0; j < pr.terms.size(); ++j) k1[p] += F.eval(pr.terms[j].coeff, cur) * inc[p][j];
    }
    for (int p = 0; p < P; ++p) {                          // :62-78
        const Process& pr = U.procs[p];
        if (!pr.levy) continue;
        double pert = 0.0;
        for (size_t j = 0; j < pr.terms.size(); ++j)
            if (pr.terms[j].kind == INC_DW) pert += F.eval(pr.terms[j].coeff, cur) * sk * sqrt_dt;
        F.set(t + 1, p, x_t[p] + k1[p] + pert);
    }
    for (int p = 0; p < P; ++p) {                          // :81-91
        const Process& pr = U.procs[p];
        if (!pr.levy) continue;
        for (size_t j = 0; j < pr.terms.size(); ++j) k2[p] += F.eval(pr.terms[j].coeff, nxt) * inc[p][j];
    }
    for (int p : U.levy_idx) F.set(t + 1, p, x_t[p] + 0.5 * (k1[p] + k2[p]));   // :94-97
    if (textbook && !U.alg_idx.empty()) F.refresh_cache(nxt);
    for (int p : U.alg_idx) F.set(t + 1, p, F.eval(U.procs[p].alg, nxt));       // :101-106
}

// Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11; the generator of Random123 and
// cuRAND's curandStatePhilox4_32_10): NOT in the reference — the counter-based stream of the engine's generator="philox"
// tier, which north_star only requires to agree STATISTICALLY with the reference's pseudo mode.  Restated here so that the
// device stream is held bit-exactly (known answers of Random123's kat_vectors in tests/test_oracle_philox.py).
static inline void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// draw #i of scenario sg: word i & 3 of block (lo32(sg), hi32(sg), i >> 2, 0) keyed by the seed's two words;
// uniform = (word + 1/2) 2^-32 in (0, 1) — the same 32-bit form as the digital-shift Sobol uniforms
static inline double philox_uniform(uint64_t seed, uint64_t sg, uint64_t i) {
    const uint32_t ctr[4] = {(uint32_t)sg, (uint32_t)(sg >> 32), (uint32_t)(i >> 2), 0u}, key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t w[4];
    philox4x32_10(ctr, key, w);
    return ((double)w[i & 3] + 0.5) * (1.0 / 4294967296.0);
}

// src/sim/mod.rs:20-92 with explicit seed and deterministic point<->path map.
static int simulate(const Universe& U, const std::vector<std::pair<std::string, double>>& init,
                    uint64_t N, const SimOptions& opt, double* out /* [N][T][P] */, std::string* err) {
    const int T = (int)U.times.size();
    const int P = (int)U.procs.size();
    const int S = T - 1;
    const int K = (int)U.stochastic_names.size();
    const size_t dims = (size_t)S * K;                     // mod.rs:31-33
    std::vector<double> dts(S), sqrt_dts(S);
    for (int t = 0; t < S; ++t) { dts[t] = U.times[t + 1] - U.times[t]; sqrt_dts[t] = std::sqrt(dts[t]); }  // increment.rs:38-41,75-79
    std::unordered_map<uint64_t, int> time_registry;       // filtration.rs:29 (later duplicates win)
    for (int t = 0; t < T; ++t) { uint64_t b; std::memcpy(&b, &U.times[t], 8); time_registry[b] = t; }
    if (opt.rng_mode >= 1 && opt.rng_mode <= 3 && !opt.sobol_V && dims > 0) { *err = "sobol direction numbers missing"; return 2; }
    if (opt.rng_mode == 4 && !opt.inject) { *err = "inject buffer missing"; return 2; }
    if (opt.scheme != 0 && K == 0) { *err = "runge-kutta needs at least one stochastic factor (reference panics, pseudo.rs:53-58)"; return 2; }
    std::vector<uint64_t> masks;
    if (opt.rng_mode == 2) {
        masks.resize(dims);
        ChaCha8F64 g(opt.seed);
        for (size_t d = 0; d < dims; ++d) masks[d] = g.next_u64();
    }
    int nthreads = opt.nthreads;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
    (void)nthreads;
#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads)
    for (int64_t s = 0; s < (int64_t)N; ++s) {
        uint64_t sg = (uint64_t)s + opt.scenario_offset;   // global scenario index
        Filtration F;                                      // ScenarioFiltration::new, filtration.rs:22-53
        F.u = &U; F.P = P; F.T = T;
        F.raw.assign((size_t)T * P, 0.0);
        F.cache.assign(P, 0.0);
        F.time_registry = &time_registry;
        F.cache_time = U.times[0];
        for (auto& kv : init) {
            auto it = U.process_registry.find(kv.first);
            if (it != U.process_registry.end()) F.set(0, it->second, kv.second);
        }
        F.refresh_cache(U.times[0]);
        std::unique_ptr<Rng> rng;
        bool injected = false;
        if (opt.rng_mode == 0) {
            rng.reset(new PseudoRng(sg + opt.seed, K));    // mod.rs:65 (wrapping add)
        } else if (opt.rng_mode == 4) {
            auto* r = new InjectRng; r->v = opt.inject + (size_t)s * S * (K + 1); r->K1 = K + 1;
            rng.reset(r); injected = true;
        } else {
            auto* r = new TableRng; r->K = K; r->values.resize(dims);
            uint64_t n = sg + 5;                           // sobol.rs:17 skip(5) + single-thread order
            if (opt.rng_mode == 5) {                       // generator = "philox": counter-based, no Sobol point
                for (size_t d = 0; d < dims; ++d) r->values[d] = philox_uniform(opt.seed, sg, d);
            } else if (opt.rng_mode == 1) {                // SobolRng::new + RandomShiftScrambler, sobol.rs:35-78
                ChaCha8F64 g(sg + opt.seed);
                for (size_t d = 0; d < dims; ++d) {
                    double rawv = (double)point_direct(opt.sobol_V + d * 64, n) * (1.0 / 18446744073709551616.0);
                    double sh = g.next_f64();
                    double v = rawv + sh;
                    r->values[d] = v - std::trunc(v);      // f64::fract
                }
            } else if (opt.rng_mode == 2) {
                for (size_t d = 0; d < dims; ++d) r->values[d] = xor_uniform(point_direct(opt.sobol_V + d * 64, n), masks[d]);
            } else {
                for (size_t d = 0; d < dims; ++d) r->values[d] = (double)point_direct(opt.sobol_V + d * 64, n) * (1.0 / 18446744073709551616.0);
            }
            rng.reset(r);
        }
        for (int t = 0; t < S; ++t) {                      // mod.rs:68-84
            if (opt.scheme == 0) euler_iteration(F, U, t, *rng, dts, sqrt_dts, injected);
            else rk_iteration(F, U, t, *rng, dts, sqrt_dts, injected, opt.scheme == 2, K);
        }
        std::memcpy(out + (size_t)s * T * P, F.raw.data(), sizeof(double) * T * P);   // to_lazyframe value column, filtration.rs:112
    }
    return 0;
}

}  // namespace orc

// =======================================================================================
// C interface for ctypes (tests/, bench.py cpu_baseline, smoke())
// =======================================================================================
extern "C" {

void orc_sobol_direction_numbers(const uint32_t* poly, const uint32_t* minit, int minit_stride, int dims, uint64_t* V) {
    orc::direction_numbers(poly, minit, minit_stride, dims, V);
}
void orc_sobol_points_sequential(const uint64_t* V, int dims, uint64_t first, uint64_t count, uint64_t* out) {
    orc::points_sequential(V, dims, first, count, out);
}
void orc_sobol_points_direct(const uint64_t* V, int dims, uint64_t first, uint64_t count, uint64_t* out) {
#pragma omp parallel for
    for (int64_t i = 0; i < (int64_t)count; ++i)
        for (int d = 0; d < dims; ++d) out[(size_t)i * dims + d] = orc::point_direct(V + (size_t)d * 64, first + (uint64_t)i);
}
void orc_chacha_block(const uint32_t* key, uint64_t counter, uint64_t stream, int rounds, uint32_t* out) {
    orc::chacha_block(key, counter, stream, rounds, out);
}
void orc_seed_from_u64(uint64_t seed, uint32_t* key) { orc::seed_from_u64(seed, key); }
void orc_chacha8_u64_stream(uint64_t seed, size_t n, uint64_t* out) {
    orc::ChaCha8F64 g(seed);
    for (size_t i = 0; i < n; ++i) out[i] = g.next_u64();
}
void orc_chacha8_f64_stream(uint64_t seed, size_t n, double* out) {
    orc::ChaCha8F64 g(seed);
    for (size_t i = 0; i < n; ++i) out[i] = g.next_f64();
}
double orc_icdf_normal(double p) { return orc::icdf_normal(p); }
void orc_icdf_normal_array(const double* p, size_t n, double* out) {
    for (size_t i = 0; i < n; ++i) out[i] = orc::icdf_normal(p[i]);
}
uint64_t orc_icdf_poisson(double u, double lambda) { return orc::icdf_poisson(u, lambda); }
void orc_philox4x32_10(const uint32_t* ctr, const uint32_t* key, uint32_t* out) { orc::philox4x32_10(ctr, key, out); }
double orc_philox_uniform(uint64_t seed, uint64_t scenario, uint64_t draw) { return orc::philox_uniform(seed, scenario, draw); }
double orc_xor_uniform(uint64_t x, uint64_t mask) { return orc::xor_uniform(x, mask); }

static thread_local std::string g_err;
const char* orc_last_error(void) { return g_err.c_str(); }

void* orc_universe_parse(const char* const* eqs, int n, const double* times, int T) {
    try {
        std::vector<std::string> v(eqs, eqs + n);
        std::vector<double> tt(times, times + T);
        return orc::parse_equations(v, tt);
    } catch (orc::ParseError& e) { g_err = e.msg; return nullptr; }
}
void orc_universe_free(void* u) { delete (orc::Universe*)u; }
int orc_universe_num_processes(const void* u) { return (int)((const orc::Universe*)u)->procs.size(); }
int orc_universe_num_factors(const void* u) { return (int)((const orc::Universe*)u)->stochastic_names.size(); }
const char* orc_universe_process_name(const void* u, int i) { return ((const orc::Universe*)u)->procs[i].name.c_str(); }
int orc_universe_process_is_levy(const void* u, int i) { return ((const orc::Universe*)u)->procs[i].levy ? 1 : 0; }
int orc_universe_num_terms(const void* u, int i) { return (int)((const orc::Universe*)u)->procs[i].terms.size(); }
const char* orc_universe_factor_name(const void* u, int i) { return ((const orc::Universe*)u)->stochastic_names[i].c_str(); }

// Evaluate one expression against named values (expression-semantics tests).
int orc_expr_eval(const char* src, const char* const* names, const double* values, int n, double t, double* out) {
    try {
        std::unordered_map<std::string, int> reg;
        for (int i = 0; i < n; ++i) reg[names[i]] = i;
        orc::Expr e;
        orc::ExprParser p(src, &reg, &e);
        p.run();
        if (!p.unresolved.empty()) { g_err = "undefined variable '" + p.unresolved[0] + "'"; return 1; }
        *out = orc::eval_node(e, e.root, t, values);
        return 0;
    } catch (orc::ParseError& e) { g_err = e.msg; return 1; }
}

int orc_simulate(const void* u, const char* const* init_names, const double* init_vals, int n_init,
                 uint64_t N, int scheme, int rng_mode, uint64_t seed, uint64_t scenario_offset,
                 const uint64_t* sobol_V, const double* inject, int nthreads, double* out) {
    orc::SimOptions o;
    o.seed = seed; o.scenario_offset = scenario_offset; o.rng_mode = rng_mode; o.scheme = scheme;
    o.sobol_V = sobol_V; o.inject = inject; o.nthreads = nthreads;
    std::vector<std::pair<std::string, double>> init;
    for (int i = 0; i < n_init; ++i) init.emplace_back(init_names[i], init_vals[i]);
    std::string err;
    int rc = orc::simulate(*(const orc::Universe*)u, init, N, o, out, &err);
    if (rc) g_err = err;
    return rc;
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
