// expr.cpp — expression parser (precedence climbing) and CUDA emitter.  See expr.h.
#include "expr.h"

#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace sde {

std::string format_double(double v) {
    if (std::isnan(v)) return "__longlong_as_double(0x7ff8000000000000ll)";
    if (std::isinf(v)) return v > 0 ? "__longlong_as_double(0x7ff0000000000000ll)" : "__longlong_as_double(0xfff0000000000000ll)";
    char buf[64];
    std::snprintf(buf, sizeof buf, "%.17g", v);
    std::string s(buf);
    if (s.find_first_of(".eE") == std::string::npos) s += ".0";
    return s;
}

namespace {
thread_local bool g_real_f32 = false;
}
void set_real_literals_f32(bool f32) { g_real_f32 = f32; }
bool real_literals_f32() { return g_real_f32; }
std::string format_real(double v) {
    if (!g_real_f32) return format_double(v);
    const float f = (float)v;
    if (std::isnan(f)) return "__int_as_float(0x7fc00000)";
    if (std::isinf(f)) return f > 0 ? "__int_as_float(0x7f800000)" : "__int_as_float(0xff800000)";
    char buf[64];
    std::snprintf(buf, sizeof buf, "%.9g", (double)f);
    std::string s(buf);
    if (s.find_first_of(".eE") == std::string::npos) s += ".0";
    return s + "f";
}

namespace {

struct BinOpInfo { const char* tok; Op op; int level; bool word; };
// fasteval gives every binary operator its own precedence (enum order), except that the six
// comparisons share one; '^' is right-associative.
const BinOpInfo kBinOps[] = {
    {"||", Op::Or, 1, false},  {"or", Op::Or, 1, true},   {"&&", Op::And, 2, false}, {"and", Op::And, 2, true},
    {"!=", Op::Ne, 3, false},  {"==", Op::Eq, 3, false},  {">=", Op::Ge, 3, false},  {"<=", Op::Le, 3, false},
    {">", Op::Gt, 3, false},   {"<", Op::Lt, 3, false},   {"+", Op::Add, 4, false},  {"-", Op::Sub, 5, false},
    {"*", Op::Mul, 6, false},  {"/", Op::Div, 7, false},  {"%", Op::Mod, 8, false},  {"^", Op::Pow, 9, false},
};
constexpr int kPowLevel = 9;

struct FuncInfo { const char* name; int min_args; int max_args; };
const FuncInfo kFuncs[] = {
    {"int", 1, 1},  {"ceil", 1, 1},  {"floor", 1, 1}, {"abs", 1, 1},   {"sign", 1, 1},  {"log", 1, 2},
    {"round", 1, 2}, {"min", 1, 1 << 20}, {"max", 1, 1 << 20}, {"e", 0, 0}, {"pi", 0, 0},
    {"sin", 1, 1},  {"cos", 1, 1},   {"tan", 1, 1},   {"asin", 1, 1},  {"acos", 1, 1},  {"atan", 1, 1},
    {"sinh", 1, 1}, {"cosh", 1, 1},  {"tanh", 1, 1},  {"asinh", 1, 1}, {"acosh", 1, 1}, {"atanh", 1, 1},
};

bool ident_char(char c) { return std::isalnum((unsigned char)c) || c == '_'; }

}  // namespace

class ExprParserImpl {
  public:
    ExprParserImpl(const std::string& s, const std::unordered_map<std::string, int>& vars, Expr& out)
        : s_(s), vars_(vars), e_(out) {}

    void run() {
        e_.src_ = s_;
        e_.root_ = climb(1);
        ws();
        if (pos_ != s_.size()) fail("unparsed tokens remaining: '" + s_.substr(pos_) + "'");
    }

  private:
    const std::string& s_;
    const std::unordered_map<std::string, int>& vars_;
    Expr& e_;
    size_t pos_ = 0;

    [[noreturn]] void fail(const std::string& m) const { throw ExprError{"Parse Error: " + m + " in '" + s_ + "'"}; }
    void ws() { while (pos_ < s_.size() && std::isspace((unsigned char)s_[pos_])) ++pos_; }
    int push(ExprNode n) { e_.nodes_.push_back(std::move(n)); return (int)e_.nodes_.size() - 1; }

    const BinOpInfo* peek_binop() {
        ws();
        for (const BinOpInfo& b : kBinOps) {
            size_t n = std::strlen(b.tok);
            if (s_.compare(pos_, n, b.tok) != 0) continue;
            if (b.word && pos_ + n < s_.size() && ident_char(s_[pos_ + n])) continue;
            return &b;
        }
        return nullptr;
    }

    // parse a chain whose operators all have level >= min_level
    int climb(int min_level) {
        int lhs = value();
        for (;;) {
            const BinOpInfo* b = peek_binop();
            if (!b || b->level < min_level) return lhs;
            pos_ += std::strlen(b->tok);
            // left-assoc: the right operand may only contain tighter operators; '^' is right-assoc
            int rhs = climb(b->level == kPowLevel ? kPowLevel : b->level + 1);
            ExprNode n; n.op = b->op; n.args = {lhs, rhs};
            lhs = push(n);
        }
    }

    int value() {
        ws();
        if (pos_ >= s_.size()) fail("unexpected end of expression");
        char c = s_[pos_];
        if (c == '(') {
            ++pos_;
            int v = climb(1);
            ws();
            if (pos_ >= s_.size() || s_[pos_] != ')') fail("missing ')'");
            ++pos_;
            return v;
        }
        if (c == '-') { ++pos_; ExprNode n; n.op = Op::Neg; n.args = {value()}; return push(n); }
        if (c == '+') { ++pos_; return value(); }
        if (c == '!') { ++pos_; ExprNode n; n.op = Op::Not; n.args = {value()}; return push(n); }
        if (std::isdigit((unsigned char)c) || c == '.') return number();
        if (std::isalpha((unsigned char)c) || c == '_') return identifier();
        fail(std::string("unexpected character '") + c + "'");
    }

    int number() {
        size_t st = pos_;
        while (pos_ < s_.size() && (std::isdigit((unsigned char)s_[pos_]) || s_[pos_] == '.')) ++pos_;
        if (pos_ < s_.size() && (s_[pos_] == 'e' || s_[pos_] == 'E')) {
            size_t j = pos_ + 1;
            if (j < s_.size() && (s_[j] == '+' || s_[j] == '-')) ++j;
            if (j < s_.size() && std::isdigit((unsigned char)s_[j])) {
                while (j < s_.size() && std::isdigit((unsigned char)s_[j])) ++j;
                pos_ = j;
            }
        }
        std::string tok = s_.substr(st, pos_ - st);
        char* endp = nullptr;
        double v = std::strtod(tok.c_str(), &endp);
        if (endp == tok.c_str() || *endp) fail("bad number '" + tok + "'");
        if (pos_ < s_.size()) {   // SI-style suffix
            double mul = 0.0; size_t adv = 1;
            switch (s_[pos_]) {
                case 'k': case 'K': mul = 1e3; break;   case 'M': mul = 1e6; break;
                case 'G': mul = 1e9; break;             case 'T': mul = 1e12; break;
                case 'm': mul = 1e-3; break;            case 'u': mul = 1e-6; break;
                case 'n': mul = 1e-9; break;            case 'p': mul = 1e-12; break;
                default: break;
            }
            if (mul == 0.0 && s_.compare(pos_, 2, "\xC2\xB5") == 0) { mul = 1e-6; adv = 2; }
            if (mul != 0.0 && !(pos_ + adv < s_.size() && ident_char(s_[pos_ + adv]))) { v *= mul; pos_ += adv; }
        }
        ExprNode n; n.op = Op::Const; n.value = v;
        return push(n);
    }

    int identifier() {
        size_t st = pos_;
        while (pos_ < s_.size() && ident_char(s_[pos_])) ++pos_;
        std::string name = s_.substr(st, pos_ - st);
        size_t save = pos_;
        ws();
        if (pos_ < s_.size() && s_[pos_] == '(') {
            const FuncInfo* fi = nullptr;
            for (const FuncInfo& f : kFuncs) if (name == f.name) fi = &f;
            if (!fi) fail("unsupported function '" + name + "'");
            ++pos_;
            ExprNode n; n.op = Op::Call; n.fn = name;
            ws();
            if (pos_ < s_.size() && s_[pos_] == ')') ++pos_;
            else for (;;) {
                n.args.push_back(climb(1));
                ws();
                if (pos_ < s_.size() && s_[pos_] == ',') { ++pos_; continue; }
                if (pos_ < s_.size() && s_[pos_] == ')') { ++pos_; break; }
                fail("missing ')' in call to " + name);
            }
            if ((int)n.args.size() < fi->min_args || (int)n.args.size() > fi->max_args)
                fail("wrong number of arguments to " + name);
            return push(n);
        }
        pos_ = save;
        ExprNode n;
        auto it = vars_.find(name);
        if (it != vars_.end()) { n.op = Op::Var; n.var = it->second; }
        else if (name == "t") n.op = Op::Time;
        else throw ExprError{"undefined variable '" + name + "' in '" + s_ + "'"};
        return push(n);
    }
};

Expr Expr::parse(const std::string& src, const std::unordered_map<std::string, int>& vars) {
    Expr e;
    ExprParserImpl p(src, vars, e);
    p.run();
    return e;
}

bool Expr::is_constant() const {
    for (const ExprNode& n : nodes_) if (n.op == Op::Var || n.op == Op::Time) return false;
    return true;
}

bool Expr::uses_var(int idx) const {
    for (const ExprNode& n : nodes_) if (n.op == Op::Var && n.var == idx) return true;
    return false;
}

void Expr::product_form(double& kappa, std::vector<int>& factors) const {
    kappa = 1.0;
    factors.clear();
    product_walk(root_, kappa, factors);
}

void Expr::product_walk(int i, double& kappa, std::vector<int>& factors) const {
    const ExprNode& n = nodes_[i];
    switch (n.op) {
        case Op::Const: kappa *= n.value; return;
        case Op::Neg: kappa = -kappa; product_walk(n.args[0], kappa, factors); return;
        case Op::Mul: product_walk(n.args[0], kappa, factors); product_walk(n.args[1], kappa, factors); return;
        case Op::Div: {
            const ExprNode& d = nodes_[n.args[1]];
            if (d.op == Op::Const && d.value != 0.0 && std::isfinite(1.0 / d.value)) {
                kappa /= d.value;
                product_walk(n.args[0], kappa, factors);
                return;
            }
            break;
        }
        case Op::Call:
            if (n.fn == "e") { kappa *= 2.718281828459045; return; }
            if (n.fn == "pi") { kappa *= 3.141592653589793; return; }
            break;
        default: break;
    }
    factors.push_back(i);
}

std::string Expr::emit_cuda(bool strict, const std::string& c, const std::string& t) const {
    return emit_node(root_, strict, c, t);
}

std::string Expr::emit_node(int i, bool strict, const std::string& c, const std::string& t) const {
    const ExprNode& n = nodes_[i];
    auto A = [&](int k) { return emit_node(n.args[k], strict, c, t); };
    auto bin = [&](const char* strict_fn, const char* op) {
        return strict ? std::string(strict_fn) + "(" + A(0) + ", " + A(1) + ")" : "(" + A(0) + " " + op + " " + A(1) + ")";
    };
    switch (n.op) {
        case Op::Const: { std::string s = format_real(n.value); return n.value < 0 || std::signbit(n.value) ? "(" + s + ")" : s; }
        case Op::Time: return t;
        case Op::Var: return c + "[" + std::to_string(n.var) + "]";
        case Op::Neg: return "(-" + A(0) + ")";
        case Op::Not: return "sde_f_not(" + A(0) + ")";
        case Op::Add: return bin("__dadd_rn", "+");
        case Op::Sub: return bin("__dsub_rn", "-");
        case Op::Mul: return bin("__dmul_rn", "*");
        case Op::Div: return bin("__ddiv_rn", "/");
        case Op::Mod: return "fmod(" + A(0) + ", " + A(1) + ")";
        case Op::Pow: {
            const ExprNode& ex = nodes_[n.args[1]];
            // powf(x, 0.5): equal to sqrt except for -0.0 / -inf; arithmetic=fast takes the MUFU seed + one cubic step (<= 1 ulp)
            if (ex.op == Op::Const && ex.value == 0.5) {
                // max(x, 0)^0.5 (the full-truncation root of square-root diffusions): one fused helper under arithmetic=fast,
                // whose range guards run on the integer pipe instead of four FP64 compares
                const ExprNode& b = nodes_[n.args[0]];
                if (!strict && !real_literals_f32() && b.op == Op::Call && b.fn == "max" && b.args.size() == 2) {
                    const ExprNode& m0 = nodes_[b.args[0]];
                    const ExprNode& m1 = nodes_[b.args[1]];
                    const bool z0 = m0.op == Op::Const && m0.value == 0.0, z1 = m1.op == Op::Const && m1.value == 0.0;
                    if (z0 != z1) return "sde_f_sqrt_max0_fast(" + emit_node(b.args[z0 ? 1 : 0], strict, c, t) + ")";
                }
                return std::string(strict ? "sqrt(" : "sde_f_sqrt_fast(") + A(0) + ")";
            }
            if (ex.op == Op::Const && ex.value == 2.0) return "sde_f_sq(" + A(0) + ")";
            if (ex.op == Op::Const && ex.value == 1.0) return A(0);
            return "pow(" + A(0) + ", " + A(1) + ")";
        }
        case Op::Lt: return "((" + A(0) + " < " + A(1) + ") ? " + format_real(1.0) + " : " + format_real(0.0) + ")";
        case Op::Gt: return "((" + A(0) + " > " + A(1) + ") ? " + format_real(1.0) + " : " + format_real(0.0) + ")";
        case Op::Le: return "((" + A(0) + " <= " + A(1) + ") ? " + format_real(1.0) + " : " + format_real(0.0) + ")";
        case Op::Ge: return "((" + A(0) + " >= " + A(1) + ") ? " + format_real(1.0) + " : " + format_real(0.0) + ")";
        case Op::Eq: return "sde_f_eq(" + A(0) + ", " + A(1) + ")";
        case Op::Ne: return "sde_f_ne(" + A(0) + ", " + A(1) + ")";
        case Op::And: return "sde_f_and(" + A(0) + ", " + A(1) + ")";
        case Op::Or: return "sde_f_or(" + A(0) + ", " + A(1) + ")";
        case Op::Call: {
            const std::string& f = n.fn;
            if (f == "int") return "trunc(" + A(0) + ")";
            if (f == "abs") return "fabs(" + A(0) + ")";
            if (f == "sign") return "sde_f_sign(" + A(0) + ")";
            if (f == "e") return format_real(2.718281828459045);
            if (f == "pi") return format_real(3.141592653589793);
            if (f == "log") {
                if (n.args.size() == 1) return "log10(" + A(0) + ")";
                if (real_literals_f32()) return "(log(" + A(1) + ") / log(" + A(0) + "))";
                return "__ddiv_rn(log(" + A(1) + "), log(" + A(0) + "))";       // log(base, x) = ln x / ln base
            }
            if (f == "round") {
                if (n.args.size() == 1) return "round(" + A(0) + ")";
                if (real_literals_f32()) return "(round(" + A(1) + " / " + A(0) + ") * " + A(0) + ")";
                return "__dmul_rn(round(__ddiv_rn(" + A(1) + ", " + A(0) + ")), " + A(0) + ")";
            }
            if (f == "min" || f == "max") {
                std::string acc = A(0);
                for (size_t k = 1; k < n.args.size(); ++k) acc = "sde_f_" + f + "(" + acc + ", " + A((int)k) + ")";
                if (n.args.size() == 1) acc = "sde_f_" + f + "(" + acc + ", " + acc + ")";
                return acc;
            }
            return f + "(" + A(0) + ")";   // ceil floor sin cos tan asin acos atan sinh cosh tanh asinh acosh atanh
        }
    }
    return "0.0";
}

}  // namespace sde
