// lower.cpp — model IR -> CUDA C++ (see lower.h).
#include "lower.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <numeric>
#include <map>
#include <sstream>

namespace sde {
namespace {

// Where the reference's ScenarioFiltrationCache points while a step is being generated.
// The cache is refreshed from the raw rows only when an expression is evaluated at a time
// different from cache.time (src/func.rs:37-39, src/filtration.rs:70-79); writes to the
// rows never touch it (filtration.rs:61-64).  Times are strictly increasing, so within the
// step for index t there are exactly three possibilities.
enum CacheAt { OLD, CUR, NEXT };

class StepGen {
  public:
    StepGen(const Universe& u, const LowerOptions& opt) : u_(u), opt_(opt) {}

    // path-independent per-step constants hoisted into the tile prologue (arithmetic=fast only):
    // each entry is a CUDA expression over t_cur, t_next, dt, sqrt_dt
    std::vector<std::string> slots;
    // the same constants evaluated on the host, [slot][step] (IEEE fma / multiply / sqrt: bit-identical to the device),
    // so that constants which do not change along the time grid can be emitted as literals instead of table reads
    std::vector<std::vector<double>> slot_values;
    // spread along the time grid (max - min) below which a slot counts as constant under arithmetic=fast: what re-rounding
    // the time points by a few ulp would do to it.  A uniform grid k/D has dt values that differ by ~ulp(t_end) only because
    // k/D is not representable; the factor 4 covers both end points of a step and the subtraction.
    std::vector<double> slot_tol;
    std::vector<bool> slot_mult;                             // the slot is a multiplier of an FMA (kept on the uniform datapath when constant)
    std::vector<double> model_consts;                        // literal coefficients placed in __constant__ memory (sde_mc[])
    bool matrix = false;                                     // the step was emitted in matrix form (wide linear model)
    std::vector<double> mat_drift, mat_M;                    // matrix form: a_i [NB*8] and loadings M [NB][K][8]
    std::vector<int> mat_kend;                               // matrix form: factors used by each block of 8 processes
    std::string matrix_decl() const { return matrix_decl_.str(); }
    std::string prelude() const { return pre_.str(); }      // declarations emitted before the step body

    // Emits the body of sde_model_step given the cache position on entry; returns it on exit.
    CacheAt generate(CacheAt enter, std::ostringstream& o) {
        state_ = enter;
        o_ = &o;
        slots.clear();
        slot_values.clear();
        slot_mult.clear();
        slot_tol.clear();
        factor_tmp_.clear();
        factor_tmp_id_ = 0;
        model_consts.clear();
        pre_.str("");
        w_declared_.assign(u_.K(), false);
        {   // hoist per-step constants into the tile prologue only while the step record stays small
            size_t n = 0;
            double lin[128];
            if (!opt_.strict && opt_.scheme == SCHEME_EULER)
                for (int p : u_.levy_indices) {
                    const Process& pr = u_.processes[p];
                    if (!linear_in_own_state(pr, p, lin)) continue;
                    std::vector<bool> used(u_.K(), false);
                    for (const Term& t : pr.terms) if (t.kind == IncKind::Wiener) used[t.factor] = true;
                    n += 1;
                    for (bool b : used) n += b ? 1 : 0;
                }
            // the factored form of the other processes (arithmetic=fast): one constant per (group, increment kind)
            if (!opt_.strict)
                for (int p : u_.levy_indices) {
                    const Process& pr = u_.processes[p];
                    if (opt_.scheme == SCHEME_EULER && linear_in_own_state(pr, p, lin)) continue;
                    n += factorise(pr).slots_needed(opt_.scheme == SCHEME_RK);
                }
            hoist_ = n <= 16;
        }
        const int P = u_.P();
        matrix_decl_.str("");
        matrix = matrix_form_ok();
        if (matrix) { emit_matrix_form(); return state_; }
        for (int p = 0; p < P; ++p) line("sde_real n" + std::to_string(p) + " = " + format_real(0.0) + ";");   // row t+1 is zero until set (filtration.rs:28)
        if (opt_.scheme == SCHEME_EULER) euler(); else if (!opt_.strict) runge_kutta_fast(); else runge_kutta();
        for (int p = 0; p < P; ++p) line("row[" + std::to_string(p) + "] = n" + std::to_string(p) + ";");
        return state_;
    }

  private:
    const Universe& u_;
    const LowerOptions& opt_;
    CacheAt state_ = OLD;
    std::ostringstream* o_ = nullptr;
    // |d slot| for a re-rounding of the time points: slot = kappa * dt or kappa * sqrt(dt)
    double grid_tol(double kappa, bool sqrt_of_dt) const {
        double tmax = 0.0, dtmin = INFINITY;
        for (double t : u_.times) tmax = std::max(tmax, std::fabs(t));
        for (size_t q = 0; q + 1 < u_.times.size(); ++q) dtmin = std::min(dtmin, u_.times[q + 1] - u_.times[q]);
        const double d = 4.0 * (std::nextafter(tmax, INFINITY) - tmax);
        if (!(dtmin > 0.0) || !std::isfinite(d)) return 0.0;
        return std::fabs(kappa) * (sqrt_of_dt ? d / (2.0 * std::sqrt(dtmin)) : d);
    }
    std::ostringstream pre_;
    std::ostringstream matrix_decl_;
    std::vector<bool> w_declared_;
    bool hoist_ = true;

    void line(const std::string& s) { *o_ << "    " << s << "\n"; }
    std::string mul(const std::string& a, const std::string& b) const { return opt_.strict ? "__dmul_rn(" + a + ", " + b + ")" : "(" + a + " * " + b + ")"; }
    std::string add(const std::string& a, const std::string& b) const { return opt_.strict ? "__dadd_rn(" + a + ", " + b + ")" : "(" + a + " + " + b + ")"; }

    // ScenarioFiltration::refresh_cache (filtration.rs:70-79): reload every registered name from the row of `at`.
    void refresh(CacheAt at) {
        line(at == CUR ? "// cache refresh from row t (func.rs:37-39)" : "// cache refresh from row t+1 as written so far (func.rs:37-39)");
        std::vector<int> idx;
        for (auto& kv : u_.process_registry) idx.push_back(kv.second);
        std::sort(idx.begin(), idx.end());
        for (int i : idx) {
            std::string s = std::to_string(i);
            line("c[" + s + "] = " + (at == CUR ? "row[" + s + "]" : "n" + s) + ";");
        }
        line(std::string("ct = ") + (at == CUR ? "t_cur;" : "t_next;"));
        state_ = at;
        factor_tmp_.clear();
    }
    // Function::eval (func.rs:32-42)
    std::string eval(const Expr& e, CacheAt at) {
        if (state_ != at) refresh(at);
        return e.emit_cuda(opt_.strict);
    }
    // Incrementor::sample (increment.rs:46-59, 89-97, 137-148)
    std::string increment(const Term& t) {
        if (t.kind == IncKind::Time) return "dt";
        std::string z = "zu[" + std::to_string(t.factor) + "]";
        if (t.kind == IncKind::Wiener) return mul("sqrt_dt", z);
        std::string lam = eval(t.lambda, CUR);               // lambda.eval(ts[time_idx], ..) * dt
        return "sde_icdf_poisson(" + z + ", " + mul(lam, "dt") + ")";
    }

    // coefficient j == lin[j] * X_p for every term (literal `c * X` or `X * c`), at most 128 terms
    bool linear_in_own_state(const Process& pr, int p, double* lin) const {
        if (pr.terms.empty() || pr.terms.size() > 128) return false;
        for (size_t j = 0; j < pr.terms.size(); ++j) {
            const Expr& e = pr.terms[j].coeff;
            const ExprNode& r = e.nodes()[e.root()];
            if (r.op != Op::Mul) return false;
            const ExprNode& a = e.nodes()[r.args[0]];
            const ExprNode& b = e.nodes()[r.args[1]];
            if (a.op == Op::Const && b.op == Op::Var && b.var == p) lin[j] = a.value;
            else if (b.op == Op::Const && a.op == Op::Var && a.var == p) lin[j] = b.value;
            else return false;
        }
        // with a stale cache (steady state entered AT times[t]) c[p] is not X_p(t): keep the literal form
        return true;
    }

    // Matrix form of a wide linear model (arithmetic=fast, Euler): every process is Levy with coefficients a_j X_p on
    // dt / dW_k only (a Cholesky-loaded GBM basket).  One step is X_i *= A_i + sum_k M[i][k] w_k with w_k = sqrt(dt) z_k.
    // Written out term by term that is thousands of FMAs with distinct literal loadings: a 150 KB step that streams through
    // the instruction cache and pins every draw in a register (255 registers, 2 warps per scheduler).  Here it is emitted
    // as loops over blocks of 8 processes with the loadings in constant memory ([block][k][8]: uniform 128-bit loads),
    // the draws and the rows in (L1-resident) local arrays: a few KB of code at ~80 registers.  Same FMA order per
    // process as the term-by-term form (k ascending onto A_i), so the results are bit-identical to it.
    bool matrix_form_ok() const {
        if (opt_.strict || opt_.scheme != SCHEME_EULER || hoist_) return false;
        if (!u_.algebraic_indices.empty() || u_.P() < 16 || u_.K() < 16) return false;
        double lin[128];
        for (int p = 0; p < u_.P(); ++p) {
            const Process& pr = u_.processes[p];
            if (!pr.levy || !linear_in_own_state(pr, p, lin)) return false;
            for (const Term& t : pr.terms) if (t.kind == IncKind::Poisson) return false;
        }
        return true;
    }
    void emit_matrix_form() {
        const int P = u_.P(), K = u_.K(), NB = (P + 7) / 8;
        std::vector<double> drift(NB * 8, 0.0), M((size_t)NB * K * 8, 0.0);
        std::vector<int> kend(NB, 0);
        double lin[128];
        for (int p = 0; p < P; ++p) {
            const Process& pr = u_.processes[p];
            linear_in_own_state(pr, p, lin);
            for (size_t j = 0; j < pr.terms.size(); ++j) {
                const Term& t = pr.terms[j];
                if (t.kind == IncKind::Time) drift[p] += lin[j];
                else { M[((size_t)(p / 8) * K + t.factor) * 8 + (p % 8)] += lin[j]; kend[p / 8] = std::max(kend[p / 8], t.factor + 1); }
            }
        }
        mat_drift = drift; mat_M = M; mat_kend = kend;
        matrix_decl_ << "__constant__ sde_real sde_ma[" << NB * 8 << "] = {";
        for (size_t i = 0; i < drift.size(); ++i) matrix_decl_ << (i ? ", " : "") << format_real(drift[i]);
        matrix_decl_ << "};\n__constant__ sde_real sde_mm[" << M.size() << "] = {";
        for (size_t i = 0; i < M.size(); ++i) matrix_decl_ << (i ? ", " : "") << format_real(M[i]);
        matrix_decl_ << "};\n";
        line("// matrix form (see lower.cpp): X_i *= A_i + sum_k M[i][k] sqrt(dt) z_k, blocks of 8 processes");
        line("sde_real w[SDE_KK];");
        line("#pragma unroll 4");
        line("for (int k = 0; k < SDE_K; ++k) w[k] = sqrt_dt * zu[k];");
        // the block loop is written out (NB copies of a small rolled k-loop with literal bounds): every constant-memory
        // address is then a loop counter plus a literal, which ptxas keeps on the uniform datapath (LDCU.128: two
        // loadings per instruction) — indexed through a runtime block number it falls back to per-thread LDC.64
        for (int b = 0; b < NB; ++b) {
            const std::string sb = std::to_string(b * 8), off = std::to_string((size_t)b * K * 8);
            line("{   // processes " + sb + " .. " + std::to_string(std::min(P, b * 8 + 8) - 1));
            line("    sde_real g[8];");
            line("#pragma unroll");
            line("    for (int j = 0; j < 8; ++j) g[j] = fma(sde_ma[" + sb + " + j], dt, " + format_real(1.0) + ");");
            line("#pragma unroll 4");
            line("    for (int k = 0; k < " + std::to_string(kend[b]) + "; ++k) {");
            line("        const sde_real wk = w[k];");
            line("#pragma unroll");
            line("        for (int j = 0; j < 8; ++j) g[j] = fma(sde_mm[" + off + " + k * 8 + j], wk, g[j]);");
            line("    }");
            for (int j = 0; j < 8 && b * 8 + j < P; ++j)
                line("    row[" + std::to_string(b * 8 + j) + "] = row[" + std::to_string(b * 8 + j) + "] * g[" + std::to_string(j) + "];");
            line("}");
        }
        line("ct = t_cur;");
        state_ = CUR;                                        // as after the refresh of the term-by-term form
    }

    // ---- arithmetic = fast: coefficients in product form, terms grouped by their state-dependent part ----------------------
    // A Levy process is  dX = sum_j C_j(X, t) dx_j.  Each C_j is flattened to kappa_j * prod(factors) (Expr::product_form).
    // Terms with the same factor product V_g form a group:  sum_j C_j dx_j = sum_g V_g * W_g,  W_g = sum_{j in g} kappa_j dx_j.
    // W_g does not depend on the state, so both Runge-Kutta stages share it, and kappa_j dt / kappa_j sqrt(dt) are per-step
    // constants (slots).  Factors common to all groups of a process are taken out of the sum:  F * sum_g (V_g / F) W_g.
    // Heston's dS = 0.05 S dt + sqrt(v+) S dW becomes S * fma(sqrt(v+), b z, a): 2 FP64 instructions per stage instead of 5.
    // Every regrouping moves a result by <= 1 ulp of its largest term (the mode's stated <= ~3 ulp per step).
    struct Group {
        std::vector<std::string> f, rest;                    // factor strings (sorted); f minus the process's common factors
        std::vector<size_t> terms;
        bool has_dt = false, has_wiener = false;
        int n_wiener_factors = 0;
    };
    struct Factored {
        std::vector<Group> groups;
        std::vector<std::string> common;
        std::vector<double> kappa;                           // per term
        size_t slots_needed(bool rk) const {
            size_t n = 0;
            for (const Group& g : groups) n += (g.has_dt ? 1 : 0) + (size_t)g.n_wiener_factors + ((rk && g.has_wiener) ? 1 : 0);
            return n;
        }
    };
    Factored factorise(const Process& pr) const {
        Factored F;
        F.kappa.resize(pr.terms.size());
        for (size_t j = 0; j < pr.terms.size(); ++j) {
            std::vector<int> nodes;
            pr.terms[j].coeff.product_form(F.kappa[j], nodes);
            std::vector<std::string> f;
            for (int nd : nodes) f.push_back(pr.terms[j].coeff.emit_cuda_node(nd, false));
            std::sort(f.begin(), f.end());
            Group* g = nullptr;
            for (Group& q : F.groups) if (q.f == f) g = &q;
            if (!g) { F.groups.emplace_back(); g = &F.groups.back(); g->f = f; }
            g->terms.push_back(j);
        }
        for (Group& g : F.groups) {
            std::vector<bool> used(u_.K(), false);
            for (size_t j : g.terms) {
                const Term& t = pr.terms[j];
                if (t.kind == IncKind::Time) g.has_dt = true;
                if (t.kind == IncKind::Wiener) { g.has_wiener = true; used[t.factor] = true; }
            }
            for (bool b : used) g.n_wiener_factors += b ? 1 : 0;
        }
        if (!F.groups.empty()) {                              // multiset intersection of the groups' factors
            F.common = F.groups[0].f;
            for (size_t q = 1; q < F.groups.size(); ++q) {
                std::vector<std::string> keep, pool = F.groups[q].f;
                for (const std::string& x : F.common) {
                    auto it = std::find(pool.begin(), pool.end(), x);
                    if (it != pool.end()) { keep.push_back(x); pool.erase(it); }
                }
                F.common = keep;
            }
            for (Group& g : F.groups) {
                g.rest = g.f;
                for (const std::string& x : F.common) g.rest.erase(std::find(g.rest.begin(), g.rest.end(), x));
            }
        }
        return F;
    }
    // A factor that is more than a plain state read is evaluated once per cache state and named (the same root or
    // payoff-like sub-expression usually sits in several terms and processes; ptxas does not merge the copies' selects).
    std::map<std::string, std::string> factor_tmp_;
    int factor_tmp_id_ = 0;
    std::string fx(const std::string& f) {
        if (f == "ct" || (f.size() >= 4 && f.compare(0, 2, "c[") == 0 && f.find(']') == f.size() - 1)) return f;
        auto it = factor_tmp_.find(f);
        if (it != factor_tmp_.end()) return it->second;
        const std::string name = "q" + std::to_string(factor_tmp_id_++);
        line("const sde_real " + name + " = " + f + ";");
        factor_tmp_[f] = name;
        return name;
    }
    std::string product(const std::vector<std::string>& f) {
        std::string acc;
        for (const std::string& x : f) { const std::string v = fx(x); acc = acc.empty() ? v : "(" + acc + " * " + v + ")"; }
        return acc;
    }
    // a per-step constant kappa * dt or kappa * sqrt(dt): table slot / literal when hoisted, else computed in place
    // `uniform`: the constant is the only non-register operand of a three-operand FMA (kept on the uniform datapath, sde_uc);
    // constants that feed a multiply, a sign flip or an FMA that has another constant stay literals — uniform registers are few
    std::string step_const(double kappa, bool sqrt_of_dt, bool uniform) {
        const std::string expr = "(" + format_real(kappa) + (sqrt_of_dt ? " * sqrt_dt)" : " * dt)");
        if (!hoist_) return expr;
        const int S = u_.T() - 1;
        std::vector<double> v(S);
        for (int q = 0; q < S; ++q) { const double dt = u_.times[q + 1] - u_.times[q]; v[q] = kappa * (sqrt_of_dt ? std::sqrt(dt) : dt); }
        slots.push_back(expr);
        slot_values.push_back(v);
        slot_mult.push_back(uniform);
        slot_tol.push_back(grid_tol(kappa, sqrt_of_dt));
        return "SDE_SLOT_" + std::to_string(slots.size() - 1);
    }
    // W_g of every group (declared as w<p>_<g>), and for Runge-Kutta the probe weights pw<p>_<g> = (sum kappa_j) sk sqrt(dt)
    // of the groups that hold Wiener terms.  `inc_name(j)` names the sampled increment of a Poisson term.
    template <class IncName>
    void emit_weights(int p, const Process& pr, const Factored& F, bool rk, IncName inc_name) {
        for (size_t gi = 0; gi < F.groups.size(); ++gi) {
            const Group& g = F.groups[gi];
            double kdt = 0.0, ksum = 0.0;
            std::vector<double> kw(u_.K(), 0.0);
            std::vector<bool> used(u_.K(), false);
            for (size_t j : g.terms) {
                const Term& t = pr.terms[j];
                if (t.kind == IncKind::Time) kdt += F.kappa[j];
                if (t.kind == IncKind::Wiener) { kw[t.factor] += F.kappa[j]; used[t.factor] = true; ksum += F.kappa[j]; }
            }
            bool alone = true;                               // W_g is the bare dt constant: an FMA operand of the stage sums
            for (size_t j : g.terms) if (pr.terms[j].kind != IncKind::Time) alone = false;
            std::string acc = g.has_dt ? step_const(kdt, false, alone) : "";
            for (int k = 0; k < u_.K(); ++k) {
                if (!used[k]) continue;
                const std::string b = step_const(kw[k], true, !acc.empty()), z = "zu[" + std::to_string(k) + "]";
                acc = acc.empty() ? "(" + b + " * " + z + ")" : "fma(" + b + ", " + z + ", " + acc + ")";
            }
            for (size_t j : g.terms) {
                if (pr.terms[j].kind != IncKind::Poisson) continue;
                const std::string kq = format_real(F.kappa[j]), x = inc_name(j);
                acc = acc.empty() ? "(" + kq + " * " + x + ")" : "fma(" + kq + ", " + x + ", " + acc + ")";
            }
            const std::string id = std::to_string(p) + "_" + std::to_string(gi);
            line("const sde_real w" + id + " = " + acc + ";");
            if (rk && g.has_wiener) line("const sde_real pw" + id + " = sde_f_xorsign((sde_real)" + step_const(ksum, true, false) + ", skm);");
        }
    }
    // sum_g (V_g / F) * weight_g over the groups that have the weight (fma chain onto the factor-free group)
    std::string inner_sum(int p, const Factored& F, const char* weight, bool wiener_only) {
        std::string acc;
        for (int pass = 0; pass < 2; ++pass)
            for (size_t gi = 0; gi < F.groups.size(); ++gi) {
                const Group& g = F.groups[gi];
                if ((wiener_only && !g.has_wiener) || (pass == 0) != g.rest.empty()) continue;
                const std::string w = weight + std::to_string(p) + "_" + std::to_string(gi);
                if (g.rest.empty()) acc = acc.empty() ? w : "(" + acc + " + " + w + ")";
                else { const std::string v = product(g.rest); acc = acc.empty() ? "(" + v + " * " + w + ")" : "fma(" + v + ", " + w + ", " + acc + ")"; }
            }
        return acc;
    }
    void ensure(CacheAt at) { if (state_ != at) refresh(at); }

    void runge_kutta_fast() {                                // runge_kutta.rs:5-107 in the factored form above
        const int P = u_.P();
        // sk = +1 where u0 > 1/2 else -1 (runge_kutta.rs:18-22), kept as a sign mask
        if (opt_.u0_bits) line("const unsigned int skm = (~u0) & 0x80000000u;   // runge_kutta.rs:18-22: sk = -1 unless u0 > 1/2 (top bit of the word)");
        else line("const unsigned int skm = (u0 > " + format_real(0.5) + ") ? 0u : 0x80000000u;   // runge_kutta.rs:18-22: sk = -1 unless u0 > 1/2");
        if (opt_.rk_textbook) state_ = OLD;
        std::vector<Factored> F(P);
        for (int p = 0; p < P; ++p) {                        // :26-35 pre-sample (Poisson intensities are evaluated here)
            const Process& pr = u_.processes[p];
            if (!pr.levy) continue;
            F[p] = factorise(pr);
            for (size_t j = 0; j < pr.terms.size(); ++j) {
                if (pr.terms[j].kind != IncKind::Poisson) continue;
                std::string x = increment(pr.terms[j]);
                line("const sde_real inc_" + std::to_string(p) + "_" + std::to_string(j) + " = " + x + ";");
            }
            emit_weights(p, pr, F[p], true, [&](size_t j) { return "inc_" + std::to_string(p) + "_" + std::to_string(j); });
        }
        for (int p = 0; p < P; ++p) line("const sde_real x" + std::to_string(p) + " = row[" + std::to_string(p) + "];");   // :38-42
        auto stage = [&](int p, const char* tag, CacheAt at) {   // k = F * i  (declares f<tag>_p, i<tag>_p, k<tag>_p)
            const Process& pr = u_.processes[p];
            const std::string sp = std::to_string(p), t = tag;
            if (!pr.terms.empty()) ensure(at);
            const std::string in = inner_sum(p, F[p], "w", false);
            line("const sde_real i" + t + "_" + sp + " = " + (in.empty() ? format_real(0.0) : in) + ";");
            if (!F[p].common.empty()) {
                const std::string f = product(F[p].common);
                line("const sde_real f" + t + "_" + sp + " = " + f + ";");
                line("const sde_real k" + t + "_" + sp + " = (f" + t + "_" + sp + " * i" + t + "_" + sp + ");");
            } else {
                line("const sde_real k" + t + "_" + sp + " = i" + t + "_" + sp + ";");
            }
        };
        for (int p = 0; p < P; ++p) if (u_.processes[p].levy) stage(p, "1", CUR);      // :45-55  k1
        for (int p = 0; p < P; ++p) {                        // :62-78  probe row: x + k1 + sum_j C_j sk sqrt(dt)
            const Process& pr = u_.processes[p];
            if (!pr.levy) continue;
            const std::string sp = std::to_string(p);
            const std::string pert = inner_sum(p, F[p], "pw", true);
            if (pert.empty()) line("n" + sp + " = (x" + sp + " + k1_" + sp + ");");
            else if (!F[p].common.empty()) line("n" + sp + " = fma(f1_" + sp + ", (i1_" + sp + " + " + pert + "), x" + sp + ");");
            else line("n" + sp + " = ((x" + sp + " + k1_" + sp + ") + " + pert + ");");
        }
        for (int p = 0; p < P; ++p) if (u_.processes[p].levy) stage(p, "2", NEXT);     // :81-91  k2 at the probe row
        for (int p : u_.levy_indices) {                      // :94-97
            std::string sp = std::to_string(p);
            line("n" + sp + " = fma(" + format_real(0.5) + ", (k1_" + sp + " + k2_" + sp + "), x" + sp + ");");
        }
        if (opt_.rk_textbook && !u_.algebraic_indices.empty()) state_ = OLD;
        for (int a : u_.algebraic_indices) {                 // :101-106 (sees the probe row: cache not refreshed)
            std::string ex = eval(u_.processes[a].algebraic, NEXT);
            line("n" + std::to_string(a) + " = " + ex + ";   // algebraic '" + u_.processes[a].name + "'");
        }
    }

    void euler() {                                           // src/sim/euler.rs:5-37
        for (int p : u_.levy_indices) {
            const Process& pr = u_.processes[p];
            std::string sp = std::to_string(p);
            line("{   // Levy process " + sp + " '" + pr.name + "' (euler.rs:15-28)");
            double lin[128];
            if (!opt_.strict && linear_in_own_state(pr, p, lin)) {
                // arithmetic=fast only: every coefficient is a_j * X_p, so
                //   X_p + sum_j (a_j X_p) dx_j = X_p * (A + sum_k B_k z_k + sum_poisson a_j N_j),
                //   A = 1 + sum_{dt terms} a_j dt,  B_k = (sum_{terms on dW_k} a_j) sqrt(dt)
                // A and B_k do not depend on the path: they are computed once per step in the tile prologue
                // (sde_model_step_consts) and read from shared memory.  <= ~3 ulp per step vs the literal order.
                if (state_ != CUR && !pr.terms.empty()) refresh(CUR);
                std::string A = format_real(1.0);
                const int S = u_.T() - 1;
                std::vector<double> Av(S, 1.0), dtv(S), sqv(S);
                for (int t = 0; t < S; ++t) { dtv[t] = u_.times[t + 1] - u_.times[t]; sqv[t] = std::sqrt(dtv[t]); }
                std::vector<double> bsum(u_.K(), 0.0);
                std::vector<bool> bused(u_.K(), false);
                for (size_t j = 0; j < pr.terms.size(); ++j) {
                    const Term& t = pr.terms[j];
                    if (t.kind == IncKind::Time) {
                        A = "fma(" + format_real(lin[j]) + ", dt, " + A + ")";
                        for (int q = 0; q < S; ++q) Av[q] = std::fma(lin[j], dtv[q], Av[q]);
                    } else if (t.kind == IncKind::Wiener) { bsum[t.factor] += lin[j]; bused[t.factor] = true; }
                }
                if (hoist_) {
                    slots.push_back(A);
                    slot_values.push_back(Av);
                    slot_mult.push_back(false);              // the addend the factor terms accumulate onto
                    {
                        double asum = 0.0;
                        for (size_t j = 0; j < pr.terms.size(); ++j) if (pr.terms[j].kind == IncKind::Time) asum += std::fabs(lin[j]);
                        slot_tol.push_back(grid_tol(asum, false));
                    }
                    line("sde_real g = SDE_SLOT_" + std::to_string(slots.size() - 1) + ";");
                    for (int k = 0; k < u_.K(); ++k) {
                        if (!bused[k]) continue;
                        slots.push_back("(" + format_real(bsum[k]) + " * sqrt_dt)");
                        std::vector<double> Bv(S);
                        for (int q = 0; q < S; ++q) Bv[q] = bsum[k] * sqv[q];
                        slot_values.push_back(Bv);
                        slot_mult.push_back(true);
                        slot_tol.push_back(grid_tol(bsum[k], true));
                        line("g = fma(SDE_SLOT_" + std::to_string(slots.size() - 1) + ", zu[" + std::to_string(k) + "], g);");
                    }
                } else {
                    // many (process, factor) pairs (e.g. a Cholesky-loaded basket): loadings stay immediates and the
                    // scaled draws w_k = sqrt(dt) z_k are shared by all processes
                    line("sde_real g = " + A + ";");
                    for (int k = 0; k < u_.K(); ++k) {
                        if (!bused[k]) continue;
                        if (!w_declared_[k]) { pre_ << "    const sde_real w" << k << " = sqrt_dt * zu[" << k << "];\n"; w_declared_[k] = true; }
                        // loadings live in constant memory: DFMA reads c[bank][offset] operands directly, whereas a 64-bit
                        // literal costs two uniform moves per use (measured on the 64-asset basket: 8 339 UMOV for 6 500 DFMA)
                        model_consts.push_back(bsum[k]);
                        line("g = fma(sde_mc[" + std::to_string(model_consts.size() - 1) + "], w" + std::to_string(k) + ", g);");
                    }
                }
                for (size_t j = 0; j < pr.terms.size(); ++j) {
                    if (pr.terms[j].kind != IncKind::Poisson) continue;
                    std::string x = increment(pr.terms[j]);
                    line("g = fma(" + format_real(lin[j]) + ", " + x + ", g);");
                }
                line("n" + sp + " = row[" + sp + "] * g; }");   // Levy slots of the cache always equal row t in Euler
                continue;
            }
            if (!opt_.strict) {                              // factored form (see factorise)
                if (!pr.terms.empty()) ensure(CUR);
                factor_tmp_.clear();                         // temporaries live in this process's block
                const Factored F = factorise(pr);
                std::vector<std::string> incs(pr.terms.size());
                for (size_t j = 0; j < pr.terms.size(); ++j) if (pr.terms[j].kind == IncKind::Poisson) incs[j] = increment(pr.terms[j]);
                emit_weights(p, pr, F, false, [&](size_t j) { return incs[j]; });
                const std::string in = inner_sum(p, F, "w", false);
                if (in.empty()) line("n" + sp + " = row[" + sp + "]; }");
                else if (!F.common.empty()) { const std::string f = product(F.common); line("n" + sp + " = fma(" + f + ", " + in + ", row[" + sp + "]); }"); }
                else line("n" + sp + " = (row[" + sp + "] + " + in + "); }");
                factor_tmp_.clear();
                continue;
            }
            line("sde_real val = row[" + sp + "];");
            for (const Term& t : pr.terms) {
                std::string cf = eval(t.coeff, CUR);
                line("{ const sde_real cf = " + cf + ";");
                std::string x = increment(t);
                line("  const sde_real x = " + x + ";");
                line("  val = " + add("val", mul("cf", "x")) + "; }");
            }
            line("n" + sp + " = val; }");
        }
        for (int a : u_.algebraic_indices) {                 // euler.rs:31-36
            std::string ex = eval(u_.processes[a].algebraic, NEXT);
            line("n" + std::to_string(a) + " = " + ex + ";   // algebraic '" + u_.processes[a].name + "'");
        }
    }

    void runge_kutta() {                                     // src/sim/runge_kutta.rs:5-107
        const int P = u_.P();
        line("const sde_real sk = (u0 > " + format_real(0.5) + ") ? " + format_real(1.0) + " : " + format_real(-1.0) + ";   // runge_kutta.rs:18-22");
        // arithmetic=fast: the probe perturbation c * sk * sqrt(dt) shares one product sk * sqrt(dt) per step (<= 1 ulp per term)
        if (!opt_.strict) line("const sde_real sksq = sk * sqrt_dt;");
        if (opt_.rk_textbook) state_ = OLD;                  // textbook variant: k1 at the settled row
        for (int p = 0; p < P; ++p) {                        // :26-35 pre-sample, reused by k1 and k2
            const Process& pr = u_.processes[p];
            if (!pr.levy) continue;
            for (size_t j = 0; j < pr.terms.size(); ++j) {
                std::string x = increment(pr.terms[j]);
                line("const sde_real inc_" + std::to_string(p) + "_" + std::to_string(j) + " = " + x + ";");
            }
        }
        for (int p = 0; p < P; ++p) line("const sde_real x" + std::to_string(p) + " = row[" + std::to_string(p) + "];");   // :38-42
        for (int p = 0; p < P; ++p) {                        // :45-55  k1
            const Process& pr = u_.processes[p];
            if (!pr.levy) continue;
            std::string k = "k1_" + std::to_string(p);
            line("sde_real " + k + " = " + format_real(0.0) + ";");
            for (size_t j = 0; j < pr.terms.size(); ++j) {
                std::string cf = eval(pr.terms[j].coeff, CUR);
                line(k + " = " + add(k, mul(cf, "inc_" + std::to_string(p) + "_" + std::to_string(j))) + ";");
            }
        }
        for (int p = 0; p < P; ++p) {                        // :62-78  probe row
            const Process& pr = u_.processes[p];
            if (!pr.levy) continue;
            std::string sp = std::to_string(p);
            line("{ sde_real pert = " + format_real(0.0) + ";");
            for (const Term& t : pr.terms) {
                if (t.kind != IncKind::Wiener) continue;
                std::string cf = eval(t.coeff, CUR);
                line("  pert = " + add("pert", opt_.strict ? mul(mul(cf, "sk"), "sqrt_dt") : mul(cf, "sksq")) + ";");
            }
            line("  n" + sp + " = " + add(add("x" + sp, "k1_" + sp), "pert") + "; }");
        }
        for (int p = 0; p < P; ++p) {                        // :81-91  k2 at the probe row
            const Process& pr = u_.processes[p];
            if (!pr.levy) continue;
            std::string k = "k2_" + std::to_string(p);
            line("sde_real " + k + " = " + format_real(0.0) + ";");
            for (size_t j = 0; j < pr.terms.size(); ++j) {
                std::string cf = eval(pr.terms[j].coeff, NEXT);
                line(k + " = " + add(k, mul(cf, "inc_" + std::to_string(p) + "_" + std::to_string(j))) + ";");
            }
        }
        for (int p : u_.levy_indices) {                      // :94-97
            std::string sp = std::to_string(p);
            line("n" + sp + " = " + add("x" + sp, mul(format_real(0.5), add("k1_" + sp, "k2_" + sp))) + ";");
        }
        if (opt_.rk_textbook && !u_.algebraic_indices.empty()) state_ = OLD;
        for (int a : u_.algebraic_indices) {                 // :101-106 (sees the probe row: cache not refreshed)
            std::string ex = eval(u_.processes[a].algebraic, NEXT);
            line("n" + std::to_string(a) + " = " + ex + ";   // algebraic '" + u_.processes[a].name + "'");
        }
    }
};

int gcd_int(int a, int b) { return b ? gcd_int(b, a % b) : a; }

}  // namespace

namespace {
struct RealLiteralMode {                                     // format_real() follows the plan's dtype while it is lowered
    explicit RealLiteralMode(bool f32) { set_real_literals_f32(f32); }
    ~RealLiteralMode() { set_real_literals_f32(false); }
};
}  // namespace

Lowered lower_model(const Universe& u, const LowerOptions& opt_in) {
    LowerOptions opt = opt_in;
    // the RK probe only reads u0 > 1/2; for uniforms (w + 1/2) 2^-32 that is the top bit of the word w (sde_u0_t)
    opt.u0_bits = !opt.strict && opt.scheme == SCHEME_RK && (opt.rng == RNG_SOBOL_XOR || opt.rng == RNG_PHILOX);
    const int P = u.P(), K = u.K();
    if (opt.f32 && opt.strict)
        throw ExprError{"dtype f32 needs arithmetic=\"fast\" (strict reproduces the reference's f64 operation order)"};
    RealLiteralMode literal_mode(opt.f32);
    if (P == 0) throw ExprError{"no equations"};
    if (u.T() < 2) throw ExprError{"time_steps needs at least two points"};
    if (opt.scheme == SCHEME_RK && K == 0)
        throw ExprError{"runge-kutta needs at least one stochastic factor (the reference panics: rng index 0 out of bounds, src/rng/pseudo.rs:53-58)"};
    Lowered L;
    const bool chacha = opt.rng == RNG_PSEUDO || opt.rng == RNG_SOBOL_CP;
    const bool sobol = opt.rng == RNG_SOBOL_CP || opt.rng == RNG_SOBOL_XOR || opt.rng == RNG_SOBOL_RAW;
    const int KK = K > 0 ? K : 1;
    L.ch = (chacha && K > 0) ? 8 / gcd_int(8, K) : 1;
    if (opt.rng == RNG_PHILOX && K > 0) L.ch = 4 / gcd_int(4, K);   // step groups start on a Philox block (4 draws) boundary

    // ---- steady-state cache position: where does one step leave the cache?
    {
        std::ostringstream scratch;
        StepGen probe(u, opt);
        CacheAt exit_state = probe.generate(OLD, scratch);
        L.enter_eq = (exit_state == NEXT);                   // this step's t+1 is the next step's t
    }
    std::ostringstream body;
    StepGen gen(u, opt);
    gen.generate(L.enter_eq ? CUR : OLD, body);
    // Per-step constants that do not move along the time grid become literals (no table read in the step loop):
    // bit-identical values always; under arithmetic=fast also values that agree to 2^-44 relative, or to what moving the
    // time points by a few ulp would change (StepGen::slot_tol: a uniform grid k/D gives dt and sqrt(dt) values a few ulp
    // of t_end apart), replaced by their median — far inside that mode's stated <= ~3 ulp per step.
    std::vector<std::string> slot_macro(gen.slots.size());
    std::vector<std::string> table_slots;
    for (size_t i = 0; i < gen.slots.size(); ++i) {
        std::vector<double> v = gen.slot_values[i];
        std::sort(v.begin(), v.end());
        const double lo = v.front(), hi = v.back(), med = v[v.size() / 2];
        const bool same = lo == hi;
        const bool close = !opt.strict && std::isfinite(lo) && std::isfinite(hi) &&
                           (hi - lo) <= std::max(std::ldexp(std::fabs(med), -44), gen.slot_tol[i]);
        if (same || close) {
            // f64 plans: as a uniform-datapath value (sde_uc, sde_expr_helpers.cuh) so that fma(SLOT, z, g) reads two
            // register pairs, not three
            // register pairs, not three.  Slot 0 of a process is the addend the factor terms accumulate onto: a plain literal.
            const bool multiplier = gen.slot_mult[i];
            slot_macro[i] = (opt.f32 || !multiplier) ? "(" + format_real(med) + ")" : "sde_uc(" + format_real(med) + ")";
        } else {
            slot_macro[i] = "((sde_real)ss[" + std::to_string(4 + table_slots.size()) + "])";
            table_slots.push_back(gen.slots[i]);
        }
    }
    const int nslot = (int)table_slots.size();

    // ---- launch shape
    // every kernel decomposes the Sobol index as CTA part ^ warp part ^ lane part and sizes shared arrays per warp:
    // a CTA is a whole number of warps, at most 1024 threads
    if (opt.block != 0 && (opt.block < 32 || opt.block > 1024 || opt.block % 32 != 0))
        throw ExprError{"block_threads must be 0 (auto) or a multiple of 32 in [32, 1024], got " + std::to_string(opt.block)};
    L.block = opt.block > 0 ? opt.block : 256;
    // steps unrolled per loop trip: whole ChaCha blocks, and >= 4 for small models so that loads, constants and
    // the state-independent inverse-CDF chains of neighbouring steps overlap
    L.unr = std::max(L.ch, K <= 2 ? 4 : (K <= 4 ? 2 : 1));
    // full paths in reference order: straight from registers as aligned 256-bit stores when a group is 4 steps
    // and no ChaCha block alignment ties groups to the time origin; otherwise through the shared-memory transpose
    const bool sector_stores_ok = opt.out == OUT_PATHS_NTP && !chacha && opt.rng != RNG_PHILOX && L.unr == 4 && P <= 8 && opt.direct != 0 && !opt.f32;
    L.direct = sector_stores_ok && (L.block % 128) == 0;
    // Sobol-driven full paths whose tables fit in shared memory for the whole time grid: persistent warps, no time tiles
    {
        const int S = u.T() - 1;
        bool wide = false;                                    // 1024-entry log table (128 KB): XOR digital shift only
        bool lane_global = false;                             // lane table in global memory (SDE_RES_LANE_GLOBAL)
        auto resident_smem = [&](int block, int nslot_) {     // mirrors the SDE_SMEM_* macros of sde_sim_resident.cuh
            const size_t sk = (size_t)S * K;
            size_t icdf = (opt.icdf == 1) ? (wide ? (size_t)1024 * 2 * 8 * 8 : (size_t)(128 * 2 * 8 + 64) * 8) : 0;
            const size_t nq = (sk + 3 + 3) / 4;               // SDE_NQ: quads of 4 dimensions (+ room for the per-CTA offset)
            return icdf + (size_t)S * (4 + nslot_) * 8 + (lane_global ? 0 : nq * 512) + (size_t)(block / 32) * nq * 16;
        };
        const bool eligible = sector_stores_ok && (opt.rng == RNG_SOBOL_XOR || opt.rng == RNG_SOBOL_RAW) && K >= 1;
        if (eligible && opt.direct != 1 && opt.direct != 3 && opt.direct != 4) {
            // 24 warps per SM (6 per scheduler), measured on B200 (profiles/r2_c2_ab.md): the step loop needs 64-80 registers,
            // and since its FP64 instructions stopped paying for three register-pair operands (sde_uc, FP32-unit seeds)
            // more resident warps pay again: 485 G path-steps/s sustained at 768 threads, 478 at 512, 480 at 1024
            int block = opt.block > 0 ? opt.block : 768;
            block = std::max(32, std::min(1024, (block / 32) * 32));   // warps are autonomous: any whole number of warps
            while (block > 32 && resident_smem(block, nslot) > 200 * 1024) block = std::max(32, (block / 64) * 32);
            if (resident_smem(block, nslot) > 200 * 1024 && opt.direct == 2) {
                // the tables of the whole time grid do not fit (C3: 2000 dimensions) and the persistent kernel was asked for
                // (ntp_direct = 3): keep the lane table in global memory, prepared by the host; the per-warp Sobol part (16
                // bytes per quad and warp) decides how many warps fit.  Not selected automatically: on C3 it runs at
                // 24.3 ms against 24.0-24.1 ms for the time-tiled kernel with the four-buffer hand-over (same box)
                lane_global = true;
                block = opt.block > 0 ? std::max(32, std::min(1024, (opt.block / 32) * 32)) : 768;
                while (block > 128 && resident_smem(block, nslot) > 200 * 1024) block -= 128;
            }
            if (resident_smem(block, nslot) <= 200 * 1024) {
                wide = opt.icdf == 1 && opt.rng == RNG_SOBOL_XOR && !std::getenv("SDE_B200_NO_WIDE_TABLE");
                if (wide && resident_smem(block, nslot) > 220 * 1024) wide = false;
                L.icdf_wide = wide;
                L.resident = true;
                L.lane_global = lane_global;
                // mirrors SDE_RES_FOLD of sde_sim_resident.cuh: the host prepares the global lane table in the same form
                L.res_fold = opt.rng == RNG_SOBOL_XOR && opt.icdf == 1 && opt.scheme != SCHEME_RK &&
                             std::all_of(u.factor_is_wiener.begin(), u.factor_is_wiener.end(), [](bool b) { return b; });
                L.direct = true;
                L.block = block;
                L.smem_bytes = resident_smem(block, nslot);
            }
        }
        if (opt.direct == 2 && !L.resident)
            throw ExprError{"the persistent-warp kernel needs Sobol (xor / none) full-path NTP output, K <= 2 and (T-1)*K*128 B of shared memory"};
    }
    // Opt-in (direct = 3): full paths of a model with an even number of processes (every row segment is 16-byte aligned)
    // staged per lane in shared memory and handed to the copy engine once per tile (SDE_TMA, sde_sim_kernel.cuh).  The
    // scattered 256-bit sector stores cost ~26 load/store data-pipe wavefronts per warp instruction (C3: the kernel is 12 %
    // slower with the stores than without, HBM at 35 %), the staged form 8 — but cp.async.bulk takes uniform-register
    // addresses, so per-lane segments serialise into a 32-trip loop of ~10 instructions per warp and tile: C3 26.9 ms
    // against 25.4 ms with sector stores on the same box.  Not selected automatically.
    {
        const bool tma_ok = opt.out == OUT_PATHS_NTP && !opt.f32 && (P % 2) == 0 && !L.resident && !gen.matrix;
        if (opt.direct == 3 && !tma_ok)
            throw ExprError{"bulk-copy stores need [N][T][P] f64 paths of a model with an even number of processes"};
        if (tma_ok && opt.direct == 3) { L.tma = true; L.direct = false; }
        // direct = 4: one 2-D tensor-map store per warp and 128-byte box row (SDE_TMA == 2): P = 2 or 4, groups of 4 steps
        const bool tma2_ok = tma_ok && (P == 2 || P == 4) && L.unr == 4;
        if (opt.direct == 4 && !tma2_ok)
            throw ExprError{"tensor-map stores need [N][T][P] f64 paths of a model with 2 or 4 processes and K <= 2"};
        if (tma2_ok && opt.direct == 4) { L.tma = true; L.tma2 = true; L.direct = false; }
    }
    // Wide linear model reduced to terminal values / moments under the XOR digital shift: the correlation product runs
    // on the FP64 tensor path (sde_sim_wide.cuh).  A lane keeps 8 wide_mt paths x (2 NB state + NKK draw) doubles in
    // registers: two row tiles per warp when that is <= 64 doubles, else one.
    if (gen.matrix && opt.wide_mma != 0 && opt.rng == RNG_SOBOL_XOR && opt.out != OUT_PATHS_TPN && !opt.f32) {
        const int NB = (P + 7) / 8, NKK = (K + 3) / 4, S = u.T() - 1;
        const int per_tile = 2 * NB + NKK;
        int mt = 2 * per_tile <= 64 ? 2 : (per_tile <= 64 ? 1 : 0);
        if (const char* g = std::getenv("SDE_B200_WIDE_MT")) mt = std::min(mt, std::max(1, std::atoi(g)));   // tuning
        // measured on B200 (64 assets, tools/run_c4.py): 16 warps per SM at a 128-register cap (a few hundred bytes of
        // L1-resident spills) beat 8 warps at 255 registers, 3.72 vs 3.47 G path-steps/s — the dependent DMMA chains
        // and the 19-deep inverse-normal chains want the extra warps more than the registers
        int block = opt.block > 0 ? std::max(32, std::min(1024, (opt.block / 32) * 32)) : 512;
        auto wide_smem = [&](int blk, bool wide_tab) {       // mirrors the SDE_SMEM_* macros of sde_sim_wide.cuh
            size_t icdf = (opt.icdf == 1) ? (wide_tab ? (size_t)1024 * 2 * 8 * 8 : (size_t)(128 * 2 * 8 + 64) * 8) : 0;
            size_t mom = opt.out == OUT_MOMENTS ? (size_t)(blk / 32) * NB * 8 * 3 * 8 : 0;
            return icdf + (size_t)NB * NKK * 32 * 8 + (size_t)NB * 8 * 8 + mom + (size_t)(blk / 32) * NKK * 4 * 4;
        };
        if (mt > 0 && (size_t)S * K < (1u << 31) && wide_smem(block, false) <= 220 * 1024) {
            L.wide = true;
            L.wide_mt = mt; L.wide_nb = NB; L.wide_nkk = NKK;
            L.icdf_wide = opt.icdf == 1 && wide_smem(block, true) <= 220 * 1024 && !std::getenv("SDE_B200_NO_WIDE_TABLE");
            L.block = block;
            L.min_blocks = 1;
            L.smem_bytes = wide_smem(block, L.icdf_wide);
            L.tt = S;
        }
    }
    if (opt.wide_mma == 1 && !L.wide)
        throw ExprError{"the tensor-core kernel for wide models needs a linear Levy model with P, K >= 16 (arithmetic=\"fast\", euler), "
                        "sobol with scramble=\"xor\", f64, and [N][T][P] paths / terminal / moments output"};
    if (!L.wide && !L.resident) {
    int tt = opt.tile_steps;
    if (tt <= 0) {
        tt = 32;
        if (opt.out == OUT_PATHS_NTP && !L.direct) tt = std::max(1, 32 / P);   // staging tile: 32 paths x (tt*P) doubles per warp
        if (L.tma) tt = std::max(tt, 16);                                      // bulk copies of >= 256 bytes per lane and tile
        if (L.tma2) tt = 16;                                                   // two 8-step boxes (P = 2) per tile
        if (sobol) tt = std::min(tt, std::max(1, 128 / KK));   // lane-table slice: 2 x tt*K*128 B of shared memory
    }
    tt = std::max(L.unr, (tt / L.unr) * L.unr);
    if (opt.tile_steps <= 0 && (opt.out != OUT_PATHS_NTP || L.direct)) {
        // prefer a tile length that divides the step count (no partial last tile): search multiples of the
        // unroll group in [tt/2, tt*9/8], nearest to the target first.  The transposing NTP path keeps tt*P a
        // divisor/multiple of 32 instead: that is what makes its flush cheap.
        const int S = u.T() - 1;
        int best = 0;
        for (int cand = L.unr; cand <= tt + tt / 8; cand += L.unr)
            if (cand * 2 >= tt && S % cand == 0 && (best == 0 || std::abs(cand - tt) <= std::abs(best - tt))) best = cand;
        if (best) tt = best;
    }
    L.tt = tt;
    const int ts = tt + (L.direct ? 3 : 0);                  // SDE_TS
    int nstage = 2;                                          // SDE_NSTAGE: stage buffers of the tile hand-over (see sde_sim_kernel.cuh)
    auto smem_for = [&](int block) {   // mirrors the SDE_SMEM_* macros of sde_sim_kernel.cuh
        const int nw = block / 32;
        size_t icdf = (opt.icdf == 1 && opt.rng != RNG_INJECT) ? (size_t)(128 * 2 * 8 + 64) * 8 : 0;   // SDE_ICDF_TABLE_DOUBLES
        const size_t tile_ld = L.tma ? (size_t)((((tt * P) + 3) & ~3) + 2) : (size_t)((tt * P) | 1);   // SDE_TILE_LD
        size_t tile = (opt.out == OUT_PATHS_NTP && !L.direct) ? (size_t)nw * 32 * tile_ld * (opt.f32 ? 4 : 8) : 0;
        if (L.tma2) { icdf = (icdf + 1023) & ~(size_t)1023; tile = (size_t)nw * 2 * 4096; }
        tile = (tile + 7) & ~(size_t)7;
        size_t stage = (size_t)ts * (4 + nslot) * 8 + (sobol ? (size_t)ts * KK * nw * 4 + (size_t)ts * KK * 32 * 4 : 0);
        size_t mom = (opt.out == OUT_MOMENTS) ? (size_t)nw * 3 * 8 : 0;
        return icdf + tile + (size_t)nstage * stage + mom + (nstage == 4 ? 16 : 0);
    };
    if (opt.block <= 0 && !L.direct) while (L.block > 32 && smem_for(L.block) > 200 * 1024) L.block /= 2;
    {
        // four stage buffers + mbarrier hand-over (a tile of slack between the warps of a CTA) where they are cheap: the
        // staged tables of four tiles within a quarter of the CTA's shared memory and the CTA still at <= 100 KB
        nstage = 4;
        const size_t with4 = smem_for(L.block);
        nstage = 2;
        const size_t with2 = smem_for(L.block);
        // (not for the ChaCha-driven modes: their step loop is several thousand instructions, and warps that drift apart
        //  stop sharing the instruction cache — C5 on the reference's ChaCha8 stream 382 -> 418 ms with four buffers;
        //  for the same reason only small models, K <= 2: the measured cases)
        bool use4 = !chacha && K <= 2 && u.T() - 1 > 2 * tt && with4 <= 112 * 1024 && (with4 - with2) * 2 <= with4;   // two CTAs per SM still fit
        if (const char* g = std::getenv("SDE_B200_NSTAGE")) use4 = std::atoi(g) == 4 && with4 <= 200 * 1024;   // tuning
        nstage = use4 ? 4 : 2;
    }
    L.nstage = nstage;
    L.smem_bytes = smem_for(L.block);
    if (L.smem_bytes > 227 * 1024) throw ExprError{"model too large for the shared-memory staging tile (P = " + std::to_string(P) + ")"};
    {
        // measured on B200 (tools/sweep.py, profiles/r1_sweep_c2.json): with 4-step groups the kernel wants ~100 registers
        // (the next tile's prefetched words stay in registers behind the last group); 2 CTAs x 256 threads at a
        // 128-register cap beat 3-4 CTAs at 85 / 64 registers (409 vs 403 / 376 G path-steps/s on C2)
        // (matrix-form steps keep the draws and the rows in local arrays: ~85 registers whatever P and K are;
        //  measured on the 64-asset basket: 3 CTAs x 256 threads 1.46-1.51 vs 1 CTA 1.03 G path-steps/s)
        const int regs_est = gen.matrix ? 85 : std::min(255, (L.unr >= 4 ? 96 : 52) + 6 * P + 2 * K + (opt.scheme == SCHEME_RK ? 4 * P : 0));
        int by_regs = std::max(1, 65536 / (L.block * regs_est));
        int by_smem = (int)std::max<size_t>(1, (size_t)(224 * 1024) / std::max<size_t>(L.smem_bytes, 1024));
        int by_threads = std::max(1, 2048 / L.block);
        L.min_blocks = std::max(1, std::min({by_regs, by_smem, by_threads}));
        if (opt.min_blocks > 0) L.min_blocks = std::min(opt.min_blocks, std::min(by_smem, by_threads));
    }

    } else if (L.resident) {
        const int by_smem = (int)std::max<size_t>(1, (size_t)(224 * 1024) / (L.smem_bytes + 1024));
        const int by_threads = std::max(1, 2048 / L.block);
        // one CTA per SM: its warps are the persistent workers (see the block-size note above); opt.min_blocks can ask
        // for more resident CTAs for tuning, within what shared memory and the thread limit allow
        L.min_blocks = 1;
        if (opt.min_blocks > 0) L.min_blocks = std::max(1, std::min(opt.min_blocks, std::min(by_smem, by_threads)));
        L.tt = u.T() - 1;
    }
    // ---- translation unit
    std::ostringstream s;
    s << "// Generated by libsde_b200 (csrc/host/lower.cpp) — model lowered from equation strings.\n";
    for (int p = 0; p < P; ++p) {
        const Process& pr = u.processes[p];
        s << "//   process " << p << " '" << pr.name << "' " << (pr.levy ? "levy" : "algebraic");
        if (pr.levy) for (const Term& t : pr.terms) s << " | (" << t.coeff.source() << ") * " << (t.kind == IncKind::Time ? "dt" : u.factor_names[t.factor]);
        else s << " = " << pr.algebraic.source();
        s << "\n";
    }
    s << "#define SDE_P " << P << "\n#define SDE_K " << K << "\n#define SDE_KK " << KK << "\n";
    s << "#define SDE_SCHEME " << opt.scheme << "\n#define SDE_RNG " << opt.rng << "\n#define SDE_OUT " << opt.out << "\n";
    if (opt.f32) s << "#define SDE_F32 1\n";
    if (const char* d = std::getenv("SDE_B200_DEFINES")) {       // tuning only: "NAME=VALUE;NAME=VALUE" -> #define lines
        std::string item;
        std::istringstream ds(d);
        while (std::getline(ds, item, ';')) {
            if (item.empty()) continue;
            const size_t eq = item.find('=');
            s << "#define " << item.substr(0, eq) << " " << (eq == std::string::npos ? "1" : item.substr(eq + 1)) << "\n";
        }
    }
    s << "#define SDE_ICDF " << opt.icdf << "\n#define SDE_STRICT " << (opt.strict ? 1 : 0) << "\n";
    s << "#define SDE_NEEDS_U0 " << (opt.scheme == SCHEME_RK ? 1 : 0) << "\n";
    if (opt.u0_bits) s << "#define SDE_U0_BITS 1\n";
    s << "#define SDE_BLOCK " << L.block << "\n#define SDE_MIN_BLOCKS " << L.min_blocks << "\n";
    if (L.wide) {
        s << "#define SDE_S " << (u.T() - 1) << "\n#define SDE_WNB " << L.wide_nb << "\n#define SDE_WNKK " << L.wide_nkk
          << "\n#define SDE_WMT " << L.wide_mt << "\n";
        if (L.icdf_wide) s << "#define SDE_ICDF_WIDE 1\n";
    }
    if (L.resident) {
        s << "#define SDE_S " << (u.T() - 1) << "\n";
        if (L.icdf_wide) s << "#define SDE_ICDF_WIDE 1\n";
        if (L.lane_global) s << "#define SDE_RES_LANE_GLOBAL 1\n";
        if (K > 0 && std::all_of(u.factor_is_wiener.begin(), u.factor_is_wiener.end(), [](bool b) { return b; })) s << "#define SDE_ALL_WIENER 1\n";
        if (std::getenv("SDE_B200_DEBUG_NOCOMPUTE")) s << "#define SDE_DEBUG_NOCOMPUTE 1\n";                                              // profiling aid
        if (std::getenv("SDE_B200_DEBUG_NOSCALAR")) s << "#define SDE_DEBUG_NOSCALAR 1\n";                                                // profiling aid
        if (std::getenv("SDE_B200_DEBUG_NOSTORE")) s << "#define SDE_DEBUG_NOSTORE 1\n";                                                  // profiling aid
        if (const char* g = std::getenv("SDE_B200_RES_PIPE")) s << "#define SDE_RES_PIPE " << (std::atoi(g) ? 1 : 0) << "\n";          // tuning
        s << "#define SDE_RES_GRP 4\n";                      // steps per unrolled group: one sector store per lane and process
    }
    s << "#define SDE_TT " << L.tt << "\n#define SDE_CH " << L.ch << "\n#define SDE_UNR " << L.unr << "\n#define SDE_NSLOT " << nslot << "\n#define SDE_DIRECT " << (L.direct ? 1 : 0) << "\n";
    if (L.tma) s << "#define SDE_TMA " << (L.tma2 ? 2 : 1) << "\n";
    if (!L.wide && !L.resident && L.nstage == 4) s << "#define SDE_NSTAGE 4\n";
    s << "#include \"sde_expr_helpers.cuh\"\n#include \"sde_device_icdf.cuh\"\n";
    s << "__device__ __forceinline__ constexpr bool sde_factor_is_wiener(int k) { return ";
    if (K > 0 && std::all_of(u.factor_is_wiener.begin(), u.factor_is_wiener.end(), [](bool b) { return b; })) {
        s << "true";                                         // (also keeps the predicate free when k is a runtime loop index)
    } else {
        bool any = false;
        for (int k = 0; k < K; ++k) if (u.factor_is_wiener[k]) { s << (any ? " || " : "") << "k == " << k; any = true; }
        if (!any) s << "false";
    }
    s << "; }\n";
    s << "// one step of " << (opt.scheme == SCHEME_EULER ? "euler_iteration (src/sim/euler.rs:5-37)" : "runge_kutta_iteration (src/sim/runge_kutta.rs:5-107)")
      << "; cache enters " << (L.enter_eq ? "AT times[t] (stale values, not refreshed)" : "behind times[t] (refreshed at the first evaluation)") << "\n";
    s << "// path-independent per-step constants, evaluated once per step in the tile prologue\n";
    s << "__device__ __forceinline__ void sde_model_step_consts(const double t_cur, const double t_next, const double dt, const double sqrt_dt, double* slots) {\n";
    s << "    (void)t_cur; (void)t_next; (void)dt; (void)sqrt_dt; (void)slots;\n";
    for (int i = 0; i < nslot; ++i) s << "    slots[" << i << "] = " << table_slots[i] << ";\n";
    s << "}\n";
    if (gen.matrix && !L.wide) s << gen.matrix_decl();
    if (L.wide) {
        // loadings in mma.m8n8k4 B-fragment order: entry (j, kk, lane) = M[process 8j + (lane >> 2)][factor 4kk + (lane & 3)]
        const int NB = L.wide_nb, NKK = L.wide_nkk;
        s << "__device__ const double sde_wm[" << NB * NKK * 32 << "] = {";
        for (int j = 0; j < NB; ++j)
            for (int kk = 0; kk < NKK; ++kk)
                for (int l = 0; l < 32; ++l) {
                    const int k = 4 * kk + (l & 3);
                    const double v = k < K ? gen.mat_M[((size_t)j * K + k) * 8 + (l >> 2)] : 0.0;
                    s << ((j || kk || l) ? ", " : "") << format_real(v);
                }
        s << "};\n__device__ const double sde_wa[" << NB * 8 << "] = {";
        for (int i = 0; i < NB * 8; ++i) s << (i ? ", " : "") << format_real(gen.mat_drift[i]);
        s << "};\n// factor steps (of 4) process tile j needs: its loadings are zero beyond\n";
        s << "__device__ __forceinline__ constexpr int sde_wkk_end(int j) {\n    constexpr int e[" << NB << "] = {";
        for (int j = 0; j < NB; ++j) s << (j ? ", " : "") << (gen.mat_kend[j] + 3) / 4;
        s << "};\n    return e[j];\n}\n";
    }
    if (!gen.model_consts.empty()) {
        s << "__constant__ sde_real sde_mc[" << gen.model_consts.size() << "] = {";
        for (size_t i = 0; i < gen.model_consts.size(); ++i) s << (i ? ", " : "") << format_real(gen.model_consts[i]);
        s << "};\n";
    }
    for (size_t i = 0; i < slot_macro.size(); ++i) s << "#define SDE_SLOT_" << i << " " << slot_macro[i] << "\n";
    s << "__device__ __forceinline__ void sde_model_step(sde_real (&row)[SDE_P], sde_real (&c)[SDE_P], double& ct, const sde_real (&zu)[SDE_KK],\n"
         "                                               const sde_u0_t u0, const double* __restrict__ ss) {\n";
    s << "    const sde_real t_cur = (sde_real)ss[0], t_next = (sde_real)ss[1], dt = (sde_real)ss[2], sqrt_dt = (sde_real)ss[3];\n";
    s << "    (void)u0; (void)t_cur; (void)t_next; (void)dt; (void)sqrt_dt; (void)zu; (void)ct;\n";
    if (L.wide) s << "    // (the step lives in sde_sim_wide.cuh: X_i *= 1 + a_i dt + sqrt(dt) sum_k M[i][k] z_k on the FP64 tensor path)\n";
    else s << gen.prelude() << body.str();
    s << "}\n#include \"" << (L.wide ? "sde_sim_wide.cuh" : (L.resident ? "sde_sim_resident.cuh" : "sde_sim_kernel.cuh")) << "\"\n";
    L.source = s.str();
    return L;
}

}  // namespace sde
