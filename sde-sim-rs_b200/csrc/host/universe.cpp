// universe.cpp — equation grammar of the reference (src/proc/util.rs:52-166) producing the
// product's model IR.  Two passes: (1) scan every equation into raw term strings while
// filling the stochastic registry in first-appearance order, (2) compile all expressions
// against the final process registry (the reference resolves names lazily at evaluation
// time against the cache map, src/filtration.rs:72-78 — same binding, later duplicates win).
#include "universe.h"

#include <cmath>

namespace sde {
namespace {

struct RawTerm { std::string coeff, inc; };
struct RawProcess { std::string name; bool levy; std::vector<RawTerm> terms; std::string rhs; };

std::string strip(const std::string& s) {
    size_t a = s.find_first_not_of(" \t\r\n\f\v");
    if (a == std::string::npos) return "";
    size_t b = s.find_last_not_of(" \t\r\n\f\v");
    return s.substr(a, b - a + 1);
}
std::string lstrip(const std::string& s) {
    size_t a = s.find_first_not_of(" \t\r\n\f\v");
    return a == std::string::npos ? "" : s.substr(a);
}

// `delimited(char('('), balanced_parens, char(')'))` (util.rs:16-38): text[open] == '(';
// returns the index one past the matching ')' or npos when unbalanced.
size_t match_paren(const std::string& text, size_t open) {
    int depth = 0;
    for (size_t i = open; i < text.size(); ++i) {
        if (text[i] == '(') ++depth;
        else if (text[i] == ')' && --depth == 0) return i + 1;
    }
    return std::string::npos;
}

RawProcess scan_equation(const std::string& eq) {
    // util.rs:73-76 — exactly one '='
    size_t first = eq.find('=');
    if (first == std::string::npos || eq.find('=', first + 1) != std::string::npos) throw ExprError{"Missing '='"};
    std::string lhs = strip(eq.substr(0, first)), rhs = strip(eq.substr(first + 1));
    RawProcess rp;
    rp.levy = !lhs.empty() && lhs[0] == 'd';                 // util.rs:80-82 (so `delta = ...` is an SDE named `elta`)
    rp.name = rp.levy ? lhs.substr(1) : lhs;
    rp.rhs = rhs;
    if (!rp.levy) return rp;
    std::string rest = rhs;
    for (;;) {                                               // util.rs:87-123
        size_t open = rest.find('(');
        if (open == std::string::npos) break;
        size_t close = match_paren(rest, open);
        if (close == std::string::npos) throw ExprError{"Unbalanced parentheses in coefficient"};
        RawTerm term;
        term.coeff = strip(rest.substr(open + 1, close - open - 2));
        std::string after = lstrip(rest.substr(close));
        if (after.empty() || after[0] != '*') break;         // silently stops (util.rs:93-96)
        std::string tok = lstrip(after.substr(1));
        if (tok.compare(0, 2, "dN") == 0) {
            size_t o = tok.find('(');
            if (o == std::string::npos) throw ExprError{"dN missing opening bracket"};
            size_t c = match_paren(tok, o);
            if (c == std::string::npos) throw ExprError{"Unbalanced parentheses in dN intensity"};
            term.inc = tok.substr(0, c);
            rest = tok.substr(c);
        } else {
            size_t sp = tok.find(' ');
            if (sp == std::string::npos) sp = tok.size();
            term.inc = tok.substr(0, sp);
            rest = tok.substr(sp);
        }
        rp.terms.push_back(std::move(term));
    }
    return rp;
}

}  // namespace

Universe parse_equations(const std::vector<std::string>& equations, const std::vector<double>& times) {
    if (times.size() < 1) throw ExprError{"time_steps must not be empty"};
    for (size_t i = 0; i < times.size(); ++i) {
        if (!std::isfinite(times[i])) throw ExprError{"time_steps must be finite"};
        if (i && !(times[i] > times[i - 1])) throw ExprError{"time_steps must be strictly increasing"};
    }
    Universe u;
    u.times = times;
    std::vector<RawProcess> raw;
    raw.reserve(equations.size());
    std::unordered_map<std::string, int> factor_reg;
    for (const std::string& eq : equations) {
        RawProcess rp = scan_equation(eq);
        for (const RawTerm& t : rp.terms) {                  // build_incrementor, util.rs:136-166
            if (t.inc == "dt") continue;
            auto ins = factor_reg.emplace(t.inc, (int)factor_reg.size());
            bool wiener = t.inc.compare(0, 2, "dW") == 0, poisson = t.inc.compare(0, 2, "dN") == 0;
            if (!wiener && !poisson) throw ExprError{"Unknown incrementor type: " + t.inc};
            if (ins.second) { u.factor_names.push_back(t.inc); u.factor_is_wiener.push_back(wiener); }
        }
        raw.push_back(std::move(rp));
    }
    for (size_t i = 0; i < raw.size(); ++i) {                // ProcessUniverse::new, mod.rs:71-89
        u.process_registry[raw[i].name] = (int)i;
        (raw[i].levy ? u.levy_indices : u.algebraic_indices).push_back((int)i);
    }
    auto compile = [&](const std::string& src, const std::string& what) {
        try { return Expr::parse(src, u.process_registry); }
        catch (ExprError& e) { throw ExprError{what + e.msg}; }
    };
    for (RawProcess& rp : raw) {
        Process p;
        p.name = rp.name;
        p.levy = rp.levy;
        if (rp.levy) {
            for (RawTerm& rt : rp.terms) {
                Term t;
                t.coeff = compile(rt.coeff, "Math error in coefficient: ");
                if (rt.inc == "dt") t.kind = IncKind::Time;
                else {
                    t.factor = factor_reg.at(rt.inc);
                    if (u.factor_is_wiener[t.factor]) t.kind = IncKind::Wiener;
                    else {
                        t.kind = IncKind::Poisson;             // extract_lambda, util.rs:41-50
                        size_t o = rt.inc.find('(');
                        size_t c = match_paren(rt.inc, o);
                        std::string lam = strip(rt.inc.substr(o + 1, c - o - 2));
                        t.lambda = compile(lam, "Math error in jump lambda '" + lam + "': ");
                    }
                }
                p.terms.push_back(std::move(t));
            }
        } else {
            p.algebraic = compile(rp.rhs, "");
        }
        u.processes.push_back(std::move(p));
    }
    return u;
}

}  // namespace sde
