// universe.h — model IR of the product: what proc::ProcessUniverse is in the reference
// (src/proc/mod.rs:7-90), built by the same equation grammar (src/proc/util.rs:52-166).
#pragma once
#include <string>
#include <unordered_map>
#include <vector>

#include "expr.h"

namespace sde {

enum class IncKind { Time, Wiener, Poisson };   // TimeIncrementor / WienerIncrementor / PoissonJumpIncrementor (increment.rs:25-157)

struct Term {
    Expr coeff;              // LevyProcess::coefficients[j]
    IncKind kind = IncKind::Time;
    int factor = -1;         // index in the stochastic registry (dW / dN), -1 for dt
    Expr lambda;             // Poisson intensity expression (increment.rs:108,146)
};

struct Process {
    std::string name;
    bool levy = false;
    std::vector<Term> terms; // Levy
    Expr algebraic;          // AlgebraicProcess::coefficients[0]
};

struct Universe {
    std::vector<Process> processes;
    std::unordered_map<std::string, int> process_registry;   // name -> idx, later duplicates win (mod.rs:74-77)
    std::vector<std::string> factor_names;                    // stochastic_registry in first-appearance order (util.rs:145-146)
    std::vector<bool> factor_is_wiener;
    std::vector<int> levy_indices, algebraic_indices;         // mod.rs:72-81
    std::vector<double> times;

    int P() const { return (int)processes.size(); }
    int K() const { return (int)factor_names.size(); }
    int T() const { return (int)times.size(); }
};

// Throws ExprError with the reference's message text where one exists.
Universe parse_equations(const std::vector<std::string>& equations, const std::vector<double>& times);

}  // namespace sde
