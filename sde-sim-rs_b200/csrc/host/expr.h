// expr.h — drift/diffusion expression front end of the product (host side).
//
// Replaces func::Function (src/func.rs:5-42): instead of a fasteval Instruction that is
// interpreted against a string-keyed BTreeMap per evaluation (src/filtration.rs:70-79), an
// expression is parsed once into an AST and *lowered to CUDA C++* that reads the cached
// state from registers.  The grammar is the fasteval 0.2.4 subset the reference can reach
// ([3P-unverified]).
#pragma once
#include <string>
#include <unordered_map>
#include <vector>

namespace sde {

struct ExprError {
    std::string msg;
};

enum class Op {
    Const, Time, Var, Neg, Not, Add, Sub, Mul, Div, Mod, Pow, Lt, Gt, Le, Ge, Eq, Ne, And, Or, Call
};

struct ExprNode {
    Op op = Op::Const;
    double value = 0.0;          // Const
    int var = -1;                // Var: process index
    std::string fn;              // Call
    std::vector<int> args;
};

class Expr {
  public:
    // `vars`: process name -> index (a process named "t" shadows time, filtration.rs:72-78).
    static Expr parse(const std::string& src, const std::unordered_map<std::string, int>& vars);

    // CUDA C++ expression over `cache[i]` (state) and `ct` (cached time).  strict = emit
    // separately rounded __dmul_rn/__dadd_rn so FMA contraction cannot move results.
    std::string emit_cuda(bool strict, const std::string& cache_name = "c", const std::string& time_name = "ct") const;

    // Product form (arithmetic = fast lowering): value = kappa * prod_i node(factors[i]).  Multiplications, negations, literal
    // factors and divisions by a literal are flattened into kappa; everything else is a factor (a node index for emit_cuda_node).
    void product_form(double& kappa, std::vector<int>& factors) const;
    std::string emit_cuda_node(int node, bool strict, const std::string& cache_name = "c", const std::string& time_name = "ct") const {
        return emit_node(node, strict, cache_name, time_name);
    }

    bool is_constant() const;
    const std::string& source() const { return src_; }
    const std::vector<ExprNode>& nodes() const { return nodes_; }
    int root() const { return root_; }
    bool uses_var(int idx) const;

  private:
    std::string src_;
    std::vector<ExprNode> nodes_;
    int root_ = -1;
    friend class ExprParserImpl;
    std::string emit_node(int i, bool strict, const std::string& c, const std::string& t) const;
    void product_walk(int i, double& kappa, std::vector<int>& factors) const;
};

std::string format_double(double v);   // shortest round-trip literal usable in CUDA source
// Literals of the generated MODEL code: f64 as format_double; in f32 mode (dtype f32 plans) single-precision
// literals ("0.05f"), so that no double arithmetic sneaks into a float model.  The mode is thread-local and is
// set by lower_model for the duration of one lowering.
std::string format_real(double v);
void set_real_literals_f32(bool f32);
bool real_literals_f32();

}  // namespace sde
