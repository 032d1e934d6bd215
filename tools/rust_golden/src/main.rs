//! Golden vectors from the unmodified reference crate (see Cargo.toml).
//!
//! Everything below goes through `pub` items of `sde_sim_rs` only:
//!   rng::pseudo::PseudoRng::{new, sample}            (src/rng/pseudo.rs:14,35)
//!   rng::sobol::{SobolEngine::{new, next_path}, SobolRng::new}   (src/rng/sobol.rs:15,23,36)
//!   rng::BaseRng                                     (src/rng/mod.rs:5) — also implemented here to inject uniforms
//!   proc::util::parse_equations                      (src/proc/util.rs:52)
//!   proc::increment::{WienerIncrementor, PoissonJumpIncrementor, Incrementor}   (src/proc/increment.rs:6,74,121)
//!   func::Function::{new, eval}                      (src/func.rs:18,32)
//!   filtration::ScenarioFiltration::{new, get}       (src/filtration.rs:22,56)
//!   sim::euler::euler_iteration, sim::runge_kutta::runge_kutta_iteration   (src/sim/euler.rs:5, src/sim/runge_kutta.rs:5)
//! The scenario loop of sim::simulate (src/sim/mod.rs:41-88) is restated here sequentially with a FIXED seed instead
//! of `rand::rng().random()` (src/sim/mod.rs:28-29) — the two deviations the B200 engine documents: with one thread,
//! scenario s takes Sobol point s + 5 (src/rng/sobol.rs:17,41-44).
//!
//! Every f64 is written as the decimal value of its bit pattern (u64), so the fixture is exact.
use ordered_float::OrderedFloat;
use sde_sim_rs::filtration::ScenarioFiltration;
use sde_sim_rs::func::Function;
use sde_sim_rs::proc::increment::{Incrementor, PoissonJumpIncrementor, WienerIncrementor};
use sde_sim_rs::proc::util::parse_equations;
use sde_sim_rs::proc::{Process, ProcessUniverse};
use sde_sim_rs::rng::pseudo::PseudoRng;
use sde_sim_rs::rng::sobol::{SobolEngine, SobolRng};
use sde_sim_rs::rng::BaseRng;
use sde_sim_rs::sim::euler::euler_iteration;
use sde_sim_rs::sim::runge_kutta::runge_kutta_iteration;
use std::collections::HashMap;
use std::fmt::Write as _;
use std::sync::{Arc, Mutex};

/// BaseRng that replays given uniforms: u[t * k + idx].
struct InjectRng {
    k: usize,
    u: Vec<f64>,
}
impl BaseRng for InjectRng {
    fn sample(&mut self, time_idx: usize, increment_idx: usize) -> f64 {
        self.u[time_idx * self.k + increment_idx]
    }
}

fn bits(v: &[f64]) -> String {
    let mut s = String::from("[");
    for (i, x) in v.iter().enumerate() {
        if i > 0 {
            s.push(',');
        }
        write!(s, "{}", x.to_bits()).unwrap();
    }
    s.push(']');
    s
}
fn jstr(x: &str) -> String {
    let mut s = String::from("\"");
    for c in x.chars() {
        match c {
            '"' => s.push_str("\\\""),
            '\\' => s.push_str("\\\\"),
            _ => s.push(c),
        }
    }
    s.push('"');
    s
}
fn jstrs(v: &[String]) -> String {
    format!("[{}]", v.iter().map(|x| jstr(x)).collect::<Vec<_>>().join(","))
}
fn grid(d: usize, steps: usize) -> Vec<f64> {
    (0..=steps).map(|k| k as f64 / d as f64).collect()
}
fn ordered(t: &[f64]) -> Vec<OrderedFloat<f64>> {
    t.iter().copied().map(OrderedFloat).collect()
}
fn eqs(v: &[&str]) -> Vec<String> {
    v.iter().map(|s| s.to_string()).collect()
}

/// The body of sim::simulate's per-scenario closure (src/sim/mod.rs:44-86), scenarios in order, fixed seed.
fn run_paths(
    universe: &ProcessUniverse,
    times: &[OrderedFloat<f64>],
    init: &HashMap<String, f64>,
    n: u64,
    scheme: &str,
    rng_method: &str,
    seed: u64,
    inject: Option<&[f64]>, // [n][S][K]
) -> Vec<f64> {
    let k = universe.stochastic_registry.len();
    let steps = times.len() - 1;
    let p = universe.processes.len();
    let engine = match (rng_method, inject) {
        ("sobol", None) => Some(Arc::new(Mutex::new(SobolEngine::new(steps * k)))),
        _ => None,
    };
    let mut out = Vec::with_capacity(n as usize * times.len() * p);
    for s_idx in 0..n {
        let local = universe.clone();
        let mut f = ScenarioFiltration::new(s_idx as i32, local.clone(), times.to_vec(), init.clone());
        let mut rng: Box<dyn BaseRng> = match (rng_method, inject) {
            (_, Some(u)) => Box::new(InjectRng { k, u: u[s_idx as usize * steps * k..(s_idx as usize + 1) * steps * k].to_vec() }),
            ("sobol", None) => Box::new(SobolRng::new(s_idx.wrapping_add(seed), Arc::clone(engine.as_ref().unwrap()), k, times.len())),
            _ => Box::new(PseudoRng::new(s_idx.wrapping_add(seed), k)),
        };
        for t_idx in 0..steps {
            match scheme {
                "euler" => euler_iteration(&mut f, &local, t_idx, rng.as_mut()),
                "runge-kutta" => runge_kutta_iteration(&mut f, &local, t_idx, rng.as_mut()),
                _ => unimplemented!(),
            }
        }
        for t in 0..times.len() {
            for pi in 0..p {
                out.push(f.get(t, pi));
            }
        }
    }
    out
}

fn main() {
    let path = std::env::args().nth(1).unwrap_or_else(|| "rust_v1.json".to_string());
    let mut j = String::new();
    j.push_str("{\n\"format\": \"sde-golden-1\",\n\"reference\": \"sde-sim-rs 0.5.1, unmodified, single thread\",\n");

    // ---- 1. ChaCha8 f64 streams through PseudoRng (K = 3: draws are consumed in (t, k) order, src/rng/pseudo.rs:22-31)
    j.push_str("\"pseudo_f64\": [\n");
    let seeds: [u64; 6] = [0, 1, 42, 123456789, u64::MAX, 42 + 7];
    for (i, seed) in seeds.iter().enumerate() {
        let (k, steps) = (3usize, 24usize);
        let mut r = PseudoRng::new(*seed, k);
        let mut v = Vec::new();
        for t in 0..steps {
            for q in 0..k {
                v.push(r.sample(t, q));
            }
        }
        write!(j, "{{\"seed\": {}, \"K\": {}, \"bits\": {}}}{}\n", seed, k, bits(&v), if i + 1 < seeds.len() { "," } else { "" }).unwrap();
    }
    j.push_str("],\n");

    // ---- 2. raw Sobol points after skip(5) (src/rng/sobol.rs:15-25): first points of several dimensionalities
    j.push_str("\"sobol_raw\": [\n");
    let cases: [(usize, usize); 5] = [(1, 64), (4, 64), (252, 24), (2000, 6), (16128, 2)];
    for (i, (dims, npts)) in cases.iter().enumerate() {
        let mut e = SobolEngine::new(*dims);
        let mut rows = Vec::new();
        for _ in 0..*npts {
            rows.push(bits(&e.next_path().expect("sobol exhausted")));
        }
        write!(j, "{{\"dims\": {}, \"first_point_index\": 5, \"bits\": [{}]}}{}\n", dims, rows.join(","), if i + 1 < cases.len() { "," } else { "" }).unwrap();
    }
    j.push_str("],\n");

    // ---- 3. shifted uniforms of SobolRng (src/rng/sobol.rs:35-53,62-79): scenario s uses seed + s and point s + 5
    j.push_str("\"sobol_shifted\": [\n");
    {
        let (k, t_len, n, seed) = (2usize, 41usize, 12u64, 42u64);
        let engine = Arc::new(Mutex::new(SobolEngine::new((t_len - 1) * k)));
        let mut rows = Vec::new();
        for s in 0..n {
            let mut r = SobolRng::new(s.wrapping_add(seed), Arc::clone(&engine), k, t_len);
            let mut v = Vec::new();
            for t in 0..t_len - 1 {
                for q in 0..k {
                    v.push(r.sample(t, q));
                }
            }
            rows.push(bits(&v));
        }
        write!(j, "{{\"seed\": {}, \"K\": {}, \"T\": {}, \"n_paths\": {}, \"bits\": [{}]}}\n", seed, k, t_len, n, rows.join(",")).unwrap();
    }
    j.push_str("],\n");

    // ---- 4. inverse CDFs through the public incrementors with dt = 1 (src/proc/increment.rs:89-97,137-148,160-200)
    {
        let times = ordered(&[0.0, 1.0]);
        let uni = parse_equations(&eqs(&["dX = ( 1.0 ) * dW1"]), times.clone()).unwrap();
        let mut f = ScenarioFiltration::new(0, uni.clone(), times.clone(), HashMap::new());
        let w = WienerIncrementor::new(0, times.clone());
        let mut ps: Vec<f64> = vec![0.5, 0.975, 0.025, 0.875, 0.375, 1e-9, 0.7090754154265618, 1.0 - 1e-12, 2f64.powi(-53), 1.0 - 2f64.powi(-53)];
        for i in 1..400 {
            ps.push(i as f64 / 400.0);
        }
        for e in 2..53 {
            ps.push(2f64.powi(-e));
            ps.push(1.0 - 2f64.powi(-e));
        }
        let mut rng = PseudoRng::new(7, 1);
        for t in 0..600 {
            ps.push(rng.sample(t, 0));
        }
        let mut zs = Vec::new();
        for p in &ps {
            let mut r = InjectRng { k: 1, u: vec![*p] };
            zs.push(w.sample(0, &mut f, &mut r)); // sqrt(1.0) * icdf(p)
        }
        write!(j, "\"icdf_normal\": {{\"p_bits\": {}, \"z_bits\": {}}},\n", bits(&ps), bits(&zs)).unwrap();
        let lams = [0.05, 0.3, 1.0, 3.0, 7.5, 40.0, 250.0, 0.0, -1.0];
        let (mut us, mut ls, mut ks) = (Vec::new(), Vec::new(), Vec::new());
        for lam in lams.iter() {
            let lf = Box::new(Function::new(&format!("{:?}", lam)).unwrap());
            let pj = PoissonJumpIncrementor::new(0, lf, times.clone());
            for u in [0.0, 0.01, 0.1, 0.5, 0.9, 0.95, 0.96, 0.99, 0.999, 0.9999, 1.0 - 1e-15].iter() {
                let mut r = InjectRng { k: 1, u: vec![*u] };
                us.push(*u);
                ls.push(*lam);
                ks.push(pj.sample(0, &mut f, &mut r));
            }
        }
        write!(j, "\"icdf_poisson\": {{\"u_bits\": {}, \"lambda_bits\": {}, \"k_bits\": {}}},\n", bits(&us), bits(&ls), bits(&ks)).unwrap();
    }

    // ---- 5. expression evaluation (fasteval 0.2.4 through func::Function, src/func.rs:18-42)
    {
        let times = ordered(&[0.25, 1.0]);
        let uni = parse_equations(&eqs(&["dx = ( 0.0 ) * dt", "dy = ( 0.0 ) * dt", "dz = ( 0.0 ) * dt"]), times.clone()).unwrap();
        let init = HashMap::from([("x".to_string(), 1.75), ("y".to_string(), -0.6), ("z".to_string(), 3.0)]);
        let mut f = ScenarioFiltration::new(0, uni.clone(), times.clone(), init);
        let exprs = [
            "1 + 2 * 3", "2 ^ 3 ^ 2", "-2 ^ 2", "2 * -x", "x - y - z", "x / y / z", "x % y", "7.5 % 2", "-7.5 % 2", "x ^ y", "x ^ 0.5", "x ^ 2",
            "1 - 2 - 3 * 4 / 5", "x < y", "x > y", "y < x", "x < y or z > 2", "x > y and z < 2", "!(x > y)", "!0", "!2", "x > 1 and y", "0 or y",
            "1k", "2.5M", "3m", "4u", "5n", "6p", "1.5e3", "1e-3 * x", "abs(y)", "sign(y)", "sign(x)", "sign(0)", "int(x)", "int(y)", "int(-1.5)",
            "ceil(x)", "ceil(y)", "floor(x)", "floor(y)", "round(x)", "round(y)", "round(0.5)", "round(1.5)", "round(-0.5)", "round(0.5, x)", "round(0.1, 0.26)",
            "log(x)", "log(100)", "log(2, 8)", "log(e(), x)", "min(x, y)", "max(x, y)", "min(x, y, z)", "max(x, y, z)", "max(x - 100.0, 0.0)",
            "e()", "pi()", "sin(x)", "cos(x)", "tan(x)", "asin(y)", "acos(y)", "atan(x)", "sinh(x)", "cosh(x)", "tanh(x)", "asinh(x)", "acosh(x)", "atanh(y)",
            "sin(t)", "0.5 * cos(t)", "0.01 * x", "2.0 * (0.5 - x)", "(x + y) * (x - y)", "x * y + z", "x + y * z", "e() ^ x", "x ^ -1", "(0 - 1) ^ 0.5",
            "max(y, 0.0) ^ 0.5 * x", "-0.21 * max(z, 0.0) ^ 0.5", "x * x * x", "1 / 3", "x / 3 * 3", "t", "t + x",
        ];
        let (mut srcs, mut vals) = (Vec::new(), Vec::new());
        for e in exprs.iter() {
            srcs.push(e.to_string());
            let v = match Function::new(e) {
                Ok(func) => func.eval(times[0], &mut f).unwrap_or(f64::NAN),
                Err(_) => f64::NAN,
            };
            vals.push(v);
        }
        write!(j, "\"expr\": {{\"t_bits\": {}, \"vars\": {{\"x\": {}, \"y\": {}, \"z\": {}}}, \"src\": {}, \"value_bits\": {}}},\n",
               0.25f64.to_bits(), 1.75f64.to_bits(), (-0.6f64).to_bits(), 3.0f64.to_bits(), jstrs(&srcs), bits(&vals)).unwrap();
    }

    // ---- 6. parser table (src/proc/util.rs:52-166, src/proc/mod.rs:29-44,70-90)
    j.push_str("\"parser\": [\n");
    {
        let tables: Vec<Vec<&str>> = vec![
            vec!["dX1 = ( 0.05 * X1 ) * dt + ( 0.1 * X1) * dW1"],
            vec!["delta = 1.0"],
            vec!["dX = ( 1.0 ) * dt - ( 2.0 ) * dW1"],
            vec!["dX = ( 1.0 ) * dt + ( 2.0 ) + ( 3.0 ) * dW1"],
            vec!["dX = ( X ) * dN1(X) + ( X ) * dN1(2*X)"],
            vec!["dX0 = ( 2.0 * (0.5 - X0) ) * dt + ( 0.1 ) * dN1(X0)", "dX1 = ( 0.05 * X1 ) * dt + ( 0.2 * X1 ) * dW1 + ( 0.5 ) * dN1(X0)", "X2 = max(X1 - 100.0, 0.0)"],
            vec!["dX1 = ( sin(t) ) * dt + (0.01 * X1) * dW1 + (0.001 * X1) * dN1(0.5 * cos(t))", "X2 = max(X1 - 100.0, 0.0)"],
            vec!["dX = ( sin(t) ) * dt + ( 0.01 * X ) * dW1+( 1 ) * dt"],
            vec!["X = 1 = 2"],
            vec!["dX 1.0"],
            vec!["dX = ( 1.0 ) * dQ"],
            vec!["dX = ( 1.0 * dt"],
        ];
        for (i, t) in tables.iter().enumerate() {
            let e = eqs(t);
            let r = parse_equations(&e, ordered(&[0.0, 0.5, 1.0]));
            match r {
                Ok(u) => {
                    let names: Vec<String> = u.processes.iter().map(|p| p.name().to_string()).collect();
                    let levy: Vec<String> = u.processes.iter().map(|p| matches!(p, Process::Levy(_)).to_string()).collect();
                    let terms: Vec<String> = u.processes.iter().map(|p| match p { Process::Levy(l) => l.coefficients.len().to_string(), Process::Algebraic(_) => "0".to_string() }).collect();
                    let mut fac: Vec<(usize, String)> = u.stochastic_registry.iter().map(|(k, v)| (*v, k.clone())).collect();
                    fac.sort();
                    let fnames: Vec<String> = fac.into_iter().map(|(_, k)| k).collect();
                    write!(j, "{{\"equations\": {}, \"ok\": true, \"names\": {}, \"is_levy\": [{}], \"num_terms\": [{}], \"factors\": {}}}", jstrs(&e), jstrs(&names), levy.join(","), terms.join(","), jstrs(&fnames)).unwrap();
                }
                Err(_) => {
                    write!(j, "{{\"equations\": {}, \"ok\": false}}", jstrs(&e)).unwrap();
                }
            }
            j.push_str(if i + 1 < tables.len() { ",\n" } else { "\n" });
        }
    }
    j.push_str("],\n");

    // ---- 7. whole paths: the BASELINE shapes (small N) and the reference's own example models, both schemes
    j.push_str("\"paths\": [\n");
    {
        let gbm = vec!["dX1 = ( 0.05 * X1 ) * dt + ( 0.1 * X1) * dW1"];
        let heston = vec![
            "dS = ( 0.05 * S ) * dt + ( max(v, 0.0)^0.5 * S ) * dW1",
            "dv = ( 2.0 * (0.04 - v) ) * dt + ( -0.21 * max(v, 0.0)^0.5 ) * dW1 + ( 0.2142428528562855 * max(v, 0.0)^0.5 ) * dW2",
        ];
        let ex_py = vec!["dX0 = ( 2.0 * (0.5 - X0) ) * dt + ( 0.1 ) * dN1(X0)", "dX1 = ( 0.05 * X1 ) * dt + ( 0.2 * X1 ) * dW1 + ( 0.5 ) * dN1(X0)", "X2 = max(X1 - 100.0, 0.0)"];
        let ex_rs = vec!["dX1 = ( sin(t) ) * dt + (0.01 * X1) * dW1 + (0.001 * X1) * dN1(0.5 * cos(t))", "X2 = max(X1 - 100.0, 0.0)"];
        let ex_gbm = vec!["dX1 = ( 0.05 * X1 ) * dt + ( 0.1 * X1) * dW1"];
        let arange = |dt: f64, n: usize| -> Vec<f64> { (0..n).map(|i| i as f64 * dt).collect() };
        struct Case<'a> { name: &'a str, eqs: Vec<&'a str>, times: Vec<f64>, init: Vec<(&'a str, f64)>, n: u64, scheme: &'a str, rng: &'a str, seed: u64 }
        let cases = vec![
            Case { name: "C1", eqs: gbm.clone(), times: grid(252, 252), init: vec![("X1", 1.0)], n: 8, scheme: "euler", rng: "pseudo", seed: 42 },
            Case { name: "C1-rk", eqs: gbm.clone(), times: grid(252, 252), init: vec![("X1", 1.0)], n: 8, scheme: "runge-kutta", rng: "pseudo", seed: 42 },
            Case { name: "C2-cp_shift", eqs: gbm.clone(), times: grid(252, 252), init: vec![("X1", 1.0)], n: 8, scheme: "euler", rng: "sobol", seed: 42 },
            Case { name: "C3-pseudo", eqs: heston.clone(), times: grid(1000, 1000), init: vec![("S", 100.0), ("v", 0.04)], n: 4, scheme: "runge-kutta", rng: "pseudo", seed: 42 },
            Case { name: "C3-sobol", eqs: heston.clone(), times: grid(1000, 1000), init: vec![("S", 100.0), ("v", 0.04)], n: 4, scheme: "runge-kutta", rng: "sobol", seed: 42 },
            Case { name: "C3-euler", eqs: heston.clone(), times: grid(1000, 200), init: vec![("S", 100.0), ("v", 0.04)], n: 4, scheme: "euler", rng: "pseudo", seed: 7 },
            Case { name: "example.py", eqs: ex_py.clone(), times: arange(0.01, 300), init: vec![("X0", 0.5), ("X1", 100.0), ("X2", 0.0)], n: 6, scheme: "runge-kutta", rng: "pseudo", seed: 42 },
            Case { name: "example.py-euler", eqs: ex_py.clone(), times: arange(0.01, 300), init: vec![("X0", 0.5), ("X1", 100.0), ("X2", 0.0)], n: 6, scheme: "euler", rng: "sobol", seed: 42 },
            Case { name: "example.rs", eqs: ex_rs.clone(), times: arange(0.1, 301), init: vec![("X1", 100.0), ("X2", 0.0)], n: 6, scheme: "euler", rng: "pseudo", seed: 42 },
            Case { name: "example.rs-rk", eqs: ex_rs.clone(), times: arange(0.1, 301), init: vec![("X1", 100.0), ("X2", 0.0)], n: 6, scheme: "runge-kutta", rng: "pseudo", seed: 42 },
            Case { name: "example_gbm.py", eqs: ex_gbm.clone(), times: arange(0.1, 100), init: vec![("X1", 1.0)], n: 8, scheme: "runge-kutta", rng: "pseudo", seed: 42 },
        ];
        for (i, c) in cases.iter().enumerate() {
            let e = eqs(&c.eqs);
            let times = ordered(&c.times);
            let uni = parse_equations(&e, times.clone()).unwrap();
            let init: HashMap<String, f64> = c.init.iter().map(|(k, v)| (k.to_string(), *v)).collect();
            let vals = run_paths(&uni, &times, &init, c.n, c.scheme, c.rng, c.seed, None);
            let init_s: Vec<String> = c.init.iter().map(|(k, v)| format!("{}: {}", jstr(k), v.to_bits())).collect();
            write!(j, "{{\"name\": {}, \"equations\": {}, \"times_bits\": {}, \"init_bits\": {{{}}}, \"n_paths\": {}, \"scheme\": {}, \"rng_method\": {}, \"seed\": {}, \"values_bits\": {}}}{}\n",
                   jstr(c.name), jstrs(&e), bits(&c.times), init_s.join(", "), c.n, jstr(c.scheme), jstr(c.rng), c.seed, bits(&vals), if i + 1 < cases.len() { "," } else { "" }).unwrap();
        }
    }
    j.push_str("],\n");

    // ---- 8. the SURVEY A.4 worked trace: GBM, 3 steps, injected uniforms, both schemes
    j.push_str("\"trace\": [\n");
    {
        let e = eqs(&["dX1 = ( 0.05 * X1 ) * dt + ( 0.1 * X1) * dW1"]);
        let t = grid(252, 3);
        let times = ordered(&t);
        let uni = parse_equations(&e, times.clone()).unwrap();
        let init = HashMap::from([("X1".to_string(), 1.0)]);
        let u = [0.7090754154265618, 0.46592172228961015, 0.6991432426747317];
        let a = run_paths(&uni, &times, &init, 1, "euler", "pseudo", 0, Some(&u));
        let b = run_paths(&uni, &times, &init, 1, "runge-kutta", "pseudo", 0, Some(&u));
        write!(j, "{{\"scheme\": \"euler\", \"u_bits\": {}, \"values_bits\": {}}},\n{{\"scheme\": \"runge-kutta\", \"u_bits\": {}, \"values_bits\": {}}}\n", bits(&u), bits(&a), bits(&u), bits(&b)).unwrap();
    }
    j.push_str("]\n}\n");
    std::fs::write(&path, j).expect("cannot write the fixture");
    eprintln!("wrote {}", path);
}
