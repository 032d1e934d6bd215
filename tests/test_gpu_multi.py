"""Several GPUs behind ONE C call (sde_device_plans_create / sde_plan_run_devices / sde_simulate_devices): the host-side
replacement of rayon's par_iter over scenarios (src/sim/mod.rs:41-43,88).  Shard union bit-identical to a one-device run;
moments all-gathered over NCCL and Chan-merged by a device kernel (no host hop).  On a one-GPU box the device is listed
several times (shards on one GPU, peer-copy gather): the sharding / merge logic is what is tested there."""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import GBM_EQ, HESTON_EQ, grid

import sde_sim_rs as S
from sde_sim_rs import _ffi

pytestmark = pytest.mark.gpu


def _devices(min_real=1):
    n = torch.cuda.device_count()
    return list(range(n)) if n >= max(2, min_real) else [0, 0, 0]


@pytest.mark.parametrize("rng_method,scheme,kw", [("sobol", "euler", dict(scramble="xor", icdf="fast", arithmetic="fast")),
                                                   ("pseudo", "runge-kutta", dict())])
def test_device_plans_union_and_merged_moments(rng_method, scheme, kw):
    devs, times, N, init = _devices(), grid(252, 30), 1001, {"S": 100.0, "v": 0.04}
    whole = S.simulate(HESTON_EQ, times, N, init, rng_method, scheme, seed=17, **kw).to_numpy()
    plans = S.DevicePlans(S.Universe(HESTON_EQ, times), scheme, rng_method, devices=devs, **kw)
    outs = plans.run(init, N, seed=17)
    assert [int(o.device.index) for o in outs] == devs
    assert np.array_equal(np.concatenate([o.cpu().numpy() for o in outs]), whole)          # bit for bit
    mplans = S.DevicePlans(S.Universe(HESTON_EQ, times), scheme, rng_method, devices=devs, output="moments", **kw)
    merged = mplans.run(init, N, seed=17)
    assert mplans.collective == ("nccl" if len(set(devs)) == len(devs) and len(devs) > 1 else "peer")
    term = whole[:, -1, :]
    for o in merged:                                                                      # every device holds the merged triples
        m = o.cpu().numpy()
        assert np.array_equal(m, merged[0].cpu().numpy())
        for p in range(2):
            assert m[p, 0] == N
            assert abs(m[p, 1] / term[:, p].mean() - 1) <= 1e-13
            assert abs(m[p, 2] / ((term[:, p] - term[:, p].mean()) ** 2).sum() - 1) <= 1e-10
    # the device merge is the host merge, bit for bit: same shards, same order
    shard_m = []
    for i in range(len(devs)):
        lo, hi = S.shard_range(N, i, len(devs))
        shard_m.append(S.simulate(HESTON_EQ, times, hi - lo, init, rng_method, scheme, seed=17, scenario_offset=lo, output="moments", **kw).to_numpy())
    assert np.array_equal(S.merge_moments(np.stack(shard_m)), merged[0].cpu().numpy())
    assert mplans.collective_ms >= 0.0 and plans.launches >= len(set(devs))


def test_fewer_scenarios_than_devices_and_offsets():
    devs, times = _devices(), grid(252, 9)
    plans = S.DevicePlans(S.Universe(GBM_EQ, times), "euler", "pseudo", devices=devs, output="moments")
    m = plans.run({"X1": 1.0}, 2, seed=3, scenario_offset=1000)[0].cpu().numpy()
    one = S.simulate(GBM_EQ, times, 2, {"X1": 1.0}, "pseudo", "euler", seed=3, scenario_offset=1000).to_numpy()[:, -1, 0]
    assert m[0, 0] == 2 and abs(m[0, 1] / one.mean() - 1) <= 1e-14
    shards = S.simulate_devices(GBM_EQ, times, 2, {"X1": 1.0}, "pseudo", "euler", devices=devs, seed=3)
    assert sum(s.shape[0] for s in shards) == 2


def test_merge_moments_device_matches_host():
    rng = np.random.default_rng(5)
    sh = np.stack([np.stack([[float(n), rng.normal(), abs(rng.normal()) * n] for _ in range(7)]) for n in (5, 0, 11, 1000, 3)])
    got = S.merge_moments_device(torch.from_numpy(sh).cuda()).cpu().numpy()
    assert np.array_equal(got, S.merge_moments(sh))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs (the driver's multi-GPU tier): NCCL all-gather of the moment triples")
def test_nccl_moment_merge_on_real_devices():
    devs, times, N = list(range(torch.cuda.device_count())), grid(365, 73), 1 << 16
    kw = dict(icdf="fast", arithmetic="fast")
    mplans = S.DevicePlans(S.Universe(GBM_EQ, times), "euler", "pseudo", devices=devs, output="moments", **kw)
    assert mplans.collective == "nccl"
    merged = [o.cpu().numpy() for o in mplans.run({"X1": 1.0}, N, seed=42)]
    whole = S.simulate(GBM_EQ, times, N, {"X1": 1.0}, "pseudo", "euler", seed=42, output="terminal", **kw).to_numpy()[:, 0]
    for m in merged:
        assert np.array_equal(m, merged[0])
        assert m[0, 0] == N and abs(m[0, 1] / whole.mean() - 1) <= 1e-13 and abs(m[0, 2] / ((whole - whole.mean()) ** 2).sum() - 1) <= 1e-10
    paths = S.DevicePlans(S.Universe(GBM_EQ, times), "euler", "pseudo", devices=devs, output="terminal", **kw).run({"X1": 1.0}, N, seed=42)
    assert np.array_equal(np.concatenate([p.cpu().numpy() for p in paths])[:, 0], whole)


def test_simulate_keywords_devices_and_compat():
    """SURVEY §8(b)'s keyword-only extras on the drop-in call itself: `devices` shards the scenarios over GPUs inside one C
    call (rayon's par_iter, src/sim/mod.rs:41-43), `compat` is the survey's name for rk_variant."""
    import sde_sim_rs as S
    from conftest import GBM_EQ, grid

    times, init, N = grid(252, 24), {"X1": 1.0}, 1001
    one = S.simulate(GBM_EQ, times, N, init, "sobol", "euler", seed=3, scramble="xor")
    devs = [0, 0, 0] if torch.cuda.device_count() < 2 else list(range(torch.cuda.device_count()))
    shards = S.simulate(GBM_EQ, times, N, init, "sobol", "euler", seed=3, scramble="xor", devices=devs)
    assert isinstance(shards, list) and len(shards) == len(devs)
    got = np.concatenate([f.to_numpy() for f in shards], axis=0)
    assert np.array_equal(got, one.to_numpy())                 # union of the shards: bit-identical
    fr = S.simulate(GBM_EQ, times, 60, init, "sobol", "euler", seed=3, scramble="xor", devices=devs, frame=True)
    ref = S.simulate(GBM_EQ, times, 60, init, "sobol", "euler", seed=3, scramble="xor", frame=True)
    assert list(fr.columns) == list(ref.columns) and len(fr) == len(ref)
    assert np.array_equal(np.asarray(fr["value"]), np.asarray(ref["value"])) and np.array_equal(np.asarray(fr["scenario"]), np.asarray(ref["scenario"]))
    mom = S.simulate(GBM_EQ, times, N, init, "pseudo", "euler", seed=3, output="moments", devices="all")
    full = S.simulate(GBM_EQ, times, N, init, "pseudo", "euler", seed=3, output="terminal").to_numpy()[:, 0]
    m = mom.to_numpy()
    assert m[0, 0] == N and abs(m[0, 1] / full.mean() - 1) < 1e-13
    a = S.simulate(GBM_EQ, times, 50, init, "pseudo", "runge-kutta", seed=5, compat="textbook").to_numpy()
    b = S.simulate(GBM_EQ, times, 50, init, "pseudo", "runge-kutta", seed=5, rk_variant="textbook").to_numpy()
    c = S.simulate(GBM_EQ, times, 50, init, "pseudo", "runge-kutta", seed=5).to_numpy()
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    with pytest.raises(ValueError, match="disagree"):
        S.simulate(GBM_EQ, times, 5, init, "pseudo", "runge-kutta", compat="textbook", rk_variant="reference")
    with pytest.raises(TypeError, match="unexpected keyword"):
        S.simulate(GBM_EQ, times, 5, init, "pseudo", "euler", device_list=[0])
