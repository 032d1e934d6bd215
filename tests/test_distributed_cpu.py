"""N > 1 host logic on CPU: world-size-2 gloo run of the sharding + moment-merge path that the GPU ranks use
(sde_sim_rs.shard_range / merge_moments around torch.distributed.all_gather).  Per-shard values come from the
CPU oracle here (test infrastructure standing in for the device kernel, which needs a GPU)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GBM_EQ, PKG, ROOT, grid


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, N, tmp):
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sde_sim_rs as S
    from oracle import oracle as orc

    times, init = grid(252, 16), {"X1": 1.0}
    lo, hi = S.shard_range(N, rank, world)
    U = orc.Universe(GBM_EQ, times)
    paths = orc.simulate(U, init, hi - lo, "euler", "sobol", seed=9, scramble="xor", scenario_offset=lo, nthreads=2)
    term = paths[:, -1, :]                                          # [n_local, P]
    mom = np.stack([[term.shape[0], term[:, p].mean(), ((term[:, p] - term[:, p].mean()) ** 2).sum()]
                    for p in range(term.shape[1])])                 # what the device moment kernel returns: [P, 3]
    t = torch.from_numpy(mom)
    gathered = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)                                    # 3*P doubles per rank (NCCL on the GPU box)
    merged = S.merge_moments(torch.stack(gathered).numpy())
    np.save(os.path.join(tmp, f"merged{rank}.npy"), merged)
    np.save(os.path.join(tmp, f"paths{rank}.npy"), paths)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_sharding_and_moment_allgather(tmp_path, oracle):
    N, world = 1001, 2
    mp.spawn(_worker, args=(world, _free_port(), N, str(tmp_path)), nprocs=world, join=True)
    m0, m1 = np.load(tmp_path / "merged0.npy"), np.load(tmp_path / "merged1.npy")
    assert np.array_equal(m0, m1)                                   # every rank holds the same merged estimator
    whole = oracle.simulate(oracle.Universe(GBM_EQ, grid(252, 16)), {"X1": 1.0}, N, "euler", "sobol", seed=9, scramble="xor")
    union = np.concatenate([np.load(tmp_path / "paths0.npy"), np.load(tmp_path / "paths1.npy")])
    assert np.array_equal(union, whole)                             # disjoint index ranges: shard union == single run, bit for bit
    term = whole[:, -1, 0]
    assert m0[0, 0] == N
    assert np.isclose(m0[0, 1], term.mean(), rtol=1e-14)
    assert np.isclose(m0[0, 2], ((term - term.mean()) ** 2).sum(), rtol=1e-11)
