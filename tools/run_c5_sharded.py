#!/usr/bin/env python3
"""BASELINE config C5 under torchrun: 1-D GBM terminal-only MC (pseudo-random, ChaCha8 streams), `PATHS_PER_GPU` paths x 365
steps per GPU, moments merged over NCCL (all_gather of 3 doubles per rank + Chan merge).  Weak scaling; rank 0 prints one
JSON line with the aggregate rate (device time, max over ranks) and the merged moments against the closed form.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/run_c5_sharded.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sde-sim-rs_b200")):
    sys.path.insert(0, p)
import sde_sim_rs as S  # noqa: E402

GBM = ["dX1 = ( 0.05 * X1 ) * dt + ( 0.1 * X1) * dW1"]
D = 365
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
per_gpu = int(os.environ.get("PATHS_PER_GPU", 1 << 30))
N = per_gpu * world
times = [k / D for k in range(D + 1)]
kw = dict(seed=2024, output="moments", icdf="fast", arithmetic="fast", device=local)

S.simulate_sharded(GBM, times, N, {"X1": 1.0}, "pseudo", "euler", **kw)          # warm-up: NVRTC, NCCL communicator
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 2
e0.record()
for _ in range(reps):
    res = S.simulate_sharded(GBM, times, N, {"X1": 1.0}, "pseudo", "euler", **kw)
e1.record()
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=f"cuda:{local}")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
ms = float(t.item())
m = res.to_numpy()[0]
if rank == 0:
    mu, sig, dt = 0.05, 0.1, 1.0 / D
    mean = (1 + mu * dt) ** D
    var = ((1 + mu * dt) ** 2 + sig * sig * dt) ** D - mean**2
    print(json.dumps({"config": "C5 GBM euler pseudo moments", "n_gpus": world, "paths": N, "steps": D, "ms": ms,
                      "path_steps_per_s": N * D / ms * 1e3, "count": m[0], "mean": m[1], "mean_closed_form": mean,
                      "mean_err_in_standard_errors": abs(m[1] - mean) / (var / N) ** 0.5, "variance_ratio": m[2] / (N - 1) / var,
                      "timing": "CUDA events around 2 calls of simulate_sharded (kernel + finalize + NCCL all_gather + host merge), max over ranks"}),
          flush=True)
if world > 1:
    dist.destroy_process_group()
