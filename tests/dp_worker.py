"""Worker of tests/test_dp_cpu.py: one of WORLD_SIZE gloo ranks.  Exercises the host side of
the data-parallel path without a GPU: rendezvous, id broadcast, column sharding, and the
identity the NCCL path relies on -- sum over ranks of shard gradients == full-batch gradients
(params), shard gradients concatenate (data), loss sums."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from oracle import yota_oracle as O  # noqa: E402
from yota_b200 import dp  # noqa: E402

rank, nranks, _ = dp.world()
d = dp.init_process_group()
# the ncclUniqueId broadcast path (bytes through a uint8 tensor)
t = torch.arange(128, dtype=torch.uint8) if rank == 0 else torch.zeros(128, dtype=torch.uint8)
d.broadcast(t, src=0)
assert t.tolist() == list(range(128))
assert dp.allreduce_host_max(float(rank)) == nranks - 1
dp.barrier()

L, W, B = 3, 16, 10  # ragged: 10 columns over 2 ranks... and over 4 ranks
args = [a.astype(np.float64) for a in cases.c3_args(L, W, B)]
f = cases.make_c3(L, B)
full_val, full_g = O.grad(f, *args, dtype=np.float64)
local = list(args)
local[-2], local[-1] = dp.shard_columns(args[-2], rank, nranks), dp.shard_columns(args[-1], rank, nranks)
val, g = O.grad(f, *local, dtype=np.float64)
tv = torch.tensor([float(val)], dtype=torch.float64)
d.all_reduce(tv)
assert abs(tv.item() - float(full_val)) < 1e-12
for i in range(2 * L):  # parameter gradients: all-reduce(sum)
    tg = torch.from_numpy(np.ascontiguousarray(g[1 + i]))
    d.all_reduce(tg)
    assert np.allclose(tg.numpy(), full_g[1 + i], rtol=1e-12, atol=1e-14), i
# data gradients stay sharded: they are the matching columns of the full gradient
for i in (2 * L, 2 * L + 1):
    assert np.allclose(g[1 + i], dp.shard_columns(full_g[1 + i], rank, nranks), rtol=1e-12, atol=1e-14)
widths = [dp.shard_columns(args[-1], r, nranks).shape[1] for r in range(nranks)]
assert sum(widths) == B and max(widths) - min(widths) <= 1
dp.barrier()
if rank == 0:
    print("DP_OK", nranks)
