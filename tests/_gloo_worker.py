"""Worker for tests/test_host_cpu.py::test_sharded_eval_logic_gloo_world2 (run under torchrun)."""
import os
import sys

import torch.distributed as dist

sys.path.insert(0, os.environ["REPO"])
from mocopci_b200 import sharding  # noqa: E402

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
pairs = sharding.shard_pairs(10, rank, world)
acc = sharding.MetricAccumulator(3)
for p in pairs:
    acc.add(frame=p % 3, cd=float(p), emd=2.0 * p)
tot = acc.reduce(dist)
if rank == 0:
    exp_cd = [sum(float(p) for p in range(10) if p % 3 == f) for f in range(3)]
    assert tot["count"] == [4, 3, 3], tot
    assert all(abs(a - b) < 1e-9 for a, b in zip(tot["cd_sum"], exp_cd)), tot
    assert all(abs(a - 2 * b) < 1e-9 for a, b in zip(tot["emd_sum"], exp_cd)), tot
    assert abs(tot["cd_mean"][0] - exp_cd[0] / 4) < 1e-9
    print("OK", pairs)
dist.destroy_process_group()
