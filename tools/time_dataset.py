"""Samples/s of the frame loader on synthetic NL-Drive-sized frames (~120k points per raw frame,
16384 kept): the reference class + the staging of test.py:73-76 against mocopci_b200.data.
python tools/time_dataset.py"""
import importlib
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baseline import fetch_ref  # noqa: E402
from mocopci_b200 import data as ours  # noqa: E402

sys.path.insert(0, fetch_ref.root())
ref = importlib.import_module("data.no_norm_datasets")
with tempfile.TemporaryDirectory() as root:
    rng = np.random.default_rng(0)
    os.makedirs(os.path.join(root, "s"))
    names = []
    for i in range(14):
        n = int(rng.integers(100000, 130000))
        name = f"s/{i:03d}.bin"
        (rng.standard_normal((n, 3)) * 30).astype(np.float32).tofile(os.path.join(root, name))
        names.append(name)
    lst = os.path.join(root, "list.txt")
    with open(lst, "w") as f:
        for s0 in range(8):
            f.write(" ".join(names[s0:s0 + 7]) + "\n")

    def run(ds, stage):
        np.random.seed(0)
        for i in range(2):
            stage(ds[i])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for rep in range(3):
            for i in range(len(ds)):
                stage(ds[i])
        torch.cuda.synchronize()
        return 3 * len(ds) / (time.perf_counter() - t0)

    def ref_stage(sample):   # test.py:73-76 (without the batch dimension the DataLoader adds)
        inp, gt = sample
        return [t.t().cuda().contiguous().float() for t in inp + gt]

    def our_stage(sample):
        inp, gt = sample
        return [t.t().contiguous() for t in inp + gt]

    kw = dict(num_points=16384, interval=4, num_frames=4)
    print(f"reference class + test.py staging : {run(ref.NLDriveDataset(root, lst, **kw), ref_stage):7.1f} samples/s")
    print(f"ours, device=None + same staging  : {run(ours.NLDriveDataset(root, lst, **kw), ref_stage):7.1f} samples/s")
    print(f"ours, device=cuda (pinned block)  : {run(ours.NLDriveDataset(root, lst, device='cuda', **kw), our_stage):7.1f} samples/s")
    print(f"ours, gather_on_device            : {run(ours.NLDriveDataset(root, lst, device='cuda', gather_on_device=True, **kw), our_stage):7.1f} samples/s")
