"""Monte-Carlo for the KNN admission-bound estimate (DESIGN.md "tau estimate"): how many refs of
the full cloud fall below the R-th smallest of 32 bucket minima of a 1-in-8 sample, and how often
that is fewer than k (=> exact redo). Pure numpy, no GPU."""
import numpy as np

rng = np.random.default_rng(0)
N, s, buckets, trials = 16384, 8, 32, 100000
m = N // s
for K, Rs in ((8, (4, 5, 6)), (16, (6, 7, 8)), (32, (10, 11, 12))):
    for R in Rs:
        u = rng.random((trials, m), dtype=np.float32)
        t = np.sort(u.reshape(trials, buckets, m // buckets).min(-1), axis=1)[:, R - 1]
        C = (u < t[:, None]).sum(1) + rng.binomial(N - m, t)
        print(f"k={K:2d} R={R}: mean admitted={C.mean():6.1f}  p99={np.percentile(C, 99):5.0f}  "
              f"P(admitted<k)={np.mean(C < K):.5f}")
