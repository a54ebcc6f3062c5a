"""Times the tensor-core PSM flush at the C5 shape (n=2e4, K<=64) and prints achieved integer tensor throughput."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from split_and_merge_gibbs_sampling_b200 import Psm

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 64
K = int(sys.argv[3]) if len(sys.argv) > 3 else 50
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
rng = np.random.default_rng(1)
labels = rng.integers(0, K, size=(T, n)).astype(np.int32)
P = Psm(n, capacity_sweeps=T)
KP = 64 if K <= 64 else (128 if K <= 128 else 256)
for rep in range(reps):
    for c in labels:
        P.push(c)
    P.flush()
    ms = P.info()["last_flush_ms"]
    ops = 2.0 * n * n * KP * T
    print(f"rep {rep}: flush of {T} sweeps n={n} KP={KP}: {ms:.3f} ms  {ops / ms / 1e9:.1f} TOPS (dense u8), "
          f"{ms / T * 1000:.1f} us per sweep; matrix RMW {2 * 4.0 * n * n / ms / 1e6:.0f} GB/s")
chk = P.read(0, 4)
ref = np.stack([(labels[:, i][:, None] == labels).sum(0) for i in range(4)]) * reps
print("check rows 0..3:", bool(np.array_equal(chk, ref)))
