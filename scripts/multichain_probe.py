"""Aggregate throughput of several independent chains stepped together on one GPU (smg_step_many)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from split_and_merge_gibbs_sampling_b200 import Chain, step_many
from split_and_merge_gibbs_sampling_b200.synth import ham_mix_gen

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
counts = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 8, 32]
p, kt = 256, 50
X, labels, cent, attr = ham_mix_gen(n, p, 5, kt, s=0.5, seed=1)
v, w = np.full(p, 6.0), np.full(p, 0.25)
for nc in counts:
    chains = [Chain(X, attr, 1.0, v, w, m=3, L=kt, t=10, r=10, neal8=True, split_merge=True, seed=100 + q, compact_init=True,
                    data_u8=True, c_i=labels, pool_size=n) for q in range(nc)]
    step_many(chains, 5)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    step_many(chains, 20)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"n={n} chains={nc}: {nc * 20 / dt:.1f} chain-sweeps/s ({dt / 20 * 1000:.3f} ms per round)", flush=True)
    for c in chains:
        c.close()
