"""Where the end-to-end time of run_markov_chain goes (diagnostic, GPU only)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from split_and_merge_gibbs_sampling_b200 import run_markov_chain
from split_and_merge_gibbs_sampling_b200.synth import ham_mix_gen
n, p = 100000, 256
X, labels, cent, attr = ham_mix_gen(n, p, 5, 50, s=0.5, seed=1)
Xd = np.asfortranarray(X.astype(np.float64))
v = np.full(p, 6.0); w = np.full(p, 0.25)
kw = dict(m=3, L=50, c_i=labels, t=10, r=10, neal8=True, split_merge=True, seed=1)
run_markov_chain(Xd, attr, 1.0, v, w, iterations=1, burnin=0, **kw)
for its in (20, 200):
    t0 = time.perf_counter()
    res = run_markov_chain(Xd, attr, 1.0, v, w, iterations=its, burnin=5, verbose=3, **kw)
    t1 = time.perf_counter()
    print(f"iterations={its}: wall {t1 - t0:.4f} s, C loop {res['seconds']:.4f} s, sweeps/s {(its + 5) / (t1 - t0):.1f}")
    del res
