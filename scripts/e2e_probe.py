import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from split_and_merge_gibbs_sampling_b200 import run_markov_chain
from split_and_merge_gibbs_sampling_b200.synth import ham_mix_gen
n, p = 100000, 256
X, labels, cent, attr = ham_mix_gen(n, p, 5, 50, s=0.5, seed=1)
Xd = np.asfortranarray(X.astype(np.float64))
v, w = np.full(p, 6.0), np.full(p, 0.25)
for rep in range(3):
    t0 = time.perf_counter()
    res = run_markov_chain(Xd, attr, 1.0, v, w, verbose=3, m=3, iterations=20, L=50, c_i=labels, burnin=5, t=10, r=10, neal8=True,
                           split_merge=True, seed=1)
    print("rep", rep, "total %.4f s" % (time.perf_counter() - t0), "loop seconds", res["seconds"])
