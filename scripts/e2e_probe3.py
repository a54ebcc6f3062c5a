"""run_markov_chain loop time with and without per-iteration snapshots, and the step-wise chain on the same start."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from split_and_merge_gibbs_sampling_b200 import run_markov_chain, Chain
from split_and_merge_gibbs_sampling_b200.synth import ham_mix_gen
n, p = 100000, 256
X, labels, cent, attr = ham_mix_gen(n, p, 5, 50, s=0.5, seed=1)
Xd = np.asfortranarray(X.astype(np.float64))
v = np.full(p, 6.0); w = np.full(p, 0.25)
kw = dict(m=3, L=50, c_i=labels, t=10, r=10, neal8=True, split_merge=True, seed=1)
run_markov_chain(Xd, attr, 1.0, v, w, iterations=1, burnin=0, **kw)
for burn, its in ((5, 400), (400, 5), (5, 400)):
    res = run_markov_chain(Xd, attr, 1.0, v, w, iterations=its, burnin=burn, verbose=3, **kw)
    print(f"burnin={burn} iterations={its}: C loop {res['seconds']:.4f} s = {1e3 * res['seconds'] / (its + burn):.4f} ms/sweep")
    del res
ch = Chain(X, attr, 1.0, v, w, data_u8=True, **kw)
ch.step(10)
ch.step(400)
print(f"step-wise chain from the same start: {ch.last_step_ms() / 400:.4f} ms/sweep (device), K={ch.snapshot(with_phi=False, with_c=False)['K']}")
tm = None
acc = {}
for _ in range(20):
    ch.step(1)
    for k, x in ch.timings().items():
        acc[k] = acc.get(k, 0.0) + x / 20
print({k: round(x, 4) for k, x in acc.items()})
