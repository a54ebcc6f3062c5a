"""Runs a few sweeps of the other BASELINE.json configs on the GPU (C3 MNIST-shaped, C4 large) and prints timings."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from split_and_merge_gibbs_sampling_b200 import Chain
from split_and_merge_gibbs_sampling_b200.synth import ham_mix_gen

which = sys.argv[1] if len(sys.argv) > 1 else "c3"
cfg = {"c3": dict(n=60000, p=784, cats=2, kt=20, s=0.6, gamma=0.1514657, v=3.0, w=0.5, init="truth"),
       "c4": dict(n=1000000, p=256, cats=5, kt=100, s=0.5, gamma=1.0, v=6.0, w=0.25, init="truth"),
       "c4r": dict(n=1000000, p=256, cats=5, kt=100, s=0.5, gamma=1.0, v=6.0, w=0.25, init="random")}[which]
t0 = time.time()
X, labels, cent, attr = ham_mix_gen(cfg["n"], cfg["p"], cfg["cats"], cfg["kt"], s=cfg["s"], seed=1)
print("generated in %.1f s" % (time.time() - t0), flush=True)
p = cfg["p"]
ch = Chain(X, attr, cfg["gamma"], np.full(p, cfg["v"]), np.full(p, cfg["w"]), m=3, L=cfg["kt"], t=10, r=10, neal8=True,
           split_merge=True, seed=1, compact_init=True, data_u8=True, c_i=labels if cfg["init"] == "truth" else None)
print("chain created in %.1f s" % (time.time() - t0), flush=True)
prev = ch.stats()
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 8):
    ch.step(1)
    st = ch.stats()
    tm = ch.timings()
    s = ch.snapshot(with_phi=False)
    print(it, "K", s["K"], "ll %.1f" % s["loglikelihood"], "ms: ll %.3f aux %.3f scan %.3f phi %.3f sm %.3f loglik %.3f tot %.3f" % (
        tm["ll_block_ms"], tm["aux_ll_ms"], tm["scan_ms"], tm["update_phi_ms"], tm["split_merge_ms"], tm["loglik_ms"], tm["total_ms"]),
        {k: st[k] - prev[k] for k in ("scan_rounds", "scan_events", "births", "deaths")}, flush=True)
    prev = st
from sklearn.metrics import adjusted_rand_score
print("ARI vs truth: %.4f" % adjusted_rand_score(labels, ch.snapshot(with_phi=False)["c_i"]))
