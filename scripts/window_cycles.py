"""Scan-CTA cycle counters over the bench's timed window (library built with -DSMG_SCAN_PROFILE)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from split_and_merge_gibbs_sampling_b200 import Chain  # noqa: E402
from split_and_merge_gibbs_sampling_b200.synth import ham_mix_gen  # noqa: E402

X, labels, cent, attr = ham_mix_gen(100000, 256, 5, 50, s=0.5, seed=1)
ch = Chain(X, attr, 1.0, np.full(256, 6.0), np.full(256, 0.25), m=3, L=50, t=10, r=10, neal8=True, split_merge=True, seed=1,
           compact_init=True, data_u8=True)
ch.step(45)
for w in range(2):
    ch.scan_profile()
    st0 = ch.stats()
    acc = 0.0
    for it in range(200):
        ch.step(1)
        acc += ch.timings()["scan_ms"]
    st1 = ch.stats()
    pr = ch.scan_profile()
    print("window", w, "scan_ms %.4f" % (acc / 200), "rounds/sweep", (st1["scan_rounds"] - st0["scan_rounds"]) / 200,
          "cycles/sweep", {k: int(v / 200) for k, v in pr.items() if v})
ch.close()
