"""Per-sweep phase times over the bench's timed window (diagnostic)."""
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
prev = ch.stats()
for it in range(200):
    ch.step(1)
    tm = ch.timings()
    st = ch.stats()
    print(it, "scan %.4f sm %.4f phi %.4f tot %.4f" % (tm["scan_ms"], tm["split_merge_ms"], tm["update_phi_ms"], tm["total_ms"]),
          st["scan_rounds"] - prev["scan_rounds"], st["scan_events"] - prev["scan_events"], st["sm_accepted"] - prev["sm_accepted"])
    prev = st
ch.close()
