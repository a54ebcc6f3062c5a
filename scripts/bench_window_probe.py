"""The bench's timed window (metric shape, 40 + 5 untimed sweeps from L = 50 random labels, then 200 sweeps in one call) under
the scan settings given in the environment; prints device ms per sweep and the scan statistics of the window."""
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
ch.step(40)
ch.step(5)
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    st0 = ch.stats()
    ch.step(200)
    ms = ch.last_step_ms() / 200
    st1 = ch.stats()
    print(f"SMG_SCAN_SPEC={os.environ.get('SMG_SCAN_SPEC', 'default')} window {rep}: {ms:.4f} ms per sweep",
          {k: st1[k] - st0[k] for k in ("scan_rounds", "scan_events", "births", "deaths", "sm_accepted")}, flush=True)
ch.close()
