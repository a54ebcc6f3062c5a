"""Per-sweep phase timings and scan statistics at the metric config (diagnostic, GPU only)."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from split_and_merge_gibbs_sampling_b200 import Chain
from split_and_merge_gibbs_sampling_b200.synth import ham_mix_gen

n, p, cats, kt = 100000, 256, 5, 50
nsw = int(sys.argv[1]) if len(sys.argv) > 1 else 30
X, labels, cent, attr = ham_mix_gen(n, p, cats, kt, s=0.5, seed=1)
ch = Chain(X, attr, 1.0, np.full(p, 6.0), np.full(p, 0.25), m=3, L=kt, t=10, r=10, neal8=True, split_merge=True, seed=1,
           compact_init=True, data_u8=True)
prev = ch.stats()
for it in range(nsw):
    ch.step(1)
    st = ch.stats()
    tm = ch.timings()
    print(it, "K", ch.snapshot(with_phi=False)["K"], "scan %.3f sm %.3f ll %.3f phi %.3f tot %.3f" % (
        tm["scan_ms"], tm["split_merge_ms"], tm["ll_block_ms"], tm["update_phi_ms"], tm["total_ms"]),
        {k: st[k] - prev[k] for k in ("scan_rounds", "scan_events", "births", "deaths")})
    prev = st
    pr = ch.scan_profile()
    print('    scan cycles:', {k: v for k, v in pr.items() if k != '_'})
