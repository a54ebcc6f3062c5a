"""Scan-kernel probe at the metric shape: cycle counters of the scan CTA (smg_debug_scan_profile), rounds / events per
sweep and the scan phase time, for a chain started from random labels and one started from the generating labels."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    from split_and_merge_gibbs_sampling_b200 import Chain
    from split_and_merge_gibbs_sampling_b200.synth import ham_mix_gen
    X, labels, cent, attr = ham_mix_gen(n, 256, 5, 50, s=0.5, seed=1)
    v, w = np.full(256, 6.0), np.full(256, 0.25)
    for start in ("random", "truth"):
        kw = dict(L=50, compact_init=True) if start == "random" else dict(c_i=labels)
        ch = Chain(X, attr, 1.0, v, w, m=3, t=10, r=10, neal8=True, split_merge=True, seed=1, data_u8=True, **kw)
        ch.step(40)
        ch.scan_profile()
        st0 = ch.stats()
        acc = 0.0
        N = 30
        for _ in range(N):
            ch.step(1)
            acc += ch.timings()["scan_ms"]
        st1 = ch.stats()
        pr = ch.scan_profile()
        s = ch.snapshot(with_phi=False)
        cnt = np.bincount(s["c_i"])
        print(f"start={start} K={s['K']} scan_ms={acc / N:.4f} rounds/sweep={(st1['scan_rounds'] - st0['scan_rounds']) / N:.1f} "
              f"events/sweep={(st1['scan_events'] - st0['scan_events']) / N:.2f} births/sweep={(st1['births'] - st0['births']) / N:.2f}")
        print("   cluster sizes (sorted):", np.sort(cnt)[:6], "...", np.sort(cnt)[-4:])
        print("   cycles/sweep:", {k: int(x / N) for k, x in pr.items()})
        ch.close()


if __name__ == "__main__":
    main()
