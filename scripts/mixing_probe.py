"""Where the scan's time goes on a chain that keeps moving (Hamming scale s = 1.6 at the metric shape): cycle counters of
the scan CTA (library built with -DSMG_SCAN_PROFILE), rounds / events / births per sweep."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    s = float(sys.argv[1]) if len(sys.argv) > 1 else 1.6
    from split_and_merge_gibbs_sampling_b200 import Chain
    from split_and_merge_gibbs_sampling_b200.synth import ham_mix_gen
    X, labels, cent, attr = ham_mix_gen(100000, 256, 5, 50, s=s, seed=555)
    v, w = np.full(256, 6.0), np.full(256, 0.25)
    ch = Chain(X, attr, 1.0, v, w, m=3, L=50, t=10, r=10, neal8=True, split_merge=True, seed=91, compact_init=True, data_u8=True)
    ch.step(30)
    ch.scan_profile()
    st0 = ch.stats()
    sp0 = ch.scan_spec()
    N = 10
    acc = 0.0
    for _ in range(N):
        ch.step(1)
        acc += ch.timings()["scan_ms"]
    st1 = ch.stats()
    pr = ch.scan_profile()
    d = {k: (st1[k] - st0[k]) / N for k in ("scan_rounds", "scan_events", "births", "deaths")}
    print(f"s={s} K={ch.snapshot(with_phi=False, with_c=False)['K']} scan_ms={acc / N:.2f} per sweep: {d}")
    tot = max(1, pr["loop"])
    print("   cycles/sweep:", {k: int(x / N) for k, x in pr.items() if x}, " shares:", {k: round(x / tot, 3) for k, x in pr.items() if x and k != "loop"})
    print(f"   us per event: {1e3 * acc / N / max(1.0, d['scan_events']):.2f}, cycles per round: {pr['loop'] / N / max(1.0, d['scan_rounds']):.0f}")
    sp1 = ch.scan_spec()
    print("   speculation per sweep:", {k: (sp1[k] - sp0[k]) / N for k in sp1}, "rounds/sweep", d["scan_rounds"])
    ch.close()


if __name__ == "__main__":
    main()
