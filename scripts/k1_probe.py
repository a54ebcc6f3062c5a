"""Likelihood-block (K1) probe: a few sweeps at the given shape, K1 device time per launch by kernel variant
(SMG_K1=tc|t16|c)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    p = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    k_true = int(sys.argv[3]) if len(sys.argv) > 3 else 50
    steps = int(sys.argv[4]) if len(sys.argv) > 4 else 20
    cats = int(sys.argv[5]) if len(sys.argv) > 5 else 5
    from split_and_merge_gibbs_sampling_b200 import Chain
    from split_and_merge_gibbs_sampling_b200.synth import ham_mix_gen
    X, labels, cent, attr = ham_mix_gen(n, p, cats, k_true, s=0.5, seed=1)
    v, w = np.full(p, 6.0), np.full(p, 0.25)
    os.environ["SMG_NO_K1_OVERLAP"] = "1"
    ch = Chain(X, attr, 1.0, v, w, m=3, L=k_true, t=10, r=10, neal8=True, split_merge=True, seed=1, c_i=labels, data_u8=True)
    ch.step(3)
    acc = 0.0
    for _ in range(steps):
        ch.step(1)
        acc += ch.timings()["ll_block_ms"]
    K = ch.snapshot(with_phi=False)["K"]
    pp = (p + 15) // 16 * 16
    ms = acc / steps
    print(f"K1 variant={os.environ.get('SMG_K1', 'tc')} n={n} p={p} K={K}: {1e3 * ms:.1f} us per launch, "
          f"{(n * pp + 8.0 * n * K) / ms / 1e6:.0f} GB/s algorithmic")
    pr = ch.scan_profile()
    if any(pr.values()):
        print("   role cycles (CTA 0) per launch: " + " ".join(f"{k}={v / (steps + 3):.0f}" for k, v in pr.items()))
    ch.close()


if __name__ == "__main__":
    main()
