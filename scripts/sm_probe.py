"""Split-merge / sweep timing probe at the metric config (or any shape): phase timings per sweep under each
split-merge device path, and the cluster kernel's phase cycle counters when the library was built with
-DSMG_SMC_PROFILE (SMG_LIB_PATH=.../libsmgibbs_prof.so)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    p = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    k_true = int(sys.argv[3]) if len(sys.argv) > 3 else 50
    steps = int(sys.argv[4]) if len(sys.argv) > 4 else 100
    modes = sys.argv[5].split(",") if len(sys.argv) > 5 else ["cluster", "coop"]
    from split_and_merge_gibbs_sampling_b200 import Chain
    from split_and_merge_gibbs_sampling_b200.synth import ham_mix_gen
    X, labels, cent, attr = ham_mix_gen(n, p, 5, k_true, s=0.5, seed=1)
    v, w = np.full(p, 6.0), np.full(p, 0.25)
    for mode in modes:
        for overlap in (("0",) if os.environ.get("SM_PROBE_OVERLAP_ONLY") else ("0", "1")):
            os.environ["SMG_SM_MODE"] = mode
            os.environ["SMG_NO_K1_OVERLAP"] = overlap
            ch = Chain(X, attr, 1.0, v, w, m=3, L=k_true, t=10, r=10, neal8=True, split_merge=True, seed=1, compact_init=True,
                       data_u8=True)
            ch.step(40)
            ch.sm_profile()
            keys = None
            acc = None
            dev = 0.0
            t0 = time.perf_counter()
            for _ in range(steps):
                ch.step(1)
                dev += ch.last_step_ms()
                tm = ch.timings()
                if keys is None:
                    keys = list(tm)
                    acc = np.zeros(len(keys))
                acc += np.array([tm[k] for k in keys])
            wall = time.perf_counter() - t0
            t0 = time.perf_counter()
            ch.step(steps)
            wall_batch = time.perf_counter() - t0
            batch_dev = ch.last_step_ms()
            pr = ch.sm_profile()
            st = ch.stats()
            print(f"mode={mode} k1_overlap={'off' if overlap == '1' else 'on'} K={ch.snapshot(with_phi=False)['K']} "
                  f"dev_ms/sweep={dev / steps:.4f} wall_ms/sweep={1e3 * wall / steps:.4f} "
                  f"batched: dev={batch_dev / steps:.4f} wall={1e3 * wall_batch / steps:.4f}")
            print("   phases(ms): " + " ".join(f"{k[:-3]}={x / steps:.4f}" for k, x in zip(keys, acc)))
            if pr.get("launches"):
                L = pr["launches"]
                for side in ("M.", "P.", "D."):
                    print(f"   {side} cycles/launch: " + " ".join(f"{k[2:]}={v / L:.0f}" for k, v in pr.items() if k.startswith(side) and v))
                    print(f"   {side} total: {sum(v for k, v in pr.items() if k.startswith(side)) / L:.0f}")
                print("   " + " ".join(f"{k}={v / L:.1f}" for k, v in pr.items() if k.startswith("nnr")))
            print(f"   accepted={st['sm_accepted']} proposals={st['sm_proposals']} launches/sweep={st['launches'] / st['sweeps']:.1f}")
            ch.close()


if __name__ == "__main__":
    main()
