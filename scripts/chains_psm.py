#!/usr/bin/env python
"""Config C5 of BASELINE.json: independent chains sharded over the GPUs of one node, PSM accumulated on the
tensor cores and NCCL-reduced together with the K histogram and split-R-hat.

  python scripts/chains_psm.py --chains 64 --n 20000                     # one GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
      scripts/chains_psm.py --chains 64 --n 20000                        # one rank per GPU
Prints one JSON line on rank 0."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains", type=int, default=64)
    ap.add_argument("--n", type=int, default=20000)
    ap.add_argument("--p", type=int, default=256)
    ap.add_argument("--cats", type=int, default=5)
    ap.add_argument("--k-true", type=int, default=50)
    ap.add_argument("--burnin", type=int, default=10)
    ap.add_argument("--iterations", type=int, default=40)
    ap.add_argument("--psm-mode", default="allreduce", choices=["allreduce", "reduce_scatter", "none"])
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from split_and_merge_gibbs_sampling_b200 import Chain, Psm, chains as mc, step_many
    from split_and_merge_gibbs_sampling_b200.synth import ham_mix_gen
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    X, labels, cent, attr = ham_mix_gen(a.n, a.p, a.cats, a.k_true, s=0.5, seed=1)  # same data on every rank
    v, w = np.full(a.p, 6.0), np.full(a.p, 0.25)

    def factory(cid):
        return Chain(X, attr, 1.0, v, w, m=3, L=a.k_true, t=10, r=10, neal8=True, split_merge=True, seed=1000 + cid,
                     device=local, compact_init=True, data_u8=True, pool_size=a.n)

    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = mc.run_chains(a.n, a.chains, factory, a.burnin, a.iterations,
                        psm_factory=(lambda t: Psm(a.n, device=local, capacity_sweeps=256, external=t)),
                        dist=dist if world > 1 else None, device=dev, step_many=step_many, psm_mode=a.psm_mode)
    torch.cuda.synchronize()
    secs = time.perf_counter() - t0
    tt = torch.tensor([secs], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    secs = float(tt[0])
    sweeps = a.chains * (a.burnin + a.iterations)
    if rank == 0:
        psm = out["psm"]
        diag_ok = None if psm is None or a.psm_mode != "allreduce" else bool((torch.diagonal(psm) == out["psm_draws"]).all())
        print(json.dumps({"config": f"{a.chains} chains n={a.n} p={a.p} K_true={a.k_true}", "n_gpus": world,
                          "chain_sweeps_per_s": sweeps / secs, "seconds": secs, "rhat_K": out["rhat_K"],
                          "rhat_loglik": out["rhat_loglik"], "K_mode": int(np.argmax(out["K_hist"])),
                          "psm_draws": out["psm_draws"], "psm_diag_ok": diag_ok, "psm_mode": a.psm_mode}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
