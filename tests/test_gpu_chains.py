"""Multi-chain path on the GPU: chains stepped together on their own streams (smg_step_many), PSM accumulated on
the tensor cores into a torch-owned matrix, summaries reduced through torch.distributed (NCCL, world_size 1 here;
the 2-rank logic is covered on CPU by test_chains_gloo.py)."""
import os
import socket

import numpy as np
import pytest

from helpers import Problem

pytestmark = pytest.mark.gpu


def test_step_many_equals_one_by_one():
    from split_and_merge_gibbs_sampling_b200 import step_many
    pb = Problem(1500, 32, 4, 5, seed=7)
    a = [pb.chain(L=5, c_i=None, compact_init=True, seed=10 + q) for q in range(3)]
    b = [pb.chain(L=5, c_i=None, compact_init=True, seed=10 + q) for q in range(3)]
    step_many(a, 4)
    for ch in b:
        ch.step(4)
    for x, y in zip(a, b):
        sx, sy = x.snapshot(), y.snapshot()
        assert sx["K"] == sy["K"] and np.array_equal(sx["c_i"], sy["c_i"])
        assert sx["loglikelihood"] == sy["loglikelihood"]
        assert np.array_equal(sx["sigmas"], sy["sigmas"])
    for ch in a + b:
        ch.close()


def test_run_chains_nccl_single_rank():
    import torch
    import torch.distributed as dist
    from split_and_merge_gibbs_sampling_b200 import Psm, chains as mc, step_many
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        pb = Problem(1200, 32, 4, 5, seed=9)
        kept = {}

        class Recorder:  # Chain that remembers every allocation pushed to the PSM
            def __init__(self, cid):
                self.cid = cid
                self.ch = pb.chain(L=5, c_i=None, compact_init=True, seed=50 + cid)
                self.h = self.ch.h

            def snapshot(self, with_phi=False):
                return self.ch.snapshot(with_phi=with_phi)

        class RecPsm(Psm):
            def push_chain(self, chain):
                kept.setdefault(chain.cid, []).append(chain.snapshot()["c_i"].copy())
                super().push_chain(chain)

        out = mc.run_chains(pb.n, 4, Recorder, 3, 6, psm_factory=lambda t: RecPsm(pb.n, capacity_sweeps=8, external=t),
                            dist=dist, device=torch.device("cuda", 0), step_many=step_many, psm_mode="allreduce")
        want = np.zeros((pb.n, pb.n), dtype=np.int64)
        for cs in kept.values():
            for c in cs:
                want += (c[:, None] == c[None, :])
        assert np.array_equal(out["psm"].cpu().numpy(), want.astype(np.int32))
        assert out["K_traces"].shape == (4, 6) and out["K_hist"].sum() == 24
        assert np.isfinite(out["rhat_loglik"])
    finally:
        dist.destroy_process_group()
