"""Host logic of the multi-GPU path on CPU: chains sharded over ranks, PSM / K-histogram /
trace reductions with torch.distributed (gloo, world_size 2).  The sampler behind each chain is the CPU oracle
(test infrastructure) on the zoo data, so the reduced summaries are those of real chains."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as orc
from split_and_merge_gibbs_sampling_b200 import chains as mc
from split_and_merge_gibbs_sampling_b200.synth import zoo_dataset

N_CHAINS, ITER, BURN = 5, 12, 4
ZOO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "zoo.data")


class OracleChain:
    """Replays one oracle chain (whole run computed up front) through the step/snapshot interface."""

    def __init__(self, cid):
        X, attr, v, w, gamma = zoo_dataset(ZOO)[:5]
        od = orc.OracleData(X, attr, gamma, v, w)
        self.tr = orc.run_chain(od, 3, ITER, 5, None, BURN, 2, 2, True, True, 100 + cid, o=orc.opts(counted=1, stable_hig=1),
                                pool_size=64, kcap=64, keep_c=True)
        self.pos = -BURN - 1

    def step(self, k):
        self.pos += k

    def snapshot(self, with_phi=False):
        i = max(self.pos, 0)
        return {"K": int(self.tr["total_cls"][i]), "c_i": self.tr["c_i"][i], "loglikelihood": float(self.tr["loglikelihood"][i])}


class TorchPsm:
    def __init__(self, mat):
        self.mat = mat

    def push_chain(self, ch):
        c = torch.as_tensor(np.asarray(ch.snapshot()["c_i"]))
        self.mat += (c[:, None] == c[None, :]).to(torch.int32)

    def flush(self):
        pass


def single_process_reference():
    n = zoo_dataset(ZOO)[0].shape[0]
    return n, mc.run_chains(n, N_CHAINS, OracleChain, BURN, ITER, psm_factory=TorchPsm)


def worker(rank, world, port, mode, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = zoo_dataset(ZOO)[0].shape[0]
    out = mc.run_chains(n, N_CHAINS, OracleChain, BURN, ITER, psm_factory=TorchPsm, dist=dist, psm_mode=mode)
    q.put((rank, out["local_chains"], out["K_traces"], out["loglik_traces"], out["rhat_K"], out["K_hist"],
           out["psm"].numpy(), out["psm_rows"]))
    dist.barrier()
    dist.destroy_process_group()


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("mode", ["allreduce", "reduce_scatter"])
def test_two_ranks_match_one_process(mode):
    n, ref = single_process_reference()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=worker, args=(r, 2, port, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got.sort(key=lambda g: g[0])
    assert got[0][1] + got[1][1] == list(range(N_CHAINS))  # every chain owned exactly once
    full = ref["psm"].numpy()
    for rank, _, Ktr, Ltr, rhat, khist, psm, rows in got:
        assert np.array_equal(Ktr, ref["K_traces"])  # traces of ALL chains on every rank
        assert np.array_equal(Ltr, ref["loglik_traces"])
        assert rhat == pytest.approx(ref["rhat_K"], nan_ok=True)
        assert np.array_equal(khist, ref["K_hist"])
        assert khist.sum() == N_CHAINS * ITER
        if mode == "allreduce":
            assert np.array_equal(psm, full)
        else:
            lo, hi = rows
            assert (lo, hi) == mc.row_block(n, 2, rank)
            assert np.array_equal(psm, full[lo:hi])
    assert np.all(np.diag(full) == N_CHAINS * ITER)


def test_shard_chains_and_moments():
    for nc, w in [(64, 8), (5, 2), (3, 4), (1, 1)]:
        parts = [mc.shard_chains(nc, w, r) for r in range(w)]
        assert sum(parts, []) == list(range(nc))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    rng = np.random.default_rng(0)
    x = rng.normal(size=(6, 50))
    m = (0.0, 0.0, 0.0)
    for row in x:
        m = mc.merge_moments(m, mc.chain_moments(row))
    assert m[0] == x.size
    assert m[1] == pytest.approx(x.mean())
    assert m[2] == pytest.approx(((x - x.mean()) ** 2).sum())
    # split R-hat: ~1 for exchangeable chains, large when one chain sits elsewhere
    assert abs(mc.split_rhat(x) - 1.0) < 0.1
    y = x.copy()
    y[0] += 10.0
    assert mc.split_rhat(y) > 2.0
