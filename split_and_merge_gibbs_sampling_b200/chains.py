"""Independent chains over GPUs, PSM / cluster-count / R-hat reductions.

What shards naturally -- and only that -- is partitioned: chains are independent (no communication while
sampling), each rank accumulates the posterior similarity matrix of its own chains on its tensor cores, and
at the end of the run `torch.distributed` (NCCL over NVLink on GPUs, gloo in the CPU tests) reduces
  * the PSM counts: all-reduce (sum, int32), or reduce-scatter so that rank g keeps rows [g*n/G, (g+1)*n/G);
  * the histogram of the number of clusters: all-reduce (sum);
  * per-chain traces of K and log-likelihood: all-gather -> split-R-hat (Gelman et al., BDA3 11.4), the
    chain-level (count, mean, M2) moments being merged with Chan's parallel update.
The reference has none of this in C++ (its R scripts call mcclust.ext::comp.psm and LaplacesDemon::ESS after the
run, realdata_analysis/zoo_simulator.R:205-215,339); counts are integers, so the reductions are exact.

The sampler and the PSM accumulator are injected (`chain_factory`, `psm_factory`): on a GPU box they are
`Chain` and `Psm` (CUDA); the CPU tests of this module's host logic inject stand-ins.
"""
import numpy as np


def shard_chains(n_chains, world_size, rank):
    """Contiguous block of chain ids owned by `rank` (sizes differ by at most one)."""
    base, extra = divmod(int(n_chains), int(world_size))
    lo = rank * base + min(rank, extra)
    return list(range(lo, lo + base + (1 if rank < extra else 0)))


def row_block(n, world_size, rank):
    """Rows [lo, hi) of the PSM owned by `rank` after a reduce-scatter (equal blocks; n is padded up by the caller).
    Trailing ranks may own no row at all (lo == hi == n) when ceil(n / world) * rank >= n."""
    per = -(-int(n) // int(world_size))
    lo = min(int(n), rank * per)
    return lo, max(lo, min(int(n), (rank + 1) * per))


def chain_moments(x):
    """(count, mean, M2) of a 1-D trace."""
    x = np.asarray(x, dtype=np.float64)
    m = x.mean() if x.size else 0.0
    return float(x.size), float(m), float(((x - m) ** 2).sum())


def merge_moments(a, b):
    """Chan et al. pairwise merge of two (count, mean, M2) triples."""
    na, ma, sa = a
    nb, mb, sb = b
    n = na + nb
    if n == 0:
        return 0.0, 0.0, 0.0
    d = mb - ma
    return n, ma + d * nb / n, sa + sb + d * d * na * nb / n


def split_rhat(traces):
    """Split-R-hat of an array [chains, draws]: every chain is cut in two halves (BDA3, eq. 11.4)."""
    x = np.asarray(traces, dtype=np.float64)
    if x.ndim != 2 or x.shape[1] < 4:
        return float("nan")
    half = x.shape[1] // 2
    parts = np.concatenate([x[:, :half], x[:, x.shape[1] - half:]], axis=0)  # 2*chains x half
    mom = [chain_moments(p) for p in parts]
    n = half
    means = np.array([m[1] for m in mom])
    W = np.mean([m[2] / (n - 1) for m in mom])
    B = n * means.var(ddof=1)
    if W == 0.0:
        return 1.0 if B == 0.0 else float("inf")
    var_plus = (n - 1) / n * W + B / n
    return float(np.sqrt(var_plus / W))


class NumpyPsm:
    """Stand-in for `Psm` used by the CPU tests of this module (exact integer counts)."""

    def __init__(self, n):
        self.n = n
        self.mat = np.zeros((n, n), dtype=np.int32)

    def push_chain(self, chain):
        c = np.asarray(chain.snapshot(with_phi=False)["c_i"])
        self.mat += (c[:, None] == c[None, :]).astype(np.int32)

    def flush(self):
        pass


def run_chains(n, n_chains, chain_factory, burnin, iterations, thinning=1, psm_factory=None, dist=None, device=None,
               step_many=None, psm_mode="allreduce", kmax=256):
    """Runs `n_chains` independent chains sharded over the ranks of `dist` (a torch.distributed module with an
    initialised process group, or None for one process) and reduces the summaries.

    chain_factory(chain_id) -> object with step(n_iters) and snapshot(with_phi=False) -> {K, c_i, loglikelihood}
    psm_factory(torch_int32_matrix_or_None) -> object with push_chain(chain), flush(); fills the given matrix
    step_many(chains, n_iters): optional overlapped stepping of all local chains (CUDA streams)
    psm_mode: "allreduce" (every rank ends with the full matrix), "reduce_scatter" (rank g keeps its row block)
              or "none".
    Returns a dict: local chain ids, traces of K and log-likelihood for ALL chains, split-R-hat of both,
    histogram of K over all kept draws, the PSM counts (full or this rank's row block) and the number of kept
    draws it was accumulated over."""
    import torch
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    mine = shard_chains(n_chains, world, rank)
    chains = [chain_factory(cid) for cid in mine]
    dev = device if device is not None else torch.device("cpu")
    per = -(-n // world)
    rows = per * world if psm_mode == "reduce_scatter" else n  # equal row blocks for the reduce-scatter
    psm_t = None
    psm = None
    if psm_mode != "none" and psm_factory is not None:
        psm_t = torch.zeros((rows, n), dtype=torch.int32, device=dev)
        psm = psm_factory(psm_t[:n])

    def advance(k):
        if step_many is not None:
            step_many(chains, k)
        else:
            for ch in chains:
                ch.step(k)

    if burnin * thinning > 0:
        advance(burnin * thinning)
    Ktr = np.zeros((len(chains), iterations), dtype=np.float64)
    Ltr = np.zeros((len(chains), iterations), dtype=np.float64)
    for it in range(iterations):
        advance(thinning)
        for q, ch in enumerate(chains):
            s = ch.snapshot(with_phi=False)
            Ktr[q, it] = s["K"]
            Ltr[q, it] = s["loglikelihood"]
            if psm is not None:
                psm.push_chain(ch)
    if psm is not None:
        psm.flush()

    # ---- reductions (collectives only here, never inside a sweep)
    khist = torch.zeros(kmax + 1, dtype=torch.int64, device=dev)
    if Ktr.size:
        if Ktr.max() > kmax:  # the histogram must sum to chains x draws: refuse to drop draws silently
            raise ValueError(f"a chain reached K = {int(Ktr.max())} > kmax = {kmax}: raise kmax")
        khist += torch.bincount(torch.as_tensor(Ktr.astype(np.int64).ravel(), device=dev), minlength=kmax + 1)[:kmax + 1]
    if dist is not None:
        dist.all_reduce(khist, op=dist.ReduceOp.SUM)
        # equal-sized trace blocks for the all-gather: pad this rank's chains up to the largest share
        cmax = -(-n_chains // world)
        buf = torch.full((cmax, 2, iterations), float("nan"), dtype=torch.float64, device=dev)
        if len(chains):
            buf[:len(chains), 0] = torch.as_tensor(Ktr, device=dev)
            buf[:len(chains), 1] = torch.as_tensor(Ltr, device=dev)
        allb = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(allb, buf)
        Kall = np.concatenate([allb[r][:len(shard_chains(n_chains, world, r)), 0].cpu().numpy() for r in range(world)])
        Lall = np.concatenate([allb[r][:len(shard_chains(n_chains, world, r)), 1].cpu().numpy() for r in range(world)])
        if psm_t is not None:
            if psm_mode == "reduce_scatter":
                out = torch.empty((per, n), dtype=torch.int32, device=dev)
                dist.reduce_scatter_tensor(out, psm_t, op=dist.ReduceOp.SUM)
                lo, hi = row_block(n, world, rank)
                psm_t = out[:max(0, hi - lo)]
            else:
                dist.all_reduce(psm_t, op=dist.ReduceOp.SUM)
    else:
        Kall, Lall = Ktr, Ltr
    # chain-level moments merged over all chains (what a streaming R-hat would keep instead of the traces)
    momK = (0.0, 0.0, 0.0)
    for row in Kall:
        momK = merge_moments(momK, chain_moments(row))
    return {
        "rank": rank, "world": world, "local_chains": mine, "K_traces": Kall, "loglik_traces": Lall,
        "rhat_K": split_rhat(Kall), "rhat_loglik": split_rhat(Lall),
        "K_hist": khist.cpu().numpy(), "K_moments": momK,
        "psm": None if psm_t is None else psm_t, "psm_rows": row_block(n, world, rank) if psm_mode == "reduce_scatter" else (0, n),
        "psm_draws": n_chains * iterations, "chains": chains,
    }


def run_chains_native(n, n_chains, chain_factory, burnin, iterations, rank=0, world=1, unique_id=None, device=0,
                      thinning=1, step_many=None, psm_mode="reduce_scatter", kmax=256, psm_capacity=64):
    """The same run with every reduction inside the C++ library (include/smgibbs.h: smg_comm_* / smg_chains_*, NCCL
    bound with dlopen): chains sharded over `world` ranks (one process per GPU), the PSM of the local chains accumulated
    on the tensor cores into a library-owned int32 matrix, then ncclReduceScatter (rank g keeps the rows [g n/G, (g+1)
    n/G)) or ncclAllReduce, ncclAllReduce of the K histogram, ncclAllGather of the half-chain moments -> split-R-hat.
    psm_mode "fused": no reduction at the end at all -- the matrix is distributed by rows over the ranks from the start and
    every flush adds straight into the owners' memories (smg_chains_psm_distribute; n % world == 0).
    `unique_id`: the 128 bytes of Comm.unique_id() from rank 0 (shipped by the caller; not needed when world == 1)."""
    import time
    from .api import Comm, Psm
    mine = shard_chains(n_chains, world, rank)
    chains = [chain_factory(cid) for cid in mine]
    comm = Comm(rank, world, unique_id, device)
    psm = Psm(n, device=device, capacity_sweeps=psm_capacity) if psm_mode != "none" else None
    if psm is not None and psm_mode == "fused":
        # accumulation and reduce-scatter in one kernel: every flush adds its tiles into the owners' memories (NVLink)
        comm.distribute_psm(psm)

    def advance(k):
        if step_many is not None:
            step_many(chains, k)
        else:
            for ch in chains:
                ch.step(k)

    t0 = time.perf_counter()
    if burnin * thinning > 0:
        advance(burnin * thinning)
    burn_s = time.perf_counter() - t0
    Ktr = np.zeros((len(chains), iterations), dtype=np.float64)
    Ltr = np.zeros((len(chains), iterations), dtype=np.float64)
    t_psm = 0.0
    for it in range(iterations):
        advance(thinning)
        for q, ch in enumerate(chains):
            s = ch.snapshot(with_phi=False, with_c=False)
            Ktr[q, it] = s["K"]
            Ltr[q, it] = s["loglikelihood"]
            if psm is not None:
                psm.push_chain(ch)
    if psm is not None:
        tp = time.perf_counter()
        psm.flush()
        t_psm = time.perf_counter() - tp
    sample_s = time.perf_counter() - t0
    out = {"rank": rank, "world": world, "local_chains": mine, "K_local": Ktr, "loglik_local": Ltr, "sample_seconds": sample_s,
           "burnin_seconds": burn_s, "kept_seconds": sample_s - burn_s - t_psm,
           "psm_flush_seconds": t_psm, "psm_draws": n_chains * iterations, "chains": chains, "psm": psm, "comm": comm}
    if psm is not None:
        r0, nr, ms, bus = comm.reduce_psm(psm, psm_mode)
        out.update(psm_rows=(r0, r0 + nr), psm_reduce_ms=ms, psm_bus_gbs=bus, psm_mode=psm_mode)
    if iterations >= 4:
        out["rhat_K"], out["n_chains_total"] = comm.split_rhat(Ktr)
        out["rhat_loglik"], _ = comm.split_rhat(Ltr)
    hist, over = comm.k_histogram(Ktr.astype(np.int32), kmax)
    out["K_hist"], out["K_hist_overflow"] = hist, over
    return out
