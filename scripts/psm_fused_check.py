"""Fused PSM accumulation + reduce-scatter over peer memory (smg_chains_psm_distribute) against the NCCL reduce-scatter of
per-rank matrices and against numpy, on G ranks:  torchrun --nproc-per-node G scripts/psm_fused_check.py [n] [T]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from split_and_merge_gibbs_sampling_b200 import Comm, Psm
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 96
    uid = [Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    C = Comm(rank, world, uid[0], local)
    rng = np.random.default_rng(100 + rank)
    labels = rng.integers(0, 50, size=(T, n)).astype(np.int32)  # this rank's kept allocations
    # (a) per-rank matrices, NCCL reduce-scatter at the end
    cap = min(T, 256)
    A = Psm(n, device=local, capacity_sweeps=cap)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for c in labels:
        A.push(c)
    A.flush(finalize=False)
    t_loc = time.perf_counter() - t0
    f_loc = A.info()["last_flush_ms"]
    r0, nr, ms, bus = C.reduce_psm(A, "reduce_scatter")
    ref_rows = A.read(r0, nr)
    # (b) distributed matrix: every flush adds into the owners' memories
    B = Psm(n, device=local, capacity_sweeps=cap)
    C.distribute_psm(B)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for c in labels:
        B.push(c)
    B.flush(finalize=False)
    t_acc = time.perf_counter() - t0
    f_acc = B.info()["last_flush_ms"]
    q0, qn, _, _ = C.reduce_psm(B, "reduce_scatter")
    got_rows = B.read(q0, qn)
    ok = (q0, qn) == (r0, nr) and np.array_equal(got_rows, ref_rows)
    # (c) numpy on rank 0 for a small case
    allab = [None] * world
    dist.all_gather_object(allab, labels if n <= 4096 else None)
    if n <= 4096:
        full = np.zeros((n, n), dtype=np.int64)
        for lab in allab:
            for c in lab:
                full += (c[:, None] == c[None, :])
        ok = ok and np.array_equal(got_rows, full[r0:r0 + nr].astype(np.int32))
    flags = [None] * world
    dist.all_gather_object(flags, bool(ok))
    if rank == 0:
        print(f"world={world} n={n} T={T}: fused == NCCL reduce-scatter{' == numpy' if n <= 4096 else ''}: {all(flags)}; "
              f"local accumulation {1e3 * t_loc:.2f} ms (last flush {f_loc:.2f} ms) + NCCL reduce-scatter {ms:.2f} ms ({bus:.0f} GB/s bus); "
              f"fused accumulation {1e3 * t_acc:.2f} ms (last flush {f_acc:.2f} ms), nothing to reduce")
    B.close()
    A.close()
    C.close()
    dist.destroy_process_group()
    if not all(flags):
        sys.exit(1)


if __name__ == "__main__":
    main()
