#!/usr/bin/env python
"""bench.py -- Gibbs + split-merge sweeps/sec on synthetic Hamming-mixture data.

One "step" = one iteration of code/launcher.cpp:85-154 on one chain: Neal-8 allocation pass over all n
observations + update_phi + one split-merge proposal (t=r=10) + full log-likelihood.
Workload (BASELINE.json metric config): n=1e5, p=256, 5 categories, K_true=50, m_aux=3, gamma=1,
v=6, w=0.25, one chain per GPU (chains are independent => weak scaling, no data-path collective).

  python bench.py --gpus N --steps K --warmup W          # this repo (CUDA, sm_100a)
  python bench.py --impl reference ...                   # the CPU oracle restating the reference,
                                                         # all host threads (one chain per thread)
Prints ONE JSON line (rank 0).  Keys the driver reads come first; long diagnostics last.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "gibbs_split_merge_sweeps_per_sec"
# dram__bytes_read.sum + dram__bytes_write.sum of one K1 launch at the metric config (ncu --set full)
K1_DRAM_TRAFFIC_BYTES = 31017216  # 27.13 MB read + 3.89 MB written
K1_DRAM_TRAFFIC_SOURCE = "profiles/r01_ncu_full_final2_raw.csv (ncu --set full, one launch at the metric config; kernel unchanged in round 2)"
UNIT = "sweeps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=100000)
    ap.add_argument("--p", type=int, default=256)
    ap.add_argument("--cats", type=int, default=5)
    ap.add_argument("--k-true", type=int, default=50)
    ap.add_argument("--m-aux", type=int, default=3)
    ap.add_argument("--t", type=int, default=10)
    ap.add_argument("--r", type=int, default=10)
    ap.add_argument("--burn", type=int, default=40, help="untimed sweeps from the random start before warm-up (the chain reaches K~50 in about 30)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--multi-chains", type=int, default=8, help="extra: this many chains of the metric shape stepped together on one GPU (0: skip)")
    ap.add_argument("--no-psm", action="store_true", help="skip the tensor-core PSM measurement (n=2e4, C5 shape)")
    ap.add_argument("--c5-chains", type=int, default=64, help="extra: BASELINE config 5 -- this many chains of n=2e4 sharded over the GPUs, PSM + NCCL reductions (0: skip)")
    ap.add_argument("--c5-kept", type=int, default=24, help="kept sweeps per chain in the C5 job")
    ap.add_argument("--c5-psm", default="fused", choices=["fused", "reduce_scatter", "all_reduce"],
                    help="C5: PSM rows distributed over the ranks and every flush added into the owners' memories over NVLink "
                         "(fused), or per-rank matrices reduced with NCCL at the end")
    ap.add_argument("--mixing-s", type=float, default=1.6, help="extra: Hamming scale of a harder data set on which the chain keeps moving (0: skip)")
    ap.add_argument("--no-random-start", action="store_true", help="skip the end-to-end call from the random start")
    ap.add_argument("--cpu-obs", type=int, default=600, help="observations of one pass timed by the faithful CPU baseline")
    ap.add_argument("--seed", type=int, default=1)
    return ap.parse_args()


def workload(a, seed, s=0.5, n=None):
    from split_and_merge_gibbs_sampling_b200.synth import ham_mix_gen
    X, labels, cent, attr = ham_mix_gen(n or a.n, a.p, a.cats, a.k_true, s=s, seed=seed)
    v = np.full(a.p, 6.0)
    w = np.full(a.p, 0.25)
    return X, labels, cent, attr, v, w, 1.0


def config_dict(a):
    return {"workload": f"synthetic Hamming mixture n={a.n} p={a.p} categories={a.cats} K_true={a.k_true} s=0.5",
            "m_aux": a.m_aux, "t": a.t, "r": a.r, "gamma": 1.0, "v": 6.0, "w": 0.25,
            "init": (f"value: L={a.k_true} random labels, {a.burn} untimed sweeps before warm-up (stationary chain, K~{a.k_true}); "
                     f"e2e: ONE run_markov_chain call (burnin=W, iterations=K) started from the generating labels, i.e. stationary "
                     f"from its first sweep like the timed region of `value`; e2e_random_start: the same call from L={a.k_true} "
                     f"random labels with the burn-in inside the call"),
            "chains_per_gpu": 1, "parallelism": "independent chains, one per GPU",
            "cache": "inputs larger than L2: X 25.6 MB + LL block 154 MB + aux pool 1.3 GB are streamed every sweep"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons (B200_PROFILING.md recipe), one process for the whole run; samples are
    time-stamped so that the ones inside a timed region can be picked out."""

    def __init__(self, indices):
        self.indices = list(indices)
        self.proc = None
        self.path = os.path.join(ROOT, "gpurun_out", "clocks.csv")
        self.t0 = None

    def start(self):
        os.makedirs(os.path.dirname(self.path), exist_ok=True)
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap,timestamp")
        try:
            self.f = open(self.path, "w")
            self.t0 = time.time()
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", ",".join(str(i) for i in self.indices)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self, windows):
        """windows: list of (t_start, t_end) wall-clock intervals (time.time()) of the timed regions."""
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        import datetime
        rows = []
        for ln in open(self.path):
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 10:
                continue
            try:
                ts = datetime.datetime.strptime(f[9], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(f[1]), float(f[2]), float(f[3]), f[5:9]))
            except ValueError:
                continue
        if not rows:
            return out
        inside = [r for r in rows if any(a - 0.05 <= r[0] <= b + 0.05 for a, b in windows)]
        scope = "timed regions"
        if len(inside) < 3:  # regions shorter than the sampling period: the samples taken under load during the whole run
            pmax = max(r[3] for r in rows)
            inside = [r for r in rows if r[3] >= 0.5 * pmax] or rows
            scope = "whole process, samples under load (timed regions are shorter than the 50 ms sampling period)"
        reasons = set()
        for r in inside:
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median([r[1] for r in inside])), "sm_max_mhz": float(max(r[2] for r in inside)),
                "reasons": sorted(reasons), "samples": len(inside), "scope": scope}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops", 1590.0)), "measured (MEASURED_PEAKS.json: hbm_gbs burst copy, bf16_tflops burst)"
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


def cpu_baseline(a, X, labels, cent, attr, v, w, gamma, n_chains, cores_label, validate=True):
    """Times the oracle (CPU restatement of the reference) on a bounded sample of the same workload:
    the Neal-8 pass over the first `cpu_obs` observations at the full n (cost per observation does not
    depend on the index), then one full update_phi, one split-merge proposal and the log-likelihood.
    `validate`: additionally ONE complete faithful sweep at n=1e4 next to its own sampled estimate, which checks the
    scaling by n/n_obs used above."""
    import oracle_lib as orc
    od = orc.OracleData(X, attr, gamma, v, w)
    K = int(labels.max() + 1)
    cen = cent.astype(np.float64)
    sig = np.full((K, a.p), 0.5)
    pc, ps = orc.draw_pool(od, 512, 7, o=orc.opts(stable_hig=1))
    out = {}
    for name, counted, nobs in (("faithful", 0, a.cpu_obs), ("counted", 1, min(a.n, 20000))):
        o = orc.opts(counted=counted, stable_hig=1, validate=1)
        t = orc.time_sweep(od, a.m_aux, a.t, a.r, labels, cen, sig, pc, ps, nobs, True, n_chains, 11, o=o)
        sweep_s = t["scan_s"] * (a.n / t["n_obs"]) + t["update_phi_s"] + t["split_merge_s"] + t["loglik_s"]
        out[name] = {"sweep_s": sweep_s, "sweeps_per_s": n_chains / sweep_s, "parts": t}
    f = out["faithful"]
    res = {"value": f["sweeps_per_s"], "unit": UNIT, "cores": n_chains, "kind": "port",
           "sample": (f"faithful oracle (same O(n^2 K) asymptotics as the reference): Neal-8 pass over the first "
                      f"{f['parts']['n_obs']} of {a.n} observations scaled to n, + full update_phi + one split-merge "
                      f"(t=r={a.t}) + log-likelihood; {cores_label}; norm_const2 in the overflow-free form because the "
                      f"reference's GSL 2F1 throws at this cluster size"),
           "sweep_seconds": f["sweep_s"],
           "counted_variant": {"value": out["counted"]["sweeps_per_s"], "sweep_seconds": out["counted"]["sweep_s"],
                               "note": "identical outputs, O(n(K+m)p): counts kept incrementally, sigma terms cached"}}
    if validate:
        # one COMPLETE faithful sweep at n=1e4 (same p, K_true) against the estimate from its first 600 observations
        n2 = 10000
        X2, lab2, cent2, attr2, v2, w2, g2 = workload(a, 77, n=n2)
        od2 = orc.OracleData(X2, attr2, g2, v2, w2)
        o = orc.opts(counted=0, stable_hig=1, validate=1)
        cen2, sig2 = cent2.astype(np.float64), np.full((int(lab2.max() + 1), a.p), 0.5)
        full = orc.time_sweep(od2, a.m_aux, a.t, a.r, lab2, cen2, sig2, pc, ps, n2, True, 1, 11, o=o)
        part = orc.time_sweep(od2, a.m_aux, a.t, a.r, lab2, cen2, sig2, pc, ps, a.cpu_obs, True, 1, 11, o=o)
        full_s = full["scan_s"] + full["update_phi_s"] + full["split_merge_s"] + full["loglik_s"]
        est_s = part["scan_s"] * (n2 / part["n_obs"]) + part["update_phi_s"] + part["split_merge_s"] + part["loglik_s"]
        res["full_sweep_check"] = {"n": n2, "full_sweep_seconds": full_s, "estimate_from_sample_seconds": est_s,
                                   "ratio": est_s / full_s, "note": "one complete faithful sweep (1 core) vs the estimate scaled from its first observations"}
    return res


def run_reference(a):
    """--impl reference: the CPU oracle (the reference cannot be built here: no R/Rcpp/GSL), one chain per
    host thread on all cores, same config / metric / unit."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    X, labels, cent, attr, v, w, gamma = workload(a, a.seed)
    cores = os.cpu_count() or 1
    t0 = time.time()
    vals = []
    cb = None
    reps = max(1, min(a.steps, 2))
    for rep in range(reps):
        cb = cpu_baseline(a, X, labels, cent, attr, v, w, gamma, cores, f"{cores} independent chains on {cores} host threads",
                          validate=(rep == 0))
        vals.append(cb["value"])
    val = float(np.mean(vals))
    line = {"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1000.0 * cores / val, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config_dict(a), "impl": "reference",
            "cpu_baseline": dict(cb, value=val),
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.time() - t0,
            "note": "aggregate over one chain per host thread; each timed step is a bounded sample (see cpu_baseline.sample)"}
    print(json.dumps(line))


def pin_to_share(local, world):
    """Each rank keeps to its own share of the host cores (8 ranks x 16 packing threads on 32 cores was round 1's e2e)."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = max(1, len(cores) // max(1, world))
        mine = cores[local * per:(local + 1) * per] or cores
        os.sched_setaffinity(0, mine)
        return len(mine)
    except Exception:
        return None


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
        return
    import torch
    from split_and_merge_gibbs_sampling_b200 import Chain, Comm, Psm, run_markov_chain, step_many
    from split_and_merge_gibbs_sampling_b200 import chains as mc
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    ncores = pin_to_share(local, world) if world > 1 else None
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    def maxred(vals):
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        if dist:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t]

    sampler = ClockSampler(range(world)) if rank == 0 else None
    if sampler:
        sampler.start()
    windows = []
    X, labels, cent, attr, v, w, gamma = workload(a, a.seed + rank)
    ch = Chain(X, attr, gamma, v, w, m=a.m_aux, L=a.k_true, t=a.t, r=a.r, neal8=True, split_merge=True, seed=a.seed + rank,
               device=local, compact_init=True, data_u8=True)
    ch.step(a.burn)  # leave the random start (untimed)
    ch.step(max(a.warmup, 3))
    st0 = ch.stats()
    barrier()
    # ---- timed region: K sweeps launched back to back by ONE call (no per-sweep host synchronisation), CUDA events on
    #      the chain's stream around them
    tw0 = time.time()
    t0 = time.perf_counter()
    ch.step(a.steps)
    dev_ms = ch.last_step_ms()
    barrier()
    wall = time.perf_counter() - t0
    windows.append((tw0, time.time()))
    st1 = ch.stats()
    K_now = ch.snapshot(with_phi=False, with_c=False)["K"]
    k1_alone_ms = ch.time_ll_block(5)  # the likelihood-block kernel timed alone on the same state (cold caches)
    # every rank's time (min / median / max), the job's time is the max
    tt = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    allms = [tt.clone() for _ in range(world)]
    if dist:
        dist.all_gather(allms, tt)
    rank_ms = sorted(float(x[0]) for x in allms) if dist else [dev_ms]
    dev_ms_max, wall_ms_max = maxred([dev_ms, wall * 1000.0])
    value = world * a.steps / (dev_ms_max / 1000.0)
    # ---- phase times: a short untimed pass stepped one sweep per call (the per-phase events are read after every sweep)
    phase = np.zeros(8)
    nph = 20
    for _ in range(nph):
        ch.step(1)
        tm = ch.timings()
        phase += np.array([tm[k] for k in ("ll_block_ms", "aux_ll_ms", "scan_ms", "update_phi_ms", "split_merge_ms",
                                           "pool_ms", "loglik_ms", "total_ms")])
    phase /= nph
    st2 = ch.stats()
    pp = (a.p + 15) // 16 * 16
    hbm_peak, bf16_peak, peak_src = peaks()
    # ---- roofline of the likelihood-block kernel (K1), timed with CUDA events on the stream it is launched on (it runs
    #      beside the split-merge proposal, on the SMs that kernel leaves free)
    alg_bytes = a.n * pp + 8.0 * a.n * K_now
    ach = alg_bytes / (k1_alone_ms / 1000.0) / 1e9 if k1_alone_ms > 0 else 0.0
    cmp_adds = float(a.n) * K_now * a.p
    roofline = {"kernel": "hamming_ll_block_t16_kernel", "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                "frac": ach / hbm_peak, "traffic": K1_DRAM_TRAFFIC_BYTES, "traffic_source": K1_DRAM_TRAFFIC_SOURCE,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": float(k1_alone_ms),
                "avg_launch_ms_inside_sweep": float(phase[0]),
                "note": ("K1 is bound by the shared-memory pipe, not HBM: n*K*p = %.3g compare-adds per launch done as 4-attribute "
                         "look-ups (1 LDS.64 + 1 DADD per 4 attributes, 64%% of the wavefront peak in the ncu capture). avg_launch_ms: the "
                         "kernel alone on the chain's stream (CUDA events, cold caches, burst peak); inside a sweep it runs on "
                         "a side stream beside the split-merge proposal (avg_launch_ms_inside_sweep, hidden time). The tcgen05 digit-plane GEMM of the same block (SMG_K1=tc, smg_lltc.cuh) is exact and passes "
                         "the parity tests but measured 0.234 ms (tensor pipe 9.7%% active, bound by building the one-hot operand): profiles/r02_summary.md section 4" % cmp_adds)}
    # ---- the dominant kernel of the sweep: the split-merge proposal (latency-bound chain of (t+1) restricted scans)
    nS_typ = 2.0 * a.n / max(K_now, 1)
    gang = int(os.environ.get("SMG_SM_CTAS", "72"))
    roofline_dominant = {"kernel": "sm_chain_kernel (cooperative, %d CTAs) -- default for one chain; sm_cluster_kernel (one 16-CTA cluster) for chains stepped together" % gang,
                         "bound": "latency", "avg_launch_ms": float(phase[4]), "share_of_sweep": float(phase[4] / max(phase[7], 1e-9)),
                         "restricted_scans": a.t + 1, "grid_barriers_max": 4 * (a.t + 1) + 11,
                         "grid_barriers_typical": "2 per settled scan (check || update, update) ... 1 when speculated; ~30 per merge proposal",
                         "ns_per_member_scan": 1e6 * float(phase[4]) / ((a.t + 1) * max(nS_typ, 1.0)),
                         "algorithmic_bytes": (a.t + 2) * nS_typ * (pp + 16),
                         "achieved_GBps": (a.t + 2) * nS_typ * (pp + 16) / max(phase[4], 1e-9) / 1e6,
                         "note": "a chain of dependent phases of 5-15 us (61% of the warp samples wait at a grid barrier for the one CTA or the "
                                 "24 CTAs that work in a phase); the floor of a phase is the double-precision latency of one Beta-rejection "
                                 "draw (~5 us).  Round 2 removed phases instead: settled scans skip the serial decision and the histograms, "
                                 "the scan after a settled one runs check and update at once.  profiles/r02_summary.md sections 2-3"}
    scan = {"ns_per_observation": 1e6 * phase[2] / a.n, "events_per_sweep": (st1["scan_events"] - st0["scan_events"]) / a.steps,
            "avg_ms": float(phase[2])}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": dev_ms_max / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "wall_ms_per_step": wall_ms_max / a.steps,
            "rank_ms_per_step": {"min": rank_ms[0] / a.steps, "median": rank_ms[len(rank_ms) // 2] / a.steps, "max": rank_ms[-1] / a.steps},
            "phase_ms": dict(zip(["ll_block", "aux_ll", "scan", "update_phi", "split_merge", "pool", "loglik", "total"],
                                 [round(float(x), 4) for x in phase])),
            "gpu_launches": int(st1["launches"] - st0["launches"]),
            "K": int(K_now), "sm_accept_rate": (st1["sm_accepted"] - st0["sm_accepted"]) / max(1, st1["sm_proposals"] - st0["sm_proposals"]),
            "config": config_dict(a)}
    ch.close()
    # ---- end to end through the public entry point (run_markov_chain mirror, host buffers in/out):
    # upload of the fp64 column-major matrix, state + pool initialisation, W+K sweeps, and a device->host
    # snapshot (c_i, centres, sigmas, log-lik, accepted) for every kept iteration -- all inside the timed region.
    Xd = np.asfortranarray(X.astype(np.float64))
    # one untimed call first: the first use of the entry point in a process pays one-off costs (page-locked staging
    # buffer, memory-pool growth) that a long-lived R session pays once, not per run
    run_markov_chain(Xd, attr, gamma, v, w, m=a.m_aux, iterations=1, L=a.k_true, c_i=labels, burnin=0, t=a.t, r=a.r,
                     neal8=True, split_merge=True, seed=a.seed + rank, device=local)
    barrier()
    tw0 = time.time()
    t0 = time.perf_counter()
    res = run_markov_chain(Xd, attr, gamma, v, w, m=a.m_aux, iterations=a.steps, L=a.k_true, c_i=labels,
                           burnin=a.warmup, t=a.t, r=a.r, neal8=True, split_merge=True, seed=a.seed + rank, device=local,
                           verbose=3 if os.environ.get("SMG_E2E_TRACE") else 0)
    e2e_local = time.perf_counter() - t0
    barrier()
    windows.append((tw0, time.time()))
    e2e_s = maxred([e2e_local])[0]
    nsw = a.steps + a.warmup
    line["e2e"] = {"value": world * nsw / e2e_s, "unit": UNIT,
                   "h2d_bytes_per_step": int(a.n * pp / nsw),  # the fp64 matrix is packed to u8 codes by host threads first
                   "d2h_bytes_per_step": int((4 * a.n + 192 * pp * 9 + 24) * a.steps / nsw),  # c_i + Kcap rows of centres/sigmas
                   "host_matrix_bytes": int(Xd.nbytes), "seconds": e2e_s, "sweeps": nsw, "loop_seconds": float(res["seconds"]),
                   "init": "generating labels (stationary from the first sweep)",
                   "note": ("whole run_markov_chain call from host buffers: data upload (amortised over the sweeps), "
                            "initialisation incl. the n*m aux pool, W+K sweeps, per-kept-iteration snapshots; "
                            "preceded by one untimed 1-iteration call of the same entry point")}
    if not a.no_random_start:
        # the same call from the declared random start: the burn-in (first pass ~0.3 s, ~30 sweeps of 2-10 ms) is inside
        barrier()
        t0 = time.perf_counter()
        res2 = run_markov_chain(Xd, attr, gamma, v, w, m=a.m_aux, iterations=a.steps, L=a.k_true, burnin=a.burn, t=a.t, r=a.r,
                                neal8=True, split_merge=True, seed=a.seed + rank, device=local)
        e2r = maxred([time.perf_counter() - t0])[0]
        line["e2e_random_start"] = {"value": world * (a.steps + a.burn) / e2r, "unit": UNIT, "seconds": e2r, "sweeps": a.steps + a.burn,
                                    "K_last": int(res2["total_cls"][-1]) if len(res2["total_cls"]) else None,
                                    "note": f"run_markov_chain(L={a.k_true}, c_i=NULL, burnin={a.burn}, iterations={a.steps}): random labels, burn-in inside the call"}
    line["roofline"] = roofline
    line["roofline_dominant"] = roofline_dominant
    line["scan"] = scan
    # ---- extra: a data set on which the chain keeps moving (larger Hamming scale => overlapping clusters)
    if a.mixing_s > 0 and rank == 0:
        Xm, labm, centm, attrm = workload(a, 555, s=a.mixing_s)[:4]
        cm = Chain(Xm, attrm, gamma, v, w, m=a.m_aux, L=a.k_true, t=a.t, r=a.r, neal8=True, split_merge=True, seed=91,
                   device=local, compact_init=True, data_u8=True)
        cm.step(30)
        s0 = cm.stats()
        sp0 = cm.scan_spec()
        nm = 20
        cm.step(nm)
        msm = cm.last_step_ms() / nm
        s1 = cm.stats()
        sp1 = cm.scan_spec()
        ph = np.zeros(8)
        for _ in range(5):
            cm.step(1)
            tm = cm.timings()
            ph += np.array([tm[k] for k in ("ll_block_ms", "aux_ll_ms", "scan_ms", "update_phi_ms", "split_merge_ms", "pool_ms", "loglik_ms", "total_ms")])
        ph /= 5
        ev = (s1["scan_events"] - s0["scan_events"]) / nm
        line["mixing"] = {"s": a.mixing_s, "K": cm.snapshot(with_phi=False, with_c=False)["K"], "ms_per_sweep": msm, "sweeps_per_s": 1000.0 / msm,
                          "events_per_sweep": ev, "scan_ms": float(ph[2]), "scan_ns_per_observation": 1e6 * ph[2] / a.n,
                          "us_per_event": 1e3 * float(ph[2]) / max(ev, 1.0), "sm_accept_rate": (s1["sm_accepted"] - s0["sm_accepted"]) / max(1, nm),
                          "scan_speculations_per_sweep": (s1["scan_rounds"] - s0["scan_rounds"]) / nm,
                          "scan_speculations_ended_early_per_sweep": (sp1["dropped"] - sp0["dropped"]) / nm,
                          "note": "same shape, Hamming scale s of the generator raised so that observations keep changing cluster at stationarity; "
                                  "the scan evaluates up to 128 undecided observations against one state with the count drift each outcome "
                                  "tolerates and applies the moves in order (DESIGN.md section 3, speculative evaluation; 6.6 us per move without it)"}
        cm.close()
    # ---- extra: the first sweeps from the declared random start, one per call (device time of each)
    if not a.no_random_start and rank == 0:
        cb = Chain(X, attr, gamma, v, w, m=a.m_aux, L=a.k_true, t=a.t, r=a.r, neal8=True, split_merge=True, seed=a.seed + rank,
                   device=local, compact_init=True, data_u8=True)
        firsts, evs, prev = [], [], cb.stats()
        for _ in range(6):
            cb.step(1)
            firsts.append(round(cb.timings()["total_ms"], 3))
            stb = cb.stats()
            evs.append(stb["scan_events"] - prev["scan_events"])
            prev = stb
        cb.close()
        line["burn_in"] = {"first_sweeps_ms": firsts, "moves": evs,
                           "note": "sweeps 1-6 from L random labels (the same chain as the timed one), CUDA-event time of each; "
                                   "round 1 / start of round 2: 480, 288, 95, 24, 11, 4.8 ms"}
    # ---- extra: several independent chains of the metric shape stepped together on this GPU
    if a.multi_chains > 1:
        group = [Chain(X, attr, gamma, v, w, m=a.m_aux, L=a.k_true, t=a.t, r=a.r, neal8=True, split_merge=True,
                       seed=1000 + 17 * rank + q, device=local, compact_init=True, data_u8=True) for q in range(a.multi_chains)]
        step_many(group, a.burn + a.warmup)
        torch.cuda.synchronize()
        nmc = max(20, a.steps // 4)
        tm0 = time.perf_counter()
        step_many(group, nmc)
        torch.cuda.synchronize()
        tq = maxred([time.perf_counter() - tm0])[0]
        line["multi_chain"] = {"chains_per_gpu": a.multi_chains, "value": world * a.multi_chains * nmc / tq,
                               "unit": UNIT, "timing": "wall clock around one smg_step_many call, synchronised both sides",
                               "note": "aggregate over independent chains sharing one GPU; the headline value is one chain per GPU"}
        for g in group:
            g.close()
    # ---- extra: BASELINE config 5 -- chains of n=2e4 sharded over the GPUs, PSM on the tensor cores, reductions by the
    #      C++ library over NCCL (smg_comm_* / smg_chains_*): reduce-scatter of the int32 row blocks, split-R-hat, K histogram
    if a.c5_chains > 0:
        n5 = 20000
        X5, lab5, cent5, attr5 = workload(a, 4242, n=n5)[:4]
        uid = None
        if dist:
            buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                buf = torch.tensor(list(Comm.unique_id()), dtype=torch.uint8, device="cuda")
            dist.broadcast(buf, 0)
            uid = bytes(buf.cpu().numpy().tolist())

        def mk(cid):
            return Chain(X5, attr5, gamma, v, w, m=a.m_aux, L=a.k_true, t=a.t, r=a.r, neal8=True, split_merge=True, seed=7000 + cid,
                         device=local, compact_init=True, data_u8=True, pool_size=n5)
        barrier()
        tw0 = time.time()
        t0 = time.perf_counter()
        out5 = mc.run_chains_native(n5, a.c5_chains, mk, 40, a.c5_kept, rank=rank, world=world, unique_id=uid, device=local,
                                    step_many=step_many, psm_mode=a.c5_psm, kmax=255, psm_capacity=256)
        c5_local = time.perf_counter() - t0
        barrier()
        windows.append((tw0, time.time()))
        c5_s, samp_s, red_ms, burn5_s, kept5_s = maxred([c5_local, out5["sample_seconds"], out5.get("psm_reduce_ms", 0.0),
                                                         out5["burnin_seconds"], out5["kept_seconds"]])
        diag = out5["psm"].read(out5["psm_rows"][0], 1)[0] if out5["psm_rows"][1] > out5["psm_rows"][0] else None
        line["c5"] = {"workload": f"{a.c5_chains} chains, n={n5}, p={a.p}, K_true={a.k_true}, L={a.k_true} random labels: 40 burn-in + {a.c5_kept} kept sweeps per chain, PSM over all kept sweeps",
                      "kept_chain_sweeps_per_s": a.c5_chains * a.c5_kept / kept5_s, "burnin_seconds": burn5_s, "kept_seconds": kept5_s,
                      "sample_seconds": samp_s, "total_seconds": c5_s,
                      "psm_flush_seconds": out5["psm_flush_seconds"], "psm_reduce": out5.get("psm_mode"),
                      "psm_reduce_ms": red_ms, "psm_bus_GBps": out5.get("psm_bus_gbs"), "psm_rows_rank0": list(out5["psm_rows"]),
                      "psm_diag_ok": (None if diag is None else bool(diag[out5["psm_rows"][0]] == a.c5_chains * a.c5_kept)),
                      "rhat_K": out5.get("rhat_K"), "rhat_loglik": out5.get("rhat_loglik"), "n_chains_total": out5.get("n_chains_total"),
                      "K_hist_total": int(out5["K_hist"].sum() + out5["K_hist_overflow"]),
                      "psm_last_flush_ms": out5["psm"].info()["last_flush_ms"],
                      "note": "reductions inside libsmgibbs.so (NCCL bound with dlopen); sampling time excludes them.  psm_reduce "
                              "'fused': the accumulation kernel adds every tile into the memory of the rank that owns its rows "
                              "(CUDA IPC over NVLink), so there is no reduction step (psm_reduce_ms 0); 'reduce_scatter': ncclReduceScatter at the end"}
        for c_ in out5["chains"]:
            c_.close()
        out5["psm"].close()
        out5["comm"].close()
    # ---- extra: posterior similarity matrix on the tensor cores at the C5 shape (n=2e4), rank 0 only
    if rank == 0 and not a.no_psm:
        npsm, T = 20000, 256
        rng = np.random.default_rng(3)
        lab = rng.integers(0, a.k_true, size=(T, npsm)).astype(np.int32)
        P = Psm(npsm, device=local, capacity_sweeps=T)
        best = None
        for _ in range(3):
            for c in lab:
                P.push(c)
            P.flush(finalize=False)
            ms = P.info()["last_flush_ms"]
            best = ms if best is None or ms < best else best
        t0 = time.perf_counter()
        P.flush()  # lower triangle <- upper triangle, once per run
        mirror_ms = 1000.0 * (time.perf_counter() - t0)
        # the matrix is symmetric: n(n+1)/2 distinct entries, each a K-term integer dot product per sweep
        ops_alg = 1.0 * npsm * (npsm + 1) * a.k_true * T
        tiles = sum(1 for by in range(-(-npsm // 128)) for bx in range(-(-npsm // 256)) if bx * 256 + 255 >= by * 128)
        ops_issued = 2.0 * tiles * 128 * 256 * 64 * T  # tiles on or above the diagonal, K padded to 64
        line["psm"] = {"kernel": "psm_accumulate_kernel<64,4>", "bound": "tensor", "n": npsm, "sweeps_per_flush": T,
                       "ms_per_flush": best, "us_per_sweep": 1000.0 * best / T, "mirror_ms_once": mirror_ms,
                       "achieved": ops_alg / best / 1e9, "unit": "TOP/s (u8 dense, algorithmic: n(n+1)/2 entries x K=%d)" % a.k_true,
                       "issued_TOPs": ops_issued / best / 1e9, "full_product_equivalent_TOPs": 2.0 * npsm * npsm * a.k_true * T / best / 1e9,
                       "peak": 2.0 * bf16_peak, "frac": ops_alg / best / 1e9 / (2.0 * bf16_peak),
                       "frac_issued": ops_issued / best / 1e9 / (2.0 * bf16_peak),
                       "peak_source": "2 x bf16_tflops of MEASURED_PEAKS.json (u8 rate = 2 x bf16)",
                       "note": "exact u8 x u8 -> s32 tcgen05.mma over one-hot allocations; only the tiles on or above the diagonal "
                               "are accumulated (symmetric matrix), the lower triangle is copied once at the end"}
        P.close()
    if rank == 0:
        line["clocks"] = sampler.stop(windows) if sampler else None
        if ncores:
            line["host_cores_per_rank"] = ncores
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(a, X, labels, cent, attr, v, w, gamma, 1, "1 chain on 1 host core")
    elif rank == 0:
        line["cpu_baseline"] = None
    if rank == 0:
        print(json.dumps(line))
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
