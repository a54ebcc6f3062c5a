"""Parity of the split-merge step (code/split_merge.cpp:542-598) with the oracle: same pair, same S,
same launch/proposal allocations (bit-exact), same centres, sigma within the bisection bracket, every
MH addend within 1e-10 relative, same accept decision and same post-accept state."""
import numpy as np
import pytest

import oracle_lib as orc
from helpers import Problem, oracle_state_full, rel_err

pytestmark = pytest.mark.gpu

SITE_CENTER_PRIOR, SITE_SIGMA, SITE_CENTER_COND = 1, 2, 5
SITE_SM_PAIR, SITE_SM_LAUNCH_ALLOC, SITE_SM_RGIBBS, SITE_SM_ACCEPT = 6, 7, 8, 9


def build_gpu_tape(log, n, p, t, r, ref, c_before):
    """Re-address the oracle's sequential draws (SURVEY Appendix A order) for the device."""
    ph, site, a, b, u = (log[k] for k in ("phase", "site", "a", "b", "u"))
    ph = ph - ph.min()  # 0 select, 1 split launch, 2 merge launch, 3 proposal, 4 accept
    i1, i2 = ref["i1"], ref["i2"]
    same = ref["is_split"]
    K = int(c_before.max() + 1)
    labA = K if same else int(c_before[i1])
    labB = int(c_before[i2])
    T = {"u_pair": u[(ph == 0) & (site == SITE_SM_PAIR)].copy(),
         "u_prior_c": np.full((3, p), 0.5), "u_prior_s": np.full((3, p), 0.5), "u_launch": np.full(n, 0.5),
         "u_rg": np.full((t + 1, n), 0.5), "u_rg_c": np.full((t + 1, 2, p), 0.5), "u_rg_s": np.full((t + 1, 2, p), 0.5),
         "u_mg_c": np.full((r + 1, p), 0.5), "u_mg_s": np.full((r + 1, p), 0.5),
         "u_accept": u[(ph == 4) & (site == SITE_SM_ACCEPT)].copy()}
    # split launch: prior draws come first (A then B), in order of appearance
    m1 = ph == 1
    pri_c = np.where(m1 & (site == SITE_CENTER_PRIOR))[0]
    assert pri_c.size == 2 * p
    T["u_prior_c"][0] = u[pri_c[:p]]
    T["u_prior_c"][1] = u[pri_c[p:]]
    first_rg = np.where(m1 & (site == SITE_SM_LAUNCH_ALLOC))[0]
    launch_start = first_rg[0] if first_rg.size else np.where(m1)[0][-1] + 1
    pri_s = np.where(m1 & (site == SITE_SIGMA) & (np.arange(u.size) < launch_start))[0]
    assert pri_s.size == 2 * p
    T["u_prior_s"][0] = u[pri_s[:p]]
    T["u_prior_s"][1] = u[pri_s[p:]]
    la = np.where(m1 & (site == SITE_SM_LAUNCH_ALLOC))[0]
    T["u_launch"][b[la]] = u[la]

    def fill_scan(mask, q_of):
        idx = np.where(mask & (site == SITE_SM_RGIBBS))[0]
        for k in idx:
            T["u_rg"][q_of(a[k]), b[k]] = u[k]
        # update_phi draws after the launch allocation: centre (cond) and sigma, tagged by label
        idx = np.where(mask & (site == SITE_CENTER_COND))[0]
        return idx

    # restricted scans of the launch: iteration index = tag a of the RGIBBS draws; the update_phi draws
    # of iteration q follow its RGIBBS draws, so walk sequentially
    q = -1
    for k in np.where(m1)[0]:
        if k < launch_start or site[k] == SITE_SM_LAUNCH_ALLOC:
            continue
        if site[k] == SITE_SM_RGIBBS:
            q = a[k]
            T["u_rg"][q, b[k]] = u[k]
        elif site[k] == SITE_CENTER_COND:
            side = 0 if a[k] == labA else 1
            T["u_rg_c"][max(q, 0), side, b[k]] = u[k]
        elif site[k] == SITE_SIGMA:
            side = 0 if a[k] == labA else 1
            T["u_rg_s"][max(q, 0), side, b[k]] = u[k]
    # t == 0 or empty S: update_phi draws of scan q still happen; handle the no-RGIBBS case by counting
    if ref["S"].size == 0:
        cc = np.where(m1 & (site == SITE_CENTER_COND))[0]
        ss = np.where(m1 & (site == SITE_SIGMA) & (np.arange(u.size) >= launch_start))[0]
        for qq in range(t):
            for side_i in range(2):
                blk = cc[(qq * 2 + side_i) * p:(qq * 2 + side_i + 1) * p]
                lab = a[blk[0]]
                side = 0 if lab == labA else 1
                T["u_rg_c"][qq, side] = u[blk]
                blk = ss[(qq * 2 + side_i) * p:(qq * 2 + side_i + 1) * p]
                T["u_rg_s"][qq, side] = u[blk]
    # merge launch: prior (row 2) then r updates
    m2 = np.where(ph == 2)[0]
    pc = [k for k in m2 if site[k] == SITE_CENTER_PRIOR]
    T["u_prior_c"][2] = u[pc]
    rest = [k for k in m2 if site[k] != SITE_CENTER_PRIOR]
    T["u_prior_s"][2] = u[rest[:p]]
    rest = rest[p:]
    for qq in range(r):
        T["u_mg_c"][qq] = u[rest[qq * 2 * p: qq * 2 * p + p]]
        T["u_mg_s"][qq] = u[rest[qq * 2 * p + p: (qq + 1) * 2 * p]]
    # proposal
    m3 = np.where(ph == 3)[0]
    if same:
        for k in m3:
            if site[k] == SITE_SM_RGIBBS:
                T["u_rg"][t, b[k]] = u[k]
            elif site[k] == SITE_CENTER_COND:
                T["u_rg_c"][t, 0 if a[k] == labA else 1, b[k]] = u[k]
            elif site[k] == SITE_SIGMA:
                T["u_rg_s"][t, 0 if a[k] == labA else 1, b[k]] = u[k]
    else:
        T["u_mg_c"][r] = u[m3[:p]]
        T["u_mg_s"][r] = u[m3[p:2 * p]]
    return T, labA, labB


MODES = ["cluster", "coop", "multi"]


def run_case(pb, state, seed, t=3, r=3, mode="cluster"):
    import os
    # the proposal runs as one thread-block cluster, as one cooperative kernel or as a sequence of launches
    # (same draws, same decisions)
    os.environ["SMG_SM_MODE"] = mode
    os.environ.pop("SMG_SM_PERSISTENT", None)
    K, c, cen, sig = state
    rng = np.random.default_rng(seed)
    tape = (rng.integers(0, 2**53, size=50 + (t + 3) * pb.n + (2 * t + r + 12) * 2 * pb.p).astype(np.float64) + 0.5) / 2.0**53
    o = orc.opts(counted=1, stable_hig=1, sigma_inverse_cdf=1, bisect_tol=0.0)
    ref = orc.split_merge(pb.od, t, r, c, cen, sig, tape, o=o)
    T, labA, labB = build_gpu_tape(ref["log"], pb.n, pb.p, t, r, ref, c)
    ch = pb.chain(t=t, r=r)
    ch.set_state(K, c, cen, sig)
    got = ch.split_merge(T)
    after = ch.snapshot()
    ch.close()
    return ref, got, after, labA, labB


def check_case(pb, ref, got, after, labA, labB):
    assert (got["i1"], got["i2"]) == (ref["i1"], ref["i2"])
    assert np.array_equal(got["S"], ref["S"])
    assert got["is_split"] == ref["is_split"]
    S = ref["S"]
    zl_ref = (ref["SL"]["c"][S] != ref["SL"]["c"][ref["i1"]]).astype(np.int32)
    assert np.array_equal(got["z_launch"], zl_ref)  # launch allocation after t restricted scans: bit-exact
    # launch parameters
    slA, slB = ref["SL"]["c"][ref["i1"]], ref["SL"]["c"][ref["i2"]]
    assert np.array_equal(got["phi"][0, 0], ref["SL"]["center"][slA])
    assert np.array_equal(got["phi"][1, 0], ref["SL"]["center"][slB])
    assert np.max(rel_err(got["phi"][0, 1], ref["SL"]["sigma"][slA])) < 1e-9
    assert np.max(rel_err(got["phi"][1, 1], ref["SL"]["sigma"][slB])) < 1e-9
    mlM = ref["ML"]["c"][ref["i2"]]
    assert np.array_equal(got["phi"][2, 0], ref["ML"]["center"][mlM])
    assert np.max(rel_err(got["phi"][2, 1], ref["ML"]["sigma"][mlM])) < 1e-9
    if ref["is_split"]:
        zs_ref = (ref["star"]["c"][S] != ref["star"]["c"][ref["i1"]]).astype(np.int32)
        assert np.array_equal(got["z_star"], zs_ref)
        stA, stB = ref["star"]["c"][ref["i1"]], ref["star"]["c"][ref["i2"]]
        assert np.array_equal(got["phi"][3, 0], ref["star"]["center"][stA])
        assert np.array_equal(got["phi"][4, 0], ref["star"]["center"][stB])
    else:
        stM = ref["star"]["c"][ref["i2"]]
        assert np.array_equal(got["phi"][5, 0], ref["star"]["center"][stM])
        assert np.max(rel_err(got["phi"][5, 1], ref["star"]["sigma"][stM])) < 1e-9
    # MH addends: 1e-10 relative to the largest addend (sigma enters with ~1e-13 relative differences)
    tr, tg = ref["terms"][:18], got["terms"][:18]
    scale = np.max(np.abs(tr[:14]))
    assert np.max(np.abs(tr - tg)) <= 1e-10 * scale
    assert got["terms"][18] == ref["terms"][18]
    assert got["accepted"] == ref["accepted"]
    assert after["K"] == ref["K"]
    assert np.array_equal(after["c_i"], ref["c"])
    assert np.array_equal(after["centers"], ref["center"])
    assert np.max(rel_err(after["sigmas"], ref["sigma"])) < 1e-9


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("seed", list(range(1, 13)))
def test_split_merge_matches_oracle(seed, mode):
    # over-merged state (K_true=6 collapsed to 3 labels) => splits get accepted; seeds hit both branches
    pb = Problem(600, 24, 4, 6, seed=100 + seed, s=0.6)
    K, c, cen, sig = oracle_state_full(pb, mode="truth", iters=1)
    if seed % 2 == 0:
        c2 = (c % 3).astype(np.int32)
        K2, c2, cen2, sig2 = 3, c2, cen[:3].copy(), sig[:3].copy()
        state = (K2, c2, cen2, sig2)
    else:
        state = (K, c, cen, sig)
    ref, got, after, labA, labB = run_case(pb, state, seed, mode=mode)
    check_case(pb, ref, got, after, labA, labB)


@pytest.mark.parametrize("mode", MODES)
def test_split_merge_accepts_happen_and_match(mode):
    # make sure both an accepted split and an accepted merge are exercised somewhere in the seed range
    acc_split = acc_merge = 0
    for seed in range(20, 60):
        pb = Problem(300, 16, 3, 4, seed=seed, s=0.5)
        K, c, cen, sig = oracle_state_full(pb, mode="truth", iters=1)
        if seed % 2 == 0:  # over-split: cut cluster 0 in two => merges are attractive
            c = c.copy()
            idx = np.where(c == 0)[0]
            c[idx[::2]] = K
            cen = np.vstack([cen, cen[0:1]])
            sig = np.vstack([sig, sig[0:1]])
            K += 1
        else:  # over-merged
            c = np.minimum(c, K - 2).astype(np.int32)
            K -= 1
            cen, sig = cen[:K].copy(), sig[:K].copy()
        ref, got, after, labA, labB = run_case(pb, (K, c, cen, sig), seed, t=2, r=2, mode=mode)
        check_case(pb, ref, got, after, labA, labB)
        if ref["accepted"]:
            if ref["is_split"]:
                acc_split += 1
            else:
                acc_merge += 1
        if acc_split and acc_merge:
            break
    assert acc_split > 0 and acc_merge > 0


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("seed", [3, 4])
def test_split_merge_large_member_set(seed, mode):
    # |S| spans several 1024-member chunks of the restricted-scan decision kernel; low-dimensional, noisy data
    # keeps many members non-robust (count-dependent), so both the parallel and the ordered part are exercised
    pb = Problem(3500 if seed % 2 == 0 else 7000, 12, 3, 2, seed=200 + seed, s=0.9)
    K, c, cen, sig = oracle_state_full(pb, mode="truth", iters=1)
    if seed % 2 == 0:  # everything in one cluster => a split proposal over n-2 members
        c = np.zeros_like(c)
        K, cen, sig = 1, cen[:1].copy(), sig[:1].copy()
    ref, got, after, labA, labB = run_case(pb, (K, c, cen, sig), seed, t=3, r=2, mode=mode)
    assert ref["S"].size > 2048
    check_case(pb, ref, got, after, labA, labB)


def _phi6_from_oracle(ref, p):
    """(centre, sigma) of split-launch A, B, merge-launch M, proposal A*, B*, M* out of the oracle's three states."""
    phi = np.zeros((6, 2, p))
    phi[:, 0, :] = 1.0
    phi[:, 1, :] = 1.0
    i1, i2 = ref["i1"], ref["i2"]
    SL, ML, st = ref["SL"], ref["ML"], ref["star"]
    for q, (stt, lab) in enumerate([(SL, SL["c"][i1]), (SL, SL["c"][i2]), (ML, ML["c"][i2])]):
        phi[q, 0], phi[q, 1] = stt["center"][lab], stt["sigma"][lab]
    if ref["is_split"]:
        phi[3, 0], phi[3, 1] = st["center"][st["c"][i1]], st["sigma"][st["c"][i1]]
        phi[4, 0], phi[4, 1] = st["center"][st["c"][i2]], st["sigma"][st["c"][i2]]
    else:
        phi[5, 0], phi[5, 1] = st["center"][st["c"][i2]], st["sigma"][st["c"][i2]]
    return phi


@pytest.mark.parametrize("seed", list(range(1, 13)))
def test_mh_terms_alone_at_1e12(seed):
    """a19-a24 in isolation (split_merge.cpp:393-540): the oracle's launch / proposal states (sides, centres, sigmas) are
    injected, so no device draw feeds the addends; every addend within 1e-12 relative, the decision identical."""
    pb = Problem(600, 24, 4, 6, seed=100 + seed, s=0.6)
    K, c, cen, sig = oracle_state_full(pb, mode="truth", iters=1)
    if seed % 2 == 0:
        c = (c % 3).astype(np.int32)
        K, cen, sig = 3, cen[:3].copy(), sig[:3].copy()
    t = r = 3
    rng = np.random.default_rng(seed)
    tape = (rng.integers(0, 2**53, size=50 + (t + 3) * pb.n + (2 * t + r + 12) * 2 * pb.p).astype(np.float64) + 0.5) / 2.0**53
    o = orc.opts(counted=1, stable_hig=1, sigma_inverse_cdf=1, bisect_tol=0.0)
    ref = orc.split_merge(pb.od, t, r, c, cen, sig, tape, o=o)
    T, _, _ = build_gpu_tape(ref["log"], pb.n, pb.p, t, r, ref, c)
    S = ref["S"]
    zl = (ref["SL"]["c"][S] != ref["SL"]["c"][ref["i1"]]).astype(np.int32)
    zs = (ref["star"]["c"][S] != ref["star"]["c"][ref["i1"]]).astype(np.int32) if ref["is_split"] else np.zeros(S.size, np.int32)
    ch = pb.chain(t=t, r=r)
    ch.set_state(K, c, cen, sig)
    got = ch.sm_terms(T["u_pair"], zl, zs, _phi6_from_oracle(ref, pb.p), float(T["u_accept"][0]))
    after = ch.snapshot()
    ch.close()
    assert (got["i1"], got["i2"], got["nS"], got["is_split"]) == (ref["i1"], ref["i2"], S.size, ref["is_split"])
    tr, tg = ref["terms"], got["terms"]
    names = ["log_alpha", "lg0", "lg1", "lg2", "pri0", "pri1", "pri2", "ll0", "ll1", "ll2", "gs_phi0", "gs_phi1", "gs_phi2",
             "gs_c", "log_prior", "log_lik", "log_prop"]
    for k, nm in enumerate(names):
        # the prior / parameter-density terms are sums of p per-attribute log-densities of order one that partly cancel:
        # 1e-12 of the term or of p (the size of the sum of their magnitudes), whichever is larger
        scale = max(abs(tr[k]), float(pb.p) if nm.startswith(("pri", "gs_phi")) or nm in ("log_prior", "log_prop") else 1.0)
        assert abs(tg[k] - tr[k]) <= 1e-12 * scale, (nm, tg[k], tr[k])
    # the ratio is a difference of the addends above: 1e-12 of the largest of them
    assert abs(tg[17] - tr[17]) <= 1e-12 * np.max(np.abs(tr[:17]))
    assert tg[18] == tr[18]
    assert got["accepted"] == ref["accepted"]
    assert after["K"] == K and np.array_equal(after["c_i"], c)  # nothing applied


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("seed,iteration", [(1, 0), (2, 7), (3, 601), (4, 1234)])
def test_deterministic_pair_selection_matches_oracle(seed, iteration, mode):
    """select_observations_deterministic (split_merge.cpp:227-261): i_1 walks over the observations (iteration mod n),
    i_2 = (int)(n u) redrawn while equal to i_1; the rest of the proposal as usual."""
    import os
    os.environ["SMG_SM_MODE"] = mode
    pb = Problem(600, 24, 4, 6, seed=300 + seed, s=0.6)
    K, c, cen, sig = oracle_state_full(pb, mode="truth", iters=1)
    if seed % 2 == 0:
        c = (c % 3).astype(np.int32)
        K, cen, sig = 3, cen[:3].copy(), sig[:3].copy()
    t = r = 2
    i1 = iteration % pb.n
    rng = np.random.default_rng(seed)
    tape = (rng.integers(0, 2**53, size=50 + (t + 3) * pb.n + (2 * t + r + 12) * 2 * pb.p).astype(np.float64) + 0.5) / 2.0**53
    if seed == 3:
        tape[0] = (i1 + 0.5) / pb.n  # the first draw of i_2 hits i_1: the reference redraws
    o = orc.opts(counted=1, stable_hig=1, sigma_inverse_cdf=1, bisect_tol=0.0, det_i1=i1)
    ref = orc.split_merge(pb.od, t, r, c, cen, sig, tape, o=o)
    assert ref["i1"] == i1 and ref["i2"] != i1
    T, labA, labB = build_gpu_tape(ref["log"], pb.n, pb.p, t, r, ref, c)
    assert T["u_pair"].size == (2 if seed == 3 else 1)
    T["u_pair"] = np.concatenate([T["u_pair"], [0.5]])[:2]
    ch = pb.chain(t=t, r=r, pair_selection="deterministic")
    ch.set_state(K, c, cen, sig)
    lb_check = ch.lib.smg_resume_at(ch.h, int(iteration))
    assert lb_check == 0
    got = ch.split_merge(T)
    after = ch.snapshot()
    ch.close()
    check_case(pb, ref, got, after, labA, labB)


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_split_merge_long_launch_on_separated_clusters(seed, mode):
    """t = r = 8 on well separated clusters: after two or three restricted scans nobody moves any more, which is the regime
    of the metric configuration -- the cluster kernel then decides every member from its two mismatch counts and skips
    the table, the ordered pass, the histogram moves and their reduction.  Same launch / proposal states as the oracle."""
    pb = Problem(2400, 64, 4, 6, seed=400 + seed, s=0.5)
    K, c, cen, sig = oracle_state_full(pb, mode="truth", iters=1)
    if seed % 2 == 0:  # over-merged: a split proposal over two true clusters
        c = (c % 3).astype(np.int32)
        K, cen, sig = 3, cen[:3].copy(), sig[:3].copy()
    ref, got, after, labA, labB = run_case(pb, (K, c, cen, sig), seed, t=8, r=8, mode=mode)
    check_case(pb, ref, got, after, labA, labB)
