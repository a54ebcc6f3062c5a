"""The allocation scan's speculative evaluation (neal8.cpp:105-159 semantics unchanged): many undecided observations are
evaluated against one state together with the drift of the log-counts their outcome tolerates, and the moves are then
applied in order.  The result has to be the one-at-a-time scan's, bit for bit: against the oracle under an injected
tape, against the plain scan (mode 0) on the device's own streams, and -- mode 2 -- every speculated outcome is
re-evaluated exactly under the state the observation meets and the differences are counted (must be 0)."""
import numpy as np
import pytest

import oracle_lib as orc
from helpers import Problem, oracle_state_full

pytestmark = pytest.mark.gpu


def _tape(n, m_aux, seed):
    rng = np.random.default_rng(seed)
    return (rng.integers(0, 2**53, size=n * (m_aux + 1)).astype(np.float64) + 0.5) / 2.0**53


@pytest.mark.parametrize("seed,start,s", [(1, "truth", 1.5), (2, "truth", 1.7), (3, "random", 1.5), (4, "random", 1.2)])
def test_speculative_scan_matches_oracle_and_plain_scan(seed, start, s):
    m_aux = 3
    pb = Problem(6000, 48, 4, 12, seed=seed, s=s)
    K, c, cen, sig = oracle_state_full(pb, mode=start, seed=seed, L=16, iters=1, m_aux=m_aux)
    pc, ps = orc.draw_pool(pb.od, 193, seed + 1, o=orc.opts(stable_hig=1))
    tape = _tape(pb.n, m_aux, seed)
    ref = orc.neal8_scan(pb.od, m_aux, c, cen, sig, pc, ps, tape, o=orc.opts(counted=1))
    events = {}
    for mode in (0, 1, 2):
        ch = pb.chain(m=m_aux)
        ch.scan_spec(mode)
        ch.set_state(K, c, cen, sig)
        ch.set_pool(pc, ps)
        ch.neal8_scan(tape)
        got = ch.snapshot()
        events[mode] = ch.stats()["scan_events"]
        sp = ch.scan_spec(mode)
        ch.close()
        assert got["K"] == ref["K"], mode
        assert np.array_equal(got["c_i"], ref["c"]), mode
        assert np.array_equal(got["centers"], ref["center"]) and np.array_equal(got["sigmas"], ref["sigma"]), mode
        if mode == 2:
            assert sp["mismatches"] == 0
    assert events[0] == events[1] == events[2] and events[0] > 300  # the draws are real decisions


def test_speculative_chain_equals_plain_chain_while_mixing():
    # a chain that keeps moving (diffuse data): thousands of moves per pass, the device's own Philox streams
    pb = Problem(20000, 64, 5, 20, seed=9, s=1.6)
    runs = {}
    for mode in (0, 2, 1):
        ch = pb.chain(m=3, L=20, c_i=None, t=2, r=2, neal8=True, split_merge=True, seed=77, compact_init=True)
        ch.scan_spec(mode)
        trace = []
        for _ in range(6):
            ch.step(1)
            s = ch.snapshot(with_phi=False)
            trace.append((s["K"], s["loglikelihood"], s["c_i"].copy()))
        st = ch.stats()
        sp = ch.scan_spec(mode)
        ch.close()
        runs[mode] = (trace, st, sp)
    assert runs[2][2]["mismatches"] == 0
    assert runs[2][2]["reevaluated"] > 0  # tolerances were used up and rows re-evaluated: the fallback is exercised
    assert runs[0][1]["scan_events"] > 5000
    for mode in (1, 2):
        assert runs[mode][1]["scan_events"] == runs[0][1]["scan_events"]
        for a, b in zip(runs[0][0], runs[mode][0]):
            assert a[0] == b[0] and a[1] == b[1] and np.array_equal(a[2], b[2])
    assert runs[1][1]["scan_rounds"] < runs[0][1]["scan_rounds"]  # the point of it: far fewer rounds of the whole block


@pytest.mark.parametrize("seed,k_true,m_aux,gamma,s,L", [(11, 58, 5, 1.0, 1.4, 58), (12, 10, 3, 60.0, 1.3, 14), (13, 30, 4, 5.0, 1.6, 40)])
def test_speculative_scan_near_capacity_and_with_births(seed, k_true, m_aux, gamma, s, L):
    # 63 of the 64 entries of the two-per-lane evaluation in use; a large concentration parameter (clusters are born and
    # closed inside the pass, so the walk hands rows back to the block and the born columns are filled between
    # speculations); pool-free auxiliary components are covered by test_gpu_modes
    pb = Problem(4000, 32, 4, k_true, seed=seed, s=s, gamma=gamma)
    K, c, cen, sig = oracle_state_full(pb, mode="random", seed=seed, L=L, iters=1, m_aux=m_aux)
    pc, ps = orc.draw_pool(pb.od, 257, seed + 1, o=orc.opts(stable_hig=1))
    tape = _tape(pb.n, m_aux, seed)
    ref = orc.neal8_scan(pb.od, m_aux, c, cen, sig, pc, ps, tape, o=orc.opts(counted=1), kcap=512)
    for mode in (1, 2):
        ch = pb.chain(m=m_aux, max_clusters=200)
        ch.scan_spec(mode)
        ch.set_state(K, c, cen, sig)
        ch.set_pool(pc, ps)
        ch.neal8_scan(tape)
        got = ch.snapshot()
        st = ch.stats()
        sp = ch.scan_spec(mode)
        ch.close()
        assert got["K"] == ref["K"] and np.array_equal(got["c_i"], ref["c"]), mode
        assert np.array_equal(got["centers"], ref["center"]) and np.array_equal(got["sigmas"], ref["sigma"]), mode
        assert sp["mismatches"] == 0
        assert st["scan_events"] > 200
    if gamma > 50.0:
        assert st["births"] > 0
