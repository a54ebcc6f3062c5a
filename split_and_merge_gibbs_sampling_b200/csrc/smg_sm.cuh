// smg_sm.cuh -- the Jain-Neal split-merge proposal (code/split_merge.cpp:542-598) on device.
//
// The whole proposal is a fixed sequence of small kernels driven by flags in device memory (no
// host round trip): pair selection + member list S (:263-301), split launch state (:303-352) with
// t restricted Gibbs scans (:163-225), merge launch state (:354-391) with r parameter updates,
// the proposal, the MH log-ratio (:438-540) and accept + relabel (clean_var, common_functions.cpp
// :296-353).  Launch states are kept in compact form: a 0/1 side per member of S (0 = with i_1,
// 1 = with i_2) and a few extra parameter slots; nothing of size n is cloned.
#pragma once
#include "smg_chain.cuh"

namespace smg {

// histogram / count indices of the split-merge workspace
enum SmHist : int { SH_L0 = 0, SH_L1 = 1, SH_M = 2, SH_S0 = 3, SH_S1 = 4, SH_P0 = 5, SH_P1 = 6, SH_N = 7 };

struct SmInfo {
  int i1, i2, nS, same, cA, cB, K, pad;
};

struct SmPlan {  // which slots / histograms feed each MH term (filled on device once the pair is known)
  int gs_hist[3], gs_sigL[3], gs_star[3];
  int pri_slot[3];
  int lg_cnt[3];
  int slotA, slotB, slotAll;  // per-row likelihood evaluations
  int zsel;                   // 0: proposal sides (split), 1: current-state sides (merge)
};

struct SmcHost;  // cluster-kernel resources (smg_smc.cuh)

struct SmWork {
  SmcHost* smc = nullptr;
  int forced_mode = 0;  // SMG_SM_MODE at creation: 0 default, 1 cluster, 2 coop, 3 multi
  int *S = nullptr, *zL = nullptr, *zStar = nullptr, *zState = nullptr;
  SmInfo* info = nullptr;
  SmPlan* plan = nullptr;
  int* anchors = nullptr;  // alias of info->i1,i2 (two ints)
  int *H = nullptr, *cnt = nullptr;
  double *rg_dl = nullptr, *rg_lgt = nullptr, *rg_lgt2 = nullptr;  // [n] LL_A - LL_B and logit(u) of the restricted scans
  double* rowvals = nullptr;  // [4][n+2]
  double* partial = nullptr;  // [4][RB]
  double* terms = nullptr;    // [24]
  size_t hist_smem = 0;       // dynamic shared memory of subset_histogram_smem_kernel (0: does not fit)
  unsigned* chain_bar = nullptr;  // [4]: two grid-barrier counters (used in turn) and the error flag of sm_chain_kernel
  int bar_flip = 0;
  int* selcnt = nullptr;          // [grid] members found per CTA in the selection phase
  bool persistent = false;        // restricted-scan chain as one cooperative kernel
  size_t chain_smem = 0;          // its dynamic shared memory: max(side histograms, decision scratch)
  // injected uniforms (device copies, allocated on first use)
  double *u_pair = nullptr, *u_prior_c = nullptr, *u_prior_s = nullptr, *u_launch = nullptr, *u_rg = nullptr;
  double *u_rg_c = nullptr, *u_rg_s = nullptr, *u_mg_c = nullptr, *u_mg_s = nullptr, *u_accept = nullptr;
  bool inj_alloc = false;
  cudaEvent_t tr[2] = {nullptr, nullptr};  // SMG_SM_TRACE: device time of the proposal kernel
};

#define SM_RB 256  // partial-sum blocks of the row reductions

// sample(indices, 2, replace=false) (split_merge.cpp:275): j1=(int)(n*u0); j2=(int)((n-1)*u1) over the swapped array
// det != 0: select_observations_deterministic (split_merge.cpp:227-261) -- i_1 walks over the observations (iteration
// mod n), i_2 = sample(indexes, 1, false)[0] = (int)(n u), redrawn while it equals i_1 (injected uniforms: the first two)
__device__ __forceinline__ void sm_pick_pair(int n, const double* u_pair, const RngKey& key, int* i1, int* i2, int det = 0) {
  if (det) {
    const int a = (int)(key.sweep % (uint32_t)n);
    int b = a;
    for (uint32_t k = 0; k < 4096u && b == a; k++) {
      const double u = get_u(k < 2u ? u_pair : nullptr, k, key, U_SM_PAIR, k, 0u);
      b = (int)((double)n * u);
      if (b >= n) b = n - 1;
    }
    *i1 = a;
    *i2 = b;
    return;
  }
  const double u0 = get_u(u_pair, 0, key, U_SM_PAIR, 0u, 0u), u1 = get_u(u_pair, 1, key, U_SM_PAIR, 1u, 0u);
  int j1 = (int)((double)n * u0);
  if (j1 >= n) j1 = n - 1;
  int j2 = (int)((double)(n - 1) * u1);
  if (j2 >= n - 1) j2 = n - 2;
  *i1 = j1;
  *i2 = (j2 == j1) ? (n - 1) : j2;
}
// proposal descriptor + which slots / histograms feed each MH term
__device__ __forceinline__ void sm_fill_info_plan(SmInfo* info, SmPlan* plan, int NS, int i1, int i2, int nS, int cA, int cB,
                                                  int K) {
  const int same = (cA == cB);
  {
    info->i1 = i1;
    info->i2 = i2;
    info->nS = nS;
    info->same = same;
    info->cA = cA;
    info->cB = cB;
    info->K = K;
    SmPlan P;
    const int B = NS;
    if (same) {  // split_acc_prob (split_merge.cpp:438-487)
      P.gs_hist[0] = SH_P0, P.gs_sigL[0] = B + SM_SL_A, P.gs_star[0] = B + SM_ST_A;
      P.gs_hist[1] = SH_P1, P.gs_sigL[1] = B + SM_SL_B, P.gs_star[1] = B + SM_ST_B;
      P.gs_hist[2] = SH_M, P.gs_sigL[2] = B + SM_ML_M, P.gs_star[2] = cA;
      P.pri_slot[0] = B + SM_ST_A, P.pri_slot[1] = B + SM_ST_B, P.pri_slot[2] = cA;
      P.lg_cnt[0] = SH_P0, P.lg_cnt[1] = SH_P1, P.lg_cnt[2] = SH_M;
      P.slotA = B + SM_ST_A, P.slotB = B + SM_ST_B, P.slotAll = cA;
      P.zsel = 0;
    } else {  // merge_acc_prob (split_merge.cpp:489-540)
      P.gs_hist[0] = SH_S0, P.gs_sigL[0] = B + SM_SL_A, P.gs_star[0] = cA;
      P.gs_hist[1] = SH_S1, P.gs_sigL[1] = B + SM_SL_B, P.gs_star[1] = cB;
      P.gs_hist[2] = SH_M, P.gs_sigL[2] = B + SM_ML_M, P.gs_star[2] = B + SM_ST_M;
      P.pri_slot[0] = cA, P.pri_slot[1] = cB, P.pri_slot[2] = B + SM_ST_M;
      P.lg_cnt[0] = SH_S0, P.lg_cnt[1] = SH_S1, P.lg_cnt[2] = SH_M;
      P.slotA = cA, P.slotB = cB, P.slotAll = B + SM_ST_M;
      P.zsel = 1;
    }
    *plan = P;
  }
}

// ------------------------------------------------------------------------------------------
// pair selection + S (split_merge.cpp:263-301).  One CTA; ordered stream compaction.
// sample(indices, 2, replace=false): j1=(int)(n*u0); j2=(int)((n-1)*u1) over the swapped array.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) sm_select_kernel(int n, const int* __restrict__ c, const int* __restrict__ Kptr,
                                                         const double* u_pair, RngKey key, int NS, int* __restrict__ S,
                                                         int* __restrict__ zState, SmInfo* info, SmPlan* plan,
                                                         int* __restrict__ cnt, double* terms, int pair_det) {
  __shared__ int s_i1, s_i2, s_cA, s_cB, s_tot;
  __shared__ int s_wcnt[32], s_woff[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    int i1, i2;
    sm_pick_pair(n, u_pair, key, &i1, &i2, pair_det);
    s_i1 = i1;
    s_i2 = i2;
    s_cA = c[i1];
    s_cB = c[i2];
  }
  for (int q = tid; q < 24; q += blockDim.x) terms[q] = 0.0;
  __syncthreads();
  const int i1 = s_i1, i2 = s_i2, cA = s_cA, cB = s_cB;
  // ordered compaction: warp w owns the contiguous rows [lo, hi); count, scan the 32 totals, write
  const int seg = ((n + 31) / 32 + 31) & ~31;
  const int lo = min(n, warp * seg), hi = min(n, lo + seg);
  int mine = 0;
  for (int i = lo + lane; i < hi; i += 32) {
    const int ci = c[i];
    mine += (i != i1 && i != i2 && (ci == cA || ci == cB));
  }
  mine = warp_sum_i(mine);
  if (lane == 0) s_wcnt[warp] = mine;
  __syncthreads();
  if (warp == 0) {
    int v = s_wcnt[lane], x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(SMG_FULL, x, o);
      if (lane >= o) x += y;
    }
    s_woff[lane] = x - v;  // exclusive
    if (lane == 31) s_tot = x;
  }
  __syncthreads();
  int off = s_woff[warp];
  for (int base = lo; base < hi; base += 32) {
    const int i = base + lane;
    const int ci = (i < hi) ? c[i] : -1;
    const bool in = (i < hi) && i != i1 && i != i2 && (ci == cA || ci == cB);
    const unsigned b = __ballot_sync(SMG_FULL, in);
    if (in) {
      const int pos = off + __popc(b & ((1u << lane) - 1));
      S[pos] = i;
      zState[pos] = (ci == cA) ? 0 : 1;
    }
    off += __popc(b);
  }
  if (tid == 0) sm_fill_info_plan(info, plan, NS, i1, i2, s_tot, cA, cB, *Kptr);
}

// random launch allocation (split_merge.cpp:346): sample({a,b}, |S|, replace) -> (int)(2u)
__global__ void sm_launch_alloc_kernel(const SmInfo* info, const double* u_inj, RngKey key, int* __restrict__ zL) {
  int pos = blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= info->nS) return;
  double u = get_u(u_inj, pos, key, U_SM_LAUNCH, (uint32_t)pos, 0u);
  int z = (int)(2.0 * u);
  zL[pos] = z > 1 ? 1 : z;
}

// Restricted Gibbs allocation scan over S (split_merge.cpp:186-216), in two kernels.
//
// The two-way Rcpp::sample (split_merge.cpp:215) walks the two probabilities in descending order
// (revsort; on a tie the SECOND entry comes first) and takes the first when u <= p_first.  With
// D = (log n_1 + LL_1) - (log n_2 + LL_2) the larger probability is 1/(1+exp(-|D|)), so the member goes
// to the larger side iff |D| >= logit(u), to the other side otherwise.  logit(u) and d0 = LL_1 - LL_2 do
// not depend on the running counts: sm_ll2prep_kernel evaluates them for every member in parallel
// (one warp per member, all SMs).
// logit of the allocation uniform of member `pos` in the scan keyed by `key` (it does not depend on the state: the
// persistent kernel evaluates it one scan ahead on CTAs that would otherwise wait; lgt == nullptr below)
__device__ __forceinline__ double sm_logit_u(const double* u_inj, int pos, const RngKey& key) {
  const double u = get_u(u_inj, pos, key, U_SM_RGIBBS, (uint32_t)pos, 0u);
  return log(u / (1.0 - u));
}
// Interval of D = (log n_1 + LL_1) - (log n_2 + LL_2) a value falls in, for logit(u) = lg; even intervals decide side 1,
// odd ones side 0 (the rule of the ordered walk below: the larger side iff |D| >= logit u, ties in D to side 1).
__device__ __forceinline__ int sm_d_region(double D, double lg) {
  if (lg <= 0.0) return D > 0.0 ? 3 : 0;
  return D >= lg ? 3 : (D > 0.0 ? 2 : (D > -lg ? 1 : 0));
}
// log((nS+1-b)/b) in single precision (absolute error below 1e-5 for nS < 2^31); callers widen it by SM_DC_MARGIN
#define SM_DC_MARGIN 1e-4
__device__ __forceinline__ double sm_dc_bound(int nS, int b) {
  return (double)(logf((float)(nS + 1 - b)) - logf((float)b));
}
__device__ __forceinline__ void sm_ll2prep_body(const uint8_t* __restrict__ X, int pp, const int* __restrict__ S, int nS,
                                                const uint8_t* cen, const double* isg, const double* sden, int slotA,
                                                int slotB, const double* u_inj, const RngKey& key, double* dl,
                                                double* lgt, int gwarp, int nwarps, const double* lgq = nullptr,
                                                const int* zcur = nullptr, int* exc = nullptr) {
  const int lane = threadIdx.x & 31;
  // exc != nullptr: also count the members that are NOT certain to keep their side whatever the running counts are
  // (decision region constant over the whole range of the count term, and equal to the current side).  When that count
  // is zero the serial scan changes nothing: the caller skips the decision and the histogram phases.
  const double dcx = nS > 0 ? sm_dc_bound(nS, 1) + SM_DC_MARGIN : 0.0, dcn = nS > 0 ? sm_dc_bound(nS, nS) - SM_DC_MARGIN : 0.0;
  int nexc = 0;
  auto settled = [&](int pos, double d0) {
    const double lg = lgq[pos];
    const int rlo = sm_d_region(dcn + d0, lg), rhi = sm_d_region(dcx + d0, lg);
    return rlo == rhi && (~rlo & 1) == zcur[pos];
  };
  const uint8_t *cA = cen + (size_t)slotA * pp, *cB = cen + (size_t)slotB * pp;
  const double *wA = isg + (size_t)slotA * pp, *wB = isg + (size_t)slotB * pp;
  const double sdA = sden[slotA], sdB = sden[slotB];
  if (pp <= 256) {
    // One 8-attribute slice per lane: the two parameter vectors go to registers once, and two members are in flight
    // per warp (their index and data loads are issued together) -- the phase is a chain of memory latencies.
    // Same per-lane order of additions and the same butterfly as warp_mismatch_dot.
    const int j0 = lane * 8;
    const bool act = j0 < pp;
    uint2 ca = make_uint2(0u, 0u), cb = make_uint2(0u, 0u);
    double wa[8], wb[8];
#pragma unroll
    for (int b = 0; b < 8; b++) wa[b] = wb[b] = 0.0;
    if (act) {
      ca = *reinterpret_cast<const uint2*>(cA + j0);
      cb = *reinterpret_cast<const uint2*>(cB + j0);
#pragma unroll
      for (int b = 0; b < 4; b++) {
        const double2 ta = reinterpret_cast<const double2*>(wA + j0)[b], tb = reinterpret_cast<const double2*>(wB + j0)[b];
        wa[2 * b] = ta.x, wa[2 * b + 1] = ta.y;
        wb[2 * b] = tb.x, wb[2 * b + 1] = tb.y;
      }
    }
    auto dot = [&](const uint2 xv, const uint2 cv, const double* w) {
      const uint32_t m0 = __vcmpne4(xv.x, cv.x), m1 = __vcmpne4(xv.y, cv.y);
      double acc = 0.0;
#pragma unroll
      for (int b = 0; b < 4; b++) {
        if (m0 & (0xffu << (8 * b))) acc += w[b];
      }
#pragma unroll
      for (int b = 0; b < 4; b++)
        if (m1 & (0xffu << (8 * b))) acc += w[4 + b];
      return acc;
    };
    for (int pos0 = gwarp; pos0 < nS; pos0 += 2 * nwarps) {
      const int pos1 = pos0 + nwarps;
      const bool has1 = pos1 < nS;
      const int r0 = S[pos0], r1 = has1 ? S[pos1] : r0;
      uint2 x0 = ca, x1 = ca;
      if (act) {
        x0 = *reinterpret_cast<const uint2*>(X + (size_t)r0 * pp + j0);
        x1 = *reinterpret_cast<const uint2*>(X + (size_t)r1 * pp + j0);
      }
      double a0 = act ? dot(x0, ca, wa) : 0.0, b0 = act ? dot(x0, cb, wb) : 0.0;
      double a1 = act ? dot(x1, ca, wa) : 0.0, b1 = act ? dot(x1, cb, wb) : 0.0;
      a0 = warp_sum(a0);
      b0 = warp_sum(b0);
      a1 = warp_sum(a1);
      b1 = warp_sum(b1);
      if (lane == 0) {
        const double d0 = (-a0 - sdA) - (-b0 - sdB);
        dl[pos0] = d0;
        if (lgt) lgt[pos0] = sm_logit_u(u_inj, pos0, key);
        if (exc && !settled(pos0, d0)) nexc++;
      }
      if (lane == 1 && has1) {
        const double d1 = (-a1 - sdA) - (-b1 - sdB);
        dl[pos1] = d1;
        if (lgt) lgt[pos1] = sm_logit_u(u_inj, pos1, key);
        if (exc && !settled(pos1, d1)) nexc++;
      }
    }
    if (exc && nexc) atomicAdd(exc, nexc);
    return;
  }
  for (int pos = gwarp; pos < nS; pos += nwarps) {
    const uint8_t* x = X + (size_t)S[pos] * pp;
    const double llA = -warp_mismatch_dot(x, cA, wA, pp, lane) - sdA;
    const double llB = -warp_mismatch_dot(x, cB, wB, pp, lane) - sdB;
    if (lane == 0) {
      dl[pos] = llA - llB;
      if (lgt) lgt[pos] = sm_logit_u(u_inj, pos, key);
      if (exc && !settled(pos, llA - llB)) nexc++;
    }
  }
  if (exc && nexc) atomicAdd(exc, nexc);
}

__global__ void __launch_bounds__(256) sm_ll2prep_kernel(const uint8_t* __restrict__ X, int pp, const int* __restrict__ S,
                                                         const SmInfo* info, const uint8_t* cen, const double* isg,
                                                         const double* sden, int slotA, int slotB, const double* u_inj,
                                                         RngKey key, double* __restrict__ dl, double* __restrict__ lgt,
                                                         const int* enable, int enable_val) {
  if (enable && *enable != enable_val) return;
  const int wpb = blockDim.x >> 5;
  sm_ll2prep_body(X, pp, S, info->nS, cen, isg, sden, slotA, slotB, u_inj, key, dl, lgt,
                  blockIdx.x * wpb + (threadIdx.x >> 5), gridDim.x * wpb);
}

// The sequential part, one CTA.  Inside a chunk of 1024 members the running side counts can only move by
// the chunk's own population, which bounds the log-count term of D; a member whose D stays on one side of
// every decision threshold over that whole interval lands on the same side no matter what happened before
// it ("robust").  Per chunk: robust members are decided in
// parallel, a block prefix sum gives the count contribution of the robust members before every
// position, and only the non-robust members are walked in order by one warp (32 at a time against the
// same counts; the first one that changes side is applied and evaluation restarts after it).  Same
// result as the one-at-a-time scan.  Also zeroes the histograms the next kernel fills and publishes
// the side counts (anchors included).
#define SM_DECIDE_T 1024
// Members per thread and chunk.  Wider chunks amortise the per-chunk barriers but loosen the count bounds, and in the
// first scans after the random launch allocation (small likelihood differences) that leaves many more members to the
// ordered walk: 2048-member chunks measured 0.84 ms per proposal against 0.72 ms with one member per thread.
#define SM_DECIDE_R 1
// ... so only the scans after the first two (sides settled, large likelihood differences) use wide chunks:
#define SM_DECIDE_WIDE_CHUNK 2048
#define SM_DECIDE_WIDE_FROM 2
static inline int sm_wide_from() {
  static const int v = [] {
    const char* e = getenv("SMG_SM_WIDE_FROM");
    return e ? atoi(e) : SM_DECIDE_WIDE_FROM;
  }();
  return v;
}
template <int N>
struct RdecideSmem {  // per-member scratch of one chunk (only the non-robust members are read back)
  double d0[N], lg[N];
  int z[N], pre[N], list[N];
  int wsum[32], wnr[32];
  int nB, nnr, cdelta;
};

// (no __restrict__ here: inside sm_chain_kernel these arrays are written by other CTAs earlier in the same launch)
// T threads, R consecutive members per thread.
template <int T, int R>
__device__ __forceinline__ void sm_rdecide_body(int nS, const double* dl, const double* lgt, int* z, int* Hzero, int hlen,
                                                int* cnt2, RdecideSmem<T * R>& M) {
  constexpr int N = T * R;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // this thread's members of the next two chunks (loads run two chunks ahead of the decisions)
  double nx_d0[R], nx_lg[R], n2_d0[R], n2_lg[R];
  int nx_z[R], n2_z[R];
#pragma unroll
  for (int k = 0; k < R; k++) {
    const int pos = tid * R + k, pos2 = N + pos;
    nx_d0[k] = pos < nS ? dl[pos] : 0.0;
    nx_lg[k] = pos < nS ? lgt[pos] : 0.0;
    nx_z[k] = pos < nS ? z[pos] : 0;
    n2_d0[k] = pos2 < nS ? dl[pos2] : 0.0;
    n2_lg[k] = pos2 < nS ? lgt[pos2] : 0.0;
    n2_z[k] = pos2 < nS ? z[pos2] : 0;
  }
  for (int q = tid; q < hlen; q += T) Hzero[q] = 0;
  {  // side-1 count on entry (anchor i_2 included)
    int c1 = 0;
    for (int pos = tid; pos < nS; pos += T) c1 += z[pos];
    c1 = warp_sum_i(c1);
    if (lane == 0) M.wsum[warp] = c1;
    __syncthreads();
    if (warp == 0) {
      int t = warp_sum_i(lane < T / 32 ? M.wsum[lane] : 0);
      if (lane == 0) M.nB = 1 + t;
    }
    __syncthreads();
  }
  int nBcur = M.nB;  // side-1 count (anchor included) on entry to the current chunk, kept by every thread
  for (int base = 0; base < nS; base += N) {
    double d0[R], lg[R];
    int zz[R];
#pragma unroll
    for (int k = 0; k < R; k++) {
      d0[k] = nx_d0[k];
      lg[k] = nx_lg[k];
      zz[k] = nx_z[k];
      nx_d0[k] = n2_d0[k];
      nx_lg[k] = n2_lg[k];
      nx_z[k] = n2_z[k];
      const int pn = base + 2 * N + tid * R + k;
      if (pn < nS) {
        n2_d0[k] = dl[pn];
        n2_lg[k] = lgt[pn];
        n2_z[k] = z[pn];
      }
    }
    // b = side-1 members other than the one being decided (anchor included) stays inside [blo, bhi] while this
    // chunk is walked: at most min(chunk size, side-1 members) leave and min(chunk size, side-0 members) join.
    // The log-count term log((nS+1-b)/b) of D is decreasing in b.
    const int nv = min(N, nS - base);
    const int blo = nBcur - min(nv, nBcur - 1), bhi = nBcur + min(nv, nS + 1 - nBcur) - 1;
    // (single-precision logs widened by a margin far above their error: these are bounds, and every thread needs them)
    const double dc_max = sm_dc_bound(nS, blo) + SM_DC_MARGIN;
    const double dc_min = sm_dc_bound(nS, bhi) - SM_DC_MARGIN;
    int newz[R], pre[R];
    unsigned nrmask = 0;  // bit k: member k of this thread needs the ordered walk
    int run = 0;          // side-1 change of this thread's robust members so far
#pragma unroll
    for (int k = 0; k < R; k++) {
      const bool valid = base + tid * R + k < nS;
      // the decision is constant on each of the D-intervals cut by {-logit u, 0, logit u} (just {0} when u <= 1/2):
      // a member whose whole D range falls inside one of them is decided whatever the counts are
      const int rlo = sm_d_region(dc_min + d0[k], lg[k]), rhi = sm_d_region(dc_max + d0[k], lg[k]);
      const bool robust = valid && rlo == rhi;
      newz[k] = robust ? (~rlo & 1) : zz[k];
      if (valid && !robust) nrmask |= 1u << k;
      pre[k] = run;
      run += newz[k] - zz[k];
    }
    const int nnr_t = __popc(nrmask);
    int incl = run, incn = nnr_t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(SMG_FULL, incl, o), yn = __shfl_up_sync(SMG_FULL, incn, o);
      if (lane >= o) {
        incl += y;
        incn += yn;
      }
    }
    if (lane == 31) {
      M.wsum[warp] = incl;
      M.wnr[warp] = incn;
    }
    __syncthreads();
    if (warp == 0) {
      int a = lane < T / 32 ? M.wsum[lane] : 0, b = lane < T / 32 ? M.wnr[lane] : 0, xa = a, xb = b;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int ya = __shfl_up_sync(SMG_FULL, xa, o), yb = __shfl_up_sync(SMG_FULL, xb, o);
        if (lane >= o) {
          xa += ya;
          xb += yb;
        }
      }
      M.wsum[lane] = xa - a;
      M.wnr[lane] = xb - b;
      if (lane == 31) {
        M.cdelta = xa;
        M.nnr = xb;
      }
    }
    __syncthreads();
    const int nnr = M.nnr;
    if (nnr == 0) {  // every member of the chunk was decided in parallel (the usual case once the sides have settled)
#pragma unroll
      for (int k = 0; k < R; k++)
        if (base + tid * R + k < nS) z[base + tid * R + k] = newz[k];
      nBcur += M.cdelta;
      continue;
    }
    {
      const int tpre = M.wsum[warp] + incl - run;      // side-1 change of the robust members of earlier threads
      int slot = M.wnr[warp] + incn - nnr_t;          // position of this thread's first non-robust member in the list
#pragma unroll
      for (int k = 0; k < R; k++) {
        const int id = tid * R + k;
        M.z[id] = newz[k];
        if ((nrmask >> k) & 1u) {
          M.pre[id] = tpre + pre[k];
          M.d0[id] = d0[k];
          M.lg[id] = lg[k];
          M.list[slot++] = id;
        }
      }
    }
    __syncthreads();
    if (warp == 0) {
      const int nB0 = nBcur;
      int extra = 0;  // side-1 change of the non-robust members decided so far
      for (int b0 = 0; b0 < nnr; b0 += 32) {
        const int q = b0 + lane;
        const bool v = q < nnr;
        const int t = v ? M.list[q] : 0;
        const double md0 = v ? M.d0[t] : 0.0, mlg = v ? M.lg[t] : 0.0;
        const int mpre = v ? M.pre[t] : 0;
        int mz = M.z[t];
        // Lanes [start, f) are settled together: lane l sees at most l - start undecided members of this block
        // before it, which bounds its count term as above; f is the first lane whose D range straddles a
        // threshold.  That one is evaluated exactly (double-precision logs, one lane) against the settled counts.
        int start = 0;
        while (start < 32 && b0 + start < nnr) {
          const bool act = v && lane >= start;
          const int nBq = nB0 + mpre + extra, b = nBq - mz, dev = lane - start;
          const int blo = max(1, b - dev), bhi = min(nS, b + dev);
          const int rlo = sm_d_region(sm_dc_bound(nS, bhi) - SM_DC_MARGIN + md0, mlg),
                    rhi = sm_d_region(sm_dc_bound(nS, blo) + SM_DC_MARGIN + md0, mlg);
          const unsigned fr = __ballot_sync(SMG_FULL, act && rlo != rhi);
          const int f = fr ? __ffs(fr) - 1 : 32;
          int d = 0;
          if (act && lane < f) {
            const int nz = ~rlo & 1;
            d = nz - mz;
            mz = nz;
          }
          extra += warp_sum_i(d);
          if (f < 32) {
            int df = 0;
            if (lane == f) {
              const int nBf = nB0 + mpre + extra, nAf = nS + 2 - nBf;
              const double dc = (mz == 0) ? log((double)(nAf - 1)) - log((double)nBf)
                                          : log((double)nAf) - log((double)(nBf - 1));
              const double D = dc + md0;
              int nz;
              if (D > 0.0)
                nz = (D >= mlg) ? 0 : 1;
              else
                nz = (-D >= mlg) ? 1 : 0;
              df = nz - mz;
              mz = nz;
            }
            extra += __shfl_sync(SMG_FULL, df, f);
          }
          start = f + 1;
        }
        if (v) M.z[t] = mz;
      }
      if (lane == 0) M.nB = extra;  // handed to every thread below
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < R; k++)
      if (base + tid * R + k < nS) z[base + tid * R + k] = M.z[tid * R + k];
    nBcur += M.cdelta + M.nB;
    __syncthreads();  // M.nB / M.cdelta / M.nnr are rewritten by the next chunk
  }
  if (tid == 0 && cnt2) {
    cnt2[0] = nS + 2 - nBcur;
    cnt2[1] = nBcur;
  }
}

__global__ void __launch_bounds__(SM_DECIDE_T) sm_rdecide_kernel(const SmInfo* info, const double* __restrict__ dl,
                                                                 const double* __restrict__ lgt, int* __restrict__ z,
                                                                 int* __restrict__ Hzero, int hlen, int* __restrict__ cnt2,
                                                                 const int* enable, int enable_val, int wide) {
  if (enable && *enable != enable_val) return;
  extern __shared__ __align__(16) unsigned char s_rd_raw[];
  if (wide) {
    RdecideSmem<SM_DECIDE_WIDE_CHUNK>& M = *reinterpret_cast<RdecideSmem<SM_DECIDE_WIDE_CHUNK>*>(s_rd_raw);
    sm_rdecide_body<SM_DECIDE_T, SM_DECIDE_WIDE_CHUNK / SM_DECIDE_T>(info->nS, dl, lgt, z, Hzero, hlen, cnt2, M);
  } else {
    RdecideSmem<SM_DECIDE_T * SM_DECIDE_R>& M = *reinterpret_cast<RdecideSmem<SM_DECIDE_T * SM_DECIDE_R>*>(s_rd_raw);
    sm_rdecide_body<SM_DECIDE_T, SM_DECIDE_R>(info->nS, dl, lgt, z, Hzero, hlen, cnt2, M);
  }
}

// proposal = copy of the split launch state (sides and the two parameter slots), split_merge.cpp:575-577
__global__ void __launch_bounds__(256) sm_begin_proposal_kernel(const SmInfo* info, const int* __restrict__ zL,
                                                                int* __restrict__ zStar, int pp, uint8_t* cen,
                                                                double* sig, double* isg, double* sden, int srcA,
                                                                int dstA, int srcB, int dstB) {
  const int nS = info->nS;
  for (int pos = blockIdx.x * blockDim.x + threadIdx.x; pos < nS; pos += gridDim.x * blockDim.x) zStar[pos] = zL[pos];
  if (blockIdx.x < 2) {
    const int src = blockIdx.x ? srcB : srcA, dst = blockIdx.x ? dstB : dstA;
    for (int j = threadIdx.x; j < pp; j += blockDim.x) {
      cen[(size_t)dst * pp + j] = cen[(size_t)src * pp + j];
      sig[(size_t)dst * pp + j] = sig[(size_t)src * pp + j];
      isg[(size_t)dst * pp + j] = isg[(size_t)src * pp + j];
    }
    if (threadIdx.x == 0) sden[dst] = sden[src];
  }
}

// H[dst] = H[a] + H[b] ; cnt likewise
__global__ void sm_hist_add_kernel(int len, int* H, int* cnt, int a, int b, int dst) {
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < len) H[(size_t)dst * len + q] = H[(size_t)a * len + q] + H[(size_t)b * len + q];
  if (q == 0) cnt[dst] = cnt[a] + cnt[b];
}

// ------------------------------------------------------------------------------------------
// The restricted-scan chain as ONE persistent cooperative kernel.
//
// t launch scans + the proposal scan are 4 dependent phases each (likelihood differences -> ordered decision ->
// side histograms -> parameter update): as separate launches every phase starts on cold SMs (ncu: 38% of the
// parameter-update kernel's stall samples are instruction-fetch starvation) and pays a launch gap.  Here a small grid
// of co-resident CTAs (cudaLaunchCooperativeKernel) walks the whole chain, phases separated by a grid barrier
// (release add / acquire poll on one global counter).  Same device functions and the same Philox keys as the
// multi-launch path: the two are interchangeable bit for bit (tests run both).  Measured at the metric config:
// 0.73 ms per proposal against 0.85 ms as ~50 launches, with 120 CTAs (24 CTAs: 0.97 ms, the likelihood phase
// starves).  A gang-scheduled 120-CTA grid does not share the GPU well, so chains that are stepped together
// (smg_step_many) use a gang of 8 CTAs when they are small (n <= 30000: fewer launches per sweep) and the multi-launch
// path otherwise.
// ------------------------------------------------------------------------------------------
enum SmJob : int { J_PRI_A = 0, J_PRI_B, J_PRI_M, J_L0, J_L1, J_MG, J_P0, J_P1, J_MSTAR, J_N };

__host__ __device__ inline PhiJob sm_job_at(int B, int which, uint32_t sub, const double* uc, const double* us, int enable_mode) {
  // (hist, current sigma slot, destination slot, count index) of J_PRI_A .. J_MSTAR
  const int hist[J_N] = {0, 0, 0, SH_L0, SH_L1, SH_M, SH_P0, SH_P1, SH_M};
  const int src[J_N] = {0, 0, 0, SM_SL_A, SM_SL_B, SM_ML_M, SM_ST_A, SM_ST_B, SM_ML_M};
  const int dst[J_N] = {SM_SL_A, SM_SL_B, SM_ML_M, SM_SL_A, SM_SL_B, SM_ML_M, SM_ST_A, SM_ST_B, SM_ST_M};
  const int cix[J_N] = {0, 0, 0, SH_L0, SH_L1, SH_M, SH_P0, SH_P1, SH_M};
  PhiJob J;
  J.hist = hist[which];
  J.src = B + src[which];
  J.dst = B + dst[which];
  J.cnt_idx = cix[which];
  J.sub = sub;
  J.prior = which <= J_PRI_M;
  J.enable_mode = enable_mode;
  J.uc = uc;
  J.us = us;
  return J;
}
struct GridBar {
  unsigned* counter;
  unsigned target, nblocks;
  int* err;
  int* status;
};
__device__ __forceinline__ void grid_sync(GridBar& B) {
  __syncthreads();
  B.target += B.nblocks;
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(B.counter) : "memory");
    unsigned seen = 0, spins = 0;
    for (;;) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(B.counter) : "memory");
      if (seen >= B.target) break;
      if (++spins > (1u << 22)) {  // a lost CTA must not hang the device: flag it and let everybody run out
        atomicExch(B.err, 1);
        if (B.status) atomicOr(B.status, ST_GRID_TIMEOUT);
        break;
      }
    }
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------
// MH terms
// ------------------------------------------------------------------------------------------
// logprobgs_phi (split_merge.cpp:20-94): block b evaluates term b; also priors (:419-436) in blocks 3..5
// (no __restrict__: inside sm_chain_kernel these arrays were written earlier in the same launch.)  Term b by the first
// 256 threads of the calling CTA; all of its threads must call.
// value of attribute j of term b (b < 3: logprobgs_phi addend of split_merge.cpp:60-92; b >= 3: priors addend :428-433)
__device__ __forceinline__ double sm_gsphi_prior_attr(int b, int j, int pp, int mmax, const int* attr, const double* v,
                                                      const double* w, const int* H, const int* cnt, const SmPlan* plan,
                                                      const uint8_t* cen, const double* sig) {
  const int len = pp * mmax;
  const int m = attr[j];
  if (b < 3) {
    const int hist = plan->gs_hist[b], sL = plan->gs_sigL[b], st = plan->gs_star[b];
    const int nm = cnt[hist];
    const int* h = H + (size_t)hist * len + (size_t)j * mmax;
    const double sg = sig[(size_t)sL * pp + j];
    // compute_prob_centers (common_functions.cpp:495-505)
    double pt[SMG_MAX_LEVELS];
    double mx = -CUDART_INF;
    for (int a = 0; a < m; a++) {
      pt[a] = -((double)nm - (double)h[a]) / sg;
      mx = pt[a] > mx ? pt[a] : mx;
    }
    double sum = 0.0;
    for (int a = 0; a < m; a++) {
      pt[a] = exp(pt[a] - mx);
      sum += pt[a];
    }
    const int cs = cen[(size_t)st * pp + j];
    const double lc = log(pt[cs - 1] / sum);
    const double sm = (double)h[cs - 1];
    const double ls = logdensity_hig_d(sig[(size_t)st * pp + j], v[j] + sm, w[j] + (double)nm - sm, (double)m);
    return lc + ls;
  }
  const int slot = plan->pri_slot[b - 3];
  return -log((double)m) + logdensity_hig_d(sig[(size_t)slot * pp + j], v[j], w[j], (double)m);
}
__device__ __forceinline__ void sm_gsphi_prior_body(int b, int pp, int p, int mmax, const int* attr, const double* v,
                                                    const double* w, const int* H, const int* cnt, const SmPlan* plan,
                                                    const uint8_t* cen, const double* sig, double* terms, double* sh) {
  double acc = 0.0;
  if (threadIdx.x < 256)
    for (int j = threadIdx.x; j < p; j += 256) acc += sm_gsphi_prior_attr(b, j, pp, mmax, attr, v, w, H, cnt, plan, cen, sig);
  if (threadIdx.x < 256) sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  // terms layout: [4..6] pri, [10..12] gs_phi
  if (threadIdx.x == 0) terms[b < 3 ? 10 + b : 4 + (b - 3)] = sh[0];
  __syncthreads();
}
__global__ void __launch_bounds__(256) sm_gsphi_prior_kernel(int pp, int p, int mmax, const int* __restrict__ attr,
                                                             const double* __restrict__ v, const double* __restrict__ w,
                                                             const int* __restrict__ H, const int* __restrict__ cnt,
                                                             const SmPlan* plan, const uint8_t* cen, const double* sig,
                                                             double* terms) {
  __shared__ double sh[256];
  sm_gsphi_prior_body(blockIdx.x, pp, p, mmax, attr, v, w, H, cnt, plan, cen, sig, terms, sh);
}

// per-member likelihood values feeding loglikelihood_hamming (:393-417) and logprobgs_c_i (:96-161)
// rowvals[0][r] = LL under its own side's parameters if side 0 else 0
// rowvals[1][r] = same for side 1 ; rowvals[2][r] = LL under the merged/old parameters
// rowvals[3][r] = log prob of the member's side under the launch counts (0 for the anchors)
__device__ __forceinline__ void sm_rowterms_body(const uint8_t* __restrict__ X, int pp, const int* S, const SmInfo* info,
                                                 const SmPlan* plan, const int* zL, const int* zStar, const int* zState,
                                                 const int* cnt, const uint8_t* cen, const double* isg, const double* sden,
                                                 double* rowvals, int stride, int gw, int nw) {
  const int nS = info->nS;
  const int lane = threadIdx.x & 31;
  for (int r = gw; r < nS + 2; r += nw) {
  const int row = r < nS ? S[r] : (r == nS ? info->i1 : info->i2);
  const int* zs = plan->zsel ? zState : zStar;
  const int side = r < nS ? zs[r] : (r - nS);
  const int sA = plan->slotA, sB = plan->slotB, sM = plan->slotAll;
  const uint8_t* x = X + (size_t)row * pp;
  double llA = -warp_mismatch_dot(x, cen + (size_t)sA * pp, isg + (size_t)sA * pp, pp, lane) - sden[sA];
  double llB = -warp_mismatch_dot(x, cen + (size_t)sB * pp, isg + (size_t)sB * pp, pp, lane) - sden[sB];
  double llM = -warp_mismatch_dot(x, cen + (size_t)sM * pp, isg + (size_t)sM * pp, pp, lane) - sden[sM];
  if (lane == 0) {
    rowvals[0 * (size_t)stride + r] = side == 0 ? llA : 0.0;
    rowvals[1 * (size_t)stride + r] = side == 1 ? llB : 0.0;
    rowvals[2 * (size_t)stride + r] = llM;
    double gc = 0.0;
    if (r < nS) {
      const int zl = zL[r];
      const int nA = cnt[SH_L0] - (zl == 0), nB = cnt[SH_L1] - (zl == 1);
      double a0 = log((double)nA) + llA, a1 = log((double)nB) + llB;
      double mx = a0 > a1 ? a0 : a1;
      double p0 = exp(a0 - mx), p1 = exp(a1 - mx);
      double sm = 0.0;
      sm += p0;
      sm += p1;
      gc = log((side == 0 ? p0 : p1) / sm);
    }
    rowvals[3 * (size_t)stride + r] = gc;
  }
  }
}
__global__ void __launch_bounds__(256) sm_rowterms_kernel(const uint8_t* __restrict__ X, int pp, const int* __restrict__ S,
                                                          const SmInfo* info, const SmPlan* plan,
                                                          const int* __restrict__ zL, const int* __restrict__ zStar,
                                                          const int* __restrict__ zState, const int* __restrict__ cnt,
                                                          const uint8_t* cen, const double* isg, const double* sden,
                                                          double* __restrict__ rowvals, int stride) {
  sm_rowterms_body(X, pp, S, info, plan, zL, zStar, zState, cnt, cen, isg, sden, rowvals, stride,
                   (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), (gridDim.x * blockDim.x) >> 5);
}

// fixed-order two-stage sums of the four row-value arrays
// partial sum bx of array q by one warp (identical to sm_rowreduce1_body)
__device__ __forceinline__ void sm_rowreduce1_warp(int nS, const double* rowvals, int stride, double* partial, int bx, int q,
                                                   int lane) {
  const int nr = nS + 2;
  const int per = (nr + SM_RB - 1) / SM_RB;
  const int lo = bx * per, hi = min(nr, lo + per);
  double v[8];
#pragma unroll
  for (int k = 0; k < 8; k++) {
    double acc = 0.0;
    for (int r = lo + lane + 32 * k; r < hi; r += 256) acc += rowvals[(size_t)q * stride + r];
    v[k] = acc;
  }
  const double t = tree256_warp(v);
  if (lane == 0) partial[q * SM_RB + bx] = t;
}
// (partial sum bx of array q by the first 256 threads of the calling CTA; all of its threads must call)
__device__ __forceinline__ void sm_rowreduce1_body(int nS, const double* rowvals, int stride, double* partial, int bx, int q,
                                                   double* sh) {
  const int nr = nS + 2;
  const int per = (nr + SM_RB - 1) / SM_RB;
  const int lo = bx * per, hi = min(nr, lo + per);
  if (threadIdx.x < 256) {
    double acc = 0.0;
    for (int r = lo + threadIdx.x; r < hi; r += 256) acc += rowvals[(size_t)q * stride + r];
    sh[threadIdx.x] = acc;
  }
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[q * SM_RB + bx] = sh[0];
  __syncthreads();
}
__global__ void __launch_bounds__(256) sm_rowreduce1_kernel(const SmInfo* info, const double* __restrict__ rowvals,
                                                            int stride, double* __restrict__ partial) {
  __shared__ double sh[256];
  sm_rowreduce1_body(info->nS, rowvals, stride, partial, blockIdx.x, blockIdx.y, sh);
}

// final sums + MH log-ratio in the reference's order of additions + accept decision
__device__ __forceinline__ void sm_accept_body(const SmInfo* info, const SmPlan* plan, const int* cnt, const double* partial,
                                               double gamma, const double* u_inj, const RngKey& key, double* terms,
                                               int* accepted, unsigned long long* stats, double* sh) {
  __shared__ double tot[4];
  static_assert(SM_RB == 256, "the final sums are 256-leaf trees");
  if (threadIdx.x < 128) {  // warp q sums array q
    const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = partial[q * SM_RB + lane + 32 * k];
    const double t = tree256_warp(v);
    if (lane == 0) tot[q] = t;
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  // terms: 0 log_alpha, 1..3 lg, 4..6 pri, 7..9 ll, 10..12 gs_phi, 13 gs_c, 14 log_prior, 15 log_lik,
  //        16 log_prop, 17 log_ratio, 18 u_accept
  const double la = log(gamma);
  double lg[3];
  for (int q = 0; q < 3; q++) lg[q] = lgamma((double)cnt[plan->lg_cnt[q]]);
  const double pri0 = terms[4], pri1 = terms[5], pri2 = terms[6];
  const double ll0 = tot[0], ll1 = tot[1], ll2 = tot[2], gsc = tot[3];
  const double g0 = terms[10], g1 = terms[11], g2 = terms[12];
  double log_prior = 0.0, log_lik = 0.0, log_prop = 0.0;
  if (info->same) {  // split_acc_prob
    log_prior += la;
    log_prior += lg[0];
    log_prior += lg[1];
    log_prior += pri0;
    log_prior += pri1;
    log_prior -= lg[2];
    log_prior -= pri2;
    log_lik += ll0;
    log_lik += ll1;
    log_lik -= ll2;
    log_prop += g2;
    log_prop -= g0;
    log_prop -= g1;
    log_prop -= gsc;
  } else {  // merge_acc_prob
    log_prior += lg[2];
    log_prior += pri2;
    log_prior -= la;
    log_prior -= lg[0];
    log_prior -= lg[1];
    log_prior -= pri0;
    log_prior -= pri1;
    log_lik += ll2;
    log_lik -= ll0;
    log_lik -= ll1;
    log_prop += g0;
    log_prop += g1;
    log_prop += gsc;
    log_prop -= g2;
  }
  const double ratio = fmin(0.0, log_prior + log_lik + log_prop);
  const double u = get_u(u_inj, 0, key, U_SM_ACCEPT, 0u, 0u);
  terms[0] = la;
  terms[1] = lg[0];
  terms[2] = lg[1];
  terms[3] = lg[2];
  terms[7] = ll0;
  terms[8] = ll1;
  terms[9] = ll2;
  terms[13] = gsc;
  terms[14] = log_prior;
  terms[15] = log_lik;
  terms[16] = log_prop;
  terms[17] = ratio;
  terms[18] = u;
  const int acc = (log(u) < ratio) ? 1 : 0;  // split_merge.cpp:591 (NaN ratio => reject)
  *accepted = acc;
  if (stats && acc) stats[7]++;
}
__global__ void __launch_bounds__(256) sm_accept_kernel(const SmInfo* info, const SmPlan* plan, const int* __restrict__ cnt,
                                                        const double* __restrict__ partial, double gamma,
                                                        const double* u_inj, RngKey key, double* terms, int* accepted,
                                                        unsigned long long* stats) {
  __shared__ double sh[256];
  sm_accept_body(info, plan, cnt, partial, gamma, u_inj, key, terms, accepted, stats, sh);
}

// ------------------------------------------------------------------------------------------
// accept: state <- proposal, labels compacted as clean_var does (common_functions.cpp:296-353)
// ------------------------------------------------------------------------------------------
// step 1 (one CTA): parameters, counts, K
__device__ __forceinline__ void sm_apply_params_body(const SmInfo* info, int NS, int Kcap, int pp, uint8_t* cen, double* sig,
                                                     double* isg, double* sden, const int* cnt, int* counts, int* Kptr,
                                                     int* status) {
  const int K = info->K, cA = info->cA, cB = info->cB;
  auto copy = [&](int src, int dst) {
    for (int j = threadIdx.x; j < pp; j += blockDim.x) {
      cen[(size_t)dst * pp + j] = cen[(size_t)src * pp + j];
      sig[(size_t)dst * pp + j] = sig[(size_t)src * pp + j];
      isg[(size_t)dst * pp + j] = isg[(size_t)src * pp + j];
    }
    if (threadIdx.x == 0) sden[dst] = sden[src];
    __syncthreads();
  };
  if (info->same) {
    if (K + 1 > Kcap) {
      if (threadIdx.x == 0) atomicOr(status, ST_LL_COLS);
      return;
    }
    copy(NS + SM_ST_A, K);   // the new label K takes the i_1 side
    copy(NS + SM_ST_B, cB);  // the old label keeps the i_2 side
    if (threadIdx.x == 0) {
      counts[K] = cnt[SH_P0];
      counts[cB] = cnt[SH_P1];
      *Kptr = K + 1;
    }
  } else {
    const int hole = cA, last = K - 1;
    copy(NS + SM_ST_M, cB);
    if (threadIdx.x == 0) counts[cB] = cnt[SH_M];
    __syncthreads();
    if (hole != last) {
      copy(last, hole);
      if (threadIdx.x == 0) counts[hole] = counts[last];
    }
    if (threadIdx.x == 0) {
      counts[last] = 0;
      *Kptr = K - 1;
    }
  }
}
__global__ void __launch_bounds__(256) sm_apply_params_kernel(const SmInfo* info, const int* accepted, int NS, int Kcap,
                                                              int pp, uint8_t* cen, double* sig, double* isg, double* sden,
                                                              const int* __restrict__ cnt, int* counts, int* Kptr,
                                                              int* status) {
  if (*accepted == 0) return;
  sm_apply_params_body(info, NS, Kcap, pp, cen, sig, isg, sden, cnt, counts, Kptr, status);
}
// step 2: labels of the members (thread gt of nt)
__device__ __forceinline__ void sm_apply_members_body(const SmInfo* info, const int* S, const int* zStar, int* c, int gt,
                                                      int nt) {
  const int nS = info->nS, same = info->same, K = info->K, cB = info->cB;
  for (int r = gt; r < nS + 2; r += nt) {
    const int row = r < nS ? S[r] : (r == nS ? info->i1 : info->i2);
    if (same) {
      const int side = r < nS ? zStar[r] : (r - nS);
      c[row] = side == 0 ? K : cB;
    } else {
      c[row] = cB;
    }
  }
}
__global__ void sm_apply_members_kernel(const SmInfo* info, const int* accepted, const int* __restrict__ S,
                                        const int* __restrict__ zStar, int* __restrict__ c) {
  if (*accepted == 0) return;
  sm_apply_members_body(info, S, zStar, c, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}
// step 3 (merge only): the last label moves into the hole
__device__ __forceinline__ void sm_apply_relabel_body(const SmInfo* info, int n, int* c, int gt, int nt) {
  if (info->same) return;
  const int hole = info->cA, last = info->K - 1;
  if (hole == last) return;
  for (int i = gt; i < n; i += nt)
    if (c[i] == last) c[i] = hole;
}
__global__ void sm_apply_relabel_kernel(const SmInfo* info, const int* accepted, int n, int* __restrict__ c) {
  if (*accepted == 0) return;
  sm_apply_relabel_body(info, n, c, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}

// The likelihood block of the next pass is evaluated CONCURRENTLY with the proposal (sweep()): when the proposal is
// accepted, the columns whose parameters it rewrote are stale (or torn) and are re-evaluated here over all rows --
// split: the new label K and label c_B; merge: label c_B (merged parameters) and the hole c_A that received the last
// label's parameters (sm_apply_params_body).  Not accepted (the usual case): returns at once.
__global__ void __launch_bounds__(256) sm_ll_patch_kernel(const int* __restrict__ accepted, const SmInfo* __restrict__ info,
                                                          const uint8_t* __restrict__ X, int n, int pp, const uint8_t* cen,
                                                          const double* isg, const double* sden, double* __restrict__ LL,
                                                          int ldl) {
  if (*accepted == 0) return;
  int col[2], nc = 0;
  if (info->same) {
    col[nc++] = info->K;
    col[nc++] = info->cB;
  } else {
    col[nc++] = info->cB;
    if (info->cA != info->K - 1) col[nc++] = info->cA;
  }
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int q = 0; q < nc; q++) {
    const int slot = col[q];
    if (slot < 0 || slot >= ldl) continue;
    const uint8_t* crow = cen + (size_t)slot * pp;
    const double* wrow = isg + (size_t)slot * pp;
    const double sd = sden[slot];
    for (int row = gw * 4; row < n; row += nw * 4) {
      double acc[4] = {0.0, 0.0, 0.0, 0.0};
      for (int j0 = lane * 8; j0 < pp; j0 += 256) {
        const uint2 cv = *reinterpret_cast<const uint2*>(crow + j0);
        uint2 xv[4];
#pragma unroll
        for (int r = 0; r < 4; r++) xv[r] = (row + r < n) ? *reinterpret_cast<const uint2*>(X + (size_t)(row + r) * pp + j0) : cv;
        const double2* w = reinterpret_cast<const double2*>(wrow + j0);
        const double2 w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3];
#pragma unroll
        for (int r = 0; r < 4; r++) {
          const uint32_t m0 = __vcmpne4(xv[r].x, cv.x), m1 = __vcmpne4(xv[r].y, cv.y);
          if (m0 & 0x000000ffu) acc[r] += w0.x;
          if (m0 & 0x0000ff00u) acc[r] += w0.y;
          if (m0 & 0x00ff0000u) acc[r] += w1.x;
          if (m0 & 0xff000000u) acc[r] += w1.y;
          if (m1 & 0x000000ffu) acc[r] += w2.x;
          if (m1 & 0x0000ff00u) acc[r] += w2.y;
          if (m1 & 0x00ff0000u) acc[r] += w3.x;
          if (m1 & 0xff000000u) acc[r] += w3.y;
        }
      }
#pragma unroll
      for (int r = 0; r < 4; r++) {
        const double dot = warp_sum(acc[r]);
        if (lane == 0 && row + r < n) LL[(size_t)(row + r) * ldl + slot] = -dot - sd;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// the persistent kernel (see the comment above sm_job_at)
// ------------------------------------------------------------------------------------------
struct SmChainArgs {
  int n, p, pp, mmax, t, r, NS, wide_from, Kcap, hist_ctas;
  double gamma;
  const uint8_t* X;
  int* c;        // labels (read by the selection, rewritten when the proposal is accepted)
  int* counts;   // cluster sizes by label
  int* Kptr;     // number of clusters
  int* S;
  SmInfo* info;
  SmPlan* plan;
  uint8_t* cen;
  double *sig, *isg, *sden;
  int *H, *cnt, *zL, *zStar, *zState;
  int* selcnt;   // [grid] members found by each CTA
  double *dl, *lgt, *lgt2;  // lgt / lgt2: logit(u) of the members for even / odd scans
  double *rowvals, *partial, *terms;
  int* accepted;
  unsigned long long* stats;
  PhiArgs phi;  // common fields of the parameter updates (jobs are filled on the device)
  const double *u_rg, *u_rg_c, *u_rg_s, *u_mg_c, *u_mg_s;  // injected uniforms (bases) or null
  const double *u_pair, *u_prior_c, *u_prior_s, *u_launch, *u_accept;
  int pair_det;  // 1: select_observations_deterministic
  int* exc;      // [2] members not certain to keep their side in the scan in progress (by scan parity)
  RngKey key;
  unsigned* bar;       // grid-barrier counter of this launch (zero on entry)
  unsigned* bar_next;  // the counter of the next launch: zeroed here
  int* err;
};

#define SM_CHAIN_CTAS 120
#define SM_CHAIN_T 512  // 128 registers per thread: the parameter-update body does not spill

__global__ void __launch_bounds__(SM_CHAIN_T, 1) sm_chain_kernel(SmChainArgs A) {
  // dynamic shared memory: the side histograms of the histogram phase and the scratch of the decision phase
  // (never live at the same time)
  extern __shared__ __align__(16) int s_hist[];
  RdecideSmem<SM_CHAIN_T * SM_DECIDE_R>& M = *reinterpret_cast<RdecideSmem<SM_CHAIN_T * SM_DECIDE_R>*>(s_hist);
  RdecideSmem<SM_DECIDE_WIDE_CHUNK>& MW = *reinterpret_cast<RdecideSmem<SM_DECIDE_WIDE_CHUNK>*>(s_hist);
  __shared__ double sh[256];
  GridBar B{A.bar, 0u, gridDim.x, A.err, A.phi.status};
  int nS = 0, same = 0;  // known after the selection phase
  const int n = A.n, p = A.p, pp = A.pp, NSB = A.NS;
  constexpr int WPB = SM_CHAIN_T / 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gwarp = blockIdx.x * WPB + warp, nwarps = gridDim.x * WPB;
  const int gtid = blockIdx.x * SM_CHAIN_T + tid, gthreads = gridDim.x * SM_CHAIN_T;
  const size_t len = (size_t)pp * A.mmax;
  auto off = [](const double* base, size_t o) -> const double* { return base ? base + o : nullptr; };
  auto fill_lgt = [&](int q, int t0, int nt) {  // thread t0 of nt: logit(u) of scan q into its buffer
    RngKey k = A.key;
    k.sub = SUB_SM_RG + q;
    double* buf = (q & 1) ? A.lgt2 : A.lgt;
    const double* ui = off(A.u_rg, (size_t)q * n);
    for (int pos = t0; pos < nS; pos += nt) buf[pos] = sm_logit_u(ui, pos, k);
  };

#ifdef SMG_PHI_PROFILE
  long long tk = clock64();
#define CHAIN_TICK(k)                                                                   \
  do {                                                                                  \
    const long long _t = clock64();                                                     \
    if (blockIdx.x == 0 && threadIdx.x == 0 && A.phi.prof) A.phi.prof[k] += (unsigned long long)(_t - tk); \
    tk = _t;                                                                            \
  } while (0)
#else
#define CHAIN_TICK(k)
#endif
#ifdef SMG_SM_HT_PROFILE
#define HT_TICK(k) CHAIN_TICK(k)
#else
#define HT_TICK(k)
#endif
  // ---- pair and S (split_merge.cpp:263-301): ordered compaction over the whole grid.  Warp w owns the rows
  // [lo, hi); counts per warp and per CTA, one grid barrier, then every CTA knows the offset of its rows.
  __shared__ int s_sel[8];
  __shared__ int s_wcnt[WPB], s_woff[WPB];
  if (tid == 0) {
    RngKey k = A.key;
    k.sub = SUB_SM_SELECT;
    int i1, i2;
    sm_pick_pair(n, A.u_pair, k, &i1, &i2, A.pair_det);
    s_sel[0] = i1;
    s_sel[1] = i2;
    s_sel[2] = A.c[i1];
    s_sel[3] = A.c[i2];
    if (blockIdx.x == 0) {
      *A.bar_next = 0u;
      if (A.exc) A.exc[0] = A.exc[1] = 0;
    }
  }
  if (blockIdx.x == 0)
    for (int q = tid; q < 24; q += SM_CHAIN_T) A.terms[q] = 0.0;
  for (size_t q = gtid; q < 2 * len; q += gthreads) A.H[(size_t)SH_S0 * len + q] = 0;  // current-state side histograms
  if (gtid == 0) A.cnt[SH_S0] = A.cnt[SH_S1] = 0;
  __syncthreads();
  const int i1 = s_sel[0], i2 = s_sel[1], cA = s_sel[2], cB = s_sel[3];
  // (the CTAs that draw the prior parameters below take no rows when the grid is large enough)
  const int pri_ctas = 3 * A.phi.nparts;
  const int sel_warps = ((int)gridDim.x - pri_ctas >= 4 ? (int)gridDim.x - pri_ctas : (int)gridDim.x) * WPB;
  const int seg = ((n + sel_warps - 1) / sel_warps + 31) & ~31;
  const int lo = gwarp < sel_warps ? (int)min((long long)n, (long long)gwarp * seg) : n, hi = min(n, lo + seg);
  {
    int mine = 0;
    for (int i = lo + lane; i < hi; i += 32) {
      const int ci = A.c[i];
      mine += (i != i1 && i != i2 && (ci == cA || ci == cB));
    }
    mine = warp_sum_i(mine);
    if (lane == 0) s_wcnt[warp] = mine;
  }
  __syncthreads();
  if (warp == 0) {
    const int v = lane < WPB ? s_wcnt[lane] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(SMG_FULL, x, o);
      if (lane >= o) x += y;
    }
    if (lane < WPB) s_woff[lane] = x - v;
    if (lane == 31) A.selcnt[blockIdx.x] = x;
  }
  {  // prior parameters of the three launch clusters (split_merge.cpp:331-343, :379-380), on the last CTAs
    const int b0 = (int)gridDim.x - pri_ctas;
    if ((int)blockIdx.x >= b0) {
      const int w = (int)blockIdx.x - b0, k = w % 3;
      const PhiJob J = sm_job_at(NSB, J_PRI_A + k, SUB_SM_PRIOR, off(A.u_prior_c, (size_t)k * p), off(A.u_prior_s, (size_t)k * p), 0);
      phi_job_body(A.phi, J, k, w / 3, A.phi.nparts, sh);
    }
  }
  grid_sync(B);
  HT_TICK(0);
  if (warp == 0) {
    int before = 0, total = 0;
    for (int b = lane; b < (int)gridDim.x; b += 32) {
      const int v = __ldcg(&A.selcnt[b]);
      total += v;
      if (b < (int)blockIdx.x) before += v;
    }
    total = warp_sum_i(total);
    before = warp_sum_i(before);
    if (lane == 0) {
      s_sel[4] = total;
      s_sel[5] = before;
    }
  }
  __syncthreads();
  nS = s_sel[4];
  same = (cA == cB);
  {
    int offp = s_sel[5] + s_woff[warp];
    for (int base = lo; base < hi; base += 32) {
      const int i = base + lane;
      const int ci = (i < hi) ? A.c[i] : -1;
      const bool in = (i < hi) && i != i1 && i != i2 && (ci == cA || ci == cB);
      const unsigned bal = __ballot_sync(SMG_FULL, in);
      if (in) {
        const int pos = offp + __popc(bal & ((1u << lane) - 1));
        A.S[pos] = i;
        A.zState[pos] = (ci == cA) ? 0 : 1;
      }
      offp += __popc(bal);
    }
  }
  if (gtid == 0) sm_fill_info_plan(A.info, A.plan, NSB, i1, i2, nS, cA, cB, *A.Kptr);
  {  // random launch allocation (split_merge.cpp:346)
    RngKey k = A.key;
    k.sub = SUB_SM_LAUNCH;
    for (int pos = gtid; pos < nS; pos += gthreads) {
      const double u = get_u(A.u_launch, pos, k, U_SM_LAUNCH, (uint32_t)pos, 0u);
      const int z = (int)(2.0 * u);
      A.zL[pos] = z > 1 ? 1 : z;
    }
  }
  fill_lgt(0, gtid, gthreads);
  grid_sync(B);
  HT_TICK(1);

  // allocation part of restricted scan q on sides z / slots (slotA, slotB) -> histograms h0, h0+1
  auto alloc_scan = [&](int* z, int slotA, int slotB, int h0, int q, bool first, bool have_dl) -> bool {
    RngKey k = A.key;
    k.sub = SUB_SM_RG + q;
    CHAIN_TICK(7);
    double* lgq = (q & 1) ? A.lgt2 : A.lgt;
    if (!have_dl) {
    // Settled scans: once the launch state has found the two groups (after two or three scans of a merge proposal, which
    // is 98% of the proposals at K = 50), every member keeps its side with certainty.  The likelihood phase counts the
    // members for which that is NOT certain; when there is none, the serial decision, the side histograms and their two
    // grid barriers are skipped -- sides, counts and histograms are what they were.  (Not in the first scan, which also
    // builds the current-state histograms, nor in the proposal scan, whose histograms do not exist yet.)
    int* excq = (first || z != A.zL || !A.exc) ? nullptr : A.exc + (q & 1);
    sm_ll2prep_body(A.X, pp, A.S, nS, A.cen, A.isg, A.sden, slotA, slotB, off(A.u_rg, (size_t)q * n), k, A.dl, nullptr, gwarp,
                    nwarps, lgq, z, excq);
    // the histograms of the current-state sides (fixed for the whole proposal) ride in the first scan
    if (first)
      subset_hist_body(A.X, pp, A.S, nS, A.zState, &A.info->i1, A.mmax, A.H + (size_t)SH_S0 * len, A.cnt + SH_S0, s_hist,
                       blockIdx.x, gridDim.x);
    grid_sync(B);
    CHAIN_TICK(4);
    if (A.exc && gtid == 0) A.exc[(q + 1) & 1] = 0;  // the next scan's counter (last read one scan ago)
    if (excq && __ldcg(excq) == 0) {
      if (q < A.t) fill_lgt(q + 1, gtid, gthreads);  // the next scan's logits (normally evaluated beside the decision)
      return true;
    }
    }  // !have_dl
    if (first && blockIdx.x != 0) {  // ... and their sum, the merged cluster's histogram, while CTA 0 decides
      const int nt = ((int)gridDim.x - 1) * SM_CHAIN_T;
      for (int q2 = gtid - SM_CHAIN_T; q2 < (int)len; q2 += nt)
        A.H[(size_t)SH_M * len + q2] = A.H[(size_t)SH_S0 * len + q2] + A.H[(size_t)SH_S1 * len + q2];
      if (gtid == SM_CHAIN_T) A.cnt[SH_M] = A.cnt[SH_S0] + A.cnt[SH_S1];
    }
    if (blockIdx.x == 0) {
      if (q >= A.wide_from)
        sm_rdecide_body<SM_CHAIN_T, SM_DECIDE_WIDE_CHUNK / SM_CHAIN_T>(nS, A.dl, lgq, z, A.H + (size_t)h0 * len, (int)(2 * len),
                                                                       A.cnt + h0, MW);
      else
        sm_rdecide_body<SM_CHAIN_T, SM_DECIDE_R>(nS, A.dl, lgq, z, A.H + (size_t)h0 * len, (int)(2 * len), A.cnt + h0, M);
    } else if (q < A.t) {
      fill_lgt(q + 1, gtid - SM_CHAIN_T, ((int)gridDim.x - 1) * SM_CHAIN_T);  // the next scan's, while CTA 0 decides
    }
    grid_sync(B);
    CHAIN_TICK(5);
    // (every participating CTA flushes 2*len counters with global atomics: fewer CTAs than the grid do it)
    if ((int)blockIdx.x < A.hist_ctas)
      subset_hist_body(A.X, pp, A.S, nS, z, &A.info->i1, A.mmax, A.H + (size_t)h0 * len, nullptr, s_hist, blockIdx.x,
                       A.hist_ctas);
    grid_sync(B);
    CHAIN_TICK(6);
    return false;
  };
  // One loop, ONE call site for the scan phases and one for the parameter jobs: the kernel runs once per sweep on cold
  // instruction caches, and every extra inlined copy of these bodies costs ~10 us of instruction fetch the first time
  // it is reached.  Iterations 0 .. nsteps-1 are the launch scans (+ merge-launch updates), iteration nsteps is the
  // proposal: copy of the split launch state, one more restricted scan for a split, the merged cluster's final update.
  const int nsteps = A.t > A.r ? A.t : A.r;
  // Slots that hold the split-launch parameters of the two sides.  After a settled scan the next one is run
  // speculatively: its check ("nobody can move") on most CTAs and its parameter update on the others AT THE SAME TIME,
  // the update writing into the spare pair of slots; one grid barrier later the check says whether the update stands
  // (the pairs swap roles) or the scan is redone step by step from the untouched current pair.
  int curA = SM_SL_A, curB = SM_SL_B;
  bool prev_settled = false;
  for (int it = 0; it <= nsteps; it++) {
    const bool prop = (it == nsteps);
    if (prop) {
      // proposal = split launch state (sides + the two parameter slots) ...
      for (int pos = gtid; pos < nS; pos += gthreads) A.zStar[pos] = A.zL[pos];
      if (blockIdx.x < 4) {  // ... copied into the proposal slots, and back into the launch slots when the spare pair holds it
        const int side = blockIdx.x & 1;
        const int src = NSB + (side ? curB : curA), dst = NSB + (blockIdx.x < 2 ? (side ? SM_ST_B : SM_ST_A) : (side ? SM_SL_B : SM_SL_A));
        if (src != dst) {
          for (int jx = threadIdx.x; jx < pp; jx += blockDim.x) {
            A.cen[(size_t)dst * pp + jx] = A.cen[(size_t)src * pp + jx];
            A.sig[(size_t)dst * pp + jx] = A.sig[(size_t)src * pp + jx];
            A.isg[(size_t)dst * pp + jx] = A.isg[(size_t)src * pp + jx];
          }
          if (threadIdx.x == 0) A.sden[dst] = A.sden[src];
        }
      }
      grid_sync(B);
    }
    const int q = prop ? A.t : it;
    const bool do_scan = prop ? (same != 0) : (it < A.t);
    PhiJob j[3];
    int nj = 0;
    bool mg_done = false;
    if (do_scan) {
      bool have_dl = false;
      // (parts per job of the speculative update: at most half of the gang works on it, the rest checks)
      const int njs = (it < A.r) ? 3 : 2;
      const int nparts = min(A.phi.nparts, phi_parts_for(pp, max(1, (int)gridDim.x / (2 * njs))));
      if (!prop && prev_settled && A.exc && (int)gridDim.x >= njs * nparts + 8) {
        // ---- speculative scan: update (current pair -> spare pair) and the merged cluster's update on the first CTAs,
        //      the check and the next scan's logits on the others
        const int nxtA = curA == SM_SL_A ? SM_TMP0 : SM_SL_A, nxtB = curB == SM_SL_B ? SM_TMP1 : SM_SL_B;
        PhiJob js[3];
        for (int side = 0; side < 2; side++) {
          js[side] = sm_job_at(NSB, J_L0 + side, SUB_SM_RG + q, off(A.u_rg_c, ((size_t)q * 2 + side) * p),
                               off(A.u_rg_s, ((size_t)q * 2 + side) * p), 0);
          js[side].src = NSB + (side ? curB : curA);
          js[side].dst = NSB + (side ? nxtB : nxtA);
        }
        if (njs == 3) js[2] = sm_job_at(NSB, J_MG, SUB_SM_MERGE + it, off(A.u_mg_c, (size_t)it * p), off(A.u_mg_s, (size_t)it * p), 0);
        const int njp = njs * nparts;
        if ((int)blockIdx.x < njp) {
          phi_job_body(A.phi, js[blockIdx.x % njs], blockIdx.x % njs, blockIdx.x / njs, nparts, sh);
        } else {
          RngKey k = A.key;
          k.sub = SUB_SM_RG + q;
          const int cw = ((int)blockIdx.x - njp) * WPB + warp, cnw = ((int)gridDim.x - njp) * WPB;
          sm_ll2prep_body(A.X, pp, A.S, nS, A.cen, A.isg, A.sden, NSB + curA, NSB + curB, off(A.u_rg, (size_t)q * n), k, A.dl,
                          nullptr, cw, cnw, (q & 1) ? A.lgt2 : A.lgt, A.zL, A.exc + (q & 1));
          if (q < A.t) fill_lgt(q + 1, gtid - njp * SM_CHAIN_T, ((int)gridDim.x - njp) * SM_CHAIN_T);
        }
        if (gtid == 0) A.exc[(q + 1) & 1] = 0;
        grid_sync(B);
        mg_done = njs == 3;
        if (__ldcg(A.exc + (q & 1)) == 0) {  // nobody moves: the update stands
          curA = nxtA;
          curB = nxtB;
          if (*(volatile int*)A.err) return;
          continue;
        }
        have_dl = true;  // redo from the decision phase: the likelihood differences are in place
      }
      prev_settled = alloc_scan(prop ? A.zStar : A.zL, NSB + (prop ? SM_ST_A : curA), NSB + (prop ? SM_ST_B : curB),
                                prop ? SH_P0 : SH_L0, q, it == 0, have_dl);
      for (int side = 0; side < 2; side++) {
        j[nj] = sm_job_at(NSB, (prop ? J_P0 : J_L0) + side, SUB_SM_RG + q, off(A.u_rg_c, ((size_t)q * 2 + side) * p),
                          off(A.u_rg_s, ((size_t)q * 2 + side) * p), 0);
        if (!prop) j[nj].src = j[nj].dst = NSB + (side ? curB : curA);
        nj++;
      }
    }
    // the job index enters the Philox counter: the merged cluster's final update keeps index 2 as in the multi-launch path
    int idx0 = 0;
    if (prop) {
      if (!do_scan) idx0 = 2;
      j[nj++] = sm_job_at(NSB, J_MSTAR, SUB_SM_MERGE + A.r, off(A.u_mg_c, (size_t)A.r * p), off(A.u_mg_s, (size_t)A.r * p), 0);
    } else if (it < A.r && !mg_done) {
      j[nj++] = sm_job_at(NSB, J_MG, SUB_SM_MERGE + it, off(A.u_mg_c, (size_t)it * p), off(A.u_mg_s, (size_t)it * p), 0);
    }
    if ((int)blockIdx.x < nj * A.phi.nparts)
      phi_job_body(A.phi, j[blockIdx.x % nj], idx0 + blockIdx.x % nj, blockIdx.x / nj, A.phi.nparts, sh);
    grid_sync(B);
    if (*(volatile int*)A.err) return;
  }
  if (*(volatile int*)A.err) return;
  HT_TICK(7);
  // ---- MH terms: the six parameter-density terms on CTAs 0..5, the per-member likelihood terms on the others
  {
    const int cb = gridDim.x >= 16 ? 6 : 0;
    if (blockIdx.x < 6)
      sm_gsphi_prior_body(blockIdx.x, pp, p, A.mmax, A.phi.attr, A.phi.v, A.phi.w, A.H, A.cnt, A.plan, A.cen, A.sig, A.terms, sh);
    if ((int)blockIdx.x >= cb)
      sm_rowterms_body(A.X, pp, A.S, A.info, A.plan, A.zL, A.zStar, A.zState, A.cnt, A.cen, A.isg, A.sden, A.rowvals, n + 2,
                       ((int)blockIdx.x - cb) * WPB + warp, ((int)gridDim.x - cb) * WPB);
  }
  grid_sync(B);
  HT_TICK(2);
  for (int vb = gwarp; vb < 4 * SM_RB; vb += nwarps) sm_rowreduce1_warp(nS, A.rowvals, n + 2, A.partial, vb % SM_RB, vb / SM_RB, lane);
  grid_sync(B);
  if (blockIdx.x == 0) {
    RngKey k = A.key;
    k.sub = SUB_SM_ACCEPT;
    sm_accept_body(A.info, A.plan, A.cnt, A.partial, A.gamma, A.u_accept, k, A.terms, A.accepted, A.stats, sh);
  }
  grid_sync(B);
  HT_TICK(3);
  // ---- accept: state <- proposal
  if (__ldcg(A.accepted) == 0) return;
  if (blockIdx.x == 0)
    sm_apply_params_body(A.info, NSB, A.Kcap, pp, A.cen, A.sig, A.isg, A.sden, A.cnt, A.counts, A.Kptr, A.phi.status);
  sm_apply_members_body(A.info, A.S, A.zStar, A.c, gtid, gthreads);
  if (same) return;
  grid_sync(B);
  sm_apply_relabel_body(A.info, n, A.c, gtid, gthreads);
}

}  // namespace smg
