// smg_smc.cuh -- the Jain-Neal split-merge proposal (code/split_merge.cpp:542-598) as ONE thread-block cluster.
//
// sm_chain_kernel (smg_sm.cuh) runs the proposal on 120 CTAs that meet ~55 times at a global-counter grid barrier
// (1.2 us each), re-read the member rows from L2 in every phase and re-fetch the long parameter-draw code from L2 in
// every scan; 57% of a sweep was spent there.  Here the same proposal lives on one cluster of CS (16) CTAs with two
// roles, so that each SM's instruction footprint per scan stays inside its instruction cache:
//   * MEMBER CTAs (the first half): the members of S (+ the two anchors) are dealt to them in contiguous position
//     blocks; a CTA keeps its members' rows (transposed: word-major, one thread per member), sides and allocation
//     logits in shared memory for the whole proposal.  Per scan it builds the table
//     T[j][a] = [a != c_Bj]/sigma_Bj - [a != c_Aj]/sigma_Aj  and gets LL_A - LL_B of a member as 256 look-ups + adds;
//     it maintains the side histograms of its members incrementally (only members that change side are moved).
//   * PARAMETER CTAs (the second half): each owns pp/(CS/2) attributes; it sums its slice of the member CTAs' side
//     histograms straight out of their shared memory (DSMEM), draws (centre, sigma) for the two sides -- PHI_G lanes
//     per attribute, same arithmetic as phi_job_body -- and publishes (centre, 1/sigma) into the member CTAs' shared
//     memory.  The r updates of the merge launch state do not depend on the scans: they run while the member CTAs
//     work.
//   * phases are separated by hardware cluster barriers (~0.2 us): three per restricted scan;
//   * a member's restricted-Gibbs decision is piecewise constant in D = (log n_1 + LL_1) - (log n_2 + LL_2)
//     (smg_sm.cuh, sm_d_region); with the count term bounded over the WHOLE scan (|log-count term| <= log nS) almost
//     every member is decided by its owner without knowing the running counts; the few that are not go through the
//     ordered walk of smg_sm.cuh on CTA 0 (one extra barrier, only when there are any).
// Same draws (Philox keys, injected uniforms) and the same decisions as the other two paths.  Sums are associated
// differently (per-thread sums over the attributes instead of warp butterflies; rank partials of the log-normalisers),
// so likelihood values agree to the last few bits, not bit for bit.  The slots an accepted proposal hands to the state
// get their log-normaliser sums from the canonical 256-leaf tree.
// The kernel leaves 132 SMs free: the likelihood block of the next pass runs beside it (sweep()).
#pragma once
#include "smg_sm.cuh"

namespace smg {

#define SMC_T 512
#define SMC_WARPS (SMC_T / 32)
#define SMC_MAXCS 16

struct SmcArgs {
  SmChainArgs A;   // same fields as the cooperative kernel
  int CS;          // CTAs in the cluster: CS/2 member CTAs (ranks 0 .. CS/2-1), CS/2 parameter CTAs
  int sl;          // attributes per parameter CTA (pp / (CS/2))
  int TL;          // table entries per attribute (power of two > largest code)
  int mcap;        // member capacity per member CTA (>= ceil((n+2)/(CS/2)))
  int rcap;        // members whose row and logit are kept in shared memory, per member CTA
  double* den;     // [slots][pp] per-attribute log-normalisers
  double* gsvals;  // [6][pp] per-attribute addends of the six parameter-density terms
  double *nr_d0, *nr_lg;         // [n + 64] members whose decision depends on the running counts, by region
  int *nr_pre, *nr_z, *nr_idx;   // [n + 64]
  int* walk_out;                 // [2] side-1 change of those members (per scan parity)
  unsigned long long* prof;      // [64] optional phase cycle counters (SMG_SMC_PROFILE)
};

struct SmcLayout {
  size_t T, isgv, sdpart, sigS, denS, vS, wS, lg, phs, phl, slh, attrS, xchg, sel, rowsT, cenv, cenS, z, lst, zig, total;
  int rstr;  // member stride of the transposed rows (odd: conflict-free both by member and by word)
};
__host__ __device__ inline size_t smc_al(size_t x) { return (x + 15) & ~(size_t)15; }
__host__ __device__ inline SmcLayout smc_layout(int pp, int mmax, int TL, int CS, int mcap, int rcap) {
  SmcLayout L;
  const int sl = pp / (CS / 2);
  size_t o = 0;
  L.rstr = rcap | 1;
  L.T = o, o = smc_al(o + (size_t)pp * TL * 8);
  L.isgv = o, o = smc_al(o + (size_t)3 * pp * 8);
  L.sdpart = o, o = smc_al(o + (size_t)3 * SMC_MAXCS * 8);
  L.sigS = o, o = smc_al(o + (size_t)3 * sl * 8);
  L.denS = o, o = smc_al(o + (size_t)3 * sl * 8);
  L.vS = o, o = smc_al(o + (size_t)sl * 8);
  L.wS = o, o = smc_al(o + (size_t)sl * 8);
  L.lg = o, o = smc_al(o + (size_t)rcap * 8);
  L.phs = o, o = smc_al(o + (size_t)2 * pp * mmax * 4);
  L.phl = o, o = smc_al(o + (size_t)2 * pp * mmax * 4);
  L.slh = o, o = smc_al(o + (size_t)SH_N * sl * mmax * 4);
  L.attrS = o, o = smc_al(o + (size_t)sl * 4);
  L.xchg = o, o = smc_al(o + (size_t)2 * SMC_MAXCS * 4 * 4);
  L.sel = o, o = smc_al(o + (size_t)SMC_MAXCS * 4);
  L.rowsT = o, o = smc_al(o + (size_t)(pp / 4) * L.rstr * 4);
  L.cenv = o, o = smc_al(o + (size_t)3 * pp);
  L.cenS = o, o = smc_al(o + (size_t)3 * sl);
  L.z = o, o = smc_al(o + (size_t)2 * mcap);
  L.lst = o, o = smc_al(o + (size_t)2 * rcap);
  L.zig = o, o = smc_al(o + (size_t)(2 * SMG_ZIG_C + 1) * 8);
  L.total = o;
  return L;
}

// ---- cluster plumbing (barrier.cluster / DSMEM through mapa + st/ld.shared::cluster)
__device__ __forceinline__ void smc_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void smc_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void smc_sync() {
  smc_arrive();
  smc_wait();
}
__device__ __forceinline__ uint32_t smc_map(const void* p, int rank) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(p), r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
  return r;
}
__device__ __forceinline__ void smc_st_s32(uint32_t a, int v) {
  asm volatile("st.shared::cluster.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void smc_st_f64(uint32_t a, double v) {
  asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}
__device__ __forceinline__ void smc_st_u8(uint32_t a, int v) {
  asm volatile("st.shared::cluster.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ int smc_ld_s32(uint32_t a) {
  int v;
  asm volatile("ld.shared::cluster.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}

struct SmcJob {
  int role;   // 0: side of i_1, 1: side of i_2, 2: merged cluster
  int hist;   // SmHist row of the slice histograms
  int nk;     // members
  int dst;    // destination slot (absolute)
  int idx;    // job index inside its step (enters the Philox counter)
  int prior;  // 1: prior draw
  uint32_t sub;
  const double *uc, *us;
};

// Parameter updates for this (parameter) CTA's attribute slice: task = (job, attribute), PHI_G lanes each.
// The arithmetic of a task is phi_job_body's (smg_kernels.cuh).  Results go to the slice arrays (sigma is the
// "current sigma" of the job's next update), to the global slot `dst`, and -- for the two sides being scanned --
// into the member CTAs' parameter vectors.
__device__ __noinline__ void smc_draw_slice(const SmcArgs& G, const SmcJob* jobs, int nj, int j0s, const int* slh,
                                            const int* attrS, const double* vS, const double* wS, double* sigS,
                                            double* denS, uint8_t* cenS, uint8_t* cenv, double* isgv, bool publish,
                                            const double* zig) {
  const SmChainArgs& A = G.A;
  const int sl = G.sl, pp = A.pp, mmax = A.mmax, nM = G.CS / 2;
  const int g = threadIdx.x & (PHI_G - 1), lane = threadIdx.x & 31, gbase = lane & ~(PHI_G - 1);
  const unsigned gmask = ((1u << PHI_G) - 1u) << gbase;
  const int ntask = nj * sl;
#ifdef SMG_SMC_PROFILE
  const bool prof_me = threadIdx.x == 0 && (int)cluster_cta_rank() == nM && G.prof;
  __shared__ unsigned long long d_prof[8];  // accumulated in shared memory, flushed once per call (a global RMW per tick
                                            // would put an L2 round trip into every phase)
  if (prof_me)
    for (int k = 0; k < 8; k++) d_prof[k] = 0ull;
  long long td = clock64();
#define DRAW_TICK(k)                                                   \
  do {                                                                 \
    const long long _t = clock64();                                    \
    if (prof_me) d_prof[(k) - 50] += (unsigned long long)(_t - td);    \
    td = _t;                                                           \
  } while (0)
#else
#define DRAW_TICK(k)
#endif
  for (int task0 = 0; task0 < ntask; task0 += SMC_T / PHI_G) {
    const int task = task0 + (int)(threadIdx.x / PHI_G);
    if (task >= ntask) continue;  // whole groups drop out together
    const SmcJob J = jobs[task / sl];
    const int jl = task % sl, j = j0s + jl;
    const size_t o = (size_t)J.dst * pp + j;
    if (j >= A.p) {  // padding attributes
      if (g == 0) {
        A.cen[o] = 0;
        A.sig[o] = 1.0;
        A.isg[o] = 0.0;
        G.den[o] = 0.0;
        sigS[J.role * sl + jl] = 1.0;
        denS[J.role * sl + jl] = 0.0;
        cenS[J.role * sl + jl] = 0;
      }
      if (publish && J.role < 2)
        for (int r = g; r < nM; r += PHI_G) {
          smc_st_u8(smc_map(cenv + J.role * pp + j, r), 0);
          smc_st_f64(smc_map(isgv + J.role * pp + j, r), 0.0);
        }
      continue;
    }
    RngKey key = A.key;
    key.sub = J.sub;
    const int m = attrS[jl];
    int center = 0;
    double s_match = 0.0;
    double sigma = 0.0, isg = 0.0;
    DRAW_TICK(50);
    bool done = false;
    if (!J.prior && !J.uc && !J.us && !A.phi.sigma_exact && m <= PHI_G) {
      // ---- the common case (conditional update, Philox draws): centre by the dominance screen of draw_center_grp when it
      //      decides alone, sigma by the inline Beta-rejection draw, and the short form of the derived quantities
      const double sg = sigS[J.role * sl + jl];
      const int* h = slh + ((size_t)J.hist * sl + jl) * mmax;
      const int hraw = g < m ? h[g] : 0;
      const int hv = g < m ? hraw : -1;
      const int h1 = __reduce_max_sync(gmask, hv);
      const unsigned top = __ballot_sync(gmask, hv == h1) & gmask;
      const int h2 = __reduce_max_sync(gmask, hv == h1 ? -1 : hv);
      const bool c_ok = __popc(top) == 1 && (double)(h1 - max(h2, 0)) >= SMG_CENTER_DOM * sg && sg > 0.0;  // draw_center_grp's screen
      if (c_ok) {
        center = (__ffs(top) - 1 - gbase) + 1;
        s_match = (double)h1;
      } else {
        const double uc = get_u(nullptr, (size_t)j, key, U_CENTER, (uint32_t)J.idx, (uint32_t)j);
        center = draw_center_grp(hraw, J.nk, sg, m, uc, g, gmask, gbase, &s_match);
      }
      const double uu = hig_draw_u_grp(key, (uint32_t)J.idx, (uint32_t)j, vS[jl] + s_match, wS[jl] + (double)J.nk - s_match,
                                       (double)m, g, gmask, gbase, zig, zig + SMG_ZIG_C + 1);
      done = true;
      if (g == 0) {
        // sigma = -1/log u as everywhere; 1/sigma and the log-normaliser through u = exp(-1/sigma) itself:
        // 1/sigma = -log u, log(1 + (m-1)/exp(1/sigma)) = log1p((m-1) u) -- three divisions and an exp shorter.
        // (Equal to the canonical expressions within an ulp or two; an accepted proposal's slots are re-derived
        // canonically from sigma before they join the state.)
        const double Lg = log(uu);
        sigma = -1.0 / Lg;
        isg = -Lg;
        const double dn = log1p((double)(m - 1) * uu);
        A.cen[o] = (uint8_t)center;
        A.sig[o] = sigma;
        A.isg[o] = isg;
        G.den[o] = dn;
        sigS[J.role * sl + jl] = sigma;
        denS[J.role * sl + jl] = dn;
        cenS[J.role * sl + jl] = (uint8_t)center;
      }
    }
    DRAW_TICK(51);
    if (!done) {
    const double uc = get_u(J.uc, (size_t)j, key, U_CENTER, (uint32_t)J.idx, (uint32_t)j);
    if (J.prior) {
      center = (int)((double)m * uc + 1.0);  // sample(m_j, 1): (int)(m*u + 1)
      if (center > m) center = m;
    } else {
      const double sg = sigS[J.role * sl + jl];
      const int* h = slh + ((size_t)J.hist * sl + jl) * mmax;
      if (m <= PHI_G) {
        center = draw_center_grp(g < m ? h[g] : 0, J.nk, sg, m, uc, g, gmask, gbase, &s_match);
      } else {
        center = 0;
        if (g == 0) {
          center = draw_center(h, J.nk, sg, m, uc);
          s_match = (double)h[center - 1];
        }
        center = __shfl_sync(gmask, center, gbase);
        s_match = __shfl_sync(gmask, s_match, gbase);
      }
    }
    const double vv = vS[jl] + s_match;
    const double ww = wS[jl] + (double)J.nk - s_match;
    double uu = 0.5;
    DRAW_TICK(52);
    if (J.us || A.phi.sigma_exact) {
      if (g == 0) {
        const double us = get_u(J.us, (size_t)j, key, U_SIGMA, (uint32_t)J.idx, (uint32_t)j);
        uu = hig_inv_u_d(us, vv, ww, (double)m);
      }
    } else {
      uu = hig_draw_u_grp(key, (uint32_t)J.idx, (uint32_t)j, vv, ww, (double)m, g, gmask, gbase, zig, zig + SMG_ZIG_C + 1);
    }
    DRAW_TICK(53);
    if (g == 0) {
      sigma = -1.0 / log(uu);
      isg = 1.0 / sigma;
      const double dn = hamming_den(sigma, m);
      A.cen[o] = (uint8_t)center;
      A.sig[o] = sigma;
      A.isg[o] = isg;
      G.den[o] = dn;
      sigS[J.role * sl + jl] = sigma;
      denS[J.role * sl + jl] = dn;
      cenS[J.role * sl + jl] = (uint8_t)center;
    }
    }
    DRAW_TICK(54);
    if (publish && J.role < 2) {
      isg = __shfl_sync(gmask, isg, gbase);
      for (int r = g; r < nM; r += PHI_G) {
        smc_st_u8(smc_map(cenv + J.role * pp + j, r), center);
        smc_st_f64(smc_map(isgv + J.role * pp + j, r), isg);
      }
    }
    DRAW_TICK(55);
  }
#ifdef SMG_SMC_PROFILE
  if (prof_me)
    for (int k = 0; k < 6; k++) G.prof[50 + k] += d_prof[k];
#endif
}

// after the draws of a step: this parameter CTA's partial sums of the log-normalisers of the two scanned sides, to
// every member CTA (k = index of this CTA among the parameter CTAs)
__device__ __forceinline__ void smc_publish_sdpart(const SmcArgs& G, int k, const double* denS, double* sdpart) {
  const int sl = G.sl, nM = G.CS / 2;
  const int t = threadIdx.x;
  if (t < 2 * nM) {
    const int role = t / nM, r = t % nM;
    double acc = 0.0;
    for (int jl = 0; jl < sl; jl++) acc += denS[role * sl + jl];
    smc_st_f64(smc_map(sdpart + role * SMC_MAXCS + k, r), acc);
  }
}

#ifdef SMG_SMC_PROFILE
// member-side phases are timed by thread 0 of CTA 0, parameter-side phases by thread 0 of the first parameter CTA
#define SMC_TICK(k)                                                                          \
  do {                                                                                       \
    const long long _t = clock64();                                                          \
    if (threadIdx.x == 0) s_prof[k] += (unsigned long long)(_t - tk);                        \
    tk = _t;                                                                                 \
  } while (0)
#else
#define SMC_TICK(k)
#endif

__global__ void __launch_bounds__(SMC_T, 1) sm_cluster_kernel(SmcArgs G) {
  const SmChainArgs& A = G.A;
  extern __shared__ __align__(16) unsigned char smc_raw[];
  const int n = A.n, p = A.p, pp = A.pp, mmax = A.mmax, NSB = A.NS;
  const int CS = G.CS, nM = CS / 2, nP = CS / 2, sl = G.sl, TL = G.TL, mcap = G.mcap, rcap = G.rcap;
  const SmcLayout L = smc_layout(pp, mmax, TL, CS, mcap, rcap);
  const int RSTR = L.rstr;
  double* T = reinterpret_cast<double*>(smc_raw + L.T);         // [pp][TL] look-up table of the scan in progress
  double* isgv = reinterpret_cast<double*>(smc_raw + L.isgv);   // [3][pp] 1/sigma of the vectors being evaluated
  double* sdpart = reinterpret_cast<double*>(smc_raw + L.sdpart);  // [3][16] partial log-normaliser sums by parameter CTA
  double* sigS = reinterpret_cast<double*>(smc_raw + L.sigS);   // [3][sl] current sigma of the three update chains
  double* denS = reinterpret_cast<double*>(smc_raw + L.denS);   // [3][sl]
  double* vS = reinterpret_cast<double*>(smc_raw + L.vS);
  double* wS = reinterpret_cast<double*>(smc_raw + L.wS);
  double* lg = reinterpret_cast<double*>(smc_raw + L.lg);       // [rcap] logit(u) of the cached members, current scan
  int* phs = reinterpret_cast<int*>(smc_raw + L.phs);           // [2][pp][mmax] own members by current-state side
  int* phl = reinterpret_cast<int*>(smc_raw + L.phl);           // [2][pp][mmax] own members by launch / proposal side
  int* slh = reinterpret_cast<int*>(smc_raw + L.slh);           // [SH_N][sl][mmax] histograms of this CTA's attributes
  int* attrS = reinterpret_cast<int*>(smc_raw + L.attrS);
  int* xchg = reinterpret_cast<int*>(smc_raw + L.xchg);         // [2][16][4] per-CTA counters of a step, from the member CTAs
  int* sel = reinterpret_cast<int*>(smc_raw + L.sel);           // [16] members found by each CTA
  uint32_t* rowsT = reinterpret_cast<uint32_t*>(smc_raw + L.rowsT);  // [pp/4][RSTR] rows of the cached members, word-major
  uint8_t* cenv = smc_raw + L.cenv;                             // [3][pp]
  uint8_t* cenS = smc_raw + L.cenS;                             // [3][sl]
  uint8_t* zc = smc_raw + L.z;                                  // [mcap] side (launch, then proposal)
  uint8_t* zn = zc + mcap;                                      // [mcap] side decided by the scan in progress
  double* zigS = reinterpret_cast<double*>(smc_raw + L.zig);    // ziggurat tables: x[129] then r[128]
  unsigned short* lst = reinterpret_cast<unsigned short*>(smc_raw + L.lst);  // [rcap] cached members that need the table
  __shared__ int s_nlist;
  __shared__ double sh[256];
  __shared__ int s_sel[8];
  __shared__ int s_wa[SMC_WARPS], s_wb[SMC_WARPS], s_tot[4];
  __shared__ double s_sd[4];

  const int rank = (int)cluster_cta_rank();
  const bool isM = rank < nM;
  const int kP = rank - nM;  // index among the parameter CTAs
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j0s = (isM ? rank : kP) * sl;  // attribute slice of this CTA's parameter updates (member CTAs: the merged cluster's)
  const size_t len = (size_t)pp * mmax;
  const int words = pp / 4;
  auto off = [](const double* base, size_t o) -> const double* { return base ? base + o : nullptr; };
#ifdef SMG_SMC_PROFILE
  __shared__ unsigned long long s_prof[32];  // phase cycle counters of this CTA's thread 0, flushed once at the end
  if (tid == 0)
    for (int k = 0; k < 32; k++) s_prof[k] = 0ull;
  long long tk = clock64();
#endif
  smc_arrive();  // (0) every CTA of the cluster is running: its shared memory may be written remotely after the wait

  // ---- pair (split_merge.cpp:275) and this CTA's share of the row scan that builds S (:280-301)
  if (tid == 0) {
    RngKey k = A.key;
    k.sub = SUB_SM_SELECT;
    int i1, i2;
    sm_pick_pair(n, A.u_pair, k, &i1, &i2, A.pair_det);
    s_sel[0] = i1;
    s_sel[1] = i2;
    s_sel[2] = A.c[i1];
    s_sel[3] = A.c[i2];
  }
  for (int jl = tid; jl < sl; jl += SMC_T) {
      attrS[jl] = A.phi.attr[j0s + jl];
      vS[jl] = A.phi.v[j0s + jl];
      wS[jl] = A.phi.w[j0s + jl];
    }
  for (int q = tid; q < 2 * SMG_ZIG_C + 1; q += SMC_T) zigS[q] = q <= SMG_ZIG_C ? g_zig_x[q] : g_zig_r[q - SMG_ZIG_C - 1];
  if (rank == 0)
    for (int q = tid; q < 24; q += SMC_T) A.terms[q] = 0.0;
  __syncthreads();
  const int i1 = s_sel[0], i2 = s_sel[1], cA = s_sel[2], cB = s_sel[3];
  const int same = (cA == cB);
  const int gwarp = rank * SMC_WARPS + warp, nwarps = CS * SMC_WARPS;
  const int seg = ((n + nwarps - 1) / nwarps + 31) & ~31;
  const int lo = (int)min((long long)n, (long long)gwarp * seg), hi = min(n, lo + seg);
  {
    int cntm = 0;
    for (int i = lo + lane; i < hi; i += 32) {
      const int ci = A.c[i];
      cntm += (i != i1 && i != i2 && (ci == cA || ci == cB));
    }
    cntm = warp_sum_i(cntm);
    if (lane == 0) s_wa[warp] = cntm;
  }
  __syncthreads();
  smc_wait();  // (0)
  if (warp == 0) {
    const int v = lane < SMC_WARPS ? s_wa[lane] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(SMG_FULL, x, o);
      if (lane >= o) x += y;
    }
    if (lane < SMC_WARPS) s_wb[lane] = x - v;
    const int tot = __shfl_sync(SMG_FULL, x, 31);
    if (lane < CS) smc_st_s32(smc_map(sel + rank, lane), tot);
  }
  // prior parameters of the three launch clusters (split_merge.cpp:331-343, :379-380): the two split-launch clusters on
  // the parameter CTAs here, the merge-launch cluster on the member CTAs (which own its update chain) before the scans
  auto prior_job = [&](int k) {
    SmcJob J;
    J.role = k;
    J.hist = 0;
    J.nk = 0;
    J.dst = NSB + (k == 0 ? SM_SL_A : (k == 1 ? SM_SL_B : SM_ML_M));
    J.idx = k;
    J.prior = 1;
    J.sub = SUB_SM_PRIOR;
    J.uc = off(A.u_prior_c, (size_t)k * p);
    J.us = off(A.u_prior_s, (size_t)k * p);
    return J;
  };
  if (!isM) {
    SmcJob jb[2];
    for (int k = 0; k < 2; k++) jb[k] = prior_job(k);
    smc_draw_slice(G, jb, 2, j0s, slh, attrS, vS, wS, sigS, denS, cenS, cenv, isgv, true, zigS);
    __syncthreads();
    smc_publish_sdpart(G, kP, denS, sdpart);
  }
  SMC_TICK(0);
  smc_sync();  // (a) member counts, launch parameters
  SMC_TICK(1);
  int nS = 0, before = 0;
  for (int r = 0; r < CS; r++) {
    const int v = sel[r];
    nS += v;
    if (r < rank) before += v;
  }
  {
    int offp = before + s_wb[warp];
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
  if (rank == 0 && tid == 0) sm_fill_info_plan(A.info, A.plan, NSB, i1, i2, nS, cA, cB, *A.Kptr);
  SMC_TICK(2);
  smc_sync();  // (b) S complete
  SMC_TICK(3);
  // ---- member CTAs: positions [pos0, pos0 + mine); the anchors are the positions nS (side 0) and nS + 1 (side 1)
  const int per = (nS + 2 + nM - 1) / nM;
  const int pos0 = isM ? rank * per : 0;
  const int mine = isM ? max(0, min(per, nS + 2 - pos0)) : 0;
  const int ndec = max(0, min(mine, nS - pos0));  // own members that are re-allocated (not anchors)
  // row of position `pos`
  auto row_of = [&](int pos) -> int { return pos < nS ? A.S[pos] : (pos == nS ? i1 : i2); };
  // word w of own member m
  auto xword = [&](int m, int w) -> uint32_t {
    return m < rcap ? rowsT[(size_t)w * RSTR + m]
                    : *reinterpret_cast<const uint32_t*>(A.X + (size_t)row_of(pos0 + m) * pp + 4 * w);
  };
  // moves member m (whole warp): -1 in histogram (ho, go), +1 in (hn, gn); ho == nullptr: only the addition
  auto hist_move = [&](int m, int* ho, int go, int* hn, int gn) {
    for (int w = lane; w < words; w += 32) {
      const uint32_t xw = xword(m, w);
#pragma unroll
      for (int b = 0; b < 4; b++) {
        const int x = (xw >> (8 * b)) & 0xff;
        if (x) {
          const size_t e = (size_t)(4 * w + b) * mmax + (x - 1);
          if (ho) atomicSub(&ho[(size_t)go * len + e], 1);
          atomicAdd(&hn[(size_t)gn * len + e], 1);
        }
      }
    }
  };
  int nBcur = 0;   // side-1 members of the launch / proposal allocation, anchor i_2 included
  int cntS1 = 0;
  if (isM) {
    RngKey k = A.key;
    k.sub = SUB_SM_LAUNCH;
    for (int m = tid; m < mine; m += SMC_T) {
      const int pos = pos0 + m;
      int zlaunch;
      if (pos < nS) {
        const double u = get_u(A.u_launch, pos, k, U_SM_LAUNCH, (uint32_t)pos, 0u);  // split_merge.cpp:346
        const int z = (int)(2.0 * u);
        zlaunch = z > 1 ? 1 : z;
      } else {
        zlaunch = pos - nS;
      }
      zc[m] = (uint8_t)zlaunch;
      zn[m] = (uint8_t)zlaunch;
    }
    for (int q = tid; q < 2 * (int)len; q += SMC_T) phs[q] = phl[q] = 0;
    {  // rows of the cached members into shared memory, word-major
      const int ncache = min(mine, rcap), chunks = pp / 16;
      for (int t = tid; t < ncache * chunks; t += SMC_T) {
        const int m = t / chunks, q = t % chunks;
        const uint4 v = *reinterpret_cast<const uint4*>(A.X + (size_t)row_of(pos0 + m) * pp + q * 16);
        rowsT[(size_t)(4 * q + 0) * RSTR + m] = v.x;
        rowsT[(size_t)(4 * q + 1) * RSTR + m] = v.y;
        rowsT[(size_t)(4 * q + 2) * RSTR + m] = v.z;
        rowsT[(size_t)(4 * q + 3) * RSTR + m] = v.w;
      }
    }
    __syncthreads();
    // both side histograms of the own members (current-state sides: fixed; launch sides: maintained by the scans).
    // Cached rows with at most 8 levels: a thread owns one 4-attribute word of the rows of every (SMC_T / words)-th
    // member and counts in registers -- for level a the four byte lanes of (x == a) are four packed 8-bit counters per
    // (histogram, side) -- then adds its counters to the shared-memory tables once.  (One shared-memory atomic per
    // attribute and member took 60 us here.)  Other rows: the atomic path.
    int c1s = 0, c1l = 0;
    const int ncache = min(mine, rcap);
    for (int m = tid; m < mine; m += SMC_T) {
      const int pos = pos0 + m;
      const int gs = pos < nS ? A.zState[pos] : pos - nS;
      zn[m] = (uint8_t)gs;  // staged for the counting loop below, restored after it
      c1s += gs;
      c1l += zc[m];
    }
    c1s = warp_sum_i(c1s);
    c1l = warp_sum_i(c1l);
    __syncthreads();
    const bool packed = mmax <= 8;
    if (packed) {
      const int ng = words <= SMC_T ? SMC_T / words : 1, g = words <= SMC_T ? tid / words : 0;
      for (int w = (words <= SMC_T ? tid % words : tid); w < words && g < ng; w += SMC_T) {
        for (int mbase = g; mbase < ncache; mbase += ng * 255) {
          uint32_t cnt[2][2][8];  // [histogram: state, launch][side][level]
#pragma unroll
          for (int h = 0; h < 2; h++)
#pragma unroll
            for (int sd = 0; sd < 2; sd++)
#pragma unroll
              for (int a = 0; a < 8; a++) cnt[h][sd][a] = 0u;
          const int mend = min(ncache, mbase + ng * 255);
          for (int m = mbase; m < mend; m += ng) {
            const uint32_t xw = rowsT[(size_t)w * RSTR + m];
            const uint32_t ms = zn[m] ? 0xffffffffu : 0u, ml = zc[m] ? 0xffffffffu : 0u;
#pragma unroll
            for (int a = 0; a < 8; a++) {
              if (a < mmax) {
                const uint32_t eq = __vcmpeq4(xw, 0x01010101u * (uint32_t)(a + 1)) & 0x01010101u;
                cnt[0][0][a] += eq & ~ms;
                cnt[0][1][a] += eq & ms;
                cnt[1][0][a] += eq & ~ml;
                cnt[1][1][a] += eq & ml;
              }
            }
          }
#pragma unroll
          for (int h = 0; h < 2; h++)
#pragma unroll
            for (int sd = 0; sd < 2; sd++)
#pragma unroll
              for (int a = 0; a < 8; a++) {
                if (a < mmax) {
                  const uint32_t v = cnt[h][sd][a];
                  int* dst = (h ? phl : phs) + (size_t)sd * len + (size_t)(4 * w) * mmax + a;
#pragma unroll
                  for (int b = 0; b < 4; b++) {
                    const int c = (int)((v >> (8 * b)) & 0xffu);
                    if (c) atomicAdd(dst + (size_t)b * mmax, c);
                  }
                }
              }
        }
      }
    }
    for (int m = (packed ? ncache : 0) + warp; m < mine; m += SMC_WARPS) {
      hist_move(m, nullptr, 0, phs, zn[m]);
      hist_move(m, nullptr, 0, phl, zc[m]);
    }
    __syncthreads();
    for (int m = tid; m < mine; m += SMC_T) zn[m] = zc[m];
    if (lane == 0) {
      s_wa[warp] = c1s;
      s_wb[warp] = c1l;
    }
    __syncthreads();
    if (warp == 0) {
      const int a = warp_sum_i(lane < SMC_WARPS ? s_wa[lane] : 0), b = warp_sum_i(lane < SMC_WARPS ? s_wb[lane] : 0);
      if (lane < CS) {
        smc_st_s32(smc_map(xchg + (0 * SMC_MAXCS + rank) * 4 + 0, lane), a);
        smc_st_s32(smc_map(xchg + (0 * SMC_MAXCS + rank) * 4 + 1, lane), b);
      }
    }
  }
  SMC_TICK(4);
  smc_sync();  // (c) side histograms and side counts of the member CTAs
  SMC_TICK(5);
  for (int r = 0; r < nM; r++) {
    cntS1 += xchg[(0 * SMC_MAXCS + r) * 4 + 0];
    nBcur += xchg[(0 * SMC_MAXCS + r) * 4 + 1];
  }
  // slice histograms: sum of the member CTAs' partial histograms over this parameter CTA's attributes (DSMEM reads)
  auto reduce_slices = [&](const int* ph, int h0, bool to_global) {
    const int per_side = sl * mmax;
    for (int t = tid; t < 2 * per_side; t += SMC_T) {
      const int s = t / per_side, e = t % per_side;
      const int* src = ph + (size_t)s * len + (size_t)j0s * mmax + e;
      int v[SMC_MAXCS / 2];  // all remote loads in flight together
#pragma unroll
      for (int r = 0; r < SMC_MAXCS / 2; r++) v[r] = r < nM ? smc_ld_s32(smc_map(src, r)) : 0;
      int acc = 0;
#pragma unroll
      for (int r = 0; r < SMC_MAXCS / 2; r++) acc += v[r];
      slh[(size_t)(h0 + s) * per_side + e] = acc;
      if (to_global) A.H[(size_t)(h0 + s) * len + (size_t)j0s * mmax + e] = acc;
    }
  };
  if (!isM) {
    reduce_slices(phs, SH_S0, true);
    __syncthreads();
    for (int e = tid; e < sl * mmax; e += SMC_T) {
      const int v = slh[(size_t)SH_S0 * sl * mmax + e] + slh[(size_t)SH_S1 * sl * mmax + e];
      slh[(size_t)SH_M * sl * mmax + e] = v;
      A.H[(size_t)SH_M * len + (size_t)j0s * mmax + e] = v;
    }
    __syncthreads();
  }
  if (isM) {
    // the merged cluster's histogram on this CTA's attribute slice (both current-state sides of every member CTA, DSMEM
    // reads; fixed for the whole proposal), then the prior draw of the merge-launch cluster: its r updates are an
    // independent chain (split_merge.cpp:386-387) that lives on the member CTAs, next to the side updates on the
    // parameter CTAs
    const int per_side = sl * mmax;
    for (int e = tid; e < per_side; e += SMC_T) {
      int v[SMC_MAXCS];
#pragma unroll
      for (int r = 0; r < SMC_MAXCS / 2; r++) {
        v[2 * r] = r < nM ? smc_ld_s32(smc_map(phs + (size_t)j0s * mmax + e, r)) : 0;
        v[2 * r + 1] = r < nM ? smc_ld_s32(smc_map(phs + len + (size_t)j0s * mmax + e, r)) : 0;
      }
      int acc = 0;
#pragma unroll
      for (int r = 0; r < SMC_MAXCS; r++) acc += v[r];
      slh[(size_t)SH_M * per_side + e] = acc;
    }
    __syncthreads();
    const SmcJob J = prior_job(2);
    smc_draw_slice(G, &J, 1, j0s, slh, attrS, vS, wS, sigS, denS, cenS, cenv, isgv, false, zigS);
  }
  if (rank == 0 && tid == 0) {
    A.cnt[SH_S0] = nS + 2 - cntS1;
    A.cnt[SH_S1] = cntS1;
    A.cnt[SH_M] = nS + 2;
  }
  // member CTAs: logit(u) of the own members for scan q (it does not depend on the state)
  auto fill_logits = [&](int q) {
    RngKey k = A.key;
    k.sub = SUB_SM_RG + q;
    const double* ui = off(A.u_rg, (size_t)q * n);
    for (int m = tid; m < ndec; m += SMC_T) {
      const double v = sm_logit_u(ui, pos0 + m, k);
      if (m < rcap)
        lg[m] = v;
      else
        A.lgt[pos0 + m] = v;
    }
  };
  // member CTAs: constants of the two sides' parameters (cenv / isgv rows 0, 1) -- the difference of their log-normaliser
  // sums and the range of 1/sigma on each side -- and, on demand, the look-up table T[j][a]
  double sdBA = 0.0;  // sden_B - sden_A
  double wAmin = 0.0, wAmax = 0.0, wBmin = 0.0, wBmax = 0.0;
  auto side_consts = [&]() {
    const double a = lane < nP ? sdpart[0 * SMC_MAXCS + lane] : 0.0, b = lane < nP ? sdpart[1 * SMC_MAXCS + lane] : 0.0;
    sdBA = warp_sum(b) - warp_sum(a);
    double amin = 1e300, amax = 0.0, bmin = 1e300, bmax = 0.0;
    for (int j = lane; j < p; j += 32) {
      const double wa = isgv[j], wb = isgv[pp + j];
      amin = fmin(amin, wa), amax = fmax(amax, wa);
      bmin = fmin(bmin, wb), bmax = fmax(bmax, wb);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      amin = fmin(amin, shfl_xor_d(amin, o)), amax = fmax(amax, shfl_xor_d(amax, o));
      bmin = fmin(bmin, shfl_xor_d(bmin, o)), bmax = fmax(bmax, shfl_xor_d(bmax, o));
    }
    wAmin = amin, wAmax = amax, wBmin = bmin, wBmax = bmax;
  };
  auto build_table = [&]() {
    const int tls = 31 - __clz(TL);  // log2(TL)
    for (int e = tid; e < pp * TL; e += SMC_T) {
      const int j = e >> tls, av = e & (TL - 1);
      double v = 0.0;
      if (av != 0) {
        if (av != cenv[pp + j]) v += isgv[pp + j];
        if (av != cenv[j]) v -= isgv[j];
      }
      T[e] = v;
    }
  };
  if (isM) {
    if (A.t > 0) fill_logits(0);
    side_consts();
    if (tid == 0) s_nlist = 0;
    __syncthreads();
  }
  SMC_TICK(6);

  // ------------------------------------------------------------------------------------------
  // launch scans (+ merge-launch updates), then the proposal
  // ------------------------------------------------------------------------------------------
  const int nsteps = A.t > A.r ? A.t : A.r;
  int cntL0 = 0, cntL1 = 0;
  int par = 1;
  for (int it = 0; it <= nsteps; it++) {
    const bool prop = (it == nsteps);
    if (prop) {
      // launch state complete: keep its sides and counts, start the proposal from a copy of it (split_merge.cpp:575-577)
      cntL1 = nBcur;
      cntL0 = nS + 2 - nBcur;
      if (isM) {
        for (int m = tid; m < ndec; m += SMC_T) A.zL[pos0 + m] = zc[m];
      } else {
        for (int t = tid; t < 2 * sl; t += SMC_T) {
          const int role = t / sl, jl = t % sl, j = j0s + jl;
          const size_t d = (size_t)(NSB + (role ? SM_ST_B : SM_ST_A)) * pp + j;
          const double sg = sigS[role * sl + jl];
          A.cen[d] = cenS[role * sl + jl];
          A.sig[d] = sg;
          A.isg[d] = (j < p) ? 1.0 / sg : 0.0;
          G.den[d] = denS[role * sl + jl];
        }
      }
      if (rank == 0 && tid == 0) {
        A.cnt[SH_L0] = cntL0;
        A.cnt[SH_L1] = cntL1;
      }
    }
    const int q = prop ? A.t : it;
    const bool do_scan = prop ? (same != 0) : (it < A.t);
    const bool do_mg = prop || it < A.r;
    // the update of the merged cluster that belongs to this step (parameter CTAs): the r updates of the merge launch
    // state (split_merge.cpp:386-387) are an independent chain on the fixed merged histogram and run while the member
    // CTAs evaluate the scan; the final one of the proposal (:584) keeps job index 2 as in the other paths
    auto mg_job = [&]() {
      SmcJob J;
      J.role = 2;
      J.hist = SH_M;
      J.nk = nS + 2;
      J.dst = NSB + (prop ? SM_ST_M : SM_ML_M);
      J.idx = prop ? 2 : (do_scan ? 2 : 0);
      J.prior = 0;
      J.sub = SUB_SM_MERGE + (prop ? A.r : it);
      J.uc = off(A.u_mg_c, (size_t)(prop ? A.r : it) * p);
      J.us = off(A.u_mg_s, (size_t)(prop ? A.r : it) * p);
      smc_draw_slice(G, &J, 1, j0s, slh, attrS, vS, wS, sigS, denS, cenS, cenv, isgv, false, zigS);
    };
    if (!do_scan) {
      if (isM && do_mg) mg_job();
      continue;
    }
    int dsum = 0, nnr_tot = 0, nnr_mine = 0, extra = 0, cta_changed = 0, any_changed = 0;
    if (isM) {
      // ---- P1: likelihood differences by table look-up (one thread per member), count-free decisions
      const double dc_max = nS > 0 ? sm_dc_bound(nS, 1) + SM_DC_MARGIN : 0.0;
      const double dc_min = nS > 0 ? sm_dc_bound(nS, nS) - SM_DC_MARGIN : 0.0;
      const unsigned Tb = (unsigned)__cvta_generic_to_shared(T);
      const int tsh = 31 - __clz(TL) + 3;  // log2(TL * 8): byte stride of an attribute's table row
      // (a) screen: with mA / mB mismatches against the two centres,
      //       mB min(1/sigma_B) - mA max(1/sigma_A)  <=  LL_A - LL_B - (sden_B - sden_A)  <=  mB max(1/sigma_B) - mA min(1/sigma_A);
      //     a member whose decision is the same at both ends of that range (and of the count term's) is decided from two
      //     integer counts -- once the launch state has settled that is every member.  The others are listed.
      const uint32_t* cwA = reinterpret_cast<const uint32_t*>(cenv);
      const uint32_t* cwB = reinterpret_cast<const uint32_t*>(cenv + pp);
      const int nscr = min(ndec, rcap);
      for (int m0 = 0; m0 < nscr; m0 += SMC_T) {
        const int m = m0 + tid;
        bool listed = false;
        if (m < nscr) {
          const uint32_t* xr = rowsT + m;
          int cA4 = 0, cB4 = 0;
#pragma unroll 8
          for (int w = 0; w < words; w++) {
            const uint32_t xw = xr[(size_t)w * RSTR];
            cA4 += __popc(__vcmpne4(xw, cwA[w]));
            cB4 += __popc(__vcmpne4(xw, cwB[w]));
          }
          const double mA = (double)(cA4 >> 3), mB = (double)(cB4 >> 3);
          const double slack = 1e-9 * (mA * wAmax + mB * wBmax) + 1e-9;
          const double d_lo = mB * wBmin - mA * wAmax + sdBA - slack, d_hi = mB * wBmax - mA * wAmin + sdBA + slack;
          const double lgm = lg[m];
          const int rlo = sm_d_region(dc_min + d_lo, lgm), rhi = sm_d_region(dc_max + d_hi, lgm);
          if (rlo == rhi)
            zn[m] = (uint8_t)(~rlo & 1);
          else
            listed = true;
        }
        const unsigned bal = __ballot_sync(SMG_FULL, listed);
        if (bal) {
          int base = 0;
          if (lane == 0) base = atomicAdd(&s_nlist, __popc(bal));
          base = __shfl_sync(SMG_FULL, base, 0);
          if (listed) lst[base + __popc(bal & ((1u << lane) - 1))] = (unsigned short)m;
        }
      }
      __syncthreads();
      const int nlist = s_nlist;
      const bool need_table = nlist > 0 || ndec > rcap;
#ifdef SMG_SMC_PROFILE
      if (tid == 0) s_prof[28] += (unsigned long long)nlist;  // members of CTA 0 the count screen did not decide
#endif
      if (need_table) build_table();
      __syncthreads();
      if (tid == 0) s_nlist = 0;
      // (b) the listed members and the ones whose rows are not cached: LL_A - LL_B by table look-up
      for (int k = tid; k < nlist + max(0, ndec - rcap); k += SMC_T) {
        const int m = k < nlist ? (int)lst[k] : rcap + (k - nlist);
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        if (m < rcap) {
          const uint32_t* xr = rowsT + m;
#pragma unroll 4
          for (int w = 0; w < words; w++) {
            const uint32_t xw = xr[(size_t)w * RSTR];
#pragma unroll
            for (int b4 = 0; b4 < 4; b4++) {
              const unsigned addr = Tb + ((unsigned)(4 * w + b4) << tsh) + (((xw >> (8 * b4)) & 0xffu) << 3);
              double v;
              asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
              acc[b4] += v;
            }
          }
        } else {
          const uint32_t* xr = reinterpret_cast<const uint32_t*>(A.X + (size_t)A.S[pos0 + m] * pp);
          for (int w = 0; w < words; w++) {
            const uint32_t xw = xr[w];
#pragma unroll
            for (int b4 = 0; b4 < 4; b4++) acc[b4] += T[(size_t)(4 * w + b4) * TL + ((xw >> (8 * b4)) & 0xffu)];
          }
        }
        const double d0 = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + sdBA;  // LL_A - LL_B
        const double lgm = m < rcap ? lg[m] : A.lgt[pos0 + m];
        const int rlo = sm_d_region(dc_min + d0, lgm), rhi = sm_d_region(dc_max + d0, lgm);
        if (rlo == rhi) {
          zn[m] = (uint8_t)(~rlo & 1);
        } else {
          zn[m] = 2;
          A.dl[pos0 + m] = d0;
        }
      }
      __syncthreads();
      SMC_TICK(7);
      // ordered pass over the own members: side-1 change of the decided ones, list of the undecided ones (by position)
      int carry_d = 0, carry_n = 0;
      int chg = 0;
      for (int m = tid; m < ndec; m += SMC_T) chg |= (zn[m] != zc[m]);
      cta_changed = __syncthreads_or(chg);
      for (int base = 0; cta_changed && base < ndec; base += SMC_T) {
        const int m = base + tid;
        const bool valid = m < ndec;
        const int znm = valid ? zn[m] : 0, zcm = valid ? zc[m] : 0;
        const int isnr = (valid && znm == 2) ? 1 : 0;
        const int delta = (valid && !isnr) ? znm - zcm : 0;
        int xd = delta, xn = isnr;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int yd = __shfl_up_sync(SMG_FULL, xd, o), yn = __shfl_up_sync(SMG_FULL, xn, o);
          if (lane >= o) {
            xd += yd;
            xn += yn;
          }
        }
        if (lane == 31) {
          s_wa[warp] = xd;
          s_wb[warp] = xn;
        }
        __syncthreads();
        if (warp == 0) {
          const int a = lane < SMC_WARPS ? s_wa[lane] : 0, b = lane < SMC_WARPS ? s_wb[lane] : 0;
          int ya = a, yb = b;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int ta = __shfl_up_sync(SMG_FULL, ya, o), tb = __shfl_up_sync(SMG_FULL, yb, o);
            if (lane >= o) {
              ya += ta;
              yb += tb;
            }
          }
          if (lane < SMC_WARPS) {
            s_wa[lane] = ya - a;
            s_wb[lane] = yb - b;
          }
          if (lane == 31) {
            s_tot[0] = ya;
            s_tot[1] = yb;
          }
        }
        __syncthreads();
        if (isnr) {
          const int k = carry_n + s_wb[warp] + xn - 1;
          const int Lx = pos0 + k;
          G.nr_idx[Lx] = m;
          G.nr_pre[Lx] = carry_d + s_wa[warp] + xd - delta;
          G.nr_z[Lx] = zcm;
          G.nr_d0[Lx] = A.dl[pos0 + m];
          G.nr_lg[Lx] = m < rcap ? lg[m] : A.lgt[pos0 + m];
        }
        carry_d += s_tot[0];
        carry_n += s_tot[1];
        __syncthreads();
      }
      if (tid < CS) {
        smc_st_s32(smc_map(xchg + (par * SMC_MAXCS + rank) * 4 + 0, tid), carry_d);
        smc_st_s32(smc_map(xchg + (par * SMC_MAXCS + rank) * 4 + 1, tid), carry_n);
        smc_st_s32(smc_map(xchg + (par * SMC_MAXCS + rank) * 4 + 2, tid), cta_changed);
      }
      SMC_TICK(8);
      smc_sync();  // (1) decisions that do not depend on the counts
      SMC_TICK(9);
    } else {
      SMC_TICK(7);
      smc_sync();  // (1)
      SMC_TICK(9);
    }
    for (int r = 0; r < nM; r++) {
      dsum += xchg[(par * SMC_MAXCS + r) * 4 + 0];
      const int v = xchg[(par * SMC_MAXCS + r) * 4 + 1];
      nnr_tot += v;
      if (r == rank) nnr_mine = v;
      any_changed |= xchg[(par * SMC_MAXCS + r) * 4 + 2];
    }
    if (nnr_tot > 0) {
      // ---- the members whose decision depends on the running counts: ordered walk on CTA 0 (split_merge.cpp:186-216
      //      one at a time; here 32 at a time with the count term bounded inside the batch, see sm_rdecide_body)
      if (rank == 0 && warp == 0) {
        int regionbase = nBcur;
        for (int r = 0; r < nM; r++) {
          const int nnr = xchg[(par * SMC_MAXCS + r) * 4 + 1];
          const int lbase = r * per;
          for (int b0 = 0; b0 < nnr; b0 += 32) {
            const int qq = b0 + lane;
            const bool v = qq < nnr;
            const double md0 = v ? __ldcg(&G.nr_d0[lbase + qq]) : 0.0, mlg = v ? __ldcg(&G.nr_lg[lbase + qq]) : 0.0;
            const int mpre = v ? __ldcg(&G.nr_pre[lbase + qq]) : 0;
            int mz = v ? __ldcg(&G.nr_z[lbase + qq]) : 0;
            int start = 0;
            while (start < 32 && b0 + start < nnr) {
              const bool act = v && lane >= start;
              const int nBq = regionbase + mpre + extra, b = nBq - mz, dev = lane - start;
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
                  const int nBf = regionbase + mpre + extra, nAf = nS + 2 - nBf;
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
            if (v) G.nr_z[lbase + qq] = mz;
          }
          regionbase += xchg[(par * SMC_MAXCS + r) * 4 + 0];
        }
        if (lane == 0) G.walk_out[par] = extra;
      }
      smc_sync();  // (1b) their sides
      extra = __ldcg(&G.walk_out[par]);
      if (isM) {
        for (int k = tid; k < nnr_mine; k += SMC_T) zn[G.nr_idx[pos0 + k]] = (uint8_t)__ldcg(&G.nr_z[pos0 + k]);
        __syncthreads();
      }
    }
    nBcur += dsum + extra;
    SMC_TICK(10);
#ifdef SMG_SMC_PROFILE
    if (rank == 0 && tid == 0 && G.prof) {
      G.prof[59] += (unsigned long long)nnr_tot;
      if (q < 4) G.prof[60 + q] += (unsigned long long)nnr_tot;
      s_prof[25] += nnr_tot > 0;     // scans that needed the ordered walk
      s_prof[26] += any_changed;     // scans in which some member changed side (or was undecided)
      s_prof[27] += 1ull;            // scans
    }
#endif
    if (isM) {
      // ---- P2: move the rows of the members that changed side between the two side histograms (a warp per mover)
      for (int base = warp * 32; cta_changed && base < ndec; base += SMC_T) {
        const int m = base + lane;
        const bool mv = m < ndec && zn[m] != zc[m];
        unsigned bal = __ballot_sync(SMG_FULL, mv);
        while (bal) {
          const int l = __ffs(bal) - 1;
          bal &= bal - 1;
          const int mm = base + l;
          hist_move(mm, phl, zc[mm], phl, zn[mm]);
        }
        __syncwarp();
        if (mv) zc[m] = zn[m];
      }
      SMC_TICK(11);
      smc_sync();  // (2) side histograms of the member CTAs
      SMC_TICK(12);
      smc_arrive();  // (3)
      if (q < A.t) fill_logits(q + 1);  // the next scan's, while the parameter CTAs draw (the proposal scan is scan t)
      SMC_TICK(13);
      if (do_mg) mg_job();  // the merged cluster's update of this step, beside the two side updates on the parameter CTAs
      SMC_TICK(15);
      smc_wait();  // (3) parameters of the two sides
      SMC_TICK(16);
      side_consts();
      SMC_TICK(6);
    } else {
      smc_sync();  // (2)
      SMC_TICK(12);
      // ---- P3: parameter updates of the two sides on this CTA's attributes
      if (prop || it == 0 || any_changed) reduce_slices(phl, prop ? SH_P0 : SH_L0, prop);  // else: nobody moved
      if (prop && kP == 0 && tid == 0) {
        A.cnt[SH_P0] = nS + 2 - nBcur;
        A.cnt[SH_P1] = nBcur;
      }
      __syncthreads();
      SMC_TICK(13);
      SmcJob jb[2];
      for (int side = 0; side < 2; side++) {
        SmcJob& J = jb[side];
        J.role = side;
        J.hist = (prop ? SH_P0 : SH_L0) + side;
        J.nk = side ? nBcur : nS + 2 - nBcur;
        J.dst = NSB + (prop ? (side ? SM_ST_B : SM_ST_A) : (side ? SM_SL_B : SM_SL_A));
        J.idx = side;
        J.prior = 0;
        J.sub = SUB_SM_RG + q;
        J.uc = off(A.u_rg_c, ((size_t)q * 2 + side) * p);
        J.us = off(A.u_rg_s, ((size_t)q * 2 + side) * p);
      }
      smc_draw_slice(G, jb, 2, j0s, slh, attrS, vS, wS, sigS, denS, cenS, cenv, isgv, true, zigS);
      __syncthreads();
      smc_publish_sdpart(G, kP, denS, sdpart);
      SMC_TICK(14);
      smc_sync();  // (3)
      SMC_TICK(16);
    }
    par ^= 1;
  }
  // ------------------------------------------------------------------------------------------
  // MH ratio (split_merge.cpp:438-540), acceptance (:591) and relabelling (clean_var, common_functions.cpp:296-353)
  // ------------------------------------------------------------------------------------------
  if (isM)
    for (int m = tid; m < ndec; m += SMC_T) A.zStar[pos0 + m] = zc[m];
  SMC_TICK(17);
  smc_sync();  // (f) every slot written by the draws is visible
  SMC_TICK(18);
  const int sA = same ? NSB + SM_ST_A : cA, sB = same ? NSB + SM_ST_B : cB, sM = same ? cA : NSB + SM_ST_M;
  if (isM) {
    // canonical log-normaliser sums of the proposal slots (the 256-leaf tree of phi_job_body), by three warps;
    // CTA 0 also stores them: an accepted proposal hands them to the state
    if (warp < 3) {
      const int slot = NSB + (warp == 0 ? SM_ST_A : (warp == 1 ? SM_ST_B : SM_ST_M));
      double v[8];
#pragma unroll
      for (int k = 0; k < 8; k++) {
        double acc = 0.0;
        for (int j = lane + 32 * k; j < p; j += 256) acc += __ldcg(&G.den[(size_t)slot * pp + j]);
        v[k] = acc;
      }
      double t = tree256_warp(v);
      t = __shfl_sync(SMG_FULL, t, 0);
      if (lane == 0) {
        s_sd[warp] = t;
        if (rank == 0) A.sden[slot] = t;
      }
    }
    const int slots3[3] = {sA, sB, sM};
    for (int t = tid; t < 3 * pp; t += SMC_T) {
      const int k = t / pp, j = t % pp;
      cenv[k * pp + j] = __ldcg(&A.cen[(size_t)slots3[k] * pp + j]);
      isgv[k * pp + j] = __ldcg(&A.isg[(size_t)slots3[k] * pp + j]);
    }
    __syncthreads();
    const double sdv[3] = {same ? s_sd[0] : A.sden[cA], same ? s_sd[1] : A.sden[cB], same ? A.sden[cA] : s_sd[2]};
    SMC_TICK(19);
    // per-member terms (sm_rowterms_body) of the own members: three passes, one per parameter vector, each with its own
    // look-up table T[j][a] = [a != c_j]/sigma_j; one thread per member
    const size_t stride = (size_t)n + 2;
    const unsigned Tb = (unsigned)__cvta_generic_to_shared(T);
    const int tsh = 31 - __clz(TL) + 3;
    for (int vec = 0; vec < 3; vec++) {
      for (int e = tid; e < pp * TL; e += SMC_T) {
        const int j = e / TL, av = e % TL;
        T[e] = (av != 0 && av != cenv[vec * pp + j]) ? isgv[vec * pp + j] : 0.0;
      }
      __syncthreads();
      for (int m = tid; m < mine; m += SMC_T) {
        const int pos = pos0 + m;
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        if (m < rcap) {
          const uint32_t* xr = rowsT + m;
#pragma unroll 4
          for (int w = 0; w < words; w++) {
            const uint32_t xw = xr[(size_t)w * RSTR];
#pragma unroll
            for (int b = 0; b < 4; b++) {
              const unsigned addr = Tb + ((unsigned)(4 * w + b) << tsh) + (((xw >> (8 * b)) & 0xffu) << 3);
              double v;
              asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
              acc[b] += v;
            }
          }
        } else {
          const uint32_t* xr = reinterpret_cast<const uint32_t*>(A.X + (size_t)row_of(pos) * pp);
          for (int w = 0; w < words; w++) {
            const uint32_t xw = xr[w];
#pragma unroll
            for (int b = 0; b < 4; b++) acc[b] += T[(size_t)(4 * w + b) * TL + ((xw >> (8 * b)) & 0xffu)];
          }
        }
        const double ll = -((acc[0] + acc[1]) + (acc[2] + acc[3])) - sdv[vec];
        if (vec == 0) {
          A.rowvals[0 * stride + pos] = ll;  // (side selection in the next pass)
        } else if (vec == 1) {
          const double llA = A.rowvals[0 * stride + pos], llB = ll;
          const int side = same ? zc[m] : (pos < nS ? A.zState[pos] : pos - nS);
          A.rowvals[0 * stride + pos] = side == 0 ? llA : 0.0;
          A.rowvals[1 * stride + pos] = side == 1 ? llB : 0.0;
          double gc = 0.0;
          if (pos < nS) {
            const int zlm = A.zL[pos];
            const int nA = cntL0 - (zlm == 0), nB = cntL1 - (zlm == 1);
            const double a0 = log((double)nA) + llA, a1 = log((double)nB) + llB;
            const double mx = a0 > a1 ? a0 : a1;
            const double p0 = exp(a0 - mx), p1 = exp(a1 - mx);
            double sm = 0.0;
            sm += p0;
            sm += p1;
            gc = log((side == 0 ? p0 : p1) / sm);
          }
          A.rowvals[3 * stride + pos] = gc;
        } else {
          A.rowvals[2 * stride + pos] = ll;
        }
      }
      __syncthreads();
    }
  } else {
    // the six parameter-density terms by attribute (sm_gsphi_prior_attr), this CTA's attributes: one thread per task
    for (int t = tid; t < 6 * sl; t += SMC_T) {
      const int b = t / sl, j = j0s + t % sl;
      if (j < p)
        G.gsvals[(size_t)b * pp + j] =
            sm_gsphi_prior_attr(b, j, pp, mmax, A.phi.attr, A.phi.v, A.phi.w, A.H, A.cnt, A.plan, A.cen, A.sig);
    }
  }
  SMC_TICK(20);
  smc_sync();  // (h) per-member and per-attribute terms
  SMC_TICK(21);
  {
    const int gw = rank * SMC_WARPS + warp;
    if (rank == 0 && warp < 6) {  // the 256-leaf trees of sm_gsphi_prior_body: leaf t = sum of attributes t, t + 256, ...
      const int b = warp;
      double v[8];
#pragma unroll
      for (int k = 0; k < 8; k++) {
        double acc = 0.0;
        for (int j = lane + 32 * k; j < p; j += 256) acc += __ldcg(&G.gsvals[(size_t)b * pp + j]);
        v[k] = acc;
      }
      const double t = tree256_warp(v);
      if (lane == 0) A.terms[b < 3 ? 10 + b : 4 + (b - 3)] = t;
    }
    for (int vb = gw; vb < 4 * SM_RB; vb += nwarps) sm_rowreduce1_warp(nS, A.rowvals, n + 2, A.partial, vb % SM_RB, vb / SM_RB, lane);
  }
  SMC_TICK(22);
  smc_sync();  // (i) partial sums
  SMC_TICK(23);
  if (rank == 0) {
    RngKey k = A.key;
    k.sub = SUB_SM_ACCEPT;
    sm_accept_body(A.info, A.plan, A.cnt, A.partial, A.gamma, A.u_accept, k, A.terms, A.accepted, A.stats, sh);
  }
#ifdef SMG_SMC_NODRAW
  if (rank == 0 && tid == 0) *A.accepted = 0;
#endif
  SMC_TICK(24);
#ifdef SMG_SMC_PROFILE
  if (tid == 0 && G.prof && (rank == 0 || rank == nM)) {
    if (rank == 0) G.prof[58] += 1ull;
    for (int k = 0; k < (rank == 0 ? 32 : 18); k++) G.prof[k + (rank ? 32 : 0)] += s_prof[k];
  }
#endif
  smc_sync();  // (j) the decision
  if (__ldcg(A.accepted) == 0) return;
  // ---- accept: state <- proposal
  if (rank == 0) {
    // the proposal slots join the state: 1/sigma, the per-attribute log-normalisers and their sum re-derived from sigma
    // by the canonical expressions (derive_terms_kernel / phi_job_body), so that a state restored from a snapshot of
    // (centres, sigmas) continues bit for bit
    for (int t = tid; t < 3 * pp; t += SMC_T) {
      const int k = t / pp, j = t % pp;
      const size_t o = (size_t)(NSB + (k == 0 ? SM_ST_A : (k == 1 ? SM_ST_B : SM_ST_M))) * pp + j;
      double w = 0.0, d = 0.0;
      if (j < p) {
        const double sg = __ldcg(&A.sig[o]);
        w = 1.0 / sg;
        d = hamming_den(sg, A.phi.attr[j]);
      }
      A.isg[o] = w;
      G.den[o] = d;
    }
    __syncthreads();
    if (warp < 3) {
      const int slot = NSB + (warp == 0 ? SM_ST_A : (warp == 1 ? SM_ST_B : SM_ST_M));
      double v[8];
#pragma unroll
      for (int k = 0; k < 8; k++) {
        double acc = 0.0;
        for (int j = lane + 32 * k; j < p; j += 256) acc += G.den[(size_t)slot * pp + j];
        v[k] = acc;
      }
      const double t = tree256_warp(v);
      if (lane == 0) A.sden[slot] = t;
    }
    __syncthreads();
    sm_apply_params_body(A.info, NSB, A.Kcap, pp, A.cen, A.sig, A.isg, A.sden, A.cnt, A.counts, A.Kptr, A.phi.status);
  }
  if (isM) {
    const int Kold = A.info->K;
    for (int m = tid; m < mine; m += SMC_T) A.c[row_of(pos0 + m)] = same ? (zc[m] == 0 ? Kold : cB) : cB;
  }
  if (same) return;
  smc_sync();  // (k) members relabelled before the last label moves into the hole
  {
    const int hole = cA, last = A.info->K - 1;
    if (hole != last)
      for (int i = rank * SMC_T + tid; i < n; i += CS * SMC_T)
        if (A.c[i] == last) A.c[i] = hole;
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct SmcHost {
  bool ok = false;
  int CS = 0, TL = 8, mcap = 0, rcap = 0;
  size_t smem = 0;
  double* gsvals = nullptr;
  double *nr_d0 = nullptr, *nr_lg = nullptr;
  int *nr_pre = nullptr, *nr_z = nullptr, *nr_idx = nullptr, *walk_out = nullptr;
  unsigned long long* prof = nullptr;
};

// decides whether the cluster kernel can run this chain (shared-memory budget, cluster size the device accepts)
static int smc_setup(smg_chain* ch, SmcHost* H) {
  const char* env = getenv("SMG_SM_MODE");  // cluster | coop | multi (see sm_step for the default)
  if (env && strcmp(env, "cluster") != 0) return 0;
  int cc_major = 0, smem_optin = 0;
  cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, ch->device);
  cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ch->device);
  if (cc_major < 9 || ch->t <= 0) return 0;
  cudaFuncAttributes fa;
  if (cudaFuncGetAttributes(&fa, sm_cluster_kernel) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  // the attribute is per function, not per chain: always the device maximum, so that chains of different shapes coexist
  const size_t budget = (size_t)smem_optin - fa.sharedSizeBytes - 1024;
  if (cudaFuncSetAttribute(sm_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)smem_optin - fa.sharedSizeBytes)) != cudaSuccess ||
      cudaFuncSetAttribute(sm_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  const char* ecs = getenv("SMG_SMC_CS");
  int TL = 8;
  while (TL <= ch->mmax) TL <<= 1;  // table entries per attribute: a power of two above the largest code
  for (int CS = ecs ? atoi(ecs) : SMC_MAXCS; CS >= 8; CS >>= 1) {
    const int nM = CS / 2;
    if (ch->pp % nM) continue;
    const int mcap = ((int)(((long long)ch->n + 2 + nM - 1) / nM) + 31) & ~31;
    const SmcLayout L0 = smc_layout(ch->pp, ch->mmax, TL, CS, mcap, 0);
    if (L0.total + (size_t)std::min(mcap, 64) * (ch->pp + 8) + 64 > budget) continue;
    int rcap = (int)std::min<size_t>((size_t)mcap, (budget - L0.total) / (ch->pp + 8));
    while (rcap > 0 && smc_layout(ch->pp, ch->mmax, TL, CS, mcap, rcap).total > budget) rcap--;
    if (rcap < std::min(mcap, 64)) continue;
    const SmcLayout L = smc_layout(ch->pp, ch->mmax, TL, CS, mcap, rcap);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CS);
    cfg.blockDim = dim3(SMC_T);
    cfg.dynamicSmemBytes = L.total;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int ncl = 0;
    if (cudaOccupancyMaxActiveClusters(&ncl, sm_cluster_kernel, &cfg) != cudaSuccess || ncl < 1) {
      (void)cudaGetLastError();
      continue;
    }
    H->CS = CS;
    H->TL = TL;
    H->mcap = mcap;
    H->rcap = rcap;
    H->smem = L.total;
    break;
  }
  if (!H->CS) return 0;
  const size_t nl = (size_t)ch->n + 64 + (size_t)SMC_MAXCS * 32;
  SMG_CUDA(dev_malloc(&H->gsvals, (size_t)6 * ch->pp * 8, ch->st));
  SMG_CUDA(dev_malloc(&H->nr_d0, nl * 8, ch->st));
  SMG_CUDA(dev_malloc(&H->nr_lg, nl * 8, ch->st));
  SMG_CUDA(dev_malloc(&H->nr_pre, nl * 4, ch->st));
  SMG_CUDA(dev_malloc(&H->nr_z, nl * 4, ch->st));
  SMG_CUDA(dev_malloc(&H->nr_idx, nl * 4, ch->st));
  SMG_CUDA(dev_malloc(&H->walk_out, 2 * 4, ch->st));
  SMG_CUDA(dev_malloc(&H->prof, 64 * 8, ch->st));
  SMG_CUDA(cudaMemsetAsync(H->prof, 0, 64 * 8, ch->st));
  SMG_CUDA(cudaMemsetAsync(H->walk_out, 0, 8, ch->st));
  H->ok = true;
  return 0;
}

static void smc_free(smg_chain* ch, SmcHost* H) {
  void* ptrs[] = {H->gsvals, H->nr_d0, H->nr_lg, H->nr_pre, H->nr_z, H->nr_idx, H->walk_out, H->prof};
  for (void* q : ptrs)
    if (q) cudaFreeAsync(q, ch->st);
}

static cudaError_t smc_launch(smg_chain* ch, const SmcHost* H, const SmChainArgs& CA) {
  SmcArgs G;
  G.A = CA;
  G.CS = H->CS;
  G.sl = ch->pp / (H->CS / 2);
  G.TL = H->TL;
  G.mcap = H->mcap;
  G.rcap = H->rcap;
  G.den = ch->den;
  G.gsvals = H->gsvals;
  G.nr_d0 = H->nr_d0;
  G.nr_lg = H->nr_lg;
  G.nr_pre = H->nr_pre;
  G.nr_z = H->nr_z;
  G.nr_idx = H->nr_idx;
  G.walk_out = H->walk_out;
  G.prof = H->prof;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(H->CS);
  cfg.blockDim = dim3(SMC_T);
  cfg.dynamicSmemBytes = H->smem;
  cfg.stream = ch->st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = H->CS;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, sm_cluster_kernel, G);
}

}  // namespace smg
