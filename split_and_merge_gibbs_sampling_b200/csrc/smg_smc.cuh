// smg_smc.cuh -- the Jain-Neal split-merge proposal (code/split_merge.cpp:542-598) as ONE thread-block cluster.
//
// sm_chain_kernel (smg_sm.cuh) runs the proposal on 120 CTAs that meet ~55 times at a global-counter grid barrier
// (1.2 us each) and re-read the member rows from L2 in every phase; 57% of a sweep was spent there, 65% of its stall
// samples on the barrier.  Here the same proposal lives on one cluster of CS (16) CTAs:
//   * the members of S (+ the two anchors) are dealt to the CTAs in contiguous position blocks; every CTA keeps its
//     members' rows (as many as fit), sides, row indices and allocation logits in shared memory for the whole proposal;
//   * the parameter updates are dealt by ATTRIBUTE slices (pp/CS attributes per CTA, all jobs of a step side by side,
//     PHI_G lanes per attribute): a CTA sums its slice of the per-CTA side histograms straight out of the other CTAs'
//     shared memory (DSMEM), draws, and publishes (centre, 1/sigma) into every CTA's shared memory;
//   * phases are separated by hardware cluster barriers (~0.2 us) -- three per restricted scan;
//   * a member's restricted-Gibbs decision is piecewise constant in D = (log n_1 + LL_1) - (log n_2 + LL_2)
//     (smg_sm.cuh, sm_d_region); with the count term bounded over the WHOLE scan (|log-count term| <= log nS) almost
//     every member is decided by its owner without knowing the running counts; the few that are not go through the
//     ordered walk of smg_sm.cuh on CTA 0 (one extra barrier, only when there are any).
// Same draws (Philox keys, injected uniforms), same decisions and the same per-attribute arithmetic as the other two
// paths; only the association of the sums of log-normalisers inside the scans differs (rank partials instead of the
// 256-leaf tree), i.e. the last bits of the likelihood differences.  The slots an accepted proposal hands to the
// state get their log-normaliser sums from the canonical tree.
// The kernel leaves 132 SMs free: the likelihood block of the next pass runs beside it (sweep()).
#pragma once
#include "smg_sm.cuh"

namespace smg {

#define SMC_T 512
#define SMC_WARPS (SMC_T / 32)
#define SMC_MAXCS 16

struct SmcArgs {
  SmChainArgs A;   // same fields as the cooperative kernel
  int CS;          // CTAs in the cluster
  int sl;          // attributes per CTA (pp / CS)
  int mcap;        // member capacity per CTA (>= ceil((n+2)/CS))
  int rcap;        // member rows cached in shared memory per CTA
  double* den;     // [slots][pp] per-attribute log-normalisers
  double* gsvals;  // [6][pp] per-attribute addends of the six parameter-density terms
  double *nr_d0, *nr_lg;         // [n + 64] members whose decision depends on the running counts, by region
  int *nr_pre, *nr_z, *nr_idx;   // [n + 64]
  int* walk_out;                 // [2] side-1 change of those members (per scan parity)
  unsigned long long* prof;      // optional phase cycle counters of CTA 0 / thread 0 (SMG_SMC_PROFILE)
};

struct SmcLayout {
  size_t isgv, sigS, denS, vS, wS, sdpart, lg, phs, phl, slh, attrS, xchg, sel, srow, cenv, cenS, z, rows, total;
};
__host__ __device__ inline size_t smc_al(size_t x) { return (x + 15) & ~(size_t)15; }
__host__ __device__ inline SmcLayout smc_layout(int pp, int mmax, int CS, int mcap, int rcap) {
  SmcLayout L;
  const int sl = pp / CS;
  size_t o = 0;
  L.isgv = o, o = smc_al(o + (size_t)3 * pp * 8);
  L.sigS = o, o = smc_al(o + (size_t)3 * sl * 8);
  L.denS = o, o = smc_al(o + (size_t)3 * sl * 8);
  L.vS = o, o = smc_al(o + (size_t)sl * 8);
  L.wS = o, o = smc_al(o + (size_t)sl * 8);
  L.sdpart = o, o = smc_al(o + (size_t)3 * SMC_MAXCS * 8);
  L.lg = o, o = smc_al(o + (size_t)mcap * 8);
  L.phs = o, o = smc_al(o + (size_t)2 * pp * mmax * 4);
  L.phl = o, o = smc_al(o + (size_t)2 * pp * mmax * 4);
  L.slh = o, o = smc_al(o + (size_t)SH_N * sl * mmax * 4);
  L.attrS = o, o = smc_al(o + (size_t)sl * 4);
  L.xchg = o, o = smc_al(o + (size_t)2 * SMC_MAXCS * 4 * 4);
  L.sel = o, o = smc_al(o + (size_t)SMC_MAXCS * 4);
  L.srow = o, o = smc_al(o + (size_t)mcap * 4);
  L.cenv = o, o = smc_al(o + (size_t)3 * pp);
  L.cenS = o, o = smc_al(o + (size_t)3 * sl);
  L.z = o, o = smc_al(o + (size_t)4 * mcap);
  L.rows = o, o = smc_al(o + (size_t)rcap * pp);
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

// One step's parameter updates for this CTA's attribute slice: task = (job, attribute), PHI_G lanes each.
// The arithmetic of a task is phi_job_body's (smg_kernels.cuh).  Results go to the slice arrays (sigma is the
// "current sigma" of the job's next update), to the global slot `dst`, and -- for the two sides being scanned --
// into every CTA's parameter vectors.
__device__ __forceinline__ void smc_draw_slice(const SmcArgs& G, const SmcJob* jobs, int nj, int rank, int j0s, const int* slh,
                                               const int* attrS, const double* vS, const double* wS, double* sigS,
                                               double* denS, uint8_t* cenS, uint8_t* cenv, double* isgv, bool publish) {
  const SmChainArgs& A = G.A;
  const int sl = G.sl, pp = A.pp, mmax = A.mmax, CS = G.CS;
  const int g = threadIdx.x & (PHI_G - 1), lane = threadIdx.x & 31, gbase = lane & ~(PHI_G - 1);
  const unsigned gmask = ((1u << PHI_G) - 1u) << gbase;
  const int ntask = nj * sl;
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
        for (int r = g; r < CS; r += PHI_G) {
          smc_st_u8(smc_map(cenv + J.role * pp + j, r), 0);
          smc_st_f64(smc_map(isgv + J.role * pp + j, r), 0.0);
        }
      continue;
    }
    RngKey key = A.key;
    key.sub = J.sub;
    const int m = attrS[jl];
    int center;
    double s_match = 0.0;
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
    if (J.us || A.phi.sigma_exact) {
      if (g == 0) {
        const double us = get_u(J.us, (size_t)j, key, U_SIGMA, (uint32_t)J.idx, (uint32_t)j);
        uu = hig_inv_u_d(us, vv, ww, (double)m);
      }
    } else {
      uu = hig_draw_u_grp(key, (uint32_t)J.idx, (uint32_t)j, vv, ww, (double)m, g, gmask, gbase);
    }
    double sigma = 0.0, isg = 0.0;
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
    if (publish && J.role < 2) {
      isg = __shfl_sync(gmask, isg, gbase);
      for (int r = g; r < CS; r += PHI_G) {
        smc_st_u8(smc_map(cenv + J.role * pp + j, r), center);
        smc_st_f64(smc_map(isgv + J.role * pp + j, r), isg);
      }
    }
  }
}

// after the draws of a step: this CTA's partial sums of the log-normalisers of the two scanned sides, to every CTA
__device__ __forceinline__ void smc_publish_sdpart(const SmcArgs& G, int rank, const double* denS, double* sdpart) {
  const int sl = G.sl, CS = G.CS;
  const int t = threadIdx.x;
  if (t < 2 * CS) {
    const int role = t / CS, r = t % CS;
    double acc = 0.0;
    for (int jl = 0; jl < sl; jl++) acc += denS[role * sl + jl];
    smc_st_f64(smc_map(sdpart + role * SMC_MAXCS + rank, r), acc);
  }
}

#ifdef SMG_SMC_PROFILE
#define SMC_TICK(k)                                                              \
  do {                                                                           \
    const long long _t = clock64();                                              \
    if (rank == 0 && threadIdx.x == 0 && G.prof) G.prof[k] += (unsigned long long)(_t - tk); \
    tk = _t;                                                                     \
  } while (0)
#else
#define SMC_TICK(k)
#endif

__global__ void __launch_bounds__(SMC_T, 1) sm_cluster_kernel(SmcArgs G) {
  const SmChainArgs& A = G.A;
  extern __shared__ __align__(16) unsigned char smc_raw[];
  const int n = A.n, p = A.p, pp = A.pp, mmax = A.mmax, NSB = A.NS;
  const int CS = G.CS, sl = G.sl, mcap = G.mcap, rcap = G.rcap;
  const SmcLayout L = smc_layout(pp, mmax, CS, mcap, rcap);
  double* isgv = reinterpret_cast<double*>(smc_raw + L.isgv);   // [3][pp] 1/sigma of the vectors being evaluated
  double* sigS = reinterpret_cast<double*>(smc_raw + L.sigS);   // [3][sl] current sigma of the three update chains
  double* denS = reinterpret_cast<double*>(smc_raw + L.denS);   // [3][sl]
  double* vS = reinterpret_cast<double*>(smc_raw + L.vS);
  double* wS = reinterpret_cast<double*>(smc_raw + L.wS);
  double* sdpart = reinterpret_cast<double*>(smc_raw + L.sdpart);  // [3][16] per-CTA partial log-normaliser sums
  double* lg = reinterpret_cast<double*>(smc_raw + L.lg);       // [mcap] logit(u) of the members, current scan
  int* phs = reinterpret_cast<int*>(smc_raw + L.phs);           // [2][pp][mmax] own members by current-state side
  int* phl = reinterpret_cast<int*>(smc_raw + L.phl);           // [2][pp][mmax] own members by launch / proposal side
  int* slh = reinterpret_cast<int*>(smc_raw + L.slh);           // [SH_N][sl][mmax] histograms of this CTA's attributes
  int* attrS = reinterpret_cast<int*>(smc_raw + L.attrS);
  int* xchg = reinterpret_cast<int*>(smc_raw + L.xchg);         // [2][16][4] per-CTA counters of a step, from every CTA
  int* sel = reinterpret_cast<int*>(smc_raw + L.sel);           // [16] members found by each CTA
  int* srow = reinterpret_cast<int*>(smc_raw + L.srow);         // [mcap] row of each own member
  uint8_t* cenv = smc_raw + L.cenv;                             // [3][pp]
  uint8_t* cenS = smc_raw + L.cenS;                             // [3][sl]
  uint8_t* zc = smc_raw + L.z;                                  // [mcap] side (launch, then proposal)
  uint8_t* zn = zc + mcap;                                      // [mcap] side decided by the scan in progress
  uint8_t* zs = zn + mcap;                                      // [mcap] side under the current state
  uint8_t* zl = zs + mcap;                                      // [mcap] launch side (kept for the proposal density)
  uint8_t* rows = smc_raw + L.rows;                             // [rcap][pp]
  __shared__ double sh[256];
  __shared__ int s_sel[8];
  __shared__ int s_wa[SMC_WARPS], s_wb[SMC_WARPS], s_tot[4];
  __shared__ double s_sd[4];

  const int rank = (int)cluster_cta_rank();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j0s = rank * sl;
  const size_t len = (size_t)pp * mmax;
  const int chunks = pp / 16;
  auto off = [](const double* base, size_t o) -> const double* { return base ? base + o : nullptr; };
#ifdef SMG_SMC_PROFILE
  long long tk = clock64();
#endif

  // ---- pair (split_merge.cpp:275) and this CTA's share of the row scan that builds S (:280-301)
  if (tid == 0) {
    RngKey k = A.key;
    k.sub = SUB_SM_SELECT;
    int i1, i2;
    sm_pick_pair(n, A.u_pair, k, &i1, &i2);
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
  // prior parameters of the three launch clusters (split_merge.cpp:331-343, :379-380)
  {
    SmcJob jb[3];
    for (int k = 0; k < 3; k++) {
      jb[k].role = k;
      jb[k].hist = 0;
      jb[k].nk = 0;
      jb[k].dst = NSB + (k == 0 ? SM_SL_A : (k == 1 ? SM_SL_B : SM_ML_M));
      jb[k].idx = k;
      jb[k].prior = 1;
      jb[k].sub = SUB_SM_PRIOR;
      jb[k].uc = off(A.u_prior_c, (size_t)k * p);
      jb[k].us = off(A.u_prior_s, (size_t)k * p);
    }
    smc_draw_slice(G, jb, 3, rank, j0s, slh, attrS, vS, wS, sigS, denS, cenS, cenv, isgv, true);
    __syncthreads();
    smc_publish_sdpart(G, rank, denS, sdpart);
  }
  smc_sync();  // (a) member counts, launch parameters
  SMC_TICK(0);
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
  smc_sync();  // (b) S complete
  // ---- own members: positions [pos0, pos0 + mine); the anchors are the positions nS (side 0) and nS + 1 (side 1)
  const int per = (nS + 2 + CS - 1) / CS;
  const int pos0 = rank * per;
  const int mine = max(0, min(per, nS + 2 - pos0));
  const int ndec = max(0, min(mine, nS - pos0));  // own members that are re-allocated (not anchors)
  {
    RngKey k = A.key;
    k.sub = SUB_SM_LAUNCH;
    for (int m = tid; m < mine; m += SMC_T) {
      const int pos = pos0 + m;
      int row, zstate, zlaunch;
      if (pos < nS) {
        row = A.S[pos];
        zstate = A.zState[pos];
        const double u = get_u(A.u_launch, pos, k, U_SM_LAUNCH, (uint32_t)pos, 0u);  // split_merge.cpp:346
        const int z = (int)(2.0 * u);
        zlaunch = z > 1 ? 1 : z;
      } else {
        row = pos == nS ? i1 : i2;
        zstate = zlaunch = pos - nS;
      }
      srow[m] = row;
      zs[m] = (uint8_t)zstate;
      zc[m] = (uint8_t)zlaunch;
      zn[m] = (uint8_t)zlaunch;
    }
  }
  for (int q = tid; q < 2 * (int)len; q += SMC_T) phs[q] = phl[q] = 0;
  __syncthreads();
  {  // rows into shared memory (as many as fit), and both side histograms of the own members
    const int ncache = min(mine, rcap);
    for (int t = tid; t < ncache * chunks; t += SMC_T) {
      const int m = t / chunks, q = t % chunks;
      *reinterpret_cast<uint4*>(rows + (size_t)m * pp + q * 16) =
          *reinterpret_cast<const uint4*>(A.X + (size_t)srow[m] * pp + q * 16);
    }
    __syncthreads();
    int c1s = 0, c1l = 0;
    for (int t = tid; t < mine * chunks; t += SMC_T) {
      const int m = t / chunks, q = t % chunks;
      const uint8_t* xr = m < rcap ? rows + (size_t)m * pp : A.X + (size_t)srow[m] * pp;
      const uint4 v = *reinterpret_cast<const uint4*>(xr + q * 16);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
      const int gs = zs[m], gl = zc[m];
      if (q == 0) {
        c1s += gs;
        c1l += gl;
      }
      int* Hs = phs + (size_t)gs * len + (size_t)q * 16 * mmax;
      int* Hl = phl + (size_t)gl * len + (size_t)q * 16 * mmax;
#pragma unroll
      for (int b = 0; b < 16; b++) {
        const int x = (w[b >> 2] >> ((b & 3) * 8)) & 0xff;
        if (x) {
          atomicAdd(&Hs[b * mmax + (x - 1)], 1);
          atomicAdd(&Hl[b * mmax + (x - 1)], 1);
        }
      }
    }
    c1s = warp_sum_i(c1s);
    c1l = warp_sum_i(c1l);
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
  smc_sync();  // (c) side histograms and side counts of every CTA
  SMC_TICK(1);
  int nBcur = 0;  // side-1 members of the launch / proposal allocation, anchor i_2 included
  int cntS1 = 0;
  for (int r = 0; r < CS; r++) {
    cntS1 += xchg[(0 * SMC_MAXCS + r) * 4 + 0];
    nBcur += xchg[(0 * SMC_MAXCS + r) * 4 + 1];
  }
  // slice histograms: sum of the other CTAs' partial histograms over this CTA's attributes (DSMEM reads)
  auto reduce_slices = [&](const int* ph, int h0, bool to_global) {
    const int per_side = sl * mmax;
    for (int t = tid; t < 2 * per_side; t += SMC_T) {
      const int s = t / per_side, e = t % per_side;
      const int* src = ph + (size_t)s * len + (size_t)j0s * mmax + e;
      int acc = 0;
      for (int r = 0; r < CS; r++) acc += smc_ld_s32(smc_map(src, r));
      slh[(size_t)(h0 + s) * per_side + e] = acc;
      if (to_global) A.H[(size_t)(h0 + s) * len + (size_t)j0s * mmax + e] = acc;
    }
  };
  reduce_slices(phs, SH_S0, true);
  __syncthreads();
  for (int e = tid; e < sl * mmax; e += SMC_T) {
    const int v = slh[(size_t)SH_S0 * sl * mmax + e] + slh[(size_t)SH_S1 * sl * mmax + e];
    slh[(size_t)SH_M * sl * mmax + e] = v;
    A.H[(size_t)SH_M * len + (size_t)j0s * mmax + e] = v;
  }
  if (rank == 0 && tid == 0) {
    A.cnt[SH_S0] = nS + 2 - cntS1;
    A.cnt[SH_S1] = cntS1;
    A.cnt[SH_M] = nS + 2;
  }
  __syncthreads();

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
      for (int m = tid; m < mine; m += SMC_T) {
        zl[m] = zc[m];
        if (pos0 + m < nS) A.zL[pos0 + m] = zc[m];
      }
      for (int t = tid; t < 2 * sl; t += SMC_T) {
        const int role = t / sl, jl = t % sl, j = j0s + jl;
        const size_t d = (size_t)(NSB + (role ? SM_ST_B : SM_ST_A)) * pp + j;
        const double sg = sigS[role * sl + jl];
        A.cen[d] = cenS[role * sl + jl];
        A.sig[d] = sg;
        A.isg[d] = (j < p) ? 1.0 / sg : 0.0;
        G.den[d] = denS[role * sl + jl];
      }
      if (rank == 0 && tid == 0) {
        A.cnt[SH_L0] = cntL0;
        A.cnt[SH_L1] = cntL1;
      }
    }
    const int q = prop ? A.t : it;
    const bool do_scan = prop ? (same != 0) : (it < A.t);
    if (do_scan) {
      // ---- P1: allocation logits, likelihood differences, count-free decisions
      {
        RngKey k = A.key;
        k.sub = SUB_SM_RG + q;
        const double* ui = off(A.u_rg, (size_t)q * n);
        for (int m = tid; m < ndec; m += SMC_T) lg[m] = sm_logit_u(ui, pos0 + m, k);
      }
      double sdA, sdB;
      {
        const double a = lane < CS ? sdpart[0 * SMC_MAXCS + lane] : 0.0, b = lane < CS ? sdpart[1 * SMC_MAXCS + lane] : 0.0;
        sdA = warp_sum(a);
        sdB = warp_sum(b);
      }
      __syncthreads();
      const double dc_max = nS > 0 ? sm_dc_bound(nS, 1) + SM_DC_MARGIN : 0.0;
      const double dc_min = nS > 0 ? sm_dc_bound(nS, nS) - SM_DC_MARGIN : 0.0;
      for (int m0 = warp; m0 < ndec; m0 += 2 * SMC_WARPS) {
        const int m1 = m0 + SMC_WARPS;
        const bool has1 = m1 < ndec;
        const uint8_t* x0 = m0 < rcap ? rows + (size_t)m0 * pp : A.X + (size_t)srow[m0] * pp;
        const uint8_t* x1 = has1 ? (m1 < rcap ? rows + (size_t)m1 * pp : A.X + (size_t)srow[m1] * pp) : x0;
        double a0 = 0.0, b0 = 0.0, a1 = 0.0, b1 = 0.0;
        for (int j0 = lane * 8; j0 < pp; j0 += 256) {
          const uint2 ca = *reinterpret_cast<const uint2*>(cenv + j0), cb = *reinterpret_cast<const uint2*>(cenv + pp + j0);
          const uint2 xv0 = *reinterpret_cast<const uint2*>(x0 + j0), xv1 = *reinterpret_cast<const uint2*>(x1 + j0);
          double wa[8], wb[8];
#pragma unroll
          for (int b = 0; b < 4; b++) {
            const double2 ta = reinterpret_cast<const double2*>(isgv + j0)[b], tb = reinterpret_cast<const double2*>(isgv + pp + j0)[b];
            wa[2 * b] = ta.x, wa[2 * b + 1] = ta.y;
            wb[2 * b] = tb.x, wb[2 * b + 1] = tb.y;
          }
          const uint32_t ma0 = __vcmpne4(xv0.x, ca.x), ma1 = __vcmpne4(xv0.y, ca.y);
          const uint32_t mb0 = __vcmpne4(xv0.x, cb.x), mb1 = __vcmpne4(xv0.y, cb.y);
          const uint32_t na0 = __vcmpne4(xv1.x, ca.x), na1 = __vcmpne4(xv1.y, ca.y);
          const uint32_t nb0 = __vcmpne4(xv1.x, cb.x), nb1 = __vcmpne4(xv1.y, cb.y);
#pragma unroll
          for (int b = 0; b < 4; b++) {
            if (ma0 & (0xffu << (8 * b))) a0 += wa[b];
            if (mb0 & (0xffu << (8 * b))) b0 += wb[b];
            if (na0 & (0xffu << (8 * b))) a1 += wa[b];
            if (nb0 & (0xffu << (8 * b))) b1 += wb[b];
          }
#pragma unroll
          for (int b = 0; b < 4; b++) {
            if (ma1 & (0xffu << (8 * b))) a0 += wa[4 + b];
            if (mb1 & (0xffu << (8 * b))) b0 += wb[4 + b];
            if (na1 & (0xffu << (8 * b))) a1 += wa[4 + b];
            if (nb1 & (0xffu << (8 * b))) b1 += wb[4 + b];
          }
        }
        a0 = warp_sum(a0);
        b0 = warp_sum(b0);
        a1 = warp_sum(a1);
        b1 = warp_sum(b1);
        if (lane < 2 && (lane == 0 || has1)) {
          const int m = lane ? m1 : m0;
          const double d0 = lane ? ((-a1 - sdA) - (-b1 - sdB)) : ((-a0 - sdA) - (-b0 - sdB));
          const double lgm = lg[m];
          const int rlo = sm_d_region(dc_min + d0, lgm), rhi = sm_d_region(dc_max + d0, lgm);
          if (rlo == rhi) {
            zn[m] = (uint8_t)(~rlo & 1);
          } else {
            zn[m] = 2;
            A.dl[pos0 + m] = d0;
          }
        }
      }
      __syncthreads();
      // ordered pass over the own members: side-1 change of the decided ones, list of the undecided ones (by position)
      int carry_d = 0, carry_n = 0;
      for (int base = 0; base < ndec; base += SMC_T) {
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
          G.nr_lg[Lx] = lg[m];
        }
        carry_d += s_tot[0];
        carry_n += s_tot[1];
        __syncthreads();
      }
      if (tid < CS) {
        smc_st_s32(smc_map(xchg + (par * SMC_MAXCS + rank) * 4 + 0, tid), carry_d);
        smc_st_s32(smc_map(xchg + (par * SMC_MAXCS + rank) * 4 + 1, tid), carry_n);
      }
      SMC_TICK(2);
      smc_sync();  // (1) decisions that do not depend on the counts
      SMC_TICK(3);
      int dsum = 0, nnr_tot = 0, nnr_mine = 0;
      for (int r = 0; r < CS; r++) {
        dsum += xchg[(par * SMC_MAXCS + r) * 4 + 0];
        const int v = xchg[(par * SMC_MAXCS + r) * 4 + 1];
        nnr_tot += v;
        if (r == rank) nnr_mine = v;
      }
      int extra = 0;
      if (nnr_tot > 0) {
        // ---- the members whose decision depends on the running counts: ordered walk on CTA 0 (split_merge.cpp:186-216
        //      one at a time; here 32 at a time with the count term bounded inside the batch, see sm_rdecide_body)
        if (rank == 0 && warp == 0) {
          int regionbase = nBcur;
          for (int r = 0; r < CS; r++) {
            const int nnr = xchg[(par * SMC_MAXCS + r) * 4 + 1];
            const int lbase = r * per;
            for (int b0 = 0; b0 < nnr; b0 += 32) {
              const int qq = b0 + lane;
              const bool v = qq < nnr;
              const double md0 = v ? G.nr_d0[lbase + qq] : 0.0, mlg = v ? G.nr_lg[lbase + qq] : 0.0;
              const int mpre = v ? G.nr_pre[lbase + qq] : 0;
              int mz = v ? G.nr_z[lbase + qq] : 0;
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
        for (int k = tid; k < nnr_mine; k += SMC_T) zn[G.nr_idx[pos0 + k]] = (uint8_t)__ldcg(&G.nr_z[pos0 + k]);
        __syncthreads();
      }
      nBcur += dsum + extra;
      // ---- P2: move the rows of the members that changed side between the two side histograms
      for (int t = tid; t < ndec * chunks; t += SMC_T) {
        const int m = t / chunks, qc = t % chunks;
        const int go = zc[m], gn = zn[m];
        if (go == gn) continue;
        const uint8_t* xr = m < rcap ? rows + (size_t)m * pp : A.X + (size_t)srow[m] * pp;
        const uint4 v = *reinterpret_cast<const uint4*>(xr + qc * 16);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        int* Ho = phl + (size_t)go * len + (size_t)qc * 16 * mmax;
        int* Hn = phl + (size_t)gn * len + (size_t)qc * 16 * mmax;
#pragma unroll
        for (int b = 0; b < 16; b++) {
          const int x = (w[b >> 2] >> ((b & 3) * 8)) & 0xff;
          if (x) {
            atomicSub(&Ho[b * mmax + (x - 1)], 1);
            atomicAdd(&Hn[b * mmax + (x - 1)], 1);
          }
        }
      }
      __syncthreads();
      for (int m = tid; m < ndec; m += SMC_T) zc[m] = zn[m];
      SMC_TICK(4);
      smc_sync();  // (2) side histograms of every CTA
      SMC_TICK(5);
      reduce_slices(phl, prop ? SH_P0 : SH_L0, prop);
      if (prop && rank == 0 && tid == 0) {
        A.cnt[SH_P0] = nS + 2 - nBcur;
        A.cnt[SH_P1] = nBcur;
      }
      __syncthreads();
    }
    // ---- P3: parameter updates of this step on this CTA's attributes
    SmcJob jb[3];
    int nj = 0;
    if (do_scan) {
      for (int side = 0; side < 2; side++) {
        SmcJob& J = jb[nj];
        J.role = side;
        J.hist = (prop ? SH_P0 : SH_L0) + side;
        J.nk = side ? nBcur : nS + 2 - nBcur;
        J.dst = NSB + (prop ? (side ? SM_ST_B : SM_ST_A) : (side ? SM_SL_B : SM_SL_A));
        J.idx = nj;
        J.prior = 0;
        J.sub = SUB_SM_RG + q;
        J.uc = off(A.u_rg_c, ((size_t)q * 2 + side) * p);
        J.us = off(A.u_rg_s, ((size_t)q * 2 + side) * p);
        nj++;
      }
    }
    if (prop || it < A.r) {
      // the r updates of the merge launch state (split_merge.cpp:386-387) are an independent chain on the fixed merged
      // histogram; the final one of the proposal (:584) keeps job index 2 as in the other paths
      SmcJob& J = jb[nj];
      J.role = 2;
      J.hist = SH_M;
      J.nk = nS + 2;
      J.dst = NSB + (prop ? SM_ST_M : SM_ML_M);
      J.idx = prop ? 2 : nj;
      J.prior = 0;
      J.sub = SUB_SM_MERGE + (prop ? A.r : it);
      J.uc = off(A.u_mg_c, (size_t)(prop ? A.r : it) * p);
      J.us = off(A.u_mg_s, (size_t)(prop ? A.r : it) * p);
      nj++;
    }
    if (nj > 0) {
      // a job on an empty side is skipped (common_functions.cpp:547): its parameters stay as they are
      int nk_ok = 0;
      SmcJob act[3];
      for (int k = 0; k < nj; k++)
        if (jb[k].nk > 0) act[nk_ok++] = jb[k];
      smc_draw_slice(G, act, nk_ok, rank, j0s, slh, attrS, vS, wS, sigS, denS, cenS, cenv, isgv, do_scan);
    }
    if (do_scan) {
      __syncthreads();
      smc_publish_sdpart(G, rank, denS, sdpart);
      SMC_TICK(6);
      smc_sync();  // (3) parameters of the two sides in every CTA
      SMC_TICK(7);
      par ^= 1;
    }
  }
  // ------------------------------------------------------------------------------------------
  // MH ratio (split_merge.cpp:438-540), acceptance (:591) and relabelling (clean_var, common_functions.cpp:296-353)
  // ------------------------------------------------------------------------------------------
  for (int m = tid; m < mine; m += SMC_T)
    if (pos0 + m < nS) A.zStar[pos0 + m] = zc[m];
  smc_sync();  // (f) every slot written by the draws is visible
  const int sA = same ? NSB + SM_ST_A : cA, sB = same ? NSB + SM_ST_B : cB, sM = same ? cA : NSB + SM_ST_M;
  {
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
  }
  const double sdA = same ? s_sd[0] : A.sden[cA], sdB = same ? s_sd[1] : A.sden[cB], sdM = same ? A.sden[cA] : s_sd[2];
  // per-member terms (sm_rowterms_body), own members; the six parameter-density terms by attribute
  for (int m = warp; m < mine; m += SMC_WARPS) {
    const int pos = pos0 + m;
    const uint8_t* x = m < rcap ? rows + (size_t)m * pp : A.X + (size_t)srow[m] * pp;
    const int side = same ? zc[m] : zs[m];
    const double llA = -warp_mismatch_dot(x, cenv, isgv, pp, lane) - sdA;
    const double llB = -warp_mismatch_dot(x, cenv + pp, isgv + pp, pp, lane) - sdB;
    const double llM = -warp_mismatch_dot(x, cenv + 2 * pp, isgv + 2 * pp, pp, lane) - sdM;
    if (lane == 0) {
      const size_t stride = (size_t)n + 2;
      A.rowvals[0 * stride + pos] = side == 0 ? llA : 0.0;
      A.rowvals[1 * stride + pos] = side == 1 ? llB : 0.0;
      A.rowvals[2 * stride + pos] = llM;
      double gc = 0.0;
      if (pos < nS) {
        const int zlm = zl[m];
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
    }
  }
  for (int t = tid; t < 6 * sl; t += SMC_T) {
    const int b = t / sl, j = j0s + t % sl;
    if (j < p)
      G.gsvals[(size_t)b * pp + j] =
          sm_gsphi_prior_attr(b, j, pp, mmax, A.phi.attr, A.phi.v, A.phi.w, A.H, A.cnt, A.plan, A.cen, A.sig);
  }
  SMC_TICK(8);
  smc_sync();  // (h) per-member and per-attribute terms
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
  smc_sync();  // (i) partial sums
  if (rank == 0) {
    RngKey k = A.key;
    k.sub = SUB_SM_ACCEPT;
    sm_accept_body(A.info, A.plan, A.cnt, A.partial, A.gamma, A.u_accept, k, A.terms, A.accepted, A.stats, sh);
  }
  SMC_TICK(9);
#ifdef SMG_SMC_PROFILE
  if (rank == 0 && tid == 0 && G.prof) G.prof[10] += 1ull;
#endif
  smc_sync();  // (j) the decision
  if (__ldcg(A.accepted) == 0) return;
  // ---- accept: state <- proposal
  if (rank == 0)
    sm_apply_params_body(A.info, NSB, A.Kcap, pp, A.cen, A.sig, A.isg, A.sden, A.cnt, A.counts, A.Kptr, A.phi.status);
  {
    const int Kold = A.info->K;
    for (int m = tid; m < mine; m += SMC_T) A.c[srow[m]] = same ? (zc[m] == 0 ? Kold : cB) : cB;
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
  int CS = 0, mcap = 0, rcap = 0;
  size_t smem = 0;
  double* gsvals = nullptr;
  double *nr_d0 = nullptr, *nr_lg = nullptr;
  int *nr_pre = nullptr, *nr_z = nullptr, *nr_idx = nullptr, *walk_out = nullptr;
  unsigned long long* prof = nullptr;
};

// decides whether the cluster kernel can run this chain (shared-memory budget, cluster size the device accepts)
static int smc_setup(smg_chain* ch, SmcHost* H) {
  const char* env = getenv("SMG_SM_MODE");  // cluster (default when feasible) | coop | multi
  if (env && strcmp(env, "cluster") != 0) return 0;
  int cc_major = 0, smem_optin = 0;
  cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, ch->device);
  cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ch->device);
  if (cc_major < 9 || ch->t <= 0) return 0;
  const size_t budget = (size_t)smem_optin - 4096;  // static shared memory of the kernel
  const char* ecs = getenv("SMG_SMC_CS");
  for (int CS = ecs ? atoi(ecs) : SMC_MAXCS; CS >= 8; CS >>= 1) {
    if (ch->pp % CS) continue;
    const int mcap = ((int)(((long long)ch->n + 2 + CS - 1) / CS) + 31) & ~31;
    const SmcLayout L0 = smc_layout(ch->pp, ch->mmax, CS, mcap, 0);
    if (L0.total + (size_t)std::min(mcap, 64) * ch->pp > budget) continue;
    const int rcap = (int)std::min<size_t>((size_t)mcap, (budget - L0.total) / ch->pp);
    const SmcLayout L = smc_layout(ch->pp, ch->mmax, CS, mcap, rcap);
    if (cudaFuncSetAttribute(sm_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total) != cudaSuccess ||
        cudaFuncSetAttribute(sm_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
      (void)cudaGetLastError();
      continue;
    }
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
  SMG_CUDA(dev_malloc(&H->prof, 16 * 8, ch->st));
  SMG_CUDA(cudaMemsetAsync(H->prof, 0, 16 * 8, ch->st));
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
  G.sl = ch->pp / H->CS;
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
