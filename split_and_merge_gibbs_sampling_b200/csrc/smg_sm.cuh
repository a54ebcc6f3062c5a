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

struct SmWork {
  int *S = nullptr, *zL = nullptr, *zStar = nullptr, *zState = nullptr;
  SmInfo* info = nullptr;
  SmPlan* plan = nullptr;
  int* anchors = nullptr;  // alias of info->i1,i2 (two ints)
  int *H = nullptr, *cnt = nullptr;
  double* LL2 = nullptr;      // [n][2]
  double *rg_dl = nullptr, *rg_lgt = nullptr, *LT = nullptr;  // [n], [n], [n+3]
  double* rowvals = nullptr;  // [4][n+2]
  double* partial = nullptr;  // [4][RB]
  double* terms = nullptr;    // [24]
  PhiJob* jobs = nullptr;     // [9]
  // injected uniforms (device copies, allocated on first use)
  double *u_pair = nullptr, *u_prior_c = nullptr, *u_prior_s = nullptr, *u_launch = nullptr, *u_rg = nullptr;
  double *u_rg_c = nullptr, *u_rg_s = nullptr, *u_mg_c = nullptr, *u_mg_s = nullptr, *u_accept = nullptr;
  bool inj_alloc = false;
};

#define SM_RB 256  // partial-sum blocks of the row reductions

// ------------------------------------------------------------------------------------------
// pair selection + S (split_merge.cpp:263-301).  One CTA; ordered stream compaction.
// sample(indices, 2, replace=false): j1=(int)(n*u0); j2=(int)((n-1)*u1) over the swapped array.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) sm_select_kernel(int n, const int* __restrict__ c, const int* __restrict__ Kptr,
                                                         const double* u_pair, RngKey key, int NS, int* __restrict__ S,
                                                         int* __restrict__ zState, SmInfo* info, SmPlan* plan,
                                                         int* __restrict__ cnt, double* terms) {
  __shared__ int s_i1, s_i2, s_cA, s_cB, s_base, s_tot;
  __shared__ int s_woff[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    double u0 = get_u(u_pair, 0, key, U_SM_PAIR, 0u, 0u), u1 = get_u(u_pair, 1, key, U_SM_PAIR, 1u, 0u);
    int j1 = (int)((double)n * u0);
    if (j1 >= n) j1 = n - 1;
    int j2 = (int)((double)(n - 1) * u1);
    if (j2 >= n - 1) j2 = n - 2;
    int i1 = j1, i2 = (j2 == j1) ? (n - 1) : j2;
    s_i1 = i1;
    s_i2 = i2;
    s_cA = c[i1];
    s_cB = c[i2];
    s_base = 0;
  }
  for (int q = tid; q < 24; q += blockDim.x) terms[q] = 0.0;
  __syncthreads();
  const int i1 = s_i1, i2 = s_i2, cA = s_cA, cB = s_cB;
  for (int t0 = 0; t0 < n; t0 += 1024) {
    int i = t0 + tid;
    int ci = (i < n) ? c[i] : -1;
    bool in = (i < n) && i != i1 && i != i2 && (ci == cA || ci == cB);
    unsigned b = __ballot_sync(SMG_FULL, in);
    if (lane == 0) s_woff[warp] = __popc(b);
    __syncthreads();
    if (warp == 0) {
      int v = s_woff[lane], x = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(SMG_FULL, x, o);
        if (lane >= o) x += y;
      }
      s_woff[lane] = x - v;  // exclusive
      if (lane == 31) s_tot = x;
    }
    __syncthreads();
    if (in) {
      int pos = s_base + s_woff[warp] + __popc(b & ((1u << lane) - 1));
      S[pos] = i;
      zState[pos] = (ci == cA) ? 0 : 1;
    }
    __syncthreads();
    if (tid == 0) s_base += s_tot;
    __syncthreads();
  }
  if (tid == 0) {
    const int nS = s_base, same = (cA == cB), K = *Kptr;
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

// random launch allocation (split_merge.cpp:346): sample({a,b}, |S|, replace) -> (int)(2u)
__global__ void sm_launch_alloc_kernel(const SmInfo* info, const double* u_inj, RngKey key, int* __restrict__ zL) {
  int pos = blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= info->nS) return;
  double u = get_u(u_inj, pos, key, U_SM_LAUNCH, (uint32_t)pos, 0u);
  int z = (int)(2.0 * u);
  zL[pos] = z > 1 ? 1 : z;
}

// LL2[pos][g] = log-likelihood of member S[pos] under parameter slot (slotA, slotB)
__global__ void __launch_bounds__(256) sm_ll2_kernel(const uint8_t* __restrict__ X, int pp, const int* __restrict__ S,
                                                     const SmInfo* info, const uint8_t* cen, const double* isg,
                                                     const double* sden, int slotA, int slotB,
                                                     double* __restrict__ LL2, const int* enable, int enable_val) {
  if (enable && *enable != enable_val) return;
  const int nS = info->nS;
  long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= 2ll * nS) return;
  const int pos = (int)(w >> 1), g = (int)(w & 1);
  const int slot = g ? slotB : slotA;
  double dot = warp_mismatch_dot(X + (size_t)S[pos] * pp, cen + (size_t)slot * pp, isg + (size_t)slot * pp, pp, lane);
  if (lane == 0) LL2[w] = -dot - sden[slot];
}

// log table LT[k] = log(k), k = 0..n+2 (member counts of the restricted scans)
__global__ void sm_logtable_kernel(int len, double* __restrict__ LT) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < len) LT[k] = k > 0 ? log((double)k) : -CUDART_INF;
}

// Restricted Gibbs allocation scan over S (split_merge.cpp:186-216), in two phases.
//
// The two-way Rcpp::sample (split_merge.cpp:215) walks the two probabilities in descending order
// (revsort; on a tie the SECOND entry comes first) and takes the first when u <= p_first.  With
// D = (log n_1 + LL_1) - (log n_2 + LL_2) the larger probability is 1/(1+exp(-|D|)), so the member goes
// to the larger side iff |D| >= logit(u), to the other side otherwise.  logit(u) and LL_1 - LL_2 do not
// depend on the running counts, so phase 1 evaluates them for every member in parallel, and the
// sequential phase 2 only needs a table look-up of log(count) and two comparisons per member.
// Phase 2 runs on one warp: 32 consecutive members are decided against the same counts, the first one
// that changes side is applied, and the evaluation restarts after it (same result as one at a time).
__global__ void __launch_bounds__(256) sm_rg_prepare_kernel(const SmInfo* info, const double* __restrict__ LL2,
                                                            const double* u_inj, RngKey key, double* __restrict__ dl,
                                                            double* __restrict__ lgt, const int* enable, int enable_val) {
  if (enable && *enable != enable_val) return;
  int pos = blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= info->nS) return;
  const double u = get_u(u_inj, pos, key, U_SM_RGIBBS, (uint32_t)pos, 0u);
  dl[pos] = LL2[2 * (size_t)pos] - LL2[2 * (size_t)pos + 1];
  lgt[pos] = log(u / (1.0 - u));
}

__global__ void __launch_bounds__(32) sm_rgibbs_kernel(const SmInfo* info, const double* __restrict__ dl,
                                                       const double* __restrict__ lgt, const double* __restrict__ LT,
                                                       int* __restrict__ z, const int* enable, int enable_val) {
  if (enable && *enable != enable_val) return;
  const int nS = info->nS, lane = threadIdx.x;
  // side counts including the anchors i_1 (side 0) and i_2 (side 1)
  int c1 = 0;
  for (int pos = lane; pos < nS; pos += 32) c1 += z[pos];
  c1 = warp_sum_i(c1);
  int nA = 1 + (nS - c1), nB = 1 + c1;
  for (int base = 0; base < nS; base += 32) {
    const int pos = base + lane;
    const bool valid = pos < nS;
    double d0 = 0.0, lg = 0.0;
    int zz = 0;
    if (valid) {
      d0 = dl[pos];
      lg = lgt[pos];
      zz = z[pos];
    }
    int start = 0;
    while (start < 32) {
      // log-count difference seen by a member currently on side 0 / side 1 (itself excluded)
      const double dA = LT[nA - 1] - LT[nB], dB = LT[nA] - LT[nB - 1];
      int newz = zz;
      if (valid && lane >= start) {
        const double D = (zz == 0 ? dA : dB) + d0;
        if (D > 0.0)
          newz = (D >= lg) ? 0 : 1;
        else
          newz = (-D >= lg) ? 1 : 0;
      }
      unsigned ch = __ballot_sync(SMG_FULL, valid && lane >= start && newz != zz);
      if (!ch) break;
      const int f = __ffs(ch) - 1;
      const int zo = __shfl_sync(SMG_FULL, zz, f), zn = __shfl_sync(SMG_FULL, newz, f);
      if (lane == f) zz = newz;
      nA += (zn == 0) - (zo == 0);
      nB += (zn == 1) - (zo == 1);
      start = f + 1;
    }
    if (valid) z[pos] = zz;
  }
}

__global__ void sm_copy_z_kernel(const SmInfo* info, const int* __restrict__ src, int* __restrict__ dst) {
  int pos = blockIdx.x * blockDim.x + threadIdx.x;
  if (pos < info->nS) dst[pos] = src[pos];
}

// copy a parameter slot (centre, sigma, 1/sigma, sum of log-normalisers)
__global__ void sm_copy_slot_kernel(int pp, uint8_t* cen, double* sig, double* isg, double* sden, int src, int dst) {
  for (int j = threadIdx.x; j < pp; j += blockDim.x) {
    cen[(size_t)dst * pp + j] = cen[(size_t)src * pp + j];
    sig[(size_t)dst * pp + j] = sig[(size_t)src * pp + j];
    isg[(size_t)dst * pp + j] = isg[(size_t)src * pp + j];
  }
  if (threadIdx.x == 0) sden[dst] = sden[src];
}

// H[dst] = H[a] + H[b] ; cnt likewise
__global__ void sm_hist_add_kernel(int len, int* H, int* cnt, int a, int b, int dst) {
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < len) H[(size_t)dst * len + q] = H[(size_t)a * len + q] + H[(size_t)b * len + q];
  if (q == 0) cnt[dst] = cnt[a] + cnt[b];
}

// ------------------------------------------------------------------------------------------
// MH terms
// ------------------------------------------------------------------------------------------
// logprobgs_phi (split_merge.cpp:20-94): block b evaluates term b; also priors (:419-436) in blocks 3..5
__global__ void __launch_bounds__(256) sm_gsphi_prior_kernel(int pp, int p, int mmax, const int* __restrict__ attr,
                                                             const double* __restrict__ v, const double* __restrict__ w,
                                                             const int* __restrict__ H, const int* __restrict__ cnt,
                                                             const SmPlan* plan, const uint8_t* cen, const double* sig,
                                                             double* terms) {
  __shared__ double sh[256];
  const int b = blockIdx.x;
  const int len = pp * mmax;
  double acc = 0.0;
  if (b < 3) {
    const int hist = plan->gs_hist[b], sL = plan->gs_sigL[b], st = plan->gs_star[b];
    const int nm = cnt[hist];
    for (int j = threadIdx.x; j < p; j += 256) {
      const int m = attr[j];
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
      acc += lc + ls;
    }
  } else {
    const int slot = plan->pri_slot[b - 3];
    for (int j = threadIdx.x; j < p; j += 256) {
      const int m = attr[j];
      acc += -log((double)m) + logdensity_hig_d(sig[(size_t)slot * pp + j], v[j], w[j], (double)m);
    }
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  // terms layout: [4..6] pri, [10..12] gs_phi
  if (threadIdx.x == 0) terms[b < 3 ? 10 + b : 4 + (b - 3)] = sh[0];
}

// per-member likelihood values feeding loglikelihood_hamming (:393-417) and logprobgs_c_i (:96-161)
// rowvals[0][r] = LL under its own side's parameters if side 0 else 0
// rowvals[1][r] = same for side 1 ; rowvals[2][r] = LL under the merged/old parameters
// rowvals[3][r] = log prob of the member's side under the launch counts (0 for the anchors)
__global__ void __launch_bounds__(256) sm_rowterms_kernel(const uint8_t* __restrict__ X, int pp, const int* __restrict__ S,
                                                          const SmInfo* info, const SmPlan* plan,
                                                          const int* __restrict__ zL, const int* __restrict__ zStar,
                                                          const int* __restrict__ zState, const int* __restrict__ cnt,
                                                          const uint8_t* cen, const double* isg, const double* sden,
                                                          double* __restrict__ rowvals, int stride) {
  const int nS = info->nS;
  long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= nS + 2) return;
  const int r = (int)w;
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

// fixed-order two-stage sums of the four row-value arrays
__global__ void __launch_bounds__(256) sm_rowreduce1_kernel(const SmInfo* info, const double* __restrict__ rowvals,
                                                            int stride, double* __restrict__ partial) {
  __shared__ double sh[256];
  const int nr = info->nS + 2, q = blockIdx.y;
  const int per = (nr + SM_RB - 1) / SM_RB;
  const int lo = blockIdx.x * per, hi = min(nr, lo + per);
  double acc = 0.0;
  for (int r = lo + threadIdx.x; r < hi; r += 256) acc += rowvals[(size_t)q * stride + r];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[q * SM_RB + blockIdx.x] = sh[0];
}

// final sums + MH log-ratio in the reference's order of additions + accept decision
__global__ void __launch_bounds__(256) sm_accept_kernel(const SmInfo* info, const SmPlan* plan, const int* __restrict__ cnt,
                                                        const double* __restrict__ partial, double gamma,
                                                        const double* u_inj, RngKey key, double* terms, int* accepted,
                                                        unsigned long long* stats) {
  __shared__ double sh[256];
  __shared__ double tot[4];
  for (int q = 0; q < 4; q++) {
    sh[threadIdx.x] = threadIdx.x < SM_RB ? partial[q * SM_RB + threadIdx.x] : 0.0;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) tot[q] = sh[0];
    __syncthreads();
  }
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

// ------------------------------------------------------------------------------------------
// accept: state <- proposal, labels compacted as clean_var does (common_functions.cpp:296-353)
// ------------------------------------------------------------------------------------------
// step 1 (one CTA): parameters, counts, K
__global__ void __launch_bounds__(256) sm_apply_params_kernel(const SmInfo* info, const int* accepted, int NS, int Kcap,
                                                              int pp, uint8_t* cen, double* sig, double* isg, double* sden,
                                                              const int* __restrict__ cnt, int* counts, int* Kptr,
                                                              int* status) {
  if (*accepted == 0) return;
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
// step 2: labels of the members
__global__ void sm_apply_members_kernel(const SmInfo* info, const int* accepted, const int* __restrict__ S,
                                        const int* __restrict__ zStar, int* __restrict__ c) {
  if (*accepted == 0) return;
  const int nS = info->nS;
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nS + 2) return;
  const int row = r < nS ? S[r] : (r == nS ? info->i1 : info->i2);
  if (info->same) {
    const int side = r < nS ? zStar[r] : (r - nS);
    c[row] = side == 0 ? info->K : info->cB;
  } else {
    c[row] = info->cB;
  }
}
// step 3 (merge only): the last label moves into the hole
__global__ void sm_apply_relabel_kernel(const SmInfo* info, const int* accepted, int n, int* __restrict__ c) {
  if (*accepted == 0 || info->same) return;
  const int hole = info->cA, last = info->K - 1;
  if (hole == last) return;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && c[i] == last) c[i] = hole;
}

// ------------------------------------------------------------------------------------------
// host orchestration
// ------------------------------------------------------------------------------------------
static inline int sm_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

static int sm_alloc(smg_chain* ch) {
  SmWork* W = new SmWork();
  ch->sm = W;
  const int n = ch->n;
  SMG_CUDA(cudaMalloc(&W->S, (size_t)n * 4));
  SMG_CUDA(cudaMalloc(&W->zL, (size_t)n * 4));
  SMG_CUDA(cudaMalloc(&W->zStar, (size_t)n * 4));
  SMG_CUDA(cudaMalloc(&W->zState, (size_t)n * 4));
  SMG_CUDA(cudaMalloc(&W->info, sizeof(SmInfo)));
  SMG_CUDA(cudaMemset(W->info, 0, sizeof(SmInfo)));
  SMG_CUDA(cudaMalloc(&W->plan, sizeof(SmPlan)));
  SMG_CUDA(cudaMalloc(&W->H, (size_t)SH_N * ch->pp * ch->mmax * 4));
  SMG_CUDA(cudaMalloc(&W->cnt, SH_N * 4 + 4));
  SMG_CUDA(cudaMalloc(&W->LL2, (size_t)n * 2 * 8));
  SMG_CUDA(cudaMalloc(&W->rg_dl, (size_t)n * 8));
  SMG_CUDA(cudaMalloc(&W->rg_lgt, (size_t)n * 8));
  SMG_CUDA(cudaMalloc(&W->LT, (size_t)(n + 3) * 8));
  sm_logtable_kernel<<<(n + 3 + 255) / 256, 256, 0, ch->st>>>(n + 3, W->LT);
  SMG_CUDA(cudaGetLastError());
  SMG_CUDA(cudaMalloc(&W->rowvals, (size_t)4 * (n + 2) * 8));
  SMG_CUDA(cudaMalloc(&W->partial, (size_t)4 * SM_RB * 8));
  SMG_CUDA(cudaMalloc(&W->terms, 24 * 8));
  SMG_CUDA(cudaMemset(W->terms, 0, 24 * 8));
  SMG_CUDA(cudaMalloc(&W->jobs, 9 * sizeof(PhiJob)));
  const int B = ch->NS;
  PhiJob j[9] = {{0, 0, B + SM_SL_A, 0},
                 {0, 0, B + SM_SL_B, 0},
                 {0, 0, B + SM_ML_M, 0},
                 {SH_L0, B + SM_SL_A, B + SM_SL_A, SH_L0},
                 {SH_L1, B + SM_SL_B, B + SM_SL_B, SH_L1},
                 {SH_M, B + SM_ML_M, B + SM_ML_M, SH_M},
                 {SH_P0, B + SM_ST_A, B + SM_ST_A, SH_P0},
                 {SH_P1, B + SM_ST_B, B + SM_ST_B, SH_P1},
                 {SH_M, B + SM_ML_M, B + SM_ST_M, SH_M}};
  SMG_CUDA(cudaMemcpy(W->jobs, j, sizeof(j), cudaMemcpyHostToDevice));
  return 0;
}

static void sm_free(smg_chain* ch) {
  SmWork* W = ch->sm;
  if (!W) return;
  void* ptrs[] = {W->S,      W->zL,     W->zStar,    W->zState,   W->info,   W->plan, W->H,      W->cnt,
                  W->LL2,    W->rg_dl, W->rg_lgt, W->LT, W->rowvals, W->partial, W->terms,    W->jobs,   W->u_pair, W->u_prior_c, W->u_prior_s,
                  W->u_launch, W->u_rg, W->u_rg_c,   W->u_rg_s,   W->u_mg_c, W->u_mg_s, W->u_accept};
  for (void* q : ptrs)
    if (q) cudaFree(q);
  delete W;
  ch->sm = nullptr;
}

static int sm_inject(smg_chain* ch, const smg_sm_tape* t) {
  SmWork* W = ch->sm;
  const int n = ch->n, p = ch->p, T1 = ch->t + 1, R1 = ch->r + 1;
  if (!W->inj_alloc) {
    SMG_CUDA(cudaMalloc(&W->u_pair, 2 * 8));
    SMG_CUDA(cudaMalloc(&W->u_prior_c, (size_t)3 * p * 8));
    SMG_CUDA(cudaMalloc(&W->u_prior_s, (size_t)3 * p * 8));
    SMG_CUDA(cudaMalloc(&W->u_launch, (size_t)n * 8));
    SMG_CUDA(cudaMalloc(&W->u_rg, (size_t)T1 * n * 8));
    SMG_CUDA(cudaMalloc(&W->u_rg_c, (size_t)T1 * 2 * p * 8));
    SMG_CUDA(cudaMalloc(&W->u_rg_s, (size_t)T1 * 2 * p * 8));
    SMG_CUDA(cudaMalloc(&W->u_mg_c, (size_t)R1 * p * 8));
    SMG_CUDA(cudaMalloc(&W->u_mg_s, (size_t)R1 * p * 8));
    SMG_CUDA(cudaMalloc(&W->u_accept, 8));
    W->inj_alloc = true;
  }
  auto up = [&](double* d, const double* h, size_t cnt) -> int {
    if (h) SMG_CUDA(cudaMemcpy(d, h, cnt * 8, cudaMemcpyHostToDevice));
    return 0;
  };
  if (up(W->u_pair, t->u_pair, 2) || up(W->u_prior_c, t->u_prior_c, (size_t)3 * p) ||
      up(W->u_prior_s, t->u_prior_s, (size_t)3 * p) || up(W->u_launch, t->u_launch, n) ||
      up(W->u_rg, t->u_rg, (size_t)T1 * n) || up(W->u_rg_c, t->u_rg_c, (size_t)T1 * 2 * p) ||
      up(W->u_rg_s, t->u_rg_s, (size_t)T1 * 2 * p) || up(W->u_mg_c, t->u_mg_c, (size_t)R1 * p) ||
      up(W->u_mg_s, t->u_mg_s, (size_t)R1 * p) || up(W->u_accept, t->u_accept, 1))
    return SMG_ERR_CUDA;
  return 0;
}

// phi_draw + sden on `nj` consecutive jobs starting at jobs[j0]
static int sm_phi(smg_chain* ch, int j0, int nj, int prior, uint32_t sub, const double* uc, const double* us,
                  const int* enable) {
  SmWork* W = ch->sm;
  const int cur = ch->cur;
  PhiArgs A;
  A.pp = ch->pp;
  A.p = ch->p;
  A.mmax = ch->mmax;
  A.attr = ch->attr;
  A.v = ch->v;
  A.w = ch->w;
  A.H = W->H;
  A.counts = W->cnt;
  A.jobs = W->jobs + j0;
  A.njobs_ptr = nullptr;
  A.njobs = nj;
  A.cen_src = ch->cen[cur];
  A.sig_src = ch->sig[cur];
  A.cen = ch->cen[cur];
  A.sig = ch->sig[cur];
  A.isg = ch->isg[cur];
  A.den = ch->den;
  A.u_center = uc;
  A.u_sigma = us;
  A.u_stride = ch->p;
  A.key = mk_key(ch, sub);
  A.prior = prior;
  A.sigma_exact = ch->sigma_exact;
  A.enable = enable;
  A.status = ch->status;
  dim3 grid(sm_cdiv(ch->pp, 128), nj);
  phi_draw_kernel<<<grid, 128, 0, ch->st>>>(A);
  phi_sden_kernel<<<nj, 256, 0, ch->st>>>(W->jobs + j0, nullptr, nj, ch->pp, ch->den, ch->sden[cur], enable);
  ch->h_launches += 2;
  SMG_CUDA(cudaGetLastError());
  return 0;
}

// histogram of S u {i1,i2} split by z into H[h0], H[h0+1] (z == nullptr: everything into H[h0])
static int sm_hist(smg_chain* ch, const int* z, int h0) {
  SmWork* W = ch->sm;
  const size_t len = (size_t)ch->pp * ch->mmax;
  const int nh = z ? 2 : 1;
  SMG_CUDA(cudaMemsetAsync(W->H + (size_t)h0 * len, 0, nh * len * 4, ch->st));
  SMG_CUDA(cudaMemsetAsync(W->cnt + h0, 0, nh * 4, ch->st));
  long long threads = (long long)(ch->n) * (ch->pp / 16);  // upper bound; the kernel trims to |S|+2
  subset_histogram_kernel<<<sm_cdiv(threads, 256), 256, 0, ch->st>>>(ch->X, ch->pp, W->S, &W->info->nS, z,
                                                                    &W->info->i1, ch->mmax, W->H + (size_t)h0 * len,
                                                                    W->cnt + h0);
  ch->h_launches++;
  SMG_CUDA(cudaGetLastError());
  return 0;
}

// one restricted scan + update_phi of the two sides on (z, slots A/B, histograms h0/h0+1, jobs j0..j0+1)
static int sm_restricted_scan(smg_chain* ch, int* z, int slotA, int slotB, int h0, int j0, int q, const double* u_rg,
                              const double* uc, const double* us, const int* enable) {
  SmWork* W = ch->sm;
  const int cur = ch->cur;
  sm_ll2_kernel<<<sm_cdiv(2ll * ch->n * 32, 256), 256, 0, ch->st>>>(ch->X, ch->pp, W->S, W->info, ch->cen[cur],
                                                                  ch->isg[cur], ch->sden[cur], slotA, slotB, W->LL2,
                                                                  enable, 1);
  sm_rg_prepare_kernel<<<sm_cdiv(ch->n, 256), 256, 0, ch->st>>>(W->info, W->LL2, u_rg, mk_key(ch, SUB_SM_RG + q), W->rg_dl,
                                                               W->rg_lgt, enable, 1);
  sm_rgibbs_kernel<<<1, 32, 0, ch->st>>>(W->info, W->rg_dl, W->rg_lgt, W->LT, z, enable, 1);
  ch->h_launches += 3;
  SMG_CUDA(cudaGetLastError());
  if (sm_hist(ch, z, h0)) return SMG_ERR_CUDA;
  return sm_phi(ch, j0, 2, 0, SUB_SM_RG + q, uc, us, enable);
}

// split_and_merge (split_merge.cpp:542-598)
static int sm_step(smg_chain* ch, const smg_sm_tape* tape) {
  SmWork* W = ch->sm;
  const int n = ch->n, p = ch->p, pp = ch->pp, B = ch->NS, cur = ch->cur;
  smg_sm_tape T;
  memset(&T, 0, sizeof(T));
  if (tape) {
    int rc = sm_inject(ch, tape);
    if (rc) return rc;
    T.u_pair = tape->u_pair ? W->u_pair : nullptr;
    T.u_prior_c = tape->u_prior_c ? W->u_prior_c : nullptr;
    T.u_prior_s = tape->u_prior_s ? W->u_prior_s : nullptr;
    T.u_launch = tape->u_launch ? W->u_launch : nullptr;
    T.u_rg = tape->u_rg ? W->u_rg : nullptr;
    T.u_rg_c = tape->u_rg_c ? W->u_rg_c : nullptr;
    T.u_rg_s = tape->u_rg_s ? W->u_rg_s : nullptr;
    T.u_mg_c = tape->u_mg_c ? W->u_mg_c : nullptr;
    T.u_mg_s = tape->u_mg_s ? W->u_mg_s : nullptr;
    T.u_accept = tape->u_accept ? W->u_accept : nullptr;
  }
  const size_t len = (size_t)pp * ch->mmax;
  // ---- pair, S, plan
  sm_select_kernel<<<1, 1024, 0, ch->st>>>(n, ch->c, ch->K, T.u_pair, mk_key(ch, SUB_SM_SELECT), B, W->S, W->zState,
                                           W->info, W->plan, W->cnt, W->terms);
  ch->h_launches++;
  SMG_CUDA(cudaGetLastError());
  const int* same = &W->info->same;
  // ---- prior parameters of the three launch clusters (split_merge.cpp:331-343, :379-380)
  if (sm_phi(ch, 0, 3, 1, SUB_SM_PRIOR, T.u_prior_c, T.u_prior_s, nullptr)) return SMG_ERR_CUDA;
  // ---- split launch: random sides then t restricted scans (split_merge.cpp:346-349)
  sm_launch_alloc_kernel<<<sm_cdiv(n, 256), 256, 0, ch->st>>>(W->info, T.u_launch, mk_key(ch, SUB_SM_LAUNCH), W->zL);
  ch->h_launches++;
  for (int q = 0; q < ch->t; q++) {
    int rc = sm_restricted_scan(ch, W->zL, B + SM_SL_A, B + SM_SL_B, SH_L0, 3, q, T.u_rg ? T.u_rg + (size_t)q * n : nullptr,
                                T.u_rg_c ? T.u_rg_c + (size_t)q * 2 * p : nullptr,
                                T.u_rg_s ? T.u_rg_s + (size_t)q * 2 * p : nullptr, nullptr);
    if (rc) return rc;
  }
  if (ch->t == 0 && sm_hist(ch, W->zL, SH_L0)) return SMG_ERR_CUDA;  // launch counts are still needed
  // ---- histograms of the current-state sides and of the merged cluster
  if (sm_hist(ch, W->zState, SH_S0)) return SMG_ERR_CUDA;
  sm_hist_add_kernel<<<sm_cdiv(len, 256), 256, 0, ch->st>>>((int)len, W->H, W->cnt, SH_S0, SH_S1, SH_M);
  ch->h_launches++;
  // ---- merge launch: r parameter updates of the merged cluster (split_merge.cpp:386-387)
  for (int q = 0; q < ch->r; q++) {
    int rc = sm_phi(ch, 5, 1, 0, SUB_SM_MERGE + q, T.u_mg_c ? T.u_mg_c + (size_t)q * p : nullptr,
                    T.u_mg_s ? T.u_mg_s + (size_t)q * p : nullptr, nullptr);
    if (rc) return rc;
  }
  // ---- proposal
  //   split (same == 1): star = split launch + one more restricted scan (split_merge.cpp:575-580)
  sm_copy_z_kernel<<<sm_cdiv(n, 256), 256, 0, ch->st>>>(W->info, W->zL, W->zStar);
  sm_copy_slot_kernel<<<1, 256, 0, ch->st>>>(pp, ch->cen[cur], ch->sig[cur], ch->isg[cur], ch->sden[cur], B + SM_SL_A,
                                             B + SM_ST_A);
  sm_copy_slot_kernel<<<1, 256, 0, ch->st>>>(pp, ch->cen[cur], ch->sig[cur], ch->isg[cur], ch->sden[cur], B + SM_SL_B,
                                             B + SM_ST_B);
  ch->h_launches += 3;
  {
    const int q = ch->t;
    int rc = sm_restricted_scan(ch, W->zStar, B + SM_ST_A, B + SM_ST_B, SH_P0, 6, q,
                                T.u_rg ? T.u_rg + (size_t)q * n : nullptr,
                                T.u_rg_c ? T.u_rg_c + (size_t)q * 2 * p : nullptr,
                                T.u_rg_s ? T.u_rg_s + (size_t)q * 2 * p : nullptr, same);
    if (rc) return rc;
  }
  //   merge (same == 0): star = merge launch + one more update_phi (split_merge.cpp:582-586).
  //   The update is drawn in both cases (it only fills the M* slot); the MH kernel ignores it for a split.
  {
    const int q = ch->r;
    int rc = sm_phi(ch, 8, 1, 0, SUB_SM_MERGE + q, T.u_mg_c ? T.u_mg_c + (size_t)q * p : nullptr,
                    T.u_mg_s ? T.u_mg_s + (size_t)q * p : nullptr, nullptr);
    if (rc) return rc;
  }
  // ---- MH terms
  sm_gsphi_prior_kernel<<<6, 256, 0, ch->st>>>(pp, p, ch->mmax, ch->attr, ch->v, ch->w, W->H, W->cnt, W->plan, ch->cen[cur],
                                               ch->sig[cur], W->terms);
  sm_rowterms_kernel<<<sm_cdiv((long long)(n + 2) * 32, 256), 256, 0, ch->st>>>(
      ch->X, pp, W->S, W->info, W->plan, W->zL, W->zStar, W->zState, W->cnt, ch->cen[cur], ch->isg[cur], ch->sden[cur],
      W->rowvals, n + 2);
  sm_rowreduce1_kernel<<<dim3(SM_RB, 4), 256, 0, ch->st>>>(W->info, W->rowvals, n + 2, W->partial);
  sm_accept_kernel<<<1, 256, 0, ch->st>>>(W->info, W->plan, W->cnt, W->partial, ch->gamma, T.u_accept,
                                          mk_key(ch, SUB_SM_ACCEPT), W->terms, ch->accepted_d, ch->stats_d);
  // ---- accept: state <- proposal
  sm_apply_params_kernel<<<1, 256, 0, ch->st>>>(W->info, ch->accepted_d, B, ch->Kcap, pp, ch->cen[cur], ch->sig[cur],
                                                ch->isg[cur], ch->sden[cur], W->cnt, ch->counts, ch->K, ch->status);
  sm_apply_members_kernel<<<sm_cdiv(n + 2, 256), 256, 0, ch->st>>>(W->info, ch->accepted_d, W->S, W->zStar, ch->c);
  sm_apply_relabel_kernel<<<sm_cdiv(n, 256), 256, 0, ch->st>>>(W->info, ch->accepted_d, n, ch->c);
  ch->h_launches += 7;
  SMG_CUDA(cudaGetLastError());
  return 0;
}

static int sm_readback(smg_chain* ch, int* info, int* S, int* z_launch, int* z_star, double* phi_out, double* terms) {
  SmWork* W = ch->sm;
  SMG_CUDA(cudaStreamSynchronize(ch->st));
  SmInfo I;
  SMG_CUDA(cudaMemcpy(&I, W->info, sizeof(I), cudaMemcpyDeviceToHost));
  int cnt[SH_N];
  SMG_CUDA(cudaMemcpy(cnt, W->cnt, sizeof(cnt), cudaMemcpyDeviceToHost));
  int acc = 0, K = 0;
  SMG_CUDA(cudaMemcpy(&acc, ch->accepted_d, 4, cudaMemcpyDeviceToHost));
  SMG_CUDA(cudaMemcpy(&K, ch->K, 4, cudaMemcpyDeviceToHost));
  if (info) {
    info[0] = I.i1;
    info[1] = I.i2;
    info[2] = I.nS;
    info[3] = I.same;
    info[4] = acc;
    info[5] = cnt[SH_P0];
    info[6] = cnt[SH_P1];
    info[7] = K;
  }
  if (S) SMG_CUDA(cudaMemcpy(S, W->S, (size_t)I.nS * 4, cudaMemcpyDeviceToHost));
  if (z_launch) SMG_CUDA(cudaMemcpy(z_launch, W->zL, (size_t)I.nS * 4, cudaMemcpyDeviceToHost));
  if (z_star) SMG_CUDA(cudaMemcpy(z_star, W->zStar, (size_t)I.nS * 4, cudaMemcpyDeviceToHost));
  if (phi_out) {
    const int slots[6] = {SM_SL_A, SM_SL_B, SM_ML_M, SM_ST_A, SM_ST_B, SM_ST_M};
    std::vector<uint8_t> hc(ch->pp);
    std::vector<double> hs(ch->pp);
    for (int q = 0; q < 6; q++) {
      const size_t off = (size_t)(ch->NS + slots[q]) * ch->pp;
      SMG_CUDA(cudaMemcpy(hc.data(), ch->cen[ch->cur] + off, ch->pp, cudaMemcpyDeviceToHost));
      SMG_CUDA(cudaMemcpy(hs.data(), ch->sig[ch->cur] + off, (size_t)ch->pp * 8, cudaMemcpyDeviceToHost));
      for (int j = 0; j < ch->p; j++) {
        phi_out[((size_t)q * 2 + 0) * ch->p + j] = hc[j];
        phi_out[((size_t)q * 2 + 1) * ch->p + j] = hs[j];
      }
    }
  }
  if (terms) SMG_CUDA(cudaMemcpy(terms, W->terms, 24 * 8, cudaMemcpyDeviceToHost));
  return 0;
}

}  // namespace smg
