// smg_sm_host.cuh -- host orchestration of the split-merge step (code/split_merge.cpp:542-598): picks one of the
// three device paths -- the cluster kernel (smg_smc.cuh), the cooperative persistent kernel or the sequence of small
// launches (smg_sm.cuh) -- and exposes the parity hooks' read-back.
#pragma once
#include "smg_sm.cuh"
#include "smg_smc.cuh"

namespace smg {

// ------------------------------------------------------------------------------------------
// host orchestration
// ------------------------------------------------------------------------------------------
static inline int sm_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

static int sm_alloc(smg_chain* ch) {
  SmWork* W = new SmWork();
  ch->sm = W;
  const int n = ch->n;
  SMG_CUDA(dev_malloc(&W->S, (size_t)n * 4, ch->st));
  SMG_CUDA(dev_malloc(&W->zL, (size_t)n * 4, ch->st));
  SMG_CUDA(dev_malloc(&W->zStar, (size_t)n * 4, ch->st));
  SMG_CUDA(dev_malloc(&W->zState, (size_t)n * 4, ch->st));
  SMG_CUDA(dev_malloc(&W->info, sizeof(SmInfo), ch->st));
  SMG_CUDA(cudaMemsetAsync(W->info, 0, sizeof(SmInfo), ch->st));
  SMG_CUDA(dev_malloc(&W->plan, sizeof(SmPlan), ch->st));
  SMG_CUDA(dev_malloc(&W->H, (size_t)SH_N * ch->pp * ch->mmax * 4, ch->st));
  SMG_CUDA(dev_malloc(&W->cnt, SH_N * 4 + 4, ch->st));
  SMG_CUDA(dev_malloc(&W->rg_dl, (size_t)n * 8, ch->st));
  SMG_CUDA(dev_malloc(&W->rg_lgt, (size_t)n * 8, ch->st));
  SMG_CUDA(dev_malloc(&W->rg_lgt2, (size_t)n * 8, ch->st));
  SMG_CUDA(dev_malloc(&W->rowvals, (size_t)4 * (n + 2) * 8, ch->st));
  SMG_CUDA(dev_malloc(&W->partial, (size_t)4 * SM_RB * 8, ch->st));
  SMG_CUDA(dev_malloc(&W->terms, 24 * 8, ch->st));
  SMG_CUDA(cudaMemsetAsync(W->terms, 0, 24 * 8, ch->st));
  // the two side histograms are privatised in shared memory when they fit
  W->hist_smem = (size_t)2 * ch->pp * ch->mmax * sizeof(int);
  if (W->hist_smem <= 160 * 1024) {
    SMG_CUDA(cudaFuncSetAttribute(subset_histogram_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)W->hist_smem));
  } else {
    W->hist_smem = 0;
  }
  SMG_CUDA(cudaFuncSetAttribute(sm_rdecide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)sizeof(RdecideSmem<SM_DECIDE_WIDE_CHUNK>)));
  SMG_CUDA(dev_malloc(&W->chain_bar, 8 * sizeof(unsigned), ch->st));  // [0..1] barrier counters, [2] error flag, [4..5] settled-scan counters
  SMG_CUDA(cudaMemsetAsync(W->chain_bar, 0, 8 * sizeof(unsigned), ch->st));
  SMG_CUDA(dev_malloc(&W->selcnt, 256 * sizeof(int), ch->st));
  {
    // the persistent chain kernel needs co-resident CTAs (cooperative launch) and the side histograms in shared memory
    int coop = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ch->device);
    const char* env = getenv("SMG_SM_PERSISTENT");
    const char* mode = getenv("SMG_SM_MODE");  // cluster | coop | multi (default: by measurement, see sm_step)
    W->forced_mode = !mode ? 0 : (strcmp(mode, "cluster") == 0 ? 1 : (strcmp(mode, "coop") == 0 ? 2 : 3));
    const bool multi_only = (env && env[0] == '0') || (mode && strcmp(mode, "multi") == 0);
    W->persistent = coop && W->hist_smem > 0 && W->hist_smem <= 64 * 1024 && !multi_only;
    if (W->persistent)
    {
      W->chain_smem = std::max(W->hist_smem, sizeof(RdecideSmem<SM_DECIDE_WIDE_CHUNK>));
      SMG_CUDA(cudaFuncSetAttribute(sm_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)W->chain_smem));
    }
    if (multi_only) return 0;  // the sequence of small launches only
  }
  W->smc = new SmcHost();
  return smc_setup(ch, W->smc);
}

static void sm_free(smg_chain* ch) {
  SmWork* W = ch->sm;
  if (!W) return;
  void* ptrs[] = {W->S,       W->zL,      W->zStar, W->zState, W->info,    W->plan,    W->H,      W->cnt,
                  W->rg_dl,   W->rg_lgt,  W->rg_lgt2, W->rowvals, W->partial, W->terms, W->chain_bar, W->u_pair,  W->u_prior_c, W->u_prior_s,
                  W->u_launch, W->u_rg,   W->u_rg_c, W->u_rg_s, W->u_mg_c,  W->u_mg_s,  W->u_accept, W->selcnt};
  for (void* q : ptrs)
    if (q) cudaFreeAsync(q, ch->st);
  for (cudaEvent_t e : W->tr)
    if (e) cudaEventDestroy(e);
  if (W->smc) {
    smc_free(ch, W->smc);
    delete W->smc;
  }
  delete W;
  ch->sm = nullptr;
}

static int sm_inject(smg_chain* ch, const smg_sm_tape* t) {
  SmWork* W = ch->sm;
  const int n = ch->n, p = ch->p, T1 = ch->t + 1, R1 = ch->r + 1;
  if (!W->inj_alloc) {
    SMG_CUDA(dev_malloc(&W->u_pair, 2 * 8, ch->st));
    SMG_CUDA(dev_malloc(&W->u_prior_c, (size_t)3 * p * 8, ch->st));
    SMG_CUDA(dev_malloc(&W->u_prior_s, (size_t)3 * p * 8, ch->st));
    SMG_CUDA(dev_malloc(&W->u_launch, (size_t)n * 8, ch->st));
    SMG_CUDA(dev_malloc(&W->u_rg, (size_t)T1 * n * 8, ch->st));
    SMG_CUDA(dev_malloc(&W->u_rg_c, (size_t)T1 * 2 * p * 8, ch->st));
    SMG_CUDA(dev_malloc(&W->u_rg_s, (size_t)T1 * 2 * p * 8, ch->st));
    SMG_CUDA(dev_malloc(&W->u_mg_c, (size_t)R1 * p * 8, ch->st));
    SMG_CUDA(dev_malloc(&W->u_mg_s, (size_t)R1 * p * 8, ch->st));
    SMG_CUDA(dev_malloc(&W->u_accept, 8, ch->st));
    SMG_CUDA(cudaStreamSynchronize(ch->st));
    W->inj_alloc = true;
  }
  auto up = [&](double* d, const double* h, size_t cnt) -> int {
    if (h) SMG_CUDA(h2d_sync(d, h, cnt * 8, ch->st));
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

static PhiJob sm_job(smg_chain* ch, int which, uint32_t sub, const double* uc, const double* us, int enable_mode) {
  return sm_job_at(ch->NS, which, sub, uc, us, enable_mode);
}

// one phi_update launch on up to PHI_MAX_INLINE_JOBS split-merge jobs (one CTA each)
static int sm_phi(smg_chain* ch, const PhiJob* jobs, int nj) {
  SmWork* W = ch->sm;
  PhiArgs A = phi_args_base(ch, 0);
  A.H = W->H;
  A.counts = W->cnt;
  A.njobs = nj;
  A.njobs_ptr = nullptr;
  for (int q = 0; q < nj; q++) A.jobs[q] = jobs[q];
  A.enable = &W->info->same;
  A.nparts = ch->phi_parts;
  phi_update_kernel<<<nj * A.nparts, 256, 0, ch->st>>>(A);
  ch->h_launches += 1;
  SMG_CUDA(cudaGetLastError());
  return 0;
}

// histogram of S u {i1,i2} split by z into H[h0], H[h0+1] (z == nullptr: everything into H[h0]).
// zeroed == true: the buffers were already cleared and the counts are known (restricted scans).
static int sm_hist(smg_chain* ch, const int* z, int h0, bool zeroed, const int* enable) {
  SmWork* W = ch->sm;
  const size_t len = (size_t)ch->pp * ch->mmax;
  const int nh = z ? 2 : 1;
  if (!zeroed) {
    SMG_CUDA(cudaMemsetAsync(W->H + (size_t)h0 * len, 0, nh * len * 4, ch->st));
    SMG_CUDA(cudaMemsetAsync(W->cnt + h0, 0, nh * 4, ch->st));
  }
  if (W->hist_smem) {
    subset_histogram_smem_kernel<<<32, 256, W->hist_smem, ch->st>>>(ch->X, ch->pp, W->S, &W->info->nS, z, &W->info->i1,
                                                                  ch->mmax, W->H + (size_t)h0 * len,
                                                                  zeroed ? nullptr : W->cnt + h0, enable, 1);
  } else {
    if (zeroed)  // the global-atomic kernel recounts: clear what the decision kernel published
      SMG_CUDA(cudaMemsetAsync(W->cnt + h0, 0, nh * 4, ch->st));
    long long threads = (long long)(ch->n) * (ch->pp / 16);  // upper bound; the kernel trims to |S|+2
    subset_histogram_kernel<<<sm_cdiv(threads, 256), 256, 0, ch->st>>>(ch->X, ch->pp, W->S, &W->info->nS, z,
                                                                      &W->info->i1, ch->mmax, W->H + (size_t)h0 * len,
                                                                      W->cnt + h0, enable, 1);
  }
  ch->h_launches++;
  SMG_CUDA(cudaGetLastError());
  return 0;
}

// allocation part of one restricted scan (split_merge.cpp:186-216) on sides z, parameter slots A/B;
// leaves the side histograms in H[h0], H[h0+1] and the side counts in cnt[h0], cnt[h0+1]
static int sm_restricted_alloc(smg_chain* ch, int* z, int slotA, int slotB, int h0, int q, const double* u_rg,
                               const int* enable) {
  SmWork* W = ch->sm;
  const int cur = ch->cur;
  const size_t len = (size_t)ch->pp * ch->mmax;
  sm_ll2prep_kernel<<<296, 256, 0, ch->st>>>(ch->X, ch->pp, W->S, W->info, ch->cen[cur], ch->isg[cur], ch->sden[cur], slotA,
                                            slotB, u_rg, mk_key(ch, SUB_SM_RG + q), W->rg_dl, W->rg_lgt, enable, 1);
  sm_rdecide_kernel<<<1, SM_DECIDE_T, sizeof(RdecideSmem<SM_DECIDE_WIDE_CHUNK>), ch->st>>>(
      W->info, W->rg_dl, W->rg_lgt, z, W->H + (size_t)h0 * len, (int)(2 * len), W->cnt + h0, enable, 1,
      q >= sm_wide_from() ? 1 : 0);
  ch->h_launches += 2;
  SMG_CUDA(cudaGetLastError());
  return sm_hist(ch, z, h0, true, enable);
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
  auto off = [](const double* base, size_t o) -> const double* { return base ? base + o : nullptr; };
  const size_t len = (size_t)pp * ch->mmax;
  const int* same = &W->info->same;
  static const int many_max_n = [] {
    const char* e = getenv("SMG_SM_MANY_MAXN");
    return e ? atoi(e) : 30000;
  }();
  // Default path by measurement (profiles/r02_summary.md): a single chain is fastest on the 120-CTA cooperative kernel
  // (0.43 ms per proposal at the metric shape against 0.56 ms on the 16-CTA cluster, whose phases run on 8 SMs each and
  // are bound by double-precision dependency latency); chains stepped together use the cluster kernel, which occupies
  // 16 SMs per chain instead of gang-scheduling 120.  SMG_SM_MODE=cluster|coop|multi forces one.
  const int forced = W->forced_mode;  // read when the chain was created
  const bool cluster_ok = W->smc && W->smc->ok && ch->t > 0;
  const bool use_cluster = cluster_ok && (forced == 1 || (forced == 0 && ch->many));
  const bool use_coop = !use_cluster && W->persistent && ch->t > 0 && !(ch->many && n > many_max_n);
  if (use_cluster || use_coop) {
    // ---- the whole proposal as one cooperative kernel: selection, launch states, t restricted scans, r merge-launch
    //      updates, proposal, MH ratio, acceptance
    SmChainArgs CA;
    memset(&CA, 0, sizeof(CA));
    CA.n = n;
    CA.p = p;
    CA.pp = pp;
    CA.mmax = ch->mmax;
    CA.t = ch->t;
    CA.r = ch->r;
    CA.NS = B;
    CA.Kcap = ch->Kcap;
    CA.gamma = ch->gamma;
    CA.c = ch->c;
    CA.counts = ch->counts;
    CA.Kptr = ch->K;
    CA.plan = W->plan;
    CA.zState = W->zState;
    CA.selcnt = W->selcnt;
    CA.rowvals = W->rowvals;
    CA.partial = W->partial;
    CA.terms = W->terms;
    CA.accepted = ch->accepted_d;
    CA.stats = ch->stats_d;
    CA.u_pair = T.u_pair;
    CA.u_prior_c = T.u_prior_c;
    CA.u_prior_s = T.u_prior_s;
    CA.u_launch = T.u_launch;
    CA.u_accept = T.u_accept;
    CA.pair_det = ch->pair_det;
    CA.wide_from = sm_wide_from();
    CA.X = ch->X;
    CA.S = W->S;
    CA.info = W->info;
    CA.cen = ch->cen[cur];
    CA.sig = ch->sig[cur];
    CA.isg = ch->isg[cur];
    CA.sden = ch->sden[cur];
    CA.H = W->H;
    CA.cnt = W->cnt;
    CA.zL = W->zL;
    CA.zStar = W->zStar;
    CA.dl = W->rg_dl;
    CA.lgt = W->rg_lgt;
    CA.lgt2 = W->rg_lgt2;
    CA.phi = phi_args_base(ch, 0);
    CA.phi.H = W->H;
    CA.phi.counts = W->cnt;
    CA.phi.njobs = 1;
    CA.phi.njobs_ptr = nullptr;
    CA.phi.enable = same;
    CA.u_rg = T.u_rg;
    CA.u_rg_c = T.u_rg_c;
    CA.u_rg_s = T.u_rg_s;
    CA.u_mg_c = T.u_mg_c;
    CA.u_mg_s = T.u_mg_s;
    CA.key = mk_key(ch, 0);
    if (use_cluster) {
      // ---- ONE thread-block cluster (smg_smc.cuh): members resident in shared memory, hardware cluster barriers
      static const bool trace_c = getenv("SMG_SM_TRACE") != nullptr;
      cudaEvent_t &tc0 = W->tr[0], &tc1 = W->tr[1];
      if (trace_c) {
        if (!tc0) {
          cudaEventCreate(&tc0);
          cudaEventCreate(&tc1);
        }
        cudaEventRecord(tc0, ch->st);
      }
      const cudaError_t ce = smc_launch(ch, W->smc, CA);
      if (trace_c && ce == cudaSuccess) {
        cudaEventRecord(tc1, ch->st);
        cudaEventSynchronize(tc1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, tc0, tc1);
        fprintf(stderr, "[smgibbs] sm_cluster_kernel %.1f us (%d CTAs, %d rows cached per CTA)\n", 1000.0 * ms, W->smc->CS,
                W->smc->rcap);
      }
      if (ce != cudaSuccess) {  // the cluster does not fit next to whatever else runs on this device: other paths from now on
        (void)cudaGetLastError();
        W->smc->ok = false;
        return sm_step(ch, tape);
      }
      ch->h_launches++;
      return 0;
    }
    // two barrier counters used in turn: each launch zeroes the one of the next (no memset between launches)
    CA.bar = W->chain_bar + W->bar_flip;
    CA.bar_next = W->chain_bar + (W->bar_flip ^ 1);
    W->bar_flip ^= 1;
    CA.err = reinterpret_cast<int*>(W->chain_bar + 2);
    {
      static const bool no_skip = [] { const char* e = getenv("SMG_SM_NO_SETTLED_SKIP"); return e && e[0] == '1'; }();
      CA.exc = no_skip ? nullptr : reinterpret_cast<int*>(W->chain_bar + 4);
    }
    void* kargs[] = {&CA};
    // grid: enough CTAs for the member-likelihood phase of a single chain; a small gang when several chains share the
    // GPU (smg_step_many), where the number of launches per sweep matters more than the latency of one proposal
    static const int many_ctas = [] {
      const char* e = getenv("SMG_SM_MANY_CTAS");
      return e ? atoi(e) : 8;
    }();
    static const int one_ctas = [] {
      // Gang size for a single chain.  Measured at the metric shape (gpurun_out/sched1.log, sched2.log;
      // profiles/r02_summary.md): the proposal is latency-bound and barely slows down with fewer CTAs (0.390 ms on 88,
      // 0.399 on 72, 0.424 on 56), but the likelihood block of the next pass, which runs beside it on the SMs the gang
      // leaves free, goes from 0.52 ms (120 CTAs: 28 SMs left) to 0.31 ms (72: 76 SMs) and ends before the proposal
      // does; 72 minimises the sweep (0.753 ms with 120, 0.629 with 72).
      const char* e = getenv("SMG_SM_CTAS");
      return e ? atoi(e) : 72;
    }();
    const int ctas = ch->many ? many_ctas : std::max(8, std::min(std::min(SM_CHAIN_CTAS, one_ctas), n / 800));
    {
      static const int want_hist = [] {
        const char* e = getenv("SMG_SM_HIST_CTAS");
        return e ? atoi(e) : 1 << 20;
      }();
      CA.hist_ctas = std::max(1, std::min(ctas, want_hist));
    }
    {  // up to 3 parameter-update jobs run side by side, each split over `nparts` CTAs
      static const int want = [] {
        const char* e = getenv("SMG_SM_PHI_PARTS");
        return e ? atoi(e) : 0;
      }();
      // default: 32 attributes (x 8 lanes) per CTA, i.e. one pass of half a CTA
      const int w = want > 0 ? want : (ch->pp + 31) / 32;
      CA.phi.nparts = phi_parts_for(ch->pp, std::max(1, std::min(w, ctas / 3)));
    }
    static const bool trace = getenv("SMG_SM_TRACE") != nullptr;  // diagnostic: device time of the persistent kernel
    cudaEvent_t &tr0 = W->tr[0], &tr1 = W->tr[1];  // per chain: chains are stepped from several host threads
    if (trace) {
      if (!tr0) {
        cudaEventCreate(&tr0);
        cudaEventCreate(&tr1);
      }
      cudaEventRecord(tr0, ch->st);
    }
    const cudaError_t ce =
        cudaLaunchCooperativeKernel((const void*)sm_chain_kernel, dim3(ctas), dim3(SM_CHAIN_T), kargs, W->chain_smem, ch->st);
    if (trace && ce == cudaSuccess) {
      cudaEventRecord(tr1, ch->st);
      cudaEventSynchronize(tr1);
      float ms = 0.f;
      cudaEventElapsedTime(&ms, tr0, tr1);
      fprintf(stderr, "[smgibbs] sm_chain_kernel %.1f us (%d CTAs)\n", 1000.0 * ms, ctas);
    }
    if (ce == cudaErrorCooperativeLaunchTooLarge || ce == cudaErrorLaunchOutOfResources) {
      // the gang does not fit next to whatever else runs on this device: use the sequence of launches from now on
      (void)cudaGetLastError();
      W->persistent = false;
      return sm_step(ch, tape);
    }
    SMG_CUDA(ce);
    ch->h_launches++;
    return 0;
  }
  // ---- pair, S, plan
  sm_select_kernel<<<1, 1024, 0, ch->st>>>(n, ch->c, ch->K, T.u_pair, mk_key(ch, SUB_SM_SELECT), B, W->S, W->zState,
                                           W->info, W->plan, W->cnt, W->terms, ch->pair_det);
  ch->h_launches++;
  SMG_CUDA(cudaGetLastError());
  // ---- histograms of the current-state sides and of the merged cluster (fixed for the whole proposal)
  if (sm_hist(ch, W->zState, SH_S0, false, nullptr)) return SMG_ERR_CUDA;
  sm_hist_add_kernel<<<sm_cdiv(len, 256), 256, 0, ch->st>>>((int)len, W->H, W->cnt, SH_S0, SH_S1, SH_M);
  ch->h_launches++;
  // ---- prior parameters of the three launch clusters (split_merge.cpp:331-343, :379-380)
  {
    PhiJob j[3];
    for (int k = 0; k < 3; k++)
      j[k] = sm_job(ch, J_PRI_A + k, SUB_SM_PRIOR, off(T.u_prior_c, (size_t)k * p), off(T.u_prior_s, (size_t)k * p), 0);
    if (sm_phi(ch, j, 3)) return SMG_ERR_CUDA;
  }
  // ---- split launch: random sides then t restricted scans (split_merge.cpp:346-349); the r parameter
  //      updates of the merge launch (split_merge.cpp:386-387) are an independent chain on the fixed merged
  //      histogram, so update q of it rides in the same launch as the update of scan q
  sm_launch_alloc_kernel<<<sm_cdiv(n, 256), 256, 0, ch->st>>>(W->info, T.u_launch, mk_key(ch, SUB_SM_LAUNCH), W->zL);
  ch->h_launches++;
  const int nsteps = ch->t > ch->r ? ch->t : ch->r;
  for (int q = 0; q < nsteps; q++) {
    PhiJob j[3];
    int nj = 0;
    if (q < ch->t) {
      int rc = sm_restricted_alloc(ch, W->zL, B + SM_SL_A, B + SM_SL_B, SH_L0, q, off(T.u_rg, (size_t)q * n), nullptr);
      if (rc) return rc;
      for (int side = 0; side < 2; side++)
        j[nj++] = sm_job(ch, J_L0 + side, SUB_SM_RG + q, off(T.u_rg_c, ((size_t)q * 2 + side) * p),
                         off(T.u_rg_s, ((size_t)q * 2 + side) * p), 0);
    }
    if (q < ch->r)
      j[nj++] = sm_job(ch, J_MG, SUB_SM_MERGE + q, off(T.u_mg_c, (size_t)q * p), off(T.u_mg_s, (size_t)q * p), 0);
    if (sm_phi(ch, j, nj)) return SMG_ERR_CUDA;
  }
  if (ch->t == 0 && sm_hist(ch, W->zL, SH_L0, false, nullptr)) return SMG_ERR_CUDA;  // launch counts are still needed
  // ---- proposal
  //   split (same == 1): star = split launch + one more restricted scan (split_merge.cpp:575-580)
  //   merge (same == 0): star = merge launch + one more update_phi (split_merge.cpp:582-586); it is drawn in
  //   both cases (it only fills the M* slot, the MH kernel ignores it for a split)
  sm_begin_proposal_kernel<<<64, 256, 0, ch->st>>>(W->info, W->zL, W->zStar, pp, ch->cen[cur], ch->sig[cur], ch->isg[cur],
                                                  ch->sden[cur], B + SM_SL_A, B + SM_ST_A, B + SM_SL_B, B + SM_ST_B);
  ch->h_launches++;
  {
    const int q = ch->t;
    int rc = sm_restricted_alloc(ch, W->zStar, B + SM_ST_A, B + SM_ST_B, SH_P0, q, off(T.u_rg, (size_t)q * n), same);
    if (rc) return rc;
    PhiJob j[3];
    for (int side = 0; side < 2; side++)
      j[side] = sm_job(ch, J_P0 + side, SUB_SM_RG + q, off(T.u_rg_c, ((size_t)q * 2 + side) * p),
                       off(T.u_rg_s, ((size_t)q * 2 + side) * p), 1);
    j[2] = sm_job(ch, J_MSTAR, SUB_SM_MERGE + ch->r, off(T.u_mg_c, (size_t)ch->r * p), off(T.u_mg_s, (size_t)ch->r * p), 0);
    if (sm_phi(ch, j, 3)) return SMG_ERR_CUDA;
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
  SMG_CUDA(d2h_sync(&I, W->info, sizeof(I), ch->st));
  int cnt[SH_N];
  SMG_CUDA(d2h_sync(cnt, W->cnt, sizeof(cnt), ch->st));
  int acc = 0, K = 0;
  SMG_CUDA(d2h_sync(&acc, ch->accepted_d, 4, ch->st));
  SMG_CUDA(d2h_sync(&K, ch->K, 4, ch->st));
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
  if (S) SMG_CUDA(d2h_sync(S, W->S, (size_t)I.nS * 4, ch->st));
  if (z_launch) SMG_CUDA(d2h_sync(z_launch, W->zL, (size_t)I.nS * 4, ch->st));
  if (z_star) SMG_CUDA(d2h_sync(z_star, W->zStar, (size_t)I.nS * 4, ch->st));
  if (phi_out) {
    const int slots[6] = {SM_SL_A, SM_SL_B, SM_ML_M, SM_ST_A, SM_ST_B, SM_ST_M};
    std::vector<uint8_t> hc(ch->pp);
    std::vector<double> hs(ch->pp);
    for (int q = 0; q < 6; q++) {
      const size_t off = (size_t)(ch->NS + slots[q]) * ch->pp;
      SMG_CUDA(d2h_sync(hc.data(), ch->cen[ch->cur] + off, ch->pp, ch->st));
      SMG_CUDA(d2h_sync(hs.data(), ch->sig[ch->cur] + off, (size_t)ch->pp * 8, ch->st));
      for (int j = 0; j < ch->p; j++) {
        phi_out[((size_t)q * 2 + 0) * ch->p + j] = hc[j];
        phi_out[((size_t)q * 2 + 1) * ch->p + j] = hs[j];
      }
    }
  }
  if (terms) SMG_CUDA(d2h_sync(terms, W->terms, 24 * 8, ch->st));
  return 0;
}

}  // namespace smg
