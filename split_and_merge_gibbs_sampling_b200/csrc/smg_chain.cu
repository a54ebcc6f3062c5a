// smg_chain.cu -- host driver of one chain + the extern "C" layer declared in include/smgibbs.h.
// Mirrors code/launcher.cpp:7-174 (state initialisation, aux pool, iteration loop, snapshots);
// all arithmetic happens in the kernels of smg_kernels.cuh / smg_sm.cuh.  There is no CPU path.
#include <atomic>
#include <thread>

#include "smg_chain.cuh"

#include <sched.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <mutex>
#include <thread>

#include "smg_psm.cuh"
#include "smg_lltc.cuh"
#include "smg_comm.cuh"
#include "smg_post.cuh"
#include "smg_sm_host.cuh"

namespace smg {
thread_local std::string g_last_error;

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

static thread_local cudaStream_t g_alloc_stream = nullptr;  // stream of the chain being allocated
template <typename T>
static int dalloc(T** ptr, size_t count) {
  SMG_CUDA(dev_malloc(ptr, std::max<size_t>(count, 1) * sizeof(T), g_alloc_stream));
  return 0;
}

// ziggurat tables of the normal generator (smg_device.cuh), once per device
static int zig_init(int device) {
  static std::mutex mu;
  static bool done[64] = {};
  std::lock_guard<std::mutex> lk(mu);
  if (device < 0 || device >= 64) return fail(SMG_ERR_ARG, "bad device ordinal");
  if (done[device]) return 0;
  double x[SMG_ZIG_C + 1], r[SMG_ZIG_C];
  double f = std::exp(-0.5 * SMG_ZIG_R * SMG_ZIG_R);
  x[0] = SMG_ZIG_V / f;
  x[1] = SMG_ZIG_R;
  x[SMG_ZIG_C] = 0.0;
  for (int i = 2; i < SMG_ZIG_C; i++) {
    x[i] = std::sqrt(-2.0 * std::log(SMG_ZIG_V / x[i - 1] + f));
    f = std::exp(-0.5 * x[i] * x[i]);
  }
  for (int i = 0; i < SMG_ZIG_C; i++) r[i] = x[i + 1] / x[i];
  SMG_CUDA(cudaSetDevice(device));
  SMG_CUDA(cudaMemcpyToSymbol(g_zig_x, x, sizeof(x)));
  SMG_CUDA(cudaMemcpyToSymbol(g_zig_r, r, sizeof(r)));
  SMG_CUDA(cudaDeviceSynchronize());
  done[device] = true;
  return 0;
}

static int status_to_error(int st) {
  if (st == 0) return SMG_OK;
  if (st & ST_BAD_PROB) return fail(SMG_ERR_PROB, "Probabilities must be finite and non-negative! (allocation draw)");
  if (st & ST_VALIDATE) return fail(SMG_ERR_STATE, "State validation failed: inconsistent cluster count");
  if (st & ST_WALKER) return fail(SMG_ERR_CAPACITY, "allocation draw hit Rcpp's Walker-alias regime (nc > 200), not supported");
  if (st & ST_TOO_MANY_ENTRIES) return fail(SMG_ERR_CAPACITY, "K + m exceeds 256 entries per allocation draw");
  if (st & ST_LL_COLS) return fail(SMG_ERR_CAPACITY, "number of clusters exceeds max_clusters");
  if (st & ST_SLOTS_EXHAUSTED) return fail(SMG_ERR_CAPACITY, "more than 1024 cluster slots used in one pass");
  if (st & ST_GRID_TIMEOUT) return fail(SMG_ERR_CUDA, "grid barrier of the split-merge chain kernel timed out");
  return fail(SMG_ERR_STATE, "device status " + std::to_string(st));
}

// ------------------------------------------------------------------------------------------
static int chain_alloc(smg_chain* ch) {
  const int n = ch->n, pp = ch->pp, NST = ch->NST;
  SMG_CUDA(cudaSetDevice(ch->device));
  SMG_CUDA(dev_pool_init(ch->device));
  if (zig_init(ch->device)) return SMG_ERR_CUDA;
  {
    // the chain's (latency-bound) stream outranks the side stream: the side stream's bandwidth-bound gather fills
    // every SM and would otherwise hold back the small kernels it is meant to overlap
    int prio_lo = 0, prio_hi = 0;
    SMG_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    SMG_CUDA(cudaStreamCreateWithPriority(&ch->st, cudaStreamNonBlocking, prio_hi));
    SMG_CUDA(cudaStreamCreateWithPriority(&ch->st_aux, cudaStreamNonBlocking, prio_lo));
    SMG_CUDA(cudaStreamCreateWithPriority(&ch->st_k1, cudaStreamNonBlocking, prio_lo));
  }
  SMG_CUDA(cudaEventCreateWithFlags(&ch->ev_phi_done, cudaEventDisableTiming));
  SMG_CUDA(cudaEventCreateWithFlags(&ch->ev_aux_go, cudaEventDisableTiming));
  SMG_CUDA(cudaEventCreateWithFlags(&ch->ev_k1_done, cudaEventDisableTiming));
  SMG_CUDA(cudaEventCreateWithFlags(&ch->ev_scan_done, cudaEventDisableTiming));
  SMG_CUDA(cudaEventCreateWithFlags(&ch->ev_aux_done, cudaEventDisableTiming));
  SMG_CUDA(cudaEventCreate(&ch->ev_k1[0]));
  SMG_CUDA(cudaEventCreate(&ch->ev_k1[1]));
  SMG_CUDA(cudaEventCreate(&ch->ev_aux_t0));
  SMG_CUDA(cudaEventCreate(&ch->ev_aux_t1));
  g_alloc_stream = ch->st;
  for (int q = 0; q < 8; q++) SMG_CUDA(cudaEventCreate(&ch->ev[q]));
  for (int q = 0; q < 2; q++) SMG_CUDA(cudaEventCreate(&ch->ev_call[q]));
  SMG_CUDA(cudaFuncSetAttribute(neal8_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SCAN_PF_BYTES));
  SMG_CUDA(cudaFuncSetAttribute(hamming_ll_block_t16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LLT_SMEM_BYTES));
  SMG_CUDA(cudaFuncSetAttribute(cluster_histogram_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
  if (dalloc(&ch->X, (size_t)n * pp)) return SMG_ERR_CUDA;
  if (dalloc(&ch->attr, pp) || dalloc(&ch->v, pp) || dalloc(&ch->w, pp)) return SMG_ERR_CUDA;
  for (int b = 0; b < 2; b++) {
    if (dalloc(&ch->cen[b], (size_t)NST * pp) || dalloc(&ch->sig[b], (size_t)NST * pp) ||
        dalloc(&ch->isg[b], (size_t)NST * pp) || dalloc(&ch->sden[b], NST))
      return SMG_ERR_CUDA;
    SMG_CUDA(cudaMemsetAsync(ch->cen[b], 0, (size_t)NST * pp, ch->st));
    SMG_CUDA(cudaMemsetAsync(ch->sig[b], 0, (size_t)NST * pp * 8, ch->st));
    SMG_CUDA(cudaMemsetAsync(ch->isg[b], 0, (size_t)NST * pp * 8, ch->st));
    SMG_CUDA(cudaMemsetAsync(ch->sden[b], 0, (size_t)NST * 8, ch->st));
  }
  if (dalloc(&ch->den, (size_t)NST * pp) || dalloc(&ch->phi_cnt, NST)) return SMG_ERR_CUDA;
  SMG_CUDA(cudaMemsetAsync(ch->phi_cnt, 0, (size_t)NST * 4, ch->st));
  {
    const char* e = getenv("SMG_PHI_PARTS");
    ch->phi_parts = phi_parts_for(pp, e ? atoi(e) : (pp + 31) / 32);  // 32 attributes x PHI_G lanes = one pass of a 256-thread CTA
  }
  if (dalloc(&ch->c, n + 4) || dalloc(&ch->c_hist, n) || dalloc(&ch->K, 1) || dalloc(&ch->counts, NST) || dalloc(&ch->counts_slot, NST) ||
      dalloc(&ch->slot2label, NST))
    return SMG_ERR_CUDA;
  if (dalloc(&ch->LL, (size_t)n * ch->ldl) || dalloc(&ch->LLaux[0], (size_t)n * ch->m_aux) || dalloc(&ch->LLaux[1], (size_t)n * ch->m_aux) || dalloc(&ch->mrg, n + 2) || dalloc(&ch->und0, ((size_t)n + 8191) / 4096 * 4096) || dalloc(&ch->und_blk, (size_t)n / 4096 + 4) ||
      dalloc(&ch->aux_e[0], (size_t)n * ch->m_aux) || dalloc(&ch->aux_e[1], (size_t)n * ch->m_aux))
    return SMG_ERR_CUDA;
  if (ch->aux_mode == 1 &&
      (dalloc(&ch->aux_sd[0], (size_t)n * ch->m_aux) || dalloc(&ch->aux_sd[1], (size_t)n * ch->m_aux)))
    return SMG_ERR_CUDA;
  const size_t P = (size_t)ch->pool_size;
  if (dalloc(&ch->pcen, P * pp) || dalloc(&ch->psig, P * pp) || dalloc(&ch->pisg, P * pp) || dalloc(&ch->pden, P * pp) ||
      dalloc(&ch->psden, P))
    return SMG_ERR_CUDA;
  if (dalloc(&ch->H, (size_t)ch->Kcap * pp * ch->mmax)) return SMG_ERR_CUDA;
  ch->loglik_blocks = std::min(1184, std::max(1, cdiv(n, 8)));
  if (dalloc(&ch->partial, ch->loglik_blocks) || dalloc(&ch->loglik_d, 1)) return SMG_ERR_CUDA;
  if (dalloc(&ch->status, 1) || dalloc(&ch->accepted_d, 1) || dalloc(&ch->stats_d, 8) || dalloc(&ch->scan_job, 4) || dalloc(&ch->scan_prof, 16))
    return SMG_ERR_CUDA;
  SMG_CUDA(cudaMemsetAsync(ch->scan_prof, 0, 128, ch->st));
  SMG_CUDA(cudaMemsetAsync(ch->und0, 0, ((size_t)n + 8191) / 4096 * 4096, ch->st));
  SMG_CUDA(cudaMemsetAsync(ch->und_blk, 0, ((size_t)n / 4096 + 4) * sizeof(int), ch->st));
  SMG_CUDA(cudaMemsetAsync(ch->status, 0, 4, ch->st));
  SMG_CUDA(cudaMemsetAsync(ch->accepted_d, 0, 4, ch->st));
  SMG_CUDA(cudaMemsetAsync(ch->stats_d, 0, 64, ch->st));
  {
    // K1 on the tensor cores when its resident operand fits the shared memory of one SM
    const char* e = getenv("SMG_K1");  // tc (default when feasible) | t16 | c
    int smem_optin = 0, sms = 0;
    cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ch->device);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ch->device);
    ch->ltc_kdp = ltc_kdp(pp, ch->mmax);
    ch->ltc_sms = sms;
    cudaFuncAttributes fa;
    bool fits = cudaFuncGetAttributes(&fa, hamming_ll_tc_kernel) == cudaSuccess;
    (void)cudaGetLastError();
    if (fits) {
      // operand ring: as many k-steps per stage as fit beside the resident B operand (8, 4, 2 or 1), >= 3 stages
      const size_t room = (size_t)smem_optin - fa.sharedSizeBytes - 1024;
      const size_t bb = ltc_b_bytes(ch->ltc_kdp) + (size_t)LTC_M * (pp + 16);  // resident B operand + the staged codes of a row tile
      const char* ekg = getenv("SMG_LTC_KG");
      fits = false;
      for (int kg = ekg ? atoi(ekg) : 8; kg >= 1; kg >>= 1) {
        const int nst = (int)std::min<size_t>(LTC_STAGES, bb < room ? (room - bb) / ((size_t)kg * LTC_A_BYTES) : 0);
        if (nst >= 3 || (kg == 1 && nst >= 2)) {
          ch->ltc_kg = kg;
          ch->ltc_nstg = nst;
          ch->ltc_smem = bb + (size_t)nst * kg * LTC_A_BYTES;
          fits = true;
          break;
        }
      }
    }
    // measured at the metric shape: 247 us against 175 us of the shared-memory look-up kernel (the one-hot operand is
    // written with generic-proxy stores and every stage pays a proxy fence; tensor pipe 10% active) -> opt-in
    ch->ltc_on = fits && e && strcmp(e, "tc") == 0;
    if (ch->ltc_on) {
      SMG_CUDA(cudaFuncSetAttribute(hamming_ll_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)((size_t)smem_optin - fa.sharedSizeBytes)));
      const int ntmax = cdiv(ch->Kcap, LTC_NC);
      if (dalloc(&ch->ltc_B, (size_t)ntmax * ltc_b_bytes(ch->ltc_kdp)) || dalloc(&ch->ltc_Q, (size_t)ntmax * LTC_NC) ||
          dalloc(&ch->ltc_scale, (size_t)ntmax * LTC_NC) || dalloc(&ch->ltc_cst, (size_t)ntmax * LTC_NC * 4) || dalloc(&ch->ltc_K, 1) ||
          dalloc(&ch->ltc_ctr, 64))
        return SMG_ERR_CUDA;
    }
  }
  if (dalloc(&ch->tape_d, (size_t)n * (ch->m_aux + 1))) return SMG_ERR_CUDA;
  if (dalloc(&ch->uc_d, (size_t)NST * ch->p) || dalloc(&ch->us_d, (size_t)NST * ch->p)) return SMG_ERR_CUDA;
  int rc = sm_alloc(ch);
  if (rc) return rc;
  SMG_CUDA(cudaStreamSynchronize(ch->st));  // every buffer is now usable from any stream
  return 0;
}

static void chain_free(smg_chain* ch) {
  if (!ch) return;
  cudaSetDevice(ch->device);
  if (ch->st_aux) cudaStreamSynchronize(ch->st_aux);
  if (ch->st_k1) cudaStreamSynchronize(ch->st_k1);
  if (ch->st) cudaStreamSynchronize(ch->st);
  sm_free(ch);
  void* ptrs[] = {ch->X,      ch->attr,   ch->v,         ch->w,          ch->cen[0], ch->cen[1], ch->sig[0], ch->sig[1],
                  ch->isg[0], ch->isg[1], ch->sden[0],   ch->sden[1],    ch->den,    ch->c,      ch->K,      ch->counts,
                  ch->counts_slot, ch->slot2label, ch->LL, ch->LLaux[0], ch->LLaux[1], ch->mrg, ch->aux_e[0], ch->aux_e[1], ch->aux_sd[0], ch->aux_sd[1], ch->pcen,   ch->psig,   ch->pisg,
                  ch->pden,   ch->psden,  ch->H,         ch->partial,    ch->loglik_d, ch->status, ch->accepted_d,
                  ch->stats_d, ch->scan_job, ch->scan_prof, ch->tape_d, ch->uc_d,     ch->us_d,
                  ch->c_hist, ch->phi_cnt, ch->und0, ch->und_blk, ch->ltc_B, ch->ltc_Q, ch->ltc_scale, ch->ltc_K, ch->ltc_ctr, ch->ltc_cst};
  for (void* q : ptrs)
    if (q) cudaFreeAsync(q, ch->st);
  if (ch->st) cudaStreamSynchronize(ch->st);
  for (int q = 0; q < 8; q++)
    if (ch->ev[q]) cudaEventDestroy(ch->ev[q]);
  for (int q = 0; q < 2; q++)
    if (ch->ev_call[q]) cudaEventDestroy(ch->ev_call[q]);
  if (ch->ev_scan_done) cudaEventDestroy(ch->ev_scan_done);
  if (ch->ev_aux_done) cudaEventDestroy(ch->ev_aux_done);
  for (int q = 0; q < 2; q++)
    if (ch->ev_k1[q]) cudaEventDestroy(ch->ev_k1[q]);
  if (ch->ev_aux_t0) cudaEventDestroy(ch->ev_aux_t0);
  if (ch->ev_aux_t1) cudaEventDestroy(ch->ev_aux_t1);
  if (ch->ev_phi_done) cudaEventDestroy(ch->ev_phi_done);
  if (ch->ev_aux_go) cudaEventDestroy(ch->ev_aux_go);
  if (ch->ev_k1_done) cudaEventDestroy(ch->ev_k1_done);
  if (ch->st_aux) cudaStreamDestroy(ch->st_aux);
  if (ch->st_k1) cudaStreamDestroy(ch->st_k1);
  if (ch->st) cudaStreamDestroy(ch->st);
  delete ch;
}

// ------------------------------------------------------------------------------------------
// phases
// ------------------------------------------------------------------------------------------
static int launch_ll_block(smg_chain* ch, cudaStream_t stream = nullptr) {
  if (!stream) stream = ch->st;
  cudaEventRecord(ch->ev_k1[0], stream);
  dim3 grid(cdiv(ch->n, LLB_ROWS), cdiv(ch->Kcap, LLB_SLOTS));
  if (ch->ltc_on) {
    // exact integer GEMM on the tensor cores: one-hot(X) x fixed-point digit planes of 1/sigma (smg_lltc.cuh)
    const int ntmax = cdiv(ch->Kcap, LTC_NC);
    SMG_CUDA(cudaMemcpyAsync(ch->ltc_K, ch->K, sizeof(int), cudaMemcpyDeviceToDevice, stream));  // one consistent K
    SMG_CUDA(cudaMemsetAsync(ch->ltc_ctr, 0, 64 * sizeof(int), stream));
    ll_tc_prep_kernel<<<ntmax * LTC_NC, 256, 0, stream>>>(ch->pp, ch->mmax, ch->ltc_kdp, ch->ltc_K, ch->cen[ch->cur],
                                                          ch->isg[ch->cur], ch->ltc_B, ch->ltc_Q, ch->ltc_scale, ch->sden[ch->cur], ch->ltc_cst);
    hamming_ll_tc_kernel<<<ch->ltc_sms, LTC_THREADS, ch->ltc_smem, stream>>>(
        ch->X, ch->n, ch->pp, ch->mmax, ch->ltc_kdp, ch->ltc_K, ch->ltc_B, ch->ltc_Q, ch->ltc_scale, ch->sden[ch->cur], ch->LL,
        ch->ldl, ch->ltc_ctr, ch->ltc_kg, ch->ltc_nstg, ch->ltc_cst, ch->scan_prof);
    ch->h_launches++;
  } else if (ch->mmax <= 7)  // every code fits 3 bits: subset-sum table form
    hamming_ll_block_t16_kernel<<<grid, 256, LLT_SMEM_BYTES, stream>>>(ch->X, ch->n, ch->pp, ch->cen[ch->cur],
                                                                      ch->isg[ch->cur], ch->sden[ch->cur], ch->K, ch->LL,
                                                                      ch->ldl);
  else
    hamming_ll_block_kernel<<<grid, 256, 0, stream>>>(ch->X, ch->n, ch->pp, ch->cen[ch->cur], ch->isg[ch->cur],
                                                      ch->sden[ch->cur], ch->K, ch->LL, ch->ldl);
  cudaEventRecord(ch->ev_k1[1], stream);
  ch->k1_timed = true;
  ch->h_launches++;
  SMG_CUDA(cudaGetLastError());
  return 0;
}

static int launch_aux_ll(smg_chain* ch, const double* tape, cudaStream_t stream, long long iter, int buf) {
  long long warps = (long long)ch->n * ch->m_aux;
  RngKey key = mk_key(ch, SUB_SCAN);
  key.sweep = (uint32_t)iter;
  if (ch->aux_mode == 1) {  // pool-free: fresh prior draws from the pass's own key, nothing stored (the tape's pool-index
                            // uniforms have no meaning here)
    key.sub = SUB_AUX_FREE;
    aux_ll_philox_kernel<<<cdiv(warps * 32, 256), 256, 0, stream>>>(ch->X, ch->n, ch->pp, ch->p, ch->m_aux, ch->attr, ch->v, ch->w,
                                                                   key, ch->sigma_exact, ch->LLaux[buf], ch->aux_e[buf],
                                                                   ch->aux_sd[buf]);
    ch->h_launches++;
    SMG_CUDA(cudaGetLastError());
    return 0;
  }
  aux_ll_kernel<<<cdiv(warps * 32, 256), 256, 0, stream>>>(ch->X, ch->n, ch->pp, ch->m_aux, ch->pcen, ch->pisg, ch->psden,
                                                          ch->pool_size, tape, ch->m_aux + 1, key, ch->LLaux[buf], ch->aux_e[buf]);
  ch->h_launches++;
  SMG_CUDA(cudaGetLastError());
  return 0;
}

// The auxiliary-component columns depend on X, the pool and the Philox key only -- not on the state -- so the ones
// of the NEXT pass are evaluated on a side stream, into the other half of the double buffer, while the (single-cluster)
// scan and update_phi of THIS pass leave the GPU almost idle: the HBM-bound gather is completely hidden.  Not done
// when the pool is re-drawn in between (launcher.cpp:123-129).
static int prefetch_next_aux(smg_chain* ch) {
  const long long next_iter = ch->iter + ch->n8_step;
  static const bool disabled = [] { const char* e = getenv("SMG_NO_AUX_PREFETCH"); return e && e[0] == '1'; }();
  if (disabled || !ch->neal8) return 0;
  // an iteration that re-draws the pool before the next pass would invalidate the columns
  for (long long it = ch->iter; it < next_iter; it++)
    if (it % 1000 == 0 && ch->aux_mode == 0) return 0;
  // everything enqueued on the chain's stream so far must be complete: the other buffer was last read by the scan of
  // the previous pass, and the pool may have been re-drawn since (launcher.cpp:123-129)
  SMG_CUDA(cudaEventRecord(ch->ev_aux_go, ch->st));
  SMG_CUDA(cudaStreamWaitEvent(ch->st_aux, ch->ev_aux_go, 0));
  cudaEventRecord(ch->ev_aux_t0, ch->st_aux);
  if (launch_aux_ll(ch, nullptr, ch->st_aux, next_iter, ch->aux_buf ^ 1)) return SMG_ERR_CUDA;
  cudaEventRecord(ch->ev_aux_t1, ch->st_aux);
  ch->aux_timed = true;
  SMG_CUDA(cudaEventRecord(ch->ev_aux_done, ch->st_aux));
  ch->aux_ready = true;
  ch->aux_iter = next_iter;
  return 0;
}

// one Neal-8 pass: launcher.cpp:95-99
static int neal8_pass(smg_chain* ch, const double* tape, bool timed, bool prefetch = false) {
  if (!ch->pool_valid) return fail(SMG_ERR_STATE, "auxiliary pool not initialised");
  if (timed) cudaEventRecord(ch->ev[0], ch->st);
  // the block of this pass was normally evaluated during the previous iteration (see sweep())
  if (tape || ch->ll_for_iter != ch->iter) {
    if (launch_ll_block(ch)) return SMG_ERR_CUDA;
  }
  ch->ll_for_iter = -1;  // the pass moves observations: the block no longer matches the state
  if (timed) cudaEventRecord(ch->ev[1], ch->st);
  if (!tape && ch->aux_ready && ch->aux_iter == ch->iter) {
    // the columns were evaluated on the side stream during the previous iteration, into the other buffer
    SMG_CUDA(cudaStreamWaitEvent(ch->st, ch->ev_aux_done, 0));
    ch->aux_buf ^= 1;
  } else {
    if (ch->aux_ready) SMG_CUDA(cudaStreamWaitEvent(ch->st, ch->ev_aux_done, 0));  // do not race a stale prefetch
    if (launch_aux_ll(ch, tape, ch->st, ch->iter, ch->aux_buf)) return SMG_ERR_CUDA;
  }
  ch->aux_ready = false;
  if (prefetch && !tape) {
    int rc = prefetch_next_aux(ch);
    if (rc) return rc;
  }
  if (timed) cudaEventRecord(ch->ev[2], ch->st);
  ScanArgs A;
  A.n = ch->n;
  A.pp = ch->pp;
  A.m_aux = ch->m_aux;
  A.ldl = ch->ldl;
  A.K0cap = ch->Kcap;
  A.X = ch->X;
  A.LL = ch->LL;
  A.LLaux = ch->LLaux[ch->aux_buf];
  A.mrg = ch->mrg;
  A.aux_e = ch->aux_e[ch->aux_buf];
  A.u_alloc = tape ? tape + ch->m_aux : nullptr;
  A.u_stride = ch->m_aux + 1;
  A.key = mk_key(ch, SUB_SCAN);
  A.c = ch->c;
  A.cen = ch->cen[ch->cur];
  A.sig = ch->sig[ch->cur];
  A.isg = ch->isg[ch->cur];
  A.sden = ch->sden[ch->cur];
  A.pool_cen = ch->pcen;
  A.pool_sig = ch->psig;
  A.pool_isg = ch->pisg;
  A.pool_sden = ch->psden;
  A.aux_free = ch->aux_mode == 1;
  A.p = ch->p;
  A.sigma_exact = ch->sigma_exact;
  A.aux_key = mk_key(ch, SUB_AUX_FREE);
  A.attr = ch->attr;
  A.hv = ch->v;
  A.hw = ch->w;
  A.aux_sd = ch->aux_sd[ch->aux_buf];
  A.Kptr = ch->K;
  A.counts = ch->counts;
  A.slot2label = ch->slot2label;
  A.NS = ch->NS;
  A.log_gamma_m = std::log(ch->gamma / ch->m_aux);
  A.status = ch->status;
  A.stats = ch->stats_d;
  A.job = ch->scan_job;
  A.und0 = ch->und0;
  A.und_blk = ch->und_blk;
  A.prof = ch->scan_prof;
  {
    // speculative evaluation of the scan (SMG_SCAN_SPEC=0 switches it off, 2 adds the self-check; the tests run all three)
    static const int spec = [] {
      const char* e = getenv("SMG_SCAN_SPEC");
      return e ? atoi(e) : 1;
    }();
    static const int rmax = [] {
      const char* e = getenv("SMG_SCAN_SPEC_RMAX");
      return e ? std::max(1, atoi(e)) : SCAN_SPEC_RMAX;
    }();
    static const double dmax = [] {
      const char* e = getenv("SMG_SCAN_SPEC_DMAX");
      return e ? std::min(0.1, atof(e)) : 0.05;
    }();
    A.spec = ch->scan_spec >= 0 ? ch->scan_spec : spec;
    A.spec_rmax = rmax;
    A.spec_dmax = dmax;
  }
  scan_margin_kernel<<<std::min(cdiv(ch->n, 8), 148 * 8), 256, 0, ch->st>>>(ch->n, ch->K, ch->ldl, ch->m_aux, ch->LL, A.LLaux,
                                                                          ch->c, ch->counts, A.log_gamma_m, A.u_alloc, A.u_stride, A.key, ch->mrg, ch->und0, ch->und_blk);
  neal8_scan_kernel<<<SCAN_CLUSTER, SMG_SCAN_WARPS * 32, SCAN_PF_BYTES, ch->st>>>(A);  // one cluster
  SMG_CUDA(cudaGetLastError());
  // back to canonical form: labels in c, parameters in label order in the other buffer
  const int nx = ch->cur ^ 1;
  SMG_CUDA(cudaMemsetAsync(ch->counts_slot, 0, (size_t)ch->NST * sizeof(int), ch->st));
  scan_finalize_labels_kernel<<<cdiv(ch->n, 256), 256, 0, ch->st>>>(ch->c, ch->n, ch->slot2label);
  scan_finalize_params_kernel<<<ch->NS, 128, 0, ch->st>>>(ch->slot2label, ch->NS, ch->pp, ch->cen[ch->cur],
                                                         ch->sig[ch->cur], ch->isg[ch->cur], ch->sden[ch->cur],
                                                         ch->counts, ch->cen[nx], ch->sig[nx], ch->isg[nx],
                                                         ch->sden[nx], ch->counts_slot);
  SMG_CUDA(cudaGetLastError());
  std::swap(ch->counts, ch->counts_slot);
  ch->cur = nx;
  ch->h_launches += 4;
  SMG_CUDA(cudaEventRecord(ch->ev_scan_done, ch->st));
  if (timed) cudaEventRecord(ch->ev[3], ch->st);
  return 0;
}

// K3: histogram + counts of the canonical state
static int launch_histogram(smg_chain* ch) {
  static const bool incremental = [] {
    const char* e = getenv("SMG_NO_INCR_HIST");
    return !(e && atoi(e) != 0);
  }();
  SMG_CUDA(cudaMemsetAsync(ch->counts, 0, (size_t)ch->NST * sizeof(int), ch->st));
  if (ch->hist_valid && incremental) {  // H matches c_hist: move only the rows whose label changed since
    const int ctas = std::max(1, std::min(148 * 4, cdiv(ch->n, 256)));
    cluster_histogram_update_kernel<<<ctas, 256, (size_t)ch->Kcap * sizeof(int), ch->st>>>(
        ch->X, ch->n, ch->pp, ch->c, ch->c_hist, ch->mmax, ch->Kcap, ch->H, ch->counts);
    ch->h_launches++;
    SMG_CUDA(cudaGetLastError());
    return 0;
  }
  SMG_CUDA(cudaMemsetAsync(ch->H, 0, (size_t)ch->Kcap * ch->pp * ch->mmax * sizeof(int), ch->st));
  SMG_CUDA(cudaMemcpyAsync(ch->c_hist, ch->c, (size_t)ch->n * sizeof(int), cudaMemcpyDeviceToDevice, ch->st));
  ch->hist_valid = true;
  const size_t smem = ((size_t)ch->Kcap * 16 * ch->mmax + ch->Kcap) * sizeof(int);
  if (smem <= 72 * 1024) {  // three CTAs per SM
    const int rg = std::max(1, std::min(32, cdiv(ch->n, 256 * 8)));
    cluster_histogram_smem_kernel<<<dim3(ch->pp / 16, rg), 256, smem, ch->st>>>(ch->X, ch->n, ch->pp, ch->c, ch->mmax, ch->K,
                                                                               ch->Kcap, ch->H, ch->counts);
  } else {
    long long threads = (long long)ch->n * (ch->pp / 16);
    cluster_histogram_kernel<<<cdiv(threads, 256), 256, 0, ch->st>>>(ch->X, ch->n, ch->pp, ch->c, ch->mmax, ch->H,
                                                                    ch->counts);
  }
  ch->h_launches++;
  SMG_CUDA(cudaGetLastError());
  return 0;
}

// common fields of a phi_update_kernel launch on the canonical parameter buffers
PhiArgs phi_args_base(smg_chain* ch, uint32_t sub) {
  PhiArgs A;
  memset(&A, 0, sizeof(A));
  A.pp = ch->pp;
  A.p = ch->p;
  A.mmax = ch->mmax;
  A.attr = ch->attr;
  A.v = ch->v;
  A.w = ch->w;
  A.H = ch->H;
  A.counts = ch->counts;
  A.njobs = 0;
  A.njobs_ptr = ch->K;
  A.cen_src = ch->cen[ch->cur];
  A.sig_src = ch->sig[ch->cur];
  A.cen = ch->cen[ch->cur];
  A.sig = ch->sig[ch->cur];
  A.isg = ch->isg[ch->cur];
  A.sden = ch->sden[ch->cur];
  A.u_stride = ch->p;
  A.key = mk_key(ch, sub);
  A.sigma_exact = ch->sigma_exact;
  A.status = ch->status;
  A.prof = ch->scan_prof + 8;
  A.nparts = 1;
  A.den = ch->den;
  A.part_cnt = ch->phi_cnt;
  return A;
}

// update_phi on every cluster (common_functions.cpp:511-591)
static int update_phi_all(smg_chain* ch, uint32_t sub, const double* uc, const double* us) {
  if (launch_histogram(ch)) return SMG_ERR_CUDA;
  PhiArgs A = phi_args_base(ch, sub);
  A.u_center = uc;
  A.u_sigma = us;
  A.prior = 0;
  A.nparts = ch->phi_parts;
  phi_update_kernel<<<ch->Kcap * A.nparts, 256, 0, ch->st>>>(A);
  ch->h_launches += 1;
  SMG_CUDA(cudaGetLastError());
  return 0;
}

// prior draws for clusters [0, K): sample_centers + sample_sigmas (launcher.cpp:46-48)
static int prior_phi_all(smg_chain* ch, int K) {
  PhiArgs A = phi_args_base(ch, SUB_INIT_PHI);
  A.prior = 1;
  phi_update_kernel<<<std::max(K, 1), 256, 0, ch->st>>>(A);
  ch->h_launches += 1;
  SMG_CUDA(cudaGetLastError());
  return 0;
}

static int launch_loglik(smg_chain* ch) {
  loglik_partial_kernel<<<ch->loglik_blocks, 256, 0, ch->st>>>(ch->X, ch->n, ch->pp, ch->c, ch->cen[ch->cur],
                                                              ch->isg[ch->cur], ch->sden[ch->cur], ch->partial);
  reduce_final_kernel<<<1, 256, 0, ch->st>>>(ch->partial, ch->loglik_blocks, ch->loglik_d);
  ch->h_launches += 2;
  SMG_CUDA(cudaGetLastError());
  return 0;
}

// aux pool (launcher.cpp:67-77, re-drawn at iter % 1000 == 0, :123-129)
static int draw_pool(smg_chain* ch, long long epoch_iter = -1) {
  long long total = ch->pool_size * ch->pp;
  RngKey key = mk_key(ch, SUB_POOL);
  if (epoch_iter >= 0) key.sweep = (uint32_t)epoch_iter;  // re-create the pool of an earlier refresh (resume)
  pool_draw_kernel<<<cdiv(total, 128), 128, 0, ch->st>>>(ch->pool_size, ch->pp, ch->p, ch->attr, ch->v, ch->w, key, ch->sigma_exact, ch->pcen,
                                                        ch->psig, ch->pisg, ch->pden);
  pool_sden_kernel<<<cdiv(ch->pool_size * 32, 256), 256, 0, ch->st>>>(ch->pool_size, ch->pp, ch->pden, ch->psden);
  ch->h_launches += 2;
  SMG_CUDA(cudaGetLastError());
  ch->pool_valid = true;
  return 0;
}

// one iteration of launcher.cpp:85-154 (without the snapshot)
//
// Stream plan of a steady iteration (Neal-8 pass and split-merge every iteration):
//   st     : [margin + scan + compaction] -> [update_phi] -> [split-merge proposal: ONE cluster] -> [patch] -> [log-lik]
//   st_aux :  aux columns of the next pass (HBM-bound gather), under the scan + update_phi
//   st_k1  :                                   likelihood block of the next pass (all other SMs), under the proposal
// The block of the next pass depends on the parameters fixed by update_phi except for the <= 2 columns an ACCEPTED
// proposal rewrites; those are re-evaluated by sm_ll_patch_kernel (a no-op otherwise; acceptance is rare at
// stationarity).
static int sweep(smg_chain* ch, bool timed) {
  SMG_CUDA(cudaMemsetAsync(ch->accepted_d, 0, sizeof(int), ch->st));
  if (timed) cudaEventRecord(ch->ev[0], ch->st);
  const bool pass_now = ch->neal8 && ch->iter % ch->n8_step == 0;
  const bool pass_next = ch->neal8 && (ch->iter + 1) % ch->n8_step == 0;
  const bool sm_now = ch->split_merge && ch->iter % ch->sam_step == 0;
  if (pass_now) {
    // where the aux columns of the next pass are evaluated: under the scan (default) or, SMG_AUX_AT=sm, next to the
    // likelihood block under the split-merge proposal (the HBM-bound gather slows the scan's dependent loads)
    static const bool aux_at_sm = [] { const char* e = getenv("SMG_AUX_AT"); return e && strcmp(e, "sm") == 0; }();
    const bool defer_aux = aux_at_sm && ch->split_merge && ch->iter % ch->sam_step == 0;
    int rc = neal8_pass(ch, nullptr, timed, !defer_aux);
    if (rc) return rc;
    rc = update_phi_all(ch, SUB_PHI_AFTER_SCAN, nullptr, nullptr);
    if (rc) return rc;
    if (defer_aux) {
      rc = prefetch_next_aux(ch);
      if (rc) return rc;
    }
  } else if (timed) {
    for (int q = 1; q <= 3; q++) cudaEventRecord(ch->ev[q], ch->st);
  }
  if (timed) cudaEventRecord(ch->ev[4], ch->st);
  const bool k1_early = ch->k1_overlap && pass_next && sm_now;
  if (k1_early) {
    SMG_CUDA(cudaEventRecord(ch->ev_phi_done, ch->st));
    SMG_CUDA(cudaStreamWaitEvent(ch->st_k1, ch->ev_phi_done, 0));
  }
  if (sm_now) {  // launched before the block so that its cluster gets its SMs first (and st outranks st_k1)
    int rc = sm_step(ch, nullptr);
    if (rc) return rc;
    ch->h_sm_props++;
  }
  if (k1_early) {
    if (launch_ll_block(ch, ch->st_k1)) return SMG_ERR_CUDA;
    SMG_CUDA(cudaEventRecord(ch->ev_k1_done, ch->st_k1));
  }
  if (timed) cudaEventRecord(ch->ev[5], ch->st);
  if (ch->iter % 1000 == 0 && ch->aux_mode == 0) {
    int rc = draw_pool(ch);
    if (rc) return rc;
  }
  if (timed) cudaEventRecord(ch->ev[6], ch->st);
  ch->k1_in_tail = false;
  if (pass_next) {
    // The next iteration starts with a Neal-8 pass on exactly this state: its likelihood block also yields the
    // full-data log-likelihood (common_functions.cpp:379-401), sum_i LL[i][c_i], without a second pass over X.
    if (k1_early) {
      SMG_CUDA(cudaStreamWaitEvent(ch->st, ch->ev_k1_done, 0));
      sm_ll_patch_kernel<<<148 * 2, 256, 0, ch->st>>>(ch->accepted_d, ch->sm->info, ch->X, ch->n, ch->pp, ch->cen[ch->cur],
                                                      ch->isg[ch->cur], ch->sden[ch->cur], ch->LL, ch->ldl);
      ch->h_launches++;
    } else {
      if (launch_ll_block(ch)) return SMG_ERR_CUDA;
      ch->k1_in_tail = true;
    }
    loglik_gather_kernel<<<ch->loglik_blocks, 256, 0, ch->st>>>(ch->LL, ch->ldl, ch->c, ch->n, ch->partial);
    reduce_final_kernel<<<1, 256, 0, ch->st>>>(ch->partial, ch->loglik_blocks, ch->loglik_d);
    ch->h_launches += 2;
    SMG_CUDA(cudaGetLastError());
    ch->ll_for_iter = ch->iter + 1;
  } else {
    int rc = launch_loglik(ch);
    if (rc) return rc;
  }
  if (timed) cudaEventRecord(ch->ev[7], ch->st);
  ch->iter++;
  ch->h_sweeps++;
  return 0;
}

static int sync_status(smg_chain* ch) {
  int st = 0, K = 0, acc = 0;
  double ll = 0;
  SMG_CUDA(cudaMemcpyAsync(&st, ch->status, 4, cudaMemcpyDeviceToHost, ch->st));
  SMG_CUDA(cudaMemcpyAsync(&K, ch->K, 4, cudaMemcpyDeviceToHost, ch->st));
  SMG_CUDA(cudaMemcpyAsync(&acc, ch->accepted_d, 4, cudaMemcpyDeviceToHost, ch->st));
  SMG_CUDA(cudaMemcpyAsync(&ll, ch->loglik_d, 8, cudaMemcpyDeviceToHost, ch->st));
  SMG_CUDA(cudaStreamSynchronize(ch->st));
  ch->h_K = K;
  ch->h_accepted = acc;
  ch->h_loglik = ll;
  if (st) {
    cudaMemsetAsync(ch->status, 0, 4, ch->st);
    return status_to_error(st);
  }
  return 0;
}

// ------------------------------------------------------------------------------------------
// creation
// ------------------------------------------------------------------------------------------
static int validate_cfg(const smg_config* cfg) {
  if (!cfg || cfg->n < 2 || cfg->p < 1) return fail(SMG_ERR_ARG, "need n >= 2 observations and p >= 1 attributes");
  if (!cfg->attrisize || !cfg->v || !cfg->w) return fail(SMG_ERR_ARG, "attrisize, v and w are required");
  if (!(cfg->gamma > 0)) return fail(SMG_ERR_ARG, "gamma must be positive");
  if (cfg->m_aux < 1 || cfg->m_aux > 32) return fail(SMG_ERR_ARG, "m (auxiliary components) must be in 1..32");
  for (int j = 0; j < cfg->p; j++) {
    if (cfg->attrisize[j] < 2 || cfg->attrisize[j] > SMG_MAX_LEVELS)
      return fail(SMG_ERR_ARG, "attrisize[j] must be in 2..64 (attribute " + std::to_string(j) + ")");
    // v_j <= 1 is accepted like the reference does (hyperg.cpp:359-376: no Beta(w+1, v-1) proposal exists, every such
    // draw goes through the inverse CDF); the density needs v_j + w_j > 0 only, v_j > 0 is what the model states
    if (!(cfg->v[j] > 0.0))
      return fail(SMG_ERR_ARG, "v[j] must be > 0 (attribute " + std::to_string(j) + ")");
    if (!(cfg->w[j] >= 0.0)) return fail(SMG_ERR_ARG, "w[j] must be >= 0");
  }
  if (cfg->thinning < 1 || cfg->n8_step_size < 1 || cfg->sam_step_size < 1)
    return fail(SMG_ERR_ARG, "thinning, n8_step_size and sam_step_size must be >= 1");
  if (cfg->t < 0 || cfg->r < 0) return fail(SMG_ERR_ARG, "t and r must be >= 0");
  if (cfg->pair_selection < 0 || cfg->pair_selection > 1) return fail(SMG_ERR_ARG, "pair_selection must be 0 or 1");
  if (cfg->aux_mode < 0 || cfg->aux_mode > 1) return fail(SMG_ERR_ARG, "aux_mode must be 0 (stored pool) or 1 (pool-free)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(SMG_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(SMG_ERR_ARG, "bad device ordinal");
  return 0;
}

static int create_common(const smg_config* cfg, smg_chain** out) {
  int rc = validate_cfg(cfg);
  if (rc) return rc;
  smg_chain* ch = new smg_chain();
  ch->n = cfg->n;
  ch->p = cfg->p;
  ch->pp = (cfg->p + 15) / 16 * 16;
  ch->m_aux = cfg->m_aux;
  ch->L = std::max(cfg->L, 1);
  ch->t = cfg->t;
  ch->r = cfg->r;
  ch->neal8 = cfg->neal8;
  ch->split_merge = cfg->split_merge;
  ch->n8_step = cfg->n8_step_size;
  ch->sam_step = cfg->sam_step_size;
  ch->thinning = cfg->thinning;
  ch->gamma = cfg->gamma;
  ch->seed = cfg->seed;
  ch->device = cfg->device;
  ch->sigma_exact = cfg->exact_sigma_inverse;
  ch->pair_det = cfg->pair_selection == 1;
  ch->aux_mode = cfg->aux_mode;
  {
    const char* e = getenv("SMG_NO_K1_OVERLAP");  // diagnostic: likelihood block after the proposal instead of beside it
    ch->k1_overlap = !(e && e[0] == '1');
  }
  ch->Kcap = cfg->max_clusters > 0 ? cfg->max_clusters : 192;
  ch->Kcap = std::min(ch->Kcap, SMG_MAX_ENTRIES - ch->m_aux);
  ch->Kcap = std::max(ch->Kcap, 8);
  ch->ldl = (ch->Kcap + 3) / 4 * 4;
  ch->NS = SMG_MAX_SLOTS;
  ch->NST = SMG_MAX_SLOTS + SM_NSLOTS;
  ch->pool_size = cfg->pool_size > 0 ? cfg->pool_size : (long long)cfg->n * cfg->m_aux * cfg->thinning;
  if (ch->aux_mode == 1) ch->pool_size = 1;  // pool-free: a one-entry placeholder keeps the pool plumbing valid
  ch->h_attr.assign(cfg->attrisize, cfg->attrisize + cfg->p);
  ch->h_v.assign(cfg->v, cfg->v + cfg->p);
  ch->h_w.assign(cfg->w, cfg->w + cfg->p);
  ch->mmax = *std::max_element(ch->h_attr.begin(), ch->h_attr.end());
  if (ch->L > ch->Kcap) {
    delete ch;
    return fail(SMG_ERR_CAPACITY, "L exceeds max_clusters");
  }
  rc = chain_alloc(ch);
  if (rc) {
    chain_free(ch);
    return rc;
  }
  std::vector<int> a(ch->pp, 2);
  std::vector<double> v(ch->pp, 2.0), w(ch->pp, 1.0);
  std::copy(ch->h_attr.begin(), ch->h_attr.end(), a.begin());
  std::copy(ch->h_v.begin(), ch->h_v.end(), v.begin());
  std::copy(ch->h_w.begin(), ch->h_w.end(), w.begin());
  if (h2d_sync(ch->attr, a.data(), ch->pp * 4, ch->st) != cudaSuccess ||
      h2d_sync(ch->v, v.data(), ch->pp * 8, ch->st) != cudaSuccess ||
      h2d_sync(ch->w, w.data(), ch->pp * 8, ch->st) != cudaSuccess) {
    chain_free(ch);
    return fail(SMG_ERR_CUDA, "upload of hyper-parameters failed");
  }
  *out = ch;
  return 0;
}

// initial allocation + prior phi + update_phi + pool (launcher.cpp:31-77)
static int init_state(smg_chain* ch, const int* c_init, int compact_init) {
  std::vector<int> c(ch->n);
  if (c_init) {
    int mn = *std::min_element(c_init, c_init + ch->n);
    for (int i = 0; i < ch->n; i++) c[i] = c_init[i] - mn;
  } else {
    init_assign_kernel<<<cdiv(ch->n, 256), 256, 0, ch->st>>>(ch->n, ch->L, nullptr, mk_key(ch, SUB_INIT), ch->c);
    ch->h_launches++;
    SMG_CUDA(cudaMemcpyAsync(c.data(), ch->c, (size_t)ch->n * 4, cudaMemcpyDeviceToHost, ch->st));
    SMG_CUDA(cudaStreamSynchronize(ch->st));
  }
  int mx = *std::max_element(c.begin(), c.end());
  if (mx >= ch->Kcap) return fail(SMG_ERR_CAPACITY, "initial labels exceed max_clusters");
  std::vector<int> cnt(mx + 1, 0);
  for (int x : c) cnt[x]++;
  int nuniq = 0;
  for (int x : cnt) nuniq += (x > 0);
  int K = c_init ? nuniq : ch->L;  // launcher.cpp:38 / :28
  if (nuniq != mx + 1 || nuniq != K) {
    if (!compact_init)
      // the reference runs into validate_state (common_functions.cpp:155-161) at the first move
      return fail(SMG_ERR_STATE, "State validation failed: inconsistent cluster count from initial assignment "
                                 "(a label in 0..K-1 is empty; pass contiguous labels or set compact_init)");
    std::vector<int> remap(mx + 1, -1);
    int nx = 0;
    for (int l = 0; l <= mx; l++)
      if (cnt[l] > 0) remap[l] = nx++;
    for (int& x : c) x = remap[x];
    K = nx;
  }
  std::vector<int> counts(ch->NST, 0);
  for (int x : c) counts[x]++;
  SMG_CUDA(cudaMemcpyAsync(ch->c, c.data(), (size_t)ch->n * 4, cudaMemcpyHostToDevice, ch->st));
  ch->hist_valid = false;
  SMG_CUDA(cudaMemcpyAsync(ch->counts, counts.data(), (size_t)ch->NST * 4, cudaMemcpyHostToDevice, ch->st));
  SMG_CUDA(cudaMemcpyAsync(ch->K, &K, 4, cudaMemcpyHostToDevice, ch->st));
  SMG_CUDA(cudaStreamSynchronize(ch->st));
  ch->h_K = K;
  int rc = prior_phi_all(ch, K);
  if (rc) return rc;
  rc = update_phi_all(ch, SUB_INIT_PHI + 1, nullptr, nullptr);
  if (rc) return rc;
  if (!ch->pool_valid) {  // normally already drawn by create_impl, under the upload
    rc = draw_pool(ch);
    if (rc) return rc;
  }
  return sync_status(ch);
}

}  // namespace smg
static int pinned_acquire(size_t bytes, uint8_t** out, int which);
static void pinned_release(uint8_t* p);
namespace smg {
// R's column-major fp64 matrix -> row-major padded u8 codes.  Small inputs are copied as they are and converted
// (and validated) on the device; large ones are packed by host threads first, which moves 8x fewer bytes over
// PCIe from pageable memory (205 MB -> 25.6 MB at the metric shape).
static int upload_colmajor(smg_chain* ch, const double* data) {
  const int n = ch->n, p = ch->p, pp = ch->pp;
  if ((long long)n * p < (1ll << 20)) {
    double* tmp = nullptr;
    int* bad = nullptr;
    SMG_CUDA(dev_malloc(&tmp, (size_t)n * p * 8, ch->st));
    SMG_CUDA(dev_malloc(&bad, 4, ch->st));
    SMG_CUDA(cudaMemsetAsync(bad, 0, 4, ch->st));
    SMG_CUDA(cudaMemcpyAsync(tmp, data, (size_t)n * p * 8, cudaMemcpyHostToDevice, ch->st));
    long long total = (long long)n * pp;
    ingest_colmajor_kernel<<<cdiv(total, 256), 256, 0, ch->st>>>(tmp, n, p, pp, ch->attr, ch->X, bad);
    ch->h_launches++;
    int hbad = 0;
    SMG_CUDA(cudaMemcpyAsync(&hbad, bad, 4, cudaMemcpyDeviceToHost, ch->st));
    SMG_CUDA(cudaStreamSynchronize(ch->st));
    cudaFreeAsync(tmp, ch->st);
    cudaFreeAsync(bad, ch->st);
    if (hbad) return fail(SMG_ERR_ARG, std::to_string(hbad) + " data entries are not integer codes in 1..attrisize[j]");
    return 0;
  }
  // packed into a cached page-locked buffer: no page faults after the first call and a faster copy
  uint8_t* buf = nullptr;
  if (pinned_acquire((size_t)n * pp, &buf, 1)) return fail(SMG_ERR_CUDA, "cudaHostAlloc of the upload staging buffer failed");
  // packing threads: the cores this process may run on (a rank pinned to its share of the host must not start 16
  // threads per rank), at most 16; SMG_PACK_THREADS overrides
  int nthr = (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  {
    cpu_set_t set;
    CPU_ZERO(&set);
    if (sched_getaffinity(0, sizeof(set), &set) == 0) nthr = std::max(1, std::min(nthr, CPU_COUNT(&set)));
    const char* e = getenv("SMG_PACK_THREADS");
    if (e && atoi(e) > 0) nthr = std::min(64, atoi(e));
  }
  std::vector<long long> nbad(nthr, 0);
  std::vector<std::thread> pool;
  const int* attr = ch->h_attr.data();
  for (int tix = 0; tix < nthr; tix++) {
    pool.emplace_back([&, tix]() {
      const int r0 = (int)((long long)n * tix / nthr), r1 = (int)((long long)n * (tix + 1) / nthr);
      long long bad = 0;
      for (int b0 = r0; b0 < r1; b0 += 64) {  // 64-row blocks: the destination block stays in cache
        const int b1 = std::min(r1, b0 + 64);
        for (int j = 0; j < p; j++) {
          const double* col = data + (size_t)n * j;
          const int mj = attr[j];
          for (int i = b0; i < b1; i++) {
            const double v = col[i];
            const int iv = (int)v;
            uint8_t o = 0;
            if ((double)iv != v || iv < 1 || iv > mj || iv > 255)
              bad++;
            else
              o = (uint8_t)iv;
            buf[(size_t)i * pp + j] = o;
          }
        }
        for (int i = b0; i < b1; i++)
          for (int j = p; j < pp; j++) buf[(size_t)i * pp + j] = 0;
      }
      nbad[tix] = bad;
    });
  }
  for (auto& t : pool) t.join();
  long long hbad = 0;
  for (long long b : nbad) hbad += b;
  if (hbad) {
    pinned_release(buf);
    return fail(SMG_ERR_ARG, std::to_string(hbad) + " data entries are not integer codes in 1..attrisize[j]");
  }
  cudaError_t ce = cudaMemcpyAsync(ch->X, buf, (size_t)n * pp, cudaMemcpyHostToDevice, ch->st);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(ch->st);
  pinned_release(buf);
  SMG_CUDA(ce);
  return 0;
}

static int upload_u8(smg_chain* ch, const unsigned char* data) {
  std::vector<uint8_t> buf((size_t)ch->n * ch->pp, 0);
  for (int i = 0; i < ch->n; i++)
    for (int j = 0; j < ch->p; j++) {
      unsigned char x = data[(size_t)i * ch->p + j];
      if (x < 1 || x > ch->h_attr[j]) return fail(SMG_ERR_ARG, "data entries must be codes in 1..attrisize[j]");
      buf[(size_t)i * ch->pp + j] = x;
    }
  SMG_CUDA(h2d_sync(ch->X, buf.data(), buf.size(), ch->st));
  return 0;
}

}  // namespace smg

using namespace smg;

// One cached page-locked staging buffer per process for the snapshot ring of smg_run_markov_chain:
// cudaHostAlloc / cudaFreeHost cost milliseconds to >100 ms per call (measured), far more than the run
// itself at small iteration counts.  A second concurrent caller simply allocates its own.
static std::mutex g_pin_mu;
struct PinSlot {
  uint8_t* buf = nullptr;
  size_t size = 0;
  bool busy = false;
};
static PinSlot g_pin[2];  // 0: snapshot ring of smg_run_markov_chain, 1: packed data of upload_colmajor
static int pinned_acquire(size_t bytes, uint8_t** out, int which) {
  std::lock_guard<std::mutex> lk(g_pin_mu);
  PinSlot& P = g_pin[which];
  if (!P.busy) {
    if (P.size < bytes) {
      if (P.buf) cudaFreeHost(P.buf);
      P.buf = nullptr;
      P.size = 0;
      if (cudaHostAlloc((void**)&P.buf, bytes, cudaHostAllocDefault) != cudaSuccess) return 1;
      P.size = bytes;
    }
    P.busy = true;
    *out = P.buf;
    return 0;
  }
  return cudaHostAlloc((void**)out, bytes, cudaHostAllocDefault) != cudaSuccess;
}
static void pinned_release(uint8_t* p) {
  std::lock_guard<std::mutex> lk(g_pin_mu);
  for (PinSlot& P : g_pin)
    if (p == P.buf) {
      P.busy = false;
      return;
    }
  if (p) cudaFreeHost(p);
}

struct smg_psm {
  int n = 0, device = 0, cap = 0, count = 0, kmax = 0;
  long long total = 0;
  int* psm = nullptr;
  bool owns = false;
  bool lower_stale = false;  // flushes accumulate the upper triangle of tiles only; smg_psm_finalize mirrors it down
  // distributed form (smg_chains_psm_distribute): rank / world of the communicator, rows per rank, the peers' matrices
  smg_comm* dist = nullptr;
  int rows_per = 0;
  int** peers_d = nullptr;          // device array [world] of peer-mapped matrix pointers (own pointer at [rank])
  std::vector<void*> peers_opened;  // IPC mappings to close
  uint8_t* labels = nullptr;
  cudaStream_t st = nullptr;
  unsigned long long launches = 0;
  double last_ms = 0.0;
  cudaEvent_t ev[2] = {};
};

// every rank of the communicator has reached this point and its stream `st` is idle
static int comm_rendezvous(smg_comm* C) {
  if (!C || C->world <= 1) return 0;
  NcclApi* N = nccl_api();
  SMG_NCCL(N->AllReduce(C->scratch, C->scratch, 1, ncclInt32, ncclSum, C->comm, C->st));
  SMG_CUDA(cudaStreamSynchronize(C->st));
  return 0;
}


template <int KP, int NST>
static int psm_launch(smg_psm* P) {
  const size_t smem = (size_t)NST * (PSM_M + PSM_N) * KP;
  SMG_CUDA(cudaFuncSetAttribute(psm_accumulate_kernel<KP, NST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(cdiv(P->n, PSM_N), cdiv(P->n, PSM_M));
  const char* mo = getenv("SMG_PSM_MMA_ONLY");
  static const bool full = [] { const char* e = getenv("SMG_PSM_FULL"); return e && e[0] == '1'; }();  // every tile (A/B measurements)
  psm_accumulate_kernel<KP, NST><<<grid, PSM_THREADS, smem, P->st>>>(P->labels, P->n, P->count, P->psm, (mo && mo[0] == '1') ? 1 : 0,
                                                                     full ? 0 : 1, P->peers_d, P->rows_per);
  P->lower_stale = true;
  SMG_CUDA(cudaGetLastError());
  return 0;
}

// ==========================================================================================
// extern "C"
// ==========================================================================================
extern "C" {

const char* smg_last_error(void) { return g_last_error.c_str(); }

int smg_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

static int create_impl(const smg_config* cfg, const double* dcol, const unsigned char* du8, const int* c_init,
                       smg_chain** out) {
  if (!out) return fail(SMG_ERR_ARG, "out is NULL");
  *out = nullptr;
  if (!dcol && !du8) return fail(SMG_ERR_ARG, "data is NULL");
  smg_chain* ch = nullptr;
  int rc = create_common(cfg, &ch);
  if (rc) return rc;
  // the initial auxiliary pool (launcher.cpp:67-77) depends on the hyper-parameters only: draw it first so that the
  // device works while the host packs and uploads the data
  rc = draw_pool(ch);
  if (!rc) rc = dcol ? upload_colmajor(ch, dcol) : upload_u8(ch, du8);
  if (!rc) rc = init_state(ch, c_init, cfg->compact_init);
  if (rc) {
    std::string keep = g_last_error;
    chain_free(ch);
    g_last_error = keep;
    return rc;
  }
  *out = ch;
  return 0;
}

int smg_create(const smg_config* cfg, const double* data_colmajor, const int* c_init, smg_chain** out) {
  return create_impl(cfg, data_colmajor, nullptr, c_init, out);
}
int smg_create_u8(const smg_config* cfg, const unsigned char* data_rowmajor, const int* c_init, smg_chain** out) {
  return create_impl(cfg, nullptr, data_rowmajor, c_init, out);
}

void smg_destroy(smg_chain* ch) { chain_free(ch); }

// launches n_iters iterations on the chain's stream without waiting for them
static int step_launch(smg_chain* ch, int n_iters) {
  SMG_CUDA(cudaSetDevice(ch->device));
  cudaEventRecord(ch->ev_call[0], ch->st);
  for (int it = 0; it < n_iters; it++) {
    int rc = sweep(ch, it == n_iters - 1);
    if (rc) return rc;
  }
  cudaEventRecord(ch->ev_call[1], ch->st);
  return 0;
}

// waits for the chain's stream, reads back status / K / log-likelihood / acceptance and the phase timings
static int step_finish(smg_chain* ch, int n_iters) {
  SMG_CUDA(cudaSetDevice(ch->device));
  int rc = sync_status(ch);
  {
    float ms = 0;
    cudaEventElapsedTime(&ms, ch->ev_call[0], ch->ev_call[1]);
    ch->h_step_ms = ms;
  }
  if (n_iters > 0) {
    const int a[7] = {0, 1, 2, 3, 4, 5, 6}, b[7] = {1, 2, 3, 4, 5, 6, 7};
    float ms;
    // [0] ll block, [1] aux, [2] scan(+finalize), [3] update_phi, [4] split-merge, [5] pool, [6] loglik ; [7] total
    for (int q = 0; q < 7; q++) {
      ms = 0;
      cudaEventElapsedTime(&ms, ch->ev[a[q]], ch->ev[b[q]]);
      ch->h_timings[q] = ms;
    }
    ms = 0;
    cudaEventElapsedTime(&ms, ch->ev[0], ch->ev[7]);
    ch->h_timings[7] = ms;
    if (ch->k1_timed) {
      // the likelihood block is launched either at the start of the pass or (normally) at the end of the previous
      // iteration, inside the [6]-[7] interval: report it as phase [0] and take it out of the log-likelihood phase
      ms = 0;
      cudaEventElapsedTime(&ms, ch->ev_k1[0], ch->ev_k1[1]);
      ch->h_timings[0] = ms;
      if (ch->k1_in_tail) ch->h_timings[6] = std::max(0.0, ch->h_timings[6] - (double)ms);
    }
    if (ch->aux_timed && cudaEventSynchronize(ch->ev_aux_t1) == cudaSuccess) {
      // the aux columns of the next pass ran on the side stream, overlapped with the split-merge step
      ms = 0;
      cudaEventElapsedTime(&ms, ch->ev_aux_t0, ch->ev_aux_t1);
      ch->h_timings[1] = ms;
    }
  }
  return rc;
}

int smg_step(smg_chain* ch, int n_iters) {
  if (!ch) return fail(SMG_ERR_ARG, "chain is NULL");
  int rc = step_launch(ch, n_iters);
  if (rc) return rc;
  return step_finish(ch, n_iters);
}

// n_iters iterations on each of `count` independent chains (possibly on different devices): the sweeps are
// launched iteration by iteration round-robin over the chains' streams so that they overlap on the GPU, and the
// host waits only once per chain at the end
int smg_step_many(smg_chain** chains, int count, int n_iters) {
  if (!chains || count < 0) return fail(SMG_ERR_ARG, "bad chain list");
  for (int q = 0; q < count; q++)
    if (!chains[q]) return fail(SMG_ERR_ARG, "chain is NULL");
  // A sweep is ~75 kernel launches and with many chains the host launch rate, not the GPU, bounds the throughput:
  // the chains are dealt to a few host threads, each launching its share iteration by iteration.
  // (large chains keep the GPU busy from one launching thread; measured at n=1e5 more threads only add contention)
  const char* env = getenv("SMG_STEP_THREADS");
  int nmax = 0;
  for (int q = 0; q < count; q++) nmax = std::max(nmax, chains[q]->n);
  int nthr = env ? atoi(env) : (nmax > 50000 ? 1 : (int)std::min(8u, std::max(1u, std::thread::hardware_concurrency())));
  nthr = std::max(1, std::min(nthr, count));
  std::vector<int> rcs(nthr, 0);
  std::vector<std::string> msgs(nthr);
  auto worker = [&](int w) {
    for (int q = w; q < count; q += nthr) {
      if (cudaSetDevice(chains[q]->device) != cudaSuccess) {
        rcs[w] = SMG_ERR_CUDA;
        msgs[w] = "cudaSetDevice failed";
        return;
      }
      cudaEventRecord(chains[q]->ev_call[0], chains[q]->st);
    }
    for (int it = 0; it < n_iters; it++)
      for (int q = w; q < count; q += nthr) {
        cudaSetDevice(chains[q]->device);
        chains[q]->many = count > 1;
        int rc = sweep(chains[q], it == n_iters - 1);
        chains[q]->many = false;
        if (rc) {
          rcs[w] = rc;
          msgs[w] = g_last_error;
          return;
        }
      }
    for (int q = w; q < count; q += nthr) {
      cudaSetDevice(chains[q]->device);
      cudaEventRecord(chains[q]->ev_call[1], chains[q]->st);
      int rc = step_finish(chains[q], n_iters);
      if (rc && !rcs[w]) {
        rcs[w] = rc;
        msgs[w] = g_last_error;
      }
    }
  };
  if (nthr == 1) {
    worker(0);
  } else {
    std::vector<std::thread> pool;
    for (int w = 0; w < nthr; w++) pool.emplace_back(worker, w);
    for (auto& t : pool) t.join();
  }
  for (int w = 0; w < nthr; w++)
    if (rcs[w]) {
      g_last_error = msgs[w];
      return rcs[w];
    }
  return 0;
}

int smg_snapshot(smg_chain* ch, int* K, int* c_i, double* centers, double* sigmas, int cap_clusters, double* loglik,
                 int* accepted) {
  if (!ch) return fail(SMG_ERR_ARG, "chain is NULL");
  SMG_CUDA(cudaSetDevice(ch->device));
  int rc = sync_status(ch);
  if (rc) return rc;
  const int Kh = ch->h_K;
  if (K) *K = Kh;
  if (loglik) *loglik = ch->h_loglik;
  if (accepted) *accepted = ch->h_accepted;
  if (c_i) SMG_CUDA(d2h_sync(c_i, ch->c, (size_t)ch->n * 4, ch->st));
  if (centers || sigmas) {
    if (Kh > cap_clusters) return fail(SMG_ERR_CAPACITY, "snapshot buffers hold fewer clusters than K");
    std::vector<uint8_t> hc((size_t)Kh * ch->pp);
    std::vector<double> hs((size_t)Kh * ch->pp);
    SMG_CUDA(d2h_sync(hc.data(), ch->cen[ch->cur], hc.size(), ch->st));
    SMG_CUDA(d2h_sync(hs.data(), ch->sig[ch->cur], hs.size() * 8, ch->st));
    for (int k = 0; k < Kh; k++)
      for (int j = 0; j < ch->p; j++) {
        if (centers) centers[(size_t)k * ch->p + j] = (double)hc[(size_t)k * ch->pp + j];
        if (sigmas) sigmas[(size_t)k * ch->p + j] = hs[(size_t)k * ch->pp + j];
      }
  }
  return 0;
}

int smg_get_stats(smg_chain* ch, unsigned long long* out8) {
  if (!ch || !out8) return fail(SMG_ERR_ARG, "NULL argument");
  SMG_CUDA(cudaSetDevice(ch->device));
  SMG_CUDA(cudaStreamSynchronize(ch->st));
  unsigned long long d[8];
  SMG_CUDA(d2h_sync(d, ch->stats_d, 64, ch->st));
  out8[0] = d[0];
  out8[1] = d[1];
  out8[2] = d[2];
  out8[3] = d[3];
  out8[4] = ch->h_sweeps;
  out8[5] = ch->h_launches;
  out8[6] = ch->h_sm_props;
  out8[7] = d[7];
  return 0;
}

// Checkpoint / resume (the reference has none: SURVEY section 5).  A chain is a deterministic function of
// (seed, iteration, allocation, parameters): smg_get_iteration + smg_snapshot save it; on a chain created with the same
// data, configuration and seed, smg_debug_set_state followed by smg_resume_at continues bit for bit -- the auxiliary
// pool is re-drawn from the Philox key of the refresh (launcher.cpp:123-129) that was current at `iteration`.
int smg_get_iteration(smg_chain* ch, long long* iteration) {
  if (!ch || !iteration) return fail(SMG_ERR_ARG, "NULL argument");
  *iteration = ch->iter;
  return 0;
}
int smg_resume_at(smg_chain* ch, long long iteration) {
  if (!ch || iteration < 0) return fail(SMG_ERR_ARG, "bad argument");
  SMG_CUDA(cudaSetDevice(ch->device));
  SMG_CUDA(cudaStreamSynchronize(ch->st_aux));
  SMG_CUDA(cudaStreamSynchronize(ch->st_k1));
  SMG_CUDA(cudaStreamSynchronize(ch->st));
  ch->iter = iteration;
  ch->ll_for_iter = -1;
  ch->aux_ready = false;
  const long long epoch = ((iteration - 1) / 1000) * 1000;  // iteration r (r % 1000 == 0) re-draws the pool AFTER its pass
  int rc = draw_pool(ch, epoch < 0 ? 0 : epoch);
  if (rc) return rc;
  return sync_status(ch);
}

// validate_state (common_functions.cpp:146-172) of the current state; the reference runs it after every move
int smg_validate_state(smg_chain* ch) {
  if (!ch) return fail(SMG_ERR_ARG, "chain is NULL");
  SMG_CUDA(cudaSetDevice(ch->device));
  int* tmp = nullptr;
  SMG_CUDA(dev_malloc(&tmp, (size_t)ch->Kcap * sizeof(int), ch->st));
  SMG_CUDA(cudaMemsetAsync(tmp, 0, (size_t)ch->Kcap * sizeof(int), ch->st));
  validate_recount_kernel<<<cdiv(ch->n, 256), 256, 0, ch->st>>>(ch->c, ch->n, ch->K, tmp, ch->status);
  validate_compare_kernel<<<cdiv(ch->Kcap, 256), 256, 0, ch->st>>>(ch->K, ch->Kcap, tmp, ch->counts, ch->status);
  ch->h_launches += 2;
  SMG_CUDA(cudaGetLastError());
  cudaFreeAsync(tmp, ch->st);
  return sync_status(ch);
}

// synthetic Hamming-mixture data generated on `device` and copied back: X [n][p] codes 1..attrisize[j], labels [n]
// (components as equal as possible, interleaved), centres [k_true][p]
int smg_synth_generate(int n, int p, const int* attrisize, int k_true, double s, unsigned long long seed, int device,
                       unsigned char* X_out, int* labels_out, unsigned char* centres_out) {
  if (n < 1 || p < 1 || !attrisize || k_true < 1 || k_true > 255 || !(s > 0) || !X_out || !labels_out || !centres_out)
    return fail(SMG_ERR_ARG, "bad argument");
  for (int j = 0; j < p; j++)
    if (attrisize[j] < 2 || attrisize[j] > SMG_MAX_LEVELS) return fail(SMG_ERR_ARG, "attrisize[j] must be in 2..64");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(SMG_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  if (device < 0 || device >= ndev) return fail(SMG_ERR_ARG, "bad device ordinal");
  SMG_CUDA(cudaSetDevice(device));
  const int pp = (p + 15) / 16 * 16;
  // labels and centres are tiny host-side draws from the same Philox generator
  std::vector<int> lab(n);
  std::vector<uint8_t> cen((size_t)k_true * pp, 0);
  for (int i = 0; i < n; i++) lab[i] = i % k_true;
  RngKey key;
  key.k0 = (uint32_t)seed;
  key.k1 = (uint32_t)(seed >> 32);
  key.sweep = 0;
  key.sub = 0;
  for (int k = 0; k < k_true; k++)
    for (int j = 0; j < p; j++) {
      uint32_t r[4];
      philox4x32_10((uint32_t)k, (uint32_t)j, 78u, 0u, key.k0, key.k1, r);
      cen[(size_t)k * pp + j] = (uint8_t)(1 + r[0] % (uint32_t)attrisize[j]);
    }
  std::vector<int> attr(pp, 2);
  std::copy(attrisize, attrisize + p, attr.begin());
  int *d_attr = nullptr, *d_lab = nullptr;
  uint8_t *d_cen = nullptr, *d_X = nullptr;
  SMG_CUDA(cudaMalloc(&d_attr, (size_t)pp * 4));
  SMG_CUDA(cudaMalloc(&d_lab, (size_t)n * 4));
  SMG_CUDA(cudaMalloc(&d_cen, cen.size()));
  SMG_CUDA(cudaMalloc(&d_X, (size_t)n * pp));
  SMG_CUDA(cudaMemcpy(d_attr, attr.data(), (size_t)pp * 4, cudaMemcpyHostToDevice));
  SMG_CUDA(cudaMemcpy(d_lab, lab.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
  SMG_CUDA(cudaMemcpy(d_cen, cen.data(), cen.size(), cudaMemcpyHostToDevice));
  synth_generate_kernel<<<cdiv((long long)n * pp, 256), 256>>>(n, p, pp, d_attr, d_cen, d_lab, s, key, d_X);
  SMG_CUDA(cudaGetLastError());
  SMG_CUDA(cudaMemcpy2D(X_out, p, d_X, pp, p, n, cudaMemcpyDeviceToHost));
  cudaFree(d_attr);
  cudaFree(d_lab);
  cudaFree(d_cen);
  cudaFree(d_X);
  std::copy(lab.begin(), lab.end(), labels_out);
  for (int k = 0; k < k_true; k++)
    for (int j = 0; j < p; j++) centres_out[(size_t)k * p + j] = cen[(size_t)k * pp + j];
  return 0;
}

int smg_debug_scan_spec(smg_chain* ch, int mode, unsigned long long* out4) {
  if (!ch) return fail(SMG_ERR_ARG, "NULL argument");
  if (mode < -1 || mode > 2) return fail(SMG_ERR_ARG, "mode must be -1, 0, 1 or 2");
  SMG_CUDA(cudaSetDevice(ch->device));
  ch->scan_spec = mode;
  if (out4) {
    SMG_CUDA(cudaStreamSynchronize(ch->st));
    unsigned long long d[8];
    SMG_CUDA(d2h_sync(d, ch->stats_d, 64, ch->st));
    out4[0] = d[4];
    out4[1] = d[5];
    out4[2] = d[6];
    out4[3] = 0;
  }
  return 0;
}

int smg_debug_scan_profile(smg_chain* ch, unsigned long long* out8) {
  if (!ch || !out8) return fail(SMG_ERR_ARG, "NULL argument");
  SMG_CUDA(cudaSetDevice(ch->device));
  SMG_CUDA(cudaStreamSynchronize(ch->st));
  SMG_CUDA(d2h_sync(out8, ch->scan_prof, 128, ch->st));  // 16 counters
#ifdef SMG_PHI_PROFILE
  SMG_CUDA(d2h_sync(out8, ch->scan_prof + 8, 64, ch->st));  // phi_update / chain counters instead
#endif
  SMG_CUDA(cudaMemsetAsync(ch->scan_prof, 0, 128, ch->st));
  SMG_CUDA(cudaStreamSynchronize(ch->st));
  return 0;
}

int smg_debug_sm_profile(smg_chain* ch, unsigned long long* out16) {
  if (!ch || !out16) return fail(SMG_ERR_ARG, "NULL argument");
  memset(out16, 0, 64 * sizeof(unsigned long long));
  SMG_CUDA(cudaSetDevice(ch->device));
  SMG_CUDA(cudaStreamSynchronize(ch->st));
  if (!ch->sm || !ch->sm->smc || !ch->sm->smc->prof) return 0;
  SMG_CUDA(d2h_sync(out16, ch->sm->smc->prof, 64 * 8, ch->st));
  SMG_CUDA(cudaMemsetAsync(ch->sm->smc->prof, 0, 64 * 8, ch->st));
  SMG_CUDA(cudaStreamSynchronize(ch->st));
  return 0;
}

int smg_last_step_ms(smg_chain* ch, double* ms) {
  if (!ch || !ms) return fail(SMG_ERR_ARG, "NULL argument");
  *ms = ch->h_step_ms;
  return 0;
}

int smg_get_timings(smg_chain* ch, double* out8) {
  if (!ch || !out8) return fail(SMG_ERR_ARG, "NULL argument");
  for (int q = 0; q < 8; q++) out8[q] = ch->h_timings[q];
  return 0;
}

// ------------------------------------------------------------------------------------------
// whole run: code/launcher.cpp:7-174
// ------------------------------------------------------------------------------------------
void smg_free_results(smg_results* r) {
  if (!r) return;
  free(r->total_cls);
  free(r->c_i);
  free(r->phi_offset);
  free(r->centers);
  free(r->sigmas);
  free(r->loglikelihood);
  free(r->final_ass);
  free(r->accepted);
  memset(r, 0, sizeof(*r));
}

int smg_run_markov_chain(const double* data, int n, int p, const int* attrisize, double gamma, const double* v,
                         const double* w, int verbose, int m, int iterations, int L, const int* c_i, int burnin, int t,
                         int r, int neal8, int split_merge, int n8_step_size, int sam_step_size, int thinning,
                         unsigned long long seed, int device, smg_results* out) {
  if (!out) return fail(SMG_ERR_ARG, "out is NULL");
  memset(out, 0, sizeof(*out));
  if (iterations < 0 || burnin < 0) return fail(SMG_ERR_ARG, "iterations and burnin must be >= 0");
  smg_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.n = n;
  cfg.p = p;
  cfg.attrisize = attrisize;
  cfg.gamma = gamma;
  cfg.v = v;
  cfg.w = w;
  cfg.m_aux = m;
  cfg.L = L;
  cfg.t = t;
  cfg.r = r;
  cfg.neal8 = neal8;
  cfg.split_merge = split_merge;
  cfg.n8_step_size = n8_step_size;
  cfg.sam_step_size = sam_step_size;
  cfg.thinning = thinning;
  cfg.seed = seed;
  cfg.device = device;
  {
    // cluster capacity: 192 by default; a start with more clusters than that (the scripts' "all different" start,
    // c_i = seq(1, n)) gets what the allocation draw can hold, K + m <= 256 (divergences)
    const char* e = getenv("SMG_MAX_CLUSTERS");
    int want = e ? atoi(e) : 0;
    if (c_i && n > 0) {
      std::vector<int> tmp(c_i, c_i + n);
      std::sort(tmp.begin(), tmp.end());
      const int k0 = (int)(std::unique(tmp.begin(), tmp.end()) - tmp.begin());
      if (k0 + 16 > 192) want = std::max(want, k0 + 16);
    } else if (L + 16 > 192) {
      want = std::max(want, L + 16);
    }
    if (want > 0) cfg.max_clusters = std::min(want, SMG_MAX_ENTRIES - m);
  }
  smg_chain* ch = nullptr;
  auto tc0 = std::chrono::steady_clock::now();
  int rc = smg_create(&cfg, data, c_i, &ch);
  if (rc) return rc;
  const double create_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - tc0).count();
  out->iterations = iterations;
  out->n = n;
  out->p = p;
  out->total_cls = (int*)calloc(std::max(iterations, 1), sizeof(int));
  out->c_i = (int*)calloc((size_t)std::max(iterations, 1) * n, sizeof(int));
  out->phi_offset = (long long*)calloc(iterations + 1, sizeof(long long));
  out->loglikelihood = (double*)calloc(std::max(iterations, 1), sizeof(double));
  out->final_ass = (int*)calloc(n, sizeof(int));
  out->accepted = (int*)calloc(std::max(iterations, 1), sizeof(int));
  // centres / sigmas of the kept iterations grow in malloc'ed buffers that are handed to the caller as they are
  // (realloc of a large block is a remap, and there is no final copy)
  struct Grow {
    double* p = nullptr;
    size_t size = 0, cap = 0;
    bool extend(size_t add) {
      if (size + add > cap) {
        const size_t nc = std::max<size_t>(std::max<size_t>(2 * cap, size + add), 1 << 16);
        double* q = (double*)realloc(p, nc * sizeof(double));
        if (!q) return false;
        p = q;
        cap = nc;
      }
      size += add;
      return true;
    }
  } cen, sg;
  // Snapshots (launcher.cpp:140-153) leave the device asynchronously: every kept iteration enqueues its copies
  // into one of RING pinned slots behind the sweep on the chain's stream and the host only waits for a slot
  // when it comes round again, so sweeps are launched back to back.
  const int RING = 8;
  const int pp = ch->pp, Kcap = ch->Kcap;
  struct Slot {
    int* hdr;  // K, status, accepted
    double* ll;
    int* c;
    uint8_t* cen;
    double* sig;
    cudaEvent_t ev;
    long long result_slot;
  };
  const size_t c_bytes = ((size_t)n * 4 + 63) & ~(size_t)63;  // every region of a slot starts 64-byte aligned
  const size_t slot_bytes = 64 + c_bytes + (((size_t)Kcap * pp + 63) & ~(size_t)63) + (size_t)Kcap * pp * 8;
  uint8_t* pinned = nullptr;
  auto th0 = std::chrono::steady_clock::now();
  if (pinned_acquire(slot_bytes * RING, &pinned, 0)) {
    smg_destroy(ch);
    smg_free_results(out);
    return fail(SMG_ERR_CUDA, "cudaHostAlloc of the snapshot ring failed");
  }
  const double pin_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - th0).count();
  Slot ring[RING];
  for (int q = 0; q < RING; q++) {
    uint8_t* base = pinned + slot_bytes * q;
    ring[q].hdr = (int*)base;
    ring[q].ll = (double*)(base + 16);
    ring[q].c = (int*)(base + 64);
    ring[q].cen = base + 64 + c_bytes;
    ring[q].sig = (double*)(ring[q].cen + (((size_t)Kcap * pp + 63) & ~(size_t)63));
    ring[q].result_slot = -1;
    cudaEventCreateWithFlags(&ring[q].ev, cudaEventDisableTiming);
  }
  const size_t cen_off = 64 + c_bytes, sig_off = cen_off + (((size_t)Kcap * pp + 63) & ~(size_t)63);
  unsigned char* dstage = nullptr;  // device-side staging blocks, one per ring slot
  cudaStream_t st_copy = nullptr;
  cudaEvent_t ev_packed[RING];
  if (dev_malloc(&dstage, slot_bytes * RING, ch->st) != cudaSuccess ||
      cudaStreamCreateWithFlags(&st_copy, cudaStreamNonBlocking) != cudaSuccess) {
    pinned_release(pinned);
    smg_destroy(ch);
    smg_free_results(out);
    return fail(SMG_ERR_CUDA, "allocation of the snapshot staging failed");
  }
  for (int q = 0; q < RING; q++) cudaEventCreateWithFlags(&ev_packed[q], cudaEventDisableTiming);
  long long n_consumed = 0;  // kept iterations already copied into the result block (in order)
  auto consume = [&](Slot& S) -> int {
    if (S.result_slot < 0) return 0;
    if (cudaEventSynchronize(S.ev) != cudaSuccess) return fail(SMG_ERR_CUDA, "snapshot copy failed");
    const long long slot = S.result_slot;
    S.result_slot = -1;
    if (S.hdr[1]) return status_to_error(S.hdr[1]);
    n_consumed = slot + 1;
    const int K = S.hdr[0];
    if (K > Kcap) return fail(SMG_ERR_CAPACITY, "number of clusters exceeds max_clusters");
    out->total_cls[slot] = K;
    out->accepted[slot] = S.hdr[2];
    out->loglikelihood[slot] = *S.ll;
    memcpy(out->c_i + (size_t)slot * n, S.c, (size_t)n * 4);
    out->phi_offset[slot + 1] = out->phi_offset[slot] + K;
    const size_t o0 = cen.size;
    if (!cen.extend((size_t)K * p) || !sg.extend((size_t)K * p)) return fail(SMG_ERR_ARG, "out of host memory");
    for (int k = 0; k < K; k++)
      for (int j = 0; j < p; j++) {
        cen.p[o0 + (size_t)k * p + j] = (double)S.cen[(size_t)k * pp + j];
        sg.p[o0 + (size_t)k * p + j] = S.sig[(size_t)k * pp + j];
      }
    return 0;
  };
  auto t0 = std::chrono::steady_clock::now();
  const long long total = (long long)(iterations + burnin) * thinning;
  long long nkept = 0;
  // The snapshots are unpacked into the result block by a second host thread, in order, while this one keeps
  // launching sweeps: at the metric shape unpacking (0.4 MB of labels, K x p parameters, first-touch page faults of the
  // result block) plus the ~13 launches of a sweep did not fit into the 0.53 ms the GPU needs for it.
  std::atomic<long long> produced{0}, consumed{0};
  std::atomic<int> cons_rc{0};
  std::atomic<bool> cons_stop{false};
  std::string cons_msg;
  std::thread consumer([&] {
    cudaSetDevice(device);
    for (long long k = 0;; k++) {
      while (k >= produced.load(std::memory_order_acquire)) {
        if (cons_stop.load(std::memory_order_acquire)) return;
        std::this_thread::sleep_for(std::chrono::microseconds(20));
      }
      const int r = consume(ring[k % RING]);
      if (r) {
        cons_msg = g_last_error;
        cons_rc.store(r, std::memory_order_release);
        return;
      }
      consumed.store(k + 1, std::memory_order_release);
    }
  });
  for (long long iter = 0; iter < total && !rc; ++iter) {
    rc = sweep(ch, false);
    if (rc) break;
    if (verbose == 1 || verbose == 2) fprintf(stderr, "[DEBUG] - Iteration %lld of %d\n", iter, iterations + burnin);
    if (iter >= (long long)thinning * burnin && iter % thinning == 0) {
      Slot& S = ring[nkept % RING];
      // the slot's previous occupant (RING snapshots old) must have been unpacked
      while (consumed.load(std::memory_order_acquire) < nkept - RING + 1 && !cons_rc.load(std::memory_order_acquire))
        std::this_thread::sleep_for(std::chrono::microseconds(10));
      if (cons_rc.load(std::memory_order_acquire)) break;
      const int cur = ch->cur;
      unsigned char* dslot = dstage + slot_bytes * (nkept % RING);
      snapshot_pack_kernel<<<148, 256, 0, ch->st>>>(ch->K, ch->status, ch->accepted_d, ch->loglik_d, ch->c, n, ch->cen[cur],
                                                    ch->sig[cur], Kcap * pp, cen_off, sig_off, dslot);
      ch->h_launches++;
      cudaEventRecord(ev_packed[nkept % RING], ch->st);
      cudaStreamWaitEvent(st_copy, ev_packed[nkept % RING], 0);
      cudaMemcpyAsync(pinned + slot_bytes * (nkept % RING), dslot, slot_bytes, cudaMemcpyDeviceToHost, st_copy);
      if (cudaEventRecord(S.ev, st_copy) != cudaSuccess) {
        rc = fail(SMG_ERR_CUDA, "cudaEventRecord failed");
        break;
      }
      S.result_slot = iter / thinning - burnin;
      nkept++;
      produced.store(nkept, std::memory_order_release);
    }
  }
  // let the consumer drain the ring in result order
  while (!rc && consumed.load(std::memory_order_acquire) < nkept && !cons_rc.load(std::memory_order_acquire))
    std::this_thread::sleep_for(std::chrono::microseconds(20));
  cons_stop.store(true, std::memory_order_release);
  consumer.join();
  if (!rc && cons_rc.load()) rc = fail(cons_rc.load(), cons_msg);
  if (!rc) rc = smg_snapshot(ch, nullptr, out->final_ass, nullptr, nullptr, 0, nullptr, nullptr);  // also checks the status
  auto t1 = std::chrono::steady_clock::now();
  out->seconds = std::chrono::duration<double>(t1 - t0).count();
  out->time = (long long)std::chrono::duration_cast<std::chrono::seconds>(t1 - t0).count();
  // A chain that outgrows the cluster capacity stops there, but what it produced is NOT thrown away (the reference
  // would have lost everything to the exception, launcher.cpp:155-161): the block keeps the iterations completed so far,
  // `iterations` says how many, and the status is still SMG_ERR_CAPACITY.
  const bool partial = rc == SMG_ERR_CAPACITY && n_consumed > 0;
  if (!rc || partial) {
    if (partial) {
      out->iterations = (int)n_consumed;
      cen.size = (size_t)out->phi_offset[n_consumed] * p;
      sg.size = cen.size;
      if (n_consumed > 0) memcpy(out->final_ass, out->c_i + (size_t)(n_consumed - 1) * n, (size_t)n * 4);
    }
    if (!cen.p) cen.extend(1), cen.size = 0;
    if (!sg.p) sg.extend(1), sg.size = 0;
    out->centers = cen.p;
    out->sigmas = sg.p;
    cen.p = sg.p = nullptr;
  }
  cudaStreamSynchronize(st_copy);
  cudaStreamSynchronize(ch->st);
  for (int q = 0; q < RING; q++) {
    cudaEventDestroy(ring[q].ev);
    cudaEventDestroy(ev_packed[q]);
  }
  cudaStreamDestroy(st_copy);
  cudaFreeAsync(dstage, ch->st);
  auto tf0 = std::chrono::steady_clock::now();
  pinned_release(pinned);
  const double unpin_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - tf0).count();
  std::string keep = g_last_error;
  auto td0 = std::chrono::steady_clock::now();
  smg_destroy(ch);
  if (verbose == 3)
    fprintf(stderr, "[smgibbs] create %.4f s, pinned alloc %.4f s, loop %.4f s, pinned free %.4f s, destroy %.4f s\n", create_s,
            pin_s, out->seconds, unpin_s, std::chrono::duration<double>(std::chrono::steady_clock::now() - td0).count());
  free(cen.p);  // (null once handed over)
  free(sg.p);
  if (rc) {
    if (!partial) smg_free_results(out);
    g_last_error = keep + (partial ? " (the result block holds the " + std::to_string(n_consumed) + " iterations kept before)" : "");
  }
  return rc;
}

// ------------------------------------------------------------------------------------------
// parity hooks
// ------------------------------------------------------------------------------------------
int smg_debug_set_state(smg_chain* ch, int K, const int* c_i, const double* centers, const double* sigmas) {
  if (!ch || !c_i || !centers || !sigmas) return fail(SMG_ERR_ARG, "NULL argument");
  if (K < 1 || K > ch->Kcap) return fail(SMG_ERR_CAPACITY, "K out of range");
  ch->ll_for_iter = -1;  // the state is about to change outside a sweep: the cached likelihood block is stale
  SMG_CUDA(cudaSetDevice(ch->device));
  SMG_CUDA(cudaStreamSynchronize(ch->st));
  std::vector<uint8_t> hc((size_t)K * ch->pp, 0);
  std::vector<double> hs((size_t)K * ch->pp, 1.0);
  for (int k = 0; k < K; k++)
    for (int j = 0; j < ch->p; j++) {
      hc[(size_t)k * ch->pp + j] = (uint8_t)centers[(size_t)k * ch->p + j];
      hs[(size_t)k * ch->pp + j] = sigmas[(size_t)k * ch->p + j];
    }
  std::vector<int> counts(ch->NST, 0);
  for (int i = 0; i < ch->n; i++) {
    if (c_i[i] < 0 || c_i[i] >= K) return fail(SMG_ERR_ARG, "label out of range");
    counts[c_i[i]]++;
  }
  SMG_CUDA(h2d_sync(ch->cen[ch->cur], hc.data(), hc.size(), ch->st));
  SMG_CUDA(h2d_sync(ch->sig[ch->cur], hs.data(), hs.size() * 8, ch->st));
  SMG_CUDA(h2d_sync(ch->c, c_i, (size_t)ch->n * 4, ch->st));
  ch->hist_valid = false;
  SMG_CUDA(h2d_sync(ch->counts, counts.data(), counts.size() * 4, ch->st));
  SMG_CUDA(h2d_sync(ch->K, &K, 4, ch->st));
  derive_terms_kernel<<<K, 256, 0, ch->st>>>(K, ch->pp, ch->p, ch->attr, ch->sig[ch->cur], ch->isg[ch->cur], ch->den,
                                             ch->sden[ch->cur]);
  ch->h_launches++;
  SMG_CUDA(cudaGetLastError());
  SMG_CUDA(cudaStreamSynchronize(ch->st));
  ch->h_K = K;
  return 0;
}

int smg_debug_set_pool(smg_chain* ch, long long pool_size, const double* pool_center, const double* pool_sigma) {
  if (!ch || !pool_center || !pool_sigma) return fail(SMG_ERR_ARG, "NULL argument");
  if (ch->aux_mode == 1) return fail(SMG_ERR_ARG, "a pool-free chain (aux_mode 1) has no pool to inject");
  if (pool_size < 1 || pool_size > ch->pool_size) return fail(SMG_ERR_ARG, "pool_size exceeds the allocated pool");
  SMG_CUDA(cudaSetDevice(ch->device));
  SMG_CUDA(cudaStreamSynchronize(ch->st_aux));  // a prefetched aux pass may still be reading the old pool
  ch->aux_ready = false;
  SMG_CUDA(cudaStreamSynchronize(ch->st));
  std::vector<uint8_t> hc((size_t)pool_size * ch->pp, 0);
  std::vector<double> hs((size_t)pool_size * ch->pp, 1.0);
  for (long long e = 0; e < pool_size; e++)
    for (int j = 0; j < ch->p; j++) {
      hc[(size_t)e * ch->pp + j] = (uint8_t)pool_center[(size_t)e * ch->p + j];
      hs[(size_t)e * ch->pp + j] = pool_sigma[(size_t)e * ch->p + j];
    }
  SMG_CUDA(h2d_sync(ch->pcen, hc.data(), hc.size(), ch->st));
  SMG_CUDA(h2d_sync(ch->psig, hs.data(), hs.size() * 8, ch->st));
  // derive 1/sigma, den and their per-entry sums, in chunks of 65535 entries (grid.x limit is larger, keep simple)
  for (long long e0 = 0; e0 < pool_size; e0 += 1 << 20) {
    int cnt = (int)std::min<long long>(1 << 20, pool_size - e0);
    derive_terms_kernel<<<cnt, 256, 0, ch->st>>>(cnt, ch->pp, ch->p, ch->attr, ch->psig + (size_t)e0 * ch->pp,
                                                 ch->pisg + (size_t)e0 * ch->pp, ch->pden + (size_t)e0 * ch->pp,
                                                 ch->psden + e0);
    ch->h_launches++;
  }
  SMG_CUDA(cudaGetLastError());
  SMG_CUDA(cudaStreamSynchronize(ch->st));
  ch->pool_size = pool_size;  // shrink the logical pool to what was injected
  ch->pool_valid = true;
  return 0;
}

int smg_debug_get_pool(smg_chain* ch, long long first, long long count, double* pool_center, double* pool_sigma) {
  if (!ch) return fail(SMG_ERR_ARG, "NULL argument");
  if (first < 0 || first + count > ch->pool_size) return fail(SMG_ERR_ARG, "pool range out of bounds");
  SMG_CUDA(cudaSetDevice(ch->device));
  SMG_CUDA(cudaStreamSynchronize(ch->st));
  std::vector<uint8_t> hc((size_t)count * ch->pp);
  std::vector<double> hs((size_t)count * ch->pp);
  SMG_CUDA(d2h_sync(hc.data(), ch->pcen + (size_t)first * ch->pp, hc.size(), ch->st));
  SMG_CUDA(d2h_sync(hs.data(), ch->psig + (size_t)first * ch->pp, hs.size() * 8, ch->st));
  for (long long e = 0; e < count; e++)
    for (int j = 0; j < ch->p; j++) {
      if (pool_center) pool_center[(size_t)e * ch->p + j] = hc[(size_t)e * ch->pp + j];
      if (pool_sigma) pool_sigma[(size_t)e * ch->p + j] = hs[(size_t)e * ch->pp + j];
    }
  return 0;
}

int smg_debug_ll_block(smg_chain* ch, double* LL, int* mism) {
  if (!ch) return fail(SMG_ERR_ARG, "NULL argument");
  SMG_CUDA(cudaSetDevice(ch->device));
  int rc = sync_status(ch);
  if (rc) return rc;
  const int K = ch->h_K;
  if (launch_ll_block(ch)) return SMG_ERR_CUDA;
  SMG_CUDA(cudaStreamSynchronize(ch->st));
  if (LL) {
    std::vector<double> h((size_t)ch->n * ch->ldl);
    SMG_CUDA(d2h_sync(h.data(), ch->LL, h.size() * 8, ch->st));
    for (int i = 0; i < ch->n; i++)
      for (int k = 0; k < K; k++) LL[(size_t)i * K + k] = h[(size_t)i * ch->ldl + k];
  }
  if (mism) {
    int* d = nullptr;
    SMG_CUDA(cudaMalloc(&d, (size_t)ch->n * K * 4));
    long long warps = (long long)ch->n * K;
    mismatch_count_kernel<<<cdiv(warps * 32, 256), 256, 0, ch->st>>>(ch->X, ch->n, ch->pp, ch->cen[ch->cur], K, d);
    ch->h_launches++;
    SMG_CUDA(cudaGetLastError());
    SMG_CUDA(cudaStreamSynchronize(ch->st));
    SMG_CUDA(d2h_sync(mism, d, (size_t)ch->n * K * 4, ch->st));
    cudaFree(d);
  }
  return 0;
}

// device time of the likelihood-block kernel ALONE (nothing else on the GPU), averaged over `reps` launches on the current
// state; a memset of the whole LL buffer (larger than L2) between launches keeps every launch cold, as inside a sweep
int smg_debug_time_ll_block(smg_chain* ch, int reps, double* ms_avg) {
  if (!ch || !ms_avg || reps < 1) return fail(SMG_ERR_ARG, "bad argument");
  SMG_CUDA(cudaSetDevice(ch->device));
  int rc = sync_status(ch);
  if (rc) return rc;
  SMG_CUDA(cudaStreamSynchronize(ch->st_aux));
  SMG_CUDA(cudaStreamSynchronize(ch->st_k1));
  double acc = 0.0;
  for (int q = 0; q < reps; q++) {
    SMG_CUDA(cudaMemsetAsync(ch->LL, 0, (size_t)ch->n * ch->ldl * sizeof(double), ch->st));
    if (launch_ll_block(ch)) return SMG_ERR_CUDA;
    SMG_CUDA(cudaStreamSynchronize(ch->st));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ch->ev_k1[0], ch->ev_k1[1]);
    acc += ms;
  }
  ch->ll_for_iter = -1;
  *ms_avg = acc / reps;
  return 0;
}

int smg_debug_neal8_scan(smg_chain* ch, const double* tape) {
  if (!ch) return fail(SMG_ERR_ARG, "NULL argument");
  SMG_CUDA(cudaSetDevice(ch->device));
  const double* td = nullptr;
  if (tape) {
    SMG_CUDA(h2d_sync(ch->tape_d, tape, (size_t)ch->n * (ch->m_aux + 1) * 8, ch->st));
    td = ch->tape_d;
  }
  int rc = neal8_pass(ch, td, true);
  if (rc) return rc;
  rc = sync_status(ch);
  float ms = 0;
  cudaEventElapsedTime(&ms, ch->ev[0], ch->ev[1]);
  ch->h_timings[0] = ms;
  cudaEventElapsedTime(&ms, ch->ev[1], ch->ev[2]);
  ch->h_timings[1] = ms;
  cudaEventElapsedTime(&ms, ch->ev[2], ch->ev[3]);
  ch->h_timings[2] = ms;
  return rc;
}

int smg_debug_histogram(smg_chain* ch, int* H, int* counts, int* mmax_out) {
  if (!ch) return fail(SMG_ERR_ARG, "NULL argument");
  SMG_CUDA(cudaSetDevice(ch->device));
  int rc = sync_status(ch);
  if (rc) return rc;
  if (launch_histogram(ch)) return SMG_ERR_CUDA;
  SMG_CUDA(cudaStreamSynchronize(ch->st));
  const int K = ch->h_K;
  if (mmax_out) *mmax_out = ch->mmax;
  if (H) {
    std::vector<int> h((size_t)K * ch->pp * ch->mmax);
    SMG_CUDA(d2h_sync(h.data(), ch->H, h.size() * 4, ch->st));
    for (int k = 0; k < K; k++)
      for (int j = 0; j < ch->p; j++)
        for (int a = 0; a < ch->mmax; a++)
          H[((size_t)k * ch->p + j) * ch->mmax + a] = h[((size_t)k * ch->pp + j) * ch->mmax + a];
  }
  if (counts) SMG_CUDA(d2h_sync(counts, ch->counts, (size_t)K * 4, ch->st));
  return 0;
}

int smg_debug_update_phi(smg_chain* ch, const double* u_center, const double* u_sigma) {
  if (!ch) return fail(SMG_ERR_ARG, "NULL argument");
  ch->ll_for_iter = -1;  // the state is about to change outside a sweep: the cached likelihood block is stale
  SMG_CUDA(cudaSetDevice(ch->device));
  int rc = sync_status(ch);
  if (rc) return rc;
  const int K = ch->h_K;
  const double *uc = nullptr, *us = nullptr;
  if (u_center) {
    SMG_CUDA(h2d_sync(ch->uc_d, u_center, (size_t)K * ch->p * 8, ch->st));
    uc = ch->uc_d;
  }
  if (u_sigma) {
    SMG_CUDA(h2d_sync(ch->us_d, u_sigma, (size_t)K * ch->p * 8, ch->st));
    us = ch->us_d;
  }
  rc = update_phi_all(ch, SUB_PHI_AFTER_SCAN, uc, us);
  if (rc) return rc;
  return sync_status(ch);
}

int smg_debug_loglik(smg_chain* ch, double* out) {
  if (!ch || !out) return fail(SMG_ERR_ARG, "NULL argument");
  SMG_CUDA(cudaSetDevice(ch->device));
  if (launch_loglik(ch)) return SMG_ERR_CUDA;
  int rc = sync_status(ch);
  *out = ch->h_loglik;
  return rc;
}

__global__ void dbg_hig_inv_kernel(int n, const double* om, const double* v, const double* w, const double* m,
                                   double* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = hig_inv_u_d(om[i], v[i], w[i], m[i]);
}
__global__ void dbg_loghig_kernel(int n, const double* s, const double* v, const double* w, const double* m,
                                  double* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = logdensity_hig_d(s[i], v[i], w[i], m[i]);
}

__global__ void dbg_rhig_kernel(int n, double v, double w, double m, RngKey key, double* out) {
  // first half: the lane-group sampler of the parameter updates; second half: the one-thread sampler of the pool
  const int t = blockIdx.x * blockDim.x + threadIdx.x, half = n / 2;
  const int i = t / PHI_G, g = t % PHI_G, lane = threadIdx.x & 31, gbase = lane & ~(PHI_G - 1);
  if (i < half) {  // whole groups
    const double u = hig_draw_u_grp(key, (uint32_t)i, 0u, v, w, m, g, ((1u << PHI_G) - 1u) << gbase, gbase);
    if (g == 0) out[i] = u;
  } else if (i < n && g == 0) {
    SubStream rs(key, U_SIGMA, (uint32_t)i, 0u);
    out[i] = hig_draw_u_d(rs, v, w, m);
  }
}

static int dbg_map4(int count, const double* a, const double* b, const double* c, const double* d, double* out, int which) {
  if (count <= 0) return 0;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(SMG_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  double* buf = nullptr;
  SMG_CUDA(cudaMalloc(&buf, (size_t)count * 5 * 8));
  const double* src[4] = {a, b, c, d};
  for (int q = 0; q < 4; q++) SMG_CUDA(cudaMemcpy(buf + (size_t)q * count, src[q], (size_t)count * 8, cudaMemcpyHostToDevice));
  if (which == 0)
    dbg_hig_inv_kernel<<<cdiv(count, 64), 64>>>(count, buf, buf + count, buf + 2 * (size_t)count, buf + 3 * (size_t)count,
                                                buf + 4 * (size_t)count);
  else
    dbg_loghig_kernel<<<cdiv(count, 64), 64>>>(count, buf, buf + count, buf + 2 * (size_t)count, buf + 3 * (size_t)count,
                                               buf + 4 * (size_t)count);
  SMG_CUDA(cudaGetLastError());
  SMG_CUDA(cudaDeviceSynchronize());
  SMG_CUDA(cudaMemcpy(out, buf + 4 * (size_t)count, (size_t)count * 8, cudaMemcpyDeviceToHost));
  cudaFree(buf);
  return 0;
}

int smg_debug_hig_inv_u(int count, const double* omega, const double* v, const double* w, const double* m, double* u_out) {
  return dbg_map4(count, omega, v, w, m, u_out, 0);
}
int smg_debug_logdensity_hig(int count, const double* sigma, const double* v, const double* w, const double* m,
                             double* out) {
  return dbg_map4(count, sigma, v, w, m, out, 1);
}

int smg_debug_rhig_u(int count, double v, double w, double m, unsigned long long seed, double* u_out) {
  if (count <= 0) return 0;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(SMG_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  double* buf = nullptr;
  {
    int dev = 0;
    SMG_CUDA(cudaGetDevice(&dev));
    if (zig_init(dev)) return SMG_ERR_CUDA;
  }
  SMG_CUDA(cudaMalloc(&buf, (size_t)count * 8));
  RngKey key;
  key.k0 = (uint32_t)seed;
  key.k1 = (uint32_t)(seed >> 32);
  key.sweep = 0;
  key.sub = 0;
  dbg_rhig_kernel<<<cdiv(count * PHI_G, 128), 128>>>(count, v, w, m, key, buf);
  SMG_CUDA(cudaGetLastError());
  SMG_CUDA(cudaDeviceSynchronize());
  SMG_CUDA(cudaMemcpy(u_out, buf, (size_t)count * 8, cudaMemcpyDeviceToHost));
  cudaFree(buf);
  return 0;
}

int smg_debug_split_merge(smg_chain* ch, const smg_sm_tape* tape, int* info, int* S, int* z_launch, int* z_star,
                          double* phi_out, double* terms) {
  if (!ch) return fail(SMG_ERR_ARG, "NULL argument");
  ch->ll_for_iter = -1;  // the state is about to change outside a sweep: the cached likelihood block is stale
  SMG_CUDA(cudaSetDevice(ch->device));
  int rc = sync_status(ch);
  if (rc) return rc;
  SMG_CUDA(cudaMemsetAsync(ch->accepted_d, 0, sizeof(int), ch->st));
  rc = sm_step(ch, tape);
  if (rc) return rc;
  rc = sync_status(ch);
  if (rc) return rc;
  return sm_readback(ch, info, S, z_launch, z_star, phi_out, terms);
}

// Pool-free auxiliary components (aux_mode 1) of the pass the chain would run next: column values of the first `count`
// (observation, component) pairs in row-major order and the parameters they were evaluated with.
int smg_debug_aux_free(smg_chain* ch, int count, double* llaux_out, double* centers_out, double* sigmas_out) {
  if (!ch || count < 1 || (long long)count > (long long)ch->n * ch->m_aux) return fail(SMG_ERR_ARG, "bad argument");
  if (ch->aux_mode != 1) return fail(SMG_ERR_ARG, "the chain was not created with aux_mode 1");
  SMG_CUDA(cudaSetDevice(ch->device));
  int rc = sync_status(ch);
  if (rc) return rc;
  SMG_CUDA(cudaStreamSynchronize(ch->st_aux));
  ch->aux_ready = false;
  const int buf = ch->aux_buf;
  if (launch_aux_ll(ch, nullptr, ch->st, ch->iter, buf)) return SMG_ERR_CUDA;
  if (llaux_out) SMG_CUDA(d2h_sync(llaux_out, ch->LLaux[buf], (size_t)count * 8, ch->st));
  if (centers_out || sigmas_out) {
    const int pp = ch->pp;
    uint8_t* dc = nullptr;
    double *ds = nullptr, *di = nullptr, *dd = nullptr;
    if (dalloc(&dc, (size_t)count * pp) || dalloc(&ds, (size_t)count * pp) || dalloc(&di, (size_t)count * pp) ||
        dalloc(&dd, (size_t)count * pp))
      return SMG_ERR_CUDA;
    RngKey key = mk_key(ch, SUB_AUX_FREE);
    pool_draw_kernel<<<cdiv((long long)count * pp, 128), 128, 0, ch->st>>>(count, pp, ch->p, ch->attr, ch->v, ch->w, key,
                                                                         ch->sigma_exact, dc, ds, di, dd);
    ch->h_launches++;
    SMG_CUDA(cudaGetLastError());
    std::vector<uint8_t> hc((size_t)count * pp);
    std::vector<double> hs((size_t)count * pp);
    SMG_CUDA(d2h_sync(hc.data(), dc, hc.size(), ch->st));
    SMG_CUDA(d2h_sync(hs.data(), ds, hs.size() * 8, ch->st));
    for (int e = 0; e < count; e++)
      for (int j = 0; j < ch->p; j++) {
        if (centers_out) centers_out[(size_t)e * ch->p + j] = hc[(size_t)e * pp + j];
        if (sigmas_out) sigmas_out[(size_t)e * ch->p + j] = hs[(size_t)e * pp + j];
      }
    void* ptrs[] = {dc, ds, di, dd};
    for (void* q : ptrs) cudaFree(q);
  }
  return 0;
}

// sample_initial_assignment (common_functions.cpp:174-183) with the uniforms injected: c_i = (int)(L u_i + 1) - 1
int smg_debug_initial_assignment(int n, int L, const double* u, int device, int* c_out) {
  if (!u || !c_out || n < 1 || L < 1) return fail(SMG_ERR_ARG, "bad argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
    return fail(SMG_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  SMG_CUDA(cudaSetDevice(device));
  double* du = nullptr;
  int* dc = nullptr;
  SMG_CUDA(cudaMalloc(&du, (size_t)n * 8));
  SMG_CUDA(cudaMalloc(&dc, (size_t)n * 4));
  SMG_CUDA(cudaMemcpy(du, u, (size_t)n * 8, cudaMemcpyHostToDevice));
  RngKey key{};
  init_assign_kernel<<<cdiv(n, 256), 256>>>(n, L, du, key, dc);
  SMG_CUDA(cudaGetLastError());
  SMG_CUDA(cudaMemcpy(c_out, dc, (size_t)n * 4, cudaMemcpyDeviceToHost));
  cudaFree(du);
  cudaFree(dc);
  return 0;
}

// The MH ratio alone (split_merge.cpp:393-540) on an injected launch / proposal: no scan and no parameter draw runs, so
// every addend depends only on the device's arithmetic of logprobgs_phi, logprobgs_c_i, loglikelihood_hamming, priors
// and logdensity_hig -- the tests hold them to 1e-12 against the oracle's on the same (centres, sigmas, sides).
int smg_debug_sm_terms(smg_chain* ch, const double* u_pair, const int* z_launch, const int* z_star, const double* phi6,
                       double u_accept, int* info, double* terms) {
  if (!ch || !u_pair || !z_launch || !z_star || !phi6) return fail(SMG_ERR_ARG, "NULL argument");
  SMG_CUDA(cudaSetDevice(ch->device));
  int rc = sync_status(ch);
  if (rc) return rc;
  SmWork* W = ch->sm;
  if (!W) return fail(SMG_ERR_ARG, "chain has no split-merge workspace");
  const int n = ch->n, p = ch->p, pp = ch->pp, B = ch->NS, cur = ch->cur;
  smg_sm_tape tp;
  memset(&tp, 0, sizeof(tp));
  tp.u_pair = u_pair;
  tp.u_accept = &u_accept;
  rc = sm_inject(ch, &tp);
  if (rc) return rc;
  SMG_CUDA(cudaMemsetAsync(ch->accepted_d, 0, sizeof(int), ch->st));
  sm_select_kernel<<<1, 1024, 0, ch->st>>>(n, ch->c, ch->K, W->u_pair, mk_key(ch, SUB_SM_SELECT), B, W->S, W->zState, W->info,
                                           W->plan, W->cnt, W->terms, ch->pair_det);
  ch->h_launches++;
  SMG_CUDA(cudaGetLastError());
  SmInfo I;
  SMG_CUDA(d2h_sync(&I, W->info, sizeof(I), ch->st));
  const size_t len = (size_t)pp * ch->mmax;
  if (sm_hist(ch, W->zState, SH_S0, false, nullptr)) return SMG_ERR_CUDA;
  sm_hist_add_kernel<<<sm_cdiv(len, 256), 256, 0, ch->st>>>((int)len, W->H, W->cnt, SH_S0, SH_S1, SH_M);
  ch->h_launches++;
  if (I.nS > 0) {
    SMG_CUDA(h2d_sync(W->zL, z_launch, (size_t)I.nS * 4, ch->st));
    SMG_CUDA(h2d_sync(W->zStar, z_star, (size_t)I.nS * 4, ch->st));
  }
  {  // the six parameter slots: split-launch A, B, merge-launch M, proposal A*, B*, M* (contiguous after the state slots)
    std::vector<uint8_t> hc((size_t)6 * pp, 0);
    std::vector<double> hs((size_t)6 * pp, 1.0);
    for (int q = 0; q < 6; q++)
      for (int j = 0; j < p; j++) {
        hc[(size_t)q * pp + j] = (uint8_t)phi6[((size_t)q * 2 + 0) * p + j];
        hs[(size_t)q * pp + j] = phi6[((size_t)q * 2 + 1) * p + j];
      }
    const size_t o = (size_t)(B + SM_SL_A) * pp;
    SMG_CUDA(h2d_sync(ch->cen[cur] + o, hc.data(), hc.size(), ch->st));
    SMG_CUDA(h2d_sync(ch->sig[cur] + o, hs.data(), hs.size() * 8, ch->st));
    derive_terms_kernel<<<6, 256, 0, ch->st>>>(6, pp, p, ch->attr, ch->sig[cur] + o, ch->isg[cur] + o, ch->den + o,
                                               ch->sden[cur] + B + SM_SL_A);
    ch->h_launches++;
    SMG_CUDA(cudaGetLastError());
  }
  if (sm_hist(ch, W->zL, SH_L0, false, nullptr)) return SMG_ERR_CUDA;
  if (sm_hist(ch, W->zStar, SH_P0, false, nullptr)) return SMG_ERR_CUDA;
  sm_gsphi_prior_kernel<<<6, 256, 0, ch->st>>>(pp, p, ch->mmax, ch->attr, ch->v, ch->w, W->H, W->cnt, W->plan, ch->cen[cur],
                                               ch->sig[cur], W->terms);
  sm_rowterms_kernel<<<sm_cdiv((long long)(n + 2) * 32, 256), 256, 0, ch->st>>>(
      ch->X, pp, W->S, W->info, W->plan, W->zL, W->zStar, W->zState, W->cnt, ch->cen[cur], ch->isg[cur], ch->sden[cur],
      W->rowvals, n + 2);
  sm_rowreduce1_kernel<<<dim3(SM_RB, 4), 256, 0, ch->st>>>(W->info, W->rowvals, n + 2, W->partial);
  sm_accept_kernel<<<1, 256, 0, ch->st>>>(W->info, W->plan, W->cnt, W->partial, ch->gamma, W->u_accept,
                                          mk_key(ch, SUB_SM_ACCEPT), W->terms, ch->accepted_d, ch->stats_d);
  ch->h_launches += 4;
  SMG_CUDA(cudaGetLastError());
  rc = sm_readback(ch, info, nullptr, nullptr, nullptr, nullptr, terms);
  SMG_CUDA(cudaMemsetAsync(ch->accepted_d, 0, sizeof(int), ch->st));  // nothing was applied
  return rc;
}


// ------------------------------------------------------------------------------------------
// posterior similarity matrix (smg_psm.cuh)
// ------------------------------------------------------------------------------------------
int smg_psm_create(int n, int device, int capacity_sweeps, void* external_psm_int32, smg_psm** out) {
  if (!out) return fail(SMG_ERR_ARG, "out is NULL");
  *out = nullptr;
  if (n < 1 || capacity_sweeps < 1) return fail(SMG_ERR_ARG, "n and capacity_sweeps must be positive");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(SMG_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  if (device < 0 || device >= ndev) return fail(SMG_ERR_ARG, "bad device ordinal");
  SMG_CUDA(cudaSetDevice(device));
  smg_psm* P = new smg_psm();
  P->n = n;
  P->device = device;
  P->cap = capacity_sweeps;
  if (cudaStreamCreateWithFlags(&P->st, cudaStreamNonBlocking) != cudaSuccess) {
    delete P;
    return fail(SMG_ERR_CUDA, "cudaStreamCreate failed");
  }
  if (external_psm_int32) {
    P->psm = (int*)external_psm_int32;
  } else {
    if (cudaMalloc(&P->psm, (size_t)n * n * sizeof(int)) != cudaSuccess) {
      cudaStreamDestroy(P->st);
      delete P;
      return fail(SMG_ERR_CUDA, "cudaMalloc of the n x n int32 matrix failed");
    }
    P->owns = true;
    // cleared on the stream the accumulate kernels run on: ordered before the first flush
    cudaMemsetAsync(P->psm, 0, (size_t)n * n * sizeof(int), P->st);
  }
  if (cudaMalloc(&P->labels, (size_t)capacity_sweeps * n) != cudaSuccess ||
      cudaEventCreate(&P->ev[0]) != cudaSuccess || cudaEventCreate(&P->ev[1]) != cudaSuccess ||
      cudaStreamSynchronize(P->st) != cudaSuccess) {
    smg_psm_destroy(P);
    return fail(SMG_ERR_CUDA, "allocation of the label buffer failed");
  }
  *out = P;
  return 0;
}

void smg_psm_destroy(smg_psm* P) {
  if (!P) return;
  cudaSetDevice(P->device);
  if (P->st) cudaStreamSynchronize(P->st);
  if (P->peers_d) {  // distributed: unmap the peers' matrices, and free the own one only when nobody maps it any more
    for (void* q : P->peers_opened) cudaIpcCloseMemHandle(q);
    cudaFree(P->peers_d);
    comm_rendezvous(P->dist);
  }
  if (P->owns && P->psm) cudaFree(P->psm);
  if (P->labels) cudaFree(P->labels);
  for (int q = 0; q < 2; q++)
    if (P->ev[q]) cudaEventDestroy(P->ev[q]);
  if (P->st) cudaStreamDestroy(P->st);
  delete P;
}

int smg_psm_flush(smg_psm* P) {
  if (!P) return fail(SMG_ERR_ARG, "psm is NULL");
  SMG_CUDA(cudaSetDevice(P->device));
  if (P->count == 0) return 0;
  cudaEventRecord(P->ev[0], P->st);
  int rc;
  if (P->kmax <= 64)
    rc = psm_launch<64, 4>(P);
  else if (P->kmax <= 128)
    rc = psm_launch<128, 3>(P);
  else
    rc = psm_launch<256, 2>(P);
  if (rc) return rc;
  cudaEventRecord(P->ev[1], P->st);
  SMG_CUDA(cudaStreamSynchronize(P->st));
  float ms = 0;
  cudaEventElapsedTime(&ms, P->ev[0], P->ev[1]);
  P->last_ms = ms;
  P->launches++;
  P->total += P->count;
  P->count = 0;
  P->kmax = 0;
  return 0;
}

// flush + lower triangle <- upper triangle: the matrix as every reader expects it
int smg_psm_finalize(smg_psm* P) {
  int rc = smg_psm_flush(P);
  if (rc) return rc;
  if (P->peers_d) {
    // Distributed matrix: every rank's flushes have added into the owners' memories (a flush returns when its kernel has
    // completed); once all ranks are here the upper triangle is final, each rank copies it into the lower triangle of
    // its own rows (reading the source rows from their owners), and nobody adds again before everybody has read.
    // COLLECTIVE: every rank of the communicator calls this.
    rc = comm_rendezvous(P->dist);
    if (rc) return rc;
    const int nb = cdiv(P->n, 32);
    psm_mirror_dist_kernel<<<dim3(nb, nb), 256, 0, P->st>>>(P->peers_d, P->n, P->rows_per, P->dist->rank);
    SMG_CUDA(cudaGetLastError());
    SMG_CUDA(cudaStreamSynchronize(P->st));
    P->launches++;
    P->lower_stale = false;
    return comm_rendezvous(P->dist);
  }
  if (P->lower_stale) {
    const int nb = cdiv(P->n, 32);
    psm_mirror_kernel<<<dim3(nb, nb), 256, 0, P->st>>>(P->psm, P->n);
    SMG_CUDA(cudaGetLastError());
    SMG_CUDA(cudaStreamSynchronize(P->st));
    P->launches++;
    P->lower_stale = false;
  }
  return 0;
}

int smg_psm_push_host(smg_psm* P, const int* c_i) {
  if (!P || !c_i) return fail(SMG_ERR_ARG, "NULL argument");
  SMG_CUDA(cudaSetDevice(P->device));
  if (P->count == P->cap) {
    int rc = smg_psm_flush(P);
    if (rc) return rc;
  }
  std::vector<uint8_t> h(P->n);
  int mx = 0;
  for (int i = 0; i < P->n; i++) {
    if (c_i[i] < 0 || c_i[i] > 255) return fail(SMG_ERR_ARG, "labels must lie in 0..255");
    h[i] = (uint8_t)c_i[i];
    mx = std::max(mx, c_i[i]);
  }
  SMG_CUDA(cudaMemcpyAsync(P->labels + (size_t)P->count * P->n, h.data(), P->n, cudaMemcpyHostToDevice, P->st));
  SMG_CUDA(cudaStreamSynchronize(P->st));
  P->kmax = std::max(P->kmax, mx + 1);
  P->count++;
  return 0;
}

int smg_psm_push_chain(smg_psm* P, smg_chain* ch) {
  if (!P || !ch) return fail(SMG_ERR_ARG, "NULL argument");
  if (ch->n != P->n || ch->device != P->device) return fail(SMG_ERR_ARG, "chain and matrix differ in n or device");
  SMG_CUDA(cudaSetDevice(P->device));
  if (P->count == P->cap) {
    int rc = smg_psm_flush(P);
    if (rc) return rc;
  }
  // the chain's allocation is final once its stream is idle (smg_step ends with a synchronisation)
  SMG_CUDA(cudaStreamSynchronize(ch->st));
  psm_pack_labels_kernel<<<cdiv(P->n, 256), 256, 0, P->st>>>(ch->c, P->n, P->labels + (size_t)P->count * P->n);
  SMG_CUDA(cudaGetLastError());
  SMG_CUDA(cudaStreamSynchronize(P->st));
  P->kmax = std::max(P->kmax, ch->h_K);
  P->count++;
  return 0;
}

int smg_psm_read(smg_psm* P, int row0, int nrows, int* out) {
  if (!P || !out) return fail(SMG_ERR_ARG, "NULL argument");
  if (row0 < 0 || nrows < 0 || row0 + nrows > P->n) return fail(SMG_ERR_ARG, "row range out of bounds");
  int rc = smg_psm_finalize(P);
  if (rc) return rc;
  SMG_CUDA(cudaMemcpy(out, P->psm + (size_t)row0 * P->n, (size_t)nrows * P->n * sizeof(int), cudaMemcpyDeviceToHost));
  return 0;
}

int smg_psm_info(smg_psm* P, long long* sweeps, double* last_flush_ms, unsigned long long* launches) {
  if (!P) return fail(SMG_ERR_ARG, "psm is NULL");
  if (sweeps) *sweeps = P->total + P->count;
  if (last_flush_ms) *last_flush_ms = P->last_ms;
  if (launches) *launches = P->launches;
  return 0;
}

// CUDA-core evaluation of the same buffered sweeps into `out_psm_host` (n x n, overwritten): device-side
// cross-check of the tensor-core kernel; does not consume the buffer
int smg_debug_psm_reference(smg_psm* P, int* out_psm_host) {
  if (!P || !out_psm_host) return fail(SMG_ERR_ARG, "NULL argument");
  SMG_CUDA(cudaSetDevice(P->device));
  int* d = nullptr;
  SMG_CUDA(cudaMalloc(&d, (size_t)P->n * P->n * sizeof(int)));
  SMG_CUDA(cudaMemsetAsync(d, 0, (size_t)P->n * P->n * sizeof(int), P->st));
  psm_reference_kernel<<<dim3(cdiv(P->n, 256), P->n), 256, 0, P->st>>>(P->labels, P->n, P->count, d);
  SMG_CUDA(cudaGetLastError());
  SMG_CUDA(cudaStreamSynchronize(P->st));
  SMG_CUDA(cudaMemcpy(out_psm_host, d, (size_t)P->n * P->n * sizeof(int), cudaMemcpyDeviceToHost));
  cudaFree(d);
  return 0;
}


// ------------------------------------------------------------------------------------------
// reductions over GPUs (smg_comm.cuh): NCCL over NVLink, called from this library
// ------------------------------------------------------------------------------------------
int smg_comm_unique_id(char* id128) {
  if (!id128) return fail(SMG_ERR_ARG, "NULL argument");
  NcclApi* N = nccl_api();
  if (!N->err.empty()) return fail(SMG_ERR_CUDA, N->err);
  ncclUniqueId id;
  SMG_NCCL(N->GetUniqueId(&id));
  memcpy(id128, id.internal, NCCL_UNIQUE_ID_BYTES);
  return 0;
}

int smg_comm_create(int rank, int world, const char* id128, int device, smg_comm** out) {
  if (!out) return fail(SMG_ERR_ARG, "out is NULL");
  *out = nullptr;
  if (world < 1 || rank < 0 || rank >= world) return fail(SMG_ERR_ARG, "bad rank / world size");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(SMG_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  if (device < 0 || device >= ndev) return fail(SMG_ERR_ARG, "bad device ordinal");
  SMG_CUDA(cudaSetDevice(device));
  smg_comm* C = new smg_comm();
  C->rank = rank;
  C->world = world;
  C->device = device;
  if (cudaStreamCreateWithFlags(&C->st, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&C->ev[0]) != cudaSuccess ||
      cudaEventCreate(&C->ev[1]) != cudaSuccess) {
    smg_comm_destroy(C);
    return fail(SMG_ERR_CUDA, "stream / event creation failed");
  }
  if (world > 1) {
    if (!id128) {
      smg_comm_destroy(C);
      return fail(SMG_ERR_ARG, "the NCCL unique id of rank 0 is required when world > 1");
    }
    NcclApi* N = nccl_api();
    if (!N->err.empty()) {
      smg_comm_destroy(C);
      return fail(SMG_ERR_CUDA, N->err);
    }
    ncclUniqueId id;
    memcpy(id.internal, id128, NCCL_UNIQUE_ID_BYTES);
    ncclResult_t r = N->CommInitRank(&C->comm, world, id, rank);
    if (r != ncclSuccess) {
      std::string msg = std::string("ncclCommInitRank: ") + N->GetErrorString(r);
      smg_comm_destroy(C);
      return fail(SMG_ERR_CUDA, msg);
    }
    // NCCL connects its channels lazily, at the first collective of each kind (~1 s measured on 2 B200): pay for that
    // here, not inside the first timed reduction
    int* w = nullptr;
    SMG_CUDA(cudaMalloc(&w, (size_t)world * 256 * sizeof(int)));
    SMG_CUDA(cudaMemsetAsync(w, 0, (size_t)world * 256 * sizeof(int), C->st));
    SMG_NCCL(N->AllReduce(w, w, 256, ncclInt32, ncclSum, C->comm, C->st));
    SMG_NCCL(N->ReduceScatter(w, w + (size_t)rank * 256, 256, ncclInt32, ncclSum, C->comm, C->st));
    SMG_NCCL(N->AllGather(w + (size_t)rank * 256, w, 256, ncclInt32, C->comm, C->st));
    SMG_CUDA(cudaStreamSynchronize(C->st));
    C->scratch = w;
  }
  *out = C;
  return 0;
}

void smg_comm_destroy(smg_comm* C) {
  if (!C) return;
  cudaSetDevice(C->device);
  if (C->st) cudaStreamSynchronize(C->st);
  if (C->comm) nccl_api()->CommDestroy(C->comm);
  for (int q = 0; q < 2; q++)
    if (C->ev[q]) cudaEventDestroy(C->ev[q]);
  if (C->scratch) cudaFree(C->scratch);
  if (C->st) cudaStreamDestroy(C->st);
  delete C;
}

// PSM counts of all ranks summed.  mode 1 (reduce-scatter): rank g ends up with the reduced rows [row0, row0 + nrows)
// in place (its other rows keep the local partial sums); mode 0 or n % world != 0: all-reduce, every rank holds the whole
// matrix.  *ms = device time of the collective, *bus_gbs = NCCL bus bandwidth (bytes * (G-1)/G [* 2 for all-reduce] / s).
int smg_chains_reduce_psm(smg_comm* C, smg_psm* P, int mode, int* row0, int* nrows, double* ms, double* bus_gbs) {
  if (!C || !P) return fail(SMG_ERR_ARG, "NULL argument");
  if (C->device != P->device) return fail(SMG_ERR_ARG, "communicator and matrix live on different devices");
  SMG_CUDA(cudaSetDevice(C->device));
  int rc = smg_psm_finalize(P);
  if (rc) return rc;
  if (P->peers_d) {  // already reduced and scattered by the accumulation kernel itself
    if (row0) *row0 = C->rank * P->rows_per;
    if (nrows) *nrows = P->rows_per;
    if (ms) *ms = 0.0;
    if (bus_gbs) *bus_gbs = 0.0;
    return 0;
  }
  const int n = P->n, G = C->world;
  const bool scatter = mode == 1 && G > 1 && n % G == 0;
  int r0 = 0, nr = n;
  float fms = 0.f;
  double bus = 0.0;
  if (G > 1) {
    NcclApi* N = nccl_api();
    const size_t total = (size_t)n * n;
    // the ranks' chains end at different times: meet first, so that *ms is the collective and not the wait for the
    // slowest rank
    SMG_NCCL(N->AllReduce(C->scratch, C->scratch, 1, ncclInt32, ncclSum, C->comm, C->st));
    SMG_CUDA(cudaEventRecord(C->ev[0], C->st));
    if (scatter) {
      const size_t per = total / G;
      SMG_NCCL(N->ReduceScatter(P->psm, P->psm + (size_t)C->rank * per, per, ncclInt32, ncclSum, C->comm, C->st));
      nr = n / G;
      r0 = C->rank * nr;
    } else {
      SMG_NCCL(N->AllReduce(P->psm, P->psm, total, ncclInt32, ncclSum, C->comm, C->st));
    }
    SMG_CUDA(cudaEventRecord(C->ev[1], C->st));
    SMG_CUDA(cudaStreamSynchronize(C->st));
    cudaEventElapsedTime(&fms, C->ev[0], C->ev[1]);
    const double bytes = (double)total * 4.0 * (double)(G - 1) / (double)G * (scatter ? 1.0 : 2.0);
    bus = fms > 0.f ? bytes / (fms * 1e-3) / 1e9 : 0.0;
  }
  if (row0) *row0 = r0;
  if (nrows) *nrows = nr;
  if (ms) *ms = fms;
  if (bus_gbs) *bus_gbs = bus;
  return 0;
}

// split-R-hat (BDA3 11.4) of a scalar trace over all chains of all ranks: every chain is cut in two halves, the
// (count, mean, M2) moments of the halves are all-gathered, W = mean within-half variance, B = n * variance of the
// half means, R-hat = sqrt(((n-1)/n W + B/n) / W).  traces: [nchains_local][T] on the host.
// Deal the rows of a fresh matrix to the ranks and map every rank's matrix into every other rank's address space (CUDA IPC;
// over NVLink between the GPUs of one box): from now on a flush adds its tiles into the owners' memories.
int smg_chains_psm_distribute(smg_comm* C, smg_psm* P) {
  if (!C || !P) return fail(SMG_ERR_ARG, "NULL argument");
  if (C->device != P->device) return fail(SMG_ERR_ARG, "communicator and matrix live on different devices");
  if (!P->owns) return fail(SMG_ERR_ARG, "an external matrix cannot be distributed (it must be the library's own allocation)");
  if (P->peers_d) return fail(SMG_ERR_ARG, "the matrix is already distributed");
  if (P->total != 0 || P->count != 0) return fail(SMG_ERR_ARG, "distribute the matrix before the first allocation is pushed");
  const int G = C->world, n = P->n;
  if (n % G) return fail(SMG_ERR_ARG, "n must be a multiple of the number of ranks");
  SMG_CUDA(cudaSetDevice(C->device));
  SMG_CUDA(cudaStreamSynchronize(P->st));  // the matrix has been cleared
  std::vector<int*> ptrs(G, nullptr);
  ptrs[C->rank] = P->psm;
  if (G > 1) {
    NcclApi* N = nccl_api();
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t mine;
    SMG_CUDA(cudaIpcGetMemHandle(&mine, P->psm));
    std::vector<cudaIpcMemHandle_t> all(G);
    unsigned char* buf = reinterpret_cast<unsigned char*>(C->scratch);  // world * 1024 bytes
    SMG_CUDA(cudaMemcpyAsync(buf + (size_t)C->rank * 64, &mine, 64, cudaMemcpyHostToDevice, C->st));
    SMG_NCCL(N->AllGather(buf + (size_t)C->rank * 64, buf, 64, ncclUint8, C->comm, C->st));
    SMG_CUDA(cudaMemcpyAsync(all.data(), buf, (size_t)G * 64, cudaMemcpyDeviceToHost, C->st));
    SMG_CUDA(cudaStreamSynchronize(C->st));
    for (int g = 0; g < G; g++) {
      if (g == C->rank) continue;
      void* q = nullptr;
      if (cudaIpcOpenMemHandle(&q, all[g], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        const std::string msg = std::string("cudaIpcOpenMemHandle of rank ") + std::to_string(g) + ": " + cudaGetErrorString(cudaGetLastError());
        for (void* o : P->peers_opened) cudaIpcCloseMemHandle(o);
        P->peers_opened.clear();
        return fail(SMG_ERR_CUDA, msg);
      }
      P->peers_opened.push_back(q);
      ptrs[g] = reinterpret_cast<int*>(q);
    }
  }
  SMG_CUDA(cudaMalloc(&P->peers_d, (size_t)G * sizeof(int*)));
  SMG_CUDA(cudaMemcpy(P->peers_d, ptrs.data(), (size_t)G * sizeof(int*), cudaMemcpyHostToDevice));
  P->rows_per = n / G;
  P->dist = C;
  return comm_rendezvous(C);  // nobody adds into a matrix that is not mapped everywhere yet
}

int smg_chains_split_rhat(smg_comm* C, const double* traces, int nchains_local, int T, double* rhat, long long* nchains_total) {
  if (!C || !rhat || (nchains_local > 0 && !traces)) return fail(SMG_ERR_ARG, "NULL argument");
  if (nchains_local < 0 || T < 4) return fail(SMG_ERR_ARG, "need T >= 4 draws per chain");
  SMG_CUDA(cudaSetDevice(C->device));
  const int G = C->world, half = T / 2;
  // local moments: per chain 2 x (mean, M2)
  std::vector<double> mom((size_t)std::max(nchains_local, 1) * 4, 0.0);
  for (int q = 0; q < nchains_local; q++)
    for (int h = 0; h < 2; h++) {
      const double* x = traces + (size_t)q * T + (h ? T - half : 0);
      double m = 0.0;
      for (int i = 0; i < half; i++) m += x[i];
      m /= half;
      double s = 0.0;
      for (int i = 0; i < half; i++) s += (x[i] - m) * (x[i] - m);
      mom[(size_t)q * 4 + 2 * h] = m;
      mom[(size_t)q * 4 + 2 * h + 1] = s;
    }
  std::vector<double> all;
  std::vector<int> cnt(G, nchains_local);
  if (G > 1) {
    NcclApi* N = nccl_api();
    int *d_cnt = nullptr, maxc = 0;
    SMG_CUDA(cudaMalloc(&d_cnt, (size_t)G * sizeof(int)));
    SMG_CUDA(cudaMemcpyAsync(d_cnt + C->rank, &nchains_local, sizeof(int), cudaMemcpyHostToDevice, C->st));
    SMG_NCCL(N->AllGather(d_cnt + C->rank, d_cnt, 1, ncclInt32, C->comm, C->st));
    SMG_CUDA(cudaMemcpyAsync(cnt.data(), d_cnt, (size_t)G * sizeof(int), cudaMemcpyDeviceToHost, C->st));
    SMG_CUDA(cudaStreamSynchronize(C->st));
    cudaFree(d_cnt);
    for (int g = 0; g < G; g++) maxc = std::max(maxc, cnt[g]);
    if (maxc == 0) return fail(SMG_ERR_ARG, "no chains");
    double* d_mom = nullptr;
    const size_t per = (size_t)maxc * 4;
    SMG_CUDA(cudaMalloc(&d_mom, per * G * sizeof(double)));
    SMG_CUDA(cudaMemsetAsync(d_mom, 0, per * G * sizeof(double), C->st));
    SMG_CUDA(cudaMemcpyAsync(d_mom + per * C->rank, mom.data(), (size_t)nchains_local * 4 * sizeof(double), cudaMemcpyHostToDevice,
                             C->st));
    SMG_NCCL(N->AllGather(d_mom + per * C->rank, d_mom, per, ncclDouble, C->comm, C->st));
    std::vector<double> buf(per * G);
    SMG_CUDA(cudaMemcpyAsync(buf.data(), d_mom, per * G * sizeof(double), cudaMemcpyDeviceToHost, C->st));
    SMG_CUDA(cudaStreamSynchronize(C->st));
    cudaFree(d_mom);
    for (int g = 0; g < G; g++) all.insert(all.end(), buf.begin() + per * g, buf.begin() + per * g + (size_t)cnt[g] * 4);
  } else {
    all.assign(mom.begin(), mom.begin() + (size_t)nchains_local * 4);
  }
  const size_t nparts = all.size() / 2;  // half-chains
  if (nchains_total) *nchains_total = (long long)(nparts / 2);
  if (nparts < 2) return fail(SMG_ERR_ARG, "no chains");
  double W = 0.0, mbar = 0.0;
  for (size_t q = 0; q < nparts; q++) {
    W += all[2 * q + 1] / (half - 1);
    mbar += all[2 * q];
  }
  W /= (double)nparts;
  mbar /= (double)nparts;
  double B = 0.0;
  for (size_t q = 0; q < nparts; q++) B += (all[2 * q] - mbar) * (all[2 * q] - mbar);
  B = (double)half * B / (double)(nparts - 1);
  if (W == 0.0)
    *rhat = (B == 0.0) ? 1.0 : INFINITY;
  else
    *rhat = std::sqrt((((double)half - 1.0) / half * W + B / half) / W);
  return 0;
}

// histogram of the number of clusters over all ranks: hist[k], k = 0..kmax; draws with K > kmax are counted in
// *overflow (so that sum(hist) + overflow = total draws)
int smg_chains_k_histogram(smg_comm* C, const int* K, long long count, int kmax, long long* hist, long long* overflow) {
  if (!C || !hist || kmax < 0 || (count > 0 && !K)) return fail(SMG_ERR_ARG, "bad argument");
  SMG_CUDA(cudaSetDevice(C->device));
  std::vector<long long> h((size_t)kmax + 2, 0);
  for (long long i = 0; i < count; i++) {
    const int k = K[i];
    if (k < 0) return fail(SMG_ERR_ARG, "negative K");
    h[k <= kmax ? k : kmax + 1]++;
  }
  if (C->world > 1) {
    NcclApi* N = nccl_api();
    long long* d = nullptr;
    SMG_CUDA(cudaMalloc(&d, h.size() * sizeof(long long)));
    SMG_CUDA(cudaMemcpyAsync(d, h.data(), h.size() * sizeof(long long), cudaMemcpyHostToDevice, C->st));
    SMG_NCCL(N->AllReduce(d, d, h.size(), ncclInt64, ncclSum, C->comm, C->st));
    SMG_CUDA(cudaMemcpyAsync(h.data(), d, h.size() * sizeof(long long), cudaMemcpyDeviceToHost, C->st));
    SMG_CUDA(cudaStreamSynchronize(C->st));
    cudaFree(d);
  }
  for (int k = 0; k <= kmax; k++) hist[k] = h[k];
  if (overflow) *overflow = h[(size_t)kmax + 1];
  return 0;
}

// ------------------------------------------------------------------------------------------
// posterior summaries on the device (smg_post.cuh): point estimate from the PSM, adjusted Rand index, IAT / ESS
// ------------------------------------------------------------------------------------------
int smg_psm_point_estimate(smg_psm* P, int row0, int nrows, const int* candidates, int ncand, long long draws,
                           long long* binder_scaled, double* vi_rowsum) {
  if (!P || !candidates || ncand < 1 || draws < 1 || !binder_scaled || !vi_rowsum) return fail(SMG_ERR_ARG, "bad argument");
  if (row0 < 0 || nrows < 0 || row0 + nrows > P->n) return fail(SMG_ERR_ARG, "row range out of bounds");
  int rc = smg_psm_finalize(P);
  if (rc) return rc;
  SMG_CUDA(cudaSetDevice(P->device));
  const int n = P->n;
  std::vector<uint8_t> hc((size_t)ncand * n);
  for (size_t e = 0; e < hc.size(); e++) {
    if (candidates[e] < 0 || candidates[e] > 255) return fail(SMG_ERR_ARG, "candidate labels must lie in 0..255");
    hc[e] = (uint8_t)candidates[e];
  }
  uint8_t* dc = nullptr;
  unsigned long long* db = nullptr;
  long long* drs = nullptr;
  double* dv = nullptr;
  SMG_CUDA(cudaMalloc(&dc, hc.size()));
  SMG_CUDA(cudaMalloc(&db, (size_t)ncand * 8));
  SMG_CUDA(cudaMalloc(&drs, (size_t)ncand * std::max(nrows, 1) * 3 * 8));
  SMG_CUDA(cudaMalloc(&dv, (size_t)ncand * 8));
  SMG_CUDA(cudaMemcpyAsync(dc, hc.data(), hc.size(), cudaMemcpyHostToDevice, P->st));
  SMG_CUDA(cudaMemsetAsync(db, 0, (size_t)ncand * 8, P->st));
  SMG_CUDA(cudaMemsetAsync(dv, 0, (size_t)ncand * 8, P->st));
  if (nrows > 0) {
    for (int q0 = 0; q0 < ncand; q0 += POST_MAXC) {
      psm_point_estimate_kernel<<<cdiv((long long)nrows * 32, 256), 256, 0, P->st>>>(P->psm + (size_t)row0 * n, n, row0, nrows, dc, ncand, q0,
                                                                                draws, db, drs);
      P->launches++;
    }
    psm_vi_reduce_kernel<<<ncand, 256, 0, P->st>>>(drs, nrows, draws, dv);
    P->launches++;
    SMG_CUDA(cudaGetLastError());
  }
  SMG_CUDA(cudaMemcpyAsync(binder_scaled, db, (size_t)ncand * 8, cudaMemcpyDeviceToHost, P->st));
  SMG_CUDA(cudaMemcpyAsync(vi_rowsum, dv, (size_t)ncand * 8, cudaMemcpyDeviceToHost, P->st));
  SMG_CUDA(cudaStreamSynchronize(P->st));
  void* ptrs[] = {dc, db, drs, dv};
  for (void* q : ptrs) cudaFree(q);
  return 0;
}

int smg_chains_point_estimate(smg_comm* C, smg_psm* P, int row0, int nrows, const int* candidates, int ncand, long long draws,
                              double* binder, double* vi_lower_bound, int* best_binder, int* best_vi) {
  if (!C || !P || !binder || !vi_lower_bound) return fail(SMG_ERR_ARG, "NULL argument");
  std::vector<long long> bs(ncand > 0 ? ncand : 1);
  std::vector<double> vs(ncand > 0 ? ncand : 1);
  int rc = smg_psm_point_estimate(P, row0, nrows, candidates, ncand, draws, bs.data(), vs.data());
  if (rc) return rc;
  if (C->world > 1) {  // Binder and VI are sums over rows: add the ranks' row blocks
    SMG_CUDA(cudaSetDevice(C->device));
    NcclApi* N = nccl_api();
    long long* d = nullptr;
    double* dd = nullptr;
    SMG_CUDA(cudaMalloc(&d, (size_t)ncand * 8));
    SMG_CUDA(cudaMalloc(&dd, (size_t)ncand * 8));
    SMG_CUDA(cudaMemcpyAsync(d, bs.data(), (size_t)ncand * 8, cudaMemcpyHostToDevice, C->st));
    SMG_CUDA(cudaMemcpyAsync(dd, vs.data(), (size_t)ncand * 8, cudaMemcpyHostToDevice, C->st));
    SMG_NCCL(N->AllReduce(d, d, (size_t)ncand, ncclInt64, ncclSum, C->comm, C->st));
    SMG_NCCL(N->AllReduce(dd, dd, (size_t)ncand, ncclDouble, ncclSum, C->comm, C->st));
    SMG_CUDA(cudaMemcpyAsync(bs.data(), d, (size_t)ncand * 8, cudaMemcpyDeviceToHost, C->st));
    SMG_CUDA(cudaMemcpyAsync(vs.data(), dd, (size_t)ncand * 8, cudaMemcpyDeviceToHost, C->st));
    SMG_CUDA(cudaStreamSynchronize(C->st));
    cudaFree(d);
    cudaFree(dd);
  }
  int bb = 0, bv = 0;
  for (int q = 0; q < ncand; q++) {
    binder[q] = (double)bs[q] / (double)draws;
    vi_lower_bound[q] = vs[q] / (double)P->n;
    if (bs[q] < bs[bb]) bb = q;
    if (vi_lower_bound[q] < vi_lower_bound[bv]) bv = q;
  }
  if (best_binder) *best_binder = bb;
  if (best_vi) *best_vi = bv;
  return 0;
}

int smg_adjusted_rand_index(const int* a, const int* b, int n, int device, double* ari) {
  if (!a || !b || !ari || n < 2) return fail(SMG_ERR_ARG, "bad argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
    return fail(SMG_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  SMG_CUDA(cudaSetDevice(device));
  int ka = 0, kb = 0;
  for (int i = 0; i < n; i++) {
    if (a[i] < 0 || b[i] < 0) return fail(SMG_ERR_ARG, "labels must be non-negative");
    ka = std::max(ka, a[i] + 1);
    kb = std::max(kb, b[i] + 1);
  }
  if ((long long)ka * kb > (1ll << 26)) return fail(SMG_ERR_ARG, "contingency table too large");
  int *da = nullptr, *db = nullptr, *tab = nullptr;
  unsigned long long* out = nullptr;
  SMG_CUDA(cudaMalloc(&da, (size_t)n * 4));
  SMG_CUDA(cudaMalloc(&db, (size_t)n * 4));
  SMG_CUDA(cudaMalloc(&tab, (size_t)ka * kb * 4));
  SMG_CUDA(cudaMalloc(&out, 3 * 8));
  SMG_CUDA(cudaMemcpy(da, a, (size_t)n * 4, cudaMemcpyHostToDevice));
  SMG_CUDA(cudaMemcpy(db, b, (size_t)n * 4, cudaMemcpyHostToDevice));
  SMG_CUDA(cudaMemset(tab, 0, (size_t)ka * kb * 4));
  SMG_CUDA(cudaMemset(out, 0, 3 * 8));
  ari_table_kernel<<<cdiv(n, 256), 256>>>(da, db, n, kb, tab);
  ari_sums_kernel<<<1, 256>>>(tab, ka, kb, out);
  SMG_CUDA(cudaGetLastError());
  unsigned long long h[3];
  SMG_CUDA(cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost));
  void* ptrs[] = {da, db, tab, out};
  for (void* q : ptrs) cudaFree(q);
  // Hubert-Arabie (mcclust::arandi)
  const double sij = (double)h[0], sa = (double)h[1], sb = (double)h[2], tot = 0.5 * (double)n * (double)(n - 1);
  const double ex = sa * sb / tot, mx = 0.5 * (sa + sb);
  *ari = (mx != ex) ? (sij - ex) / (mx - ex) : 1.0;
  return 0;
}

int smg_trace_ess(const double* traces, int ntraces, int T, int device, double* iat_out, double* ess_out) {
  if (!traces || ntraces < 1 || T < 2) return fail(SMG_ERR_ARG, "bad argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
    return fail(SMG_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  SMG_CUDA(cudaSetDevice(device));
  double *dx = nullptr, *dm = nullptr, *dg = nullptr, *di = nullptr;
  SMG_CUDA(cudaMalloc(&dx, (size_t)ntraces * T * 8));
  SMG_CUDA(cudaMalloc(&dm, (size_t)ntraces * 8));
  SMG_CUDA(cudaMalloc(&dg, (size_t)ntraces * T * 8));
  SMG_CUDA(cudaMalloc(&di, (size_t)ntraces * 8));
  SMG_CUDA(cudaMemcpy(dx, traces, (size_t)ntraces * T * 8, cudaMemcpyHostToDevice));
  trace_mean_kernel<<<ntraces, 256>>>(dx, T, dm);
  trace_autocov_kernel<<<dim3(T, ntraces), 256>>>(dx, T, dm, dg);
  trace_iat_kernel<<<cdiv(ntraces, 64), 64>>>(dg, T, ntraces, di);
  SMG_CUDA(cudaGetLastError());
  std::vector<double> h(ntraces);
  SMG_CUDA(cudaMemcpy(h.data(), di, (size_t)ntraces * 8, cudaMemcpyDeviceToHost));
  void* ptrs[] = {dx, dm, dg, di};
  for (void* q : ptrs) cudaFree(q);
  for (int r = 0; r < ntraces; r++) {
    if (iat_out) iat_out[r] = h[r];
    if (ess_out) ess_out[r] = (double)T / h[r];
  }
  return 0;
}

}  // extern "C"
