// hb_api.cu — the C ABI of libhb_b200.so (declared in include/hb_b200.h).
//
// Host side only: argument checks, host-libm parameter preparation (logs, emission table), the
// per-device scratch cache, host<->device staging for the *_host entry points, kernel launches.
// There is no CPU implementation of the path behind these functions.
#include <cuda_runtime.h>
#include <emmintrin.h>
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <future>
#include <mutex>
#include <thread>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/hb_b200.h"
#include "hb_binom2.cuh"
#include "hb_binom2_fast.cuh"
#include "hb_host_math.h"
#include "hb_kernels.h"
#include "hb_longchain.cuh"
#include "hb_norm3.cuh"
#include "hb_norm3_fast.cuh"
#include "hb_norm3_resident.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return fail(e_ == cudaErrorMemoryAllocation ? HB_ERR_NOMEM : HB_ERR_CUDA,             \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__,        \
                        __LINE__);                                                                \
    } while (0)

#define HBTRY(call)                                                                               \
    do {                                                                                          \
        int rc_ = (call);                                                                         \
        if (rc_ != HB_OK) return rc_;                                                             \
    } while (0)

// Slots of the host-buffer pipeline (staging buffers, events, copy-out jobs).  Two are enough to overlap upload |
// compute | download, but the enqueueing thread waits for the copy-out job that last used a slot before it reuses
// it, and with two slots that wait sat on the critical path (a 2.6 ms bubble on the link per 4 ms download at the
// bench sizes, traced with HB_PIPE_TRACE): four let it run two sub-batches further ahead.
#define HB_PIPE_SLOTS 4

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return HB_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            e = cudaMalloc(&p, bytes);
            want = bytes;
        }
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(HB_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        }
        cap = want;
        return HB_OK;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T *as() { return (T *)p; }
};

// Per-device scratch, grown on demand and reused across calls (no allocation in steady state).
struct Workspace {
    DevBuf psi, ckpt, sdaux, segtab, status, flagged, fbr_gate, pool_fast, lut_lb, lut_rho, lut_rho_d, cellmask, gsize;
    DevBuf lc, lc_fb, lc_lb, lc_y;
    int32_t lc_K = 0;
    int64_t lc_nb = 0;
    int32_t *lc_cstate = nullptr;
    DevBuf st_in, st_in2, st_x, st_x2, st_y, st_g, st_out, st_sd, st_ll, st_member, st_pool, st_pool2, st_pre, st_x80, st_part, st_inp, st_inp2, st_tab;
    // host-buffer entry points: three-stream pipeline (upload | layout + chains | download), staging x2
    DevBuf pin[HB_PIPE_SLOTS], pin2[HB_PIPE_SLOTS], pout_y[HB_PIPE_SLOTS], pout_g[HB_PIPE_SLOTS], pout_ll[HB_PIPE_SLOTS];
    // pinned bounce buffers for ordinary (pageable) host memory, e.g. R vectors: [slot]
    void *bounce_in[HB_PIPE_SLOTS] = {}, *bounce_out[HB_PIPE_SLOTS] = {};
    size_t bounce_in_cap = 0, bounce_out_cap = 0;
    void *bounce_y[HB_PIPE_SLOTS] = {}; // states cross PCIe as bytes and are widened to R integers here
    size_t bounce_y_cap = 0;
    cudaEvent_t ev_bin[HB_PIPE_SLOTS] = {}, ev_bout[HB_PIPE_SLOTS] = {};
    cudaStream_t s_in = nullptr, s_cmp = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[HB_PIPE_SLOTS] = {}, ev_free_in[HB_PIPE_SLOTS] = {},
                ev_out[HB_PIPE_SLOTS] = {}, ev_free_out[HB_PIPE_SLOTS] = {};
    std::vector<int64_t> seg_cached;
    int64_t nchunks = 0;
    int32_t last_norm3_fast = -1, last_binom2_lean = -1; // kernels of the last launches (hb_last_kernels)
    int64_t io_h2d = 0, io_d2h = 0, io_rebuilt = 0;      // bytes of the last *_host pipeline call (hb_last_host_io)
    // Stream hand-over: the scratch above is shared by every call on this device.  Each entry point records
    // ev_use on its stream when it has enqueued its work; a call on a DIFFERENT stream first makes its stream
    // wait for that event, so two streams never overlap on the same scratch (WsUse below).
    cudaEvent_t ev_use = nullptr;
    cudaStream_t last_st = nullptr;
    bool used = false;
    // Pinned staging for small host tables (segment offsets, emission tables): uploaded with
    // cudaMemcpyAsync on the caller's stream, no device-wide or stream-wide synchronisation.
    void *stage = nullptr;
    size_t stage_cap = 0, stage_off = 0;
    cudaEvent_t ev_stage = nullptr;
    bool stage_pending = false;
    double lut_key[3] = {-1, -1, -1};
    int32_t lut_nmax = -1;
    void release_all()
    {
        DevBuf *all[] = {&pool_fast, &fbr_gate, &flagged, &lc, &lc_fb, &lc_lb, &lc_y, &psi, &ckpt, &sdaux, &segtab, &status, &lut_lb, &lut_rho, &lut_rho_d, &cellmask, &gsize,
                         &st_in, &st_in2, &st_x, &st_x2, &st_y, &st_g, &st_out, &st_sd, &st_ll,
                         &st_member, &st_pool, &st_pool2, &st_pre, &st_x80, &st_part, &st_inp, &st_inp2, &st_tab};
        for (DevBuf *b : all) b->release();
        for (int k = 0; k < HB_PIPE_SLOTS; k++) {
            DevBuf *slot[] = {&pin[k], &pin2[k], &pout_y[k], &pout_g[k], &pout_ll[k]};
            for (DevBuf *b : slot) b->release();
        }
        lc_cstate = nullptr; // pointed into `lc`
        lc_K = 0;
        lc_nb = 0;
        if (s_in) {
            cudaStreamDestroy(s_in);
            cudaStreamDestroy(s_cmp);
            cudaStreamDestroy(s_out);
            for (int k = 0; k < HB_PIPE_SLOTS; k++) {
                cudaEventDestroy(ev_in[k]);
                cudaEventDestroy(ev_free_in[k]);
                cudaEventDestroy(ev_out[k]);
                cudaEventDestroy(ev_free_out[k]);
            }
            s_in = s_cmp = s_out = nullptr;
        }
        for (int k = 0; k < HB_PIPE_SLOTS; k++) {
            if (bounce_in[k]) cudaFreeHost(bounce_in[k]);
            if (bounce_out[k]) cudaFreeHost(bounce_out[k]);
            if (bounce_y[k]) cudaFreeHost(bounce_y[k]);
            bounce_in[k] = bounce_out[k] = bounce_y[k] = nullptr;
            if (ev_bin[k]) cudaEventDestroy(ev_bin[k]);
            if (ev_bout[k]) cudaEventDestroy(ev_bout[k]);
            ev_bin[k] = ev_bout[k] = nullptr;
        }
        bounce_in_cap = bounce_out_cap = bounce_y_cap = 0;
        if (stage) cudaFreeHost(stage);
        stage = nullptr;
        stage_cap = stage_off = 0;
        stage_pending = false;
        if (ev_stage) cudaEventDestroy(ev_stage);
        if (ev_use) cudaEventDestroy(ev_use);
        ev_stage = ev_use = nullptr;
        used = false;
        seg_cached.clear();
        lut_nmax = -1;
        lut_key[0] = -1;
    }
};

#define HB_FLAGGED_CAP 16384 // re-run chains recorded per Gaussian launch (hb_norm3_flagged_dev)
std::mutex g_mu;
Workspace g_ws[16];
int g_lut_nmax_cfg = 511;
int g_host_threads = 0; // copy-out threads of the host pipeline; 0 = all but two of this process's CPUs (<= 16)
// Kernel-choice switches are per calling thread (test / experiment hooks; two callers with different
// settings do not interfere).
thread_local int g_binom2_mode = 0; // 0 auto (lean kernel, general kernel behind its gate), 1 general kernel only
thread_local int g_dist_mode = 0;      // 0 pipelined dist kernel (cp.async ring, flag pre-pass), 1 the plain one
thread_local int g_ward_mode = 0;      // 0 auto (cluster kernel from 256 cells up), 1 single block, 2 cluster
thread_local int g_longchain_mode = 0; // 0 auto (pooled allele rounds), 1 never, 2 also the pooled expression rounds
thread_local int g_norm3_mode = 0; // 0 auto, 1 literal kernel only, 2 speculative kernel with every chain re-run exactly, 3 speculative kernel whenever eligible

int cur_ws(Workspace **out)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(HB_ERR_NODEVICE, "no CUDA device: %s", cudaGetErrorString(e));
    }
    if (dev < 0 || dev >= 16) return fail(HB_ERR_ARG, "device index %d out of range", dev);
    *out = &g_ws[dev];
    return HB_OK;
}

// Host table -> device on `st` without synchronising the stream: the bytes are parked in a pinned arena
// that lives until the copy has run.  The arena is recycled per API call (stage_begin, from WsUse::enter); the only host wait
// is for an EARLIER call's table copy that has not executed yet, and only when tables change (they are
// cached per workspace), never in steady state.
int stage_begin(Workspace *ws)
{
    if (ws->stage_pending) {
        CU(cudaEventSynchronize(ws->ev_stage));
        ws->stage_pending = false;
    }
    ws->stage_off = 0;
    return HB_OK;
}
int upload_async(Workspace *ws, void *dst, const void *src, size_t bytes, cudaStream_t st)
{
    if (!bytes) return HB_OK;
    const size_t need = ws->stage_off + ((bytes + 255) & ~(size_t)255);
    if (need > ws->stage_cap) {
        // earlier copies of this call may still read the old arena: let them finish before it goes
        if (ws->stage_off) CU(cudaStreamSynchronize(st));
        void *np_ = nullptr;
        const size_t cap = std::max(need + need / 2, (size_t)1 << 20);
        CU(cudaHostAlloc(&np_, cap, cudaHostAllocDefault));
        if (ws->stage) cudaFreeHost(ws->stage);
        ws->stage = np_;
        ws->stage_cap = cap;
        ws->stage_off = 0;
    }
    if (!ws->ev_stage) CU(cudaEventCreateWithFlags(&ws->ev_stage, cudaEventDisableTiming));
    char *b = (char *)ws->stage + ws->stage_off;
    memcpy(b, src, bytes);
    ws->stage_off += (bytes + 255) & ~(size_t)255;
    CU(cudaMemcpyAsync(dst, b, bytes, cudaMemcpyHostToDevice, st));
    CU(cudaEventRecord(ws->ev_stage, st));
    ws->stage_pending = true;
    return HB_OK;
}

// RAII: orders this call's use of the per-device scratch after the previous user's (other stream).
struct WsUse {
    Workspace *ws = nullptr;
    cudaStream_t st = nullptr;
    int enter(Workspace *w, cudaStream_t s)
    {
        if (!w->ev_use) CU(cudaEventCreateWithFlags(&w->ev_use, cudaEventDisableTiming));
        if (w->used && w->last_st != s) CU(cudaStreamWaitEvent(s, w->ev_use, 0));
        HBTRY(stage_begin(w));
        ws = w;
        st = s;
        return HB_OK;
    }
    ~WsUse()
    {
        if (!ws) return;
        if (cudaEventRecord(ws->ev_use, st) == cudaSuccess) {
            ws->last_st = st;
            ws->used = true;
        } else {
            cudaGetLastError();
        }
    }
};
#define WS_USE(stream_)                                                                           \
    WsUse ws_use_;                                                                                \
    HBTRY(ws_use_.enter(ws, (cudaStream_t)(stream_)))

// Segment tables on the device: seg_off[nseg+1] | chunk_off[nseg+1] | order[nseg] (int32).
int prepare_segments(Workspace *ws, const int64_t *seg_off, int32_t nseg, int64_t T,
                     cudaStream_t st, const int64_t **d_seg, const int64_t **d_chunk,
                     const int32_t **d_order)
{
    if (nseg < 1 || !seg_off) return fail(HB_ERR_ARG, "need nseg >= 1 and seg_off");
    if (seg_off[0] < 0 || seg_off[nseg] > T) return fail(HB_ERR_ARG, "seg_off outside [0, T]");
    for (int32_t s = 0; s < nseg; s++)
        if (seg_off[s + 1] < seg_off[s]) return fail(HB_ERR_ARG, "seg_off must be non-decreasing");
    const size_t n64 = (size_t)(nseg + 1);
    const size_t bytes = 2 * n64 * sizeof(int64_t) + (size_t)nseg * sizeof(int32_t);
    const bool same = ws->seg_cached.size() == n64 &&
                      memcmp(ws->seg_cached.data(), seg_off, n64 * sizeof(int64_t)) == 0 &&
                      ws->segtab.cap >= bytes;
    if (!same) {
        HBTRY(ws->segtab.ensure(bytes));
        std::vector<int64_t> host(2 * n64);
        std::vector<int32_t> order(nseg);
        int64_t acc = 0;
        for (int32_t s = 0; s <= nseg; s++) {
            host[s] = seg_off[s];
            host[n64 + s] = acc;
            if (s < nseg) acc += (seg_off[s + 1] - seg_off[s] + HB_CHUNK - 1) / HB_CHUNK;
        }
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) {
            return (seg_off[a + 1] - seg_off[a]) > (seg_off[b + 1] - seg_off[b]);
        });
        // through the pinned arena: the temporaries die at the end of this scope, the stream is not waited for
        HBTRY(upload_async(ws, ws->segtab.p, host.data(), 2 * n64 * sizeof(int64_t), st));
        HBTRY(upload_async(ws, (char *)ws->segtab.p + 2 * n64 * sizeof(int64_t), order.data(),
                           (size_t)nseg * sizeof(int32_t), st));
        ws->seg_cached.assign(seg_off, seg_off + n64);
        ws->nchunks = acc;
    }
    *d_seg = ws->segtab.as<int64_t>();
    *d_chunk = ws->segtab.as<int64_t>() + n64;
    *d_order = (const int32_t *)((char *)ws->segtab.p + 2 * n64 * sizeof(int64_t));
    return HB_OK;
}

bool all_finite(const double *v, int n)
{
    for (int i = 0; i < n; i++)
        if (!std::isfinite(v[i])) return false;
    return true;
}

int fill_norm3_const(hb::Norm3Const &c, const double *mean, const double *Pi, const double *delta,
                     bool *sym)
{
    if (!mean || !Pi || !delta) return fail(HB_ERR_ARG, "mean, Pi and delta are required");
    if (!all_finite(mean, 3) || !all_finite(Pi, 9) || !all_finite(delta, 3))
        return fail(HB_ERR_ARG, "non-finite HMM parameter");
    for (int i = 0; i < 3; i++) {
        c.mean[i] = mean[i];
        c.delta[i] = delta[i];
        c.logdelta[i] = log(delta[i]); // log(0) = -inf, as in R
    }
    for (int i = 0; i < 9; i++) {
        if (Pi[i] < 0) return fail(HB_ERR_ARG, "negative transition probability");
        c.Pi[i] = Pi[i];
        c.logPi[i] = log(Pi[i]);
    }
    const bool sym_pi = Pi[0] == Pi[4] && Pi[0] == Pi[8] && Pi[1] == Pi[2] && Pi[1] == Pi[3] &&
                        Pi[1] == Pi[5] && Pi[1] == Pi[6] && Pi[1] == Pi[7] && Pi[1] > 0;
    const bool sym_mean = (mean[0] - mean[1]) == -(mean[2] - mean[1]);
    *sym = sym_pi && sym_mean;
    c.la = c.logPi[0];
    c.lt = c.logPi[1];
    c.amt = Pi[0] - Pi[1];
    c.tt = Pi[1];
    c.tau = c.lt - c.la;
    c.dl0 = c.logdelta[0] - c.logdelta[1];
    c.dl2 = c.logdelta[2] - c.logdelta[1];
    c.mumax2 = std::max(mean[0] * mean[0], std::max(mean[1] * mean[1], mean[2] * mean[2]));
    c.dmu2 = std::max((mean[0] - mean[1]) * (mean[0] - mean[1]), (mean[2] - mean[1]) * (mean[2] - mean[1]));
    c.ldabs = 0;
    for (int i = 0; i < 3; i++)
        if (std::isfinite(c.logdelta[i])) c.ldabs = std::max(c.ldabs, fabs(c.logdelta[i]));
    c.eps_scale = 1.0;
    hb::norm3_fill_k(c);
    return HB_OK;
}

int check_flags(int32_t flags, const void *y, const void *gamma, const void *ll, bool *vit,
                bool *fb)
{
    *vit = (flags & HB_WANT_VITERBI) != 0;
    *fb = (flags & (HB_WANT_GAMMA | HB_WANT_LL)) != 0;
    if (!*vit && !*fb) return fail(HB_ERR_ARG, "flags select nothing to compute");
    if (*vit && !y) return fail(HB_ERR_ARG, "HB_WANT_VITERBI needs y");
    if ((flags & HB_WANT_GAMMA) && !gamma) return fail(HB_ERR_ARG, "HB_WANT_GAMMA needs gamma");
    if ((flags & HB_WANT_LL) && !ll) return fail(HB_ERR_ARG, "HB_WANT_LL needs ll");
    return HB_OK;
}

// Core launch on device-resident gene-major data.  inv_sd/log_sd may be NULL (computed on device).
int norm3_run(Workspace *ws, const double *x, int64_t ld, int64_t T, int64_t B,
              const int64_t *seg_off, int32_t nseg, const double *mean, const double *sd,
              const double *inv_sd, const double *log_sd, int32_t sd_per_seg, const double *Pi,
              const double *delta, int32_t flags, uint8_t *y, int64_t ldy, double *gamma,
              int64_t ldg, double *ll, cudaStream_t st, bool reset_status = true, bool tiled = false)
{
    if (T < 1 || B < 1 || (!tiled && ld < B))
        return fail(HB_ERR_ARG, "bad shape T=%lld B=%lld ld=%lld", (long long)T, (long long)B, (long long)ld);
    if (!x || !sd) return fail(HB_ERR_ARG, "x and sd are required");
    bool vit, fb, sym;
    HBTRY(check_flags(flags, y, gamma, ll, &vit, &fb));
    if (!tiled && vit && ldy < B) return fail(HB_ERR_ARG, "ldy < B");
    if (!tiled && (flags & HB_WANT_GAMMA) && ldg < B) return fail(HB_ERR_ARG, "ldg < B");
    const int64_t Bp = tiled ? (B + 31) / 32 * 32 : B; // columns the scratch is laid out for
    hb::Norm3Args a;
    memset(&a, 0, sizeof a);
    HBTRY(fill_norm3_const(a.c, mean, Pi, delta, &sym));
    HBTRY(prepare_segments(ws, seg_off, nseg, T, st, &a.seg_off, &a.chunk_off, &a.seg_order));
    const int64_t nsd = sd_per_seg ? (int64_t)nseg * B : B;
    if (!inv_sd || !log_sd) {
        HBTRY(ws->sdaux.ensure((size_t)nsd * 2 * sizeof(double)));
        double *aux = ws->sdaux.as<double>();
        CU(hb::launch_sd_prep(sd, nsd, aux, aux + nsd, st));
        inv_sd = aux;
        log_sd = aux + nsd;
    }
    HBTRY(ws->status.ensure(4 * sizeof(int32_t)));
    if (vit) HBTRY(ws->psi.ensure((size_t)ws->nchunks * Bp * sizeof(uint64_t)));
    if (fb) HBTRY(ws->ckpt.ensure((size_t)ws->nchunks * 3 * Bp * sizeof(double)));
    if (reset_status) CU(cudaMemsetAsync(ws->status.p, 0, 4 * sizeof(int32_t), st));
    a.x = x;
    a.ld = ld;
    a.T = T;
    a.B = B;
    a.tiled = tiled ? 1 : 0;
    a.gplane = tiled ? T * Bp : T * ldg;
    a.nchunks = ws->nchunks;
    a.nseg = nseg;
    a.nblk_per_seg = (int32_t)((B + HB_BLOCK - 1) / HB_BLOCK);
    a.sd = sd;
    a.inv_sd = inv_sd;
    a.log_sd = log_sd;
    a.sd_per_seg = sd_per_seg ? 1 : 0;
    a.y = y;
    a.ldy = ldy;
    a.gamma = (flags & HB_WANT_GAMMA) ? gamma : nullptr;
    a.ldg = ldg;
    a.ll = (flags & HB_WANT_LL) ? ll : nullptr;
    a.psi = ws->psi.as<uint64_t>();
    a.ckpt = ws->ckpt.as<double>();
    a.status = ws->status.as<int32_t>();
    HBTRY(ws->flagged.ensure((size_t)HB_FLAGGED_CAP * sizeof(int64_t)));
    a.flagged = ws->flagged.as<int64_t>();
    a.flagged_cap = HB_FLAGGED_CAP;
    // HB_WANT_LL without HB_WANT_GAMMA: the forward sweep alone (the kernels return after storing LL when
    // no posterior buffer is given); with Viterbi as well, the Viterbi-only launch precedes it.
    const bool ll_only = fb && !a.gamma;
    // Speculate-and-verify kernel: the reference's symmetric parameter block, state 1 reachable at
    // t = 0 (the differences are taken against nu_1), chains short enough for a useful error bound.
    // Auto mode uses it when every row starts on a 1 KB boundary (ld, ldy/8... multiples of 128
    // elements): measured 94 vs 90.5 G steps/s there, and 88.5 vs 90.7 with unpadded rows, because
    // both kernels then sit on the same memory-system ceiling.
    int64_t lmax = 0;
    for (int32_t s = 0; s < nseg; s++) lmax = std::max(lmax, seg_off[s + 1] - seg_off[s]);
    const bool fast_ok = sym && delta[1] > 0 && Pi[0] > 0 && lmax <= HB_FAST_LMAX;
    const bool padded = tiled || (ld % 128 == 0 && (!a.gamma || ldg % 128 == 0));
    const bool fast = fast_ok && (g_norm3_mode >= 2 || (g_norm3_mode == 0 && padded));
    if (fast && g_norm3_mode == 2) a.c.eps_scale = INFINITY;
    ws->last_norm3_fast = fast ? 1 : 0;
    auto launch = [&](bool v, bool f) {
        return fast ? hb::launch_norm3_fast(a, v, f, st) : hb::launch_norm3(a, sym, v, f, st);
    };
    if (ll_only) {
        if (vit) CU(launch(true, false));
        CU(launch(false, true));
    } else {
        CU(launch(vit, fb));
    }
    return HB_OK;
}

// Few long chains, Viterbi parallel along the gene axis (hb_longchain.cuh).  lb: device [K][T][S]
// literal log-emissions; y: device [K][T] int32, 1-based.  Enqueues ~12 small kernels, no sync.
int longchain_run(Workspace *ws, int32_t S, const double *lb, int64_t T, int32_t K, const double *Pi,
                  const double *delta, int32_t *y, cudaStream_t st, bool reset_status = true, int32_t L = 0)
{
    if ((S != 2 && S != 3) || T < 1 || K < 1 || !lb || !y || !Pi || !delta)
        return fail(HB_ERR_ARG, "bad arguments");
    if (!all_finite(Pi, S * S) || !all_finite(delta, S)) return fail(HB_ERR_ARG, "bad HMM parameter");
    hb::LcArgs a;
    memset(&a, 0, sizeof a);
    a.lb = lb;
    a.T = T;
    a.K = K;
    a.S = S;
    a.L = L > 0 ? L : hb::lc_block_len(T);
    a.nb = T > 1 ? (T - 1 + a.L - 1) / a.L : 0;
    for (int i = 0; i < S * S; i++) {
        if (Pi[i] < 0) return fail(HB_ERR_ARG, "negative transition probability");
        a.lp[i] = log(Pi[i]);
    }
    for (int i = 0; i < S; i++) a.ld[i] = log(delta[i]);
    const size_t nkb = (size_t)K * a.nb + 1, al = 256;
    auto up = [&](size_t v) { return (v + al - 1) / al * al; };
    size_t off = 0;
    const size_t o_f = off; off += up(nkb * S * S * sizeof(double));
    const size_t o_im = off; off += up(nkb * 2 * S * S * sizeof(int64_t));
    const size_t o_en = off; off += up(nkb * S * sizeof(double));
    const size_t o_le = off; off += up(nkb * S * sizeof(double));
    const size_t o_be = off; off += up(nkb * sizeof(int32_t));
    const size_t o_cs = off; off += up((size_t)K * HB_LC_CS * sizeof(int32_t));
    const size_t o_io = off; off += up(nkb);
    const size_t o_bm = off; off += up(nkb);
    const size_t o_yb = off; off += up(nkb);
    const size_t o_ps = off; off += up((size_t)K * T);
    HBTRY(ws->lc.ensure(off));
    HBTRY(ws->status.ensure(4 * sizeof(int32_t)));
    char *base = ws->lc.as<char>();
    a.fmat = (double *)(base + o_f);
    a.imat = (int64_t *)(base + o_im);
    a.entry = (double *)(base + o_en);
    a.lexit = (double *)(base + o_le);
    a.bexp = (int32_t *)(base + o_be);
    a.cstate = (int32_t *)(base + o_cs);
    a.iok = (uint8_t *)(base + o_io);
    a.bmap = (uint8_t *)(base + o_bm);
    a.ybound = (uint8_t *)(base + o_yb);
    a.psi = (uint8_t *)(base + o_ps);
    a.y = y;
    a.status = ws->status.as<int32_t>();
    if (reset_status) CU(cudaMemsetAsync(ws->status.p, 0, 4 * sizeof(int32_t), st));
    CU(cudaMemsetAsync(a.cstate, 0, (size_t)K * HB_LC_CS * sizeof(int32_t), st));
    ws->lc_K = K;
    ws->lc_nb = a.nb;
    ws->lc_cstate = a.cstate;
    CU(hb::launch_longchain(a, st));
    return HB_OK;
}

// Posteriors of a few long chains (lc_fb_* in hb_longchain.cuh).  gamma: device [K][T][S]; ll: device [K] or NULL.
int longchain_fb_run(Workspace *ws, int32_t S, const double *lb, int64_t T, int32_t K, const double *Pi,
                     const double *delta, double *gamma, double *ll, cudaStream_t st, int32_t L = 0,
                     bool planes = false)
{
    if ((S != 2 && S != 3) || T < 1 || K < 1 || !lb || !gamma || !Pi || !delta)
        return fail(HB_ERR_ARG, "bad arguments");
    if (!all_finite(Pi, S * S) || !all_finite(delta, S)) return fail(HB_ERR_ARG, "bad HMM parameter");
    hb::LcArgs a;
    memset(&a, 0, sizeof a);
    a.lb = lb;
    a.T = T;
    a.K = K;
    a.S = S;
    a.L = L > 0 ? L : hb::lc_block_len(T);
    a.nb = T > 1 ? (T - 1 + a.L - 1) / a.L : 0;
    for (int i = 0; i < S * S; i++) {
        if (Pi[i] < 0) return fail(HB_ERR_ARG, "negative transition probability");
        a.pi[i] = Pi[i];
    }
    for (int i = 0; i < S; i++) a.dl[i] = delta[i];
    const size_t nkb = (size_t)K * a.nb + 1, al = 256;
    auto up = [&](size_t v) { return (v + al - 1) / al * al; };
    size_t off = 0;
    const size_t o_f = off; off += up(nkb * S * S * sizeof(double));
    const size_t o_ae = off; off += up(nkb * S * sizeof(double));
    const size_t o_be = off; off += up(nkb * S * sizeof(double));
    const size_t o_ls = off; off += up(nkb * sizeof(double));
    const size_t o_e = off; off += up(nkb * S * sizeof(int32_t));
    HBTRY(ws->lc_fb.ensure(off));
    char *base = ws->lc_fb.as<char>();
    a.ffm = (double *)(base + o_f);
    a.fae = (double *)(base + o_ae);
    a.fbe = (double *)(base + o_be);
    a.fls = (double *)(base + o_ls);
    a.ffe = (int32_t *)(base + o_e);
    a.gamma = gamma;
    a.ll = ll;
    a.g_sc = planes ? T : T * S; // planes: R's [T x K x S] column-major, one plane per state
    a.g_st = planes ? 1 : S;
    a.g_sk = planes ? (int64_t)K * T : 1;
    CU(hb::launch_longchain_fb(a, st));
    return HB_OK;
}

int read_status(Workspace *ws, cudaStream_t st)
{
    int32_t s = 0;
    CU(cudaMemcpyAsync(&s, ws->status.p, sizeof s, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (s == 1) return fail(HB_ERR_UNDERFLOW, "Problems With Underflow");
    if (s == 2) return fail(HB_ERR_ARG, "beta-binomial count above the emission table");
    return HB_OK;
}

// Emission table for the binomial chains, cached per (p0, p1, rho, nmax).
int prepare_lut(Workspace *ws, const double *prob, double rho, cudaStream_t st)
{
    const int32_t nmax = g_lut_nmax_cfg;
    if (ws->lut_nmax == nmax && ws->lut_key[0] == prob[0] && ws->lut_key[1] == prob[1] &&
        ws->lut_key[2] == rho)
        return HB_OK;
    const size_t cnt = (size_t)(nmax + 1) * (nmax + 2) / 2;
    std::vector<double> lb(2 * cnt);
    std::vector<hb::RhoME> rr(cnt);
    std::vector<double> rd(cnt);
    const hbhost::BetaBinomCorr bb0(prob[0], rho, nmax), bb1(prob[1], rho, nmax);
    for (int32_t n = 0; n <= nmax; n++)
        for (int32_t x = 0; x <= n; x++) {
            size_t ix = (size_t)n * (n + 1) / 2 + x;
            double l0, l1;
            if (rho > 0) {
                l0 = hbhost::betabinom_logpmf(bb0, x, n, prob[0]);
                l1 = hbhost::betabinom_logpmf(bb1, x, n, prob[1]);
            } else {
                l0 = hbhost::binom_logpmf(x, n, prob[0]);
                l1 = hbhost::binom_logpmf(x, n, prob[1]);
            }
            lb[2 * ix] = l0;
            lb[2 * ix + 1] = l1;
            rr[ix] = hb::rho_from_logratio(l0 - l1);
            // impossible under both states: the general path uses rho = 1 (binom2_emit_slow)
            rd[ix] = (l0 == -INFINITY && l1 == -INFINITY) ? 1.0 : hb::b2f_rho_double(rr[ix]);
        }
    HBTRY(ws->lut_lb.ensure(lb.size() * sizeof(double)));
    HBTRY(ws->lut_rho.ensure(rr.size() * sizeof(hb::RhoME)));
    HBTRY(ws->lut_rho_d.ensure(rd.size() * sizeof(double)));
    HBTRY(upload_async(ws, ws->lut_lb.p, lb.data(), lb.size() * sizeof(double), st));
    HBTRY(upload_async(ws, ws->lut_rho.p, rr.data(), rr.size() * sizeof(hb::RhoME), st));
    HBTRY(upload_async(ws, ws->lut_rho_d.p, rd.data(), rd.size() * sizeof(double), st));
    ws->lut_nmax = nmax;
    ws->lut_key[0] = prob[0];
    ws->lut_key[1] = prob[1];
    ws->lut_key[2] = rho;
    return HB_OK;
}

int binom2_run(Workspace *ws, const int32_t *x, const int32_t *n, int64_t ld, int64_t T, int64_t B,
               const int64_t *seg_off, int32_t nseg, const double *prob, double rho,
               const double *Pi, const double *delta, int32_t flags, uint8_t *y, int64_t ldy,
               double *gamma, int64_t ldg, double *ll, cudaStream_t st, bool reset_status = true,
               bool pooled = false, bool tiled = false, bool packed = false)
{
    if (T < 1 || B < 1 || (!tiled && ld < B))
        return fail(HB_ERR_ARG, "bad shape T=%lld B=%lld ld=%lld", (long long)T, (long long)B, (long long)ld);
    const int64_t Bp = tiled ? (B + 31) / 32 * 32 : B;
    if (!x || (!n && !packed) || !prob || !Pi || !delta) return fail(HB_ERR_ARG, "missing argument");
    if (!all_finite(prob, 2) || !all_finite(Pi, 4) || !all_finite(delta, 2) || !(rho >= 0) ||
        !(rho < 1))
        return fail(HB_ERR_ARG, "bad HMM parameter");
    for (int k = 0; k < 2; k++)
        if (prob[k] < 0 || prob[k] > 1) return fail(HB_ERR_ARG, "prob outside [0,1]");
    bool vit, fb;
    HBTRY(check_flags(flags, y, gamma, ll, &vit, &fb));
    if (!tiled && vit && ldy < B) return fail(HB_ERR_ARG, "ldy < B");
    if (!tiled && (flags & HB_WANT_GAMMA) && ldg < B) return fail(HB_ERR_ARG, "ldg < B");
    const bool ll_only = fb && !(flags & HB_WANT_GAMMA);
    hb::Binom2Args a;
    memset(&a, 0, sizeof a);
    for (int k = 0; k < 2; k++) {
        a.c.prob[k] = prob[k];
        a.c.q[k] = 1 - prob[k];
        a.c.logp[k] = log(prob[k]);
        a.c.logq[k] = log(1 - prob[k]);
        a.c.delta[k] = delta[k];
        a.c.logdelta[k] = log(delta[k]);
    }
    for (int i = 0; i < 4; i++) {
        if (Pi[i] < 0) return fail(HB_ERR_ARG, "negative transition probability");
        a.c.Pi[i] = Pi[i];
        a.c.logPi[i] = log(Pi[i]);
    }
    a.c.bb = rho > 0 ? 1 : 0;
    HBTRY(prepare_segments(ws, seg_off, nseg, T, st, &a.seg_off, &a.chunk_off, &a.seg_order));
    HBTRY(prepare_lut(ws, prob, rho, st));
    HBTRY(ws->status.ensure(4 * sizeof(int32_t)));
    if (vit) HBTRY(ws->psi.ensure((size_t)ws->nchunks * Bp * sizeof(uint64_t)));
    if (fb) HBTRY(ws->ckpt.ensure((size_t)ws->nchunks * 3 * Bp * sizeof(double)));
    if (reset_status) CU(cudaMemsetAsync(ws->status.p, 0, 4 * sizeof(int32_t), st));
    a.x = x;
    a.n = n;
    a.ld = ld;
    a.T = T;
    a.B = B;
    a.tiled = tiled ? 1 : 0;
    a.packed = packed ? 1 : 0;
    a.gplane = tiled ? T * Bp : T * ldg;
    a.nchunks = ws->nchunks;
    a.nseg = nseg;
    a.nblk_per_seg = (int32_t)((B + HB_BLOCK - 1) / HB_BLOCK);
    a.lut_lb = ws->lut_lb.as<double2>();
    a.lut_rho = ws->lut_rho.as<hb::RhoME>();
    a.lut_rho_d = ws->lut_rho_d.as<double>();
    a.lut_nmax = ws->lut_nmax;
    a.y = y;
    a.ldy = ldy;
    a.gamma = (flags & HB_WANT_GAMMA) ? gamma : nullptr;
    a.ldg = ldg;
    a.ll = (flags & HB_WANT_LL) ? ll : nullptr;
    a.psi = ws->psi.as<uint64_t>();
    a.ckpt = ws->ckpt.as<double>();
    a.status = ws->status.as<int32_t>();
    // Lean kernel first (plain-double filter, rows of Pi positive); it raises the gate when some
    // position carries evidence beyond 2^+-900, and the general kernel then redoes the launch
    // (otherwise its blocks exit at once).  Both are enqueued: no host synchronisation here.
    if (pooled && vit && !fb) {
        // Pooled chains (a handful of long chains, deep counts at every position): evaluate all
        // emissions in a parallel pre-pass, then run the literal Viterbi over the stored values.
        HBTRY(ws->st_pre.ensure((size_t)T * ld * sizeof(double2)));
        CU(hb::launch_binom2_emit(a, ws->st_pre.as<double2>(), st));
        a.pre_lb = ws->st_pre.as<double2>();
        CU(hb::launch_binom2(a, vit, fb, st));
        return HB_OK;
    }
    const bool lean_ok = Pi[0] > 0 && Pi[1] > 0 && Pi[2] > 0 && Pi[3] > 0;
    const bool lean = lean_ok && g_binom2_mode != 1;
    ws->last_binom2_lean = lean ? 1 : 0;
    if (lean) a.gate = ws->status.as<int32_t>() + 2;
    auto launch = [&](bool v, bool f) -> cudaError_t {
        hb::Binom2Args c = a;
        if (lean) {
            cudaError_t e = hb::launch_binom2_fast(c, v, f, st);
            if (e != cudaSuccess) return e;
            c.gate_in = c.gate;
        }
        return hb::launch_binom2(c, v, f, st);
    };
    // HB_WANT_LL without HB_WANT_GAMMA: Viterbi-only launch (if wanted), then the forward sweep alone
    if (ll_only) {
        if (vit) CU(launch(true, false));
        CU(launch(false, true));
    } else {
        CU(launch(vit, fb));
    }
    return HB_OK;
}

// Pooled allele round on the long-chain path: literal emissions of all (chain, position) pairs in
// parallel, then the gene-axis parallel Viterbi.  x, n: device [K][G] pooled counts; y: device [K][G].
// Literal dbinom log-emissions of pooled count vectors x, n (device [K][G]) into ws->lc_lb ([K][G][2]);
// resets the status words.
int binom2_emissions_cm(Workspace *ws, const int32_t *x, const int32_t *n, int64_t G, int32_t K,
                        const double *prob, cudaStream_t st)
{
    if (!all_finite(prob, 2) || prob[0] < 0 || prob[0] > 1 || prob[1] < 0 || prob[1] > 1)
        return fail(HB_ERR_ARG, "prob outside [0,1]");
    hb::Binom2Args a;
    memset(&a, 0, sizeof a);
    for (int k = 0; k < 2; k++) {
        a.c.prob[k] = prob[k];
        a.c.q[k] = 1 - prob[k];
        a.c.logp[k] = log(prob[k]);
        a.c.logq[k] = log(1 - prob[k]);
    }
    HBTRY(prepare_lut(ws, prob, 0.0, st));
    HBTRY(ws->status.ensure(4 * sizeof(int32_t)));
    CU(cudaMemsetAsync(ws->status.p, 0, 4 * sizeof(int32_t), st));
    a.lut_lb = ws->lut_lb.as<double2>();
    a.lut_rho = ws->lut_rho.as<hb::RhoME>();
    a.lut_rho_d = ws->lut_rho_d.as<double>();
    a.lut_nmax = ws->lut_nmax;
    a.status = ws->status.as<int32_t>();
    HBTRY(ws->lc_lb.ensure((size_t)K * G * sizeof(double2)));
    CU(hb::launch_binom2_emit_cm(a, x, n, (int64_t)K * G, ws->lc_lb.as<double2>(), st));
    return HB_OK;
}

// Pooled allele round on the long-chain path: literal emissions of all (chain, position) pairs in
// parallel, then the gene-axis parallel Viterbi.  x, n: device [K][G] pooled counts; y: device [K][G].
int binom2_longchain(Workspace *ws, const int32_t *x, const int32_t *n, int64_t G, int32_t K,
                     const double *prob, const double *Pi, const double *delta, int32_t *y, cudaStream_t st)
{
    HBTRY(binom2_emissions_cm(ws, x, n, G, K, prob, st));
    return longchain_run(ws, 2, ws->lc_lb.as<double>(), G, K, Pi, delta, y, st, false);
}

// Host calls shaped like the reference's own: a few chains, each long, one segment, Viterbi only.
bool long_chain_shape(int64_t T, int64_t B, int32_t nseg, int32_t flags)
{
    (void)flags; // Viterbi, posteriors and LL all have a long-chain form
    return g_longchain_mode != 1 && nseg == 1 && B <= 32 && T >= 1024;
}

// Shared tail of the long-chain host routes: lb = ws->lc_lb holds the literal log-emissions [B][T][S].
int longchain_host_outputs(Workspace *ws, int32_t S, int64_t T, int64_t B, const double *Pi, const double *delta,
                           int32_t flags, int32_t *y, double *gamma, double *ll, bool status_reset)
{
    const bool vit = (flags & HB_WANT_VITERBI) != 0, want_g = (flags & HB_WANT_GAMMA) != 0,
               want_ll = (flags & HB_WANT_LL) != 0;
    if (vit) {
        HBTRY(ws->st_out.ensure((size_t)T * B * sizeof(int32_t)));
        HBTRY(longchain_run(ws, S, ws->lc_lb.as<double>(), T, (int32_t)B, Pi, delta, ws->st_out.as<int32_t>(), 0,
                            status_reset));
        CU(cudaMemcpyAsync(y, ws->st_out.p, (size_t)T * B * sizeof(int32_t), cudaMemcpyDeviceToHost, 0));
    }
    if (want_g || want_ll) { // (LL alone: the posteriors land in scratch and stay on the device)
        HBTRY(ws->st_g.ensure((size_t)T * B * S * sizeof(double)));
        HBTRY(ws->st_ll.ensure((size_t)B * sizeof(double)));
        HBTRY(longchain_fb_run(ws, S, ws->lc_lb.as<double>(), T, (int32_t)B, Pi, delta, ws->st_g.as<double>(),
                               want_ll ? ws->st_ll.as<double>() : nullptr, 0, 0, true));
        if (want_g)
            CU(cudaMemcpyAsync(gamma, ws->st_g.p, (size_t)T * B * S * sizeof(double), cudaMemcpyDeviceToHost, 0));
        if (want_ll)
            CU(cudaMemcpyAsync(ll, ws->st_ll.p, (size_t)B * sizeof(double), cudaMemcpyDeviceToHost, 0));
    }
    return read_status(ws, 0);
}

int64_t pad_ld(int64_t B) { return (B + 15) / 16 * 16; } // 128-byte rows for fp64

// membership -> group sizes (host) ; returns HB_ERR_ARG on empty groups
int group_sizes(const uint8_t *member, int32_t K, int64_t N, std::vector<int32_t> &sz)
{
    sz.assign(K, 0);
    for (int32_t k = 0; k < K; k++) {
        int32_t c = 0;
        for (int64_t i = 0; i < N; i++) c += member[(int64_t)k * N + i] != 0;
        if (c < 1) return fail(HB_ERR_ARG, "group %d has no cells", k);
        sz[k] = c;
    }
    return HB_OK;
}

} // namespace

// =====================================================================================
extern "C" {

int hb_version(void) { return 100; }

const char *hb_last_error(void) { return g_err.c_str(); }

int hb_device_count(int *count)
{
    if (!count) return fail(HB_ERR_ARG, "count is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *count = 0;
        return fail(HB_ERR_NODEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    *count = n;
    return HB_OK;
}

int hb_set_device(int device)
{
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(HB_ERR_NODEVICE, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
    }
    return HB_OK;
}

int hb_release_workspace(void)
{
    std::lock_guard<std::mutex> lk(g_mu);
    Workspace *ws;
    HBTRY(cur_ws(&ws));
    cudaDeviceSynchronize();
    ws->release_all();
    return HB_OK;
}

int hb_set_binom_lut_max(int32_t nmax)
{
    if (nmax < 0 || nmax > 4095) return fail(HB_ERR_ARG, "nmax must be in [0, 4095]");
    std::lock_guard<std::mutex> lk(g_mu);
    g_lut_nmax_cfg = nmax;
    return HB_OK;
}

int hb_set_host_threads(int32_t n)
{
    if (n < 0 || n > 256) return fail(HB_ERR_ARG, "n must be in [0, 256]");
    std::lock_guard<std::mutex> lk(g_mu);
    g_host_threads = n;
    return HB_OK;
}

int hb_set_norm3_mode(int32_t mode)
{
    if (mode < 0 || mode > 3)
        return fail(HB_ERR_ARG, "mode must be 0 (auto), 1 (literal), 2 (speculative, re-run all) or 3 (speculative)");
    g_norm3_mode = mode; // thread-local
    return HB_OK;
}

int hb_set_binom2_mode(int32_t mode)
{
    if (mode < 0 || mode > 1) return fail(HB_ERR_ARG, "mode must be 0 (auto) or 1 (general kernel only)");
    g_binom2_mode = mode; // thread-local
    return HB_OK;
}

int hb_binom2_stats_dev(void *stream, int32_t *general_rerun)
{
    std::lock_guard<std::mutex> lk(g_mu);
    Workspace *ws;
    HBTRY(cur_ws(&ws));
    WS_USE(stream);
    if (!general_rerun) return fail(HB_ERR_ARG, "general_rerun is NULL");
    *general_rerun = 0;
    if (!ws->status.p) return HB_OK;
    int32_t s[3] = {0, 0, 0};
    CU(cudaMemcpyAsync(s, ws->status.p, sizeof s, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CU(cudaStreamSynchronize((cudaStream_t)stream));
    *general_rerun = s[2];
    return HB_OK;
}

int hb_norm3_stats_dev(void *stream, int64_t *rerun_chains)
{
    std::lock_guard<std::mutex> lk(g_mu);
    Workspace *ws;
    HBTRY(cur_ws(&ws));
    WS_USE(stream);
    if (!rerun_chains) return fail(HB_ERR_ARG, "rerun_chains is NULL");
    *rerun_chains = 0;
    if (!ws->status.p) return HB_OK;
    int32_t s[2] = {0, 0};
    CU(cudaMemcpyAsync(s, ws->status.p, sizeof s, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CU(cudaStreamSynchronize((cudaStream_t)stream));
    *rerun_chains = s[1];
    return HB_OK;
}

int hb_last_kernels(int32_t *norm3_fast, int32_t *binom2_lean)
{
    std::lock_guard<std::mutex> lk(g_mu);
    Workspace *ws;
    HBTRY(cur_ws(&ws));
    if (norm3_fast) *norm3_fast = ws->last_norm3_fast;
    if (binom2_lean) *binom2_lean = ws->last_binom2_lean;
    return HB_OK;
}

int hb_norm3_flagged_dev(void *stream, int64_t *chains, int64_t cap, int64_t *n)
{
    std::lock_guard<std::mutex> lk(g_mu);
    Workspace *ws;
    HBTRY(cur_ws(&ws));
    WS_USE(stream);
    if (!n || cap < 0 || (cap > 0 && !chains)) return fail(HB_ERR_ARG, "bad arguments");
    *n = 0;
    if (!ws->status.p || !ws->flagged.p) return HB_OK;
    int32_t s[2] = {0, 0};
    CU(cudaMemcpyAsync(s, ws->status.p, sizeof s, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CU(cudaStreamSynchronize((cudaStream_t)stream));
    const int64_t have = std::min<int64_t>(s[1], HB_FLAGGED_CAP);
    *n = std::min(have, cap);
    if (*n > 0) {
        CU(cudaMemcpyAsync(chains, ws->flagged.p, (size_t)*n * sizeof(int64_t), cudaMemcpyDeviceToHost,
                           (cudaStream_t)stream));
        CU(cudaStreamSynchronize((cudaStream_t)stream));
    }
    return HB_OK;
}

int hb_last_host_io(int64_t *h2d_bytes, int64_t *d2h_bytes, int64_t *rebuilt_bytes)
{
    std::lock_guard<std::mutex> lk(g_mu);
    Workspace *ws;
    HBTRY(cur_ws(&ws));
    if (h2d_bytes) *h2d_bytes = ws->io_h2d;
    if (d2h_bytes) *d2h_bytes = ws->io_d2h;
    if (rebuilt_bytes) *rebuilt_bytes = ws->io_rebuilt;
    return HB_OK;
}

int hb_status_dev(void *stream)
{
    std::lock_guard<std::mutex> lk(g_mu);
    Workspace *ws;
    HBTRY(cur_ws(&ws));
    WS_USE(stream);
    if (!ws->status.p) return HB_OK;
    return read_status(ws, (cudaStream_t)stream);
}

// ------------------------------------------------------------------ Gaussian chains
int hb_hmm_norm_dev(const double *x, int64_t ld, int64_t T, int64_t B, const int64_t *seg_off,
                    int32_t nseg, const double *mean, const double *sd, int32_t sd_per_seg,
                    const double *Pi, const double *delta, int32_t flags, uint8_t *y, int64_t ldy,
                    double *gamma, int64_t ldg, double *ll, void *stream)
{
    std::lock_guard<std::mutex> lk(g_mu);
    Workspace *ws;
    HBTRY(cur_ws(&ws));
    WS_USE(stream);
    return norm3_run(ws, x, ld, T, B, seg_off, nseg, mean, sd, nullptr, nullptr, sd_per_seg, Pi,
                     delta, flags, y, ldy, gamma, ldg, ll, (cudaStream_t)stream);
}

// Chain-resident posteriors on CHAIN-MAJOR data (hb_norm3_resident.cuh): x[b*T + t] (R's column-major
// [T x B]), gamma[(k*B + b)*T + t] (R's [T x B x 3]).  One launch per segment, longest first.
int hb_forwardback_norm_cm_dev(const double *x, int64_t T, int64_t B, const int64_t *seg_off, int32_t nseg,
                               const double *mean, const double *sd, int32_t sd_per_seg, const double *Pi,
                               const double *delta, double *gamma, void *stream)
{
    std::lock_guard<std::mutex> lk(g_mu);
    Workspace *ws;
    HBTRY(cur_ws(&ws));
    // no WS_USE: this path touches none of the shared scratch (its only device state is its own gate
    // word), so it may run on another stream BESIDE a chain launch — e.g. next to the Viterbi-only kernel
    cudaStream_t st = (cudaStream_t)stream;
    if (!x || !sd || !gamma || T < 1 || B < 1) return fail(HB_ERR_ARG, "bad arguments");
    const int64_t one_seg[2] = {0, T};
    if (!seg_off) {
        seg_off = one_seg;
        nseg = 1;
    }
    if (nseg < 1 || seg_off[0] < 0 || seg_off[nseg] > T) return fail(HB_ERR_ARG, "seg_off outside [0, T]");
    hb::FbrArgs a;
    memset(&a, 0, sizeof a);
    bool sym;
    HBTRY(fill_norm3_const(a.c, mean, Pi, delta, &sym));
    if (!sym || !(Pi[1] > 0)) return fail(HB_ERR_ARG, "the chain-resident kernel serves the symmetric parameter block only");
    HBTRY(ws->fbr_gate.ensure(sizeof(int32_t)));
    CU(cudaMemsetAsync(ws->fbr_gate.p, 0, sizeof(int32_t), st));
    a.x = x;
    a.T = T;
    a.B = B;
    a.sd = sd;
    a.sd_per_seg = sd_per_seg ? 1 : 0;
    a.gamma = gamma;
    a.gate = ws->fbr_gate.as<int32_t>();
    std::vector<int32_t> order(nseg);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int32_t p, int32_t q) {
        return (seg_off[p + 1] - seg_off[p]) > (seg_off[q + 1] - seg_off[q]);
    });
    for (int32_t s : order) {
        const int64_t L = seg_off[s + 1] - seg_off[s];
        if (L < 0) return fail(HB_ERR_ARG, "seg_off must be non-decreasing");
        if (L == 0) continue;
        // lanes per chain: pieces of <= ~32 positions while a block (<= 4 warps) can hold the chain
        int nw = L <= 1024 ? 1 : (L <= 2048 ? 2 : 4);
        if (const char *e = getenv("HB_FBR_NW")) // experiment knob: warps per chain (1, 2, 4)
            if (atoi(e) == 1 || atoi(e) == 2 || atoi(e) == 4) nw = atoi(e);
        int m = 0;
        const size_t smem = hb::fbr_smem_bytes((int)L, nw, &m);
        if (smem > 220 * 1024)
            return fail(HB_ERR_ARG, "segment of %lld positions does not fit the chain-resident kernel (<= %d)",
                        (long long)L, 220 * 1024 / 36);
        a.t0 = seg_off[s];
        a.L = (int32_t)L;
        a.m = m;
        a.seg = s;
        CU(hb::launch_norm3_resident(a, nw, smem, st));
    }
    return HB_OK;
}

int hb_fbr_stats_dev(void *stream, int32_t *gate)
{
    std::lock_guard<std::mutex> lk(g_mu);
    Workspace *ws;
    HBTRY(cur_ws(&ws));
    if (!gate) return fail(HB_ERR_ARG, "gate is NULL");
    *gate = 0;
    if (!ws->fbr_gate.p) return HB_OK;
    CU(cudaMemcpyAsync(gate, ws->fbr_gate.p, sizeof(int32_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CU(cudaStreamSynchronize((cudaStream_t)stream));
    return HB_OK;
}

// Pipeline of the host-buffer entry points.  Columns are cut into sub-batches; three streams work on
// consecutive sub-batches at once: s_in uploads sub-batch i+1 (H2D), s_cmp changes the layout and
// runs the chains of sub-batch i, s_out downloads sub-batch i-1 (D2H) — PCIe is full duplex, so the
// call approaches max(H2D, D2H) instead of their sum.  Staging buffers are double-buffered and
// handed over with events; with pageable host memory the copies degrade to the driver's staged
// path (still correct, less overlap).
// ---------------------------------------------------------------------------------------------
// Host memory that is not page-locked (every R vector) cannot be the source or target of an
// asynchronous copy: cudaMemcpyAsync silently degrades to a synchronous, staged transfer (measured:
// 15 GB/s instead of 67 GB/s, no overlap).  HostStager routes such buffers through the library's own
// pinned bounce buffers: the upload side copies user memory -> pinned with several host threads right
// before the H2D of a sub-batch, the download side lands D2H in pinned memory and a background task
// copies it out to the user's buffers while the GPU works on the next sub-batch.  Pinned user
// buffers (bench.py) take the direct path.
static bool host_is_pinned(const void *p)
{
    if (!p) return true;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}

static void parallel_memcpy(void *dst, const void *src, size_t bytes, int nthreads)
{
    if (bytes < ((size_t)4 << 20) || nthreads <= 1) {
        memcpy(dst, src, bytes);
        return;
    }
    std::vector<std::thread> th;
    const size_t per = (bytes / nthreads + 4095) & ~(size_t)4095;
    for (int i = 0; i < nthreads; i++) {
        const size_t a = (size_t)i * per;
        if (a >= bytes) break;
        const size_t n = std::min(per, bytes - a);
        th.emplace_back([=] { memcpy((char *)dst + a, (const char *)src + a, n); });
    }
    for (auto &t : th) t.join();
}

static void parallel_widen(int32_t *dst, const uint8_t *src, size_t count, int nthreads)
{
    auto body = [=](size_t a, size_t b) {
        for (size_t i = a; i < b; i++) dst[i] = src[i];
    };
    if (count < ((size_t)1 << 20) || nthreads <= 1) {
        body(0, count);
        return;
    }
    std::vector<std::thread> th;
    const size_t per = (count / nthreads + 4095) & ~(size_t)4095;
    for (int i = 0; i < nthreads; i++) {
        const size_t a = (size_t)i * per;
        if (a >= count) break;
        th.emplace_back(body, a, std::min(count, a + per));
    }
    for (auto &t : th) t.join();
}

// The neutral-state posterior plane does not cross PCIe: the chain kernels define it as
// gamma_1 = |1 - (gamma_0 + gamma_2)| (3 states) / gamma_1 = |1 - gamma_0| (2 states) — the host rebuilds it
// with the same two IEEE operations from the planes that did come down (streaming stores: the plane is
// write-only here).  a, c: the downloaded planes (c may be NULL), out: the rebuilt one.
static void recon_range(const double *a, const double *c, double *out, size_t n)
{
    size_t i = 0;
    auto one = [&](size_t k) { out[k] = fabs(1.0 - (c ? a[k] + c[k] : a[k])); };
    while (i < n && ((uintptr_t)(out + i) & 15)) one(i++);
    const __m128d ones = _mm_set1_pd(1.0), sign = _mm_set1_pd(-0.0);
    for (; i + 2 <= n; i += 2) {
        __m128d s = _mm_loadu_pd(a + i);
        if (c) s = _mm_add_pd(s, _mm_loadu_pd(c + i));
        _mm_stream_pd(out + i, _mm_andnot_pd(sign, _mm_sub_pd(ones, s)));
    }
    for (; i < n; i++) one(i);
    _mm_sfence();
}

static void parallel_recon(const double *a, const double *c, double *out, size_t count, int nthreads)
{
    if (count < ((size_t)1 << 18) || nthreads <= 1) {
        recon_range(a, c, out, count);
        return;
    }
    std::vector<std::thread> th;
    const size_t per = (count / nthreads + 511) & ~(size_t)511;
    for (int i = 0; i < nthreads; i++) {
        const size_t lo = (size_t)i * per;
        if (lo >= count) break;
        const size_t n = std::min(per, count - lo);
        th.emplace_back([=] { recon_range(a + lo, c ? c + lo : nullptr, out + lo, n); });
    }
    for (auto &t : th) t.join();
}

// HB_E2E_FULL_PLANES=1: download every posterior plane instead (A/B measurements);
// HB_E2E_ADAPTIVE=1 switches HostStager::plan()'s balancing on (off by default: every sub-batch rebuilds the
// plane).  Measured with HB_PIPE_TRACE on a 55 GB/s box, 10 k x 20 k: the copy-out threads need 58 ms per pass
// (0.07 + 0.22 ns per element), the downloads 79-87 ms at the 39-44 GB/s the link delivers while uploads and the
// host threads run beside it — the call is link-bound with the plane rebuilt, so there is nothing to balance,
// and with four slots in flight the per-job timings plan() relies on are inflated by the jobs overlapping
// (100.9 ms with balancing vs 90.1 ms without).  A first heuristic ("download the plane when the slot's previous
// copy-out job is still running") was worse still: the enqueueing thread runs ahead of the GPU, so that test is
// nearly always true.
static bool recon_enabled()
{
    const char *e = getenv("HB_E2E_FULL_PLANES");
    return !(e && atoi(e) != 0);
}
static bool recon_adaptive()
{
    const char *e = getenv("HB_E2E_ADAPTIVE");
    return e && atoi(e) != 0;
}

struct HostStager {
    Workspace *ws;
    bool bounce_in, bounce_out;
    int nthreads;
    struct Copy {
        const void *src;
        void *dst;
        size_t bytes;
    };
    struct Widen {
        const uint8_t *src;
        int32_t *dst;
        size_t count;
    };
    struct Recon {
        const double *a, *c;
        double *out;
        size_t count;
    };
    std::vector<Copy> pend[HB_PIPE_SLOTS];
    std::vector<Widen> pend_w[HB_PIPE_SLOTS];
    std::vector<Recon> pend_r[HB_PIPE_SLOTS];
    std::future<void> fut[HB_PIPE_SLOTS];
    size_t off_out[HB_PIPE_SLOTS] = {}, off_in[HB_PIPE_SLOTS] = {};
    // ---- balance between the download link and the copy-out threads (plan()): per sub-batch the neutral
    // posterior plane is either rebuilt by the host threads (less PCIe, more host work) or downloaded (the
    // reverse); both sides are timed — CUDA events around the sub-batch's D2H copies, a clock inside the
    // copy-out job — and each sub-batch takes the option that keeps the two cumulative loads level.
    cudaEvent_t ev_t0[HB_PIPE_SLOTS] = {}, ev_t1[HB_PIPE_SLOTS] = {};
    double job_other_s[HB_PIPE_SLOTS] = {}, job_recon_s[HB_PIPE_SLOTS] = {};
    size_t slot_bytes[HB_PIPE_SLOTS] = {}, slot_elems[HB_PIPE_SLOTS] = {};
    bool slot_recon[HB_PIPE_SLOTS] = {}, slot_used[HB_PIPE_SLOTS] = {};
    double s_per_byte = 0, other_per_elem = 0, recon_per_elem = 0, cumD = 0, cumH = 0;
    int n_link = 0, n_other = 0, n_recon = 0;
    bool balance = false;

    // Decide for sub-batch i (slot k, `elems` = positions x chains, S states): true = rebuild on the host.
    bool plan(int k, size_t elems, int S, bool recon_allowed)
    {
        if (slot_used[k]) { // harvest the measurements of this slot's previous sub-batch
            if (fut[k].valid()) fut[k].get();
            float ms = 0;
            if (ev_t0[k] && cudaEventElapsedTime(&ms, ev_t0[k], ev_t1[k]) == cudaSuccess && ms > 0 && slot_bytes[k]) {
                const double v = ms * 1e-3 / (double)slot_bytes[k];
                s_per_byte = n_link ? 0.7 * s_per_byte + 0.3 * v : v;
                n_link++;
            } else {
                cudaGetLastError();
            }
            if (slot_elems[k]) {
                const double o = job_other_s[k] / (double)slot_elems[k];
                other_per_elem = n_other ? 0.7 * other_per_elem + 0.3 * o : o;
                n_other++;
                if (slot_recon[k]) {
                    const double r = job_recon_s[k] / (double)slot_elems[k];
                    recon_per_elem = n_recon ? 0.7 * recon_per_elem + 0.3 * r : r;
                    n_recon++;
                }
            }
        }
        bool recon = recon_allowed;
        if (recon_allowed && balance && n_link >= 1 && n_recon >= 1) {
            const double b_r = (double)elems * (1 + 8.0 * (S - 1)), b_f = (double)elems * (1 + 8.0 * S);
            const double Dr = b_r * s_per_byte, Df = b_f * s_per_byte;
            const double Hr = (double)elems * (other_per_elem + recon_per_elem), Hf = (double)elems * other_per_elem;
            recon = std::max(cumD + Dr, cumH + Hr) <= std::max(cumD + Df, cumH + Hf);
            cumD += recon ? Dr : Df;
            cumH += recon ? Hr : Hf;
            const double base = std::min(cumD, cumH); // only the difference matters
            cumD -= base;
            cumH -= base;
        }
        slot_used[k] = true;
        slot_recon[k] = recon;
        slot_elems[k] = elems;
        slot_bytes[k] = 0;
        return recon;
    }
    int link_begin(int k, cudaStream_t st)
    {
        if (!ev_t0[k]) {
            CU(cudaEventCreate(&ev_t0[k]));
            CU(cudaEventCreate(&ev_t1[k]));
        }
        CU(cudaEventRecord(ev_t0[k], st));
        return HB_OK;
    }
    int link_end(int k, cudaStream_t st)
    {
        CU(cudaEventRecord(ev_t1[k], st));
        return HB_OK;
    }
    void trace(const char *what, int64_t T, int64_t B, double wall_s)
    {
        if (!getenv("HB_PIPE_TRACE")) return;
        fprintf(stderr, "[hb pipe] %s T=%lld B=%lld wall %.1f ms | link %.1f GB/s (%d obs) | copy-out other %.2f ns/elem, "
                        "rebuild %.2f ns/elem (%d threads) | d2h %.0f MB rebuilt %.0f MB\n",
                what, (long long)T, (long long)B, wall_s * 1e3, s_per_byte > 0 ? 1e-9 / s_per_byte : 0.0, n_link,
                other_per_elem * 1e9, recon_per_elem * 1e9, nthreads, ws->io_d2h / 1e6, ws->io_rebuilt / 1e6);
    }
    ~HostStager()
    {
        for (int k = 0; k < HB_PIPE_SLOTS; k++) {
            if (fut[k].valid()) fut[k].get();
            if (ev_t0[k]) cudaEventDestroy(ev_t0[k]);
            if (ev_t1[k]) cudaEventDestroy(ev_t1[k]);
        }
    }

    int init(Workspace *w, bool b_in, bool b_out, size_t in_bytes, size_t out_bytes, size_t y_count = 0)
    {
        ws = w;
        w->io_h2d = w->io_d2h = w->io_rebuilt = 0;
        if (y_count && w->bounce_y_cap < y_count) {
            for (int k = 0; k < HB_PIPE_SLOTS; k++) {
                if (w->bounce_y[k]) cudaFreeHost(w->bounce_y[k]);
                w->bounce_y[k] = nullptr;
                CU(cudaHostAlloc(&w->bounce_y[k], y_count, cudaHostAllocDefault));
            }
            w->bounce_y_cap = y_count;
        }
        bounce_in = b_in;
        bounce_out = b_out;
        // copy-out / widen / rebuild threads: the rebuilt posterior plane makes the download side host-bound on
        // fast PCIe (measured 16 threads 2.22 vs 8 threads 2.05 G steps/s on a 16-vCPU box): all but two cores
        {
            const unsigned hw = std::thread::hardware_concurrency();
            nthreads = g_host_threads > 0 ? g_host_threads
                                          : (int)std::max(1u, std::min(16u, hw > 4 ? hw - 2 : hw / 2));
        }
        if (const char *e = getenv("HB_STAGER_THREADS")) nthreads = std::max(1, atoi(e));
        for (int k = 0; k < HB_PIPE_SLOTS; k++) {
            if (b_in && ws->bounce_in_cap < in_bytes) {
                if (ws->bounce_in[k]) cudaFreeHost(ws->bounce_in[k]);
                ws->bounce_in[k] = nullptr;
                CU(cudaHostAlloc(&ws->bounce_in[k], in_bytes, cudaHostAllocDefault));
            }
            if (b_out && ws->bounce_out_cap < out_bytes) {
                if (ws->bounce_out[k]) cudaFreeHost(ws->bounce_out[k]);
                ws->bounce_out[k] = nullptr;
                CU(cudaHostAlloc(&ws->bounce_out[k], out_bytes, cudaHostAllocDefault));
            }
            if (!ws->ev_bin[k]) CU(cudaEventCreateWithFlags(&ws->ev_bin[k], cudaEventDisableTiming));
            if (!ws->ev_bout[k]) CU(cudaEventCreateWithFlags(&ws->ev_bout[k], cudaEventDisableTiming));
        }
        if (b_in) ws->bounce_in_cap = std::max(ws->bounce_in_cap, in_bytes);
        if (b_out) ws->bounce_out_cap = std::max(ws->bounce_out_cap, out_bytes);
        return HB_OK;
    }
    // slot k is about to be filled again: its previous H2D must have left the bounce buffer
    int begin_in(int k, bool reused)
    {
        off_in[k] = 0;
        if (bounce_in && reused) CU(cudaEventSynchronize(ws->ev_bin[k]));
        return HB_OK;
    }
    int h2d(int k, void *dev, const void *host, size_t bytes, cudaStream_t st)
    {
        ws->io_h2d += (int64_t)bytes;
        const void *src = host;
        if (bounce_in) {
            char *b = (char *)ws->bounce_in[k] + off_in[k];
            parallel_memcpy(b, host, bytes, nthreads);
            off_in[k] += (bytes + 255) & ~(size_t)255;
            src = b;
        }
        CU(cudaMemcpyAsync(dev, src, bytes, cudaMemcpyHostToDevice, st));
        return HB_OK;
    }
    int end_in(int k, cudaStream_t st)
    {
        if (bounce_in) CU(cudaEventRecord(ws->ev_bin[k], st));
        return HB_OK;
    }
    // slot k's previous download must have been copied out to the user before D2H lands in it again
    void begin_out(int k)
    {
        if (fut[k].valid()) fut[k].get();
        pend[k].clear();
        pend_w[k].clear();
        pend_r[k].clear();
        off_out[k] = 0;
    }
    // Viterbi states: one byte per position over PCIe (a quarter of R's integers), landed in pinned memory
    // and widened into the caller's int32 array by the background task while the GPU works on
    int d2h_widen(int k, int32_t *host, const uint8_t *dev, size_t count, cudaStream_t st)
    {
        ws->io_d2h += (int64_t)count;
        slot_bytes[k] += count;
        CU(cudaMemcpyAsync(ws->bounce_y[k], dev, count, cudaMemcpyDeviceToHost, st));
        pend_w[k].push_back({(const uint8_t *)ws->bounce_y[k], host, count});
        return HB_OK;
    }
    int d2h(int k, void *host, const void *dev, size_t bytes, cudaStream_t st)
    {
        ws->io_d2h += (int64_t)bytes;
        slot_bytes[k] += bytes;
        void *dst = host;
        if (bounce_out) {
            char *b = (char *)ws->bounce_out[k] + off_out[k];
            off_out[k] += (bytes + 255) & ~(size_t)255;
            pend[k].push_back({b, host, bytes});
            dst = b;
        }
        CU(cudaMemcpyAsync(dst, dev, bytes, cudaMemcpyDeviceToHost, st));
        return HB_OK;
    }
    // after this slot's downloads (and copy-outs): out = |1 - (a + c)| on the caller's arrays
    void recon(int k, const double *a, const double *c, double *out, size_t count)
    {
        pend_r[k].push_back({a, c, out, count});
    }
    int end_out(int k, cudaStream_t st)
    {
        if (pend[k].empty() && pend_w[k].empty() && pend_r[k].empty()) return HB_OK;
        CU(cudaEventRecord(ws->ev_bout[k], st));
        cudaEvent_t ev = ws->ev_bout[k];
        std::vector<Copy> jobs = pend[k];
        std::vector<Widen> wjobs = pend_w[k];
        std::vector<Recon> rjobs = pend_r[k];
        const int nt = nthreads;
        int dev = 0;
        cudaGetDevice(&dev);
        double *t_other = &job_other_s[k], *t_recon = &job_recon_s[k];
        fut[k] = std::async(std::launch::async, [ev, jobs, wjobs, rjobs, nt, dev, t_other, t_recon] {
            cudaSetDevice(dev);
            cudaEventSynchronize(ev);
            const auto c0 = std::chrono::steady_clock::now();
            for (const Copy &c : jobs) parallel_memcpy(c.dst, c.src, c.bytes, nt);
            for (const Widen &w : wjobs) parallel_widen(w.dst, w.src, w.count, nt);
            const auto c1 = std::chrono::steady_clock::now();
            for (const Recon &r : rjobs) parallel_recon(r.a, r.c, r.out, r.count, nt);
            const auto c2 = std::chrono::steady_clock::now();
            *t_other = std::chrono::duration<double>(c1 - c0).count();
            *t_recon = std::chrono::duration<double>(c2 - c1).count();
        });
        return HB_OK;
    }
    void finish()
    {
        for (int k = 0; k < HB_PIPE_SLOTS; k++)
            if (fut[k].valid()) fut[k].get();
    }
};

static int pipe_init(Workspace *ws)
{
    if (ws->s_in) return HB_OK;
    CU(cudaStreamCreateWithFlags(&ws->s_in, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&ws->s_cmp, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&ws->s_out, cudaStreamNonBlocking));
    for (int k = 0; k < HB_PIPE_SLOTS; k++) {
        CU(cudaEventCreateWithFlags(&ws->ev_in[k], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ws->ev_free_in[k], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ws->ev_out[k], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ws->ev_free_out[k], cudaEventDisableTiming));
    }
    return HB_OK;
}

// Sub-batch width: ~8+ sub-batches per call when B allows, a multiple of 32 columns, bounded by
// the free device memory (`bytes_per_col` covers staging, resident copy, outputs and scratch).
static int64_t pipe_batch(int64_t B, double bytes_per_col, int64_t want_cap = 1024)
{
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) {
        cudaGetLastError();
        free_b = (size_t)8 << 30;
    }
    int64_t cap = std::max<int64_t>(1, (int64_t)(free_b * 0.6 / bytes_per_col));
    int64_t want = std::max<int64_t>(32, (B / 8 + 31) / 32 * 32);
    want = std::min<int64_t>(want, want_cap);
    if (const char *e = getenv("HB_PIPE_COLS")) // tuning knob: columns per sub-batch of the host pipeline
        if (atoi(e) >= 32) want = atoi(e) / 32 * 32;
    int64_t Bc = std::min(std::min(cap, want), B);
    if (Bc >= 32) Bc = Bc / 32 * 32;
    return std::max<int64_t>(1, Bc);
}

int hb_hmm_norm_tiled_dev(const double *x, int64_t T, int64_t B, const int64_t *seg_off, int32_t nseg,
                          const double *mean, const double *sd, int32_t sd_per_seg, const double *Pi,
                          const double *delta, int32_t flags, uint8_t *y, double *gamma, double *ll,
                          void *stream)
{
    std::lock_guard<std::mutex> lk(g_mu);
    Workspace *ws;
    HBTRY(cur_ws(&ws));
    WS_USE(stream);
    // tiles are moved in 16-byte pieces and the states of a chunk as one 8-byte store per lane
    if (((uintptr_t)x & 15) || ((uintptr_t)y & 7)) return fail(HB_ERR_ARG, "tiled buffers must be 16-byte aligned");
    return norm3_run(ws, x, 32, T, B, seg_off, nseg, mean, sd, nullptr, nullptr, sd_per_seg, Pi, delta,
                     flags, y, 32, gamma, 32, ll, (cudaStream_t)stream, true, true);
}

int hb_hmm_norm_host(const double *x, int64_t T, int64_t B, const int64_t *seg_off, int32_t nseg,
                     const double *mean, const double *sd, int32_t sd_per_seg, const double *Pi,
                     const double *delta, int32_t flags, int32_t *y, double *gamma, double *ll)
{
    std::lock_guard<std::mutex> lk(g_mu);
    Workspace *ws;
    HBTRY(cur_ws(&ws));
    if (T < 1 || B < 1 || !x || !sd) return fail(HB_ERR_ARG, "bad arguments");
    const int64_t one_seg[2] = {0, T};
    if (!seg_off) {
        seg_off = one_seg;
        nseg = 1;
    }
    bool vit, fb;
    HBTRY(check_flags(flags, y, gamma, ll, &vit, &fb));
    const bool want_g = (flags & HB_WANT_GAMMA) != 0, want_ll = (flags & HB_WANT_LL) != 0;
    WsUse ws_use_;
    if (long_chain_shape(T, B, nseg, flags) && seg_off[0] == 0 && seg_off[1] == T) {
        HBTRY(ws_use_.enter(ws, 0));
        // the drop-in seam itself: Viterbi(dthmm(x, ...)) on one (or a few) long pooled vectors.  R's
        // column-major [T x B] is already chain-major: emissions in parallel, chains along the gene axis.
        hb::Norm3Const nc;
        bool sym;
        memset(&nc, 0, sizeof nc);
        HBTRY(fill_norm3_const(nc, mean, Pi, delta, &sym));
        std::vector<double> h_aux(3 * (size_t)B);
        for (int64_t b = 0; b < B; b++) {
            h_aux[b] = sd[b];
            h_aux[B + b] = 1.0 / sd[b];
            h_aux[2 * B + b] = log(sd[b]);
        }
        HBTRY(ws->st_in.ensure((size_t)T * B * sizeof(double)));
        HBTRY(ws->sdaux.ensure(h_aux.size() * sizeof(double)));
        HBTRY(ws->status.ensure(4 * sizeof(int32_t)));
        CU(cudaMemsetAsync(ws->status.p, 0, 4 * sizeof(int32_t), 0));
        HBTRY(ws->lc_lb.ensure((size_t)T * B * 3 * sizeof(double)));
        CU(cudaMemcpyAsync(ws->st_in.p, x, (size_t)T * B * sizeof(double), cudaMemcpyHostToDevice, 0));
        CU(cudaMemcpyAsync(ws->sdaux.p, h_aux.data(), h_aux.size() * sizeof(double), cudaMemcpyHostToDevice, 0));
        const double *aux = ws->sdaux.as<double>();
        CU(hb::launch_norm3_emit_cm(nc, ws->st_in.as<double>(), T, (int32_t)B, aux, aux + B, aux + 2 * B,
                                    ws->lc_lb.as<double>(), 0));
        return longchain_host_outputs(ws, 3, T, B, Pi, delta, flags, y, gamma, ll, false);
    }
    HBTRY(pipe_init(ws));
    HBTRY(ws_use_.enter(ws, ws->s_cmp));
    const bool recon_all = want_g && recon_enabled();
    // <= 512 columns per sub-batch: the pipeline's fill (first upload) and drain (last download + copy-out) are
    // one sub-batch each and are not overlapped with anything
    const int64_t Bc = pipe_batch(B, 200.0 * (double)T, T >= 4096 ? 512 : 1024);
    const int64_t nb = (B + Bc - 1) / Bc;
    const int64_t ldc = (Bc + 127) / 128 * 128;
    const int64_t nsd_seg = sd_per_seg ? nseg : 1;
    // every allocation before the first enqueue (cudaMalloc / cudaFree synchronise the device)
    for (int k = 0; k < HB_PIPE_SLOTS; k++) {
        HBTRY(ws->pin[k].ensure((size_t)T * Bc * sizeof(double)));
        if (vit) HBTRY(ws->pout_y[k].ensure((size_t)T * Bc));
        if (want_g) HBTRY(ws->pout_g[k].ensure((size_t)3 * T * Bc * sizeof(double)));
        if (want_ll) HBTRY(ws->pout_ll[k].ensure((size_t)nseg * Bc * sizeof(double)));
    }
    HBTRY(ws->st_x.ensure((size_t)T * ldc * sizeof(double)));
    if (vit) HBTRY(ws->st_y.ensure((size_t)T * ldc));
    if (want_g) HBTRY(ws->st_g.ensure((size_t)3 * T * ldc * sizeof(double)));
    HBTRY(ws->st_sd.ensure((size_t)nsd_seg * B * 3 * sizeof(double)));
    HostStager hs;
    hs.balance = recon_adaptive();
    const auto wall0 = std::chrono::steady_clock::now();
    HBTRY(hs.init(ws, !host_is_pinned(x),
                  !(host_is_pinned(want_g ? gamma : nullptr) && host_is_pinned(want_ll ? ll : nullptr)),
                  (size_t)T * Bc * sizeof(double) + 256,
                  (want_g ? 3 * ((size_t)T * Bc * sizeof(double) + 256) : 0) +
                      (want_ll ? (size_t)nseg * (Bc * sizeof(double) + 256) : 0) + 256,
                  vit ? (size_t)T * Bc : 0));
    // chain scales of all sub-batches in one upload: sd, 1/sd (IEEE), log(sd) (host libm, as R);
    // sub-batch i occupies [3 * nsd_seg * b0, +3 * nsd_seg * bn) as three [nsd_seg][bn] blocks
    std::vector<double> h_aux((size_t)nsd_seg * B * 3);
    for (int64_t b0 = 0; b0 < B; b0 += Bc) {
        const int64_t bn = std::min(Bc, B - b0);
        double *blk = h_aux.data() + (size_t)3 * nsd_seg * b0;
        for (int64_t sg = 0; sg < nsd_seg; sg++)
            for (int64_t b = 0; b < bn; b++) {
                const double v = sd[sg * B + b0 + b];
                blk[(0 * nsd_seg + sg) * bn + b] = v;
                blk[(1 * nsd_seg + sg) * bn + b] = 1.0 / v;
                blk[(2 * nsd_seg + sg) * bn + b] = log(v);
            }
    }
    CU(cudaMemcpyAsync(ws->st_sd.p, h_aux.data(), h_aux.size() * sizeof(double), cudaMemcpyHostToDevice,
                       ws->s_cmp));
    int rc = HB_OK;
    for (int64_t i = 0; i < nb && rc == HB_OK; i++) {
        const int k = (int)(i % HB_PIPE_SLOTS);
        const int64_t b0 = i * Bc, bn = std::min(Bc, B - b0);
        const int64_t ld = (bn + 127) / 128 * 128;
        // neutral plane: rebuilt on the host or downloaded, whichever keeps link and host threads level (plan())
        const bool recon = want_g ? hs.plan(k, (size_t)T * bn, 3, recon_all) : false;
        // upload
        if (i >= HB_PIPE_SLOTS) CU(cudaStreamWaitEvent(ws->s_in, ws->ev_free_in[k], 0));
        HBTRY(hs.begin_in(k, i >= HB_PIPE_SLOTS));
        HBTRY(hs.h2d(k, ws->pin[k].p, x + b0 * T, (size_t)T * bn * sizeof(double), ws->s_in));
        HBTRY(hs.end_in(k, ws->s_in));
        CU(cudaEventRecord(ws->ev_in[k], ws->s_in));
        // layout + chains
        CU(cudaStreamWaitEvent(ws->s_cmp, ws->ev_in[k], 0));
        CU(hb::launch_transpose_f64(ws->pin[k].as<double>(), T, bn, ws->st_x.as<double>(), ld, ws->s_cmp));
        CU(cudaEventRecord(ws->ev_free_in[k], ws->s_cmp));
        const double *d_sd = ws->st_sd.as<double>() + (size_t)3 * nsd_seg * b0;
        uint8_t *d_y = vit ? ws->st_y.as<uint8_t>() : nullptr;
        double *d_g = want_g ? ws->st_g.as<double>() : nullptr;
        double *d_ll = want_ll ? ws->pout_ll[k].as<double>() : nullptr;
        if (i >= HB_PIPE_SLOTS) CU(cudaStreamWaitEvent(ws->s_cmp, ws->ev_free_out[k], 0));
        rc = norm3_run(ws, ws->st_x.as<double>(), ld, T, bn, seg_off, nseg, mean, d_sd,
                       d_sd + nsd_seg * bn, d_sd + 2 * nsd_seg * bn, sd_per_seg, Pi, delta, flags, d_y,
                       ld, d_g, ld, d_ll, ws->s_cmp, i == 0);
        if (rc != HB_OK) break;
        if (vit)
            CU(hb::launch_transpose_u8_back_u8(d_y, ld, T, bn, ws->pout_y[k].as<uint8_t>(), ws->s_cmp));
        if (want_g)
            for (int pl = 0; pl < 3; pl++) {
                if (recon && pl == 1) continue; // the neutral-state plane is rebuilt on the host
                CU(hb::launch_transpose_f64_back(d_g + (int64_t)pl * T * ld, ld, T, bn,
                                                 ws->pout_g[k].as<double>() + (int64_t)pl * T * bn,
                                                 ws->s_cmp));
            }
        CU(cudaEventRecord(ws->ev_out[k], ws->s_cmp));
        // download
        CU(cudaStreamWaitEvent(ws->s_out, ws->ev_out[k], 0));
        hs.begin_out(k);
        HBTRY(hs.link_begin(k, ws->s_out));
        if (vit)
            HBTRY(hs.d2h_widen(k, y + b0 * T, ws->pout_y[k].as<uint8_t>(), (size_t)T * bn, ws->s_out));
        if (want_g) {
            for (int pl = 0; pl < 3; pl++) {
                if (recon && pl == 1) continue;
                HBTRY(hs.d2h(k, gamma + ((int64_t)pl * B + b0) * T,
                             ws->pout_g[k].as<double>() + (int64_t)pl * T * bn,
                             (size_t)T * bn * sizeof(double), ws->s_out));
            }
            if (recon) {
                hs.recon(k, gamma + b0 * T, gamma + (2 * B + b0) * T, gamma + (B + b0) * T, (size_t)T * bn);
                ws->io_rebuilt += (int64_t)T * bn * (int64_t)sizeof(double);
            }
        }
        if (want_ll)
            for (int32_t sg = 0; sg < nseg; sg++)
                HBTRY(hs.d2h(k, ll + (int64_t)sg * B + b0, d_ll + (int64_t)sg * bn,
                             (size_t)bn * sizeof(double), ws->s_out));
        HBTRY(hs.link_end(k, ws->s_out));
        CU(cudaEventRecord(ws->ev_free_out[k], ws->s_out));
        HBTRY(hs.end_out(k, ws->s_out));
    }
    // drain (also on the error path: host buffers must not be touched after we return)
    cudaStreamSynchronize(ws->s_in);
    cudaStreamSynchronize(ws->s_cmp);
    cudaError_t e = cudaStreamSynchronize(ws->s_out);
    hs.finish();
    hs.trace("host pipeline", T, B, std::chrono::duration<double>(std::chrono::steady_clock::now() - wall0).count());
    if (rc != HB_OK) return rc;
    if (e != cudaSuccess) return fail(HB_ERR_CUDA, "pipeline failed: %s", cudaGetErrorString(e));
    return read_status(ws, ws->s_cmp);
}

// ------------------------------------------------------------------ binomial chains
int hb_hmm_binom_dev(const int32_t *x, const int32_t *size, int64_t ld, int64_t T, int64_t B,
                     const int64_t *seg_off, int32_t nseg, const double *prob, double rho,
                     const double *Pi, const double *delta, int32_t flags, uint8_t *y,
                     int64_t ldy, double *gamma, int64_t ldg, double *ll, void *stream)
{
    std::lock_guard<std::mutex> lk(g_mu);
    Workspace *ws;
    HBTRY(cur_ws(&ws));
    WS_USE(stream);
    return binom2_run(ws, x, size, ld, T, B, seg_off, nseg, prob, rho, Pi, delta, flags, y, ldy,
                      gamma, ldg, ll, (cudaStream_t)stream);
}

int hb_hmm_binom_tiled_dev(const int32_t *x, const int32_t *size, int64_t T, int64_t B,
                           const int64_t *seg_off, int32_t nseg, const double *prob, double rho,
                           const double *Pi, const double *delta, int32_t flags, uint8_t *y,
                           double *gamma, double *ll, void *stream)
{
    std::lock_guard<std::mutex> lk(g_mu);
    Workspace *ws;
    HBTRY(cur_ws(&ws));
    WS_USE(stream);
    if (((uintptr_t)x & 15) || ((uintptr_t)size & 15)) return fail(HB_ERR_ARG, "tiled buffers must be 16-byte aligned");
    return binom2_run(ws, x, size, 32, T, B, seg_off, nseg, prob, rho, Pi, delta, flags, y, 32, gamma,
                      32, ll, (cudaStream_t)stream, true, false, true);
}

int hb_hmm_binom_packed_tiled_dev(const uint32_t *xn, int64_t T, int64_t B, const int64_t *seg_off,
                                  int32_t nseg, const double *prob, double rho, const double *Pi,
                                  const double *delta, int32_t flags, uint8_t *y, double *gamma,
                                  double *ll, void *stream)
{
    std::lock_guard<std::mutex> lk(g_mu);
    Workspace *ws;
    HBTRY(cur_ws(&ws));
    WS_USE(stream);
    if ((uintptr_t)xn & 15) return fail(HB_ERR_ARG, "tiled buffers must be 16-byte aligned");
    return binom2_run(ws, (const int32_t *)xn, nullptr, 32, T, B, seg_off, nseg, prob, rho, Pi, delta, flags,
                      y, 32, gamma, 32, ll, (cudaStream_t)stream, true, false, true, true);
}

int hb_pack_counts_dev(const int32_t *x, const int32_t *size, int64_t count, uint32_t *out, void *stream)
{
    std::lock_guard<std::mutex> lk(g_mu);
    Workspace *ws;
    HBTRY(cur_ws(&ws));
    WS_USE(stream);
    if (!x || !size || !out || count < 1) return fail(HB_ERR_ARG, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    HBTRY(ws->status.ensure(4 * sizeof(int32_t)));
    int32_t *flag = ws->status.as<int32_t>() + 3;
    CU(cudaMemsetAsync(flag, 0, sizeof(int32_t), st));
    CU(hb::launch_pack_counts(x, size, count, out, flag, st));
    int32_t h = 0;
    CU(cudaMemcpyAsync(&h, flag, sizeof h, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (h) return fail(HB_ERR_ARG, "a count does not fit 16 bits: use the two-plane entry points");
    return HB_OK;
}

int hb_hmm_binom_host(const int32_t *x, const int32_t *size, int64_t T, int64_t B,
                      const int64_t *seg_off, int32_t nseg, const double *prob, double rho,
                      const double *Pi, const double *delta, int32_t flags, int32_t *y,
                      double *gamma, double *ll)
{
    std::lock_guard<std::mutex> lk(g_mu);
    Workspace *ws;
    HBTRY(cur_ws(&ws));
    if (T < 1 || B < 1 || !x || !size) return fail(HB_ERR_ARG, "bad arguments");
    const int64_t one_seg[2] = {0, T};
    if (!seg_off) {
        seg_off = one_seg;
        nseg = 1;
    }
    bool vit, fb;
    HBTRY(check_flags(flags, y, gamma, ll, &vit, &fb));
    const bool want_g = (flags & HB_WANT_GAMMA) != 0, want_ll = (flags & HB_WANT_LL) != 0;
    WsUse ws_use_;
    if (rho == 0.0 && long_chain_shape(T, B, nseg, flags) && seg_off[0] == 0 && seg_off[1] == T) {
        HBTRY(ws_use_.enter(ws, 0));
        // Viterbi(dthmm(mafl, ..., "binom", ...)) on one (or a few) pooled count vectors: see hb_hmm_norm_host
        HBTRY(ws->st_in.ensure((size_t)T * B * sizeof(int32_t)));
        HBTRY(ws->st_in2.ensure((size_t)T * B * sizeof(int32_t)));
        CU(cudaMemcpyAsync(ws->st_in.p, x, (size_t)T * B * sizeof(int32_t), cudaMemcpyHostToDevice, 0));
        CU(cudaMemcpyAsync(ws->st_in2.p, size, (size_t)T * B * sizeof(int32_t), cudaMemcpyHostToDevice, 0));
        HBTRY(binom2_emissions_cm(ws, ws->st_in.as<int32_t>(), ws->st_in2.as<int32_t>(), T, (int32_t)B, prob, 0));
        return longchain_host_outputs(ws, 2, T, B, Pi, delta, flags, y, gamma, ll, false);
    }
    HBTRY(pipe_init(ws));
    HBTRY(ws_use_.enter(ws, ws->s_cmp));
    const bool recon_all = want_g && recon_enabled();
    const int64_t Bc = pipe_batch(B, 160.0 * (double)T);
    const int64_t nb = (B + Bc - 1) / Bc;
    const int64_t ldc = (Bc + 127) / 128 * 128;
    for (int k = 0; k < HB_PIPE_SLOTS; k++) {
        HBTRY(ws->pin[k].ensure((size_t)T * Bc * sizeof(int32_t)));
        HBTRY(ws->pin2[k].ensure((size_t)T * Bc * sizeof(int32_t)));
        if (vit) HBTRY(ws->pout_y[k].ensure((size_t)T * Bc));
        if (want_g) HBTRY(ws->pout_g[k].ensure((size_t)2 * T * Bc * sizeof(double)));
        if (want_ll) HBTRY(ws->pout_ll[k].ensure((size_t)nseg * Bc * sizeof(double)));
    }
    HBTRY(ws->st_x.ensure((size_t)T * ldc * sizeof(int32_t)));
    HBTRY(ws->st_x2.ensure((size_t)T * ldc * sizeof(int32_t)));
    if (vit) HBTRY(ws->st_y.ensure((size_t)T * ldc));
    if (want_g) HBTRY(ws->st_g.ensure((size_t)2 * T * ldc * sizeof(double)));
    HostStager hs;
    hs.balance = recon_adaptive();
    const auto wall0 = std::chrono::steady_clock::now();
    HBTRY(hs.init(ws, !(host_is_pinned(x) && host_is_pinned(size)),
                  !(host_is_pinned(want_g ? gamma : nullptr) && host_is_pinned(want_ll ? ll : nullptr)),
                  2 * ((size_t)T * Bc * sizeof(int32_t) + 256),
                  (want_g ? 2 * ((size_t)T * Bc * sizeof(double) + 256) : 0) +
                      (want_ll ? (size_t)nseg * (Bc * sizeof(double) + 256) : 0) + 256,
                  vit ? (size_t)T * Bc : 0));
    int rc = HB_OK;
    for (int64_t i = 0; i < nb && rc == HB_OK; i++) {
        const int k = (int)(i % HB_PIPE_SLOTS);
        const int64_t b0 = i * Bc, bn = std::min(Bc, B - b0);
        const int64_t ld = (bn + 127) / 128 * 128;
        const bool recon = want_g ? hs.plan(k, (size_t)T * bn, 2, recon_all) : false; // see hb_hmm_norm_host
        if (i >= HB_PIPE_SLOTS) CU(cudaStreamWaitEvent(ws->s_in, ws->ev_free_in[k], 0));
        HBTRY(hs.begin_in(k, i >= HB_PIPE_SLOTS));
        HBTRY(hs.h2d(k, ws->pin[k].p, x + b0 * T, (size_t)T * bn * sizeof(int32_t), ws->s_in));
        HBTRY(hs.h2d(k, ws->pin2[k].p, size + b0 * T, (size_t)T * bn * sizeof(int32_t), ws->s_in));
        HBTRY(hs.end_in(k, ws->s_in));
        CU(cudaEventRecord(ws->ev_in[k], ws->s_in));
        CU(cudaStreamWaitEvent(ws->s_cmp, ws->ev_in[k], 0));
        CU(hb::launch_transpose_i32(ws->pin[k].as<int32_t>(), T, bn, ws->st_x.as<int32_t>(), ld, ws->s_cmp));
        CU(hb::launch_transpose_i32(ws->pin2[k].as<int32_t>(), T, bn, ws->st_x2.as<int32_t>(), ld, ws->s_cmp));
        CU(cudaEventRecord(ws->ev_free_in[k], ws->s_cmp));
        uint8_t *d_y = vit ? ws->st_y.as<uint8_t>() : nullptr;
        double *d_g = want_g ? ws->st_g.as<double>() : nullptr;
        double *d_ll = want_ll ? ws->pout_ll[k].as<double>() : nullptr;
        if (i >= HB_PIPE_SLOTS) CU(cudaStreamWaitEvent(ws->s_cmp, ws->ev_free_out[k], 0));
        rc = binom2_run(ws, ws->st_x.as<int32_t>(), ws->st_x2.as<int32_t>(), ld, T, bn, seg_off, nseg,
                        prob, rho, Pi, delta, flags, d_y, ld, d_g, ld, d_ll, ws->s_cmp, i == 0);
        if (rc != HB_OK) break;
        if (vit)
            CU(hb::launch_transpose_u8_back_u8(d_y, ld, T, bn, ws->pout_y[k].as<uint8_t>(), ws->s_cmp));
        if (want_g)
            for (int pl = 0; pl < (recon ? 1 : 2); pl++)
                CU(hb::launch_transpose_f64_back(d_g + (int64_t)pl * T * ld, ld, T, bn,
                                                 ws->pout_g[k].as<double>() + (int64_t)pl * T * bn,
                                                 ws->s_cmp));
        CU(cudaEventRecord(ws->ev_out[k], ws->s_cmp));
        CU(cudaStreamWaitEvent(ws->s_out, ws->ev_out[k], 0));
        hs.begin_out(k);
        HBTRY(hs.link_begin(k, ws->s_out));
        if (vit)
            HBTRY(hs.d2h_widen(k, y + b0 * T, ws->pout_y[k].as<uint8_t>(), (size_t)T * bn, ws->s_out));
        if (want_g) {
            for (int pl = 0; pl < (recon ? 1 : 2); pl++)
                HBTRY(hs.d2h(k, gamma + ((int64_t)pl * B + b0) * T,
                             ws->pout_g[k].as<double>() + (int64_t)pl * T * bn,
                             (size_t)T * bn * sizeof(double), ws->s_out));
            if (recon) {
                hs.recon(k, gamma + b0 * T, nullptr, gamma + (B + b0) * T, (size_t)T * bn);
                ws->io_rebuilt += (int64_t)T * bn * (int64_t)sizeof(double);
            }
        }
        if (want_ll)
            for (int32_t sg = 0; sg < nseg; sg++)
                HBTRY(hs.d2h(k, ll + (int64_t)sg * B + b0, d_ll + (int64_t)sg * bn,
                             (size_t)bn * sizeof(double), ws->s_out));
        HBTRY(hs.link_end(k, ws->s_out));
        CU(cudaEventRecord(ws->ev_free_out[k], ws->s_out));
        HBTRY(hs.end_out(k, ws->s_out));
    }
    cudaStreamSynchronize(ws->s_in);
    cudaStreamSynchronize(ws->s_cmp);
    cudaError_t e = cudaStreamSynchronize(ws->s_out);
    hs.finish();
    hs.trace("host pipeline", T, B, std::chrono::duration<double>(std::chrono::steady_clock::now() - wall0).count());
    if (rc != HB_OK) return rc;
    if (e != cudaSuccess) return fail(HB_ERR_CUDA, "pipeline failed: %s", cudaGetErrorString(e));
    return read_status(ws, ws->s_cmp);
}

// ------------------------------------------------------------------ layout
int hb_transpose_f64_dev(const double *src, int64_t G, int64_t N, double *dst, int64_t ld,
                         void *stream)
{
    if (!src || !dst || G < 1 || N < 1 || ld < N) return fail(HB_ERR_ARG, "bad arguments");
    CU(hb::launch_transpose_f64(src, G, N, dst, ld, (cudaStream_t)stream));
    return HB_OK;
}

int hb_transpose_f64_to_i32_dev(const double *src, int64_t G, int64_t N, int32_t *dst, int64_t ld,
                                void *stream)
{
    if (!src || !dst || G < 1 || N < 1 || ld < N) return fail(HB_ERR_ARG, "bad arguments");
    CU(hb::launch_transpose_f64_i32(src, G, N, dst, ld, (cudaStream_t)stream));
    return HB_OK;
}

int hb_transpose_u8_back_dev(const uint8_t *src, int64_t ld, int64_t G, int64_t N, int32_t *dst,
                             void *stream)
{
    if (!src || !dst || G < 1 || N < 1 || ld < N) return fail(HB_ERR_ARG, "bad arguments");
    CU(hb::launch_transpose_u8_back(src, ld, G, N, dst, (cudaStream_t)stream));
    return HB_OK;
}

int hb_transpose_f64_back_dev(const double *src, int64_t ld, int64_t G, int64_t N, double *dst,
                              void *stream)
{
    if (!src || !dst || G < 1 || N < 1 || ld < N) return fail(HB_ERR_ARG, "bad arguments");
    CU(hb::launch_transpose_f64_back(src, ld, G, N, dst, (cudaStream_t)stream));
    return HB_OK;
}

int hb_tile_dev(const void *src, int64_t ld, int64_t T, int64_t B, int32_t elem, void *dst, void *stream)
{
    if (!src || !dst || T < 1 || B < 1 || ld < B || (elem != 1 && elem != 4 && elem != 8))
        return fail(HB_ERR_ARG, "bad arguments");
    CU(hb::launch_tile(src, ld, T, B, elem, dst, true, (cudaStream_t)stream));
    return HB_OK;
}

int hb_untile_dev(const void *src, int64_t T, int64_t B, int32_t elem, void *dst, int64_t ld, void *stream)
{
    if (!src || !dst || T < 1 || B < 1 || ld < B || (elem != 1 && elem != 4 && elem != 8))
        return fail(HB_ERR_ARG, "bad arguments");
    CU(hb::launch_tile(src, ld, T, B, elem, dst, false, (cudaStream_t)stream));
    return HB_OK;
}

// ------------------------------------------------------------------ statistics
int hb_chain_sd_dev(const double *x, int64_t ld, int64_t T, int64_t B, const int64_t *seg_off,
                    int32_t nseg, int32_t per_seg, double *sd, void *stream)
{
    std::lock_guard<std::mutex> lk(g_mu);
    Workspace *ws;
    HBTRY(cur_ws(&ws));
    WS_USE(stream);
    if (!x || !sd || T < 2 || B < 1 || ld < B) return fail(HB_ERR_ARG, "bad arguments");
    const int64_t one_seg[2] = {0, T};
    if (!per_seg || !seg_off) {
        seg_off = one_seg;
        nseg = 1;
    }
    const int64_t *d_seg, *d_chunk;
    const int32_t *d_order;
    HBTRY(prepare_segments(ws, seg_off, nseg, T, (cudaStream_t)stream, &d_seg, &d_chunk, &d_order));
    CU(hb::launch_chain_sd(x, ld, B, d_seg, nseg, sd, (cudaStream_t)stream));
    return HB_OK;
}

// group sizes + launch order (largest group first) -> device, scratch of the throughput form (falls back to the
// block-per-gene-tile kernel when the cell-major copy does not fit), then the two passes.  Caller holds g_mu.
static int pool_mean_launch(Workspace *ws, const double *x, int64_t ld, int64_t G, int64_t N, const uint8_t *member_dev,
                            const std::vector<int32_t> &sz, int32_t K, double *out, cudaStream_t st)
{
    std::vector<int32_t> tab(2 * (size_t)K);
    std::vector<int32_t> order(K);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int32_t p, int32_t q) { return sz[p] > sz[q]; });
    for (int32_t k = 0; k < K; k++) {
        tab[k] = sz[k];
        tab[K + k] = order[k];
    }
    HBTRY(ws->gsize.ensure(tab.size() * sizeof(int32_t)));
    HBTRY(upload_async(ws, ws->gsize.p, tab.data(), tab.size() * sizeof(int32_t), st));
    HBTRY(ws->st_x80.ensure((size_t)2 * K * G * 16));
    void *fast = nullptr;
    if (ws->pool_fast.ensure(hb::pool_accum_scratch_bytes(G, N, K)) == HB_OK) fast = ws->pool_fast.p;
    CU(hb::launch_pool_mean(x, ld, G, N, member_dev, ws->gsize.as<int32_t>(), K, out, ws->st_x80.p, fast,
                            ws->gsize.as<int32_t>() + K, st));
    return HB_OK;
}

int hb_pool_mean_dev(const double *x, int64_t ld, int64_t G, int64_t N, const uint8_t *member,
                     int32_t K, double *out, void *stream)
{
    std::lock_guard<std::mutex> lk(g_mu);
    Workspace *ws;
    HBTRY(cur_ws(&ws));
    WS_USE(stream);
    if (!x || !member || !out || G < 1 || N < 1 || K < 1 || ld < N)
        return fail(HB_ERR_ARG, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    // group sizes: tiny D2H of the membership (K*N bytes)
    std::vector<uint8_t> h_mem((size_t)K * N);
    CU(cudaMemcpyAsync(h_mem.data(), member, h_mem.size(), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    std::vector<int32_t> sz;
    HBTRY(group_sizes(h_mem.data(), K, N, sz));
    HBTRY(pool_mean_launch(ws, x, ld, G, N, member, sz, K, out, st));
    return HB_OK;
}

int hb_pool_mean_relay_dev(const double *x, int64_t ld, int64_t G, int64_t N, const uint8_t *member,
                           int32_t K, const int32_t *n_total, int32_t phase, void *state,
                           void *mean80, double *out, void *stream)
{
    if (G < 1 || K < 1 || !state || phase < 0 || phase > 3) return fail(HB_ERR_ARG, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (phase == 0 || phase == 2) {
        if (N < 0 || (N > 0 && (!x || !member || ld < N))) return fail(HB_ERR_ARG, "bad arguments");
        if (phase == 2 && !mean80) return fail(HB_ERR_ARG, "phase 2 needs mean80");
        if (N > 0) {
            std::lock_guard<std::mutex> lk(g_mu);
            Workspace *ws;
            HBTRY(cur_ws(&ws));
            WS_USE(stream);
            if (ws->pool_fast.ensure(hb::pool_accum_scratch_bytes(G, N, K)) == HB_OK) {
                CU(hb::launch_pool_accum_prepare(x, ld, G, N, member, K, ws->pool_fast.p, st));
                CU(hb::launch_pool_accum_fast(G, N, K, nullptr, phase == 2 ? mean80 : nullptr, state,
                                              ws->pool_fast.p, st));
            } else {
                CU(hb::launch_pool_accum(x, ld, G, N, member, K, phase == 2 ? mean80 : nullptr, state, st));
            }
        }
        return HB_OK;
    }
    if (!n_total || !mean80) return fail(HB_ERR_ARG, "phases 1 and 3 need n_total and mean80");
    if (phase == 1) {
        CU(hb::launch_pool_div(state, n_total, G, K, mean80, st));
    } else {
        if (!out) return fail(HB_ERR_ARG, "phase 3 needs out");
        CU(hb::launch_pool_finish(mean80, state, n_total, G, K, out, st));
    }
    return HB_OK;
}

int hb_pooled_sd_dev(const double *pooled, int64_t G, int32_t K, double *out_sd, void *stream)
{
    if (!pooled || !out_sd || G < 2 || K < 1) return fail(HB_ERR_ARG, "bad arguments");
    CU(hb::launch_pooled_sd(pooled, G, K, out_sd, (cudaStream_t)stream));
    return HB_OK;
}

int hb_pool_count_pos_dev(const int32_t *v, int64_t ld, int64_t G, int64_t N,
                          const uint8_t *member, int32_t K, int32_t *out, void *stream)
{
    std::lock_guard<std::mutex> lk(g_mu);
    Workspace *ws;
    HBTRY(cur_ws(&ws));
    WS_USE(stream);
    if (!v || !member || !out || G < 1 || N < 1 || K < 1 || ld < N)
        return fail(HB_ERR_ARG, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    HBTRY(ws->cellmask.ensure((size_t)N * sizeof(uint32_t)));
    for (int32_t k0 = 0; k0 < K; k0 += 32) {
        const int32_t kn = std::min(32, K - k0);
        CU(hb::launch_build_cellmask(member + (int64_t)k0 * N, N, kn, ws->cellmask.as<uint32_t>(), st));
        CU(hb::launch_pool_count(v, ld, G, N, ws->cellmask.as<uint32_t>(), kn, out + (int64_t)k0 * G,
                                 st));
    }
    return HB_OK;
}

// ------------------------------------------------------------------ grouping of cells (8f-2)
int hb_runmean_dev(const double *x, int64_t ld, int64_t G, int64_t N, int32_t k, double *out, int64_t ldo,
                   void *stream)
{
    if (!x || !out || G < 1 || N < 1 || ld < N || ldo < N || k < 1 || (k & 1) == 0)
        return fail(HB_ERR_ARG, "bad arguments (k must be odd)");
    CU(hb::launch_runmean(x, nullptr, nullptr, ld, G, N, k, out, ldo, nullptr, (cudaStream_t)stream));
    return HB_OK;
}

int hb_runmean_ratio_dev(const int32_t *r, const int32_t *n, int64_t ld, int64_t G, int64_t N, int32_t k,
                         double *out, int64_t ldo, void *stream)
{
    std::lock_guard<std::mutex> lk(g_mu);
    Workspace *ws;
    HBTRY(cur_ws(&ws));
    WS_USE(stream);
    if (!r || !n || !out || G < 1 || N < 1 || ld < N || ldo < N || k < 1 || (k & 1) == 0)
        return fail(HB_ERR_ARG, "bad arguments (k must be odd)");
    HBTRY(ws->st_out.ensure((size_t)G * ldo * sizeof(double))); // the ratios, formed once
    CU(hb::launch_runmean(nullptr, r, n, ld, G, N, k, out, ldo, ws->st_out.as<double>(), (cudaStream_t)stream));
    return HB_OK;
}

int hb_dist_sq_dev(const double *s, int64_t ld, int64_t G, int64_t N, double *D, void *stream)
{
    std::lock_guard<std::mutex> lk(g_mu);
    Workspace *ws;
    HBTRY(cur_ws(&ws));
    WS_USE(stream);
    if (!s || !D || G < 1 || N < 1 || ld < N) return fail(HB_ERR_ARG, "bad arguments");
    void *flags = nullptr;
    if (g_dist_mode != 1) {
        HBTRY(ws->st_part.ensure(hb::dist_scratch_bytes(G, N)));
        flags = ws->st_part.p;
    }
    CU(hb::launch_dist_sq(s, ld, G, N, D, flags, (cudaStream_t)stream));
    return HB_OK;
}

int hb_dist_sq_rect_dev(const double *sa, int64_t lda, int64_t na, const double *sb, int64_t ldb, int64_t nb,
                        int64_t G, double *D, int64_t ldd, void *stream)
{
    std::lock_guard<std::mutex> lk(g_mu);
    Workspace *ws;
    HBTRY(cur_ws(&ws));
    WS_USE(stream);
    if (!sa || !sb || !D || G < 1 || na < 2 || nb < 2 || lda < na || ldb < nb || ldd < nb)
        return fail(HB_ERR_ARG, "bad arguments");
    if ((lda & 1) || (ldb & 1) || ((uintptr_t)sa & 15) || ((uintptr_t)sb & 15))
        return fail(HB_ERR_ARG, "hb_dist_sq_rect_dev needs even row strides and 16-byte aligned matrices");
    HBTRY(ws->st_part.ensure(hb::dist_scratch_bytes(G, na) + hb::dist_scratch_bytes(G, nb) + 512));
    CU(hb::launch_dist_sq_rect(sa, lda, na, sb, ldb, nb, G, D, ldd, ws->st_part.p, (cudaStream_t)stream));
    return HB_OK;
}

int hb_ward_linkage_dev(double *D, int64_t N, int32_t *merge_a, int32_t *merge_b, double *height,
                        void *stream)
{
    std::lock_guard<std::mutex> lk(g_mu);
    Workspace *ws;
    HBTRY(cur_ws(&ws));
    WS_USE(stream);
    if (!D || !merge_a || !merge_b || !height || N < 2) return fail(HB_ERR_ARG, "bad arguments");
    HBTRY(ws->st_part.ensure(hb::ward_scratch_bytes(N)));
    CU(hb::launch_ward(D, N, ws->st_part.p, merge_a, merge_b, height, g_ward_mode, (cudaStream_t)stream));
    return HB_OK;
}

// cutree(hc, k) for each k from the chain-ordered merge list: the N - k lowest merges (stable sort by
// height keeps children before parents), union-find over leaves, clusters numbered by first appearance.
static void cut_merges_host(const int32_t *ma, const int32_t *mb, const double *h, int64_t N,
                            const int32_t *ks, int32_t nk, int32_t *labels)
{
    std::vector<int32_t> ord((size_t)(N - 1));
    std::iota(ord.begin(), ord.end(), 0);
    std::stable_sort(ord.begin(), ord.end(), [&](int32_t a, int32_t b) { return h[a] < h[b]; });
    std::vector<int32_t> parent((size_t)N), remap((size_t)N);
    for (int32_t q = 0; q < nk; q++) {
        std::iota(parent.begin(), parent.end(), 0);
        auto find = [&](int32_t i) {
            while (parent[i] != i) {
                parent[i] = parent[parent[i]];
                i = parent[i];
            }
            return i;
        };
        const int64_t nm = std::max<int64_t>(0, N - std::max(1, ks[q]));
        for (int64_t m = 0; m < nm; m++) {
            const int32_t ra = find(ma[ord[m]]), rb = find(mb[ord[m]]);
            if (ra != rb) parent[ra] = rb;
        }
        std::fill(remap.begin(), remap.end(), 0);
        int32_t next = 0;
        for (int64_t i = 0; i < N; i++) {
            const int32_t r = find((int32_t)i);
            if (!remap[r]) remap[r] = ++next;
            labels[(int64_t)q * N + i] = remap[r];
        }
    }
}

int hb_cutree_host(const int32_t *merge_a, const int32_t *merge_b, const double *height, int64_t N,
                   const int32_t *ks, int32_t nk, int32_t *labels)
{
    if (!merge_a || !merge_b || !height || !ks || !labels || N < 1 || nk < 1) return fail(HB_ERR_ARG, "bad arguments");
    for (int64_t m = 0; m + 1 < N; m++)
        if (merge_a[m] < 0 || merge_a[m] >= N || merge_b[m] < 0 || merge_b[m] >= N)
            return fail(HB_ERR_ARG, "merge list names a leaf outside [0, N)");
    if (N == 1) {
        for (int32_t q = 0; q < nk; q++) labels[q] = 1;
        return HB_OK;
    }
    cut_merges_host(merge_a, merge_b, height, N, ks, nk, labels);
    return HB_OK;
}

// Grouping core on device-resident gene-major data (caller holds g_mu and a WsUse): runmean -> dist ->
// Ward.D2 -> cutree for each k.  Either xg (fp64) or rg / ng (int32 counts) is given.
static int group_core(Workspace *ws, const double *xg, const int32_t *rg, const int32_t *ng, int64_t ld, int64_t G,
                      int64_t N, int32_t window, const int32_t *ks, int32_t nk, int32_t *labels, cudaStream_t st)
{
    if (N == 1) {
        for (int32_t q = 0; q < nk; q++) labels[q] = 1;
        return HB_OK;
    }
    const int64_t lds = (N + 127) / 128 * 128;
    HBTRY(ws->st_g.ensure((size_t)G * lds * sizeof(double))); // smoothed matrix
    double *S = ws->st_g.as<double>();
    double *ratio = nullptr;
    if (!xg) {
        HBTRY(ws->st_out.ensure((size_t)G * lds * sizeof(double))); // r / n formed once, not once per window position
        ratio = ws->st_out.as<double>();
    }
    CU(hb::launch_runmean(xg, rg, ng, ld, G, N, window, S, lds, ratio, st));
    HBTRY(ws->st_pool.ensure((size_t)N * N * sizeof(double)));
    HBTRY(ws->st_pool2.ensure((size_t)N * 16 + 64));
    HBTRY(ws->st_part.ensure(std::max(hb::ward_scratch_bytes(N), hb::dist_scratch_bytes(G, N))));
    double *D = ws->st_pool.as<double>();
    CU(hb::launch_dist_sq(S, lds, G, N, D, g_dist_mode != 1 ? ws->st_part.p : nullptr, st));
    int32_t *d_ma = ws->st_pool2.as<int32_t>(), *d_mb = d_ma + N;
    double *d_h = reinterpret_cast<double *>(d_mb + N);
    CU(hb::launch_ward(D, N, ws->st_part.p, d_ma, d_mb, d_h, g_ward_mode, st));
    std::vector<int32_t> ma((size_t)N), mb((size_t)N);
    std::vector<double> hh((size_t)N);
    CU(cudaMemcpyAsync(ma.data(), d_ma, (size_t)(N - 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(mb.data(), d_mb, (size_t)(N - 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(hh.data(), d_h, (size_t)(N - 1) * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    cut_merges_host(ma.data(), mb.data(), hh.data(), N, ks, nk, labels);
    return HB_OK;
}

// Host-buffer form of the grouping step: x (or r, n) column-major [G x N] as R holds them.
static int group_host(const double *x, const double *r, const double *n, int64_t G, int64_t N,
                      int32_t window, const int32_t *ks, int32_t nk, int32_t *labels)
{
    std::lock_guard<std::mutex> lk(g_mu);
    Workspace *ws;
    HBTRY(cur_ws(&ws));
    WS_USE(0);
    if (G < 1 || N < 1 || !ks || nk < 1 || !labels || window < 1 || (window & 1) == 0)
        return fail(HB_ERR_ARG, "bad arguments");
    cudaStream_t st = 0;
    const int64_t ld = (N + 127) / 128 * 128;
    HBTRY(ws->st_in.ensure((size_t)G * N * sizeof(double)));
    if (x) {
        HBTRY(ws->st_x.ensure((size_t)G * ld * sizeof(double)));
        CU(cudaMemcpyAsync(ws->st_in.p, x, (size_t)G * N * sizeof(double), cudaMemcpyHostToDevice, st));
        CU(hb::launch_transpose_f64(ws->st_in.as<double>(), G, N, ws->st_x.as<double>(), ld, st));
        return group_core(ws, ws->st_x.as<double>(), nullptr, nullptr, ld, G, N, window, ks, nk, labels, st);
    }
    HBTRY(ws->st_x.ensure((size_t)G * ld * sizeof(int32_t)));
    HBTRY(ws->st_x2.ensure((size_t)G * ld * sizeof(int32_t)));
    CU(cudaMemcpyAsync(ws->st_in.p, r, (size_t)G * N * sizeof(double), cudaMemcpyHostToDevice, st));
    CU(hb::launch_transpose_f64_i32(ws->st_in.as<double>(), G, N, ws->st_x.as<int32_t>(), ld, st));
    CU(cudaMemcpyAsync(ws->st_in.p, n, (size_t)G * N * sizeof(double), cudaMemcpyHostToDevice, st));
    CU(hb::launch_transpose_f64_i32(ws->st_in.as<double>(), G, N, ws->st_x2.as<int32_t>(), ld, st));
    return group_core(ws, nullptr, ws->st_x.as<int32_t>(), ws->st_x2.as<int32_t>(), ld, G, N, window, ks, nk, labels,
                      st);
}

int hb_group_gexp_host(const double *gexp_norm, int64_t G, int64_t N, int32_t window, const int32_t *ks,
                       int32_t nk, int32_t *labels)
{
    if (!gexp_norm) return fail(HB_ERR_ARG, "bad arguments");
    return group_host(gexp_norm, nullptr, nullptr, G, N, window, ks, nk, labels);
}

int hb_group_allele_host(const double *r_maf, const double *n_sc, int64_t G, int64_t N, int32_t window,
                         const int32_t *ks, int32_t nk, int32_t *labels)
{
    if (!r_maf || !n_sc) return fail(HB_ERR_ARG, "bad arguments");
    return group_host(nullptr, r_maf, n_sc, G, N, window, ks, nk, labels);
}

// ------------------------------------------------------------------ input construction (8f-3)
int hb_row_means_dev(const double *x, int64_t ld, int64_t G, int64_t N, double *out, void *stream)
{
    if (!x || !out || G < 1 || N < 1 || ld < N || N >= ((int64_t)1 << 32)) return fail(HB_ERR_ARG, "bad arguments");
    CU(hb::launch_row_means(x, ld, G, N, out, (cudaStream_t)stream));
    return HB_OK;
}

int hb_col_scale_stats_dev(const double *x, int64_t ld, int64_t G, int64_t N, const uint8_t *keep,
                           double *center, double *scl, void *stream)
{
    if (!x || !center || G < 1 || N < 1 || ld < N || G >= ((int64_t)1 << 32)) return fail(HB_ERR_ARG, "bad arguments");
    CU(hb::launch_col_stats(x, ld, G, N, keep, center, scl, (cudaStream_t)stream));
    return HB_OK;
}

int hb_gexp_apply_dev(const double *x, int64_t ld, int64_t N, const int32_t *rows, int64_t Gk,
                      const double *center, const double *scl, const double *refmean, double *out,
                      int64_t ldo, void *stream)
{
    if (!x || !out || N < 1 || Gk < 0 || ld < N || ldo < N) return fail(HB_ERR_ARG, "bad arguments");
    if (Gk == 0) return HB_OK;
    CU(hb::launch_gexp_apply(x, ld, N, rows, Gk, center, scl, refmean, out, ldo, (cudaStream_t)stream));
    return HB_OK;
}

// mean(x[x != 0]).  exact = 0: parallel double-double sum; *margin bounds |result - R's value| (R's
// serial evaluation can differ from the true mean by n * 2^-62 * max|x|).  exact = 1: R's loop replayed.
static int mean_nonzero(Workspace *ws, const double *x, int64_t ld, int64_t G, int64_t N, int32_t exact,
                        double *mean, double *margin, cudaStream_t st)
{
    const int nblk = 148 * 4;
    HBTRY(ws->st_part.ensure((size_t)nblk * 4 * sizeof(double) + 64));
    double *ph = ws->st_part.as<double>(), *pl = ph + nblk, *pm = pl + nblk;
    unsigned long long *pn = (unsigned long long *)(pm + nblk);
    CU(hb::launch_nz_sum(x, ld, G, N, nblk, ph, pl, pn, pm, st));
    std::vector<double> h((size_t)nblk * 4);
    CU(cudaMemcpyAsync(h.data(), ph, h.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    // combine the block partials in double-double (error-free additions, libm-free)
    double H = 0.0, L = 0.0, mx = 0.0;
    unsigned long long n = 0;
    auto dd = [&](double v) {
        volatile double s = H + v;
        volatile double bb = s - H;
        volatile double e1 = H - (s - bb), e2 = v - bb;
        H = s;
        L += e1 + e2;
    };
    for (int i = 0; i < nblk; i++) {
        dd(h[i]);
        dd(h[nblk + i]);
        mx = std::max(mx, h[2 * nblk + i]);
        unsigned long long c;
        memcpy(&c, &h[3 * nblk + i], sizeof c);
        n += c;
    }
    if (n == 0) { // mean(numeric(0)) is NaN in R
        *mean = NAN;
        if (margin) *margin = 0.0;
        return HB_OK;
    }
    if (n >= (1ull << 32)) return fail(HB_ERR_ARG, "more than 2^32 non-zero entries");
    if (!exact) {
        const long double m = ((long double)H + (long double)L) / (long double)n;
        *mean = (double)m;
        if (margin) *margin = (double)n * ldexp(mx, -62) + ldexp(fabs(*mean), -50);
        return HB_OK;
    }
    double *d_out = ws->st_part.as<double>();
    CU(hb::launch_nz_mean_seq(x, ld, G, N, n, d_out, st));
    CU(cudaMemcpyAsync(mean, d_out, sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (margin) *margin = 0.0;
    return HB_OK;
}

int hb_mean_nonzero_dev(const double *x, int64_t ld, int64_t G, int64_t N, int32_t exact, double *mean,
                        double *margin, void *stream)
{
    std::lock_guard<std::mutex> lk(g_mu);
    Workspace *ws;
    HBTRY(cur_ws(&ws));
    WS_USE(stream);
    if (!x || !mean || G < 1 || N < 1 || ld < N) return fail(HB_ERR_ARG, "bad arguments");
    return mean_nonzero(ws, x, ld, G, N, exact, mean, margin, (cudaStream_t)stream);
}

// setGexpMats (R/HoneyBADGER.R:124-188) on device-resident gene-major matrices; caller holds g_mu.
static int set_gexp_mats(Workspace *ws, const double *sc, int64_t ld, int64_t G, int64_t N, const double *ref,
                         int64_t ldr, int64_t R, int32_t filter, double min_both, const double *min_test,
                         const double *min_ref, int32_t scale, double *norm, int64_t ldo, int32_t *keep_rows,
                         int64_t *Gk_out, int32_t *info, cudaStream_t st)
{
    if (info) *info = 0;
    // scratch: rm_s[G] rm_r[G] refmean[G] | center_s[N] scl_s[N] | center_r[R] scl_r[R] | rows[G] | keep[G]
    const size_t nd = (size_t)3 * G + 2 * N + 2 * R;
    HBTRY(ws->st_inp.ensure(nd * sizeof(double) + (size_t)G * sizeof(int32_t) + (size_t)G + 64));
    double *rm_s = ws->st_inp.as<double>(), *rm_r = rm_s + G, *refmean = rm_r + G;
    double *cen_s = refmean + G, *scl_s = cen_s + N, *cen_r = scl_s + N, *scl_r = cen_r + R;
    int32_t *d_rows = (int32_t *)(scl_r + R);
    uint8_t *d_keep = (uint8_t *)(d_rows + G);
    CU(hb::launch_row_means(ref, ldr, G, R, rm_r, st));
    std::vector<int32_t> rows;
    rows.reserve((size_t)G);
    if (filter) {
        CU(hb::launch_row_means(sc, ld, G, N, rm_s, st));
        std::vector<double> h((size_t)2 * G);
        CU(cudaMemcpyAsync(h.data(), rm_s, h.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
        double mt = 0, mr = 0, mgt = 0, mgr = 0;
        if (min_test) mt = *min_test;
        else HBTRY(mean_nonzero(ws, sc, ld, G, N, 0, &mt, &mgt, st));
        if (min_ref) mr = *min_ref;
        else HBTRY(mean_nonzero(ws, ref, ldr, G, R, 0, &mr, &mgr, st));
        CU(cudaStreamSynchronize(st));
        // a row mean inside the safety margin of a default threshold: only R's serial evaluation decides
        bool near_t = false, near_r = false;
        for (int64_t g = 0; g < G; g++) {
            if (!min_test && fabs(h[g] - mt) <= mgt) near_t = true;
            if (!min_ref && fabs(h[G + g] - mr) <= mgr) near_r = true;
        }
        if (near_t) HBTRY(mean_nonzero(ws, sc, ld, G, N, 1, &mt, nullptr, st));
        if (near_r) HBTRY(mean_nonzero(ws, ref, ldr, G, R, 1, &mr, nullptr, st));
        if (info) *info = (near_t ? 1 : 0) | (near_r ? 2 : 0);
        for (int64_t g = 0; g < G; g++) {
            const double a = h[g], b = h[G + g];
            if ((a > min_both && b > min_both) || a > mt || b > mr) rows.push_back((int32_t)g);
        }
    } else {
        rows.resize((size_t)G);
        std::iota(rows.begin(), rows.end(), 0);
    }
    const int64_t Gk = (int64_t)rows.size();
    *Gk_out = Gk;
    if (keep_rows) memcpy(keep_rows, rows.data(), (size_t)Gk * sizeof(int32_t));
    if (Gk == 0) return HB_OK;
    std::vector<uint8_t> keep((size_t)G, 0);
    for (int32_t g : rows) keep[(size_t)g] = 1;
    CU(cudaMemcpyAsync(d_rows, rows.data(), (size_t)Gk * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_keep, keep.data(), (size_t)G, cudaMemcpyHostToDevice, st));
    if (scale) {
        // scale(gexp.ref): centre / scale the kept rows of every reference column, then rowMeans
        const int64_t ldr2 = (R + 31) / 32 * 32;
        HBTRY(ws->st_inp2.ensure((size_t)Gk * ldr2 * sizeof(double)));
        double *ref2 = ws->st_inp2.as<double>();
        CU(hb::launch_col_stats(ref, ldr, G, R, d_keep, cen_r, scl_r, st));
        CU(hb::launch_gexp_apply(ref, ldr, R, d_rows, Gk, cen_r, scl_r, nullptr, ref2, ldr2, st));
        CU(hb::launch_row_means(ref2, ldr2, Gk, R, refmean, st));
        CU(hb::launch_col_stats(sc, ld, G, N, d_keep, cen_s, scl_s, st));
        CU(hb::launch_gexp_apply(sc, ld, N, d_rows, Gk, cen_s, scl_s, refmean, norm, ldo, st));
    } else {
        // rowMeans of a row subset is the subset of the rowMeans: gather rm_r (N = 1 "matrix")
        CU(hb::launch_gexp_apply(rm_r, 1, 1, d_rows, Gk, nullptr, nullptr, nullptr, refmean, 1, st));
        CU(hb::launch_gexp_apply(sc, ld, N, d_rows, Gk, nullptr, nullptr, refmean, norm, ldo, st));
    }
    CU(cudaStreamSynchronize(st)); // `rows` / `keep` are host temporaries of this call
    return HB_OK;
}

int hb_set_gexp_mats_dev(const double *sc, int64_t ld, int64_t G, int64_t N, const double *ref, int64_t ldr,
                         int64_t R, int32_t filter, double min_mean_both, const double *min_mean_test,
                         const double *min_mean_ref, int32_t scale, double *norm, int64_t ldo,
                         int32_t *keep_rows, int64_t *Gk, int32_t *info, void *stream)
{
    std::lock_guard<std::mutex> lk(g_mu);
    Workspace *ws;
    HBTRY(cur_ws(&ws));
    WS_USE(stream);
    if (!sc || !ref || !norm || !Gk || G < 1 || N < 1 || R < 1 || ld < N || ldr < R || ldo < N ||
        G >= ((int64_t)1 << 31) || N >= ((int64_t)1 << 32))
        return fail(HB_ERR_ARG, "bad arguments");
    return set_gexp_mats(ws, sc, ld, G, N, ref, ldr, R, filter, min_mean_both, min_mean_test, min_mean_ref,
                         scale, norm, ldo, keep_rows, Gk, info, (cudaStream_t)stream);
}

int hb_set_gexp_mats_host(const double *sc, int64_t G, int64_t N, const double *ref, int64_t R, int32_t filter,
                          double min_mean_both, const double *min_mean_test, const double *min_mean_ref,
                          int32_t scale, double *norm, int32_t *keep_rows, int64_t *Gk, int32_t *info)
{
    std::lock_guard<std::mutex> lk(g_mu);
    Workspace *ws;
    HBTRY(cur_ws(&ws));
    WS_USE(0);
    if (!sc || !ref || !norm || !Gk || G < 1 || N < 1 || R < 1 || G >= ((int64_t)1 << 31) ||
        N >= ((int64_t)1 << 32))
        return fail(HB_ERR_ARG, "bad arguments");
    const int64_t ld = (N + 127) / 128 * 128, ldr = (R + 31) / 32 * 32;
    // R's column-major G x N is [cell][gene] in memory: upload, turn gene-major, build, turn back
    HBTRY(ws->st_in.ensure((size_t)G * N * sizeof(double)));
    HBTRY(ws->st_in2.ensure((size_t)G * R * sizeof(double)));
    HBTRY(ws->st_x.ensure((size_t)G * ld * sizeof(double)));
    HBTRY(ws->st_x2.ensure((size_t)G * ldr * sizeof(double)));
    HBTRY(ws->st_g.ensure((size_t)G * ld * sizeof(double)));
    cudaStream_t st = 0;
    CU(cudaMemcpyAsync(ws->st_in.p, sc, (size_t)G * N * sizeof(double), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ws->st_in2.p, ref, (size_t)G * R * sizeof(double), cudaMemcpyHostToDevice, st));
    CU(hb::launch_transpose_f64(ws->st_in.as<double>(), G, N, ws->st_x.as<double>(), ld, st));
    CU(hb::launch_transpose_f64(ws->st_in2.as<double>(), G, R, ws->st_x2.as<double>(), ldr, st));
    HBTRY(set_gexp_mats(ws, ws->st_x.as<double>(), ld, G, N, ws->st_x2.as<double>(), ldr, R, filter,
                        min_mean_both, min_mean_test, min_mean_ref, scale, ws->st_g.as<double>(), ld,
                        keep_rows, Gk, info, st));
    if (*Gk == 0) return HB_OK;
    CU(hb::launch_transpose_f64_back(ws->st_g.as<double>(), ld, *Gk, N, ws->st_in.as<double>(), st));
    CU(cudaMemcpyAsync(norm, ws->st_in.p, (size_t)*Gk * N * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return HB_OK;
}

// setAlleleMats (R/HoneyBADGER.R:486-614) on gene-major int32 count matrices; caller holds g_mu.
static int set_allele_mats(Workspace *ws, const int32_t *r, const int32_t *n, int64_t ld, int64_t G, int64_t N,
                           const double *l_init, const double *n_bulk_init, int32_t filter, double thr,
                           int32_t min_cell, int32_t *r_maf, int32_t *n_sc, int64_t ldo, double *l_out,
                           double *n_bulk_out, double *l_maf, int32_t *keep_rows, int64_t *Gk_out,
                           int64_t *n_het, cudaStream_t st)
{
    // rowSums(r.init > 0), rowSums(n.sc.init > 0): the pooled-count kernel with one all-cells group
    HBTRY(ws->cellmask.ensure((size_t)N * sizeof(uint32_t)));
    HBTRY(ws->st_inp.ensure((size_t)2 * G * sizeof(int32_t) + (size_t)G * sizeof(int32_t) + (size_t)G + 64));
    int32_t *d_cnt = ws->st_inp.as<int32_t>(), *d_rows = d_cnt + 2 * G;
    uint8_t *d_code = (uint8_t *)(d_rows + G);
    CU(cudaMemsetAsync(ws->cellmask.p, 0x01, (size_t)N * sizeof(uint32_t), st));
    CU(hb::launch_pool_count(r, ld, G, N, ws->cellmask.as<uint32_t>(), 1, d_cnt, st));
    CU(hb::launch_pool_count(n, ld, G, N, ws->cellmask.as<uint32_t>(), 1, d_cnt + G, st));
    std::vector<int32_t> cnt((size_t)2 * G);
    CU(cudaMemcpyAsync(cnt.data(), d_cnt, cnt.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    const bool silico = !l_init || !n_bulk_init;
    std::vector<int32_t> rows;
    std::vector<uint8_t> code;
    rows.reserve((size_t)G);
    code.reserve((size_t)G);
    const double hi = 1 - thr;
    int64_t het = 0;
    for (int64_t g = 0; g < G; g++) {
        const double l = silico ? (double)cnt[g] : l_init[g];
        const double nb = silico ? (double)cnt[G + g] : n_bulk_init[g];
        const double E = l / nb; // NaN for 0/0: dropped by which(), coded "NA" without the filter
        if (filter) {
            if (!(E > thr && E < hi)) continue;
            het++;                               // what "heterozygous SNPs identified" reports (:527)
            if (cnt[G + g] < min_cell) continue; // rowSums(n.sc > 0) >= min.cell
        }
        const int64_t k = (int64_t)rows.size();
        rows.push_back((int32_t)g);
        const uint8_t cd = (E != E) ? 2 : (E <= 0.5 ? 0 : 1);
        code.push_back(cd);
        if (l_out) l_out[k] = l;
        if (n_bulk_out) n_bulk_out[k] = nb;
        if (l_maf) l_maf[k] = cd == 2 ? 0.0 : (cd == 0 ? l : nb - l);
    }
    const int64_t Gk = (int64_t)rows.size();
    *Gk_out = Gk;
    if (n_het) *n_het = filter ? het : Gk;
    if (keep_rows) memcpy(keep_rows, rows.data(), (size_t)Gk * sizeof(int32_t));
    if (Gk == 0) return HB_OK;
    CU(cudaMemcpyAsync(d_rows, rows.data(), (size_t)Gk * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_code, code.data(), (size_t)Gk, cudaMemcpyHostToDevice, st));
    CU(hb::launch_allele_flip(r, n, ld, N, d_rows, Gk, d_code, r_maf, n_sc, ldo, st));
    CU(cudaStreamSynchronize(st));
    return HB_OK;
}

int hb_set_allele_mats_dev(const int32_t *r, const int32_t *n_sc_init, int64_t ld, int64_t G, int64_t N,
                           const double *l_init, const double *n_bulk_init, int32_t filter,
                           double het_deviance_threshold, int32_t min_cell, int32_t *r_maf, int32_t *n_sc,
                           int64_t ldo, double *l, double *n_bulk, double *l_maf, int32_t *keep_rows,
                           int64_t *Gk, int64_t *n_het, void *stream)
{
    std::lock_guard<std::mutex> lk(g_mu);
    Workspace *ws;
    HBTRY(cur_ws(&ws));
    WS_USE(stream);
    if (!r || !n_sc_init || !r_maf || !Gk || G < 1 || N < 1 || ld < N || ldo < N || G >= ((int64_t)1 << 31))
        return fail(HB_ERR_ARG, "bad arguments");
    return set_allele_mats(ws, r, n_sc_init, ld, G, N, l_init, n_bulk_init, filter, het_deviance_threshold,
                           min_cell, r_maf, n_sc, ldo, l, n_bulk, l_maf, keep_rows, Gk, n_het,
                           (cudaStream_t)stream);
}

int hb_set_allele_mats_host(const int32_t *r, const int32_t *n_sc_init, int64_t G, int64_t N,
                            const double *l_init, const double *n_bulk_init, int32_t filter,
                            double het_deviance_threshold, int32_t min_cell, int32_t *r_maf, int32_t *n_sc,
                            double *l, double *n_bulk, double *l_maf, int32_t *keep_rows, int64_t *Gk,
                            int64_t *n_het)
{
    std::lock_guard<std::mutex> lk(g_mu);
    Workspace *ws;
    HBTRY(cur_ws(&ws));
    WS_USE(0);
    if (!r || !n_sc_init || !r_maf || !Gk || G < 1 || N < 1 || G >= ((int64_t)1 << 31))
        return fail(HB_ERR_ARG, "bad arguments");
    const int64_t ld = (N + 127) / 128 * 128;
    const size_t mat = (size_t)G * ld * sizeof(int32_t);
    HBTRY(ws->st_in.ensure((size_t)G * N * sizeof(int32_t)));
    HBTRY(ws->st_in2.ensure((size_t)G * N * sizeof(int32_t)));
    HBTRY(ws->st_x.ensure(mat));
    HBTRY(ws->st_x2.ensure(mat));
    HBTRY(ws->st_g.ensure(2 * mat));
    cudaStream_t st = 0;
    CU(cudaMemcpyAsync(ws->st_in.p, r, (size_t)G * N * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ws->st_in2.p, n_sc_init, (size_t)G * N * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CU(hb::launch_transpose_i32(ws->st_in.as<int32_t>(), G, N, ws->st_x.as<int32_t>(), ld, st));
    CU(hb::launch_transpose_i32(ws->st_in2.as<int32_t>(), G, N, ws->st_x2.as<int32_t>(), ld, st));
    int32_t *o_r = ws->st_g.as<int32_t>(), *o_n = o_r + (size_t)G * ld;
    HBTRY(set_allele_mats(ws, ws->st_x.as<int32_t>(), ws->st_x2.as<int32_t>(), ld, G, N, l_init, n_bulk_init,
                          filter, het_deviance_threshold, min_cell, o_r, n_sc ? o_n : nullptr, ld, l, n_bulk,
                          l_maf, keep_rows, Gk, n_het, st));
    if (*Gk == 0) return HB_OK;
    CU(hb::launch_transpose_i32_back(o_r, ld, *Gk, N, ws->st_in.as<int32_t>(), st));
    CU(cudaMemcpyAsync(r_maf, ws->st_in.p, (size_t)*Gk * N * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    if (n_sc) {
        CU(hb::launch_transpose_i32_back(o_n, ld, *Gk, N, ws->st_in2.as<int32_t>(), st));
        CU(cudaMemcpyAsync(n_sc, ws->st_in2.p, (size_t)*Gk * N * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    }
    CU(cudaStreamSynchronize(st));
    return HB_OK;
}

// ------------------------------------------------------------------ deterministic retest (8f-1)
int hb_retest_comb_dev(const double *x, int64_t ld, int64_t G, int64_t N, const int64_t *rows,
                       int64_t nrows, double mag0, const double *sigma0, const double *del_logodds,
                       double *p_amp, double *p_del, double *mu0, double *D_local, const double *D_total,
                       int32_t phase, void *stream)
{
    std::lock_guard<std::mutex> lk(g_mu);
    Workspace *ws;
    HBTRY(cur_ws(&ws));
    WS_USE(stream);
    if (!x || !rows || !p_amp || !p_del || G < 1 || N < 1 || nrows < 1 || ld < N || phase < 0 || phase > 2)
        return fail(HB_ERR_ARG, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t nb = (size_t)((N + 127) / 128);
    HBTRY(ws->st_part.ensure(nb * sizeof(double)));
    double *part = ws->st_part.as<double>();
    if (phase == 0 || phase == 1)
        CU(hb::launch_retest_gexp_a(x, ld, N, rows, nrows, mag0, sigma0, del_logodds, p_amp, p_del, mu0, part,
                                    D_local, st));
    if (phase == 0 || phase == 2)
        CU(hb::launch_retest_gexp_b(N, part, phase == 2 ? D_total : nullptr, p_amp, p_del, st));
    return HB_OK;
}

int hb_retest_gexp_dev(const double *x, int64_t ld, int64_t G, int64_t N, const int64_t *rows,
                       int64_t nrows, double mag0, const double *sigma0, double *p_amp, double *p_del,
                       double *mu0, double *D_local, const double *D_total, int32_t phase, void *stream)
{
    return hb_retest_comb_dev(x, ld, G, N, rows, nrows, mag0, sigma0, nullptr, p_amp, p_del, mu0, D_local,
                              D_total, phase, stream);
}

int hb_retest_gexp_host(const double *gexp_rows, int64_t n, int64_t N, double mag0, const double *sigma0,
                        double *p_amp, double *p_del, double *mu0)
{
    if (!gexp_rows || !p_amp || !p_del || n < 1 || N < 1) return fail(HB_ERR_ARG, "bad arguments");
    const int64_t ld = (N + 127) / 128 * 128;
    double *d_x, *d_out, *d_sig = nullptr;
    int64_t *d_rows;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        Workspace *ws;
        HBTRY(cur_ws(&ws));
        WS_USE(0);
        HBTRY(ws->st_in.ensure((size_t)n * N * sizeof(double)));
        HBTRY(ws->st_x.ensure((size_t)n * ld * sizeof(double)));
        HBTRY(ws->st_out.ensure((size_t)4 * N * sizeof(double)));
        HBTRY(ws->st_member.ensure((size_t)n * sizeof(int64_t)));
        CU(cudaMemcpyAsync(ws->st_in.p, gexp_rows, (size_t)n * N * sizeof(double), cudaMemcpyHostToDevice, 0));
        CU(hb::launch_transpose_f64(ws->st_in.as<double>(), n, N, ws->st_x.as<double>(), ld, 0));
        std::vector<int64_t> rows((size_t)n);
        std::iota(rows.begin(), rows.end(), (int64_t)0);
        CU(cudaMemcpy(ws->st_member.p, rows.data(), (size_t)n * sizeof(int64_t), cudaMemcpyHostToDevice));
        d_x = ws->st_x.as<double>();
        d_out = ws->st_out.as<double>();
        d_rows = ws->st_member.as<int64_t>();
        if (sigma0) {
            d_sig = d_out + 3 * N;
            CU(cudaMemcpy(d_sig, sigma0, (size_t)N * sizeof(double), cudaMemcpyHostToDevice));
        }
    }
    HBTRY(hb_retest_gexp_dev(d_x, ld, n, N, d_rows, n, mag0, d_sig, d_out, d_out + N, d_out + 2 * N, nullptr,
                             nullptr, 0, nullptr));
    CU(cudaMemcpy(p_amp, d_out, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(p_del, d_out + N, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost));
    if (mu0) CU(cudaMemcpy(mu0, d_out + 2 * N, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost));
    return HB_OK;
}

int hb_retest_allele_lr_dev(const int32_t *r_maf, const int32_t *n_sc, int64_t ld, int64_t G, int64_t N,
                            const int64_t *rows, int64_t nrows, double pd, double pn, double *p_del,
                            void *stream)
{
    if (!r_maf || !n_sc || !rows || !p_del || G < 1 || N < 1 || nrows < 1 || ld < N || !(pd > 0 && pd < 1) ||
        !(pn > 0 && pn < 1))
        return fail(HB_ERR_ARG, "bad arguments");
    CU(hb::launch_retest_allele(r_maf, n_sc, ld, N, rows, nrows, pd, pn, p_del, (cudaStream_t)stream));
    return HB_OK;
}

// Gauss-Legendre nodes / weights on [-1, 1] (Newton on P_n; n small).
static void gauss_legendre(int n, std::vector<double> &x, std::vector<double> &w)
{
    x.assign(n, 0.0);
    w.assign(n, 0.0);
    for (int i = 0; i < (n + 1) / 2; i++) {
        double z = cos(M_PI * (i + 0.75) / (n + 0.5)), pp = 1.0;
        for (int it = 0; it < 100; it++) {
            double p1 = 1.0, p2 = 0.0;
            for (int j = 0; j < n; j++) {
                const double p3 = p2;
                p2 = p1;
                p1 = ((2.0 * j + 1.0) * z * p2 - j * p3) / (j + 1.0);
            }
            pp = n * (z * p1 - p2) / (z * z - 1.0);
            const double z1 = z;
            z = z1 - p1 / pp;
            if (fabs(z - z1) < 1e-15) break;
        }
        x[i] = -z;
        x[n - 1 - i] = z;
        w[i] = w[n - 1 - i] = 2.0 / ((1.0 - z * z) * pp * pp);
    }
}

// Exact marginal of inst/bug/snpModel.bug per cell (see retest_allele_model_kernel).
int hb_retest_allele_model_dev(const int32_t *r_maf, const int32_t *n_sc, int64_t ld, int64_t G, int64_t N,
                               const int64_t *rows, int64_t nrows, const uint8_t *gene_end,
                               const double *l_maf, const double *n_bulk, double pseudo, double mono,
                               int32_t logodds, double *p_del, void *stream)
{
    std::lock_guard<std::mutex> lk(g_mu);
    Workspace *ws;
    HBTRY(cur_ws(&ws));
    WS_USE(stream);
    if (!r_maf || !n_sc || !rows || !l_maf || !n_bulk || !p_del || G < 1 || N < 1 || nrows < 1 || ld < N ||
        !(pseudo > 0 && pseudo < 0.5) || !(mono > 0 && mono < 1))
        return fail(HB_ERR_ARG, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    // h ~ dnorm(0.5, 0.1)T(pseudo, 1 - pseudo): JAGS' second argument is a precision (sd = 1/sqrt(0.1))
    const int panels = 16, gl = 16, nq = panels * gl;
    const double sd = 1.0 / sqrt(0.1), lo = pseudo, hi = 1.0 - pseudo;
    const double Z = sd * sqrt(2.0 * M_PI) * 0.5 * (erf((hi - 0.5) / (sd * M_SQRT2)) - erf((lo - 0.5) / (sd * M_SQRT2)));
    std::vector<double> gx, gw, tab((size_t)3 * nq);
    gauss_legendre(gl, gx, gw);
    const double pw = (hi - lo) / panels;
    for (int p = 0; p < panels; p++)
        for (int i = 0; i < gl; i++) {
            const double h = lo + pw * (p + 0.5 * (gx[i] + 1.0));
            const double dens = exp(-0.5 * (h - 0.5) * (h - 0.5) / (sd * sd)) / Z;
            const int q = p * gl + i;
            tab[q] = log(h);
            tab[nq + q] = log1p(-h);
            tab[2 * nq + q] = log(0.5 * pw * gw[i] * dens);
        }
    // bulk posterior of the allele indicator: l ~ Bin(n.bulk, fma), fma = pseudo (ma = 1) | 1 - pseudo (ma = 0)
    std::vector<double> lq((size_t)2 * nrows);
    const double lr = log((1.0 - pseudo) / pseudo);
    std::vector<uint8_t> ge((size_t)nrows, 1); // no gene factor: every SNP is its own gene
    if (gene_end) memcpy(ge.data(), gene_end, (size_t)nrows);
    ge[(size_t)nrows - 1] = 1;
    for (int64_t j = 0; j < nrows; j++) {
        const double d = (n_bulk[j] - 2.0 * l_maf[j]) * lr; // log odds of ma = 1
        lq[j] = d >= 0 ? -log1p(exp(-d)) : d - log1p(exp(d));
        lq[nrows + j] = d >= 0 ? -d - log1p(exp(-d)) : -log1p(exp(d));
    }
    const size_t bytes = ((size_t)3 * nq + 2 * nrows) * sizeof(double) + (size_t)nrows;
    HBTRY(ws->st_tab.ensure(bytes + 64));
    double *d_tab = ws->st_tab.as<double>(), *d_lq = d_tab + 3 * nq;
    uint8_t *d_ge = (uint8_t *)(d_lq + 2 * nrows);
    CU(cudaMemcpyAsync(d_tab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_lq, lq.data(), lq.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_ge, ge.data(), ge.size(), cudaMemcpyHostToDevice, st));
    CU(hb::launch_retest_allele_model(r_maf, n_sc, ld, N, rows, nrows, d_ge, d_lq, d_lq + nrows, d_tab, nq, pseudo,
                                      mono, logodds, p_del, st));
    CU(cudaStreamSynchronize(st)); // tab / lq / ge are host temporaries
    return HB_OK;
}

// Host form for the R side: r_rows, n_rows are the region's [n x N] column-major integer sub-matrices
// (`r.maf[bound.snps.new, ]`), gene[n] the gene factor codes (any integers; NULL = one gene per SNP).
int hb_retest_allele_model_host(const int32_t *r_rows, const int32_t *n_rows, int64_t n, int64_t N,
                                const double *l_maf, const double *n_bulk, const int32_t *gene, double pseudo,
                                double mono, double *p_del)
{
    if (!r_rows || !n_rows || !l_maf || !n_bulk || !p_del || n < 1 || N < 1) return fail(HB_ERR_ARG, "bad arguments");
    // group the SNPs by gene in order of first appearance (unique(geneFactor), R/HoneyBADGER.R:963)
    std::vector<int64_t> order((size_t)n);
    std::iota(order.begin(), order.end(), (int64_t)0);
    std::vector<uint8_t> ge((size_t)n, 1);
    if (gene) {
        std::vector<int32_t> seen;
        std::vector<int64_t> rank((size_t)n);
        for (int64_t j = 0; j < n; j++) {
            auto it = std::find(seen.begin(), seen.end(), gene[j]);
            rank[j] = it - seen.begin();
            if (it == seen.end()) seen.push_back(gene[j]);
        }
        std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return rank[a] < rank[b]; });
        for (int64_t j = 0; j < n; j++) ge[j] = (j + 1 == n) || rank[order[j + 1]] != rank[order[j]];
    }
    std::vector<double> l2((size_t)n), nb2((size_t)n);
    for (int64_t j = 0; j < n; j++) {
        l2[j] = l_maf[order[j]];
        nb2[j] = n_bulk[order[j]];
    }
    const int64_t ld = (N + 127) / 128 * 128;
    int32_t *d_r, *d_n;
    int64_t *d_rows;
    double *d_out;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        Workspace *ws;
        HBTRY(cur_ws(&ws));
        WS_USE(0);
        HBTRY(ws->st_in.ensure((size_t)n * N * sizeof(int32_t)));
        HBTRY(ws->st_in2.ensure((size_t)n * N * sizeof(int32_t)));
        HBTRY(ws->st_x.ensure((size_t)n * ld * sizeof(int32_t)));
        HBTRY(ws->st_x2.ensure((size_t)n * ld * sizeof(int32_t)));
        HBTRY(ws->st_out.ensure((size_t)N * sizeof(double)));
        HBTRY(ws->st_member.ensure((size_t)n * sizeof(int64_t)));
        CU(cudaMemcpyAsync(ws->st_in.p, r_rows, (size_t)n * N * sizeof(int32_t), cudaMemcpyHostToDevice, 0));
        CU(cudaMemcpyAsync(ws->st_in2.p, n_rows, (size_t)n * N * sizeof(int32_t), cudaMemcpyHostToDevice, 0));
        CU(hb::launch_transpose_i32(ws->st_in.as<int32_t>(), n, N, ws->st_x.as<int32_t>(), ld, 0));
        CU(hb::launch_transpose_i32(ws->st_in2.as<int32_t>(), n, N, ws->st_x2.as<int32_t>(), ld, 0));
        CU(cudaMemcpy(ws->st_member.p, order.data(), (size_t)n * sizeof(int64_t), cudaMemcpyHostToDevice));
        d_r = ws->st_x.as<int32_t>();
        d_n = ws->st_x2.as<int32_t>();
        d_rows = ws->st_member.as<int64_t>();
        d_out = ws->st_out.as<double>();
    }
    HBTRY(hb_retest_allele_model_dev(d_r, d_n, ld, n, N, d_rows, n, ge.data(), l2.data(), nb2.data(), pseudo, mono,
                                     0, d_out, nullptr));
    CU(cudaMemcpy(p_del, d_out, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost));
    return HB_OK;
}

// ------------------------------------------------------------------ few long chains
int hb_set_ward_mode(int32_t mode)
{
    if (mode < 0 || mode > 2) return fail(HB_ERR_ARG, "mode must be 0, 1 or 2");
    g_ward_mode = mode;
    return HB_OK;
}

int hb_set_dist_mode(int32_t mode)
{
    if (mode < 0 || mode > 1) return fail(HB_ERR_ARG, "mode must be 0 or 1");
    g_dist_mode = mode;
    return HB_OK;
}

int hb_set_longchain_mode(int32_t mode)
{
    if (mode < 0 || mode > 2) return fail(HB_ERR_ARG, "mode must be 0, 1 or 2");
    g_longchain_mode = mode;
    return HB_OK;
}

int hb_viterbi_long_logb_dev(const double *logb, int32_t S, int64_t T, int32_t K, const double *Pi,
                             const double *delta, int32_t block_len, int32_t *y, void *stream)
{
    std::lock_guard<std::mutex> lk(g_mu);
    Workspace *ws;
    HBTRY(cur_ws(&ws));
    WS_USE(stream);
    if (block_len < 0 || block_len > 4096) return fail(HB_ERR_ARG, "block_len out of range");
    return longchain_run(ws, S, logb, T, K, Pi, delta, y, (cudaStream_t)stream, true, block_len);
}

int hb_forwardback_long_logb_dev(const double *logb, int32_t S, int64_t T, int32_t K, const double *Pi,
                                 const double *delta, int32_t block_len, double *gamma, double *ll,
                                 void *stream)
{
    std::lock_guard<std::mutex> lk(g_mu);
    Workspace *ws;
    HBTRY(cur_ws(&ws));
    WS_USE(stream);
    if (block_len < 0 || block_len > 4096) return fail(HB_ERR_ARG, "block_len out of range");
    return longchain_fb_run(ws, S, logb, T, K, Pi, delta, gamma, ll, (cudaStream_t)stream, block_len);
}

int hb_longchain_stats_dev(void *stream, int64_t *out)
{
    std::lock_guard<std::mutex> lk(g_mu);
    Workspace *ws;
    HBTRY(cur_ws(&ws));
    WS_USE(stream);
    if (!out) return fail(HB_ERR_ARG, "out is NULL");
    for (int i = 0; i < 5; i++) out[i] = 0;
    if (!ws->lc_cstate || ws->lc_K < 1) return HB_OK;
    std::vector<int32_t> cs((size_t)ws->lc_K * HB_LC_CS);
    CU(cudaMemcpyAsync(cs.data(), ws->lc_cstate, cs.size() * sizeof(int32_t), cudaMemcpyDeviceToHost,
                       (cudaStream_t)stream));
    CU(cudaStreamSynchronize((cudaStream_t)stream));
    out[0] = ws->lc_K;
    out[1] = ws->lc_nb;
    for (int32_t c = 0; c < ws->lc_K; c++) {
        out[2] += cs[(size_t)c * HB_LC_CS + 1] == 2;             // chains finished on the sequential path
        out[3] = std::max<int64_t>(out[3], cs[(size_t)c * HB_LC_CS + 2]); // repair rounds used (max over chains)
        out[4] += cs[(size_t)c * HB_LC_CS + 3];                  // blocks run literally by the sequential walkers
    }
    return HB_OK;
}

// ------------------------------------------------------------------ the four reference call sites
// One round of pooled chains on DEVICE-RESIDENT matrices: membership and the small results are
// host buffers (K <= a few dozen chains), the [G x N] matrices never leave the GPU.
int hb_gexp_pooled_viterbi_dev(const double *x, int64_t ld, int64_t G, int64_t N,
                               const uint8_t *member, int32_t K, const double *mean,
                               const double *Pi, const double *delta, int32_t *y, double *pooled_x,
                               double *pooled_sd, void *stream)
{
    std::lock_guard<std::mutex> lk(g_mu);
    Workspace *ws;
    HBTRY(cur_ws(&ws));
    WS_USE(stream);
    if (!x || !member || !y || G < 2 || N < 1 || K < 1 || ld < N)
        return fail(HB_ERR_ARG, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<int32_t> sz;
    HBTRY(group_sizes(member, K, N, sz));
    const int64_t ldk = pad_ld(K);
    HBTRY(ws->st_member.ensure((size_t)K * N));
    HBTRY(ws->st_pool.ensure((size_t)K * G * sizeof(double)));    // [K][G]
    HBTRY(ws->st_pool2.ensure((size_t)G * ldk * sizeof(double))); // gene-major [G][ldk]
    HBTRY(ws->st_sd.ensure((size_t)K * sizeof(double)));
    HBTRY(ws->st_y.ensure((size_t)G * ldk));
    HBTRY(ws->st_out.ensure((size_t)G * K * sizeof(int32_t)));
    CU(cudaMemcpyAsync(ws->st_member.p, member, (size_t)K * N, cudaMemcpyHostToDevice, st));
    HBTRY(pool_mean_launch(ws, x, ld, G, N, ws->st_member.as<uint8_t>(), sz, K, ws->st_pool.as<double>(), st));
    CU(hb::launch_pooled_sd(ws->st_pool.as<double>(), G, K, ws->st_sd.as<double>(), st));
    // [K][G] row-major is column-major [G x K]: the same transpose as an R matrix upload
    CU(hb::launch_transpose_f64(ws->st_pool.as<double>(), G, K, ws->st_pool2.as<double>(), ldk, st));
    // log(sd) must be the host libm's (R evaluates it on the CPU): finish the K scales here
    std::vector<double> h_sd(K), h_aux(3 * (size_t)K);
    CU(cudaMemcpyAsync(h_sd.data(), ws->st_sd.p, (size_t)K * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st)); // also keeps `sz` alive until its copy is done
    for (int32_t k = 0; k < K; k++) {
        h_aux[k] = h_sd[k];
        h_aux[K + k] = 1.0 / h_sd[k];
        h_aux[2 * K + k] = log(h_sd[k]);
    }
    HBTRY(ws->sdaux.ensure(h_aux.size() * sizeof(double)));
    CU(cudaMemcpy(ws->sdaux.p, h_aux.data(), h_aux.size() * sizeof(double), cudaMemcpyHostToDevice));
    const int64_t seg[2] = {0, G};
    const double *aux = ws->sdaux.as<double>();
    if (g_longchain_mode == 2) {
        hb::Norm3Const nc;
        bool sym;
        memset(&nc, 0, sizeof nc);
        HBTRY(fill_norm3_const(nc, mean, Pi, delta, &sym));
        HBTRY(ws->lc_lb.ensure((size_t)K * G * 3 * sizeof(double)));
        CU(hb::launch_norm3_emit_cm(nc, ws->st_pool.as<double>(), G, K, aux, aux + K, aux + 2 * K,
                                    ws->lc_lb.as<double>(), st));
        HBTRY(longchain_run(ws, 3, ws->lc_lb.as<double>(), G, K, Pi, delta, ws->st_out.as<int32_t>(), st));
    } else {
        HBTRY(norm3_run(ws, ws->st_pool2.as<double>(), ldk, G, K, seg, 1, mean, aux, aux + K, aux + 2 * K,
                        0, Pi, delta, HB_WANT_VITERBI, ws->st_y.as<uint8_t>(), ldk, nullptr, 0, nullptr,
                        st));
        CU(hb::launch_transpose_u8_back(ws->st_y.as<uint8_t>(), ldk, G, K, ws->st_out.as<int32_t>(), st));
    }
    CU(cudaMemcpyAsync(y, ws->st_out.p, (size_t)G * K * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    if (pooled_x)
        CU(cudaMemcpyAsync(pooled_x, ws->st_pool.p, (size_t)K * G * sizeof(double),
                           cudaMemcpyDeviceToHost, st));
    if (pooled_sd) memcpy(pooled_sd, h_sd.data(), (size_t)K * sizeof(double));
    return read_status(ws, st);
}

int hb_allele_pooled_viterbi_dev(const int32_t *r_maf, const int32_t *n_sc, int64_t ld, int64_t G,
                                 int64_t N, const uint8_t *member, int32_t K, const double *prob,
                                 const double *Pi, const double *delta, int32_t *y, int32_t *mafl,
                                 int32_t *sizel, void *stream)
{
    std::lock_guard<std::mutex> lk(g_mu);
    Workspace *ws;
    HBTRY(cur_ws(&ws));
    WS_USE(stream);
    if (!r_maf || !n_sc || !member || !y || G < 2 || N < 1 || K < 1 || ld < N)
        return fail(HB_ERR_ARG, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<int32_t> sz;
    HBTRY(group_sizes(member, K, N, sz));
    const int64_t ldk = (K + 31) / 32 * 32;
    HBTRY(ws->st_member.ensure((size_t)K * N));
    HBTRY(ws->cellmask.ensure((size_t)N * sizeof(uint32_t)));
    HBTRY(ws->st_pool.ensure((size_t)K * G * sizeof(int32_t) * 2));
    HBTRY(ws->st_pool2.ensure((size_t)G * ldk * sizeof(int32_t) * 2));
    HBTRY(ws->st_y.ensure((size_t)G * ldk));
    HBTRY(ws->st_out.ensure((size_t)G * K * sizeof(int32_t)));
    CU(cudaMemcpyAsync(ws->st_member.p, member, (size_t)K * N, cudaMemcpyHostToDevice, st));
    int32_t *cnt[2] = {ws->st_pool.as<int32_t>(), ws->st_pool.as<int32_t>() + (int64_t)K * G};
    int32_t *gm[2] = {ws->st_pool2.as<int32_t>(), ws->st_pool2.as<int32_t>() + G * ldk};
    const int32_t *src[2] = {r_maf, n_sc};
    for (int i = 0; i < 2; i++) {
        for (int32_t k0 = 0; k0 < K; k0 += 32) {
            const int32_t kn = std::min(32, K - k0);
            CU(hb::launch_build_cellmask(ws->st_member.as<uint8_t>() + (int64_t)k0 * N, N, kn,
                                         ws->cellmask.as<uint32_t>(), st));
            CU(hb::launch_pool_count(src[i], ld, G, N, ws->cellmask.as<uint32_t>(), kn,
                                     cnt[i] + (int64_t)k0 * G, st));
        }
    }
    if (g_longchain_mode != 1) {
        // a handful of chains of G positions: parallel along the gene axis (hb_longchain.cuh), straight
        // from the chain-major pooled counts into the [G x K] result
        HBTRY(binom2_longchain(ws, cnt[0], cnt[1], G, K, prob, Pi, delta, ws->st_out.as<int32_t>(), st));
    } else {
        for (int i = 0; i < 2; i++) CU(hb::launch_transpose_i32(cnt[i], G, K, gm[i], ldk, st));
        const int64_t seg[2] = {0, G};
        HBTRY(binom2_run(ws, gm[0], gm[1], ldk, G, K, seg, 1, prob, 0.0, Pi, delta, HB_WANT_VITERBI,
                         ws->st_y.as<uint8_t>(), ldk, nullptr, 0, nullptr, st, true, true));
        CU(hb::launch_transpose_u8_back(ws->st_y.as<uint8_t>(), ldk, G, K, ws->st_out.as<int32_t>(), st));
    }
    CU(cudaMemcpyAsync(y, ws->st_out.p, (size_t)G * K * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    if (mafl)
        CU(cudaMemcpyAsync(mafl, cnt[0], (size_t)K * G * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    if (sizel)
        CU(cudaMemcpyAsync(sizel, cnt[1], (size_t)K * G * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    return read_status(ws, st);
}

// Host-buffer forms: upload the R matrices (column-major doubles), change layout, run the round.
int hb_gexp_pooled_viterbi_host(const double *gexp_norm, int64_t G, int64_t N,
                                const uint8_t *member, int32_t K, const double *mean,
                                const double *Pi, const double *delta, int32_t *y,
                                double *pooled_x, double *pooled_sd)
{
    if (!gexp_norm || G < 2 || N < 1) return fail(HB_ERR_ARG, "bad arguments");
    const int64_t ld = pad_ld(N);
    double *d_x = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        Workspace *ws;
        HBTRY(cur_ws(&ws));
        WS_USE(0);
        HBTRY(ws->st_in.ensure((size_t)G * N * sizeof(double)));
        HBTRY(ws->st_x.ensure((size_t)G * ld * sizeof(double)));
        CU(cudaMemcpyAsync(ws->st_in.p, gexp_norm, (size_t)G * N * sizeof(double),
                           cudaMemcpyHostToDevice, 0));
        CU(hb::launch_transpose_f64(ws->st_in.as<double>(), G, N, ws->st_x.as<double>(), ld, 0));
        d_x = ws->st_x.as<double>();
    }
    return hb_gexp_pooled_viterbi_dev(d_x, ld, G, N, member, K, mean, Pi, delta, y, pooled_x,
                                      pooled_sd, nullptr);
}

int hb_allele_pooled_viterbi_host(const double *r_maf, const double *n_sc, int64_t G, int64_t N,
                                  const uint8_t *member, int32_t K, const double *prob,
                                  const double *Pi, const double *delta, int32_t *y, int32_t *mafl,
                                  int32_t *sizel)
{
    if (!r_maf || !n_sc || G < 2 || N < 1) return fail(HB_ERR_ARG, "bad arguments");
    const int64_t ld = (N + 31) / 32 * 32;
    int32_t *d_r = nullptr, *d_n = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        Workspace *ws;
        HBTRY(cur_ws(&ws));
        WS_USE(0);
        HBTRY(ws->st_in.ensure((size_t)G * N * sizeof(double)));
        HBTRY(ws->st_x.ensure((size_t)G * ld * sizeof(int32_t)));
        HBTRY(ws->st_x2.ensure((size_t)G * ld * sizeof(int32_t)));
        CU(cudaMemcpyAsync(ws->st_in.p, r_maf, (size_t)G * N * sizeof(double), cudaMemcpyHostToDevice, 0));
        CU(hb::launch_transpose_f64_i32(ws->st_in.as<double>(), G, N, ws->st_x.as<int32_t>(), ld, 0));
        CU(cudaMemcpyAsync(ws->st_in.p, n_sc, (size_t)G * N * sizeof(double), cudaMemcpyHostToDevice, 0));
        CU(hb::launch_transpose_f64_i32(ws->st_in.as<double>(), G, N, ws->st_x2.as<int32_t>(), ld, 0));
        d_r = ws->st_x.as<int32_t>();
        d_n = ws->st_x2.as<int32_t>();
    }
    return hb_allele_pooled_viterbi_dev(d_r, d_n, ld, G, N, member, K, prob, Pi, delta, y, mafl,
                                        sizel, nullptr);
}


// ------------------------------------------------------------------ device-resident handle
// The state the reference keeps across its recursive calls is the matrices themselves
// (R/HoneyBADGER.R:1098-1102; sub-matrices handed down at :1304-1309).  Here they are uploaded ONCE and stay
// on the device behind an opaque handle; every later call names rows / cells by index sets.
} // extern "C"

struct hb_handle {
    int32_t kind = 0; // 1 expression (fp64), 2 allele counts (two int32 planes)
    int dev = 0;
    int64_t G = 0, N = 0, ld = 0, h2d_bytes = 0;
    DevBuf a, b;         // resident gene-major matrices
    DevBuf sa, sb, idx;  // gathered sub-matrix of the current call, index sets
    int64_t sub_ld = 0;
};

static int handle_check(const hb_handle *h, int kind)
{
    if (!h || h->kind != kind) return fail(HB_ERR_ARG, "not a %s handle", kind == 1 ? "gexp" : "counts");
    int dev = -1;
    cudaGetDevice(&dev);
    if (dev != h->dev) return fail(HB_ERR_ARG, "handle lives on device %d, current device is %d", h->dev, dev);
    return HB_OK;
}

// column slabs: upload [G x n] column-major pieces and transpose them into the resident rows
template <typename T>
static int upload_slabs(Workspace *ws, const double *src, int64_t G, int64_t N, T *dst, int64_t ld, int64_t *bytes)
{
    const int64_t slab = std::max<int64_t>(1, std::min<int64_t>(N, ((int64_t)256 << 20) / (G * (int64_t)sizeof(double))));
    HBTRY(ws->st_in.ensure((size_t)G * slab * sizeof(double)));
    for (int64_t c0 = 0; c0 < N; c0 += slab) {
        const int64_t n = std::min(slab, N - c0);
        CU(cudaMemcpyAsync(ws->st_in.p, src + c0 * G, (size_t)G * n * sizeof(double), cudaMemcpyHostToDevice, 0));
        if (sizeof(T) == 8)
            CU(hb::launch_transpose_f64(ws->st_in.as<double>(), G, n, (double *)dst + c0, ld, 0));
        else
            CU(hb::launch_transpose_f64_i32(ws->st_in.as<double>(), G, n, (int32_t *)dst + c0, ld, 0));
        *bytes += G * n * (int64_t)sizeof(double);
    }
    CU(cudaStreamSynchronize(0));
    return HB_OK;
}

extern "C" {

int hb_upload_gexp_host(const double *gexp_norm, int64_t G, int64_t N, hb_handle **out)
{
    std::lock_guard<std::mutex> lk(g_mu);
    Workspace *ws;
    HBTRY(cur_ws(&ws));
    WS_USE(0);
    if (!gexp_norm || !out || G < 1 || N < 1) return fail(HB_ERR_ARG, "bad arguments");
    hb_handle *h = new hb_handle;
    h->kind = 1;
    cudaGetDevice(&h->dev);
    h->G = G;
    h->N = N;
    h->ld = (N + 127) / 128 * 128;
    int rc = h->a.ensure((size_t)G * h->ld * sizeof(double));
    if (rc == HB_OK) rc = upload_slabs<double>(ws, gexp_norm, G, N, h->a.as<double>(), h->ld, &h->h2d_bytes);
    if (rc != HB_OK) {
        h->a.release();
        delete h;
        return rc;
    }
    *out = h;
    return HB_OK;
}

int hb_upload_counts_host(const double *r_maf, const double *n_sc, int64_t G, int64_t N, hb_handle **out)
{
    std::lock_guard<std::mutex> lk(g_mu);
    Workspace *ws;
    HBTRY(cur_ws(&ws));
    WS_USE(0);
    if (!r_maf || !n_sc || !out || G < 1 || N < 1) return fail(HB_ERR_ARG, "bad arguments");
    hb_handle *h = new hb_handle;
    h->kind = 2;
    cudaGetDevice(&h->dev);
    h->G = G;
    h->N = N;
    h->ld = (N + 127) / 128 * 128;
    int rc = h->a.ensure((size_t)G * h->ld * sizeof(int32_t));
    if (rc == HB_OK) rc = h->b.ensure((size_t)G * h->ld * sizeof(int32_t));
    if (rc == HB_OK) rc = upload_slabs<int32_t>(ws, r_maf, G, N, h->a.as<int32_t>(), h->ld, &h->h2d_bytes);
    if (rc == HB_OK) rc = upload_slabs<int32_t>(ws, n_sc, G, N, h->b.as<int32_t>(), h->ld, &h->h2d_bytes);
    if (rc != HB_OK) {
        h->a.release();
        h->b.release();
        delete h;
        return rc;
    }
    *out = h;
    return HB_OK;
}

int hb_free_handle(hb_handle *h)
{
    if (!h) return HB_OK;
    std::lock_guard<std::mutex> lk(g_mu);
    int cur = -1;
    cudaGetDevice(&cur);
    if (cur != h->dev) cudaSetDevice(h->dev);
    cudaDeviceSynchronize();
    DevBuf *all[] = {&h->a, &h->b, &h->sa, &h->sb, &h->idx};
    for (DevBuf *b : all) b->release();
    if (cur != h->dev && cur >= 0) cudaSetDevice(cur);
    delete h;
    return HB_OK;
}

int hb_handle_info(const hb_handle *h, int32_t *kind, int64_t *G, int64_t *N, int64_t *h2d_bytes)
{
    if (!h) return fail(HB_ERR_ARG, "handle is NULL");
    if (kind) *kind = h->kind;
    if (G) *G = h->G;
    if (N) *N = h->N;
    if (h2d_bytes) *h2d_bytes = h->h2d_bytes;
    return HB_OK;
}

// rows / cols (host, 0-based, NULL = all) -> gathered sub-matrix in the handle's own buffers; the index sets
// are the only bytes that cross PCIe.  Caller holds g_mu.
static int handle_gather(Workspace *ws, hb_handle *h, const int32_t *rows, int64_t nrows, const int32_t *cols,
                         int64_t ncols, int64_t *G2, int64_t *N2, cudaStream_t st)
{
    *G2 = rows ? nrows : h->G;
    *N2 = cols ? ncols : h->N;
    if (*G2 < 1 || *N2 < 1) return fail(HB_ERR_ARG, "empty row or cell set");
    for (int64_t i = 0; rows && i < nrows; i++)
        if (rows[i] < 0 || rows[i] >= h->G) return fail(HB_ERR_ARG, "row index %d outside [0, %lld)", rows[i], (long long)h->G);
    for (int64_t j = 0; cols && j < ncols; j++)
        if (cols[j] < 0 || cols[j] >= h->N) return fail(HB_ERR_ARG, "cell index %d outside [0, %lld)", cols[j], (long long)h->N);
    h->sub_ld = (*N2 + 127) / 128 * 128;
    const size_t el = h->kind == 1 ? sizeof(double) : sizeof(int32_t);
    HBTRY(h->sa.ensure((size_t)*G2 * h->sub_ld * el));
    if (h->kind == 2) HBTRY(h->sb.ensure((size_t)*G2 * h->sub_ld * el));
    HBTRY(h->idx.ensure((size_t)(*G2 + *N2) * sizeof(int32_t)));
    int32_t *d_rows = rows ? h->idx.as<int32_t>() : nullptr, *d_cols = cols ? h->idx.as<int32_t>() + *G2 : nullptr;
    if (rows) HBTRY(upload_async(ws, d_rows, rows, (size_t)nrows * sizeof(int32_t), st));
    if (cols) HBTRY(upload_async(ws, d_cols, cols, (size_t)ncols * sizeof(int32_t), st));
    CU(hb::launch_gather_rc(h->a.p, (int32_t)el, h->ld, d_rows, *G2, d_cols, *N2, h->sa.p, h->sub_ld, st));
    if (h->kind == 2)
        CU(hb::launch_gather_rc(h->b.p, (int32_t)el, h->ld, d_rows, *G2, d_cols, *N2, h->sb.p, h->sub_ld, st));
    return HB_OK;
}

int hb_gexp_round_handle(hb_handle *h, const int32_t *rows, int64_t nrows, const int32_t *cols, int64_t ncols,
                         const uint8_t *member, int32_t K, const double *mean, const double *Pi,
                         const double *delta, int32_t *y, double *pooled_x, double *pooled_sd)
{
    int64_t G2, N2;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        Workspace *ws;
        HBTRY(cur_ws(&ws));
        HBTRY(handle_check(h, 1));
        WS_USE(0);
        HBTRY(handle_gather(ws, h, rows, nrows, cols, ncols, &G2, &N2, 0));
    }
    return hb_gexp_pooled_viterbi_dev(h->sa.as<double>(), h->sub_ld, G2, N2, member, K, mean, Pi, delta, y, pooled_x,
                                      pooled_sd, nullptr);
}

int hb_allele_round_handle(hb_handle *h, const int32_t *rows, int64_t nrows, const int32_t *cols, int64_t ncols,
                           const uint8_t *member, int32_t K, const double *prob, const double *Pi,
                           const double *delta, int32_t *y, int32_t *mafl, int32_t *sizel)
{
    int64_t G2, N2;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        Workspace *ws;
        HBTRY(cur_ws(&ws));
        HBTRY(handle_check(h, 2));
        WS_USE(0);
        HBTRY(handle_gather(ws, h, rows, nrows, cols, ncols, &G2, &N2, 0));
    }
    return hb_allele_pooled_viterbi_dev(h->sa.as<int32_t>(), h->sb.as<int32_t>(), h->sub_ld, G2, N2, member, K, prob,
                                        Pi, delta, y, mafl, sizel, nullptr);
}

int hb_group_handle(hb_handle *h, const int32_t *rows, int64_t nrows, const int32_t *cols, int64_t ncols,
                    int32_t window, const int32_t *ks, int32_t nk, int32_t *labels)
{
    std::lock_guard<std::mutex> lk(g_mu);
    Workspace *ws;
    HBTRY(cur_ws(&ws));
    if (!h || (h->kind != 1 && h->kind != 2)) return fail(HB_ERR_ARG, "bad handle");
    HBTRY(handle_check(h, h->kind));
    if (!ks || nk < 1 || !labels || window < 1 || (window & 1) == 0) return fail(HB_ERR_ARG, "bad arguments");
    WS_USE(0);
    int64_t G2, N2;
    HBTRY(handle_gather(ws, h, rows, nrows, cols, ncols, &G2, &N2, 0));
    if (h->kind == 1)
        return group_core(ws, h->sa.as<double>(), nullptr, nullptr, h->sub_ld, G2, N2, window, ks, nk, labels, 0);
    return group_core(ws, nullptr, h->sa.as<int32_t>(), h->sb.as<int32_t>(), h->sub_ld, G2, N2, window, ks, nk, labels, 0);
}

int hb_retest_gexp_handle(hb_handle *h, const int32_t *rows, int64_t nrows, const int32_t *cols, int64_t ncols,
                          double mag0, const double *sigma0, double *p_amp, double *p_del, double *mu0)
{
    int64_t G2, N2;
    double *d_out;
    int64_t *d_rows;
    double *d_sig = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        Workspace *ws;
        HBTRY(cur_ws(&ws));
        HBTRY(handle_check(h, 1));
        if (!p_amp || !p_del) return fail(HB_ERR_ARG, "bad arguments");
        WS_USE(0);
        HBTRY(handle_gather(ws, h, rows, nrows, cols, ncols, &G2, &N2, 0));
        HBTRY(ws->st_out.ensure((size_t)4 * N2 * sizeof(double)));
        HBTRY(ws->st_member.ensure((size_t)G2 * sizeof(int64_t)));
        std::vector<int64_t> iota_rows((size_t)G2);
        std::iota(iota_rows.begin(), iota_rows.end(), (int64_t)0);
        HBTRY(upload_async(ws, ws->st_member.p, iota_rows.data(), (size_t)G2 * sizeof(int64_t), 0));
        d_out = ws->st_out.as<double>();
        d_rows = ws->st_member.as<int64_t>();
        if (sigma0) {
            d_sig = d_out + 3 * N2;
            HBTRY(upload_async(ws, d_sig, sigma0, (size_t)N2 * sizeof(double), 0));
        }
    }
    HBTRY(hb_retest_gexp_dev(h->sa.as<double>(), h->sub_ld, G2, N2, d_rows, G2, mag0, d_sig, d_out, d_out + N2,
                             d_out + 2 * N2, nullptr, nullptr, 0, nullptr));
    CU(cudaMemcpy(p_amp, d_out, (size_t)N2 * sizeof(double), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(p_del, d_out + N2, (size_t)N2 * sizeof(double), cudaMemcpyDeviceToHost));
    if (mu0) CU(cudaMemcpy(mu0, d_out + 2 * N2, (size_t)N2 * sizeof(double), cudaMemcpyDeviceToHost));
    return HB_OK;
}

int hb_retest_allele_handle(hb_handle *h, const int32_t *rows, int64_t nrows, const int32_t *cols, int64_t ncols,
                            const double *l_maf, const double *n_bulk, const uint8_t *gene_end, double pseudo,
                            double mono, double *p_del)
{
    int64_t G2, N2;
    double *d_out;
    int64_t *d_rows;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        Workspace *ws;
        HBTRY(cur_ws(&ws));
        HBTRY(handle_check(h, 2));
        if (!l_maf || !n_bulk || !p_del) return fail(HB_ERR_ARG, "bad arguments");
        WS_USE(0);
        HBTRY(handle_gather(ws, h, rows, nrows, cols, ncols, &G2, &N2, 0));
        HBTRY(ws->st_out.ensure((size_t)N2 * sizeof(double)));
        HBTRY(ws->st_member.ensure((size_t)G2 * sizeof(int64_t)));
        std::vector<int64_t> iota_rows((size_t)G2);
        std::iota(iota_rows.begin(), iota_rows.end(), (int64_t)0);
        HBTRY(upload_async(ws, ws->st_member.p, iota_rows.data(), (size_t)G2 * sizeof(int64_t), 0));
        d_out = ws->st_out.as<double>();
        d_rows = ws->st_member.as<int64_t>();
    }
    HBTRY(hb_retest_allele_model_dev(h->sa.as<int32_t>(), h->sb.as<int32_t>(), h->sub_ld, G2, N2, d_rows, G2, gene_end,
                                     l_maf, n_bulk, pseudo, mono, 0, d_out, nullptr));
    CU(cudaMemcpy(p_del, d_out, (size_t)N2 * sizeof(double), cudaMemcpyDeviceToHost));
    return HB_OK;
}

} // extern "C"
