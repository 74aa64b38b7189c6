// common.cuh — context, error plumbing and device-side reduction helpers shared by all kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <vector>
#include "../../include/lsob200.h"

#define LSO_NUM_SMS_DEFAULT 148
#define LSO_MAX_DEVICES 64

struct lso_ctx {
    int device = 0;
    int num_sms = LSO_NUM_SMS_DEFAULT;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;    // host-to-device copies that run under the factorisation (lso_qr_factor_keep_host)
    cudaEvent_t copy_ev[18] = {};          // [0] = compute stream -> copy stream, [1 + k] = chunk k has landed
    // scratch for two-stage deterministic reductions
    double* d_partials = nullptr;      // LSO_PARTIALS doubles
    unsigned int* d_counters = nullptr; // ticket counters for last-block reductions (zeroed)
    double* d_scalars = nullptr;       // small device scalar slots
    double* h_scalars = nullptr;       // pinned mirror
    int64_t launches = 0;
    // statistics (lso_ctx_stat): algorithmic work issued since the last reset
    double stat_qr_update_flops = 0.0;  // trailing-update flops of the QR factorisations: sum over panels of 4*QB*rows*trailing columns
    double stat_qr_flops = 0.0;         // 2 M n^2 - 2/3 n^3 per factorisation
    double stat_syrk_flops = 0.0;       // m n (n + 1) per J'J
    double stat_syrk_i8_macs = 0.0;     // int8 multiply-accumulates issued to tcgen05 by the digit-matrix syrk
    double stat_spmv_bytes = 0.0;       // 12 B per stored entry + 8 B per vector element, per sparse product
    std::string last_error;
    // options
    int64_t opt_qr_apply = 2;          // 0 = plain-FMA apply kernel, 1 = DMMA apply kernel, 2 = ping-pong DMMA kernel (one launch
                                       // per tree level), 3 = ping-pong kernel, all tree levels of a panel in one launch
                                       // (correct, but 12 % slower at C2 — per-level tails add up), 4 = levels 0, 1 one launch each + the
                                       // latency-bound levels above chained in one launch (saves 0.05 ms per solve at C2)
    int64_t opt_syrk = 3;              // 0 = plain syrk, 1 = DMMA syrk, 2 = tcgen05 int8 digit products (Ozaki scheme, ozaki.cu),
                                       // 3 = tcgen05 for m >= 8192 and n >= 512, DMMA otherwise (default)
    int64_t opt_ozaki_slices = 8;      // 7-bit digits per fp64 value in the tcgen05 syrk (2..8; 8 = below fp64 rounding)
    int64_t opt_qr_tune = 1;           // 1 = small QR plans time their launch schedules once and keep the fastest (cleared by
                                       // an explicit "qr_apply" / "qr_lookahead" setting)
    int64_t opt_qr_shard_pipeline = 1; // row-sharded QR: 1 = the stack QR runs panel by panel behind the local one (one small
                                       // all-gather per panel), 0 = local QR, one all-gather, stack QR
    int64_t opt_qr_twin = 2;           // host-fed chunked factorisation: extra workspaces / streams the chunks rotate through
    int64_t opt_qr_lookahead = 0;      // 1 = panel trees on a second stream under the previous update (+2% at C2)
    int64_t opt_trisolve = 1;          // 1 = multi-CTA triangular solves (mailbox-chained diagonal blocks), 0 = one CTA
    void* d_trimail = nullptr;         // mailbox of the multi-CTA triangular solve: 24 032 x {lo32, tag, hi32, tag}
    unsigned tri_tag = 0, tri_ticket_base = 0;
    int64_t opt_spmv = 2;              // 0 = first-generation sparse products, 1 = stream kernels (shared-memory staging),
                                       // 2 = warp kernels (registers + shuffles, persistent grid; default)
    int64_t opt_lsmr_fused = 1;        // 0 = LSMR scalars on the host, 1 = fused device-resident LSMR (sparse operator)
    int64_t opt_profile = 0;           // 1 = bracket every launch of the dominant kernel with CUDA events
    std::vector<cudaEvent_t> prof_events;   // pairs (begin, end)
    size_t prof_used = 0;
    std::vector<cudaEvent_t> prof2_events;  // second channel: the collective of a sharded solve (all-reduce / all-gather)
    size_t prof2_used = 0;
    // scratch of the rank-revealing small-R finish (qr_finish.cu), grown on demand
    double* d_finish = nullptr;
    size_t finish_cap = 0;
    double* d_gemv = nullptr;          // column-split gemv partial row sums
    size_t gemv_cap = 0;
    // NCCL (lazily loaded)
    void* nccl_comm = nullptr;
    int nranks = 1;
    int rank = 0;
};

#define LSO_PARTIALS (1 << 20)
#define LSO_NSCALARS 256

extern std::string g_lso_last_error;

int lso_set_error(lso_ctx* ctx, int code, const char* fmt, ...);

#define LSO_CHECK_CUDA(ctx, expr)                                                               \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            return lso_set_error((ctx), LSO_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,           \
                                 cudaGetErrorString(_e), __FILE__, __LINE__);                   \
    } while (0)

#define LSO_CHECK_LAUNCH(ctx)                                                                   \
    do {                                                                                        \
        (ctx)->launches++;                                                                      \
        cudaError_t _e = cudaGetLastError();                                                    \
        if (_e != cudaSuccess)                                                                  \
            return lso_set_error((ctx), LSO_ERR_CUDA, "kernel launch failed: %s (%s:%d)",       \
                                 cudaGetErrorString(_e), __FILE__, __LINE__);                   \
    } while (0)

#define LSO_REQUIRE(ctx, cond, msg)                                                             \
    do {                                                                                        \
        if (!(cond)) return lso_set_error((ctx), LSO_ERR_ARG, "%s (%s:%d)", msg, __FILE__, __LINE__); \
    } while (0)

// Every entry point makes its context's device current first: a process may hold contexts on several GPUs
// (cudaSetDevice is a no-op costing ~50 ns when the device is already current).
#define LSO_ENTER(ctx)                                                                          \
    do {                                                                                        \
        if ((ctx) != nullptr) {                                                                 \
            cudaError_t _e = cudaSetDevice((ctx)->device);                                      \
            if (_e != cudaSuccess)                                                              \
                return lso_set_error((ctx), LSO_ERR_CUDA, "cudaSetDevice(%d) failed: %s", (ctx)->device, cudaGetErrorString(_e)); \
        }                                                                                       \
    } while (0)

#define LSO_TRY(expr)                                                                           \
    do {                                                                                        \
        int _s = (expr);                                                                        \
        if (_s != 0) return _s;                                                                 \
    } while (0)

// CUDA-event bracketing of the dominant kernel's launches (bench.py roofline: live, on the launching stream)
static inline void lso_prof_mark(lso_ctx* ctx) {
    if (!ctx->opt_profile) return;
    if (ctx->prof_used == ctx->prof_events.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        ctx->prof_events.push_back(e);
    }
    cudaEventRecord(ctx->prof_events[ctx->prof_used++], ctx->stream);
}

static inline void lso_prof_mark2(lso_ctx* ctx) {
    if (!ctx->opt_profile) return;
    if (ctx->prof2_used == ctx->prof2_events.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        ctx->prof2_events.push_back(e);
    }
    cudaEventRecord(ctx->prof2_events[ctx->prof2_used++], ctx->stream);
}

static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t roundup64(int64_t a, int64_t b) { return cdiv64(a, b) * b; }

#ifdef __CUDACC__
// ---- warp / block reductions (fixed order => deterministic) ------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// NaN-propagating max of |x| (Julia's maximum(abs, x) returns NaN if any element is NaN)
__device__ __forceinline__ double nanmax(double a, double b) {
    return (a != a) ? a : ((b != b) ? b : (a > b ? a : b));
}
__device__ __forceinline__ double warp_nanmax(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = nanmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// Block-wide sum; result valid in thread 0 (and all threads of warp 0). blockDim.x multiple of 32, <= 1024.
__device__ __forceinline__ double block_sum(double v, double* sm /* >= 32 doubles */) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();  // protect sm reuse
    if (lane == 0) sm[w] = v;
    __syncthreads();
    v = (threadIdx.x < nw) ? sm[threadIdx.x] : 0.0;
    if (w == 0) v = warp_sum(v);
    return v;
}
__device__ __forceinline__ double block_nanmax(double v, double* sm) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_nanmax(v);
    __syncthreads();
    if (lane == 0) sm[w] = v;
    __syncthreads();
    v = (threadIdx.x < nw) ? sm[threadIdx.x] : 0.0;
    if (w == 0) v = warp_nanmax(v);
    return v;
}
#endif
