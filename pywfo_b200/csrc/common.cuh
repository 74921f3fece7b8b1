// Shared helpers of the wfo_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/wfo_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "wfo_b200 is written for sm_100a (B200) only"
#endif

namespace wfo {

extern thread_local char g_err[512];

// printf-style; integer arguments are passed as long long (use %lld)
template <typename... Args>
inline int set_err(int code, const char *fmt, Args... args) {
    if constexpr (sizeof...(Args) == 0) snprintf(g_err, sizeof(g_err), "%s", fmt);
    else snprintf(g_err, sizeof(g_err), fmt, static_cast<long long>(args)...);
    return code;
}

#define WFO_CUDA(call)                                                                   \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) {                                                         \
            snprintf(wfo::g_err, sizeof(wfo::g_err), "%s:%d %s: %s", __FILE__, __LINE__, \
                     #call, cudaGetErrorString(e_));                                     \
            return WFO_ERR_CUDA;                                                         \
        }                                                                                \
    } while (0)

#define WFO_TRY(call)            \
    do {                         \
        int r_ = (call);         \
        if (r_ != WFO_OK) return r_; \
    } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// FP64 tensor-core MMA, D(8x8) += A(8x4) * B(4x8).  SASS: DMMA.8x8x4.
// Fragment ownership (lane = 4*g + t):  a = A[g][t],  b = B[t][g],
// c0/c1 = C[g][2t], C[g][2t+1].
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// 1/x without the library's special-case slow path: MUFU.RCP64H seed (~20 bits) and two Newton
// steps (4 dependent DFMA).  x = 0 / subnormal gives inf/NaN, which callers treat as singular.
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}

__device__ __forceinline__ unsigned hi_abs(double x) {
    return (unsigned)__double2hiint(x) & 0x7fffffffu;
}

}  // namespace wfo

namespace wfo { struct BaseInfo; struct Comm; }

constexpr int TRAJ_MAX_LANES = 4;   // worker contexts of the trajectory driver

struct wfo_ctx {
    int device;
    int sm_count;
    size_t smem_optin;
    cudaStream_t stream;   // private stream of the *_host entry points
    cudaStream_t stream2;  // side stream: work that overlaps the single-CTA base factorisation
    // early base block (wfo_set_occ_hint): wfo_smo_inv_dev forms S_MO[:occ,:occ] first and inverts it on stream3
    // while the rest of S_MO is still being multiplied; the next state-overlap call on that S_MO picks it up
    cudaStream_t stream3, stream4;
    int *pre_flag;           // device: raised (to pre_epoch) when the base block is ready for the waiting inverse kernel
    int pre_epoch;
    int pre_probe;           // 0 not probed yet, 1 kernels on two streams run concurrently (early base block usable), -1 they do not
    cudaEvent_t ev_pre_in, ev_pre;
    int occ_hint;
    char *pre_buf;
    size_t pre_bytes;
    const double *pre_smo;   // the S_MO buffer the stashed inverse belongs to (valid until consumed)
    int pre_n;
    double *pre_Ai;
    const double *pre_A;     // the base block as the early chain multiplied it (K slabs), input of the early kernel AND its guard
    int pre_nslab;
    wfo::BaseInfo *pre_info;
    cudaEvent_t ev_fork, ev_join;
    double *ns_slots;      // device: norms + per-iteration residuals of the Newton-Schulz inverse (128 doubles)
    double *h_slots;       // pinned host mirror
    int last_inverse_iters;
    int last_inverse_pivoted;   // 1: the last inverse came from the pivoted Gauss-Jordan fallback
    char *gj_buf;          // its bookkeeping
    size_t gj_bytes;
    wfo::BaseInfo *h_info;  // pinned host copy of the base-block verdict
    cudaEvent_t ev_info;
    char *scratch;         // device arena, grown on demand
    size_t scratch_bytes;
    double *gscratch;      // per-CTA LU working matrices (global-scratch variant of the T0 kernel)
    unsigned long long *t0_next;   // device: work counter of the brute-force pair kernel (zeroed before each launch)
    size_t gscratch_bytes;
    char *resident;        // inputs + phase-1 workspace of the raw-CI / trajectory entry points
    size_t resident_bytes;
    long long launches;
    int last_tier;
    wfo::Comm *comm;       // in-library collective over peer memory (wfo_comm_init), NULL = single rank
    wfo_ctx *lanes[TRAJ_MAX_LANES];   // [0] = this context, others created on demand (wfo_trajectory_ci_host)
    int last_lanes;
    int lu_variant;        // WFO_LU_*: which brute-force kernel serves 32 < n <= 128
    int timing;
    float last_ms;
    cudaEvent_t ev0, ev1;
    // phase breakdown of the last state-overlap call (timing level 2): events at phase boundaries
    cudaEvent_t pev[24];
    const char *pname[24];
    int n_phase;
};

// record a phase boundary on `st` (timing level 2 only); the name labels the phase that ENDS here
#define WFO_PHASE(ctx, st, name)                                            \
    do {                                                                    \
        if ((ctx)->timing >= 2 && (ctx)->n_phase < 24) {                    \
            WFO_CUDA(cudaEventRecord((ctx)->pev[(ctx)->n_phase], st));      \
            (ctx)->pname[(ctx)->n_phase++] = name;                          \
        }                                                                   \
    } while (0)
