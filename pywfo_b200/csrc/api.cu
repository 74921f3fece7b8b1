// extern "C" entry points of libwfo_b200.so (see include/wfo_b200.h).
#include <math.h>
#include <stdlib.h>
#include <new>
#include <vector>

#include <algorithm>
#include <map>
#include <thread>
#include <vector>

#include "common.cuh"
#include "comm.cuh"
#include "det_lu_warp.cuh"
#include "det_lu_ll.cuh"
#include "det_lu_wpm.cuh"
#include "gemm_f64.cuh"
#include "inverse_gj.cuh"
#include "lists.cuh"
#include "lu_base.cuh"
#include "misc_kernels.cuh"
#include "pair_update.cuh"
#include "tables.cuh"

namespace wfo {
thread_local char g_err[512] = "";

struct Arena {
    char *base;
    size_t cap, off;
    template <typename T>
    T *take(size_t count) {
        size_t bytes = align_up(count * sizeof(T), 256);
        T *p = reinterpret_cast<T *>(base + off);
        off += bytes;
        return p;
    }
};
struct Sizer {   // same interface, only adds up
    size_t off = 0;
    template <typename T>
    T *take(size_t count) {
        off += align_up(count * sizeof(T), 256);
        return nullptr;
    }
};

static int ensure_scratch(wfo_ctx *ctx, size_t bytes) {
    if (bytes <= ctx->scratch_bytes) return WFO_OK;
    WFO_CUDA(cudaDeviceSynchronize());   // earlier async work may still read the old arena
    if (ctx->scratch) WFO_CUDA(cudaFree(ctx->scratch));
    ctx->scratch = nullptr;
    ctx->scratch_bytes = 0;
    size_t want = bytes + bytes / 4 + (1 << 20);
    WFO_CUDA(cudaMalloc(&ctx->scratch, want));
    ctx->scratch_bytes = want;
    return WFO_OK;
}

// grow the arena but keep its first keep_bytes (device-to-device copy)
static int ensure_scratch_keep(wfo_ctx *ctx, size_t bytes, size_t keep_bytes) {
    if (bytes <= ctx->scratch_bytes) return WFO_OK;
    WFO_CUDA(cudaDeviceSynchronize());
    size_t want = bytes + bytes / 4 + (1 << 20);
    char *fresh = nullptr;
    WFO_CUDA(cudaMalloc(&fresh, want));
    if (ctx->scratch && keep_bytes) WFO_CUDA(cudaMemcpy(fresh, ctx->scratch, keep_bytes, cudaMemcpyDeviceToDevice));
    if (ctx->scratch) WFO_CUDA(cudaFree(ctx->scratch));
    ctx->scratch = fresh;
    ctx->scratch_bytes = want;
    return WFO_OK;
}

#define WFO_LAUNCH_CHECK(ctx)                 \
    do {                                      \
        (ctx)->launches++;                    \
        WFO_CUDA(cudaGetLastError());         \
    } while (0)

static int gemm(wfo_ctx *ctx, bool ta, bool tb, int m, int n, int k, double alpha, const double *A, int lda,
                const double *B, int ldb, double beta, double *C, int ldc, int kslab, long long slab_stride,
                cudaStream_t st, int batch = 0, long long batch_a = 0, long long batch_b = 0) {
    if (m <= 0 || n <= 0) return WFO_OK;
    if (kslab <= 0 || batch > 0) kslab = k > 0 ? k : 1;
    int nz = batch > 0 ? batch : (k > 0 ? cdiv(k, kslab) : 1);
    // 64x64 CTA tiles unless that leaves SMs idle: then 32x64, then 32x32
    int gm = 64, gn = 64;
    if ((long long)cdiv(n, 64) * cdiv(m, 64) * nz < ctx->sm_count) gm = 32;
    if (gm == 32 && (long long)cdiv(n, 64) * cdiv(m, 32) * nz < ctx->sm_count) gn = 32;
    dim3 grid(cdiv(n, gn), cdiv(m, gm), nz), block(128);
    // 16-byte async copies need even leading dimensions / slab starts / batch strides and aligned bases
    const bool vec = !ta && !((lda | ldb) & 1) && !(((uintptr_t)A | (uintptr_t)B) & 15) &&
                     !((batch_a | batch_b) & 1) && (nz == 1 || batch > 0 || !(kslab & 1));
#define GEMM_LAUNCH2(TA_, TB_, GM_, GN_, VEC_)                                                                    \
    do {                                                                                                          \
        static bool attr_done_[64];   /* per device: function attributes belong to the device's context */      \
        if (!attr_done_[ctx->device & 63]) {                                                                      \
            WFO_CUDA(cudaFuncSetAttribute(gemm_f64_kernel<TA_, TB_, GM_, GN_, VEC_>,                              \
                                          cudaFuncAttributeMaxDynamicSharedMemorySize,                            \
                                          (int)gemm_smem_bytes<GM_, GN_>()));                                     \
            attr_done_[ctx->device & 63] = true;                                                                  \
        }                                                                                                         \
        gemm_f64_kernel<TA_, TB_, GM_, GN_, VEC_><<<grid, block, gemm_smem_bytes<GM_, GN_>(), st>>>(              \
            m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, kslab, slab_stride, batch_a, batch_b);                  \
    } while (0)
#define GEMM_LAUNCH(TA_, TB_, VEC_)                                       \
    do {                                                                  \
        if (gm == 64) { GEMM_LAUNCH2(TA_, TB_, 64, 64, VEC_); }           \
        else if (gn == 64) { GEMM_LAUNCH2(TA_, TB_, 32, 64, VEC_); }      \
        else { GEMM_LAUNCH2(TA_, TB_, 32, 32, VEC_); }                    \
    } while (0)
    if (vec && !tb) GEMM_LAUNCH(false, false, true);
    else if (vec && tb) GEMM_LAUNCH(false, true, true);
    else if (!ta && !tb) GEMM_LAUNCH(false, false, false);
    else if (!ta && tb) GEMM_LAUNCH(false, true, false);
    else if (ta && !tb) GEMM_LAUNCH(true, false, false);
    else GEMM_LAUNCH(true, true, false);
#undef GEMM_LAUNCH
#undef GEMM_LAUNCH2
    WFO_LAUNCH_CHECK(ctx);
    return WFO_OK;
}

// ---- base block of any size: register-resident Gauss-Jordan up to 128, above that a 2x2 block inverse
// through the Schur complement (recursive), everything else DMMA GEMMs:
//   inv(A) = [[A11i + T1 Si T2, -T1 Si], [-Si T2, Si]],  T1 = A11i A12, T2 = A21 A11i, S = A22 - A21 T1
constexpr int WFO_MAX_OCC = 4096;
static size_t base_work_doubles(int n) {
    if (n <= LUB_MAXN) return 0;
    const size_t n1 = n / 2, n2 = n - n1;
    const size_t a = base_work_doubles((int)n1), b = base_work_doubles((int)n2);
    return n1 * n1 + 2 * n1 * n2 + 2 * n2 * n2 + 64 + (a > b ? a : b);
}
static int copy2d(wfo_ctx *ctx, const double *src, int ld_src, double *dst, int ld_dst, int rows, int cols, cudaStream_t st) {
    const long long tot = (long long)rows * cols;
    if (tot <= 0) return WFO_OK;
    copy2d_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(src, ld_src, dst, ld_dst, rows, cols);
    WFO_LAUNCH_CHECK(ctx);
    return WFO_OK;
}
static int base_factor(wfo_ctx *ctx, const double *A, int lda, int n, double *Ai, int ldo, BaseInfo *info, double *work,
                       cudaStream_t st, int n_slabs = 1, long long slab_stride = 0) {
    if (n <= LUB_MAXN) {
        base_inverse_kernel<<<1, LUB_THREADS, lub_smem_bytes(n), st>>>(A, lda, n, Ai, ldo, info, n_slabs, slab_stride, nullptr, 0);
        WFO_LAUNCH_CHECK(ctx);
        return WFO_OK;
    }
    const int n1 = n / 2, n2 = n - n1;
    double *A11i = work, *T1 = A11i + (size_t)n1 * n1, *T2 = T1 + (size_t)n1 * n2, *S = T2 + (size_t)n2 * n1,
           *Si = S + (size_t)n2 * n2;
    BaseInfo *infos = reinterpret_cast<BaseInfo *>(Si + (size_t)n2 * n2);
    double *rest = Si + (size_t)n2 * n2 + 64;
    const double *A12 = A + n1, *A21 = A + (size_t)n1 * lda, *A22 = A21 + n1;
    WFO_TRY(base_factor(ctx, A, lda, n1, A11i, n1, &infos[0], rest, st));
    WFO_TRY(gemm(ctx, false, false, n1, n2, n1, 1.0, A11i, n1, A12, lda, 0.0, T1, n2, 0, 0, st));
    WFO_TRY(gemm(ctx, false, false, n2, n1, n1, 1.0, A21, lda, A11i, n1, 0.0, T2, n1, 0, 0, st));
    WFO_TRY(copy2d(ctx, A22, lda, S, n2, n2, n2, st));
    WFO_TRY(gemm(ctx, false, false, n2, n2, n1, -1.0, A21, lda, T1, n2, 1.0, S, n2, 0, 0, st));
    WFO_TRY(base_factor(ctx, S, n2, n2, Si, n2, &infos[1], rest, st));
    double *B12 = Ai + n1, *B21 = Ai + (size_t)n1 * ldo, *B22 = B21 + n1;
    WFO_TRY(copy2d(ctx, Si, n2, B22, ldo, n2, n2, st));
    WFO_TRY(gemm(ctx, false, false, n1, n2, n2, -1.0, T1, n2, Si, n2, 0.0, B12, ldo, 0, 0, st));
    WFO_TRY(gemm(ctx, false, false, n2, n1, n2, -1.0, Si, n2, T2, n1, 0.0, B21, ldo, 0, 0, st));
    WFO_TRY(copy2d(ctx, A11i, n1, Ai, ldo, n1, n1, st));
    WFO_TRY(gemm(ctx, false, false, n1, n1, n2, -1.0, B12, ldo, T2, n1, 1.0, Ai, ldo, 0, 0, st));
    combine_info_kernel<<<1, 32, 0, st>>>(&infos[0], &infos[1], info);
    WFO_LAUNCH_CHECK(ctx);
    return WFO_OK;
}

static int ensure_gscratch(wfo_ctx *ctx, size_t bytes) {
    if (bytes <= ctx->gscratch_bytes) return WFO_OK;
    WFO_CUDA(cudaDeviceSynchronize());
    if (ctx->gscratch) WFO_CUDA(cudaFree(ctx->gscratch));
    ctx->gscratch = nullptr;
    ctx->gscratch_bytes = 0;
    WFO_CUDA(cudaMalloc(&ctx->gscratch, bytes));
    ctx->gscratch_bytes = bytes;
    return WFO_OK;
}

// Brute-force tier, 32 < n <= 128: two warp-per-matrix kernels.
//   left-looking (det_lu_ll.cuh), the default: M panels in shared memory up to n = 64 (21 % of the FP64 peak at
//     n = 64 against 15 %), in an L2-resident scratch above (eight matrices per SM: 16-17 % at n = 120-128
//     against 14-16 %, level at n = 100);
//   right-looking with an L2-resident scratch (det_lu_wpm.cuh): round 1's kernel, kept as the cross-check.
// wfo_set_lu_variant() pins one of them (tests run both at every size).
static bool use_ll_kernels(const wfo_ctx *ctx, int n) {
    if (n <= 32 || n > 128) return false;
    if (ctx->lu_variant == WFO_LU_LEFT) return true;
    if (ctx->lu_variant == WFO_LU_RIGHT) return false;
    return true;
}
#define WFO_LL_DISPATCH(n, CALL)                                                         \
    do {                                                                                 \
        switch (((n) + 7) / 8) {                                                         \
            case 5: CALL(5, false); break;   case 6: CALL(6, false); break;              \
            case 7: CALL(7, false); break;   case 8: CALL(8, false); break;              \
            case 9: CALL(9, true); break;    case 10: CALL(10, true); break;             \
            case 11: CALL(11, true); break;  case 12: CALL(12, true); break;             \
            case 13: CALL(13, true); break;  case 14: CALL(14, true); break;             \
            case 15: CALL(15, true); break;  default: CALL(16, true); break;             \
        }                                                                                \
    } while (0)
#define WFO_WPM_DISPATCH(n, CALL)            \
    do {                                     \
        if ((n) <= 64) { CALL(2); }          \
        else if ((n) <= 96) { CALL(3); }     \
        else { CALL(4); }                    \
    } while (0)
// register-resident segment kernels for n <= 32: (padded columns, lanes per determinant)
#define WFO_WARP_DISPATCH(n, CALL)             \
    do {                                       \
        if ((n) <= 8) { CALL(8, 8); }          \
        else if ((n) <= 16) { CALL(16, 16); }  \
        else if ((n) <= 24) { CALL(24, 32); }  \
        else { CALL(32, 32); }                 \
    } while (0)

template <typename K>
static int persistent_grid(wfo_ctx *ctx, K kernel, int threads, size_t smem, long long items, unsigned *grid) {
    WFO_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    WFO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
    if (per_sm < 1) per_sm = 1;
    long long g = (long long)ctx->sm_count * per_sm;
    if (g > items) g = items;
    *grid = (unsigned)g;
    return WFO_OK;
}

static int det_table(wfo_ctx *ctx, const double *S, int ld_s, int n, const int32_t *rows, int kb,
                     const int32_t *cols, int kk, double *D, cudaStream_t st) {
    long long npairs = (long long)kb * kk;
    if (npairs <= 0) return WFO_OK;
    if (n < 1 || n > 128)
        return set_err(WFO_ERR_ARG, "det_table: n_occ=%lld outside the supported range 1..128", n);
    unsigned grid = 1;
    if (use_ll_kernels(ctx, n)) {
#define CALL_DL(NT_, MG_)                                                                                       \
    WFO_TRY(persistent_grid(ctx, det_table_ll_kernel<NT_, MG_>, LLWarps<NT_, MG_>::W * 32,                           \
                            LLWarps<NT_, MG_>::W * LLCfg<NT_, MG_>::BYTES,                                           \
                            (npairs + LLWarps<NT_, MG_>::W - 1) / LLWarps<NT_, MG_>::W, &grid));                     \
    if (MG_) WFO_TRY(ensure_gscratch(ctx, (size_t)grid * LLWarps<NT_, MG_>::W * ll_scratch_doubles(NT_) * 8));       \
    det_table_ll_kernel<NT_, MG_><<<grid, LLWarps<NT_, MG_>::W * 32, LLWarps<NT_, MG_>::W * LLCfg<NT_, MG_>::BYTES, st>>>( \
        S, ld_s, n, rows, kb, cols, kk, D, ctx->gscratch)
        WFO_LL_DISPATCH(n, CALL_DL);
#undef CALL_DL
    } else if (n > 32) {
#define CALL_DP(SL_)                                                                                          \
    WFO_TRY(persistent_grid(ctx, det_table_wpm_kernel<SL_>, WPM_WARPS * 32, 0, (npairs + WPM_WARPS - 1) / WPM_WARPS, &grid)); \
    WFO_TRY(ensure_gscratch(ctx, (size_t)grid * WPM_WARPS * wpm_mat_bytes(n)));                                \
    det_table_wpm_kernel<SL_><<<grid, WPM_WARPS * 32, 0, st>>>(S, ld_s, n, rows, kb, cols, kk, D, ctx->gscratch)
        WFO_WPM_DISPATCH(n, CALL_DP);
#undef CALL_DP
    } else {
#define CALL_DW(NC_, SEG_)                                                                                 \
    WFO_TRY(persistent_grid(ctx, det_table_warp_kernel<NC_, SEG_>, 128, 0, (npairs + 3) / 4, &grid));      \
    det_table_warp_kernel<NC_, SEG_><<<grid, 128, 0, st>>>(S, ld_s, n, rows, kb, cols, kk, D)
        WFO_WARP_DISPATCH(n, CALL_DW);
#undef CALL_DW
    }
    WFO_LAUNCH_CHECK(ctx);
    return WFO_OK;
}

// Variants of the rank-update kernel: (row stride in tiles, DMMA tiles, DFMA columns) for n_ket = 1..55.
// DFMA columns (REM > 0) are compiled out of the dispatch: measured on B200, a DFMA issued between
// DMMAs costs ~11 pipe cycles per SMSP instead of 2, so 2 leftover columns as DFMA are slower
// (+0.94 ms at C4) than padding them to a seventh DMMA tile (+0.69 ms).
#define WFO_T1_VARIANTS(X)                                            \
    X(1, 1, 0)                                                        \
    X(3, 1, 0) X(3, 2, 0) X(3, 3, 0)                                  \
    X(5, 3, 0) X(5, 4, 0) X(5, 5, 0)                                  \
    X(7, 5, 0) X(7, 6, 0) X(7, 7, 0)

static int pick_ldg(int n_ket) {
    int nt = cdiv(n_ket + 1, 8);
    if ((nt & 1) == 0) nt++;   // odd tile count: conflict-free B-fragment loads (stride = 8 mod 16)
    return nt * 8;
}

// how the n_ket state columns are split between the tensor cores and DFMA (see pair_update.cuh)
static void t1_split(int n_ket, int *nt, int *nd, int *rem) {
    *nt = pick_ldg(n_ket) / 8;
    *nd = (n_ket + 7) / 8;   // the D(k,0) column behind the states is filled in by the epilogue, not here
    *rem = 0;
}

// the rank-update kernel runs one persistent CTA per SM; opt in to its dynamic shared memory once
static int t1_prepare(const wfo_ctx *ctx, int n_ket) {
    static bool done_all[64][1000];   // per device
    bool *done = done_all[ctx->device & 63];
    int nt, nd, rem;
    t1_split(n_ket, &nt, &nd, &rem);
    const int key = nt * 100 + nd * 10 + rem;
    if (nt < 1 || nt > 7) return set_err(WFO_ERR_ARG, "state_overlaps: unsupported number of ket states %lld per pass", n_ket);
    if (!done[key]) {
        switch (key) {
#define T1_ATTR(NT_, ND_, REM_) case NT_ * 100 + ND_ * 10 + REM_: WFO_CUDA(cudaFuncSetAttribute(t1_fused_kernel<NT_, ND_, REM_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)t1_smem_bytes<NT_>())); break;
            WFO_T1_VARIANTS(T1_ATTR)
#undef T1_ATTR
            default: return set_err(WFO_ERR_ARG, "state_overlaps: no rank-update kernel for %lld ket states", n_ket);
        }
        done[key] = true;
    }
    return WFO_OK;
}

struct Plan {   // everything wfo_state_overlaps needs from the arena
    // common
    double *ckt, *gpart, *P, *wvec, *dk0, *d0l, *d00, *out_dev;
    // T0
    int32_t *bra_lists, *ket_lists, *gs_list;
    // T1
    double *Abuf, *Ai, *X, *Y, *Wb, *bscale, *bwork, *gjwork;
    BraMeta *bmeta;
    KetMeta *kmeta;
    BaseInfo *info;
};

template <typename AR>
static void carve(AR &ar, Plan &p, int n, int v, int ksh, int kk, int n_bra, int ldg, int n_parts, int n_slabs,
                  bool t0, bool t1) {
    const size_t kk_pad = align_up((size_t)(kk > 0 ? kk : 1), T1_LT);   // whole ket tiles for the TMA ring
    p.ckt = ar.template take<double>(kk_pad * ldg);
    const size_t ksh_g = (size_t)ksh + 1;   // + the virtual ground-state row of the rank-update tier
    p.gpart = ar.template take<double>((size_t)n_parts * ksh_g * ldg);
    p.wvec = ar.template take<double>((size_t)KGS_SLABS * ldg);
    p.P = ar.template take<double>((size_t)n_slabs * n_bra * ldg);
    p.dk0 = ar.template take<double>(ksh_g);
    p.d0l = ar.template take<double>(kk > 0 ? kk : 1);
    p.d00 = ar.template take<double>(1);
    if (t0) {
        p.bra_lists = ar.template take<int32_t>((size_t)(ksh > 0 ? ksh : 1) * n);
        p.ket_lists = ar.template take<int32_t>((size_t)(kk > 0 ? kk : 1) * n);
        p.gs_list = ar.template take<int32_t>(n);
    }
    if (t1) {
        p.Ai = ar.template take<double>((size_t)n * n);
        p.X = ar.template take<double>((size_t)(v > 0 ? v : 1) * n);
        p.Y = ar.template take<double>((size_t)n * (v > 0 ? v : 1));
        p.Wb = ar.template take<double>((size_t)(v + 1) * (v + 1));
        p.bscale = ar.template take<double>(ksh_g);
        p.bmeta = ar.template take<BraMeta>(ksh_g);
        p.kmeta = ar.template take<KetMeta>(kk_pad);
        p.info = ar.template take<BaseInfo>(1);
        p.bwork = ar.template take<double>(base_work_doubles(n) + 1);
        // base blocks above 128 (no brute-force tier behind them): work matrices of the pivoted fallback inverse
        p.gjwork = (n > LUB_MAXN) ? ar.template take<double>(2 * (size_t)n * n) : nullptr;
    }
}



struct Part {   // how the pair space of one call is cut up
    int ldg, t0_chunks, chunk_len, n_parts, n_slabs;
    PartMap t1;   // stream-K decomposition of the rank-update kernel
    bool may_t0, want_t1;
};

static size_t base_smem_bytes(int n) { return n <= WFO_MAX_OCC ? 0 : (size_t)1 << 40; }   // block inverse above 128

// which tiers can / should run for this call
static int resolve_tier(const wfo_ctx *ctx, int n, int tier, bool *want_t1, bool *may_t0) {
    if (tier != WFO_TIER_AUTO && tier != WFO_TIER_LU && tier != WFO_TIER_UPDATE && tier != WFO_TIER_CLOSED)
        return set_err(WFO_ERR_ARG, "state_overlaps: unknown tier %lld", tier);
    if (tier == WFO_TIER_CLOSED) {   // shares the base-block machinery (and its size limit) with UPDATE
        *want_t1 = true;
        *may_t0 = false;
        if (base_smem_bytes(n) > ctx->smem_optin)
            return set_err(WFO_ERR_ARG, "state_overlaps: tier CLOSED supports n_occ <= 4096, got %lld", n);
        return WFO_OK;
    }
    *want_t1 = (tier == WFO_TIER_AUTO || tier == WFO_TIER_UPDATE);
    *may_t0 = (tier == WFO_TIER_AUTO || tier == WFO_TIER_LU);
    const bool t0_fits = n <= 128;
    const bool t1_fits = base_smem_bytes(n) <= ctx->smem_optin;
    if (tier == WFO_TIER_LU && !t0_fits)
        return set_err(WFO_ERR_ARG, "state_overlaps: tier LU supports n_occ <= 128, got %lld", n);
    if (tier == WFO_TIER_UPDATE && !t1_fits)
        return set_err(WFO_ERR_ARG, "state_overlaps: tier UPDATE supports n_occ <= 4096, got %lld", n);
    if (tier == WFO_TIER_AUTO) {
        if (!t1_fits) *want_t1 = false;
        if (!t0_fits) *may_t0 = false;
        if (!*want_t1 && !*may_t0)
            return set_err(WFO_ERR_ARG, "state_overlaps: n_occ=%lld too large for every tier", n);
    }
    return WFO_OK;
}

static int make_part(wfo_ctx *ctx, int ksh, int kk, int n_ket, bool may_t0, bool want_t1, Part *out) {
    Part q;
    q.may_t0 = may_t0;
    q.want_t1 = want_t1;
    q.ldg = pick_ldg(n_ket);
    q.n_parts = 1;
    memset(&q.t1, 0, sizeof(q.t1));
    if (want_t1) {
        // T1: iterations (bra tile, ket tile) in equal contiguous ranges over the resident CTAs
        WFO_TRY(t1_prepare(ctx, n_ket));
        const int slots = ctx->sm_count;
        const int ktiles = cdiv(ksh + 1, T1_KTILE), n_ltiles = cdiv(kk, T1_LT);
        const long long I = (long long)ktiles * n_ltiles;
        long long G = slots;
        if (G > I / 4) G = I / 4;   // at least 4 ket tiles per CTA
        if (G < 1) G = 1;
        q.t1.I = I;
        q.t1.G = (int)G;
        q.t1.n_ltiles = n_ltiles;
        q.t1.ktile_rows = T1_KTILE;
        q.t1.uniform_parts = 0;
        const long long per = I / G;   // >= 1
        const int parts = (int)((n_ltiles + per - 1) / per) + 1;
        if (parts > q.n_parts) q.n_parts = parts;
    }
    // T0: items (k, chunk of l)
    // (about 32 items per resident warp: the items are handed out dynamically and a coarse last item is idle time;
    // C5: 96 items per SM 20.9 ms, 192: 20.3, 384: 20.0, 768: 19.9)
    const int target_ctas = ctx->sm_count * 384;
    int chunks = cdiv(target_ctas, ksh);
    if (chunks > kk) chunks = kk;
    if (chunks < 1) chunks = 1;
    q.chunk_len = cdiv(kk, chunks);
    q.t0_chunks = cdiv(kk, q.chunk_len);
    if (may_t0 && q.t0_chunks > q.n_parts) q.n_parts = q.t0_chunks;
    q.n_slabs = cdiv(ksh + 1, RC_ROWS);
    *out = q;
    return WFO_OK;
}

__global__ void iota_kernel(int32_t *p, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}

}  // namespace wfo

using namespace wfo;

extern "C" {

int wfo_version(void) { return 100; }
const char *wfo_last_error(void) { return g_err; }

int wfo_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int wfo_create(int device, wfo_ctx **out) {
    if (!out) return set_err(WFO_ERR_ARG, "wfo_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    WFO_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return set_err(WFO_ERR_ARG, "wfo_create: no CUDA device %lld", device);
    WFO_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    WFO_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return set_err(WFO_ERR_ARG, "wfo_create: device is sm_%lld%lld, this library is sm_100a only", prop.major, prop.minor);
    wfo_ctx *c = new (std::nothrow) wfo_ctx();
    if (!c) return set_err(WFO_ERR_ARG, "wfo_create: out of host memory");
    memset(c, 0, sizeof(*c));
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->smem_optin = prop.sharedMemPerBlockOptin;
    c->last_tier = -1;
    c->lu_variant = WFO_LU_AUTO;
    WFO_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    WFO_CUDA(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
    {   // the early base block is a short dependent chain next to wide GEMMs: its CTAs go first
        int prio_lo = 0, prio_hi = 0;
        WFO_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        WFO_CUDA(cudaStreamCreateWithPriority(&c->stream3, cudaStreamNonBlocking, prio_hi));
        WFO_CUDA(cudaStreamCreateWithPriority(&c->stream4, cudaStreamNonBlocking, prio_hi));
        WFO_CUDA(cudaMalloc(&c->pre_flag, sizeof(int)));
        WFO_CUDA(cudaMemset(c->pre_flag, 0, sizeof(int)));
    }
    WFO_CUDA(cudaEventCreateWithFlags(&c->ev_pre_in, cudaEventDisableTiming));
    WFO_CUDA(cudaEventCreateWithFlags(&c->ev_pre, cudaEventDisableTiming));
    WFO_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    WFO_CUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    WFO_CUDA(cudaEventCreate(&c->ev0));
    WFO_CUDA(cudaEventCreate(&c->ev1));
    for (int i = 0; i < 24; i++) WFO_CUDA(cudaEventCreate(&c->pev[i]));
    WFO_CUDA(cudaFuncSetAttribute(base_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(LUB_EXCL_SMEM > lub_smem_bytes(LUB_MAXN) ? LUB_EXCL_SMEM : lub_smem_bytes(LUB_MAXN))));
    WFO_CUDA(cudaMalloc(&c->ns_slots, 128 * sizeof(double)));
    WFO_CUDA(cudaMalloc(&c->t0_next, sizeof(unsigned long long)));
    WFO_CUDA(cudaMallocHost(&c->h_slots, 128 * sizeof(double)));
    WFO_CUDA(cudaMallocHost(&c->h_info, sizeof(BaseInfo)));
    WFO_CUDA(cudaEventCreateWithFlags(&c->ev_info, cudaEventDisableTiming));
    *out = c;
    return WFO_OK;
}

int wfo_destroy(wfo_ctx *ctx) {
    if (!ctx) return WFO_OK;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    wfo_comm_destroy(ctx);
    for (int l = 1; l < TRAJ_MAX_LANES; l++)
        if (ctx->lanes[l]) { wfo_destroy(ctx->lanes[l]); ctx->lanes[l] = nullptr; }
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->gscratch) cudaFree(ctx->gscratch);
    if (ctx->gj_buf) cudaFree(ctx->gj_buf);
    if (ctx->resident) cudaFree(ctx->resident);
    if (ctx->ns_slots) cudaFree(ctx->ns_slots);
    if (ctx->t0_next) cudaFree(ctx->t0_next);
    if (ctx->h_slots) cudaFreeHost(ctx->h_slots);
    if (ctx->h_info) cudaFreeHost(ctx->h_info);
    cudaEventDestroy(ctx->ev_info);
    cudaEventDestroy(ctx->ev0);
    cudaEventDestroy(ctx->ev1);
    for (int i = 0; i < 24; i++) cudaEventDestroy(ctx->pev[i]);
    cudaEventDestroy(ctx->ev_fork);
    cudaEventDestroy(ctx->ev_join);
    if (ctx->pre_buf) cudaFree(ctx->pre_buf);
    cudaEventDestroy(ctx->ev_pre_in);
    cudaEventDestroy(ctx->ev_pre);
    cudaStreamDestroy(ctx->stream3);
    cudaStreamDestroy(ctx->stream4);
    if (ctx->pre_flag) cudaFree(ctx->pre_flag);
    cudaStreamDestroy(ctx->stream2);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return WFO_OK;
}

long long wfo_launch_count(const wfo_ctx *ctx) { return ctx ? ctx->launches : 0; }
/* ------------------------------------------------------------ K5: in-library collective */
int wfo_comm_destroy(wfo_ctx *ctx) {
    if (!ctx || !ctx->comm) return WFO_OK;
    Comm *c = ctx->comm;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < c->world; r++)
        if (r != c->rank && c->peers.base[r]) cudaIpcCloseMemHandle(c->peers.base[r]);
    if (c->buf) cudaFree(c->buf);
    unlink(comm_path(c->tag, c->rank).c_str());
    delete c;
    ctx->comm = nullptr;
    return WFO_OK;
}

int wfo_comm_init(wfo_ctx *ctx, int rank, int world, const char *tag, long long gather_bytes) {
    if (!ctx || !tag) return set_err(WFO_ERR_ARG, "comm_init: NULL argument");
    if (world < 1 || world > COMM_MAX_RANKS || rank < 0 || rank >= world)
        return set_err(WFO_ERR_ARG, "comm_init: rank %lld of %lld (at most 16 ranks of one node)", rank, world);
    if (strlen(tag) >= sizeof(Comm::tag)) return set_err(WFO_ERR_ARG, "comm_init: tag too long");
    WFO_TRY(wfo_comm_destroy(ctx));
    if (world == 1) return WFO_OK;
    WFO_CUDA(cudaSetDevice(ctx->device));
    Comm *c = new (std::nothrow) Comm();
    if (!c) return set_err(WFO_ERR_ARG, "comm_init: out of host memory");
    memset(c, 0, sizeof(*c));
    c->rank = rank;
    c->world = world;
    strcpy(c->tag, tag);
    c->gat_cap = align_up((size_t)(gather_bytes > 0 ? gather_bytes : (64ll << 20)), 4096);
    c->bytes = comm_gat_off(world) + c->gat_cap;
    ctx->comm = c;   // from here on wfo_comm_destroy cleans up
    WFO_CUDA(cudaMalloc(&c->buf, c->bytes));
    WFO_CUDA(cudaMemset(c->buf, 0, COMM_HDR_BYTES));
    WFO_CUDA(cudaDeviceSynchronize());
    CommCard mine;
    memset(&mine, 0, sizeof(mine));
    WFO_CUDA(cudaIpcGetMemHandle(&mine.handle, c->buf));
    mine.device = ctx->device;
    mine.pid = (int)getpid();
    mine.bytes = c->bytes;
    if (!comm_write_card(comm_path(tag, rank), mine)) {
        wfo_comm_destroy(ctx);
        return set_err(WFO_ERR_COMM, "comm_init: cannot write the rendezvous file in /dev/shm");
    }
    c->peers.base[rank] = c->buf;
    for (int r = 0; r < world; r++) {
        if (r == rank) continue;
        CommCard card;
        if (!comm_read_card(comm_path(tag, r), &card, 120.0)) {
            wfo_comm_destroy(ctx);
            return set_err(WFO_ERR_COMM, "comm_init: rank %lld did not show up within 120 s", r);
        }
        if (card.bytes != c->bytes) {
            wfo_comm_destroy(ctx);
            return set_err(WFO_ERR_COMM, "comm_init: rank %lld uses a different buffer size", r);
        }
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, card.handle, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            wfo_comm_destroy(ctx);
            return set_err(WFO_ERR_COMM, "comm_init: cannot map the buffer of rank %lld (no peer access between the GPUs?)", r);
        }
        c->peers.base[r] = static_cast<char *>(p);
    }
    return WFO_OK;
}

int wfo_comm_world(const wfo_ctx *ctx) { return (ctx && ctx->comm) ? ctx->comm->world : 0; }

// raise the epoch of both collectives (every rank, every call, also on an error path: the counters of
// all ranks stay in step)
static void comm_next(Comm *c, int *ge, int *re) {
    *ge = ++c->gat_epoch;
    *re = ++c->red_epoch;
}

static int comm_allreduce(wfo_ctx *ctx, double *buf, size_t count, int epoch, cudaStream_t st) {
    Comm *c = ctx->comm;
    if (count * sizeof(double) > COMM_RED_CAP)
        return set_err(WFO_ERR_ARG, "allreduce: %lld doubles exceed the slot size", (long long)count);
    const int parity = epoch & 1;
    const int bpp = count > 4096 ? 4 : 1;
    comm_push_kernel<<<c->world * bpp, 256, 0, st>>>(buf, count, c->peers, comm_red_off(c->world, parity, c->rank),
                                                      (int)(offsetof(CommHdr, red_flag) / 4), c->rank, c->world, epoch, bpp,
                                                      true, reinterpret_cast<CommHdr *>(c->buf));
    WFO_LAUNCH_CHECK(ctx);
    const int gs = (int)std::min<size_t>(8, (count + 255) / 256);
    comm_reduce_sum_kernel<<<gs, 256, 0, st>>>(c->buf, c->world, parity, epoch, count, buf);
    WFO_LAUNCH_CHECK(ctx);
    return WFO_OK;
}

// my flat slice [lo, hi) (doubles) of the gather buffer is in place: push it to the peers, wait for theirs
static int comm_allgather(wfo_ctx *ctx, size_t lo, size_t hi, int epoch, cudaStream_t st) {
    Comm *c = ctx->comm;
    double *g = reinterpret_cast<double *>(c->buf + comm_gat_off(c->world));
    const int bpp = 16;
    if (hi > lo) {
        comm_push_kernel<<<c->world * bpp, 256, 0, st>>>(g + lo, hi - lo, c->peers, comm_gat_off(c->world) + lo * 8,
                                                          (int)(offsetof(CommHdr, gat_flag) / 4), c->rank, c->world, epoch,
                                                          bpp, false, reinterpret_cast<CommHdr *>(c->buf));
    } else {   // nothing to send: raise the flags only
        comm_push_kernel<<<c->world, 32, 0, st>>>(g, 0, c->peers, comm_gat_off(c->world), (int)(offsetof(CommHdr, gat_flag) / 4),
                                                   c->rank, c->world, epoch, 1, false, reinterpret_cast<CommHdr *>(c->buf));
    }
    WFO_LAUNCH_CHECK(ctx);
    comm_wait_kernel<<<1, 32, 0, st>>>(reinterpret_cast<CommHdr *>(c->buf), (int)(offsetof(CommHdr, gat_flag) / 4), c->world,
                                       epoch, c->rank);
    WFO_LAUNCH_CHECK(ctx);
    return WFO_OK;
}

// after a stream sync: did a wait time out?
static int comm_check(wfo_ctx *ctx) {
    int err = 0;
    WFO_CUDA(cudaMemcpy(&err, ctx->comm->buf + offsetof(CommHdr, error), sizeof(int), cudaMemcpyDeviceToHost));
    if (err) {
        WFO_CUDA(cudaMemset(ctx->comm->buf + offsetof(CommHdr, error), 0, sizeof(int)));
        return set_err(WFO_ERR_COMM, "collective timed out: a rank did not take part (failed or called with other inputs)");
    }
    return WFO_OK;
}

int wfo_allreduce_dev(wfo_ctx *ctx, double *buf, long long count, void *stream) {
    if (!ctx || !buf || count < 0) return set_err(WFO_ERR_ARG, "allreduce_dev: bad argument");
    if (!ctx->comm) return WFO_OK;   // one rank
    WFO_CUDA(cudaSetDevice(ctx->device));
    int ge, re;
    comm_next(ctx->comm, &ge, &re);
    return comm_allreduce(ctx, buf, (size_t)count, re, (cudaStream_t)stream);
}

int wfo_comm_status(wfo_ctx *ctx) {
    if (!ctx || !ctx->comm) return WFO_OK;
    WFO_CUDA(cudaSetDevice(ctx->device));
    return comm_check(ctx);
}

int wfo_last_inverse_info(const wfo_ctx *ctx, int *iterations, int *pivoted) {
    if (!ctx) return set_err(WFO_ERR_ARG, "ctx is NULL");
    if (iterations) *iterations = ctx->last_inverse_iters;
    if (pivoted) *pivoted = ctx->last_inverse_pivoted;
    return WFO_OK;
}

/* ------------------------------------------------------------ general determinant lists */
namespace {

// Block-determinant table D (nb x nk) of one spin on the device, with super-block reuse on the bra side:
// bra blocks that share their first n-1 orbitals (pywfo/dets.py:119-156, `super_map`) are served by one
// pivoted factorisation per (super-block, ket block) plus one dot product per member
// (cofactor_rows_kernel); super-blocks with fewer than SUPER_MIN members and everything the cofactor path
// flags as ill conditioned go through the brute-force LU kernels.  stats: [super-blocks used, blocks
// served by them, blocks redone by brute force].
constexpr int SUPER_MIN = 4;

int det_table_super(wfo_ctx *ctx, const double *d_S, int ld_s, int n, const int32_t *h_bblk, int nb,
                    const int32_t *d_bblk, const int32_t *d_kblk, int nk, double *d_D, Arena &ar, cudaStream_t st,
                    int *stats) {
    stats[0] = stats[1] = stats[2] = 0;
    if (nb == 0 || nk == 0) return WFO_OK;
    std::vector<int> solo;   // blocks that take the brute-force path
    std::vector<int32_t> rows_rep, mem_off(1, 0), mem_block, mem_row;
    if (n >= 2 && n <= LUB_MAXN) {
        std::map<std::vector<int32_t>, std::vector<int>> groups;   // prefix (n-1 orbitals) -> blocks
        for (int b = 0; b < nb; b++)
            groups[std::vector<int32_t>(h_bblk + (size_t)b * n, h_bblk + (size_t)b * n + n - 1)].push_back(b);
        for (auto &g : groups) {
            if ((int)g.second.size() < SUPER_MIN) {
                solo.insert(solo.end(), g.second.begin(), g.second.end());
                continue;
            }
            const int rep = g.second[0];
            rows_rep.insert(rows_rep.end(), h_bblk + (size_t)rep * n, h_bblk + (size_t)rep * n + n);
            for (int b : g.second) {
                mem_block.push_back(b);
                mem_row.push_back(h_bblk[(size_t)b * n + n - 1]);
            }
            mem_off.push_back((int32_t)mem_block.size());
        }
    } else {
        for (int b = 0; b < nb; b++) solo.push_back(b);
    }
    const int n_super = (int)mem_off.size() - 1;
    stats[0] = n_super;
    stats[1] = (int)mem_block.size();

    auto brute = [&](const std::vector<int> &blocks) -> int {   // rows `blocks` of D by one LU per pair
        if (blocks.empty()) return WFO_OK;
        const int m = (int)blocks.size();
        if (m == nb) return det_table(ctx, d_S, ld_s, n, d_bblk, nb, d_kblk, nk, d_D, st);
        std::vector<int32_t> lists((size_t)m * n), row_of(m);
        for (int i = 0; i < m; i++) {
            std::copy(h_bblk + (size_t)blocks[i] * n, h_bblk + (size_t)blocks[i] * n + n, lists.begin() + (size_t)i * n);
            row_of[i] = blocks[i];
        }
        int32_t *d_lists = ar.take<int32_t>((size_t)m * n), *d_row_of = ar.take<int32_t>(m);
        double *d_tmp = ar.take<double>((size_t)m * nk);
        WFO_CUDA(cudaMemcpyAsync(d_lists, lists.data(), lists.size() * 4, cudaMemcpyHostToDevice, st));
        WFO_CUDA(cudaMemcpyAsync(d_row_of, row_of.data(), row_of.size() * 4, cudaMemcpyHostToDevice, st));
        WFO_CUDA(cudaStreamSynchronize(st));   // the staging vectors go out of scope
        WFO_TRY(det_table(ctx, d_S, ld_s, n, d_lists, m, d_kblk, nk, d_tmp, st));
        const long long tot = (long long)m * nk;
        scatter_rows_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(d_tmp, m, nk, d_row_of, d_D, nk);
        WFO_LAUNCH_CHECK(ctx);
        return WFO_OK;
    };
    WFO_TRY(brute(solo));
    if (n_super > 0) {
        int32_t *d_rep = ar.take<int32_t>(rows_rep.size()), *d_off = ar.take<int32_t>(mem_off.size()),
                *d_mb = ar.take<int32_t>(mem_block.size()), *d_mr = ar.take<int32_t>(mem_row.size());
        unsigned char *d_flags = ar.take<unsigned char>((size_t)n_super * nk);
        WFO_CUDA(cudaMemcpyAsync(d_rep, rows_rep.data(), rows_rep.size() * 4, cudaMemcpyHostToDevice, st));
        WFO_CUDA(cudaMemcpyAsync(d_off, mem_off.data(), mem_off.size() * 4, cudaMemcpyHostToDevice, st));
        WFO_CUDA(cudaMemcpyAsync(d_mb, mem_block.data(), mem_block.size() * 4, cudaMemcpyHostToDevice, st));
        WFO_CUDA(cudaMemcpyAsync(d_mr, mem_row.data(), mem_row.size() * 4, cudaMemcpyHostToDevice, st));
        static bool attr_done[64];
        if (!attr_done[ctx->device & 63]) {
            WFO_CUDA(cudaFuncSetAttribute(cofactor_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)lub_smem_bytes(LUB_MAXN)));
            attr_done[ctx->device & 63] = true;
        }
        const long long pairs = (long long)n_super * nk;
        int per_sm = 1;
        WFO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cofactor_rows_kernel, LUB_THREADS, lub_smem_bytes(n)));
        const unsigned grid = (unsigned)std::min<long long>(pairs, (long long)ctx->sm_count * (per_sm < 1 ? 1 : per_sm));
        cofactor_rows_kernel<<<grid, LUB_THREADS, lub_smem_bytes(n), st>>>(d_S, ld_s, n, d_rep, n_super, d_off, d_mb, d_mr,
                                                                          d_kblk, nk, d_D, nk, d_flags);
        WFO_LAUNCH_CHECK(ctx);
        std::vector<unsigned char> flags((size_t)n_super * nk);
        WFO_CUDA(cudaMemcpyAsync(flags.data(), d_flags, flags.size(), cudaMemcpyDeviceToHost, st));
        WFO_CUDA(cudaStreamSynchronize(st));
        std::vector<int> redo;   // every member of a super-block with a flagged ket block: the whole rows again
        for (int sb = 0; sb < n_super; sb++) {
            bool bad = false;
            for (int k = 0; k < nk && !bad; k++) bad = flags[(size_t)sb * nk + k] != 0;
            if (bad)
                for (int m = mem_off[sb]; m < mem_off[sb + 1]; m++) redo.push_back(mem_block[m]);
        }
        stats[2] = (int)redo.size();
        WFO_TRY(brute(redo));
    }
    return WFO_OK;
}

}  // namespace

int wfo_overlap_from_lists_host(wfo_ctx *ctx, const double *s_mo, int n_rows, int n_cols, int n_alpha,
                                const int32_t *ba_blocks, int nba, const int32_t *ka_blocks, int nka, int n_beta,
                                const int32_t *bb_blocks, int nbb, const int32_t *kb_blocks, int nkb, int nd_b,
                                const int32_t *ba_of, const int32_t *bb_of, const double *bc, int ns_b, int nd_k,
                                const int32_t *ka_of, const int32_t *kb_of, const double *kc, int ns_k, double *out,
                                int32_t *info) {
    if (!ctx || !s_mo || !bc || !kc || !out) return set_err(WFO_ERR_ARG, "overlap_from_lists: NULL argument");
    if (n_rows < 1 || n_cols < 1 || n_alpha < 0 || n_beta < 0 || nd_b < 0 || nd_k < 0 || ns_b < 1 || ns_k < 1)
        return set_err(WFO_ERR_ARG, "overlap_from_lists: bad sizes");
    if (n_alpha > 128 || n_beta > 128) return set_err(WFO_ERR_ARG, "overlap_from_lists: at most 128 electrons per spin");
    if ((n_alpha > 0 && (!ba_blocks || !ka_blocks || !ba_of || !ka_of)) || (n_beta > 0 && (!bb_blocks || !kb_blocks || !bb_of || !kb_of)))
        return set_err(WFO_ERR_ARG, "overlap_from_lists: block lists / maps missing");
    auto in_range = [](const int32_t *v, size_t cnt, int hi) {
        for (size_t i = 0; i < cnt; i++)
            if (v[i] < 0 || v[i] >= hi) return false;
        return true;
    };
    if (n_alpha > 0 && !(in_range(ba_blocks, (size_t)nba * n_alpha, n_rows) && in_range(ka_blocks, (size_t)nka * n_alpha, n_cols) &&
                         in_range(ba_of, nd_b, nba) && in_range(ka_of, nd_k, nka)))
        return set_err(WFO_ERR_ARG, "overlap_from_lists: alpha index out of range");
    if (n_beta > 0 && !(in_range(bb_blocks, (size_t)nbb * n_beta, n_rows) && in_range(kb_blocks, (size_t)nkb * n_beta, n_cols) &&
                        in_range(bb_of, nd_b, nbb) && in_range(kb_of, nd_k, nkb)))
        return set_err(WFO_ERR_ARG, "overlap_from_lists: beta index out of range");
    if (info) info[0] = info[1] = info[2] = info[3] = 0;
    if (nd_b == 0 || nd_k == 0) {
        memset(out, 0, (size_t)ns_b * ns_k * sizeof(double));
        return WFO_OK;
    }
    WFO_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    // arena: inputs, tables, temporaries of the brute-force / cofactor paths (generous upper bound)
    const size_t nS = (size_t)n_rows * n_cols;
    size_t bytes = 0;
    auto add = [&](size_t b) { bytes += align_up(b, 256); };
    add(nS * 8);
    for (int spin = 0; spin < 2; spin++) {
        const int ne = spin ? n_beta : n_alpha, nb = spin ? nbb : nba, nk = spin ? nkb : nka;
        if (ne == 0) continue;
        add((size_t)nb * ne * 4); add((size_t)nk * ne * 4); add((size_t)nb * nk * 8);
        add((size_t)nd_b * 4); add((size_t)nd_k * 4);
        // brute (twice: solo + redo): lists, row map, temp table; cofactor: representatives, members, flags
        add((size_t)nb * ne * 4); add((size_t)nb * 4); add((size_t)nb * nk * 8);
        add((size_t)nb * ne * 4); add((size_t)nb * 4); add((size_t)nb * nk * 8);
        add((size_t)nb * ne * 4); add((size_t)(nb + 1) * 4); add((size_t)nb * 4); add((size_t)nb * 4); add((size_t)nb * nk);
    }
    add((size_t)nd_b * ns_b * 8); add((size_t)nd_k * ns_k * 8); add((size_t)nd_k * ns_b * 8); add((size_t)ns_b * ns_k * 8);
    WFO_TRY(ensure_scratch(ctx, bytes + 4096));
    Arena ar{ctx->scratch, ctx->scratch_bytes, 0};
    double *d_S = ar.take<double>(nS);
    WFO_CUDA(cudaMemcpyAsync(d_S, s_mo, nS * 8, cudaMemcpyHostToDevice, st));
    double *d_tab[2] = {nullptr, nullptr};
    int32_t *d_bof[2] = {nullptr, nullptr}, *d_kof[2] = {nullptr, nullptr};
    int ld_tab[2] = {0, 0};
    for (int spin = 0; spin < 2; spin++) {
        const int ne = spin ? n_beta : n_alpha, nb = spin ? nbb : nba, nk = spin ? nkb : nka;
        if (ne == 0) continue;
        const int32_t *hb = spin ? bb_blocks : ba_blocks, *hk = spin ? kb_blocks : ka_blocks;
        int32_t *d_b = ar.take<int32_t>((size_t)nb * ne), *d_k = ar.take<int32_t>((size_t)nk * ne);
        d_tab[spin] = ar.take<double>((size_t)nb * nk);
        ld_tab[spin] = nk;
        d_bof[spin] = ar.take<int32_t>(nd_b);
        d_kof[spin] = ar.take<int32_t>(nd_k);
        WFO_CUDA(cudaMemcpyAsync(d_b, hb, (size_t)nb * ne * 4, cudaMemcpyHostToDevice, st));
        WFO_CUDA(cudaMemcpyAsync(d_k, hk, (size_t)nk * ne * 4, cudaMemcpyHostToDevice, st));
        WFO_CUDA(cudaMemcpyAsync(d_bof[spin], spin ? bb_of : ba_of, (size_t)nd_b * 4, cudaMemcpyHostToDevice, st));
        WFO_CUDA(cudaMemcpyAsync(d_kof[spin], spin ? kb_of : ka_of, (size_t)nd_k * 4, cudaMemcpyHostToDevice, st));
        int stats[3];
        WFO_TRY(det_table_super(ctx, d_S, n_cols, ne, hb, nb, d_b, d_k, nk, d_tab[spin], ar, st, stats));
        if (info) { info[0] += stats[0]; info[1] += stats[1]; info[2] += stats[2]; }
    }
    double *d_bc = ar.take<double>((size_t)nd_b * ns_b), *d_kc = ar.take<double>((size_t)nd_k * ns_k);
    double *d_tmp = ar.take<double>((size_t)nd_k * ns_b), *d_out = ar.take<double>((size_t)ns_b * ns_k);
    WFO_CUDA(cudaMemcpyAsync(d_bc, bc, (size_t)nd_b * ns_b * 8, cudaMemcpyHostToDevice, st));
    WFO_CUDA(cudaMemcpyAsync(d_kc, kc, (size_t)nd_k * ns_k * 8, cudaMemcpyHostToDevice, st));
    lists_contract_kernel<<<nd_k, 256, 0, st>>>(d_tab[0], ld_tab[0], d_tab[1], ld_tab[1], d_bof[0], d_bof[1], nd_b, d_kof[0],
                                                d_kof[1], d_bc, ns_b, d_tmp);
    WFO_LAUNCH_CHECK(ctx);
    // out (ns_b x ns_k) = tmp^T (ns_b x nd_k) kc (nd_k x ns_k)
    WFO_TRY(gemm(ctx, true, false, ns_b, ns_k, nd_k, 1.0, d_tmp, ns_b, d_kc, ns_k, 0.0, d_out, ns_k, 0, 0, st));
    WFO_CUDA(cudaMemcpyAsync(out, d_out, (size_t)ns_b * ns_k * 8, cudaMemcpyDeviceToHost, st));
    WFO_CUDA(cudaStreamSynchronize(st));
    return WFO_OK;
}

int wfo_set_lu_variant(wfo_ctx *ctx, int variant) {
    if (!ctx || variant < WFO_LU_AUTO || variant > WFO_LU_RIGHT) return set_err(WFO_ERR_ARG, "set_lu_variant: bad argument");
    ctx->lu_variant = variant;
    return WFO_OK;
}

int wfo_last_tier(const wfo_ctx *ctx) { return ctx ? ctx->last_tier : -1; }
int wfo_set_occ_hint(wfo_ctx *ctx, int n_occ) {
    if (!ctx) return set_err(WFO_ERR_ARG, "ctx is NULL");
    if (n_occ < 0) return set_err(WFO_ERR_ARG, "wfo_set_occ_hint: n_occ must be >= 0");
    ctx->occ_hint = n_occ;
    ctx->pre_smo = nullptr;
    return WFO_OK;
}
int wfo_set_timing(wfo_ctx *ctx, int on) {
    if (!ctx) return set_err(WFO_ERR_ARG, "ctx is NULL");
    ctx->timing = on;
    return WFO_OK;
}
int wfo_last_phases(wfo_ctx *ctx, double *ms, char *names, int names_cap) {
    if (!ctx) return 0;
    if (names && names_cap > 0) names[0] = 0;
    if (ctx->timing < 2 || ctx->n_phase < 2) return 0;
    if (cudaEventSynchronize(ctx->pev[ctx->n_phase - 1]) != cudaSuccess) return 0;
    size_t used = 0;
    for (int i = 1; i < ctx->n_phase; i++) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, ctx->pev[i - 1], ctx->pev[i]) != cudaSuccess) return 0;
        if (ms) ms[i - 1] = t;
        if (names) {
            int w = snprintf(names + used, used < (size_t)names_cap ? names_cap - used : 0, "%s%s", i > 1 ? "," : "",
                             ctx->pname[i]);
            if (w > 0) used += (size_t)w;
        }
    }
    return ctx->n_phase - 1;
}

double wfo_last_kernel_ms(const wfo_ctx *ctx) {
    if (!ctx || !ctx->timing) return -1.0;
    float ms = -1.f;
    if (cudaEventSynchronize(ctx->ev1) != cudaSuccess) return -1.0;
    if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) != cudaSuccess) return -1.0;
    return ms;
}

/* ---------------------------------------------------------------- stage 1 */
int wfo_gemm_dev(wfo_ctx *ctx, int ta, int tb, int m, int n, int k, double alpha, const double *A, int lda,
                 const double *B, int ldb, double beta, double *C, int ldc, void *stream) {
    if (!ctx) return set_err(WFO_ERR_ARG, "ctx is NULL");
    WFO_CUDA(cudaSetDevice(ctx->device));
    return gemm(ctx, ta != 0, tb != 0, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, 0, 0, (cudaStream_t)stream);
}

int wfo_sao_dev(wfo_ctx *ctx, const double *c_inv, int n, double *s_ao, void *stream) {
    if (!ctx) return set_err(WFO_ERR_ARG, "ctx is NULL");
    WFO_CUDA(cudaSetDevice(ctx->device));
    // S_AO = Cinv Cinv^T  (pywfo/main.py:304)
    return gemm(ctx, false, true, n, n, n, 1.0, c_inv, n, c_inv, n, 0.0, s_ao, n, 0, 0, (cudaStream_t)stream);
}

int wfo_smo_dev(wfo_ctx *ctx, const double *bra_mos, const double *s_ao, const double *ket_mos, int p, int u,
                int v, int q, double *s_mo, double *work, void *stream) {
    if (!ctx) return set_err(WFO_ERR_ARG, "ctx is NULL");
    WFO_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (!work) {
        WFO_TRY(ensure_scratch(ctx, (size_t)p * v * sizeof(double) + 256));
        work = reinterpret_cast<double *>(ctx->scratch);
    }
    // (bra S_AO) then (.) ket^T, the order of pywfo/main.py:532
    WFO_TRY(gemm(ctx, false, false, p, v, u, 1.0, bra_mos, u, s_ao, v, 0.0, work, v, 0, 0, st));
    WFO_TRY(gemm(ctx, false, true, p, q, v, 1.0, work, v, ket_mos, v, 0.0, s_mo, q, 0, 0, st));
    return WFO_OK;
}

/* Newton-Schulz inverse on the DMMA GEMM (replaces np.linalg.inv, pywfo/main.py:303) */
constexpr int NS_MAX_ITERS = 51;   // enough for cond ~ 1e6; anything slower goes to the pivoted elimination
}  // extern "C"

// Pivoted Gauss-Jordan inverse of the n x n matrix a (leading dimension lda) on one cooperative grid (csrc/inverse_gj.cuh):
// a_inv (n x n, contiguous), W and V n x n work matrices.  Queued on st, no synchronisation.  The pivot rows, pivot
// values and the singular flag stay on the device (*piv_out, *dval_out, *state_out) for the caller to look at.
static int gj_invert(wfo_ctx *ctx, const double *a, int lda, int n, double *a_inv, double *W, double *V, cudaStream_t st,
                     int **piv_out, double **dval_out, GjState **state_out) {
    WFO_CUDA(cudaMemcpy2DAsync(W, (size_t)n * sizeof(double), a, (size_t)lda * sizeof(double), (size_t)n * sizeof(double), n,
                               cudaMemcpyDeviceToDevice, st));
    int per_sm = 1;
    WFO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gj_inverse_kernel, GJ_THREADS, 0));
    int G = ctx->sm_count * (per_sm < 1 ? 1 : (per_sm > 2 ? 2 : per_sm));
    if (G > n) G = n;
    // bookkeeping: pivots, pivot values, candidates, used flags, state
    const size_t bytes = align_up((size_t)n * 4, 16) + (size_t)n * 8 + (size_t)G * sizeof(GjCand) + align_up((size_t)n, 16) + 16;
    if (bytes > ctx->gj_bytes) {
        WFO_CUDA(cudaStreamSynchronize(st));
        if (ctx->gj_buf) WFO_CUDA(cudaFree(ctx->gj_buf));
        ctx->gj_buf = nullptr;
        ctx->gj_bytes = 0;
        WFO_CUDA(cudaMalloc(&ctx->gj_buf, bytes));
        ctx->gj_bytes = bytes;
    }
    char *q = ctx->gj_buf;
    int *piv = reinterpret_cast<int *>(q); q += align_up((size_t)n * 4, 16);
    double *dval = reinterpret_cast<double *>(q); q += (size_t)n * 8;
    GjCand *cand = reinterpret_cast<GjCand *>(q); q += (size_t)G * sizeof(GjCand);
    unsigned char *used = reinterpret_cast<unsigned char *>(q); q += align_up((size_t)n, 16);
    GjState *state = reinterpret_cast<GjState *>(q);
    int nn_i = n;
    double *outp = a_inv;
    void *args[] = {&W, &V, &nn_i, &outp, &piv, &dval, &cand, &used, &state};
    WFO_CUDA(cudaLaunchCooperativeKernel((void *)gj_inverse_kernel, dim3(G), dim3(GJ_THREADS), args, 0, st));
    ctx->launches++;
    if (piv_out) *piv_out = piv;
    if (dval_out) *dval_out = dval;
    if (state_out) *state_out = state;
    return WFO_OK;
}

extern "C" {

int wfo_inverse_dev(wfo_ctx *ctx, const double *a, int n, double *a_inv, double *work, void *stream) {
    if (!ctx || !a || !a_inv) return set_err(WFO_ERR_ARG, "inverse: NULL argument");
    if (n < 1) return set_err(WFO_ERR_ARG, "inverse: n=%lld", n);
    WFO_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t nn = (size_t)n * n;
    if (!work) {
        WFO_TRY(ensure_scratch(ctx, 2 * align_up(nn * sizeof(double), 256)));
        work = reinterpret_cast<double *>(ctx->scratch);
    }
    ctx->last_inverse_iters = 0;
    ctx->last_inverse_pivoted = 0;
    if (n <= LUB_MAXN) {
        // small matrices: ONE launch of the register-resident Gauss-Jordan inverse with partial pivoting (the
        // kernel behind the base block of the rank-update tier) instead of a dozen launches of Newton-Schulz:
        // these sizes are launch-latency bound (bundled pywfo tests: n_mo = 24 ... 38)
        BaseInfo *d_info = reinterpret_cast<BaseInfo *>(ctx->ns_slots);
        base_inverse_kernel<<<1, LUB_THREADS, lub_smem_bytes(n), st>>>(a, n, n, a_inv, n, d_info, 1, 0, nullptr, 0);
        WFO_LAUNCH_CHECK(ctx);
        WFO_CUDA(cudaMemcpyAsync(ctx->h_slots, d_info, sizeof(BaseInfo), cudaMemcpyDeviceToHost, st));
        WFO_CUDA(cudaStreamSynchronize(st));
        ctx->last_inverse_pivoted = 1;
        if (reinterpret_cast<const BaseInfo *>(ctx->h_slots)->singular)
            return set_err(WFO_ERR_NOCONV, "inverse: singular matrix (zero pivot column)");
        return WFO_OK;
    }
    double *T = work, *Xb = work + align_up(nn * sizeof(double), 256) / sizeof(double);
    double *X = a_inv, *Xn = Xb;   // ping-pong; the result must end in a_inv
    double *slots = ctx->ns_slots;
    WFO_CUDA(cudaMemsetAsync(slots, 0, 128 * sizeof(double), st));
    // starting scale 1 / |A^T A|_inf: a rigorous bound on sigma_max^2 that is never worse than
    // |A|_1 |A|_inf and EXACT for orthonormal MO matrices (A^T A = I: the first iterate is the answer)
    WFO_TRY(gemm(ctx, true, false, n, n, n, 1.0, a, n, a, n, 0.0, T, n, 0, 0, st));
    ns_norms_kernel<<<n, 256, 0, st>>>(T, n, slots);
    WFO_LAUNCH_CHECK(ctx);
    ns_init_kernel<<<dim3(cdiv(n, 32), cdiv(n, 32)), dim3(32, 8), 0, st>>>(a, n, slots, X);
    WFO_LAUNCH_CHECK(ctx);
    const int rblocks = (int)((nn + 255) / 256 < (size_t)ctx->sm_count * 4 ? (nn + 255) / 256 : (size_t)ctx->sm_count * 4);
    int it = 0, checked = 0;
    bool converged = false, give_up = false;
    double prev = 1e300;
    while (it < NS_MAX_ITERS && !converged && !give_up) {
        // iterations between two looks at the residuals; the first look comes after ONE: orthonormal MO matrices
        // (A^T A = I) start at the answer, and four GEMMs on the way to the first synchronisation are 30-50 us
        const int batch = (it == 0) ? 1 : 3;
        for (int b = 0; b < batch && it < NS_MAX_ITERS; b++, it++) {
            // T = A X; residual max|T - I| -> slots[2 + it], Xn = X; Xn = 2 Xn - X T
            WFO_TRY(gemm(ctx, false, false, n, n, n, 1.0, a, n, X, n, 0.0, T, n, 0, 0, st));
            ns_residual_kernel<<<rblocks, 256, 0, st>>>(T, X, n, slots + 2 + it, Xn);
            WFO_LAUNCH_CHECK(ctx);
            WFO_TRY(gemm(ctx, false, false, n, n, n, -1.0, X, n, T, n, 2.0, Xn, n, 0, 0, st));
            double *tmp = X; X = Xn; Xn = tmp;
        }
        WFO_CUDA(cudaMemcpyAsync(ctx->h_slots, slots, 128 * sizeof(double), cudaMemcpyDeviceToHost, st));
        WFO_CUDA(cudaStreamSynchronize(st));
        // the residual seen at iteration i is squared by the update that follows it (max norm: at most
        // n r^2); once it is below 1e-4 it must keep collapsing: if it does not even halve, rounding has
        // taken over (cond >~ 1e7)
        for (; checked < it && !converged; checked++) {
            const double r = ctx->h_slots[2 + checked];
            if (!(r < 1e250)) give_up = true;                        // diverged / not finite
            else if (r < 1e-8) converged = true;
            else if (prev < 1e-4 && r > 0.5 * prev) give_up = true;   // stagnation
            prev = r;
        }
    }
    ctx->last_inverse_iters = it;
    if (converged) {
        if (X != a_inv) WFO_CUDA(cudaMemcpyAsync(a_inv, X, nn * sizeof(double), cudaMemcpyDeviceToDevice, st));
        return WFO_OK;
    }
    // ---- fallback: Gauss-Jordan with partial pivoting (what LAPACK's inverse is to the reference)
    {
        GjState *state = nullptr;
        WFO_TRY(gj_invert(ctx, a, n, n, a_inv, T, Xb, st, nullptr, nullptr, &state));
        GjState hs;
        WFO_CUDA(cudaMemcpyAsync(&hs, state, sizeof(hs), cudaMemcpyDeviceToHost, st));
        WFO_CUDA(cudaStreamSynchronize(st));
        ctx->last_inverse_pivoted = 1;
        if (hs.singular) return set_err(WFO_ERR_NOCONV, "inverse: singular matrix (zero pivot column in the pivoted elimination)");
    }
    return WFO_OK;
}

int wfo_inverse_host(wfo_ctx *ctx, const double *a, int n, double *a_inv) {
    if (!ctx || !a || !a_inv) return set_err(WFO_ERR_ARG, "inverse: NULL argument");
    if (n < 1) return set_err(WFO_ERR_ARG, "inverse: n=%lld", n);
    WFO_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t nn = (size_t)n * n, blk = align_up(nn * sizeof(double), 256);
    WFO_TRY(ensure_scratch(ctx, 4 * blk));
    double *d_a = reinterpret_cast<double *>(ctx->scratch), *d_x = reinterpret_cast<double *>(ctx->scratch + blk),
           *d_w = reinterpret_cast<double *>(ctx->scratch + 2 * blk);
    WFO_CUDA(cudaMemcpyAsync(d_a, a, nn * sizeof(double), cudaMemcpyHostToDevice, st));
    WFO_TRY(wfo_inverse_dev(ctx, d_a, n, d_x, d_w, st));
    WFO_CUDA(cudaMemcpyAsync(a_inv, d_x, nn * sizeof(double), cudaMemcpyDeviceToHost, st));
    WFO_CUDA(cudaStreamSynchronize(st));
    return WFO_OK;
}

/* S_MO straight from the inverse behind S_AO: S_MO = bra (Cinv Cinv^T) ket^T = (bra Cinv)(ket Cinv)^T.
 * Same product as pywfo/main.py:304 + :532 with the matrix chain associated the other way: two
 * launches (the two left factors as one batched GEMM, then P Q^T) instead of three dependent ones,
 * and S_AO is never materialised.  work: 2 n*n doubles. */
#ifdef WFO_DEBUG_PRE   // scratch builds: timeline of the early base-block chain against the main products
static cudaEvent_t g_dbg[8];
static bool g_dbg_init = false;
#define DBG_EV(i, s) do { if (!g_dbg_init) { for (auto &e_ : g_dbg) cudaEventCreate(&e_); g_dbg_init = true; } cudaEventRecord(g_dbg[i], s); } while (0)
#else
#define DBG_EV(i, s)
#endif
int wfo_smo_inv_dev(wfo_ctx *ctx, const double *bra_mos, const double *ket_mos, const double *c_inv, int n,
                    double *s_mo, double *work, void *stream) {
    if (!ctx || !bra_mos || !ket_mos || !c_inv || !s_mo) return set_err(WFO_ERR_ARG, "smo_inv: NULL argument");
    WFO_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t nn = (size_t)n * n;
    if (!work) {
        WFO_TRY(ensure_scratch(ctx, 2 * nn * sizeof(double) + 512));
        work = reinterpret_cast<double *>(ctx->scratch);
    }
    double *P = work, *Q = work + nn;
    const ptrdiff_t dbytes = reinterpret_cast<const char *>(ket_mos) - reinterpret_cast<const char *>(bra_mos);
    ctx->pre_smo = nullptr;
    const int h = ctx->occ_hint;
    if (h >= 1 && h <= LUB_MAXN && h <= n && n >= 96 && ctx->pre_probe == 0) {
        // once per context: the early chain needs its waiting kernel and the producers behind it to run concurrently
        int *d_res = nullptr, h_res = 0;
        WFO_CUDA(cudaMalloc(&d_res, sizeof(int)));
        for (int pass = 0; pass < 2; pass++) {
            // pass 0 loads the two kernels (lazy module loading takes longer than the probe waits): flag first
            const int epoch = ++ctx->pre_epoch;
            if (pass == 0) {
                set_flag_kernel<<<1, 32, 0, ctx->stream3>>>(ctx->pre_flag, epoch);
                WFO_LAUNCH_CHECK(ctx);
                WFO_CUDA(cudaStreamSynchronize(ctx->stream3));
            }
            concurrency_probe_kernel<<<1, 32, 0, ctx->stream4>>>(ctx->pre_flag, epoch, d_res);
            WFO_LAUNCH_CHECK(ctx);
            if (pass == 1) {
                set_flag_kernel<<<1, 32, 0, ctx->stream3>>>(ctx->pre_flag, epoch);
                WFO_LAUNCH_CHECK(ctx);
            }
            WFO_CUDA(cudaStreamSynchronize(ctx->stream3));
            WFO_CUDA(cudaStreamSynchronize(ctx->stream4));
        }
        WFO_CUDA(cudaMemcpy(&h_res, d_res, sizeof(int), cudaMemcpyDeviceToHost));
        WFO_CUDA(cudaFree(d_res));
        ctx->pre_probe = h_res ? 1 : -1;
    }
    if (h >= 1 && h <= LUB_MAXN && h <= n && n >= 96 && ctx->pre_probe > 0) {   // (tiny matrices: the five extra launches cost more than they hide)
        // the base block first, on its own stream: A = (bra[:h] Cinv)(ket[:h] Cinv)^T, then its inverse (one CTA for
        // ~0.8 us per column) next to the full products below instead of behind them
        const size_t hn = (size_t)h * n, hh = (size_t)h * h;
        // the base block as a split-K product (its K loop is the latency): at most 8 slabs, summed by the inverse kernel
        const int kslab = n <= 512 ? 64 : (int)align_up((size_t)cdiv(n, 8), 32), nslab = cdiv(n, kslab);
        const size_t want = (2 * hn + (size_t)(nslab + 1) * hh) * sizeof(double) + 256;
        if (want > ctx->pre_bytes) {
            WFO_CUDA(cudaDeviceSynchronize());
            if (ctx->pre_buf) WFO_CUDA(cudaFree(ctx->pre_buf));
            ctx->pre_buf = nullptr;
            ctx->pre_bytes = 0;
            WFO_CUDA(cudaMalloc(&ctx->pre_buf, want + want / 4));
            ctx->pre_bytes = want + want / 4;
        }
        double *Ls = reinterpret_cast<double *>(ctx->pre_buf), *Rs = Ls + hn, *A = Rs + hn, *Ai = A + (size_t)nslab * hh;
        BaseInfo *info = reinterpret_cast<BaseInfo *>(Ai + hh);
        cudaStream_t s3 = ctx->stream3, s4 = ctx->stream4;
        DBG_EV(0, st);
        WFO_CUDA(cudaEventRecord(ctx->ev_pre_in, st));
        WFO_CUDA(cudaStreamWaitEvent(s3, ctx->ev_pre_in, 0));
        WFO_CUDA(cudaStreamWaitEvent(s4, ctx->ev_pre_in, 0));
        // the inverse kernel goes FIRST: it takes an SM for itself and waits there for the flag
        const int epoch = ++ctx->pre_epoch;
        base_inverse_kernel<<<1, LUB_THREADS, LUB_EXCL_SMEM > lub_smem_bytes(h) ? LUB_EXCL_SMEM : lub_smem_bytes(h), s4>>>(
            A, h, h, Ai, h, info, nslab, (long long)hh, ctx->pre_flag, epoch);
        WFO_LAUNCH_CHECK(ctx);
        DBG_EV(3, s4);
        WFO_CUDA(cudaEventRecord(ctx->ev_pre, s4));
        if (bra_mos != ket_mos && dbytes % (ptrdiff_t)sizeof(double) == 0) {
            WFO_TRY(gemm(ctx, false, false, h, n, n, 1.0, bra_mos, n, c_inv, n, 0.0, Ls, n, 0, (long long)hn, s3, 2,
                         (long long)(dbytes / (ptrdiff_t)sizeof(double)), 0));
        } else {
            WFO_TRY(gemm(ctx, false, false, h, n, n, 1.0, bra_mos, n, c_inv, n, 0.0, Ls, n, 0, 0, s3));
            WFO_TRY(gemm(ctx, false, false, h, n, n, 1.0, ket_mos, n, c_inv, n, 0.0, Rs, n, 0, 0, s3));
        }
        DBG_EV(1, s3);
        WFO_TRY(gemm(ctx, false, true, h, h, n, 1.0, Ls, n, Rs, n, 0.0, A, h, kslab, (long long)hh, s3));
        set_flag_kernel<<<1, 32, 0, s3>>>(ctx->pre_flag, epoch);
        WFO_LAUNCH_CHECK(ctx);
        DBG_EV(2, s3);
        // the wide products below would fill every SM and leave this short dependent chain waiting for CTA slots
        // (measured: base block ready after 56 us instead of 12): they start behind it
        WFO_CUDA(cudaEventRecord(ctx->ev_pre_in, s3));
        WFO_CUDA(cudaStreamWaitEvent(st, ctx->ev_pre_in, 0));
        ctx->pre_smo = s_mo;
        ctx->pre_n = h;
        ctx->pre_Ai = Ai;
        ctx->pre_info = info;
        ctx->pre_A = A;
        ctx->pre_nslab = nslab;
    }
    if (bra_mos == ket_mos) {
        WFO_TRY(gemm(ctx, false, false, n, n, n, 1.0, bra_mos, n, c_inv, n, 0.0, P, n, 0, 0, st));
        Q = P;
    } else if (dbytes % (ptrdiff_t)sizeof(double) == 0) {
        // one launch, grid.z = 2: z = 0 -> P = bra Cinv, z = 1 -> Q = ket Cinv
        WFO_TRY(gemm(ctx, false, false, n, n, n, 1.0, bra_mos, n, c_inv, n, 0.0, P, n, 0, (long long)nn, st, 2,
                     (long long)(dbytes / (ptrdiff_t)sizeof(double)), 0));
    } else {
        WFO_TRY(gemm(ctx, false, false, n, n, n, 1.0, bra_mos, n, c_inv, n, 0.0, P, n, 0, 0, st));
        WFO_TRY(gemm(ctx, false, false, n, n, n, 1.0, ket_mos, n, c_inv, n, 0.0, Q, n, 0, 0, st));
    }
    DBG_EV(4, st);
    const int rc_s = gemm(ctx, false, true, n, n, n, 1.0, P, n, Q, n, 0.0, s_mo, n, 0, 0, st);
#ifdef WFO_DEBUG_PRE
    DBG_EV(5, st);
    if (ctx->pre_smo) {
        cudaDeviceSynchronize();
        float t[6] = {0};
        for (int i = 1; i < 6; i++) cudaEventElapsedTime(&t[i], g_dbg[0], g_dbg[i]);
        fprintf(stderr, "[pre] small LR done %.1f us, S_oo done %.1f, base done %.1f | main LR done %.1f, S done %.1f\n",
                t[1] * 1e3, t[2] * 1e3, t[3] * 1e3, t[4] * 1e3, t[5] * 1e3);
    }
#endif
    return rc_s;
}

int wfo_smo_host(wfo_ctx *ctx, const double *bra_mos, const double *s_ao, const double *ket_mos, int p, int u,
                 int v, int q, double *s_mo) {
    if (!ctx) return set_err(WFO_ERR_ARG, "ctx is NULL");
    WFO_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    size_t nb = (size_t)p * u, ns = (size_t)u * v, nk = (size_t)q * v, nw = (size_t)p * v, no = (size_t)p * q;
    Sizer sz;
    sz.take<double>(nb); sz.take<double>(ns); sz.take<double>(nk); sz.take<double>(nw); sz.take<double>(no);
    WFO_TRY(ensure_scratch(ctx, sz.off));
    Arena ar{ctx->scratch, ctx->scratch_bytes, 0};
    double *db = ar.take<double>(nb), *ds = ar.take<double>(ns), *dk = ar.take<double>(nk),
           *dw = ar.take<double>(nw), *dout = ar.take<double>(no);
    WFO_CUDA(cudaMemcpyAsync(db, bra_mos, nb * 8, cudaMemcpyHostToDevice, st));
    WFO_CUDA(cudaMemcpyAsync(ds, s_ao, ns * 8, cudaMemcpyHostToDevice, st));
    WFO_CUDA(cudaMemcpyAsync(dk, ket_mos, nk * 8, cudaMemcpyHostToDevice, st));
    WFO_TRY(wfo_smo_dev(ctx, db, ds, dk, p, u, v, q, dout, dw, st));
    WFO_CUDA(cudaMemcpyAsync(s_mo, dout, no * 8, cudaMemcpyDeviceToHost, st));
    WFO_CUDA(cudaStreamSynchronize(st));
    return WFO_OK;
}

/* ---------------------------------------------------------------- stage 2 */
int wfo_exc_lists_dev(wfo_ctx *ctx, const int32_t *exc, int k, int n_occ, int32_t *lists, void *stream) {
    if (!ctx) return set_err(WFO_ERR_ARG, "ctx is NULL");
    if (k <= 0) return WFO_OK;
    WFO_CUDA(cudaSetDevice(ctx->device));
    long long tot = (long long)k * n_occ;
    exc_lists_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(exc, k, n_occ, lists);
    WFO_LAUNCH_CHECK(ctx);
    return WFO_OK;
}

int wfo_det_table_dev(wfo_ctx *ctx, const double *s_mo, int n_mo, int ld_smo, int n, const int32_t *rows,
                      int kb, const int32_t *cols, int kk, double *D, void *stream) {
    if (!ctx) return set_err(WFO_ERR_ARG, "ctx is NULL");
    if (n > n_mo) return set_err(WFO_ERR_ARG, "det_table: n=%lld > n_mo=%lld", n, n_mo);
    WFO_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (ctx->timing) WFO_CUDA(cudaEventRecord(ctx->ev0, st));
    WFO_TRY(det_table(ctx, s_mo, ld_smo, n, rows, kb, cols, kk, D, st));
    if (ctx->timing) WFO_CUDA(cudaEventRecord(ctx->ev1, st));
    return WFO_OK;
}

int wfo_det_table_host(wfo_ctx *ctx, const double *s_mo, int n_mo, int n, const int32_t *rows, int kb,
                       const int32_t *cols, int kk, double *D) {
    if (!ctx) return set_err(WFO_ERR_ARG, "ctx is NULL");
    if (kb <= 0 || kk <= 0) return WFO_OK;
    WFO_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    size_t ns = (size_t)n_mo * n_mo, nr = (size_t)kb * n, nc = (size_t)kk * n, nd = (size_t)kb * kk;
    Sizer sz;
    sz.take<double>(ns); sz.take<int32_t>(nr); sz.take<int32_t>(nc); sz.take<double>(nd);
    WFO_TRY(ensure_scratch(ctx, sz.off));
    Arena ar{ctx->scratch, ctx->scratch_bytes, 0};
    double *ds = ar.take<double>(ns);
    int32_t *dr = ar.take<int32_t>(nr), *dc = ar.take<int32_t>(nc);
    double *dd = ar.take<double>(nd);
    WFO_CUDA(cudaMemcpyAsync(ds, s_mo, ns * 8, cudaMemcpyHostToDevice, st));
    WFO_CUDA(cudaMemcpyAsync(dr, rows, nr * 4, cudaMemcpyHostToDevice, st));
    WFO_CUDA(cudaMemcpyAsync(dc, cols, nc * 4, cudaMemcpyHostToDevice, st));
    WFO_TRY(wfo_det_table_dev(ctx, ds, n_mo, n_mo, n, dr, kb, dc, kk, dd, st));
    WFO_CUDA(cudaMemcpyAsync(D, dd, nd * 8, cudaMemcpyDeviceToHost, st));
    WFO_CUDA(cudaStreamSynchronize(st));
    return WFO_OK;
}

static size_t closed_form_bytes(int n, int v, int n_bra, int n_ket) {
    const size_t nv = (size_t)n * v;
    Sizer z;
    z.take<double>((size_t)n * n); z.take<double>(nv); z.take<double>(nv); z.take<double>((size_t)(v + 1) * (v + 1));
    z.take<double>((size_t)n_bra * nv); z.take<double>((size_t)n_ket * nv); z.take<double>((size_t)n_bra * nv);
    z.take<double>((size_t)n_bra * nv); z.take<double>((size_t)cdiv((int)nv, 1024) * n_bra * n_ket);
    z.take<double>(n_bra); z.take<double>(n_ket); z.take<BaseInfo>(1); z.take<int>(1);
    z.take<double>(base_work_doubles(n) + 1);
    return z.off;
}

// T2: closed-form contraction: pair determinants are never formed.
//   S_ij = det(A)^2 [ (1+sigma) <c~_i, X^T> <c~'_j, Y> + < Ai c~_i W , c~'_j > ]
// with dense signed CI tensors c~[s][f][t]; everything is DMMA GEMMs.
static int closed_form_impl(wfo_ctx *ctx, size_t arena_off, const double *s_mo, int n_mo, int n,
                            const int32_t *bra_exc, int kb, const double *Cb, int n_bra, const int32_t *ket_exc,
                            int kk, const double *Ck, int n_ket, double sigma, int k_begin, int k_end, double *out,
                            int ld_out, cudaStream_t st) {
    const int v = n_mo - n;
    const size_t nv = (size_t)n * v;
    const int kslab = 1024, n_slabs = cdiv((int)nv, kslab);
    Sizer sz;
    sz.off = arena_off;
    double *bwork = nullptr;
    auto carve2 = [&](auto &ar, double *&Ai, double *&Xt, double *&Y, double *&Wb, double *&cb, double *&ck, double *&Z,
                      double *&T, double *&P, double *&bx, double *&ky, BaseInfo *&info, int *&flag) {
        Ai = ar.template take<double>((size_t)n * n);
        Xt = ar.template take<double>(nv);
        Y = ar.template take<double>(nv);
        Wb = ar.template take<double>((size_t)(v + 1) * (v + 1));
        cb = ar.template take<double>((size_t)n_bra * nv);
        ck = ar.template take<double>((size_t)n_ket * nv);
        Z = ar.template take<double>((size_t)n_bra * nv);
        T = ar.template take<double>((size_t)n_bra * nv);
        P = ar.template take<double>((size_t)n_slabs * n_bra * n_ket);
        bx = ar.template take<double>(n_bra);
        ky = ar.template take<double>(n_ket);
        info = ar.template take<BaseInfo>(1);
        flag = ar.template take<int>(1);
        bwork = ar.template take<double>(base_work_doubles(n) + 1);
    };
    double *Ai, *Xt, *Y, *Wb, *cb, *ck, *Z, *T, *P, *bx, *ky;
    BaseInfo *info;
    int *flag;
    carve2(sz, Ai, Xt, Y, Wb, cb, ck, Z, T, P, bx, ky, info, flag);
    WFO_TRY(ensure_scratch(ctx, sz.off));
    Arena ar{ctx->scratch, ctx->scratch_bytes, arena_off};
    carve2(ar, Ai, Xt, Y, Wb, cb, ck, Z, T, P, bx, ky, info, flag);

    WFO_TRY(base_factor(ctx, s_mo, n_mo, n, Ai, n, info, bwork, st));
    WFO_CUDA(cudaMemsetAsync(cb, 0, (size_t)n_bra * nv * 8, st));
    WFO_CUDA(cudaMemsetAsync(ck, 0, (size_t)n_ket * nv * 8, st));
    WFO_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), st));
    scatter_signed_ci_kernel<<<dim3(cdiv(k_end - k_begin, 256), n_bra), 256, 0, st>>>(bra_exc, Cb, n_bra, kb, k_begin,
                                                                                   k_end, n, v, cb, flag);
    WFO_LAUNCH_CHECK(ctx);
    scatter_signed_ci_kernel<<<dim3(cdiv(kk, 256), n_ket), 256, 0, st>>>(ket_exc, Ck, n_ket, kk, 0, kk, n, v, ck, flag);
    WFO_LAUNCH_CHECK(ctx);
    struct { BaseInfo info; int flag; } h;
    WFO_CUDA(cudaMemcpyAsync(&h.info, info, sizeof(BaseInfo), cudaMemcpyDeviceToHost, st));
    WFO_CUDA(cudaMemcpyAsync(&h.flag, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    WFO_CUDA(cudaStreamSynchronize(st));
    if (h.flag) return set_err(WFO_ERR_ARG, "state_overlaps: tier CLOSED takes excitation tables without ground-state entries");
    if (h.info.singular || !(h.info.min_piv > 1e-6 * h.info.max_piv))
        return set_err(WFO_ERR_SINGULAR, "state_overlaps: base block S_MO[:occ,:occ] is (numerically) singular, "
                                         "tier CLOSED not applicable");
    const double *Bblk = s_mo + n;                    // occ  x virt
    const double *Cblk = s_mo + (size_t)n * n_mo;     // virt x occ
    // Xt = (C Ai)^T = Ai^T C^T (n x v), Y = Ai B (n x v), W = Dv - C Y (v x v, ld v+1)
    WFO_TRY(gemm(ctx, true, true, n, v, n, 1.0, Ai, n, Cblk, n_mo, 0.0, Xt, v, 0, 0, st));
    WFO_TRY(gemm(ctx, false, false, n, v, n, 1.0, Ai, n, Bblk, n_mo, 0.0, Y, v, 0, 0, st));
    long long wtot = (long long)(v + 1) * (v + 1);
    w_init_kernel<<<(unsigned)((wtot + 255) / 256), 256, 0, st>>>(s_mo, n_mo, n, v, Wb);
    WFO_LAUNCH_CHECK(ctx);
    WFO_TRY(gemm(ctx, false, false, v, v, n, -1.0, Cblk, n_mo, Y, v, 1.0, Wb, v + 1, 0, 0, st));
    if (ctx->timing) WFO_CUDA(cudaEventRecord(ctx->ev0, st));
    // Z = c~_bra W  ((n_bra n) x v);  T_i = Ai Z_i (batched);  P = T c~_ket^T (split K)
    WFO_TRY(gemm(ctx, false, false, n_bra * n, v, v, 1.0, cb, v, Wb, v + 1, 0.0, Z, v, 0, 0, st));
    WFO_TRY(gemm(ctx, false, false, n, v, n, 1.0, Ai, n, Z, v, 0.0, T, v, 0, (long long)nv, st, n_bra, 0, (long long)nv));
    WFO_TRY(gemm(ctx, false, true, n_bra, n_ket, (int)nv, 1.0, T, (int)nv, ck, (int)nv, 0.0, P, n_ket, kslab,
                 (long long)n_bra * n_ket, st));
    if (ctx->timing) WFO_CUDA(cudaEventRecord(ctx->ev1, st));
    rowdot_kernel<<<n_bra, 256, 0, st>>>(cb, (long long)nv, Xt, bx);
    WFO_LAUNCH_CHECK(ctx);
    rowdot_kernel<<<n_ket, 256, 0, st>>>(ck, (long long)nv, Y, ky);
    WFO_LAUNCH_CHECK(ctx);
    closed_finalize_kernel<<<cdiv(n_bra * n_ket, 128), 128, 0, st>>>(P, n_slabs, n_bra, n_ket, bx, ky, info, sigma, out,
                                                                     ld_out);
    WFO_LAUNCH_CHECK(ctx);
    ctx->last_tier = WFO_TIER_CLOSED;
    return WFO_OK;
}

/* ----------------------------------------------------------- stages 2 + 3 */
// `arena_off`: bytes at the start of the context arena owned by the caller (host wrappers)
constexpr int MAX_KET_BLOCK = 55;   // ket states per pass of the fused kernels: 55 + the D(k,0) column = 7 tiles of 8

// one block of at most MAX_KET_BLOCK ket states (any number for the closed form); out has row stride ld_out
static int state_overlaps_block(wfo_ctx *ctx, size_t arena_off, const double *s_mo, int n_mo, int n,
                                const int32_t *bra_exc, int kb, const double *Cb, int n_bra,
                                const int32_t *ket_exc, int kk, const double *Ck, int n_ket, double sigma,
                                int tier, int k_begin, int k_end, double *out, int ld_out, cudaStream_t st) {
    if (n < 1 || n > n_mo) return set_err(WFO_ERR_ARG, "state_overlaps: n_occ=%lld must be in 1..n_mo=%lld", n, n_mo);
    if (n_bra < 1 || n_ket < 1) return set_err(WFO_ERR_ARG, "state_overlaps: need at least one bra and one ket state");
    if (k_begin < 0 || k_end > kb || k_begin > k_end)
        return set_err(WFO_ERR_ARG, "state_overlaps: bad shard [%lld,%lld)", k_begin, k_end);
    bool want_t1 = false, may_t0 = false;
    WFO_TRY(resolve_tier(ctx, n, tier, &want_t1, &may_t0));
    const int v = n_mo - n;
    const int ksh = k_end - k_begin;
    if (ksh == 0 || kk == 0) {   // empty shard / no ket determinants: the partial sum is zero
        WFO_CUDA(cudaMemset2DAsync(out, (size_t)ld_out * sizeof(double), 0, (size_t)n_ket * sizeof(double), n_bra, st));
        ctx->last_tier = want_t1 ? WFO_TIER_UPDATE : WFO_TIER_LU;
        return WFO_OK;
    }
    if (tier == WFO_TIER_CLOSED)
        return closed_form_impl(ctx, arena_off, s_mo, n_mo, n, bra_exc, kb, Cb, n_bra, ket_exc, kk, Ck, n_ket, sigma,
                                k_begin, k_end, out, ld_out, st);
    Part q;
    WFO_TRY(make_part(ctx, ksh, kk, n_ket, may_t0, want_t1, &q));
    const int ldg = q.ldg, t0_chunks = q.t0_chunks, chunk_len = q.chunk_len, n_slabs = q.n_slabs;

    Plan p;
    memset(&p, 0, sizeof(p));
    Sizer sz;
    sz.off = arena_off;
    carve(sz, p, n, v, ksh, kk, n_bra, ldg, q.n_parts, n_slabs, may_t0, want_t1);
    WFO_TRY(ensure_scratch(ctx, sz.off));
    Arena ar{ctx->scratch, ctx->scratch_bytes, arena_off};
    carve(ar, p, n, v, ksh, kk, n_bra, ldg, q.n_parts, n_slabs, may_t0, want_t1);

    const int32_t *bra_sh = bra_exc + 2 * (size_t)k_begin;
    const double *d00_ptr = nullptr;
    int used = -1;
    PartMap pmap;
    memset(&pmap, 0, sizeof(pmap));
    ctx->n_phase = 0;
    WFO_PHASE(ctx, st, "start");
    const size_t ckt_smem = (size_t)CKT_TL * (ldg + 1) * sizeof(double);

    if (want_t1) {
        // ---- side stream: what does not depend on the base factorisation (signed, transposed ket
        // coefficients; the Dv block of W) runs next to the single-CTA base inverse
        cudaStream_t s2 = ctx->stream2;
        WFO_CUDA(cudaEventRecord(ctx->ev_fork, st));
        WFO_CUDA(cudaStreamWaitEvent(s2, ctx->ev_fork, 0));
        const int kk_pad = (int)align_up((size_t)kk, T1_LT);
        ckt_kernel<<<cdiv(kk_pad, CKT_TL), 256, ckt_smem, s2>>>(Ck, kk, kk_pad, n_ket, ket_exc, n, ldg, p.ckt);
        WFO_LAUNCH_CHECK(ctx);
        {
            long long wtot = (long long)(v + 1) * (v + 1);
            w_init_kernel<<<(unsigned)((wtot + 255) / 256), 256, 0, s2>>>(s_mo, n_mo, n, v, p.Wb);
            WFO_LAUNCH_CHECK(ctx);
        }
        WFO_CUDA(cudaEventRecord(ctx->ev_join, s2));
        // ---- one base factorisation (or the one wfo_smo_inv_dev started early for this S_MO), then the three block products
        if (ctx->pre_smo == s_mo && ctx->pre_n == n) {
            p.Ai = ctx->pre_Ai;
            p.info = ctx->pre_info;
            WFO_CUDA(cudaStreamWaitEvent(st, ctx->ev_pre, 0));
        } else {
            WFO_TRY(base_factor(ctx, s_mo, n_mo, n, p.Ai, n, p.info, p.bwork, st));
        }
        ctx->pre_smo = nullptr;   // consumed (or not ours): never reused
        // the conditioning of the base block decides whether the rank-update tier is usable.  The
        // 32-byte verdict travels to pinned host memory behind the kernel; the rank-update kernels
        // are queued right away (speculatively) and the host looks at the verdict after queueing
        // them: no bubble on the GPU, and the brute-force tier simply overwrites `out` if needed.
        WFO_PHASE(ctx, st, "base_inverse");
        WFO_CUDA(cudaStreamWaitEvent(st, ctx->ev_join, 0));
        WFO_CUDA(cudaEventRecord(ctx->ev_fork, st));            // base inverse done ...
        WFO_CUDA(cudaStreamWaitEvent(s2, ctx->ev_fork, 0));     // ... the copy rides on the side stream
        WFO_CUDA(cudaMemcpyAsync(ctx->h_info, p.info, sizeof(BaseInfo), cudaMemcpyDeviceToHost, s2));
        WFO_CUDA(cudaEventRecord(ctx->ev_info, s2));
        // everything behind the base inverse (queued speculatively; queued again if the pivoted fallback had to step in)
        auto queue_t1 = [&]() -> int {
            const double *Bblk = s_mo + n;                    // occ  x virt
            const double *Cblk = s_mo + (size_t)n * n_mo;     // virt x occ
            if (v > 0) {
                // X = C Ai on the side stream (behind the verdict copy) next to Y = Ai B -> W = D - C Y
                WFO_TRY(gemm(ctx, false, false, v, n, n, 1.0, Cblk, n_mo, p.Ai, n, 0.0, p.X, n, 0, 0, s2));
                WFO_CUDA(cudaEventRecord(ctx->ev_join, s2));
                WFO_TRY(gemm(ctx, false, false, n, v, n, 1.0, p.Ai, n, Bblk, n_mo, 0.0, p.Y, v, 0, 0, st));
                WFO_TRY(gemm(ctx, false, false, v, v, n, -1.0, Cblk, n_mo, p.Y, v, 1.0, p.Wb, v + 1, 0, 0, st));
                WFO_CUDA(cudaStreamWaitEvent(st, ctx->ev_join, 0));
            }
            WFO_PHASE(ctx, st, "xyw_gemms");
            t1_prep_kernel<<<cdiv(ksh + 1, 256) + cdiv(kk_pad, 256), 256, 0, st>>>(bra_sh, ksh, ket_exc, kk, kk_pad, n, v, p.X, p.Y, p.info,
                                                                          p.bmeta, p.bscale, p.dk0, p.kmeta, p.d0l);
            WFO_LAUNCH_CHECK(ctx);
            WFO_PHASE(ctx, st, "prep");
            if (ctx->timing) WFO_CUDA(cudaEventRecord(ctx->ev0, st));
            pmap = q.t1;
            {
                int nt, nd, rem;
                t1_split(n_ket, &nt, &nd, &rem);
                switch (nt * 100 + nd * 10 + rem) {
#define T1_CASE(NT_, ND_, REM_) case NT_ * 100 + ND_ * 10 + REM_: t1_fused_kernel<NT_, ND_, REM_><<<pmap.G, T1_WARPS * 32, t1_smem_bytes<NT_>(), st>>>(p.Ai, p.Wb, p.bmeta, p.bscale, ksh + 1, p.kmeta, p.ckt, kk, pmap, p.gpart); break;
                    WFO_T1_VARIANTS(T1_CASE)
#undef T1_CASE
                    default: return set_err(WFO_ERR_ARG, "state_overlaps: no rank-update kernel for %lld ket states", n_ket);
                }
            }
            WFO_LAUNCH_CHECK(ctx);
            if (ctx->timing) WFO_CUDA(cudaEventRecord(ctx->ev1, st));
            WFO_PHASE(ctx, st, "t1_fused");
            d00_ptr = &p.info->det;
            used = WFO_TIER_UPDATE;
            return WFO_OK;
        };
        WFO_TRY(queue_t1());
        WFO_CUDA(cudaEventSynchronize(ctx->ev_info));
        const BaseInfo &hinfo = *ctx->h_info;
        if (hinfo.singular == 2) {
            // the early inverse kernel (wfo_smo_inv_dev with the occupation hint) gave up waiting for its input:
            // kernels do not run concurrently here (a profiler, CUDA_LAUNCH_BLOCKING, ...).  Same input, same
            // kernel, now behind the producer; then the rank-update stages again.
            base_inverse_kernel<<<1, LUB_THREADS, lub_smem_bytes(n), st>>>(ctx->pre_A, n, n, p.Ai, n, p.info, ctx->pre_nslab,
                                                                        (long long)n * n, nullptr, 0);
            WFO_LAUNCH_CHECK(ctx);
            WFO_CUDA(cudaMemcpyAsync(ctx->h_info, p.info, sizeof(BaseInfo), cudaMemcpyDeviceToHost, st));
            WFO_CUDA(cudaStreamSynchronize(st));
            if (v > 0) {   // W's D block again (the first pass subtracted garbage from it)
                long long wtot = (long long)(v + 1) * (v + 1);
                w_init_kernel<<<(unsigned)((wtot + 255) / 256), 256, 0, st>>>(s_mo, n_mo, n, v, p.Wb);
                WFO_LAUNCH_CHECK(ctx);
            }
            WFO_TRY(queue_t1());
        }
        bool bad_base = hinfo.singular || !(hinfo.min_piv > 1e-6 * hinfo.max_piv);
        if (bad_base && n > LUB_MAXN) {
            // the blockwise (Schur) inverse pivots inside its 128-blocks only: a well-conditioned base block with a
            // singular leading half (occupied orbitals reordered between the geometries) fails there although
            // LAPACK, i.e. the reference, handles it.  Fall back to the pivoted Gauss-Jordan inverse of the whole
            // block (csrc/inverse_gj.cuh) and queue the rank-update stages again.
            int *piv = nullptr;
            double *dval = nullptr;
            GjState *state = nullptr;
            WFO_TRY(gj_invert(ctx, s_mo, n_mo, n, p.Ai, p.gjwork, p.gjwork + (size_t)n * n, st, &piv, &dval, &state));
            gj_info_kernel<<<1, 256, 0, st>>>(piv, dval, n, state, p.info);
            WFO_LAUNCH_CHECK(ctx);
            WFO_CUDA(cudaMemcpyAsync(ctx->h_info, p.info, sizeof(BaseInfo), cudaMemcpyDeviceToHost, st));
            WFO_CUDA(cudaStreamSynchronize(st));
            bad_base = hinfo.singular || !(hinfo.min_piv > 1e-6 * hinfo.max_piv);
            if (!bad_base) {
                if (v > 0) {   // W's D block again (the first pass subtracted garbage from it)
                    long long wtot = (long long)(v + 1) * (v + 1);
                    w_init_kernel<<<(unsigned)((wtot + 255) / 256), 256, 0, st>>>(s_mo, n_mo, n, v, p.Wb);
                    WFO_LAUNCH_CHECK(ctx);
                }
                WFO_TRY(queue_t1());
            }
        }
        if (bad_base) {
            used = -1;   // what was queued computes garbage into buffers the brute-force tier rewrites
            if (!may_t0)
                return set_err(WFO_ERR_SINGULAR,
                               "state_overlaps: base block S_MO[:occ,:occ] is (numerically) singular, "
                               "tier UPDATE not applicable");
        }
    }
    if (used < 0) {
        if (!may_t0)
            return set_err(WFO_ERR_SINGULAR, "state_overlaps: base block singular and n_occ too large for tier LU");
        // ---- brute force: ground-state row/column/corner first, then the fused pair kernel
        iota_kernel<<<cdiv(n, 128), 128, 0, st>>>(p.gs_list, n);
        WFO_LAUNCH_CHECK(ctx);
        WFO_TRY(wfo_exc_lists_dev(ctx, bra_sh, ksh, n, p.bra_lists, st));
        WFO_TRY(wfo_exc_lists_dev(ctx, ket_exc, kk, n, p.ket_lists, st));
        WFO_TRY(det_table(ctx, s_mo, n_mo, n, p.bra_lists, ksh, p.gs_list, 1, p.dk0, st));
        WFO_TRY(det_table(ctx, s_mo, n_mo, n, p.gs_list, 1, p.ket_lists, kk, p.d0l, st));
        WFO_TRY(det_table(ctx, s_mo, n_mo, n, p.gs_list, 1, p.gs_list, 1, p.d00, st));
        ckt_kernel<<<cdiv(kk, CKT_TL), 256, ckt_smem, st>>>(Ck, kk, kk, n_ket, nullptr, n, ldg, p.ckt);
        WFO_LAUNCH_CHECK(ctx);
        const long long items = (long long)ksh * t0_chunks;
        unsigned grid = 1;
        if (use_ll_kernels(ctx, n)) {
#define CALL_FL(NT_, MG_)                                                                                         \
    WFO_TRY(persistent_grid(ctx, t0_fused_ll_kernel<NT_, MG_>, LLWarps<NT_, MG_>::W * 32,                             \
                            LLWarps<NT_, MG_>::W * LLCfg<NT_, MG_>::BYTES,                                            \
                            (items + LLWarps<NT_, MG_>::W - 1) / LLWarps<NT_, MG_>::W, &grid));                       \
    if (MG_) WFO_TRY(ensure_gscratch(ctx, (size_t)grid * LLWarps<NT_, MG_>::W * ll_scratch_doubles(NT_) * 8));        \
    WFO_CUDA(cudaMemsetAsync(ctx->t0_next, 0, sizeof(unsigned long long), st));                                       \
    if (ctx->timing) WFO_CUDA(cudaEventRecord(ctx->ev0, st));                                                         \
    t0_fused_ll_kernel<NT_, MG_><<<grid, LLWarps<NT_, MG_>::W * 32, LLWarps<NT_, MG_>::W * LLCfg<NT_, MG_>::BYTES, st>>>( \
        s_mo, n_mo, n, p.bra_lists, ksh, p.ket_lists, kk, p.ckt, ldg, chunk_len, t0_chunks, p.gpart, ctx->gscratch, ctx->t0_next)
            WFO_LL_DISPATCH(n, CALL_FL);
#undef CALL_FL
        } else if (n > 32) {
#define CALL_FP(SL_)                                                                                              \
    WFO_TRY(persistent_grid(ctx, t0_fused_wpm_kernel<SL_>, WPM_WARPS * 32, 0, (items + WPM_WARPS - 1) / WPM_WARPS, &grid)); \
    WFO_TRY(ensure_gscratch(ctx, (size_t)grid * WPM_WARPS * wpm_mat_bytes(n)));                                    \
    if (ctx->timing) WFO_CUDA(cudaEventRecord(ctx->ev0, st));                                                      \
    t0_fused_wpm_kernel<SL_><<<grid, WPM_WARPS * 32, 0, st>>>(s_mo, n_mo, n, p.bra_lists, ksh, p.ket_lists, kk,    \
                                                              p.ckt, ldg, chunk_len, t0_chunks, p.gpart, ctx->gscratch)
            WFO_WPM_DISPATCH(n, CALL_FP);
#undef CALL_FP
        } else {
#define CALL_FW(NC_, SEG_)                                                                                      \
    WFO_TRY(persistent_grid(ctx, t0_fused_warp_kernel<NC_, SEG_>, 128, 0, (items + 3) / 4, &grid));             \
    if (ctx->timing) WFO_CUDA(cudaEventRecord(ctx->ev0, st));                                                   \
    t0_fused_warp_kernel<NC_, SEG_><<<grid, 128, 0, st>>>(s_mo, n_mo, n, p.bra_lists, ksh, p.ket_lists, kk,     \
                                                          p.ckt, ldg, chunk_len, t0_chunks, p.gpart)
            WFO_WARP_DISPATCH(n, CALL_FW);
#undef CALL_FW
        }
        WFO_LAUNCH_CHECK(ctx);
        if (ctx->timing) WFO_CUDA(cudaEventRecord(ctx->ev1, st));
        WFO_PHASE(ctx, st, "t0_path");
        d00_ptr = p.d00;
        used = WFO_TIER_LU;
        pmap.I = 0;
        pmap.uniform_parts = t0_chunks;
    }

    // ---- CI contraction: P[z] = Cb[:, slab z] (sum of partials of slab z), then the final sums
    const int ksh_g = (used == WFO_TIER_UPDATE) ? ksh + 1 : ksh;
    int n_wparts = 1;
    if (used == WFO_TIER_LU) {   // the rank-update tier gets w from its virtual ground-state row
        ket_gs_dot_kernel<<<dim3(n_ket, KGS_SLABS), 256, 0, st>>>(Ck, kk, n_ket, ldg, p.d0l, p.wvec);
        WFO_LAUNCH_CHECK(ctx);
        n_wparts = KGS_SLABS;
    }
    reduce_contract_kernel<<<dim3(cdiv(ksh_g, RC_ROWS), ldg / 8), 256, 0, st>>>(
        p.gpart, pmap, ksh_g, ksh, ldg, n_ket, p.dk0, Cb + k_begin, kb, n_bra, p.P, p.wvec,
        used == WFO_TIER_UPDATE ? ksh : -1);
    WFO_LAUNCH_CHECK(ctx);
    WFO_PHASE(ctx, st, "reduce_contract");
    finalize3_kernel<<<n_bra, 256, 0, st>>>(p.P, cdiv(ksh_g, RC_ROWS), n_bra, ldg, n_ket, p.wvec, n_wparts, d00_ptr, sigma,
                                            out, ld_out);
    WFO_LAUNCH_CHECK(ctx);
    WFO_PHASE(ctx, st, "finalize");
    ctx->last_tier = used;
    return WFO_OK;
}

// any number of ket states: the fused tiers run in passes of MAX_KET_BLOCK states (the bra side and
// the determinants are the same in every pass; only the ket coefficient columns change)
static int state_overlaps_impl(wfo_ctx *ctx, size_t arena_off, const double *s_mo, int n_mo, int n,
                               const int32_t *bra_exc, int kb, const double *Cb, int n_bra,
                               const int32_t *ket_exc, int kk, const double *Ck, int n_ket, double sigma,
                               int tier, int k_begin, int k_end, double *out, cudaStream_t st) {
    if (tier == WFO_TIER_CLOSED || n_ket <= MAX_KET_BLOCK)
        return state_overlaps_block(ctx, arena_off, s_mo, n_mo, n, bra_exc, kb, Cb, n_bra, ket_exc, kk, Ck, n_ket,
                                    sigma, tier, k_begin, k_end, out, n_ket, st);
    for (int j0 = 0; j0 < n_ket; j0 += MAX_KET_BLOCK) {
        const int nb = n_ket - j0 < MAX_KET_BLOCK ? n_ket - j0 : MAX_KET_BLOCK;
        WFO_TRY(state_overlaps_block(ctx, arena_off, s_mo, n_mo, n, bra_exc, kb, Cb, n_bra, ket_exc, kk,
                                     Ck + (size_t)j0 * kk, nb, sigma, tier, k_begin, k_end, out + j0, n_ket, st));
    }
    return WFO_OK;
}

int wfo_state_overlaps_dev(wfo_ctx *ctx, const double *s_mo, int n_mo, int n_occ, const int32_t *bra_exc, int kb,
                           const double *Cb, int n_bra, const int32_t *ket_exc, int kk, const double *Ck,
                           int n_ket, double sigma, int tier, int k_begin, int k_end, double *out, void *stream) {
    if (!ctx) return set_err(WFO_ERR_ARG, "ctx is NULL");
    WFO_CUDA(cudaSetDevice(ctx->device));
    return state_overlaps_impl(ctx, 0, s_mo, n_mo, n_occ, bra_exc, kb, Cb, n_bra, ket_exc, kk, Ck, n_ket, sigma,
                               tier, k_begin, k_end, out, (cudaStream_t)stream);
}

// shared by the two host front ends: uploads tables, runs, downloads
static int host_front(wfo_ctx *ctx, const double *s_mo_host, const double *bra_mos, const double *ket_mos,
                      const double *c_inv, int n_mo, int n, const int32_t *bra_exc, int kb, const double *Cb,
                      int n_bra, const int32_t *ket_exc, int kk, const double *Ck, int n_ket, double sigma,
                      int tier, int k_begin, int k_end, double *out) {
    if (!ctx) return set_err(WFO_ERR_ARG, "ctx is NULL");
    if (kb < 0 || kk < 0 || n_bra < 1 || n_ket < 1 || n_mo < 1)
        return set_err(WFO_ERR_ARG, "state_overlaps: negative or empty sizes");
    WFO_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    size_t nsq = (size_t)n_mo * n_mo;
    Sizer sz;
    sz.take<double>(nsq);                                     // S_MO
    if (!s_mo_host) { sz.take<double>(nsq); sz.take<double>(nsq); sz.take<double>(nsq); sz.take<double>(nsq); sz.take<double>(nsq); }
    sz.take<int32_t>(2 * (size_t)(kb + 1)); sz.take<int32_t>(2 * (size_t)(kk + 1));
    sz.take<double>((size_t)n_bra * (kb + 1)); sz.take<double>((size_t)n_ket * (kk + 1));
    sz.take<double>((size_t)n_bra * n_ket);
    size_t front = sz.off;
    {   // size the arena once for the front buffers plus the implementation's plan
        size_t total = front;
        const int v = n_mo - n, ksh = k_end - k_begin;
        bool want_t1 = false, may_t0 = false;
        if (n >= 1 && n <= n_mo && ksh > 0 && kk > 0 && k_begin >= 0 && k_end <= kb &&
            resolve_tier(ctx, n, tier, &want_t1, &may_t0) == WFO_OK) {
            Part q;
            WFO_TRY(make_part(ctx, ksh, kk, n_ket < MAX_KET_BLOCK ? n_ket : MAX_KET_BLOCK, may_t0, want_t1, &q));
            Plan p;
            Sizer s2;
            s2.off = front;
            carve(s2, p, n, v, ksh, kk, n_bra, q.ldg, q.n_parts, q.n_slabs, may_t0, want_t1);
            total = s2.off;
            if (tier == WFO_TIER_CLOSED) total = front + closed_form_bytes(n, v, n_bra, n_ket);
        }
        WFO_TRY(ensure_scratch(ctx, total));
    }
    Arena ar{ctx->scratch, ctx->scratch_bytes, 0};
    double *d_smo = ar.take<double>(nsq);
    if (!s_mo_host) {
        double *d_bra = ar.take<double>(nsq), *d_ket = ar.take<double>(nsq), *d_inv = ar.take<double>(nsq),
               *d_sao = ar.take<double>(nsq), *d_work = ar.take<double>(nsq);
        WFO_CUDA(cudaMemcpyAsync(d_bra, bra_mos, nsq * 8, cudaMemcpyHostToDevice, st));
        WFO_CUDA(cudaMemcpyAsync(d_ket, ket_mos, nsq * 8, cudaMemcpyHostToDevice, st));
        WFO_CUDA(cudaMemcpyAsync(d_inv, c_inv, nsq * 8, cudaMemcpyHostToDevice, st));
        WFO_TRY(wfo_sao_dev(ctx, d_inv, n_mo, d_sao, st));
        WFO_TRY(wfo_smo_dev(ctx, d_bra, d_sao, d_ket, n_mo, n_mo, n_mo, n_mo, d_smo, d_work, st));
    } else {
        WFO_CUDA(cudaMemcpyAsync(d_smo, s_mo_host, nsq * 8, cudaMemcpyHostToDevice, st));
    }
    int32_t *d_be = ar.take<int32_t>(2 * (size_t)(kb + 1)), *d_ke = ar.take<int32_t>(2 * (size_t)(kk + 1));
    double *d_cb = ar.take<double>((size_t)n_bra * (kb + 1)), *d_ck = ar.take<double>((size_t)n_ket * (kk + 1));
    double *d_out = ar.take<double>((size_t)n_bra * n_ket);
    if (kb > 0) {
        WFO_CUDA(cudaMemcpyAsync(d_be, bra_exc, 2 * (size_t)kb * 4, cudaMemcpyHostToDevice, st));
        WFO_CUDA(cudaMemcpyAsync(d_cb, Cb, (size_t)n_bra * kb * 8, cudaMemcpyHostToDevice, st));
    }
    if (kk > 0) {
        WFO_CUDA(cudaMemcpyAsync(d_ke, ket_exc, 2 * (size_t)kk * 4, cudaMemcpyHostToDevice, st));
        WFO_CUDA(cudaMemcpyAsync(d_ck, Ck, (size_t)n_ket * kk * 8, cudaMemcpyHostToDevice, st));
    }
    char *before = ctx->scratch;
    WFO_TRY(state_overlaps_impl(ctx, ar.off, d_smo, n_mo, n, d_be, kb, d_cb, n_bra, d_ke, kk, d_ck, n_ket, sigma,
                                tier, k_begin, k_end, d_out, st));
    if (ctx->scratch != before) return set_err(WFO_ERR_CUDA, "internal: arena moved during a host call");
    WFO_CUDA(cudaMemcpyAsync(out, d_out, (size_t)n_bra * n_ket * 8, cudaMemcpyDeviceToHost, st));
    WFO_CUDA(cudaStreamSynchronize(st));
    return WFO_OK;
}

int wfo_state_overlaps_host(wfo_ctx *ctx, const double *s_mo, int n_mo, int n_occ, const int32_t *bra_exc, int kb,
                            const double *Cb, int n_bra, const int32_t *ket_exc, int kk, const double *Ck,
                            int n_ket, double sigma, int tier, int k_begin, int k_end, double *out) {
    if (!s_mo) return set_err(WFO_ERR_ARG, "state_overlaps_host: s_mo is NULL");
    return host_front(ctx, s_mo, nullptr, nullptr, nullptr, n_mo, n_occ, bra_exc, kb, Cb, n_bra, ket_exc, kk, Ck,
                      n_ket, sigma, tier, k_begin, k_end, out);
}

int wfo_overlaps_host(wfo_ctx *ctx, const double *bra_mos, const double *ket_mos, const double *c_inv, int n_mo,
                      int n_occ, const int32_t *bra_exc, int kb, const double *Cb, int n_bra,
                      const int32_t *ket_exc, int kk, const double *Ck, int n_ket, double sigma, int tier,
                      int k_begin, int k_end, double *out) {
    if (!bra_mos || !ket_mos || !c_inv) return set_err(WFO_ERR_ARG, "overlaps_host: NULL MO pointer");
    return host_front(ctx, nullptr, bra_mos, ket_mos, c_inv, n_mo, n_occ, bra_exc, kb, Cb, n_bra, ket_exc, kk, Ck,
                      n_ket, sigma, tier, k_begin, k_end, out);
}

/* ---- the whole path from RAW CI tensors: selection of excitations on the device ---- */
struct DevTables {
    int32_t *exc;     // (K, 2)
    double *coeff;    // (n_states, K)
    int K;
};

// build one side's tables.  ci_dev: (n_states, n, v) on the device.  flags/offs/exc_buf are
// caller-provided device buffers of n*v ints / n*v ints / 2*n*v ints; totals: 1 int + n_states ints.
static int build_tables_dev(wfo_ctx *ctx, const double *ci_dev, int n_states, int n, int v, double thr, int strict,
                            int order, int *flags, int *offs, int *totals, cudaStream_t st, int pos0 = 0, int cnt = -1,
                            long long stride = -1) {
    const int count = n * v;
    if (cnt < 0) { cnt = count; stride = count; }
    if (count <= 0) return WFO_OK;
    if (cnt > 0) {
        union_flags_kernel<<<cdiv(cnt, 256), 256, 0, st>>>(ci_dev, n_states, n, v, thr, strict, order, flags, pos0, cnt, stride);
        WFO_LAUNCH_CHECK(ctx);
    }
    scan_flags_kernel<<<1, 1024, 0, st>>>(flags, cnt, offs, totals);
    WFO_LAUNCH_CHECK(ctx);
    if (cnt == count && strict) {   // per-state totals: only pywfo.main.overlaps looks at them (main.py:186), whole tensor only
        state_counts_kernel<<<n_states, 256, 0, st>>>(ci_dev, n, v, thr, strict, totals + 1);
        WFO_LAUNCH_CHECK(ctx);
    }
    return WFO_OK;
}

}  // extern "C"

// ---- raw-CI path: phase-1 workspace (sizes depend on n_mo, n_occ and the state counts only) lives in
// the context's RESIDENT buffer next to the inputs; the arena holds what is sized by the number of
// surviving excitations.  Neither moves during a call.
struct CiWork {
    double *smo, *sao, *work, *work2, *inv;   // n_mo^2 each; work and work2 are contiguous (device inverse)
    int *fl_b, *of_b, *fl_k, *of_k, *tot_b, *tot_k;
    int32_t *exc_b, *exc_k;
    double *out;
};
template <typename AR>
static void carve_ciwork(AR &ar, CiWork &w, size_t nsq, int count, int n_bra, int n_ket) {
    w.smo = ar.template take<double>(nsq);
    w.sao = ar.template take<double>(nsq);
    w.work = ar.template take<double>(nsq);
    w.work2 = ar.template take<double>(nsq);
    w.inv = ar.template take<double>(nsq);
    w.fl_b = ar.template take<int>(count);
    w.of_b = ar.template take<int>(count);
    w.fl_k = ar.template take<int>(count);
    w.of_k = ar.template take<int>(count);
    w.tot_b = ar.template take<int>(1 + n_bra);
    w.tot_k = ar.template take<int>(1 + n_ket);
    w.exc_b = ar.template take<int32_t>(2 * (size_t)count);
    w.exc_k = ar.template take<int32_t>(2 * (size_t)count);
    w.out = ar.template take<double>((size_t)n_bra * n_ket);
}

static int ensure_resident(wfo_ctx *ctx, size_t bytes) {
    if (bytes <= ctx->resident_bytes) return WFO_OK;
    WFO_CUDA(cudaDeviceSynchronize());
    if (ctx->resident) WFO_CUDA(cudaFree(ctx->resident));
    ctx->resident = nullptr;
    ctx->resident_bytes = 0;
    const size_t want = bytes + bytes / 8 + (1 << 20);
    WFO_CUDA(cudaMalloc(&ctx->resident, want));
    ctx->resident_bytes = want;
    return WFO_OK;
}

// One (bra, ket) pair from DEVICE inputs: S_AO, S_MO on the side stream (from d_inv if given, else the
// inverse of the chosen MO matrix is formed first), excitation tables on the main stream, then the
// fused stages on this rank's shard.  Result in w.out (device).  `zero` is set when the result is
// identically zero (no excitation survives) and w.out was not written.
static int overlaps_ci_core(wfo_ctx *ctx, const CiWork &w, const double *d_bra, const double *d_ket,
                            const double *d_inv, int ao_side, int n_mo, int n_occ, const double *d_bci, int n_bra,
                            const double *d_kci, int n_ket, double ci_thresh, int semantics, int tier, int rank,
                            int world, bool side_after_main, int32_t *info, bool *zero, cudaStream_t st,
                            int bra_pos0 = 0, int bra_cnt = -1, long long bra_stride = -1) {
    // bra window: d_bci points at position bra_pos0 of state 0 and covers bra_cnt positions (a rank's
    // slice of the (from, to) space, see wfo_overlaps_ci_host); default = the whole tensor
    const int n = n_occ, v = n_mo - n_occ, count = n * v;
    if (bra_cnt < 0) { bra_cnt = count; bra_stride = count; }
    const int strict = semantics == 1;
    const double sigma = semantics == 1 ? -1.0 : 1.0;
    *zero = false;
    cudaStream_t s2 = ctx->stream2;
    if (side_after_main) {   // the side stream may not overwrite S_MO / the scratch while earlier main-stream work reads them
        WFO_CUDA(cudaEventRecord(ctx->ev_fork, st));
        WFO_CUDA(cudaStreamWaitEvent(s2, ctx->ev_fork, 0));
    }
    WFO_TRY(build_tables_dev(ctx, d_bci, n_bra, n, v, ci_thresh, strict, 0, w.fl_b, w.of_b, w.tot_b, st, bra_pos0, bra_cnt,
                             bra_stride));
    WFO_TRY(build_tables_dev(ctx, d_kci, n_ket, n, v, ci_thresh, strict, 1, w.fl_k, w.of_k, w.tot_k, st));
    if (!d_inv) {
        WFO_TRY(wfo_inverse_dev(ctx, ao_side == WFO_AO_BRA ? d_bra : d_ket, n_mo, w.inv, w.work, s2));
        d_inv = w.inv;
    }
    {   // the base block early (the rank-update tier will want it), whatever hint the caller has set
        const int hint0 = ctx->occ_hint;
        static const bool no_early = getenv("WFO_NO_EARLY_BASE") != nullptr;   // tuning switch
        ctx->occ_hint = (!no_early && (tier == WFO_TIER_AUTO || tier == WFO_TIER_UPDATE)) ? n : 0;
        const int rc_ = wfo_smo_inv_dev(ctx, d_bra, d_ket, d_inv, n_mo, w.smo, w.work, s2);   // work, work2 contiguous
        ctx->occ_hint = hint0;
        if (rc_ != WFO_OK) return rc_;
    }
    WFO_CUDA(cudaEventRecord(ctx->ev_join, s2));
    WFO_CUDA(cudaStreamWaitEvent(st, ctx->ev_join, 0));
    // ---- the table sizes (and per-state counts) come back: 2 small copies, one sync
    std::vector<int> hb_v(1 + n_bra), hk_v(1 + n_ket);
    int *hb = hb_v.data(), *hk = hk_v.data();
    WFO_CUDA(cudaMemcpyAsync(hb, w.tot_b, (1 + n_bra) * sizeof(int), cudaMemcpyDeviceToHost, st));
    WFO_CUDA(cudaMemcpyAsync(hk, w.tot_k, (1 + n_ket) * sizeof(int), cudaMemcpyDeviceToHost, st));
    WFO_CUDA(cudaStreamSynchronize(st));
    const int kb = hb[0], kk = hk[0];
    if (info) { info[0] = kb; info[1] = kk; info[2] = 0; info[3] = -1; }
    if (semantics == 1 && bra_cnt == count) {   // pywfo/main.py:186: a state without excitations cannot be unpacked
        bool empty = false;
        for (int i = 0; i < n_bra; i++) empty = empty || hb[1 + i] == 0;
        for (int i = 0; i < n_ket; i++) empty = empty || hk[1 + i] == 0;
        if (empty) {
            if (info) info[2] = 1;
            return set_err(WFO_ERR_EMPTY_STATE, "overlaps: a state keeps no excitation above the threshold");
        }
    }
    if (kb == 0 || kk == 0) {
        *zero = true;
        return WFO_OK;
    }
    // ---- phase 2: coefficient matrices, then the fused stages on this rank's shard
    const int base = kb / world, rem = kb % world;
    const int k_begin = rank * base + (rank < rem ? rank : rem);
    const int k_end = k_begin + base + (rank < rem ? 1 : 0);
    size_t total = align_up((size_t)n_bra * kb * 8, 256) + align_up((size_t)n_ket * kk * 8, 256);
    {
        bool want_t1 = false, may_t0 = false;
        WFO_TRY(resolve_tier(ctx, n, tier, &want_t1, &may_t0));
        const int ksh = k_end - k_begin;
        if (ksh > 0) {
            Part q;
            WFO_TRY(make_part(ctx, ksh, kk, n_ket < MAX_KET_BLOCK ? n_ket : MAX_KET_BLOCK, may_t0, want_t1, &q));
            Plan p;
            Sizer s2z;
            s2z.off = total;
            carve(s2z, p, n, v, ksh, kk, n_bra, q.ldg, q.n_parts, q.n_slabs, may_t0, want_t1);
            if (tier == WFO_TIER_CLOSED) s2z.off = total + closed_form_bytes(n, v, n_bra, n_ket);
            total = s2z.off;
        }
    }
    WFO_TRY(ensure_scratch(ctx, total));   // nothing of this call lives in the arena yet
    Arena ar2{ctx->scratch, ctx->scratch_bytes, 0};
    double *d_cb = ar2.take<double>((size_t)n_bra * kb), *d_ck = ar2.take<double>((size_t)n_ket * kk);
    emit_exc_kernel<<<cdiv(bra_cnt, 256), 256, 0, st>>>(w.fl_b, w.of_b, n, v, 0, w.exc_b, bra_pos0, bra_cnt);
    WFO_LAUNCH_CHECK(ctx);
    emit_exc_kernel<<<cdiv(count, 256), 256, 0, st>>>(w.fl_k, w.of_k, n, v, 1, w.exc_k, 0, count);
    WFO_LAUNCH_CHECK(ctx);
    emit_coeff_kernel<<<dim3(cdiv(bra_cnt, 256), n_bra), 256, 0, st>>>(d_bci, w.fl_b, w.of_b, n, v, ci_thresh, strict, 0,
                                                                      semantics == 1, kb, d_cb, bra_pos0, bra_cnt, bra_stride);
    WFO_LAUNCH_CHECK(ctx);
    emit_coeff_kernel<<<dim3(cdiv(count, 256), n_ket), 256, 0, st>>>(d_kci, w.fl_k, w.of_k, n, v, ci_thresh, strict, 1,
                                                                    semantics == 1, kk, d_ck, 0, count, (long long)count);
    WFO_LAUNCH_CHECK(ctx);
    char *before = ctx->scratch;
    WFO_TRY(state_overlaps_impl(ctx, ar2.off, w.smo, n_mo, n, w.exc_b, kb, d_cb, n_bra, w.exc_k, kk, d_ck, n_ket, sigma,
                                tier, k_begin, k_end, w.out, st));
    if (ctx->scratch != before) return set_err(WFO_ERR_CUDA, "internal: arena moved during a host call");
    if (info) info[3] = ctx->last_tier;
    return WFO_OK;
}

static int check_ci_args(const char *who, int ao_side, int n_mo, int n_occ, int n_bra, int n_ket, int semantics,
                         int rank, int world) {
    if (ao_side != WFO_AO_GIVEN && ao_side != WFO_AO_BRA && ao_side != WFO_AO_KET)
        return set_err(WFO_ERR_ARG, "ao_side must be WFO_AO_GIVEN, WFO_AO_BRA or WFO_AO_KET");
    if (n_mo < 1 || n_occ < 1 || n_occ > n_mo || n_bra < 1 || n_ket < 1 || world < 1 || rank < 0 || rank >= world)
        return set_err(WFO_ERR_ARG, "bad sizes");
    if (semantics != 0 && semantics != 1) return set_err(WFO_ERR_ARG, "semantics must be 0 or 1");
    (void)who;
    return WFO_OK;
}

extern "C" {

int wfo_overlaps_ci_host(wfo_ctx *ctx, const double *bra_mos, const double *ket_mos, const double *c_inv, int ao_side,
                         int n_mo, int n_occ, const double *bra_ci, int n_bra, const double *ket_ci, int n_ket,
                         double ci_thresh, int semantics, int tier, int rank, int world, double *out,
                         int32_t *info) {
    if (!ctx) return set_err(WFO_ERR_ARG, "ctx is NULL");
    if (!bra_mos || !ket_mos || !bra_ci || !ket_ci || !out)
        return set_err(WFO_ERR_ARG, "overlaps_ci_host: NULL pointer");
    WFO_TRY(check_ci_args("overlaps_ci_host", ao_side, n_mo, n_occ, n_bra, n_ket, semantics, rank, world));
    if (ao_side == WFO_AO_GIVEN && !c_inv) return set_err(WFO_ERR_ARG, "overlaps_ci_host: c_inv is NULL");
    WFO_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream, s2 = ctx->stream2;
    const int n = n_occ, v = n_mo - n_occ, count = n * v;
    const size_t nsq = (size_t)n_mo * n_mo;
    if (info) { info[0] = info[1] = 0; info[2] = 0; info[3] = -1; }
    if (count == 0) {   // no virtual orbitals: no excitations, all sums are empty
        memset(out, 0, (size_t)n_bra * n_ket * sizeof(double));
        return WFO_OK;
    }
    CiWork w;
    Sizer sz;
    sz.take<double>(nsq); sz.take<double>(nsq); sz.take<double>(nsq);
    sz.take<double>((size_t)n_bra * count); sz.take<double>((size_t)n_ket * count);
    carve_ciwork(sz, w, nsq, count, n_bra, n_ket);
    WFO_TRY(ensure_resident(ctx, sz.off));
    Arena ar{ctx->resident, ctx->resident_bytes, 0};
    double *d_bra = ar.take<double>(nsq), *d_ket = ar.take<double>(nsq), *d_inv = ar.take<double>(nsq);
    double *d_bci = ar.take<double>((size_t)n_bra * count), *d_kci = ar.take<double>((size_t)n_ket * count);
    carve_ciwork(ar, w, nsq, count, n_bra, n_ket);
    // main stream: raw CI tensors up (then the tables); side stream, concurrently: MO matrices up (then the
    // inverse behind S_AO, S_AO, S_MO).  With several ranks and the singlet semantics each rank uploads and
    // thresholds only ITS slice of the bra (from, to) space -- any partition of the bra excitations gives
    // partial matrices that add up -- so the bra-side copies and table work shrink with the rank count.
    const bool slice = world > 1 && semantics == 0;
    // With a communicator (wfo_comm_init) the ranks also share the ket side: every rank uploads 1/world
    // of the ket CI tensor and pushes it to its peers over NVLink, and the partial matrices are summed
    // on the device, so `out` is the full matrix on every rank.
    Comm *cm = (ctx->comm && ctx->comm->world == world && world > 1) ? ctx->comm : nullptr;
    if (ctx->comm && !cm) return set_err(WFO_ERR_ARG, "overlaps_ci_host: world does not match the communicator");
    if (cm && (size_t)n_bra * n_ket * 8 > COMM_RED_CAP) return set_err(WFO_ERR_ARG, "overlaps_ci_host: result too large for the collective");
    int gat_epoch = 0, red_epoch = 0;
    if (cm) comm_next(cm, &gat_epoch, &red_epoch);
    const size_t kci_n = (size_t)n_ket * count;
    const bool gather = cm && kci_n * 8 <= cm->gat_cap;
    if (gather) d_kci = reinterpret_cast<double *>(cm->buf + comm_gat_off(world));
    int p_lo = 0, p_len = count;
    if (slice) {
        const int base = count / world, rem = count % world;
        p_lo = rank * base + (rank < rem ? rank : rem);
        p_len = base + (rank < rem ? 1 : 0);
        if (p_len > 0)
            WFO_CUDA(cudaMemcpy2DAsync(d_bci, (size_t)p_len * 8, bra_ci + p_lo, (size_t)count * 8, (size_t)p_len * 8,
                                       n_bra, cudaMemcpyHostToDevice, st));
    } else {
        WFO_CUDA(cudaMemcpyAsync(d_bci, bra_ci, (size_t)n_bra * count * 8, cudaMemcpyHostToDevice, st));
    }
    if (gather) {
        const size_t lo = kci_n * rank / world, hi = kci_n * (rank + 1) / world;
        if (hi > lo) WFO_CUDA(cudaMemcpyAsync(d_kci + lo, ket_ci + lo, (hi - lo) * 8, cudaMemcpyHostToDevice, st));
        WFO_TRY(comm_allgather(ctx, lo, hi, gat_epoch, st));
    } else {
        WFO_CUDA(cudaMemcpyAsync(d_kci, ket_ci, kci_n * 8, cudaMemcpyHostToDevice, st));
    }
    WFO_CUDA(cudaMemcpyAsync(d_bra, bra_mos, nsq * 8, cudaMemcpyHostToDevice, s2));
    WFO_CUDA(cudaMemcpyAsync(d_ket, ket_mos, nsq * 8, cudaMemcpyHostToDevice, s2));
    if (ao_side == WFO_AO_GIVEN) WFO_CUDA(cudaMemcpyAsync(d_inv, c_inv, nsq * 8, cudaMemcpyHostToDevice, s2));
    bool zero = false;
    if (slice)
        WFO_TRY(overlaps_ci_core(ctx, w, d_bra, d_ket, ao_side == WFO_AO_GIVEN ? d_inv : nullptr, ao_side, n_mo, n_occ, d_bci,
                                 n_bra, d_kci, n_ket, ci_thresh, semantics, tier, 0, 1, false, info, &zero, st, p_lo, p_len,
                                 (long long)p_len));
    else
        WFO_TRY(overlaps_ci_core(ctx, w, d_bra, d_ket, ao_side == WFO_AO_GIVEN ? d_inv : nullptr, ao_side, n_mo, n_occ, d_bci,
                                 n_bra, d_kci, n_ket, ci_thresh, semantics, tier, rank, world, false, info, &zero, st));
    if (cm) {   // every rank takes part, also one whose slice kept no excitation
        if (zero) WFO_CUDA(cudaMemsetAsync(w.out, 0, (size_t)n_bra * n_ket * 8, st));
        WFO_TRY(comm_allreduce(ctx, w.out, (size_t)n_bra * n_ket, red_epoch, st));
    } else if (zero) {
        memset(out, 0, (size_t)n_bra * n_ket * sizeof(double));
        return WFO_OK;
    }
    WFO_CUDA(cudaMemcpyAsync(out, w.out, (size_t)n_bra * n_ket * 8, cudaMemcpyDeviceToHost, st));
    WFO_CUDA(cudaStreamSynchronize(st));
    if (cm) WFO_TRY(comm_check(ctx));
    return WFO_OK;
}

/* The same from DEVICE buffers (bra_mos, ket_mos, c_inv or NULL, bra_ci, ket_ci, out all on the device):
 * for callers whose orbitals and CI vectors already live in HBM.  Runs on `stream` plus the context's side
 * stream; synchronises `stream` once (the number of surviving excitations sizes the second phase) and
 * returns with the remaining work queued.  *zero_out is set when no excitation survives (out is zeroed). */
int wfo_overlaps_ci_dev(wfo_ctx *ctx, const double *bra_mos, const double *ket_mos, const double *c_inv, int ao_side,
                        int n_mo, int n_occ, const double *bra_ci, int n_bra, const double *ket_ci, int n_ket,
                        double ci_thresh, int semantics, int tier, int rank, int world, double *out, int32_t *info,
                        void *stream) {
    if (!ctx) return set_err(WFO_ERR_ARG, "ctx is NULL");
    if (!bra_mos || !ket_mos || !bra_ci || !ket_ci || !out) return set_err(WFO_ERR_ARG, "overlaps_ci_dev: NULL pointer");
    WFO_TRY(check_ci_args("overlaps_ci_dev", ao_side, n_mo, n_occ, n_bra, n_ket, semantics, rank, world));
    if (ao_side == WFO_AO_GIVEN && !c_inv) return set_err(WFO_ERR_ARG, "overlaps_ci_dev: c_inv is NULL");
    WFO_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int n = n_occ, v = n_mo - n_occ, count = n * v;
    const size_t nsq = (size_t)n_mo * n_mo, nout = (size_t)n_bra * n_ket;
    if (info) { info[0] = info[1] = 0; info[2] = 0; info[3] = -1; }
    if (count == 0) {
        WFO_CUDA(cudaMemsetAsync(out, 0, nout * sizeof(double), st));
        return WFO_OK;
    }
    CiWork w;
    Sizer sz;
    carve_ciwork(sz, w, nsq, count, n_bra, n_ket);
    WFO_TRY(ensure_resident(ctx, sz.off));
    Arena ar{ctx->resident, ctx->resident_bytes, 0};
    carve_ciwork(ar, w, nsq, count, n_bra, n_ket);
    bool zero = false;
    WFO_TRY(overlaps_ci_core(ctx, w, bra_mos, ket_mos, ao_side == WFO_AO_GIVEN ? c_inv : nullptr, ao_side, n_mo, n_occ,
                             bra_ci, n_bra, ket_ci, n_ket, ci_thresh, semantics, tier, rank, world, true, info, &zero, st));
    if (zero) WFO_CUDA(cudaMemsetAsync(out, 0, nout * sizeof(double), st));
    else WFO_CUDA(cudaMemcpyAsync(out, w.out, nout * sizeof(double), cudaMemcpyDeviceToDevice, st));
    return WFO_OK;
}

/* Trajectory driver: n_cycles geometries, the state overlaps of n_pairs (bra cycle, ket
 * cycle) pairs.  Every cycle's MO matrix and CI tensor is uploaded ONCE and stays resident, the inverse
 * behind S_AO is formed once per cycle that needs it; per pair only S_AO, S_MO, the tables and the fused
 * stages run.  upstream: a loop over pywfo.main.overlaps_cache / overlaps calls (pysisyphus' use of it). */
int wfo_trajectory_ci_host(wfo_ctx *ctx, const double *mos, const double *ci, int n_cycles, int n_mo, int n_occ,
                           int n_states, double ci_thresh, int semantics, int tier, int ao_side,
                           const int32_t *pairs, int n_pairs, int rank, int world, double *out, int32_t *info) {
    if (!ctx) return set_err(WFO_ERR_ARG, "ctx is NULL");
    if (!mos || !ci || !pairs || !out) return set_err(WFO_ERR_ARG, "trajectory: NULL pointer");
    if (n_cycles < 1 || n_pairs < 0) return set_err(WFO_ERR_ARG, "trajectory: bad sizes");
    if (ao_side != WFO_AO_BRA && ao_side != WFO_AO_KET)
        return set_err(WFO_ERR_ARG, "trajectory: ao_side must be WFO_AO_BRA or WFO_AO_KET");
    WFO_TRY(check_ci_args("trajectory", ao_side, n_mo, n_occ, n_states, n_states, semantics, rank, world));
    for (int p = 0; p < n_pairs; p++)
        if (pairs[2 * p] < 0 || pairs[2 * p] >= n_cycles || pairs[2 * p + 1] < 0 || pairs[2 * p + 1] >= n_cycles)
            return set_err(WFO_ERR_ARG, "trajectory: pair %lld names a cycle outside 0..n_cycles-1", p);
    WFO_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream, s2 = ctx->stream2;
    const int n = n_occ, v = n_mo - n_occ, count = n * v;
    const size_t nsq = (size_t)n_mo * n_mo, ncc = (size_t)n_states * count, nout = (size_t)n_states * n_states;
    if (info) { info[0] = info[1] = 0; info[2] = 0; info[3] = -1; }
    if (count == 0 || n_pairs == 0) {
        memset(out, 0, (size_t)n_pairs * nout * sizeof(double));
        return WFO_OK;
    }
    // Pairs of a trajectory are independent and, at the sizes of real trajectories (cytosine: 137 MOs, 2 states),
    // each is a chain of ~40 small launches with two host synchronisations: launch latency, not the GPU, sets the
    // pace.  So small problems run on several LANES -- worker contexts on the same device with their own streams
    // and arenas, one host thread each -- whose kernels overlap on the GPU; the cycles (MOs, CI, inverses) are
    // shared and resident.  Large problems (one pair fills the GPU) keep one lane.
    int n_lanes = 1;
    if ((long long)n_states * count <= 200000 && n_mo <= 512) n_lanes = n_pairs < TRAJ_MAX_LANES ? n_pairs : TRAJ_MAX_LANES;
    {
        const char *e = getenv("WFO_TRAJ_LANES");
        if (e && atoi(e) >= 1) n_lanes = atoi(e) < TRAJ_MAX_LANES ? atoi(e) : TRAJ_MAX_LANES;
        if (n_lanes > n_pairs) n_lanes = n_pairs;
    }
    for (int l = 1; l < n_lanes; l++)
        if (!ctx->lanes[l]) WFO_TRY(wfo_create(ctx->device, &ctx->lanes[l]));
    ctx->lanes[0] = ctx;
    Sizer sz;
    sz.take<double>(nsq * n_cycles); sz.take<double>(ncc * n_cycles); sz.take<double>(nsq * n_cycles);
    sz.take<double>(nout * n_pairs);
    const size_t shared_bytes = sz.off;
    {
        CiWork w0;
        carve_ciwork(sz, w0, nsq, count, n_states, n_states);
    }
    const size_t work_bytes = sz.off - shared_bytes;
    WFO_TRY(ensure_resident(ctx, sz.off));
    for (int l = 1; l < n_lanes; l++) WFO_TRY(ensure_resident(ctx->lanes[l], work_bytes));
    Arena ar{ctx->resident, ctx->resident_bytes, 0};
    double *d_mos = ar.take<double>(nsq * n_cycles), *d_ci = ar.take<double>(ncc * n_cycles);
    double *d_inv = ar.take<double>(nsq * n_cycles), *d_outs = ar.take<double>(nout * n_pairs);
    std::vector<CiWork> works(n_lanes);
    carve_ciwork(ar, works[0], nsq, count, n_states, n_states);
    for (int l = 1; l < n_lanes; l++) {
        Arena al{ctx->lanes[l]->resident, ctx->lanes[l]->resident_bytes, 0};
        carve_ciwork(al, works[l], nsq, count, n_states, n_states);
    }
    WFO_CUDA(cudaMemcpyAsync(d_mos, mos, nsq * n_cycles * 8, cudaMemcpyHostToDevice, st));
    WFO_CUDA(cudaMemcpyAsync(d_ci, ci, ncc * n_cycles * 8, cudaMemcpyHostToDevice, st));
    WFO_CUDA(cudaStreamSynchronize(st));   // the lanes read the cycles from their own streams
    // the cycles whose MO matrix defines an S_AO, in order of first use
    std::vector<int> inv_cycles;
    {
        std::vector<char> seen(n_cycles, 0);
        for (int p = 0; p < n_pairs; p++) {
            const int c = ao_side == WFO_AO_BRA ? pairs[2 * p] : pairs[2 * p + 1];
            if (!seen[c]) { seen[c] = 1; inv_cycles.push_back(c); }
        }
    }
    struct LaneResult { int rc; char msg[512]; int32_t info[4]; };
    std::vector<LaneResult> res(n_lanes);
    auto run_lanes = [&](auto &&body) {   // body(lane) -> rc, on the lane's own host thread
        auto wrapped = [&](int l) {
            cudaSetDevice(ctx->device);
            res[l].rc = body(l);
            if (res[l].rc != WFO_OK) snprintf(res[l].msg, sizeof(res[l].msg), "%s", g_err);
        };
        std::vector<std::thread> th;
        for (int l = 1; l < n_lanes; l++) th.emplace_back(wrapped, l);
        wrapped(0);
        for (auto &t : th) t.join();
        for (int l = 0; l < n_lanes; l++)
            if (res[l].rc != WFO_OK) {
                snprintf(g_err, sizeof(g_err), "%s", res[l].msg);
                return res[l].rc;
            }
        return (int)WFO_OK;
    };
    // phase 1: one inverse per cycle, dealt over the lanes
    WFO_TRY(run_lanes([&](int l) -> int {
        wfo_ctx *lc = ctx->lanes[l];
        for (size_t i = l; i < inv_cycles.size(); i += n_lanes) {
            const int c = inv_cycles[i];
            WFO_TRY(wfo_inverse_dev(lc, d_mos + nsq * c, n_mo, d_inv + nsq * c, works[l].work, lc->stream));
        }
        WFO_CUDA(cudaStreamSynchronize(lc->stream));
        return WFO_OK;
    }));
    // phase 2: the pairs, dealt over the lanes
    for (int l = 0; l < n_lanes; l++) { res[l].info[0] = res[l].info[1] = 0; res[l].info[2] = 0; res[l].info[3] = -1; }
    const int rc2 = run_lanes([&](int l) -> int {
        wfo_ctx *lc = ctx->lanes[l];
        cudaStream_t ls = lc->stream;
        for (int p = l; p < n_pairs; p += n_lanes) {
            const int cb = pairs[2 * p], ck = pairs[2 * p + 1];
            const int ci_c = ao_side == WFO_AO_BRA ? cb : ck;
            bool zero = false;
            int32_t pinfo[4];
            WFO_TRY(overlaps_ci_core(lc, works[l], d_mos + nsq * cb, d_mos + nsq * ck, d_inv + nsq * ci_c, WFO_AO_GIVEN, n_mo,
                                     n_occ, d_ci + ncc * cb, n_states, d_ci + ncc * ck, n_states, ci_thresh, semantics, tier,
                                     rank, world, true, pinfo, &zero, ls));
            int32_t *li = res[l].info;
            li[0] = pinfo[0] > li[0] ? pinfo[0] : li[0];
            li[1] = pinfo[1] > li[1] ? pinfo[1] : li[1];
            li[3] = pinfo[3];
            if (zero) WFO_CUDA(cudaMemsetAsync(d_outs + nout * p, 0, nout * 8, ls));
            else WFO_CUDA(cudaMemcpyAsync(d_outs + nout * p, works[l].out, nout * 8, cudaMemcpyDeviceToDevice, ls));
        }
        WFO_CUDA(cudaStreamSynchronize(ls));
        return WFO_OK;
    });
    for (int l = 1; l < n_lanes; l++) {   // bookkeeping of the worker contexts belongs to the caller's context
        ctx->launches += ctx->lanes[l]->launches;
        ctx->lanes[l]->launches = 0;
    }
    if (rc2 != WFO_OK) {
        if (info && rc2 == WFO_ERR_EMPTY_STATE) info[2] = 1;
        return rc2;
    }
    if (info)
        for (int l = 0; l < n_lanes; l++) {
            info[0] = res[l].info[0] > info[0] ? res[l].info[0] : info[0];
            info[1] = res[l].info[1] > info[1] ? res[l].info[1] : info[1];
            if (res[l].info[3] >= 0) info[3] = res[l].info[3];
        }
    ctx->last_lanes = n_lanes;
    WFO_CUDA(cudaMemcpyAsync(out, d_outs, nout * n_pairs * 8, cudaMemcpyDeviceToHost, st));
    WFO_CUDA(cudaStreamSynchronize(st));
    return WFO_OK;
}

#ifdef WFO_WPM_CLOCKS
int wfo_debug_wpm_clocks(long long *out8, int reset) {
    if (out8) cudaMemcpyFromSymbol(out8, g_wpm_clk, sizeof(long long) * 8);
    if (reset) { long long z[8] = {0}; cudaMemcpyToSymbol(g_wpm_clk, z, sizeof(z)); }
    return 0;
}
#endif

#ifdef WFO_LL_CLOCKS
extern "C" int wfo_debug_ll_clocks(long long *out8, int reset) {
    if (out8) cudaMemcpyFromSymbol(out8, g_ll_clk, sizeof(long long) * 8);
    if (reset) { long long z[8] = {0}; cudaMemcpyToSymbol(g_ll_clk, z, sizeof(z)); }
    return 0;
}
#endif

#ifdef WFO_T1_CLOCKS
int wfo_debug_t1_clocks(long long *out, int n_ctas) {
    return (int)cudaMemcpyFromSymbol(out, g_t1_clk, sizeof(long long) * 4 * (size_t)n_ctas);
}
#endif

/* ------------------------------------------------------------ measurement */
int wfo_fp64_peak(wfo_ctx *ctx, int kind, int iters, double *tflops) {
    if (!ctx || !tflops) return set_err(WFO_ERR_ARG, "fp64_peak: NULL argument");
    WFO_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    WFO_TRY(ensure_scratch(ctx, 256));
    double *out = reinterpret_cast<double *>(ctx->scratch);
    const int blocks = ctx->sm_count * 8, threads = 256;
    double flops_per_thread_iter = kind == 0 ? 16 * 2.0 : (kind == 1 ? 8 * 512.0 / 32 : 4 * 2048.0 / 32);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        WFO_CUDA(cudaEventRecord(ctx->ev0, st));
        if (kind == 0) dfma_peak_kernel<<<blocks, threads, 0, st>>>(iters, out);
        else if (kind == 1) dmma884_peak_kernel<<<blocks, threads, 0, st>>>(iters, out);
        else if (kind == 2) dmma1688_peak_kernel<<<blocks, threads, 0, st>>>(iters, out);
        else return set_err(WFO_ERR_ARG, "fp64_peak: kind must be 0, 1 or 2");
        WFO_LAUNCH_CHECK(ctx);
        WFO_CUDA(cudaEventRecord(ctx->ev1, st));
        WFO_CUDA(cudaEventSynchronize(ctx->ev1));
        float ms = 0.f;
        WFO_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        if (rep > 0 && ms < best) best = ms;
    }
    *tflops = flops_per_thread_iter * iters * (double)blocks * threads / (best * 1e-3) / 1e12;
    return WFO_OK;
}

}  // extern "C"
