// T0 (brute force), blocked kernel for n <= 128: one CTA per determinant pair.
//
// The gathered n x n minor of S_MO lives in shared memory (row-major, padded stride).
// Right-looking LU with partial pivoting in 16-column panels:
//   * panel factorisation by the first ceil(n/32) warps, register resident (one row per
//     lane, 16 columns per lane): per column ONE REDUX.MAX over a key that packs the high
//     word of |a| with (warp, lane); each warp's candidate publishes its row, one named
//     barrier, every lane reads the winning row; branch-free elimination; the reciprocal
//     of the pivot is computed speculatively by every lane for its own candidate while
//     the search is in flight;
//   * rows are physically compacted after every panel (pivot rows to the top, displaced
//     rows into the vacated positions); an original-row-id column travels with the data
//     and gives the permutation sign at the end;
//   * triangular solve for the 16-row U block: one thread per trailing column;
//   * trailing update A22 -= L21 U12 on the FP64 tensor cores (DMMA.8x8x4), rank 16,
//     operands read from shared memory as fragments, 16x32 macro tiles per warp.
// The determinant is the signed product of the pivots -- the quantity np.linalg.det
// returns for S_MO[rows][:, cols] (pywfo/main.py:583-584).
#pragma once
#include "common.cuh"

namespace wfo {

#ifdef WFO_PHASE_TIMING
// scratch-build instrumentation: cycles per phase, accumulated by thread 0 of block 0
__device__ long long g_phase[16];
#define PH_DECL long long ph_t = clock64();
#define PH_MARK(i)                                                        \
    do {                                                                  \
        if (threadIdx.x == 0 && blockIdx.x == 0) {                        \
            long long now_ = clock64();                                   \
            g_phase[i] += now_ - ph_t;                                    \
            ph_t = now_;                                                  \
        }                                                                 \
    } while (0)
#else
#define PH_DECL
#define PH_MARK(i)
#endif

struct BlkShared {
    double det;                  // product of pivots (written by thread 0)
    int singular;                // an exactly zero pivot was met
    unsigned key[2][4];          // per-warp pivot candidates, double buffered by column parity
    double cand[2][4][18];       // per-warp candidate rows (+ reciprocal of the candidate at [16])
    int piv_row[16];             // row chosen as pivot of step jj of the current panel
    int aff_dst[32];             // row positions that receive another row (-1: unused)
    int aff_src[32];             // ... and the position the row comes from
};

__device__ __forceinline__ void panel_bar(int nthreads) {   // named barrier 1: the panel warps only
    asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
}

// ---------------------------------------------------------------------------------
// panel factorisation: rows [j0, n), columns [j0, j0+16) of which w are inside the matrix,
// by the first PW warps of the CTA (warp p owns rows j0+32p..j0+32p+31, one row per lane).
// On exit the panel columns are written back at the rows' FINAL positions, the affected-row
// list is published in sh, det is multiplied by the w pivots, zero is set on a zero pivot.
template <int PW>
__device__ __forceinline__ void factor_panel16(double *A, int LD, int n, int j0, int w, BlkShared *sh, int warp,
                                               int lane, double &det, bool &zero) {
    const unsigned FULL = 0xffffffffu;
    constexpr int NT = PW * 32;
    const int r = j0 + 32 * warp + lane;
    const bool valid = r < n;
    double a[16];
    {
        const double2 *src = reinterpret_cast<const double2 *>(A + (size_t)(valid ? r : j0) * LD + j0);
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const double2 v = src[c];
            a[2 * c] = v.x;
            a[2 * c + 1] = v.y;
        }
    }
    bool done = !valid;
    int pstep = -1;
    const unsigned tag = 0x80u | (unsigned)(warp << 5) | (unsigned)lane;
#pragma unroll
    for (int jj = 0; jj < 16; jj++) {
        if (jj < w) {   // uniform
            const int b = jj & 1;
            const unsigned key = done ? 0u : ((hi_abs(a[jj]) & 0xffffff00u) | tag);
            const unsigned kw = __reduce_max_sync(FULL, key);
            const double rc_own = fast_rcp(done ? 1.0 : a[jj]);   // speculative, overlaps the search
            unsigned kmax = kw;
            double u[16], rcp;
            if (PW > 1) {
                if (key == kw && key != 0u) {            // this warp's candidate publishes its row
                    sh->key[b][warp] = kw;
#pragma unroll
                    for (int c = jj; c < 16; c++) sh->cand[b][warp][c] = a[c];
                    sh->cand[b][warp][16] = rc_own;
                }
                if (kw == 0u && lane == 0) sh->key[b][warp] = 0u;
                panel_bar(NT);
#pragma unroll
                for (int q = 0; q < PW; q++) kmax = max(kmax, sh->key[b][q]);
                const int ww = (kmax >> 5) & 3;
#pragma unroll
                for (int c = jj; c < 16; c++) u[c] = sh->cand[b][ww][c];
                rcp = sh->cand[b][ww][16];
            } else {
                const int wl = kmax & 31;
#pragma unroll
                for (int c = jj; c < 16; c++) u[c] = __shfl_sync(FULL, a[c], wl);
                rcp = __shfl_sync(FULL, rc_own, wl);
            }
            const double piv = u[jj];
            det *= piv;
            zero = zero || (piv == 0.0);
            const bool win = (key == kmax) && (key != 0u);
            if (win) sh->piv_row[jj] = r;
            pstep = win ? jj : pstep;
            done = done || win;
            // branch-free elimination: finished rows use a zero multiplier
            const double l = done ? 0.0 : a[jj] * rcp;
            a[jj] = done ? a[jj] : l;
#pragma unroll
            for (int c = jj + 1; c < 16; c++) a[c] = fma(-l, u[c], a[c]);
        }
    }
    if (PW > 1) panel_bar(NT); else __syncwarp();   // piv_row[] complete
    // ---- final positions: pivot of step jj -> j0+jj; displaced rows (top w rows that are not
    //      pivots, all in warp 0) -> the positions vacated by pivots from below, in step order
    int newpos = (pstep >= 0) ? j0 + pstep : r;
    if (warp == 0) {
        const bool displaced = (lane < w) && pstep < 0 && valid;
        const unsigned dbal = __ballot_sync(FULL, displaced);
        const int drank = __popc(dbal & ((1u << lane) - 1));
        if (displaced) {
            int seen = 0;
#pragma unroll
            for (int q = 0; q < 16; q++) {
                const int pr = sh->piv_row[q];
                if (q < w && pr >= j0 + w) {
                    if (seen == drank) newpos = pr;
                    seen++;
                }
            }
        }
        // affected list: entries 0..15 pivots, 16..31 displaced rows
        if (lane < 16) {
            const int pr = (lane < w) ? sh->piv_row[lane] : -1;
            const bool moved = (lane < w) && pr != j0 + lane;
            sh->aff_dst[lane] = moved ? j0 + lane : -1;
            sh->aff_src[lane] = moved ? pr : -1;
            sh->aff_dst[16 + lane] = -1;
            sh->aff_src[16 + lane] = -1;
        }
        __syncwarp();
        if (displaced) { sh->aff_dst[16 + drank] = newpos; sh->aff_src[16 + drank] = r; }
    }
    if (valid) {
        double2 *dst = reinterpret_cast<double2 *>(A + (size_t)newpos * LD + j0);
#pragma unroll
        for (int c = 0; c < 8; c++) dst[c] = make_double2(a[2 * c], a[2 * c + 1]);
    }
}

// Apply the published row moves to the columns [c0, c1) and to rowid.  All threads; one warp
// per moved row, lanes over columns (CH chunks of 32); one block barrier between the reads
// and the writes.
template <int WARPS, int CH>
__device__ __forceinline__ void permute_rows(double *A, int *rowid, int LD, const BlkShared *sh, int c0, int c1) {
    constexpr int EPW = 32 / WARPS;          // entries per warp
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wt = c1 - c0;
    double tmp[EPW][CH];
    int rid[EPW];
#pragma unroll
    for (int e = 0; e < EPW; e++) {
        const int src = sh->aff_src[warp * EPW + e];
        rid[e] = -1;
#pragma unroll
        for (int k = 0; k < CH; k++) {
            const int cc = lane + 32 * k;
            tmp[e][k] = 0.0;
            if (src >= 0 && cc < wt) tmp[e][k] = A[(size_t)src * LD + c0 + cc];
        }
        if (src >= 0) rid[e] = rowid[src];
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < EPW; e++) {
        const int dst = sh->aff_dst[warp * EPW + e];
        if (dst >= 0) {
#pragma unroll
            for (int k = 0; k < CH; k++) {
                const int cc = lane + 32 * k;
                if (cc < wt) A[(size_t)dst * LD + c0 + cc] = tmp[e][k];
            }
            if (lane == 0) rowid[dst] = rid[e];
        }
    }
}

// U block row: rows [r0, r0+16) are the pivot rows of the panel; for every trailing column
// solve with the unit-lower block L[r0.., r0..].  One thread per column.
template <int THREADS>
__device__ __forceinline__ void trsm_block_row16(double *A, int LD, int n, int r0) {
    for (int c = r0 + 16 + threadIdx.x; c < n; c += THREADS) {
        double x[16];
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] = A[(size_t)(r0 + i) * LD + c];
#pragma unroll
        for (int q = 0; q < 15; q++) {
#pragma unroll
            for (int i = q + 1; i < 16; i++) x[i] = fma(-A[(size_t)(r0 + i) * LD + r0 + q], x[q], x[i]);
        }
#pragma unroll
        for (int i = 1; i < 16; i++) A[(size_t)(r0 + i) * LD + c] = x[i];
    }
}

// Trailing update C[R0.., C0..C1) -= L[R0.., KC..KC+KW) * U[KC..KC+KW, C0..C1) with DMMA.
// KW in {8, 16}.  Rows/columns are processed in 8x8 tiles from R0/C0; the shared matrix has
// round_up(n, 8) rows and LD >= round_up(n, 8) columns, so edge tiles stay in bounds.
template <int WARPS>
__device__ __forceinline__ void dmma_update(double *A, int LD, int n, int R0, int C0, int C1, int KC, int KW) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int TR = (n - R0 + 7) >> 3, TC = (C1 - C0 + 7) >> 3;
    if (TR <= 0 || TC <= 0) return;
    const int MR = (TR + 1) >> 1, MC = (TC + 3) >> 2;   // macro tiles: 2 tile rows x 4 tile cols
    const int ksteps = KW >> 2;
    for (int m = warp; m < MR * MC; m += WARPS) {
        const int mr = m / MC, mc = m - mr * MC;
        const int tr0 = mr * 2, tc0 = mc * 4;
        const int nr = min(2, TR - tr0), nc = min(4, TC - tc0);
        double acc[2][4][2];
        // load C fragments
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (i < nr && j < nc) {
                    const double2 v = *reinterpret_cast<const double2 *>(
                        A + (size_t)(R0 + 8 * (tr0 + i) + g) * LD + C0 + 8 * (tc0 + j) + 2 * t);
                    acc[i][j][0] = v.x;
                    acc[i][j][1] = v.y;
                } else {
                    acc[i][j][0] = acc[i][j][1] = 0.0;
                }
            }
        for (int ks = 0; ks < ksteps; ks++) {
            double af[2], bf[4];
#pragma unroll
            for (int i = 0; i < 2; i++)
                af[i] = (i < nr) ? -A[(size_t)(R0 + 8 * (tr0 + i) + g) * LD + KC + 4 * ks + t] : 0.0;
#pragma unroll
            for (int j = 0; j < 4; j++)
                bf[j] = (j < nc) ? A[(size_t)(KC + 4 * ks + t) * LD + C0 + 8 * (tc0 + j) + g] : 0.0;
#pragma unroll
            for (int i = 0; i < 2; i++)
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if (i < nr && j < nc) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (i < nr && j < nc)
                    *reinterpret_cast<double2 *>(A + (size_t)(R0 + 8 * (tr0 + i) + g) * LD + C0 + 8 * (tc0 + j) + 2 * t) =
                        make_double2(acc[i][j][0], acc[i][j][1]);
    }
}

// Whole factorisation; all threads of the CTA call it.  A: round_up(n,8) x LD in shared
// memory (destroyed), rowid: n ints.  Returns the determinant to every thread.
template <int WARPS, int SLOTS>
__device__ __forceinline__ double cta_lu_det_blocked(double *A, int *rowid, int LD, int n, BlkShared *sh) {
    constexpr int THREADS = WARPS * 32;
    constexpr int CH = SLOTS;   // column chunks of 32 that cover n
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < n; i += THREADS) rowid[i] = i;
    double det = 1.0;
    bool zero = false;
    __syncthreads();
    PH_DECL
    for (int j = 0; j < n; j += 16) {
        const int w = min(16, n - j);
        if (warp < SLOTS) factor_panel16<SLOTS>(A, LD, n, j, w, sh, warp, lane, det, zero);
        __syncthreads();
        PH_MARK(1);
        const bool last = (j + 16 >= n);
        permute_rows<WARPS, CH>(A, rowid, LD, sh, last ? 0 : j + 16, last ? 0 : n);
        __syncthreads();
        PH_MARK(2);
        if (last) break;
        trsm_block_row16<THREADS>(A, LD, n, j);
        __syncthreads();
        PH_MARK(3);
        dmma_update<WARPS>(A, LD, n, j + 16, j + 16, n, j, 16);
        __syncthreads();
        PH_MARK(5);
    }
    // ---- |det| from thread 0 (a panel lane), sign from the row permutation (inversion count)
    if (tid == 0) { sh->det = det; sh->singular = zero ? 1 : 0; }
    int inv = 0;
    for (int i = tid; i < n; i += THREADS) {
        const int ri = rowid[i];
        for (int k = i + 1; k < n; k++) inv += (rowid[k] < ri);
    }
    const int par = __syncthreads_count(inv & 1) & 1;
    PH_MARK(6);
    if (sh->singular) return 0.0;
    const double d = sh->det;
    return par ? -d : d;
}

}  // namespace wfo

// ---------------------------------------------------------------------------------
// kernels: same two front ends as the generic kernel (det_lu_v1.cuh)
namespace wfo {

__host__ __device__ inline int blk_rows(int n) { return (n + 7) & ~7; }
__host__ __device__ inline int blk_ld_h(int n) {
    int ld = (n + 7) & ~7;
    if ((ld & 15) == 0) ld += 8;
    return ld;
}
inline size_t blk_mat_bytes(int n) { return (size_t)blk_rows(n) * blk_ld_h(n) * sizeof(double); }
inline size_t blk_smem_bytes(int n, bool global_scratch = false) {
    return (global_scratch ? 0 : blk_mat_bytes(n)) + 3 * (size_t)n * sizeof(int) + 64;
}

// gather the minor: warps walk rows, lanes walk columns (column indices cached in registers),
// four rows in flight per warp so that 16 independent L2 loads are outstanding per thread
__device__ __forceinline__ void gather_minor_ld(const double *__restrict__ S, int ld_s, const int *rows,
                                                const int *cols, int n, double *a, int ld) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int cc[4];
#pragma unroll
    for (int k = 0; k < 4; k++) cc[k] = (lane + 32 * k < n) ? cols[lane + 32 * k] : 0;
    for (int r0 = warp * 4; r0 < n; r0 += nw * 4) {
        double v[4][4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int r = r0 + i;
            const double *src = S + (size_t)rows[r < n ? r : 0] * ld_s;
#pragma unroll
            for (int k = 0; k < 4; k++) v[i][k] = (r < n && lane + 32 * k < n) ? __ldg(src + cc[k]) : 0.0;
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int r = r0 + i;
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (r < n && lane + 32 * k < n) a[(size_t)r * ld + lane + 32 * k] = v[i][k];
        }
    }
}

// GLOBAL = true keeps the working matrix in a per-CTA global scratch (L2 resident) instead of
// shared memory: more CTAs per SM for the larger n, at L2 instead of shared-memory latency.
template <int WARPS, int SLOTS, bool GLOBAL>
__global__ void __launch_bounds__(WARPS * 32)
det_table_blk_kernel(const double *__restrict__ S, int ld_s, int n, const int32_t *__restrict__ rows, int kb,
                     const int32_t *__restrict__ cols, int kk, double *__restrict__ D, double *gscratch) {
    extern __shared__ __align__(16) double sm[];
    const int LD = blk_ld_h(n);
    double *a = GLOBAL ? gscratch + (size_t)blockIdx.x * blk_rows(n) * LD : sm;
    int *ridx = GLOBAL ? (int *)sm : (int *)(sm + (size_t)blk_rows(n) * LD);
    int *cidx = ridx + n;
    int *rowid = cidx + n;
    __shared__ BlkShared sh;
    const long long npairs = (long long)kb * kk;
    for (long long pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
        const int k = (int)(pair / kk), l = (int)(pair - (long long)k * kk);
        for (int i = threadIdx.x; i < n; i += WARPS * 32) {
            ridx[i] = rows[(size_t)k * n + i];
            cidx[i] = cols[(size_t)l * n + i];
        }
        __syncthreads();
#ifdef WFO_PHASE_TIMING
        long long g0_ = clock64();
#endif
        gather_minor_ld(S, ld_s, ridx, cidx, n, a, LD);
        __syncthreads();
#ifdef WFO_PHASE_TIMING
        if (threadIdx.x == 0 && blockIdx.x == 0) { g_phase[0] += clock64() - g0_; g_phase[7] += 1; }
#endif
        double det = cta_lu_det_blocked<WARPS, SLOTS>(a, rowid, LD, n, &sh);
        if (threadIdx.x == 0) D[pair] = det;
        __syncthreads();
    }
}

template <int WARPS, int SLOTS, bool GLOBAL>
__global__ void __launch_bounds__(WARPS * 32)
t0_fused_blk_kernel(const double *__restrict__ S, int ld_s, int n, const int32_t *__restrict__ rows, int ksh,
                    const int32_t *__restrict__ cols, int kk, const double *__restrict__ ckt, int ldg,
                    int chunk_len, int n_chunks, double *__restrict__ gpart, double *gscratch) {
    extern __shared__ __align__(16) double sm[];
    const int LD = blk_ld_h(n);
    double *a = GLOBAL ? gscratch + (size_t)blockIdx.x * blk_rows(n) * LD : sm;
    int *ridx = GLOBAL ? (int *)sm : (int *)(sm + (size_t)blk_rows(n) * LD);
    int *cidx = ridx + n;
    int *rowid = cidx + n;
    __shared__ BlkShared sh;
    const long long nitems = (long long)ksh * n_chunks;
    for (long long item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int k = (int)(item / n_chunks), ch = (int)(item - (long long)k * n_chunks);
        for (int i = threadIdx.x; i < n; i += WARPS * 32) ridx[i] = rows[(size_t)k * n + i];
        double g = 0.0;
        const int l0 = ch * chunk_len, l1 = min(kk, l0 + chunk_len);
        for (int l = l0; l < l1; l++) {
            for (int i = threadIdx.x; i < n; i += WARPS * 32) cidx[i] = cols[(size_t)l * n + i];
            __syncthreads();
            gather_minor_ld(S, ld_s, ridx, cidx, n, a, LD);
            __syncthreads();
            const double det = cta_lu_det_blocked<WARPS, SLOTS>(a, rowid, LD, n, &sh);
            if (threadIdx.x < ldg) g += det * __ldg(ckt + (size_t)l * ldg + threadIdx.x);
            __syncthreads();
        }
        if (threadIdx.x < ldg) gpart[((size_t)ch * ksh + k) * ldg + threadIdx.x] = g;
    }
}

}  // namespace wfo
