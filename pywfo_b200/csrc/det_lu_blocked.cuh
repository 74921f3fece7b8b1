// T0 (brute force), blocked kernel for 16 < n <= 128: one CTA per determinant pair.
//
// The gathered n x n minor of S_MO lives in shared memory (row-major, padded
// stride).  Right-looking LU with partial pivoting in 16-column panels, each
// factored as two 8-column sub-panels:
//   * sub-panel factorisation by ONE warp, register resident (lane = row,
//     up to 4 rows per lane, 8 columns each): the pivot search is a single
//     REDUX.MAX over a key that packs the high word of |a| with (slot, lane),
//     the pivot row is broadcast with warp shuffles, no block barrier inside;
//   * rows are physically compacted after every sub-panel (pivot rows to the
//     top, displaced rows into the vacated positions), an original-row-id
//     column travels with the data and gives the permutation sign at the end;
//   * triangular solve for the U block row: one thread per trailing column;
//   * trailing update A22 -= L21 U12 on the FP64 tensor cores (DMMA.8x8x4),
//     rank 16, operands read from shared memory as fragments, each warp
//     working on 16x32 macro tiles (8 accumulators tiles in registers).
// The determinant is the signed product of the pivots -- the quantity
// np.linalg.det returns for S_MO[rows][:, cols] (pywfo/main.py:583-584).
#pragma once
#include "common.cuh"

namespace wfo {

struct BlkShared {
    double det;          // product of pivots (written by warp 0)
    int singular;
    int aff_dst[16];     // row positions that receive another row (-1: unused)
    int aff_src[16];     // ... and the position the row comes from
    int vac[8];          // scratch: vacated positions
    double red[8];
};

__device__ __forceinline__ int blk_ld(int n) {   // leading dimension, == 8 (mod 16), >= n rounded to 8
    int ld = (n + 7) & ~7;
    if ((ld & 15) == 0) ld += 8;
    return ld;
}

// ---------------------------------------------------------------------------------
// sub-panel factorisation: rows [j0, n), columns [j0, j0+w), w <= 8.  Warp 0 only.
// On exit the panel columns are written back at the rows' FINAL positions, the
// affected-row list is published in sh, det is multiplied by the w pivots.
template <int SLOTS>
__device__ __forceinline__ void factor_subpanel(double *A, int *rowid, int LD, int n, int j0, int w,
                                                BlkShared *sh, int lane, double &det, bool &singular) {
    const unsigned FULL = 0xffffffffu;
    double a[SLOTS][8];
    bool done[SLOTS];
    int pstep[SLOTS];
#pragma unroll
    for (int s = 0; s < SLOTS; s++) {
        const int r = j0 + lane + 32 * s;
        done[s] = !(r < n);
        pstep[s] = -1;
        if (r < n) {
            const double2 *src = reinterpret_cast<const double2 *>(A + (size_t)r * LD + j0);
#pragma unroll
            for (int c = 0; c < 4; c++) {
                double2 v = src[c];
                a[s][2 * c] = v.x;
                a[s][2 * c + 1] = v.y;
            }
        } else {
#pragma unroll
            for (int c = 0; c < 8; c++) a[s][c] = 0.0;
        }
    }
    int piv_pos[8];
#pragma unroll
    for (int jj = 0; jj < 8; jj++) {
        piv_pos[jj] = j0 + jj;
        if (jj < w && !singular) {
            // ---- pivot search: one REDUX over (|a| high word, slot, lane)
            unsigned key = 0;
#pragma unroll
            for (int s = 0; s < SLOTS; s++) {
                unsigned k = (hi_abs(a[s][jj]) & 0xffffff00u) | 0x80u | (unsigned)(s << 5) | (unsigned)lane;
                if (!done[s]) key = max(key, k);
            }
            const unsigned kmax = __reduce_max_sync(FULL, key);
            const int wl = kmax & 31, ws = (kmax >> 5) & 3;
            // ---- broadcast the pivot row (ws is warp-uniform: no divergence)
            double u[8];
#pragma unroll
            for (int s = 0; s < SLOTS; s++) {
                if (ws == s) {
#pragma unroll
                    for (int c = jj; c < 8; c++) u[c] = __shfl_sync(FULL, a[s][c], wl);
                }
            }
            const double piv = u[jj];
            if (piv == 0.0) {
                singular = true;
            } else {
                det *= piv;
                const double rcp = 1.0 / piv;
                piv_pos[jj] = j0 + wl + 32 * ws;
#pragma unroll
                for (int s = 0; s < SLOTS; s++) {
                    if (lane == wl && ws == s) { done[s] = true; pstep[s] = jj; }
                    if (!done[s]) {
                        const double l = a[s][jj] * rcp;
                        a[s][jj] = l;
#pragma unroll
                        for (int c = jj + 1; c < 8; c++) a[s][c] = fma(-l, u[c], a[s][c]);
                    }
                }
            }
        }
    }
    if (singular) return;   // warp-uniform: nothing below is needed, det = 0
    // ---- final positions: pivot row jj -> j0+jj; displaced rows -> vacated positions
    int newpos[SLOTS];
    unsigned vac_before = 0;
#pragma unroll
    for (int s = 0; s < SLOTS; s++) {
        const int r = j0 + lane + 32 * s;
        newpos[s] = (pstep[s] >= 0) ? j0 + pstep[s] : r;
        const bool vacated = pstep[s] >= 0 && r >= j0 + w;
        const unsigned bal = __ballot_sync(FULL, vacated);
        if (vacated) sh->vac[vac_before + __popc(bal & ((1u << lane) - 1))] = r;
        vac_before += __popc(bal);
    }
    const bool displaced = (lane < w) && pstep[0] < 0 && (j0 + lane < n);
    const unsigned dbal = __ballot_sync(FULL, displaced);
    const int drank = __popc(dbal & ((1u << lane) - 1));
    __syncwarp();
    if (displaced) newpos[0] = sh->vac[drank];
    // ---- affected list: entries 0..7 pivots, 8..15 displaced
    if (lane < 16) { sh->aff_dst[lane] = -1; sh->aff_src[lane] = -1; }
    __syncwarp();
#pragma unroll
    for (int s = 0; s < SLOTS; s++) {
        const int r = j0 + lane + 32 * s;
        if (pstep[s] >= 0 && newpos[s] != r) { sh->aff_dst[pstep[s]] = newpos[s]; sh->aff_src[pstep[s]] = r; }
    }
    if (displaced) { sh->aff_dst[8 + drank] = newpos[0]; sh->aff_src[8 + drank] = j0 + lane; }
    // ---- write the panel columns back at the final positions
#pragma unroll
    for (int s = 0; s < SLOTS; s++) {
        const int r = j0 + lane + 32 * s;
        if (r < n) {
            double2 *dst = reinterpret_cast<double2 *>(A + (size_t)newpos[s] * LD + j0);
#pragma unroll
            for (int c = 0; c < 4; c++) dst[c] = make_double2(a[s][2 * c], a[s][2 * c + 1]);
        }
    }
    (void)piv_pos;
    (void)rowid;
}

// Apply the published row moves to the column ranges [c0a,c1a) and [c0b,c1b) and to rowid.
// All threads; one warp per moved row, lanes over columns; one block barrier between the
// reads and the writes.
template <int WARPS>
__device__ __forceinline__ void permute_rows(double *A, int *rowid, int LD, const BlkShared *sh, int c0a, int c1a,
                                             int c0b, int c1b) {
    constexpr int EPW = 16 / WARPS;          // entries per warp
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wa = c1a - c0a, wt = wa + (c1b - c0b);
    double tmp[EPW][4];
    int rid[EPW];
#pragma unroll
    for (int e = 0; e < EPW; e++) {
        const int src = sh->aff_src[warp * EPW + e];
        rid[e] = -1;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int cc = lane + 32 * k;
            tmp[e][k] = 0.0;
            if (src >= 0 && cc < wt) tmp[e][k] = A[(size_t)src * LD + (cc < wa ? c0a + cc : c0b + (cc - wa))];
        }
        if (src >= 0) rid[e] = rowid[src];
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < EPW; e++) {
        const int dst = sh->aff_dst[warp * EPW + e];
        if (dst >= 0) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int cc = lane + 32 * k;
                if (cc < wt) A[(size_t)dst * LD + (cc < wa ? c0a + cc : c0b + (cc - wa))] = tmp[e][k];
            }
            if (lane == 0) rowid[dst] = rid[e];
        }
    }
}

// U block row: rows [r0, r0+w) are pivot rows; for every trailing column solve with the
// unit-lower block L[r0.., r0..]; if r0 > p0 first subtract L[r0.., p0..r0) * U[p0..r0) (rank-8
// contribution of the first sub-panel of this panel).  One thread per column.
template <int THREADS>
__device__ __forceinline__ void trsm_block_row(double *A, int LD, int n, int p0, int r0, int w, int c_begin) {
    for (int c = c_begin + threadIdx.x; c < n; c += THREADS) {
        double x[8];
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = (i < w) ? A[(size_t)(r0 + i) * LD + c] : 0.0;
        for (int q = p0; q < r0; q++) {
            const double uq = A[(size_t)q * LD + c];
#pragma unroll
            for (int i = 0; i < 8; i++)
                if (i < w) x[i] = fma(-A[(size_t)(r0 + i) * LD + q], uq, x[i]);
        }
#pragma unroll
        for (int q = 0; q < 8; q++) {
            if (q < w) {
#pragma unroll
                for (int i = q + 1; i < 8; i++)
                    if (i < w) x[i] = fma(-A[(size_t)(r0 + i) * LD + r0 + q], x[q], x[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < 8; i++)
            if (i < w) A[(size_t)(r0 + i) * LD + c] = x[i];
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
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < n; i += THREADS) rowid[i] = i;
    if (tid == 0) sh->singular = 0;
    double det = 1.0;
    bool singular = false;
    __syncthreads();
    for (int j = 0; j < n; j += 16) {
        const int w1 = min(8, n - j);
        // ---- first sub-panel
        if (warp == 0) {
            factor_subpanel<SLOTS>(A, rowid, LD, n, j, w1, sh, lane, det, singular);
            if (lane == 0 && singular) sh->singular = 1;
        }
        __syncthreads();
        if (sh->singular) break;
        if (j + w1 >= n) { permute_rows<WARPS>(A, rowid, LD, sh, 0, 0, 0, 0); __syncthreads(); break; }
        permute_rows<WARPS>(A, rowid, LD, sh, j + 8, n, 0, 0);
        __syncthreads();
        trsm_block_row<THREADS>(A, LD, n, j, j, 8, j + 8);
        __syncthreads();
        const int w2 = min(8, n - j - 8);
        // rank-8 update of the second sub-panel's columns only
        dmma_update<WARPS>(A, LD, n, j + 8, j + 8, j + 8 + w2, j, 8);
        __syncthreads();
        // ---- second sub-panel
        if (warp == 0) {
            factor_subpanel<SLOTS>(A, rowid, LD, n, j + 8, w2, sh, lane, det, singular);
            if (lane == 0 && singular) sh->singular = 1;
        }
        __syncthreads();
        if (sh->singular) break;
        if (j + 8 + w2 >= n) { permute_rows<WARPS>(A, rowid, LD, sh, 0, 0, 0, 0); __syncthreads(); break; }
        permute_rows<WARPS>(A, rowid, LD, sh, j, j + 8, j + 16, n);
        __syncthreads();
        trsm_block_row<THREADS>(A, LD, n, j, j + 8, 8, j + 16);
        __syncthreads();
        // ---- rank-16 trailing update on the tensor cores
        dmma_update<WARPS>(A, LD, n, j + 16, j + 16, n, j, 16);
        __syncthreads();
    }
    // ---- publish |det| from warp 0, sign from the row permutation (inversion count)
    if (tid == 0) sh->det = det;
    __syncthreads();
    if (sh->singular) return 0.0;
    int inv = 0;
    for (int i = tid; i < n; i += THREADS) {
        const int ri = rowid[i];
        for (int k = i + 1; k < n; k++) inv += (rowid[k] < ri);
    }
    const int par = __syncthreads_count(inv & 1) & 1;
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
inline size_t blk_smem_bytes(int n) {
    return (size_t)blk_rows(n) * blk_ld_h(n) * sizeof(double) + 3 * (size_t)n * sizeof(int) + 64;
}

__device__ __forceinline__ void gather_minor_ld(const double *__restrict__ S, int ld_s, const int *rows,
                                                const int *cols, int n, double *a, int ld) {
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
        int r = idx / n, c = idx - r * n;
        a[r * ld + c] = __ldg(S + (size_t)rows[r] * ld_s + cols[c]);
    }
}

template <int WARPS, int SLOTS>
__global__ void __launch_bounds__(WARPS * 32)
det_table_blk_kernel(const double *__restrict__ S, int ld_s, int n, const int32_t *__restrict__ rows, int kb,
                     const int32_t *__restrict__ cols, int kk, double *__restrict__ D) {
    extern __shared__ __align__(16) double sm[];
    const int LD = blk_ld_h(n);
    double *a = sm;
    int *ridx = (int *)(a + (size_t)blk_rows(n) * LD);
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
        gather_minor_ld(S, ld_s, ridx, cidx, n, a, LD);
        __syncthreads();
        double det = cta_lu_det_blocked<WARPS, SLOTS>(a, rowid, LD, n, &sh);
        if (threadIdx.x == 0) D[pair] = det;
        __syncthreads();
    }
}

template <int WARPS, int SLOTS>
__global__ void __launch_bounds__(WARPS * 32)
t0_fused_blk_kernel(const double *__restrict__ S, int ld_s, int n, const int32_t *__restrict__ rows, int ksh,
                    const int32_t *__restrict__ cols, int kk, const double *__restrict__ ckt, int ldg,
                    int chunk_len, int n_chunks, double *__restrict__ gpart) {
    extern __shared__ __align__(16) double sm[];
    const int LD = blk_ld_h(n);
    double *a = sm;
    int *ridx = (int *)(a + (size_t)blk_rows(n) * LD);
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
