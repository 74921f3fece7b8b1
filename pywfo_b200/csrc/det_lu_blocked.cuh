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
//   * no triangular solve: the panel lanes turn their multipliers L21 into
//     M = L21 L11^-1 in registers (120 FMA per row), so that the Schur complement is
//     A22 - M A12 with the RAW pivot rows A12 (U12 is never needed for a determinant);
//   * trailing update A22 -= M A12 on the FP64 tensor cores (DMMA.8x8x4), rank 16,
//     operands read from shared memory as fragments, 16x32 macro tiles per warp,
//     branch-free (clamped edge tiles, predicated stores).
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
    double L11[16][16];          // unit-lower factor of the pivot block (strict lower part used)
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
            if (win) {
                sh->piv_row[jj] = r;
#pragma unroll
                for (int c = 0; c < jj; c++) sh->L11[jj][c] = a[c];   // its multipliers: row jj of L11
            }
            pstep = win ? jj : pstep;
            done = done || win;
            // branch-free elimination: finished rows use a zero multiplier
            const double l = done ? 0.0 : a[jj] * rcp;
            a[jj] = done ? a[jj] : l;
#pragma unroll
            for (int c = jj + 1; c < 16; c++) a[c] = fma(-l, u[c], a[c]);
        }
    }
    if (PW > 1) panel_bar(NT); else __syncwarp();   // piv_row[], L11[][] complete
    // ---- M = L21 L11^-1 (row-wise back substitution x L11 = l) for the rows that stay in the
    //      trailing matrix; only needed when a trailing matrix exists (full panel)
    if (w == 16 && j0 + 16 < n && pstep < 0) {
        // column-oriented (axpy) form: once x[q] is final, it updates all x[c], c < q, independently
#pragma unroll
        for (int q = 15; q >= 1; q--) {
#pragma unroll
            for (int c = 0; c < q; c++) a[c] = fma(-a[q], sh->L11[q][c], a[c]);
        }
    }
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
        __syncwarp();
        // sign of this panel's row permutation = parity of the inversions among the moved rows
        const int md = sh->aff_dst[lane], ms = sh->aff_src[lane];
        int cnt = 0;
#pragma unroll 8
        for (int o = 0; o < 32; o++) {
            const int d2 = __shfl_sync(FULL, md, o), s2 = __shfl_sync(FULL, ms, o);
            cnt += (md >= 0 && d2 > md && s2 < ms && d2 >= 0) ? 1 : 0;
        }
        if (__popc(__ballot_sync(FULL, cnt & 1)) & 1) det = -det;
    }
    if (valid) {
        double2 *dst = reinterpret_cast<double2 *>(A + (size_t)newpos * LD + j0);
#pragma unroll
        for (int c = 0; c < 8; c++) dst[c] = make_double2(a[2 * c], a[2 * c + 1]);
    }
}

// Apply the published row moves to the columns [c0, c1).  All threads; one warp
// per moved row, lanes over columns (CH chunks of 32); one block barrier between the reads
// and the writes.
template <int WARPS, int CH>
__device__ __forceinline__ void permute_rows(double *A, int LD, const BlkShared *sh, int c0, int c1) {
    constexpr int EPW = 32 / WARPS;          // entries per warp
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wt = c1 - c0;
    double tmp[EPW][CH];
#pragma unroll
    for (int e = 0; e < EPW; e++) {
        const int src = sh->aff_src[warp * EPW + e];
#pragma unroll
        for (int k = 0; k < CH; k++) {
            const int cc = lane + 32 * k;
            tmp[e][k] = 0.0;
            if (src >= 0 && cc < wt) tmp[e][k] = A[(size_t)src * LD + c0 + cc];
        }
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

// Trailing update A[R0.., R0..) -= M[R0.., J..J+16) * A[J..J+16, R0..) with R0 = J+16, on the
// tensor cores.  8x8 tiles from R0; macro tile = 2 tile rows x 4 tile columns per warp.  Edge
// macro tiles clamp their tile indices (a duplicate tile is computed twice, stored once), so
// there is no branch between the loads and the 32 DMMAs.  The shared matrix has
// round_up(n, 8) rows and LD >= round_up(n, 8) columns, so every tile stays in bounds.
template <int WARPS>
__device__ __forceinline__ void dmma_update16(double *A, int LD, int n, int J) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int R0 = J + 16;
    const int TR = (n - R0 + 7) >> 3;
    if (TR <= 0) return;
    const int MR = (TR + 1) >> 1, MC = (TR + 3) >> 2;
    const double *Brow = A + (size_t)(J + t) * LD + R0 + g;      // B fragment base: row J+t, col R0+g
    for (int m = warp; m < MR * MC; m += WARPS) {
        const int mr = m / MC, mc = m - mr * MC;
        int ri[2], cj[4];
        bool vr[2], vc[4];
#pragma unroll
        for (int i = 0; i < 2; i++) { vr[i] = 2 * mr + i < TR; ri[i] = min(2 * mr + i, TR - 1); }
#pragma unroll
        for (int j = 0; j < 4; j++) { vc[j] = 4 * mc + j < TR; cj[j] = min(4 * mc + j, TR - 1); }
        double *crow[2];
#pragma unroll
        for (int i = 0; i < 2; i++) crow[i] = A + (size_t)(R0 + 8 * ri[i] + g) * LD;
        double acc[2][4][2];
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const double2 v = *reinterpret_cast<const double2 *>(crow[i] + R0 + 8 * cj[j] + 2 * t);
                acc[i][j][0] = v.x;
                acc[i][j][1] = v.y;
            }
#pragma unroll
        for (int ks = 0; ks < 4; ks++) {
            double af[2], bf[4];
#pragma unroll
            for (int i = 0; i < 2; i++) af[i] = -crow[i][J + 4 * ks + t];
#pragma unroll
            for (int j = 0; j < 4; j++) bf[j] = Brow[(size_t)(4 * ks) * LD + 8 * cj[j]];
#pragma unroll
            for (int i = 0; i < 2; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (vr[i] && vc[j])
                    *reinterpret_cast<double2 *>(crow[i] + R0 + 8 * cj[j] + 2 * t) =
                        make_double2(acc[i][j][0], acc[i][j][1]);
    }
}

// Whole factorisation; all threads of the CTA call it.  A: round_up(n,8) x LD in shared
// memory (destroyed).  Returns the determinant to every thread.
template <int WARPS, int SLOTS>
__device__ __forceinline__ double cta_lu_det_blocked(double *A, int LD, int n, BlkShared *sh) {
    constexpr int CH = SLOTS;   // column chunks of 32 that cover n
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double det = 1.0;           // signed: the panel warps fold the permutation sign in
    bool zero = false;
    PH_DECL
    for (int j = 0; j < n; j += 16) {
        const int w = min(16, n - j);
        if (warp < SLOTS) factor_panel16<SLOTS>(A, LD, n, j, w, sh, warp, lane, det, zero);
        if (j + 16 >= n) break;   // last panel: nothing left to permute or update
        __syncthreads();
        PH_MARK(1);
        permute_rows<WARPS, CH>(A, LD, sh, j + 16, n);
        __syncthreads();
        PH_MARK(2);
        dmma_update16<WARPS>(A, LD, n, j);
        __syncthreads();
        PH_MARK(5);
    }
    if (tid == 0) { sh->det = det; sh->singular = zero ? 1 : 0; }
    __syncthreads();
    PH_MARK(6);
    return sh->singular ? 0.0 : sh->det;
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
    return (global_scratch ? 0 : blk_mat_bytes(n)) + 2 * (size_t)n * sizeof(int) + 64;
}

// gather the minor: warps walk rows, lanes walk columns (column indices cached in registers),
// eight rows in flight per warp so that 32 independent L2 loads are outstanding per thread
__device__ __forceinline__ void gather_minor_ld(const double *__restrict__ S, int ld_s, const int *rows,
                                                const int *cols, int n, double *a, int ld) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int cc[4];
#pragma unroll
    for (int k = 0; k < 4; k++) cc[k] = (lane + 32 * k < n) ? cols[lane + 32 * k] : 0;
    constexpr int RF = 8;   // rows in flight per warp: 32 independent L2 loads per thread
    for (int r0 = warp * RF; r0 < n; r0 += nw * RF) {
        double v[RF][4];
#pragma unroll
        for (int i = 0; i < RF; i++) {
            const int r = r0 + i;
            const double *src = S + (size_t)rows[r < n ? r : 0] * ld_s;
#pragma unroll
            for (int k = 0; k < 4; k++) v[i][k] = (r < n && lane + 32 * k < n) ? __ldg(src + cc[k]) : 0.0;
        }
#pragma unroll
        for (int i = 0; i < RF; i++) {
            const int r = r0 + i;
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (r < n && lane + 32 * k < n) a[(size_t)r * ld + lane + 32 * k] = v[i][k];
        }
    }
}

// GLOBAL = true keeps the working matrix in a per-CTA global scratch (L2 resident) instead of
// shared memory: more CTAs per SM for the larger n, at L2 instead of shared-memory latency.
#ifndef WFO_T0_MINB64
#define WFO_T0_MINB64 5
#endif
// minimum resident CTAs asked of the compiler (caps the registers): 2 x 8 warps at 128 registers for
// n > 64 (shared memory allows two 86 KB matrices anyway), 5 x 4 warps for n <= 64
constexpr int blk_minb(int warps) { return warps == 8 ? 2 : (warps == 4 ? WFO_T0_MINB64 : (warps == 2 ? 8 : 16)); }

template <int WARPS, int SLOTS, bool GLOBAL>
__global__ void __launch_bounds__(WARPS * 32, blk_minb(WARPS))
det_table_blk_kernel(const double *__restrict__ S, int ld_s, int n, const int32_t *__restrict__ rows, int kb,
                     const int32_t *__restrict__ cols, int kk, double *__restrict__ D, double *gscratch) {
    extern __shared__ __align__(16) double sm[];
    const int LD = blk_ld_h(n);
    double *a = GLOBAL ? gscratch + (size_t)blockIdx.x * blk_rows(n) * LD : sm;
    int *ridx = GLOBAL ? (int *)sm : (int *)(sm + (size_t)blk_rows(n) * LD);
    int *cidx = ridx + n;
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
        double det = cta_lu_det_blocked<WARPS, SLOTS>(a, LD, n, &sh);
        if (threadIdx.x == 0) D[pair] = det;
        __syncthreads();
    }
}

template <int WARPS, int SLOTS, bool GLOBAL>
__global__ void __launch_bounds__(WARPS * 32, blk_minb(WARPS))
t0_fused_blk_kernel(const double *__restrict__ S, int ld_s, int n, const int32_t *__restrict__ rows, int ksh,
                    const int32_t *__restrict__ cols, int kk, const double *__restrict__ ckt, int ldg,
                    int chunk_len, int n_chunks, double *__restrict__ gpart, double *gscratch) {
    extern __shared__ __align__(16) double sm[];
    const int LD = blk_ld_h(n);
    double *a = GLOBAL ? gscratch + (size_t)blockIdx.x * blk_rows(n) * LD : sm;
    int *ridx = GLOBAL ? (int *)sm : (int *)(sm + (size_t)blk_rows(n) * LD);
    int *cidx = ridx + n;
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
            const double det = cta_lu_det_blocked<WARPS, SLOTS>(a, LD, n, &sh);
            if (threadIdx.x < ldg) g += det * __ldg(ckt + (size_t)l * ldg + threadIdx.x);
            __syncthreads();
        }
        if (threadIdx.x < ldg) gpart[((size_t)ch * ksh + k) * ldg + threadIdx.x] = g;
    }
}

}  // namespace wfo
