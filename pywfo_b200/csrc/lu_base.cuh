// Base-block factorisation for the rank-update tiers: in-place Gauss-Jordan inverse with
// partial (row) pivoting of A = S_MO[:n,:n] (n <= 128) in ONE CTA of 512 threads, REGISTER
// resident (32 matrix elements per thread).  Pivoting is implicit (rows never move), the
// row/column permutation is undone when the inverse is written out.
// Produces Ai = A^-1, det(A) and min/max |pivot| as a conditioning indicator.
//
// This is the "one base LU" that north_star's rank-1 (Sherman-Morrison / cofactor) updates
// reuse; upstream the same numbers would come from np.linalg.det / np.linalg.inv of
// S_MO[:occ,:occ] (pywfo/main.py:548).
#pragma once
#include "common.cuh"

namespace wfo {

struct BaseInfo {     // device-resident result block
    double det;       // det(A)
    double min_piv;   // min |pivot|
    double max_piv;   // max |pivot|
    int singular;     // 1 if an exactly zero / non-finite pivot was met
    int pad;
};

#ifdef WFO_LUB_CLOCKS
__device__ long long g_lub_clk[16];
#define LUB_T(i) do { long long now_ = clock64(); t_acc[i] += now_ - t_prev; t_prev = now_; } while (0)
#else
#define LUB_T(i)
#endif
constexpr int LUB_MAXN = 128;
constexpr int LUB_THREADS = 512;
constexpr int LUB_WARPS = LUB_THREADS / 32;   // 16
constexpr int LUB_CS = LUB_MAXN / LUB_WARPS;  // 8 column slots per thread (columns warp + 16 c)
constexpr int LUB_RS = 4;                     // 4 row slots per thread (rows 4 lane + s)
constexpr int LUB_LDS = LUB_MAXN + 1;         // row stride of the staging tile
inline size_t lub_smem_bytes(int n) { return (size_t)n * LUB_LDS * sizeof(double); }

// A_src: n x n block with leading dimension ld_src.  Ai_out: n x n, leading dimension ld_out.
// Dynamic shared memory: lub_smem_bytes(n) (staging tile for coalesced load / store).
//
// Ownership: lane l holds rows 4l..4l+3, warp w holds columns w, w+16, ..., w+112, so
//   * column j lives entirely in ONE warp (j mod 16): the pivot search is 4 compares per lane and
//     one REDUX inside that warp -- no shared memory, no barrier;
//   * row p lives in lane p/4 of EVERY warp: each warp passes its part of the pivot row through a
//     warp-private shared buffer (one lane writes, the warp reads it back) -- no block barrier;
//   * the update is branch- and select-free: the pivot row's own multiplier is published as d - 1,
//     so that a_p - (d-1) (a_p / d) = a_p / d comes out of the same multiply-add as every other row;
//   * only the multiplier column (and p, 1/pivot) goes through block-shared memory: ONE block barrier
//     per elimination step, and the warp that owns the next column updates that column first,
//     searches its pivot and publishes it while the other warps still work on the current step;
//   * det / min / max pivot are accumulated by the owning warps and combined at the end.
// The factorisation itself: lub_tile (n x n, row stride LUB_LDS, shared memory) holds A on entry and, unless
// the matrix is singular, A^-1 on return (all threads synchronised).  Every thread returns det(A), min/max |pivot|
// and the singular flag.  Shared by base_inverse_kernel (one matrix) and cofactor_rows_kernel (a batch).
__device__ __forceinline__ void lub_invert_tile(double *lub_tile, int n, double &det_out, double &min_out, double &max_out,
                                                int &sing_out) {
    __shared__ __align__(16) double mcol[2][LUB_MAXN];    // multiplier column, double buffered by step parity
    __shared__ __align__(16) double prow[LUB_WARPS][LUB_CS];   // warp-private: this warp's part of the pivot row
    __shared__ double prd[2];                             // reciprocal of the pivot
    __shared__ int pp[2];                                 // pivot row (-1: no candidate left)
    __shared__ int piv_of_col[LUB_MAXN];                  // p_j: pivot row chosen for column j
    __shared__ int col_of_piv[LUB_MAXN];                  // inverse map
    __shared__ double wdet[LUB_WARPS], wmin[LUB_WARPS], wmax[LUB_WARPS];
    __shared__ int wsing[LUB_WARPS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned FULL = 0xffffffffu;
#ifdef WFO_LUB_CLOCKS
    long long t_acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long t_prev = clock64();
#endif
    double a[LUB_RS][LUB_CS];
    unsigned donemask = 0;   // bit s: row 4 lane + s is a used pivot row or beyond n
#pragma unroll
    for (int s = 0; s < LUB_RS; s++) {
        const int r = 4 * lane + s;
        if (!(r < n)) donemask |= 1u << s;
#pragma unroll
        for (int c = 0; c < LUB_CS; c++) {
            const int col = warp + LUB_WARPS * c;
            a[s][c] = (r < n && col < n) ? lub_tile[r * LUB_LDS + col] : 0.0;
        }
    }
    double det = 1.0, minp = 1e300, maxp = 0.0;   // over the columns this warp owns
    int singular = 0;

    // pivot search in column slot C of this warp + publication (multipliers, pivot row index,
    // reciprocal) for step J; then the slot restarts as the unit column e_p (in-place Gauss-Jordan)
#define LUB_PUBLISH(C, B, J)                                                                               \
    do {                                                                                                   \
        unsigned key = 0;                                                                                  \
        _Pragma("unroll") for (int s = 0; s < LUB_RS; s++) {                                               \
            const unsigned k = (hi_abs(a[s][C]) & 0xffffff00u) | 0x80u | (unsigned)(4 * lane + s);         \
            key = max(key, ((donemask >> s) & 1u) ? 0u : k);                                               \
        }                                                                                                  \
        const unsigned kmax = __reduce_max_sync(FULL, key);                                                \
        const int p_ = (kmax != 0u) ? (int)(kmax & 0x7fu) : -1;                                            \
        const int ps_ = p_ & 3, pl_ = (p_ >> 2) & 31;                                                      \
        double cand = a[0][C];                                                                             \
        cand = (ps_ == 1) ? a[1][C] : cand;                                                                \
        cand = (ps_ == 2) ? a[2][C] : cand;                                                                \
        cand = (ps_ == 3) ? a[3][C] : cand;                                                                \
        const double d_ = __shfl_sync(FULL, cand, pl_);                                                    \
        /* multipliers; the pivot row's own entry is d - 1, so that a_p - (d-1) (a_p/d) = a_p/d */        \
        double m_[LUB_RS];                                                                                 \
        _Pragma("unroll") for (int s = 0; s < LUB_RS; s++) m_[s] = (4 * lane + s == p_) ? d_ - 1.0 : a[s][C]; \
        *reinterpret_cast<double2 *>(&mcol[B][4 * lane]) = make_double2(m_[0], m_[1]);                     \
        *reinterpret_cast<double2 *>(&mcol[B][4 * lane + 2]) = make_double2(m_[2], m_[3]);                 \
        if (lane == 0) {                                                                                   \
            prd[B] = fast_rcp(d_);                                                                         \
            pp[B] = p_;                                                                                    \
            piv_of_col[J] = p_;                                                                            \
            col_of_piv[p_ & 127] = J;                                                                      \
        }                                                                                                  \
        const double ad_ = fabs(d_);                                                                       \
        singular |= (p_ < 0 || !(ad_ > 0.0) || !(ad_ < 1e300)) ? 1 : 0;                                    \
        det *= d_;                                                                                         \
        minp = fmin(minp, ad_);                                                                            \
        maxp = fmax(maxp, ad_);                                                                            \
        _Pragma("unroll") for (int s = 0; s < LUB_RS; s++) a[s][C] = (4 * lane + s == p_) ? 1.0 : 0.0;    \
    } while (0)

    LUB_T(0);
    if (warp == 0 && n > 0) LUB_PUBLISH(0, 0, 0);
    __syncthreads();
    LUB_T(1);

#pragma unroll
    for (int cj = 0; cj < LUB_CS; cj++) {           // column slot of the current step (static register index)
        const int nsteps = min(LUB_WARPS, n - LUB_WARPS * cj);
        for (int wj = 0; wj < nsteps; wj++) {       // owning warp
            const int j = wj + LUB_WARPS * cj;
            const int b = j & 1;
            LUB_T(2);
            const int p = pp[b];
            const double rd = prd[b];
            const int ps = p & 3, pl = (p >> 2) & 31;
            // this warp's part of the pivot row: lane pl writes it (slot ps), everybody reads it back
            if (lane == pl) {
                switch (ps) {   // static register indices
#define LUB_ROW(S_)                                                                              \
    _Pragma("unroll") for (int c = 0; c < LUB_CS; c += 2)                                        \
        *reinterpret_cast<double2 *>(&prow[warp][c]) = make_double2(a[S_][c], a[S_][c + 1]);
                    case 0: LUB_ROW(0) break;
                    case 1: LUB_ROW(1) break;
                    case 2: LUB_ROW(2) break;
                    default: LUB_ROW(3) break;
#undef LUB_ROW
                }
            }
            __syncwarp();
            double pr[LUB_CS];
#pragma unroll
            for (int c = 0; c < LUB_CS; c += 2) {
                const double2 v = *reinterpret_cast<const double2 *>(&prow[warp][c]);
                pr[c] = v.x * rd;
                pr[c + 1] = v.y * rd;
            }
            LUB_T(3);
            // multipliers of this lane's rows
            const double2 f01 = *reinterpret_cast<const double2 *>(&mcol[b][4 * lane]);
            const double2 f23 = *reinterpret_cast<const double2 *>(&mcol[b][4 * lane + 2]);
            const double f[LUB_RS] = {f01.x, f01.y, f23.x, f23.y};
            donemask |= (lane == pl) ? (1u << ps) : 0u;
            LUB_T(4);
            // look-ahead: the owner of the next column updates it first and publishes its pivot
            const int jn = j + 1;
            const bool own_next = (jn < n) && (warp == (jn & (LUB_WARPS - 1)));
            if (jn < n) {
                if (wj + 1 < LUB_WARPS) {           // next column is in the same slot cj
                    if (own_next) {
#pragma unroll
                        for (int s = 0; s < LUB_RS; s++) a[s][cj] = fma(-f[s], pr[cj], a[s][cj]);
                        LUB_PUBLISH(cj, b ^ 1, jn);
                    }
                } else if (cj + 1 < LUB_CS) {       // next column is slot cj + 1 of warp 0
                    if (own_next) {
#pragma unroll
                        for (int s = 0; s < LUB_RS; s++)
                            a[s][(cj + 1) % LUB_CS] = fma(-f[s], pr[(cj + 1) % LUB_CS], a[s][(cj + 1) % LUB_CS]);
                        LUB_PUBLISH((cj + 1) % LUB_CS, b ^ 1, jn);
                    }
                }
            }
            LUB_T(5);
            // the rest of the update (the column handled by the look-ahead is skipped by its owner)
            const int skip = !own_next ? -1 : ((wj + 1 < LUB_WARPS) ? cj : cj + 1);
#pragma unroll
            for (int c = 0; c < LUB_CS; c++) {
                if (c == skip) continue;
#pragma unroll
                for (int s = 0; s < LUB_RS; s++) a[s][c] = fma(-f[s], pr[c], a[s][c]);
            }
            // keep the bulk update on this side of the barrier: sunk into the next step it would sit in
            // front of that step's critical path (pivot row exchange -> look-ahead publish)
#pragma unroll
            for (int c = 0; c < LUB_CS; c++)
#pragma unroll
                for (int s = 0; s < LUB_RS; s++) asm volatile("" : "+d"(a[s][c]));
            LUB_T(6);
            __syncthreads();
            LUB_T(7);
        }
    }
#undef LUB_PUBLISH
    LUB_T(8);
    // ---- combine the per-warp pivot statistics
    if (lane == 0) { wdet[warp] = det; wmin[warp] = minp; wmax[warp] = maxp; wsing[warp] = singular; }
    __syncthreads();
    double tdet = 1.0, tmin = 1e300, tmax = 0.0;
    int tsing = 0;
#pragma unroll
    for (int w = 0; w < LUB_WARPS; w++) {
        tdet *= wdet[w];
        tmin = fmin(tmin, wmin[w]);
        tmax = fmax(tmax, wmax[w]);
        tsing |= wsing[w];
    }
    // ---- write out: stored[r][j'] = M[r][p_j'],  Ai[j][c] = M[p_j][c]  =>  Ai[col_of_piv[r]][piv_of_col[j']] = stored[r][j']
    // (permuted into the staging tile, then a coalesced copy)
    if (!tsing) {
#pragma unroll
        for (int s = 0; s < LUB_RS; s++) {
            const int r = 4 * lane + s;
            if (r < n) {
                const int orow = col_of_piv[r];
#pragma unroll
                for (int c = 0; c < LUB_CS; c++) {
                    const int col = warp + LUB_WARPS * c;
                    if (col < n) lub_tile[orow * LUB_LDS + piv_of_col[col]] = a[s][c];
                }
            }
        }
    }
    // sign of the permutation j -> piv_of_col[j]: parallel inversion count
    int inv = 0;
    if (!tsing && tid < n)
        for (int k = tid + 1; k < n; k++) inv += piv_of_col[k] < piv_of_col[tid];
    const int par = __syncthreads_count(inv & 1) & 1;   // also: the permuted inverse in lub_tile is complete
    LUB_T(9);
#ifdef WFO_LUB_CLOCKS
    if (tid == 32) for (int i = 0; i < 10; i++) g_lub_clk[i] = t_acc[i];
#endif
    det_out = tsing ? 0.0 : (par ? -tdet : tdet);
    min_out = tmin;
    max_out = tmax;
    sing_out = tsing;
}

// A_src: n x n block with leading dimension ld_src.  Ai_out: n x n, leading dimension ld_out.
// Dynamic shared memory: lub_smem_bytes(n) (staging tile for coalesced load / store).
__global__ void __launch_bounds__(LUB_THREADS)
base_inverse_kernel(const double *__restrict__ A_src, int ld_src, int n, double *__restrict__ Ai_out, int ld_out,
                    BaseInfo *__restrict__ info) {
    extern __shared__ __align__(16) double lub_tile[];    // n x LUB_LDS
    const int tid = threadIdx.x;
    // coalesced load through the staging tile
    for (int idx = tid; idx < n * n; idx += LUB_THREADS) {
        const int r = idx / n, c = idx - r * n;
        lub_tile[r * LUB_LDS + c] = A_src[(size_t)r * ld_src + c];
    }
    __syncthreads();
    double det, minp, maxp;
    int sing;
    lub_invert_tile(lub_tile, n, det, minp, maxp, sing);
    if (!sing)
        for (int idx = tid; idx < n * n; idx += LUB_THREADS) {
            const int r = idx / n, c = idx - r * n;
            Ai_out[(size_t)r * ld_out + c] = lub_tile[r * LUB_LDS + c];
        }
    if (tid == 0) {
        info->det = det;
        info->min_piv = minp;
        info->max_piv = maxp;
        info->singular = sing;
    }
}

// Super-blocks of a general determinant list (pywfo/dets.py:119-156: blocks that are equal up to the LAST
// orbital; the reference prepares `super_map` and counts the super-blocks, dets.py:233-236, but then takes one
// np.linalg.det per block pair): the n - 1 shared rows are factored ONCE per (super-block, ket block) and
// every member costs one dot product -- the Laplace expansion along the last row that laplace.py sketches,
//     det([A; r_m]) = sum_j r_m[j] C_j,   C_j = cofactor of (n, j) = det(M) (M^-1)[j][n]   for M = [A; r_rep],
// with the cofactors taken from the pivoted Gauss-Jordan inverse of the representative member's matrix
// (the cofactors of the last row do not depend on that row).
// One CTA per (super-block, ket block).  rows_rep: (n_super x n) index lists of the representatives;
// mem_off (n_super + 1), mem_block / mem_row: member block ids and their last-row indices; cols: (n_kblk x n).
// D (n_bblk x n_kblk, row stride ld_d) gets the members' determinants; flags (n_super x n_kblk) is set
// when the representative's matrix is singular or ill conditioned (pivot ratio below 1e-6): the caller
// recomputes those members by brute force.
__global__ void __launch_bounds__(LUB_THREADS)
cofactor_rows_kernel(const double *__restrict__ S, int ld_s, int n, const int32_t *__restrict__ rows_rep, int n_super,
                     const int32_t *__restrict__ mem_off, const int32_t *__restrict__ mem_block,
                     const int32_t *__restrict__ mem_row, const int32_t *__restrict__ cols, int n_kblk,
                     double *__restrict__ D, int ld_d, unsigned char *__restrict__ flags) {
    extern __shared__ __align__(16) double lub_tile[];    // n x LUB_LDS
    __shared__ double cof[LUB_MAXN];
    __shared__ int colv[LUB_MAXN];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (long long pair = blockIdx.x; pair < (long long)n_super * n_kblk; pair += gridDim.x) {
        const int sb = (int)(pair / n_kblk), kb = (int)(pair - (long long)sb * n_kblk);
        const int32_t *rl = rows_rep + (size_t)sb * n, *cl = cols + (size_t)kb * n;
        __syncthreads();   // the previous pair is done with the tile
        for (int idx = tid; idx < n * n; idx += LUB_THREADS) {
            const int r = idx / n, c = idx - r * n;
            lub_tile[r * LUB_LDS + c] = __ldg(S + (size_t)rl[r] * ld_s + cl[c]);
        }
        if (tid < n) colv[tid] = cl[tid];
        __syncthreads();
        double det, minp, maxp;
        int sing;
        lub_invert_tile(lub_tile, n, det, minp, maxp, sing);
        const bool bad = sing || !(minp > 1e-6 * maxp);
        if (tid == 0) flags[pair] = bad ? 1 : 0;
        if (bad) continue;   // uniform
        if (tid < n) cof[tid] = det * lub_tile[tid * LUB_LDS + (n - 1)];
        __syncthreads();
        // members: one warp per member, dot product of its last row with the cofactors
        for (int m = mem_off[sb] + warp; m < mem_off[sb + 1]; m += LUB_WARPS) {
            const double *srow = S + (size_t)mem_row[m] * ld_s;
            double acc = 0.0;
            for (int j = lane; j < n; j += 32) acc = fma(__ldg(srow + colv[j]), cof[j], acc);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) D[(size_t)mem_block[m] * ld_d + kb] = acc;
        }
    }
}

// verdict of a 2x2 block inverse: A = [[A11, A12], [A21, A22]], det(A) = det(A11) det(S), S the Schur complement
__global__ void combine_info_kernel(const BaseInfo *__restrict__ a, const BaseInfo *__restrict__ b, BaseInfo *__restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        BaseInfo r;
        r.det = a->det * b->det;
        r.min_piv = fmin(a->min_piv, b->min_piv);
        r.max_piv = fmax(a->max_piv, b->max_piv);
        r.singular = a->singular | b->singular;
        r.pad = 0;
        *out = r;
    }
}

// dst (rows x cols, ld_dst) = src (ld_src)
__global__ void copy2d_kernel(const double *__restrict__ src, int ld_src, double *__restrict__ dst, int ld_dst, int rows,
                              int cols) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)rows * cols) return;
    const int r = (int)(idx / cols), c = (int)(idx - (long long)r * cols);
    dst[(size_t)r * ld_dst + c] = src[(size_t)r * ld_src + c];
}

}  // namespace wfo
