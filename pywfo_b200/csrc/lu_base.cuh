// Base-block factorisation for the rank-update tiers: Gauss-Jordan inverse with
// partial (row) pivoting of A = S_MO[:n,:n] (n <= 128) in ONE CTA of 512 threads,
// REGISTER resident: warp w owns rows {w, w+16, ..., w+112}, lane l owns columns
// {l, l+32, l+64, l+96} -> 32 matrix elements per thread.  Per elimination step only
// the pivot row and the multiplier column travel through shared memory (2 block
// barriers per step); pivoting is implicit (rows never move), the row/column
// permutation is undone when the inverse is written out.
// Produces Ai = A^-1, det(A) and min/max |pivot| as a conditioning indicator.
//
// This is the "one base LU" that north_star's rank-1 (Sherman-Morrison / cofactor)
// updates reuse; upstream the same numbers would come from np.linalg.det /
// np.linalg.inv of S_MO[:occ,:occ] (pywfo/main.py:548).
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

constexpr int LUB_MAXN = 128;
constexpr int LUB_THREADS = 512;
constexpr int LUB_WARPS = LUB_THREADS / 32;   // 16
constexpr int LUB_RS = LUB_MAXN / LUB_WARPS;     // 8 row slots per thread

// A_src: n x n block with leading dimension ld_src.  Ai_out: n x n, ld = n.
__global__ void __launch_bounds__(LUB_THREADS)
base_inverse_kernel(const double *__restrict__ A_src, int ld_src, int n, double *__restrict__ Ai_out,
                    BaseInfo *__restrict__ info) {
    // double buffered by step parity: a step's writes never race the previous step's reads
    __shared__ unsigned keys[2][32];   // LUB_WARPS used, rest stays 0
    __shared__ double prow[2][LUB_MAXN];   // scaled pivot row
    __shared__ double mcol[2][LUB_MAXN];   // multiplier column
    __shared__ int piv_of_col[LUB_MAXN];   // p_j: pivot row chosen for column j
    __shared__ int col_of_piv[LUB_MAXN];   // inverse map
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    double a[LUB_RS][4];
    bool done[LUB_RS];
    if (tid < 64) keys[tid >> 5][tid & 31] = 0;
    __syncthreads();
#pragma unroll
    for (int s = 0; s < LUB_RS; s++) {
        const int r = warp + LUB_WARPS * s;
        done[s] = !(r < n);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int c = lane + 32 * k;
            a[s][k] = (r < n && c < n) ? A_src[(size_t)r * ld_src + c] : 0.0;
        }
    }
    double det = 1.0, minp = 1e300, maxp = 0.0;
    int singular = 0;

    // straight-line step body (no early exits, selects instead of branches): the compiler keeps
    // the 32 matrix registers in place; a zero pivot only sets a sticky flag
#pragma unroll
    for (int cj = 0; cj < 4; cj++) {          // column slot (static register index)
        const int nsteps = min(32, n - 32 * cj);
        for (int lj = 0; lj < nsteps; lj++) { // owning lane
            const int j = lj + 32 * cj;
            const int b = j & 1;
            // ---- pivot search: per-warp candidate from the lane that owns column j; the owner
            // publishes the column (the multipliers) and restarts its column slot from zero
            if (lane == lj) {
                unsigned key = 0;
#pragma unroll
                for (int s = 0; s < LUB_RS; s++) {
                    const unsigned k = (hi_abs(a[s][cj]) & 0xffffff00u) | 0x80u | (unsigned)(warp + LUB_WARPS * s);
                    key = max(key, done[s] ? 0u : k);
                    mcol[b][warp + LUB_WARPS * s] = a[s][cj];
                    a[s][cj] = 0.0;
                }
                keys[b][warp] = key;
            }
            __syncthreads();
            const unsigned kmax = __reduce_max_sync(0xffffffffu, keys[b][lane]);
            const int p = (int)(kmax & 0x7fu);          // pivot row
            const int wp = p % LUB_WARPS, sp = p / LUB_WARPS;
            const double d = mcol[b][p];
            const double ad = fabs(d);
            singular |= (kmax == 0u || !(ad > 0.0) || !(ad < 1e300)) ? 1 : 0;
            det *= d;
            minp = fmin(minp, ad);
            maxp = fmax(maxp, ad);
            const double rd = fast_rcp(d);
            if (warp == wp) {                           // publish the scaled pivot row
                double row[4];
#pragma unroll
                for (int k = 0; k < 4; k++) row[k] = a[0][k];
#pragma unroll
                for (int s = 1; s < LUB_RS; s++)
#pragma unroll
                    for (int k = 0; k < 4; k++) row[k] = (s == sp) ? a[s][k] : row[k];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int c = lane + 32 * k;
                    prow[b][c] = ((c == j) ? 1.0 : row[k]) * rd;
                }
                if (lane == 0) { piv_of_col[j] = p; col_of_piv[p] = j; }
            }
            __syncthreads();
            double pr[4];
#pragma unroll
            for (int k = 0; k < 4; k++) pr[k] = prow[b][lane + 32 * k];
            if (warp != wp) {                           // 15 of 16 warps: pure multiply-add
#pragma unroll
                for (int s = 0; s < LUB_RS; s++) {
                    const double f = mcol[b][warp + LUB_WARPS * s];
#pragma unroll
                    for (int k = 0; k < 4; k++) a[s][k] = fma(-f, pr[k], a[s][k]);
                }
            } else {                                    // the warp that holds the pivot row
#pragma unroll
                for (int s = 0; s < LUB_RS; s++) {
                    const bool is_p = (s == sp);
                    done[s] = done[s] || is_p;
                    const double f = is_p ? 0.0 : mcol[b][warp + LUB_WARPS * s];
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const double upd = fma(-f, pr[k], a[s][k]);
                        a[s][k] = is_p ? pr[k] : upd;
                    }
                }
            }
        }
    }
    // ---- write out: stored[r][j'] = M[r][p_j'],  Ai[j][c] = M[p_j][c]  =>  Ai[col_of_piv[r]][piv_of_col[j']] = stored[r][j']
    if (!singular) {
#pragma unroll
        for (int s = 0; s < LUB_RS; s++) {
            const int r = warp + LUB_WARPS * s;
            if (r < n) {
                const int orow = col_of_piv[r];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int c = lane + 32 * k;
                    if (c < n) Ai_out[(size_t)orow * n + piv_of_col[c]] = a[s][k];
                }
            }
        }
    }
    // sign of the permutation j -> piv_of_col[j]: parallel inversion count
    __syncthreads();
    int inv = 0;
    if (!singular && tid < n)
        for (int k = tid + 1; k < n; k++) inv += piv_of_col[k] < piv_of_col[tid];
    const int par = __syncthreads_count(inv & 1) & 1;
    if (tid == 0) {
        info->det = singular ? 0.0 : (par ? -det : det);
        info->min_piv = minp;
        info->max_piv = maxp;
        info->singular = singular;
    }
}

}  // namespace wfo
