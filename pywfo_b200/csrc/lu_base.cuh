// Base-block factorisation for the rank-update tiers: in-place Gauss-Jordan
// inverse with partial (row) pivoting of A = S_MO[:n,:n] (n x n, n <= 160) in
// one CTA, shared-memory resident.  Produces Ai = A^-1, det(A) and a
// conditioning indicator (min |pivot| / max |pivot|).
//
// This is the "one base LU" that north_star's rank-1 (Sherman-Morrison /
// cofactor) updates reuse; in the reference the same numbers would come from
// np.linalg.det / np.linalg.inv of S_MO[:occ,:occ] (pywfo/main.py:548).
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

constexpr int LUB_THREADS = 256;

// A_src: n x n block with leading dimension ld_src.  Ai_out: n x n, ld = n.
__global__ void __launch_bounds__(LUB_THREADS)
base_inverse_kernel(const double *__restrict__ A_src, int ld_src, int n, double *__restrict__ Ai_out,
                    BaseInfo *__restrict__ info) {
    extern __shared__ double sm[];
    const int ld = n | 1;                  // odd stride: conflict-free column walks
    double *a = sm;                        // n x ld
    double *colj = a + (size_t)n * ld;     // n   (eliminated column)
    int *piv = (int *)(colj + n);          // n
    __shared__ double red_v[LUB_THREADS / 32];
    __shared__ int red_i[LUB_THREADS / 32];
    __shared__ int s_p;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int idx = tid; idx < n * n; idx += LUB_THREADS) {
        int r = idx / n, c = idx % n;
        a[r * ld + c] = A_src[(size_t)r * ld_src + c];
    }
    double det = 1.0, minp = 1e300, maxp = 0.0;
    int singular = 0;
    __syncthreads();

    for (int j = 0; j < n; j++) {
        // pivot search over rows j..n-1 of column j
        double best = -1.0;
        int bi = j;
        for (int r = j + tid; r < n; r += LUB_THREADS) {
            double v = fabs(a[r * ld + j]);
            if (v > best) { best = v; bi = r; }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            double ov = __shfl_xor_sync(0xffffffffu, best, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (lane == 0) { red_v[warp] = best; red_i[warp] = bi; }
        __syncthreads();
        if (tid == 0) {
            double bv = red_v[0];
            int bb = red_i[0];
            for (int w = 1; w < LUB_THREADS / 32; w++)
                if (red_v[w] > bv || (red_v[w] == bv && red_i[w] < bb)) { bv = red_v[w]; bb = red_i[w]; }
            s_p = bb;
            piv[j] = bb;
        }
        __syncthreads();
        const int p = s_p;
        if (p != j) {
            for (int c = tid; c < n; c += LUB_THREADS) {
                double x = a[j * ld + c];
                a[j * ld + c] = a[p * ld + c];
                a[p * ld + c] = x;
            }
            det = -det;
        }
        __syncthreads();
        const double d = a[j * ld + j];
        det *= d;
        const double ad = fabs(d);
        minp = fmin(minp, ad);
        maxp = fmax(maxp, ad);
        if (!(ad > 0.0) || !isfinite(d)) { singular = 1; break; }   // uniform: d is shared
        const double rd = 1.0 / d;
        // save column j, scale pivot row
        for (int r = tid; r < n; r += LUB_THREADS) colj[r] = a[r * ld + j];
        __syncthreads();
        for (int c = tid; c < n; c += LUB_THREADS) a[j * ld + c] = (c == j ? 1.0 : a[j * ld + c]) * rd;
        __syncthreads();
        // eliminate column j from every other row:  row_i -= colj[i] * row_j, a[i][j] = -colj[i]*rd
        for (int idx = tid; idx < n * n; idx += LUB_THREADS) {
            int r = idx / n, c = idx % n;
            if (r == j) continue;
            double f = colj[r];
            double base = (c == j) ? 0.0 : a[r * ld + c];
            a[r * ld + c] = base - f * a[j * ld + c];
        }
        __syncthreads();
    }
    if (!singular) {
        // undo the row interchanges as column interchanges, last first
        for (int j = n - 1; j >= 0; j--) {
            int p = piv[j];
            if (p != j) {
                for (int r = tid; r < n; r += LUB_THREADS) {
                    double x = a[r * ld + j];
                    a[r * ld + j] = a[r * ld + p];
                    a[r * ld + p] = x;
                }
                __syncthreads();
            }
        }
        for (int idx = tid; idx < n * n; idx += LUB_THREADS) {
            int r = idx / n, c = idx % n;
            Ai_out[(size_t)r * n + c] = a[r * ld + c];
        }
    }
    if (tid == 0) {
        info->det = singular ? 0.0 : det;
        info->min_piv = minp;
        info->max_piv = maxp;
        info->singular = singular;
    }
}

}  // namespace wfo
