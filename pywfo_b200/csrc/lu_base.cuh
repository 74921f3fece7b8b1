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
    double *colj = a + (size_t)n * ld;     // n   (unused spare column)
    int *piv = (int *)(colj + n);          // n
    __shared__ int red_i[LUB_THREADS / 32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int idx = tid; idx < n * n; idx += LUB_THREADS) {
        int r = idx / n, c = idx % n;
        a[r * ld + c] = A_src[(size_t)r * ld_src + c];
    }
    double det = 1.0, minp = 1e300, maxp = 0.0;
    int singular = 0;
    __syncthreads();

    const int nwarps = LUB_THREADS / 32;
    for (int j = 0; j < n; j++) {
        // ---- pivot search over rows j..n-1 of column j: REDUX over (|a| high word | row)
        unsigned key = 0;
        for (int r = j + tid; r < n; r += LUB_THREADS)
            key = max(key, (hi_abs(a[r * ld + j]) & 0xffffff00u) | (unsigned)r);   // n <= 160 < 256
        key = __reduce_max_sync(0xffffffffu, key);
        if (lane == 0) red_i[warp] = (int)key;
        __syncthreads();
        unsigned kmax = 0;
#pragma unroll
        for (int w = 0; w < nwarps; w++) kmax = max(kmax, (unsigned)red_i[w]);
        const int p = (int)(kmax & 0xffu);
        if (tid == 0) piv[j] = p;
        const double d = a[p * ld + j];          // pivot value (before the swap)
        const double ad = fabs(d);
        if (!(ad > 0.0) || !isfinite(d)) { singular = 1; break; }   // uniform
        det *= (p != j) ? -d : d;
        minp = fmin(minp, ad);
        maxp = fmax(maxp, ad);
        const double rd = 1.0 / d;
        // ---- lanes own columns (4 per lane), warps walk rows.  Row swap folded in:
        //      new row j = old row p scaled, new row p = old row j eliminated.
        double pr[5], oj[5];
#pragma unroll
        for (int k = 0; k < 5; k++) {
            const int c = lane + 32 * k;
            pr[k] = (c < n) ? ((c == j) ? 1.0 : a[p * ld + c]) * rd : 0.0;   // scaled pivot row
            oj[k] = (c < n) ? a[j * ld + c] : 0.0;                            // old row j
        }
        const double fj = a[j * ld + j];           // old a[j][j]: multiplier of the displaced row
        __syncthreads();                          // everyone has read rows j and p
        for (int r = warp; r < n; r += nwarps) {
            if (r == j) {
#pragma unroll
                for (int k = 0; k < 5; k++) { const int c = lane + 32 * k; if (c < n) a[j * ld + c] = pr[k]; }
            } else {
                const bool is_p = (r == p);
                const double f = is_p ? fj : a[r * ld + j];
#pragma unroll
                for (int k = 0; k < 5; k++) {
                    const int c = lane + 32 * k;
                    if (c < n) {
                        const double base = (c == j) ? 0.0 : (is_p ? oj[k] : a[r * ld + c]);
                        a[r * ld + c] = fma(-f, pr[k], base);
                    }
                }
            }
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
