// T0 (brute force), generic kernel: one CTA per determinant pair, the gathered
// n x n minor of S_MO staged in shared memory, right-looking LU with partial
// pivoting (block arg-max by warp shuffles), determinant = signed product of
// pivots.  Valid for any index lists and any n with n*(n|1)*8 bytes of shared
// memory (n <= 160).  This is the reference's np.linalg.det(S_MO[rows][:, cols])
// (pywfo/main.py:583-584) evaluated pair by pair.
//
// Two front ends share the device function:
//   det_table_kernel : D[k][l]                       (test / bench / GS row+col)
//   t0_fused_kernel  : G[k][j] += D(k,l) * CkT[l][j] (per-pair value stays on chip)
#pragma once
#include "common.cuh"

namespace wfo {

constexpr int T0_MAX_WARPS = 8;

struct LuShared {
    double red_v[T0_MAX_WARPS];
    int red_i[T0_MAX_WARPS];
};

// All threads of the CTA call this; `a` is n x ld in shared memory and is
// destroyed.  Returns det to every thread.  Requires n <= 128.
__device__ __forceinline__ double cta_lu_det(double *a, int n, int ld, LuShared *sh) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nthreads = blockDim.x, nw = nthreads >> 5;
    double det = 1.0;
    for (int j = 0; j < n; j++) {
        // ---- pivot search in column j, rows j..n-1
        double best = -1.0;
        int bi = j;
        for (int r = j + tid; r < n; r += nthreads) {
            double v = fabs(a[r * ld + j]);
            if (v > best) { best = v; bi = r; }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            double ov = __shfl_xor_sync(0xffffffffu, best, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (lane == 0) { sh->red_v[warp] = best; sh->red_i[warp] = bi; }
        __syncthreads();
        double bv = sh->red_v[0];
        int p = sh->red_i[0];
        for (int w = 1; w < nw; w++) {
            double v = sh->red_v[w];
            int i = sh->red_i[w];
            if (v > bv || (v == bv && i < p)) { bv = v; p = i; }
        }
        if (!(bv > 0.0)) return 0.0;   // exactly singular (uniform across the CTA)
        // ---- row interchange on the trailing columns
        if (p != j) {
            for (int c = j + tid; c < n; c += nthreads) {
                double x = a[j * ld + c];
                a[j * ld + c] = a[p * ld + c];
                a[p * ld + c] = x;
            }
            det = -det;
        }
        __syncthreads();
        const double d = a[j * ld + j];
        det *= d;
        const double rd = 1.0 / d;
        // ---- rank-1 update of the trailing block: lanes walk columns, warps walk rows
        double u[4];
#pragma unroll
        for (int s = 0; s < 4; s++) {
            int c = j + 1 + lane + 32 * s;
            u[s] = (c < n) ? a[j * ld + c] : 0.0;
        }
        for (int r = j + 1 + warp; r < n; r += nw) {
            const double l = a[r * ld + j] * rd;
#pragma unroll
            for (int s = 0; s < 4; s++) {
                int c = j + 1 + lane + 32 * s;
                if (c < n) a[r * ld + c] -= l * u[s];
            }
        }
        __syncthreads();
    }
    return det;
}

// gather the minor rows[0..n) x cols[0..n) of S (leading dimension ld_s) into a (n x ld)
__device__ __forceinline__ void gather_minor(const double *__restrict__ S, int ld_s, const int *rows,
                                             const int *cols, int n, double *a, int ld) {
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
        int r = idx / n, c = idx - r * n;
        a[r * ld + c] = __ldg(S + (size_t)rows[r] * ld_s + cols[c]);
    }
}

__global__ void det_table_kernel(const double *__restrict__ S, int ld_s, int n,
                                 const int32_t *__restrict__ rows, int kb,
                                 const int32_t *__restrict__ cols, int kk, double *__restrict__ D) {
    extern __shared__ double sm[];
    const int ld = n | 1;
    double *a = sm;
    int *ridx = (int *)(a + (size_t)n * ld);
    int *cidx = ridx + n;
    __shared__ LuShared sh;
    const long long npairs = (long long)kb * kk;
    for (long long pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
        const int k = (int)(pair / kk), l = (int)(pair - (long long)k * kk);
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            ridx[i] = rows[(size_t)k * n + i];
            cidx[i] = cols[(size_t)l * n + i];
        }
        __syncthreads();
        gather_minor(S, ld_s, ridx, cidx, n, a, ld);
        __syncthreads();
        double det = cta_lu_det(a, n, ld, &sh);
        if (threadIdx.x == 0) D[pair] = det;
        __syncthreads();
    }
}

// One work item = (bra determinant k of the shard, chunk of ket determinants).
// gpart[(chunk*ksh + k)*ldg + j] = sum_{l in chunk} D(k,l) * ckt[l*ldg + j]
__global__ void t0_fused_kernel(const double *__restrict__ S, int ld_s, int n,
                                const int32_t *__restrict__ rows, int ksh,
                                const int32_t *__restrict__ cols, int kk,
                                const double *__restrict__ ckt, int ldg, int chunk_len, int n_chunks,
                                double *__restrict__ gpart) {
    extern __shared__ double sm[];
    const int ld = n | 1;
    double *a = sm;
    int *ridx = (int *)(a + (size_t)n * ld);
    int *cidx = ridx + n;
    __shared__ LuShared sh;
    const long long nitems = (long long)ksh * n_chunks;
    for (long long item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int k = (int)(item / n_chunks), ch = (int)(item - (long long)k * n_chunks);
        for (int i = threadIdx.x; i < n; i += blockDim.x) ridx[i] = rows[(size_t)k * n + i];
        double g = 0.0;
        const int l0 = ch * chunk_len, l1 = min(kk, l0 + chunk_len);
        for (int l = l0; l < l1; l++) {
            for (int i = threadIdx.x; i < n; i += blockDim.x) cidx[i] = cols[(size_t)l * n + i];
            __syncthreads();
            gather_minor(S, ld_s, ridx, cidx, n, a, ld);
            __syncthreads();
            double det = cta_lu_det(a, n, ld, &sh);
            if (threadIdx.x < ldg) g += det * __ldg(ckt + (size_t)l * ldg + threadIdx.x);
            __syncthreads();
        }
        if (threadIdx.x < ldg) gpart[((size_t)ch * ksh + k) * ldg + threadIdx.x] = g;
    }
}

}  // namespace wfo
