// General determinant lists (pywfo/dets.py:119-283): the contraction
//     wfo[i][j] = sum_{d,e} bc[d][i] * alpha[ba_of[d]][ka_of[e]] * beta[bb_of[d]][kb_of[e]] * kc[e][j]
// on the device, with the block maps as gather indices: neither the block-determinant tables nor the
// (bra determinants x ket determinants) matrix SS of dets.py:280 ever leaves HBM -- SS is never formed.
#pragma once
#include "common.cuh"

namespace wfo {

constexpr int LISTS_SCHUNK = 8;   // bra states per pass

// tmp[e][i] = sum_d bc[d][i] * alpha[ba_of[d]][ka_of[e]] * beta[bb_of[d]][kb_of[e]]; one CTA per ket determinant e.
// alpha / beta may be NULL (no electrons of that spin: the factor is 1).
__global__ void __launch_bounds__(256)
lists_contract_kernel(const double *__restrict__ alpha, int ld_a, const double *__restrict__ beta, int ld_b,
                      const int32_t *__restrict__ ba_of, const int32_t *__restrict__ bb_of, int nd_b,
                      const int32_t *__restrict__ ka_of, const int32_t *__restrict__ kb_of,
                      const double *__restrict__ bc, int ns_b, double *__restrict__ tmp) {
    __shared__ double red[8][LISTS_SCHUNK];
    const int e = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ka = alpha ? ka_of[e] : 0, kb = beta ? kb_of[e] : 0;
    for (int i0 = 0; i0 < ns_b; i0 += LISTS_SCHUNK) {
        double acc[LISTS_SCHUNK];
#pragma unroll
        for (int i = 0; i < LISTS_SCHUNK; i++) acc[i] = 0.0;
        for (int d = tid; d < nd_b; d += 256) {
            double ss = 1.0;
            if (alpha) ss *= __ldg(alpha + (size_t)ba_of[d] * ld_a + ka);
            if (beta) ss *= __ldg(beta + (size_t)bb_of[d] * ld_b + kb);
#pragma unroll
            for (int i = 0; i < LISTS_SCHUNK; i++)
                if (i0 + i < ns_b) acc[i] = fma(__ldg(bc + (size_t)d * ns_b + i0 + i), ss, acc[i]);
        }
#pragma unroll
        for (int i = 0; i < LISTS_SCHUNK; i++) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
        }
        __syncthreads();
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < LISTS_SCHUNK; i++) red[warp][i] = acc[i];
        }
        __syncthreads();
        if (tid < LISTS_SCHUNK && i0 + tid < ns_b) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < 8; w++) s += red[w][tid];   // fixed order: deterministic
            tmp[(size_t)e * ns_b + i0 + tid] = s;
        }
    }
}

// dst[row_of[r]][:] = src[r][:]
__global__ void scatter_rows_kernel(const double *__restrict__ src, int n_rows, int n_cols, const int32_t *__restrict__ row_of,
                                    double *__restrict__ dst, int ld_dst) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)n_rows * n_cols) return;
    const int r = (int)(idx / n_cols), c = (int)(idx - (long long)r * n_cols);
    dst[(size_t)row_of[r] * ld_dst + c] = src[idx];
}

}  // namespace wfo
