// T0 (brute force), small matrices (n <= 32): one determinant per warp SEGMENT, register resident.
//
// A segment of SEG lanes (SEG = 8, 16 or 32 >= n) owns one gathered n x n minor: lane = row,
// the row's NC >= n columns live in registers.  Unblocked LU with partial pivoting, no shared
// memory and no barrier: per column one segmented REDUX.MAX over (|a| high word, lane) finds
// the pivot row, its remaining entries (and the reciprocal of the pivot, computed
// speculatively by every lane) are broadcast with shuffles, the elimination is branch-free
// with implicit pivoting (finished rows get a zero multiplier).  The sign is the parity of
// the pivot order (inversion count with shuffles).  Four (SEG=8) or two (SEG=16) minors are
// factored per warp at once -- this is the bundled-test regime of pywfo (n_occ = 5, 9).
// Same quantity as np.linalg.det(S_MO[rows][:, cols]) (pywfo/main.py:583-584).
#pragma once
#include "common.cuh"

namespace wfo {

// lane-local row a[0..NC) of the minor; sl = row index inside the segment; returns det to
// every lane of the segment
template <int NC, int SEG>
__device__ __forceinline__ double seg_lu_det(double (&a)[NC], int n, int sl, unsigned mask) {
    bool done = sl >= n;
    int my_step = 32 + sl;          // rows beyond n: never pivots, no inversions with real rows
    double det = 1.0;
    bool zero = false;
#pragma unroll
    for (int j = 0; j < NC; j++) {
        if (j < n) {   // uniform
            const unsigned key = done ? 0u : ((hi_abs(a[j]) & 0xffffffc0u) | 0x20u | (unsigned)sl);
            const double rc_own = fast_rcp(done ? 1.0 : a[j]);
            const unsigned kmax = __reduce_max_sync(mask, key);
            const int wl = kmax & 31;                      // segment-relative pivot lane
            double u[NC];
#pragma unroll
            for (int c = j; c < NC; c++) u[c] = __shfl_sync(mask, a[c], wl, SEG);
            const double rcp = __shfl_sync(mask, rc_own, wl, SEG);
            det *= u[j];
            zero = zero || (u[j] == 0.0);
            const bool win = (key == kmax) && (key != 0u);
            my_step = win ? j : my_step;
            done = done || win;
            const double l = done ? 0.0 : a[j] * rcp;
            a[j] = done ? a[j] : l;
#pragma unroll
            for (int c = j + 1; c < NC; c++) a[c] = fma(-l, u[c], a[c]);
        }
    }
    // sign: parity of the inversions of row -> step
    int cnt = 0;
#pragma unroll
    for (int o = 0; o < SEG; o++) {
        const int s2 = __shfl_sync(mask, my_step, o, SEG);
        cnt += (o > sl && s2 < my_step) ? 1 : 0;
    }
    const unsigned odd = __ballot_sync(mask, cnt & 1) & mask;
    if (__popc(odd) & 1) det = -det;
    return zero ? 0.0 : det;
}

// gather row sl of the minor (rows[sl], cols[0..n)) into registers; columns >= n are zero
template <int NC, int SEG>
__device__ __forceinline__ void seg_gather(const double *__restrict__ S, int ld_s, const int32_t *__restrict__ rows,
                                           const int32_t *__restrict__ cols, int n, int sl, unsigned mask,
                                           double (&a)[NC]) {
    const int myrow = (sl < n) ? rows[sl] : 0;
    const int mycol = (sl < n) ? cols[sl] : 0;
    const double *src = S + (size_t)myrow * ld_s;
#pragma unroll
    for (int c = 0; c < NC; c++) {
        const int cc = __shfl_sync(mask, mycol, c < SEG ? c : 0, SEG);
        a[c] = (c < n && sl < n) ? __ldg(src + cc) : 0.0;
    }
}

template <int NC, int SEG>
__global__ void __launch_bounds__(128)
det_table_warp_kernel(const double *__restrict__ S, int ld_s, int n, const int32_t *__restrict__ rows, int kb,
                      const int32_t *__restrict__ cols, int kk, double *__restrict__ D) {
    constexpr int SPW = 32 / SEG;   // segments per warp
    const int lane = threadIdx.x & 31, sg = lane / SEG, sl = lane % SEG;
    const unsigned mask = (SEG == 32) ? 0xffffffffu : (((1u << SEG) - 1u) << (sg * SEG));
    const long long nseg = (long long)gridDim.x * (blockDim.x >> 5) * SPW;
    const long long seg0 = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * SPW + sg;
    const long long npairs = (long long)kb * kk;
    for (long long pair = seg0; pair < npairs; pair += nseg) {
        const int k = (int)(pair / kk), l = (int)(pair - (long long)k * kk);
        double a[NC];
        seg_gather<NC, SEG>(S, ld_s, rows + (size_t)k * n, cols + (size_t)l * n, n, sl, mask, a);
        const double det = seg_lu_det<NC, SEG>(a, n, sl, mask);
        if (sl == 0) D[pair] = det;
    }
}

// One work item per WARP = (bra determinant k, chunk of ket determinants); the SPW segments
// of the warp take consecutive ket determinants, then all 32 lanes accumulate
// g[j] += D(k,l) * CkT[l][j] for their columns j = lane, lane + 32.
template <int NC, int SEG>
__global__ void __launch_bounds__(128)
t0_fused_warp_kernel(const double *__restrict__ S, int ld_s, int n, const int32_t *__restrict__ rows, int ksh,
                     const int32_t *__restrict__ cols, int kk, const double *__restrict__ ckt, int ldg,
                     int chunk_len, int n_chunks, double *__restrict__ gpart) {
    constexpr int SPW = 32 / SEG;
    const int lane = threadIdx.x & 31, sg = lane / SEG, sl = lane % SEG;
    const unsigned mask = (SEG == 32) ? 0xffffffffu : (((1u << SEG) - 1u) << (sg * SEG));
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nitems = (long long)ksh * n_chunks;
    for (long long item = warp0; item < nitems; item += nwarps) {
        const int k = (int)(item / n_chunks), ch = (int)(item - (long long)k * n_chunks);
        const int l0 = ch * chunk_len, l1 = min(kk, l0 + chunk_len);
        double g0 = 0.0, g1 = 0.0;
        for (int lb = l0; lb < l1; lb += SPW) {
            const int l = lb + sg;
            const bool live = l < l1;
            double a[NC];
            seg_gather<NC, SEG>(S, ld_s, rows + (size_t)k * n, cols + (size_t)(live ? l : l0) * n, n, sl, mask, a);
            double det = seg_lu_det<NC, SEG>(a, n, sl, mask);
            if (!live) det = 0.0;
#pragma unroll
            for (int s = 0; s < SPW; s++) {
                const double ds = __shfl_sync(0xffffffffu, det, s * SEG);
                const int ls = lb + s;
                if (ls < l1) {
                    if (lane < ldg) g0 = fma(ds, __ldg(ckt + (size_t)ls * ldg + lane), g0);
                    if (lane + 32 < ldg) g1 = fma(ds, __ldg(ckt + (size_t)ls * ldg + lane + 32), g1);
                }
            }
        }
        double *dst = gpart + ((size_t)ch * ksh + k) * ldg;
        if (lane < ldg) dst[lane] = g0;
        if (lane + 32 < ldg) dst[lane + 32] = g1;
    }
}

}  // namespace wfo
