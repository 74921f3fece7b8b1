// Device-side excitation selection (replaces the host loops pywfo/main.py:539-540,
// 554-560, 589-611 and main.py:109-134): from the raw CI tensor ci[state][from][to]
// build the union excitation table and the per-state coefficient matrix with exact
// zeros where a state drops an excitation -- the form the determinant kernels consume.
//
//   flags   : union over states of |c| >= thr (">" when strict), one per (from, to)
//   scan    : exclusive prefix sum in traversal order (single CTA, n_occ*n_virt flags)
//   emit    : exc[k] = (from, to);  coeff[s][k] = keep(s) ? c * factor(from) : 0
//
// Traversal order 0 = row major (from, to) like np.nonzero; order 1 = "locality"
// (to/32, from, to%32), the order the rank-update kernel wants for its ket list
// (a CTA's window of W stays within 32 columns at a time).  factor(from) is 1 for
// the singlet family and ((-1)^(occ-from+1))^occ for pywfo.main.overlaps
// (main.py:154, 176: det(sign*M) = sign^occ det(M)).
#pragma once
#include "common.cuh"

namespace wfo {

__device__ __forceinline__ void pos_to_ft(int pos, int n, int v, int order, int &f, int &t) {
    if (order == 0) {
        f = pos / v;
        t = pos - f * v;
    } else {
        const int nfull = v >> 5;
        const int per = n << 5;
        int b = pos / per;
        if (b > nfull) b = nfull;
        const int q = pos - b * per;
        const int wb = (b < nfull) ? 32 : (v & 31);
        f = q / wb;
        t = (b << 5) + (q - f * wb);
    }
}

__device__ __forceinline__ bool ci_keep(double c, double thr, int strict) {
    const double a = fabs(c);
    return strict ? (a > thr) : (a >= thr);
}

// A WINDOW of the (from, to) space can be processed on its own (one rank's slice of the bra side):
// positions [pos0, pos0 + cnt) in traversal order 0, `ci` then points at the window's first element
// of state 0 and `stride` is the distance between states (cnt for a compact slice, n*v in place).
// The whole tensor is pos0 = 0, cnt = n*v, stride = n*v.
__global__ void union_flags_kernel(const double *__restrict__ ci, int n_states, int n, int v, double thr,
                                   int strict, int order, int *__restrict__ flags, int pos0, int cnt, long long stride) {
    const int local = blockIdx.x * blockDim.x + threadIdx.x;
    if (local >= cnt) return;
    int f, t;
    pos_to_ft(pos0 + local, n, v, order, f, t);
    const long long off = (long long)f * v + t - pos0;
    int any = 0;
    for (int s = 0; s < n_states; s++) any |= ci_keep(ci[s * stride + off], thr, strict) ? 1 : 0;
    flags[local] = any;
}

// kept[s] = number of excitations state s keeps (main.overlaps raises on 0, main.py:186)
__global__ void __launch_bounds__(256)
state_counts_kernel(const double *__restrict__ ci, int n, int v, double thr, int strict, int *__restrict__ kept) {
    __shared__ int red[256];
    const int s = blockIdx.x;
    const size_t stride = (size_t)n * v;
    int cnt = 0;
    for (size_t i = threadIdx.x; i < stride; i += 256) cnt += ci_keep(ci[s * stride + i], thr, strict) ? 1 : 0;
    red[threadIdx.x] = cnt;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) kept[s] = red[0];
}

// exclusive scan of `count` flags by one CTA of 1024 threads; total -> *total_out.
// Tiles of 4096 flags: every thread takes four consecutive flags (coalesced across the CTA), the tile is
// scanned with shuffles (warp, then the 32 warp totals) and a running carry; the next tile's loads are
// issued before the current one is scanned.  (Round 1's version gave every thread one long strided run:
// 40 us for the 40 000 flags of config C4, all of it load latency.)
__global__ void __launch_bounds__(1024)
scan_flags_kernel(const int *__restrict__ flags, int count, int *__restrict__ offs, int *__restrict__ total_out) {
    __shared__ int wsum[32];
    __shared__ int carry_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    auto load4 = [&](int base, int (&f)[4]) {
#pragma unroll
        for (int e = 0; e < 4; e++) f[e] = (base + e < count) ? flags[base + e] : 0;
    };
    int carry = 0;
    int cur[4], nxt[4];
    load4(4 * tid, cur);
    for (int t0 = 0; t0 < count; t0 += 4096) {
        const int base = t0 + 4 * tid;
        load4(base + 4096, nxt);   // next tile in flight
        const int s1 = cur[0], s2 = s1 + cur[1], s3 = s2 + cur[2], mine = s3 + cur[3];
        int inc = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int w = wsum[lane], winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += v;
            }
            wsum[lane] = winc - w;            // exclusive prefix of the warp totals
            if (lane == 31) carry_s = winc;   // tile total
        }
        __syncthreads();
        const int ex = carry + wsum[warp] + inc - mine;   // exclusive prefix of this thread's first flag
        if (base < count) offs[base] = ex;
        if (base + 1 < count) offs[base + 1] = ex + s1;
        if (base + 2 < count) offs[base + 2] = ex + s2;
        if (base + 3 < count) offs[base + 3] = ex + s3;
        carry += carry_s;
        __syncthreads();   // wsum / carry_s are rewritten by the next tile
#pragma unroll
        for (int e = 0; e < 4; e++) cur[e] = nxt[e];
    }
    if (tid == 0) *total_out = carry;
}

__global__ void emit_exc_kernel(const int *__restrict__ flags, const int *__restrict__ offs, int n, int v,
                                int order, int32_t *__restrict__ exc, int pos0, int cnt) {
    const int local = blockIdx.x * blockDim.x + threadIdx.x;
    if (local >= cnt || !flags[local]) return;
    int f, t;
    pos_to_ft(pos0 + local, n, v, order, f, t);
    exc[2 * offs[local]] = f;
    exc[2 * offs[local] + 1] = t;
}

// coeff is n_states x K row-major.  grid.y = state.
__global__ void emit_coeff_kernel(const double *__restrict__ ci, const int *__restrict__ flags,
                                  const int *__restrict__ offs, int n, int v, double thr, int strict,
                                  int order, int minus_sign, int K, double *__restrict__ coeff, int pos0, int cnt,
                                  long long stride) {
    const int local = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = blockIdx.y;
    if (local >= cnt || !flags[local]) return;
    int f, t;
    pos_to_ft(pos0 + local, n, v, order, f, t);
    const double c = ci[s * stride + (long long)f * v + t - pos0];
    double val = ci_keep(c, thr, strict) ? c : 0.0;
    if (minus_sign && (n & 1) && ((n - f + 1) & 1)) val = -val;   // ((-1)^(n-f+1))^n
    coeff[(size_t)s * K + offs[local]] = val;
}

}  // namespace wfo
