// Small data-movement / prologue / epilogue kernels around the determinant kernels.
#pragma once
#include "common.cuh"
#include "lu_base.cuh"
#include "pair_update.cuh"

namespace wfo {

// (f,t) -> appended-order MO index list (pywfo/main.py:565-576); (-1,-1) -> ground state.
__global__ void exc_lists_kernel(const int32_t *__restrict__ exc, int k, int n, int32_t *__restrict__ lists) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)k * n) return;
    int e = (int)(idx / n), i = (int)(idx - (long long)e * n);
    int f = exc[2 * e], t = exc[2 * e + 1];
    int v;
    if (f < 0) v = i;
    else if (i == n - 1) v = n + t;
    else v = (i < f) ? i : i + 1;
    lists[idx] = v;
}

// CkT[l][j] = scale[l] * Ck[j][l] for j < n_ket, 0 for the padding columns.
__global__ void ckt_kernel(const double *__restrict__ Ck, int kk, int n_ket, const double *__restrict__ scale,
                           int ldg, double *__restrict__ ckt) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)kk * ldg) return;
    int j = (int)(idx / kk), l = (int)(idx - (long long)j * kk);   // l fastest: coalesced reads
    double v = 0.0;
    if (j < n_ket) v = Ck[(size_t)j * kk + l] * (scale ? scale[l] : 1.0);
    ckt[(size_t)l * ldg + j] = v;
}

// W buffer (v+1) x (v+1), ld = v+1: top-left = Dv block of S_MO, zero border
// (the border row/column serve ground-state entries of the excitation tables).
__global__ void w_init_kernel(const double *__restrict__ S, int ld_s, int n, int v, double *__restrict__ Wb) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int ldw = v + 1;
    if (idx >= (long long)ldw * ldw) return;
    int r = (int)(idx / ldw), c = (int)(idx - (long long)r * ldw);
    Wb[idx] = (r < v && c < v) ? S[(size_t)(n + r) * ld_s + n + c] : 0.0;
}

__device__ __forceinline__ double app_sign(int f, int n) { return ((n - 1 - f) & 1) ? -1.0 : 1.0; }

// bra side of the rank-update tier.  exc points at the first excitation of the shard.
// X is v x n.  dk0[k] = D(k, GS) = s_k det(A) X[t,f].
__global__ void t1_prep_bra_kernel(const int32_t *__restrict__ exc, int ksh, int n, int v,
                                   const double *__restrict__ X, const BaseInfo *__restrict__ info,
                                   BraMeta *__restrict__ meta, double *__restrict__ scale,
                                   double *__restrict__ dk0) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= ksh) return;
    int f = exc[2 * k], t = exc[2 * k + 1];
    const double detA = info->det;
    BraMeta m;
    double s;
    if (f < 0) { m.f = 0; m.woff = v * (v + 1); m.x = 1.0; s = 1.0; }
    else { m.f = f; m.woff = t * (v + 1); m.x = X[(size_t)t * n + f]; s = app_sign(f, n); }
    meta[k] = m;
    scale[k] = s * detA;
    dk0[k] = s * detA * m.x;
}

// ket side: Y is n x v.  d0l[l] = D(GS, l) = s_l det(A) Y[f',t'];  sgn[l] = s_l.
__global__ void t1_prep_ket_kernel(const int32_t *__restrict__ exc, int kk, int n, int v,
                                   const double *__restrict__ Y, const BaseInfo *__restrict__ info,
                                   KetMeta *__restrict__ meta, double *__restrict__ sgn,
                                   double *__restrict__ d0l) {
    int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= kk) return;
    int f = exc[2 * l], t = exc[2 * l + 1];
    const double detA = info->det;
    KetMeta m;
    double s;
    if (f < 0) { m.aoff = 0; m.tp = v; m.y = 1.0; s = 1.0; }
    else { m.aoff = f * n; m.tp = t; m.y = Y[(size_t)f * v + t]; s = app_sign(f, n); }
    meta[l] = m;
    sgn[l] = s;
    d0l[l] = s * detA * m.y;
}

// G[k][j] = sum_p gpart[p][k][j]  (fixed order);  column n_ket carries dk0[k].
__global__ void sum_parts_kernel(const double *__restrict__ gpart, int n_parts, int ksh, int ldg,
                                 int n_ket, const double *__restrict__ dk0, double *__restrict__ G) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long tot = (long long)ksh * ldg;
    if (idx >= tot) return;
    int j = (int)(idx % ldg);
    double s = 0.0;
    for (int p = 0; p < n_parts; p++) s += gpart[(size_t)p * tot + idx];
    if (j == n_ket) s = dk0[idx / ldg];
    G[idx] = s;
}

// wpart[z][j] = sum over slab z of Ck[j][l] * d0l[l]  (grid (n_ket, KGS_SLABS), fixed-shape tree)
constexpr int KGS_SLABS = 8;
__global__ void __launch_bounds__(256)
ket_gs_dot_kernel(const double *__restrict__ Ck, int kk, int n_ket, const double *__restrict__ d0l,
                  double *__restrict__ wpart) {
    __shared__ double red[256];
    const int j = blockIdx.x, z = blockIdx.y;
    const int per = (kk + KGS_SLABS - 1) / KGS_SLABS;
    const int lb = z * per, le = min(kk, lb + per);
    double s = 0.0;
    for (int l = lb + threadIdx.x; l < le; l += 256) s += Ck[(size_t)j * kk + l] * d0l[l];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) wpart[z * n_ket + j] = red[0];
}

// out[i][j] = d00 * sum_z P[z][i][j] + sigma * (sum_z P[z][i][n_ket]) * sum_z wpart[z][j]
__global__ void finalize_kernel(const double *__restrict__ P, int n_slabs, int n_bra, int ldg, int n_ket,
                                const double *__restrict__ wpart, const double *__restrict__ d00,
                                double sigma, double *__restrict__ out, int ld_out) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_bra * n_ket) return;
    int i = idx / n_ket, j = idx - i * n_ket;
    double a = 0.0, u = 0.0, wj = 0.0;
#pragma unroll
    for (int z = 0; z < KGS_SLABS; z++) wj += wpart[z * n_ket + j];
    size_t slab = (size_t)n_bra * ldg;
    for (int z = 0; z < n_slabs; z++) {
        a += P[z * slab + (size_t)i * ldg + j];
        u += P[z * slab + (size_t)i * ldg + n_ket];
    }
    out[(size_t)i * ld_out + j] = (*d00) * a + sigma * u * wj;
}

// ---- closed-form tier (T2) helpers ---------------------------------------------------------
// dense signed CI tensor: dense[s][f][t] += coeff[s][k] * (-1)^(n-1-f) for k in [k0, k1)
// (duplicate excitations, as pywfo.dets.get_dets produces them, add up).  flag is set when a
// ground-state entry (f < 0) is met: the closed form takes excitations only.
__global__ void scatter_signed_ci_kernel(const int32_t *__restrict__ exc, const double *__restrict__ coeff,
                                         int n_states, int K, int k0, int k1, int n, int v,
                                         double *__restrict__ dense, int *__restrict__ flag) {
    const int k = k0 + blockIdx.x * blockDim.x + threadIdx.x;
    const int s = blockIdx.y;
    if (k >= k1) return;
    const int f = exc[2 * k], t = exc[2 * k + 1];
    if (f < 0) { *flag = 1; return; }
    const double c = coeff[(size_t)s * K + k];
    if (c != 0.0) atomicAdd(dense + ((size_t)s * n + f) * v + t, c * app_sign(f, n));
}

// out[r] = sum_i M[r][i] * vec[i]   (one CTA per row, fixed-shape tree)
__global__ void __launch_bounds__(256)
rowdot_kernel(const double *__restrict__ M, long long len, const double *__restrict__ vec, double *__restrict__ out) {
    __shared__ double red[256];
    const double *row = M + (size_t)blockIdx.x * len;
    double s = 0.0;
    for (long long i = threadIdx.x; i < len; i += 256) s += row[i] * vec[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = red[0];
}

// out[i][j] = det^2 * ( (1+sigma) bx[i] ky[j] + sum_z P[z][i][j] )
__global__ void closed_finalize_kernel(const double *__restrict__ P, int n_slabs, int n_bra, int n_ket,
                                       const double *__restrict__ bx, const double *__restrict__ ky,
                                       const BaseInfo *__restrict__ info, double sigma, double *__restrict__ out,
                                       int ld_out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_bra * n_ket) return;
    const int i = idx / n_ket, j = idx - i * n_ket;
    double a = 0.0;
    for (int z = 0; z < n_slabs; z++) a += P[(size_t)z * n_bra * n_ket + idx];
    const double d = info->det;
    out[(size_t)i * ld_out + j] = d * d * ((1.0 + sigma) * bx[i] * ky[j] + a);
}

// ---- FP64 peak microbenchmarks (roofline denominators, bench.py) -------------
__global__ void __launch_bounds__(256) dfma_peak_kernel(int iters, double *out) {
    double a[16];
    double x = 1.0 + 1e-9 * threadIdx.x, y = 1e-9;
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = i * 0.5 + threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) a[i] = fma(a[i], x, y);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += a[i];
    if (s == 12345.678) out[0] = s;   // keep the loop alive
}

__global__ void __launch_bounds__(256) dmma884_peak_kernel(int iters, double *out) {
    double c[8][2];
    double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9 * threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; i++) c[i][0] = c[i][1] = i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}

// DMMA m16n8k8 (sm_90+ shape): A 16x8 (4 regs), B 8x8 (2 regs), C 16x8 (4 regs)
__global__ void __launch_bounds__(256) dmma1688_peak_kernel(int iters, double *out) {
    double c[4][4];
    double a0 = 1.0 + 1e-9 * threadIdx.x, a1 = a0 * 0.5, a2 = a0 * 0.25, a3 = a0 * 0.125;
    double b0 = 1e-9 * threadIdx.x, b1 = 2e-9;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) c[i][j] = i + j;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++)
            asm volatile(
                "mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
                "{%0,%1,%2,%3};\n"
                : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) s += c[i][j];
    if (s == 12345.678) out[0] = s;
}

}  // namespace wfo
