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

__device__ __forceinline__ double app_sign(int f, int n) { return ((n - 1 - f) & 1) ? -1.0 : 1.0; }

// CkT[l][j] = s_l * Ck[j][l] for j < n_ket, 0 for the padding columns; s_l = (-1)^(n-1-f'_l) when an
// excitation table is given (rank-update tier), 1 otherwise.  Tiled through shared memory: both
// the reads (rows of Ck) and the writes (a contiguous block of CkT rows) are coalesced.
constexpr int CKT_TL = 64;   // ket excitations per CTA
__global__ void __launch_bounds__(256)
ckt_kernel(const double *__restrict__ Ck, int kk, int kk_pad, int n_ket, const int32_t *__restrict__ exc, int n,
           int ldg, double *__restrict__ ckt) {
    extern __shared__ double ckt_sm[];   // CKT_TL x (ldg + 1)
    const int l0 = blockIdx.x * CKT_TL;
    const int rows = min(CKT_TL, kk_pad - l0);   // rows in [kk, kk_pad) are written as zeros
    const int lds = ldg + 1;
    for (int idx = threadIdx.x; idx < ldg * CKT_TL; idx += 256) {
        const int j = idx / CKT_TL, r = idx - j * CKT_TL;
        double v = 0.0;
        if (j < n_ket && l0 + r < kk) {
            v = Ck[(size_t)j * kk + l0 + r];
            if (exc) {
                const int f = exc[2 * (l0 + r)];
                if (f >= 0) v *= app_sign(f, n);
            }
        }
        ckt_sm[r * lds + j] = v;
    }
    __syncthreads();
    double *dst = ckt + (size_t)l0 * ldg;
    for (int idx = threadIdx.x; idx < rows * ldg; idx += 256) {
        const int r = idx / ldg, j = idx - r * ldg;
        dst[idx] = ckt_sm[r * lds + j];
    }
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

// both sides in one launch: the first cdiv(ksh+1,256) CTAs do the bra table, the rest the ket table.
// Bra row ksh is a virtual ground-state row (x = 1, zero row of W): its pair values are y_l, so its
// row of G is w_j = sum_l D(0,l) Ck[j][l], the ket-side factor of the cross term, for free.
__global__ void __launch_bounds__(256)
t1_prep_kernel(const int32_t *__restrict__ bra_exc, int ksh, const int32_t *__restrict__ ket_exc, int kk, int kk_pad,
               int n, int v,
               const double *__restrict__ X, const double *__restrict__ Y, const BaseInfo *__restrict__ info,
               BraMeta *__restrict__ bmeta, double *__restrict__ bscale, double *__restrict__ dk0,
               KetMeta *__restrict__ kmeta, double *__restrict__ d0l) {
    const int nb_bra = (ksh + 1 + 255) / 256;   // + the virtual ground-state row k = ksh
    const double detA = info->det;
    if ((int)blockIdx.x < nb_bra) {
        const int k = blockIdx.x * 256 + threadIdx.x;
        if (k > ksh) return;
        const int f = (k < ksh) ? bra_exc[2 * k] : -1, t = (k < ksh) ? bra_exc[2 * k + 1] : -1;
        BraMeta m;
        double s;
        if (f < 0) { m.f = 0; m.woff = v * (v + 1); m.x = 1.0; s = 1.0; }
        else { m.f = f; m.woff = t * (v + 1); m.x = X[(size_t)t * n + f]; s = app_sign(f, n); }
        bmeta[k] = m;
        bscale[k] = s * detA;
        dk0[k] = s * detA * m.x;
    } else {
        const int l = (blockIdx.x - nb_bra) * 256 + threadIdx.x;
        if (l >= kk_pad) return;
        if (l >= kk) {   // padding of the last ket tile: a finite pair value times zero coefficients
            KetMeta z;
            z.aoff = 0; z.tp = 0; z.y = 0.0;
            kmeta[l] = z;
            return;
        }
        const int f = ket_exc[2 * l], t = ket_exc[2 * l + 1];
        KetMeta m;
        double s;
        if (f < 0) { m.aoff = 0; m.tp = v; m.y = 1.0; s = 1.0; }
        else { m.aoff = f * n; m.tp = t; m.y = Y[(size_t)f * v + t]; s = app_sign(f, n); }
        kmeta[l] = m;
        d0l[l] = s * detA * m.y;
    }
}

// ---- epilogue, two launches ----------------------------------------------------------------------
// (1) reduce_contract: CTA (z, nt) owns RC_ROWS rows of G and one 8-column tile.  It adds the
//     partial sums of those elements (fixed order; column n_ket is replaced by D(k,0)) into a shared
//     slab and contracts it with the bra coefficients on the FP64 tensor cores:
//     P[z][:, tile] = Cb[:, rows of z] . G_slab (every coefficient is used exactly once per CTA, so
//     the A fragments come straight from global memory).  Row `gs_row` of G (rank-update tier: the
//     virtual ground-state bra row, whose G is w_j = sum_l D(0,l) Ck[j][l]) is copied to wvec instead.
// (2) finalize: CTA i adds the slabs (fixed order) and writes
//     out[i][j] = D00 * sum_z P[z][i][j] + sigma * (sum_z P[z][i][n_ket]) * w_j.
constexpr int RC_ROWS = 128;

__global__ void __launch_bounds__(256)
reduce_contract_kernel(const double *__restrict__ gpart, PartMap pm, int ksh_g, int ksh, int ldg, int n_ket,
                       const double *__restrict__ dk0, const double *__restrict__ Cb, int ld_cb, int n_bra,
                       double *__restrict__ P, double *__restrict__ wvec, int gs_row) {
    __shared__ double Gs[RC_ROWS * 8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int k0 = blockIdx.x * RC_ROWS, j0 = blockIdx.y * 8;
    const size_t part_stride = (size_t)ksh_g * ldg;
#pragma unroll
    for (int e = 0; e < RC_ROWS * 8 / 256; e++) {
        const int idx = tid + e * 256;
        const int r = idx >> 3, j = j0 + (idx & 7);
        const int k = k0 + r;
        double v = 0.0;
        if (k < ksh_g) {
            if (j == n_ket) {
                v = (k < ksh) ? dk0[k] : 0.0;
            } else {
                const int np = parts_of_row(pm, k);
                const double *src = gpart + (size_t)k * ldg + j;
                for (int p = 0; p < np; p++) v += src[p * part_stride];
                if (k == gs_row && j < n_ket) wvec[j] = v;
            }
        }
        Gs[idx] = v;
    }
    __syncthreads();
    double *Pz = P + (size_t)blockIdx.x * n_bra * ldg;
    const int kmax = ksh - k0;   // rows r >= kmax carry no bra coefficient (padding / the ground-state row)
    for (int mt = warp; mt * 8 < n_bra; mt += 8) {
        const int i = 8 * mt + g;
        const double *arow = Cb + (size_t)(i < n_bra ? i : 0) * ld_cb + k0 + t;
        double c0 = 0.0, c1 = 0.0;
#pragma unroll 8
        for (int ks = 0; ks < RC_ROWS / 4; ks++) {
            const double a = (i < n_bra && 4 * ks + t < kmax) ? __ldg(arow + 4 * ks) : 0.0;
            dmma884(c0, c1, a, Gs[(4 * ks + t) * 8 + g]);
        }
        if (i < n_bra) *reinterpret_cast<double2 *>(Pz + (size_t)i * ldg + j0 + 2 * t) = make_double2(c0, c1);
    }
}

__global__ void __launch_bounds__(256)
finalize3_kernel(const double *__restrict__ P, int n_slabs, int n_bra, int ldg, int n_ket,
                 const double *__restrict__ wparts, int n_wparts, const double *__restrict__ d00,
                 double sigma, double *__restrict__ out, int ld_out) {
    __shared__ double ra[4][64];
    __shared__ double tot[64];
    const int i = blockIdx.x, tid = threadIdx.x;
    const int j = tid & 63, q = tid >> 6;
    const size_t slab = (size_t)n_bra * ldg;
    double a = 0.0;
    if (j < ldg)
        for (int z = q; z < n_slabs; z += 4) a += P[z * slab + (size_t)i * ldg + j];
    ra[q][j] = a;
    __syncthreads();
    if (q == 0) tot[j] = (ra[0][j] + ra[1][j]) + (ra[2][j] + ra[3][j]);
    __syncthreads();
    if (q == 0 && j < n_ket) {
        double w = 0.0;
        for (int p = 0; p < n_wparts; p++) w += wparts[p * ldg + j];
        out[(size_t)i * ld_out + j] = (*d00) * tot[j] + sigma * tot[n_ket] * w;
    }
}

// wpart[z][j] = sum over slab z of Ck[j][l] * d0l[l]  (grid (n_ket, KGS_SLABS), fixed-shape tree)
constexpr int KGS_SLABS = 8;
__global__ void __launch_bounds__(256)
ket_gs_dot_kernel(const double *__restrict__ Ck, int kk, int n_ket, int ldw, const double *__restrict__ d0l,
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
    if (threadIdx.x == 0) wpart[z * ldw + j] = red[0];
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

// ---- Newton-Schulz inverse helpers (device-side replacement of np.linalg.inv, pywfo/main.py:303) ----
__device__ __forceinline__ void atomic_max_nonneg(double *slot, double v) {   // v >= 0: IEEE order = integer order
    atomicMax(reinterpret_cast<unsigned long long *>(slot), (unsigned long long)__double_as_longlong(v));
}
// slots[0] = max row sum of |A| (inf-norm), slots[1] = max column sum (1-norm); grid = n CTAs
__global__ void __launch_bounds__(256) ns_norms_kernel(const double *__restrict__ A, int n, double *__restrict__ slots) {
    __shared__ double red[2][256];
    const int r = blockIdx.x;
    double rs = 0.0, cs = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) {
        rs += fabs(A[(size_t)r * n + i]);
        cs += fabs(A[(size_t)i * n + r]);
    }
    red[0][threadIdx.x] = rs;
    red[1][threadIdx.x] = cs;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if (threadIdx.x < o) {
            red[0][threadIdx.x] += red[0][threadIdx.x + o];
            red[1][threadIdx.x] += red[1][threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        atomic_max_nonneg(slots, red[0][0]);
        atomic_max_nonneg(slots + 1, red[1][0]);
    }
}
// X0 = A^T / slots[0]   (slots[0] = |A^T A|_inf >= sigma_max(A)^2)
__global__ void ns_init_kernel(const double *__restrict__ A, int n, const double *__restrict__ slots, double *__restrict__ X) {
    __shared__ double tile[32][33];
    const double sc = 1.0 / slots[0];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
        const int r = by + dy, c = bx + threadIdx.x;
        tile[dy][threadIdx.x] = (r < n && c < n) ? A[(size_t)r * n + c] : 0.0;
    }
    __syncthreads();
    for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
        const int r = bx + dy, c = by + threadIdx.x;
        if (r < n && c < n) X[(size_t)r * n + c] = tile[threadIdx.x][dy] * sc;
    }
}
// slot = max |T - I|;  Xn = X  (the next iterate starts as a copy: Xn = 2 Xn - X T follows as a GEMM)
__global__ void ns_residual_kernel(const double *__restrict__ T, const double *__restrict__ X, int n,
                                   double *__restrict__ slot, double *__restrict__ Xn) {
    const size_t tot = (size_t)n * n;
    double m = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / n), c = (int)(i - (size_t)r * n);
        const double d = fabs(T[i] - (r == c ? 1.0 : 0.0));
        m = (d > m || d != d) ? d : m;   // NaN propagates
        Xn[i] = X[i];
    }
    for (int o = 16; o; o >>= 1) {
        const double x = __shfl_xor_sync(0xffffffffu, m, o);
        m = (x > m || x != x) ? x : m;
    }
    if ((threadIdx.x & 31) == 0) {
        if (m != m) m = 1e300;
        atomic_max_nonneg(slot, m);
    }
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
