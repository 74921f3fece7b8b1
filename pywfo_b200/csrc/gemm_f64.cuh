// K1: FP64 tensor-core GEMM  C = alpha * op(A) op(B) + beta * C  (row-major).
//
// Used for S_AO = Cinv Cinv^T and S_MO = C_bra S_AO C_ket^T (pywfo/main.py:304, 532),
// for the base-block products X = C Ai, Y = Ai B, W = Dv - C Y of the rank-update
// tier, and (with split-K slabs) for the final CI contraction S = Cb G.
//
// CTA tile GM x GN (64x64, or 32x64 / 32x32 when that is needed to fill the 148 SMs), 4 warps
// in a 2x2 grid, warp tile (GM/2) x (GN/2) = (GM/16) x (GN/16) DMMA.8x8x4 tiles, K tile 32.  Operand tiles
// are brought in with cp.async (16-byte copies when the operands allow it, else 8-byte elements;
// zero fill at the edges) through a 3-stage shared-memory ring, so the L2 latency of tile t+2 hides behind the DMMAs of tile t with one
// block barrier per K tile.  Padded shared tiles make both fragment loads conflict-free
// (k-contiguous tiles: stride 36 doubles, n-contiguous B tile: stride GN + 8 doubles).  blockIdx.z selects a K slab (split-K): slab z
// covers k in [z*kslab, min(k,(z+1)*kslab)) and writes C + z*slab_stride; deterministic,
// summed by a later kernel in fixed order.  With batch strides blockIdx.z is a batch index.
#pragma once
#include "common.cuh"

namespace wfo {

constexpr int GK = 32;            // GM (64 / 32 rows) and GN (64 / 32 columns per CTA) are template parameters
constexpr int GLDA = GK + 4;      // 36: row stride of k-contiguous tiles (A, and B when it is stored transposed)
constexpr int GSTAGES = 3;
template <int GN> constexpr int gemm_ldb() { return GN + 8; }        // 72 / 40: row stride of the n-contiguous B tile
template <int GN> constexpr int gemm_bs() { return GN * GLDA > GK * (GN + 8) ? GN * GLDA : GK * (GN + 8); }   // B tile of a stage, either layout

template <int GM, int GN>
constexpr size_t gemm_smem_bytes() { return (size_t)GSTAGES * (GM * GLDA + gemm_bs<GN>()) * sizeof(double); }

// async copies global -> shared; a source size below the copy size zero-fills the rest
__device__ __forceinline__ void cp_async8(double *dst, const double *src, bool ok) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    const int bytes = ok ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(src), "r"(bytes));
}
__device__ __forceinline__ void cp_async16(double *dst, const double *src, int bytes) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(src), "r"(bytes));
}

// VEC: 16-byte copies (needs !TA, even leading dimensions and K-slab starts, 16-byte aligned bases)
template <bool TA, bool TB, int GM, int GN, bool VEC>
__global__ void __launch_bounds__(128)
gemm_f64_kernel(int m, int n, int k, double alpha, const double *__restrict__ A, int lda,
                const double *__restrict__ B, int ldb, double beta, double *__restrict__ C, int ldc,
                int kslab, long long slab_stride, long long batch_a, long long batch_b) {
    extern __shared__ __align__(16) double gsm[];
    constexpr int GLDB = GN + 8;
    constexpr int STAGE = GM * GLDA + (GN * GLDA > GK * (GN + 8) ? GN * GLDA : GK * (GN + 8));   // doubles per stage
    constexpr int NTW = GN / 16;                       // 8-column tiles per warp

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    constexpr int MT = GM / 16;               // 8-row tiles per warp (warps are 2 x 2)
    const int wm = (warp >> 1) * (GM / 2), wn = (warp & 1) * (GN / 2);
    const int m0 = blockIdx.y * GM, n0 = blockIdx.x * GN;
    // blockIdx.z is either a K slab (batch strides 0) or a batch index (kslab >= k, batch strides set)
    const bool batched = (batch_a | batch_b) != 0;
    const int kbeg = batched ? 0 : blockIdx.z * kslab;
    const int kend = min(k, kbeg + kslab);
    A += (long long)blockIdx.z * batch_a;
    B += (long long)blockIdx.z * batch_b;
    C += (long long)blockIdx.z * slab_stride;

    double acc[MT][NTW][2];
#pragma unroll
    for (int i = 0; i < MT; i++)
#pragma unroll
        for (int j = 0; j < NTW; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

    // shared layouts: As[mm][kk] (stride GLDA); Bs[kb][nn] (stride GLDB) for B as stored k x n,
    // Bs[nn][kb] (stride GLDA) for B stored n x k (TB)
    auto issue_tile = [&](int stage, int k0) {
        double *As = gsm + stage * STAGE;
        double *Bs = As + GM * GLDA;
        if (VEC) {
#pragma unroll
            for (int e = 0; e < GM * (GK / 2) / 128; e++) {
                const int idx = tid + e * 128;
                const int mm = idx / (GK / 2), kc = (idx % (GK / 2)) * 2;
                const int gm = m0 + mm, gk = k0 + kc;
                const int bytes = (gm < m) ? max(0, min(16, (kend - gk) * 8)) : 0;
                cp_async16(As + mm * GLDA + kc, bytes ? A + (long long)gm * lda + gk : A, bytes);
            }
#pragma unroll
            for (int e = 0; e < GN * (GK / 2) / 128; e++) {
                const int idx = tid + e * 128;
                if (TB) {
                    const int nn = idx / (GK / 2), kc = (idx % (GK / 2)) * 2;
                    const int gn = n0 + nn, gk = k0 + kc;
                    const int bytes = (gn < n) ? max(0, min(16, (kend - gk) * 8)) : 0;
                    cp_async16(Bs + nn * GLDA + kc, bytes ? B + (long long)gn * ldb + gk : B, bytes);
                } else {
                    const int kb = idx / (GN / 2), nc = (idx % (GN / 2)) * 2;
                    const int gkb = k0 + kb, gn = n0 + nc;
                    const int bytes = (gkb < kend) ? max(0, min(16, (n - gn) * 8)) : 0;
                    cp_async16(Bs + kb * GLDB + nc, bytes ? B + (long long)gkb * ldb + gn : B, bytes);
                }
            }
        } else {
#pragma unroll
            for (int e = 0; e < GM * GK / 128; e++) {
                const int idx = tid + e * 128;
                int mm, kk;
                if (TA) { kk = idx / GM; mm = idx % GM; } else { mm = idx / GK; kk = idx % GK; }
                const int gm = m0 + mm, gk = k0 + kk;
                const bool ok = gm < m && gk < kend;
                const double *src = ok ? (TA ? A + (long long)gk * lda + gm : A + (long long)gm * lda + gk) : A;
                cp_async8(As + mm * GLDA + kk, src, ok);
            }
#pragma unroll
            for (int e = 0; e < GN * GK / 128; e++) {
                const int idx = tid + e * 128;
                int nn, kb;
                if (TB) { nn = idx / GK; kb = idx % GK; } else { kb = idx / GN; nn = idx % GN; }
                const int gn = n0 + nn, gkb = k0 + kb;
                const bool ok = gn < n && gkb < kend;
                const double *src = ok ? (TB ? B + (long long)gn * ldb + gkb : B + (long long)gkb * ldb + gn) : B;
                cp_async8(TB ? Bs + nn * GLDA + kb : Bs + kb * GLDB + nn, src, ok);
            }
        }
    };

    const int ntiles = (kend > kbeg) ? (kend - kbeg + GK - 1) / GK : 0;
    // prologue: tiles 0 and 1 in flight (one commit group per tile, empty groups keep the count)
#pragma unroll
    for (int s = 0; s < GSTAGES - 1; s++) {
        if (s < ntiles) issue_tile(s, kbeg + s * GK);
        asm volatile("cp.async.commit_group;\n");
    }
    for (int it = 0; it < ntiles; it++) {
        asm volatile("cp.async.wait_group %0;\n" ::"n"(GSTAGES - 2));   // tile `it` has landed
        __syncthreads();                                                  // ... for everyone; tile it-1 is consumed
        if (it + GSTAGES - 1 < ntiles) issue_tile((it + GSTAGES - 1) % GSTAGES, kbeg + (it + GSTAGES - 1) * GK);
        asm volatile("cp.async.commit_group;\n");
        const double *As = gsm + (it % GSTAGES) * STAGE;
        const double *Bs = As + GM * GLDA;
#pragma unroll
        for (int ks = 0; ks < GK; ks += 4) {
            double a[MT], b[NTW];
#pragma unroll
            for (int i = 0; i < MT; i++) a[i] = As[(wm + 8 * i + g) * GLDA + ks + t];
#pragma unroll
            for (int j = 0; j < NTW; j++)
                b[j] = TB ? Bs[(wn + 8 * j + g) * GLDA + ks + t] : Bs[(ks + t) * GLDB + wn + 8 * j + g];
#pragma unroll
            for (int i = 0; i < MT; i++)
#pragma unroll
                for (int j = 0; j < NTW; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    asm volatile("cp.async.wait_group 0;\n");

#pragma unroll
    for (int i = 0; i < MT; i++) {
        int row = m0 + wm + 8 * i + g;
        if (row >= m) continue;
#pragma unroll
        for (int j = 0; j < NTW; j++) {
            int col = n0 + wn + 8 * j + 2 * t;
#pragma unroll
            for (int c = 0; c < 2; c++) {
                if (col + c < n) {
                    double *p = C + (long long)row * ldc + col + c;
                    double v = alpha * acc[i][j][c];
                    if (beta != 0.0) v += beta * (*p);
                    *p = v;
                }
            }
        }
    }
}

}  // namespace wfo
