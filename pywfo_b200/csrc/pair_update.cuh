// T1 (rank-update tier), K3+K4 fused: every pair determinant from ONE base LU.
//
// With A = S_MO[:n,:n], Ai = A^-1, X = C Ai, Y = Ai B, W = Dv - C Ai B
// (B, C, Dv the occ-virt, virt-occ, virt-virt blocks of S_MO) the determinant of
// the minor with bra excitation k=(f,t) and ket excitation l=(f',t') is, by a
// rank-1 row replacement followed by a rank-1 column replacement of A
// (2x2 capacitance determinant):
//     D(k,l) = s_k s_l det(A) * ( X[t,f] Y[f',t'] + Ai[f',f] W[t,t'] )
// s = (-1)^(n-1-f) converts to the reference's appended ordering
// (pywfo/main.py:565-576).  It replaces np.linalg.det of the gathered minor
// (pywfo/main.py:583-584) for single excitations.
//
// The kernel forms each pair value in registers, directly in the A-fragment
// layout of DMMA.8x8x4, and contracts it on the FP64 tensor cores with the ket
// coefficient tile:  G[k][j] = s_k det(A) * sum_l d(k,l) * CkTs[l][j]
// (CkTs = s_l * Ck^T, zero padded to ldg columns).  Pair values never reach HBM
// (pywfo/main.py:613-626 is the loop this replaces).
//
// CTA = 4 warps, 96 bra excitations (3 m-tiles of 8 per warp, 168 registers, 3 CTAs/SM);
// persistent CTAs fetch (k tile, l split) work items from an atomic counter; the ket
// coefficient tiles (32 excitations x ldg, one contiguous block of CkTs) are staged in shared
// memory by TMA bulk copies (cp.async.bulk + mbarrier, SASS UBLKCP), double buffered, and
// shared by the 4 warps.  Ai and W are gathered
// through L1 (the host orders the ket excitations so that a CTA's working set
// of W is a few KB at a time, see pywfo_b200/excitations.py).
#pragma once
#include "common.cuh"

namespace wfo {

#ifndef WFO_T1_WARPS
#define WFO_T1_WARPS 4
#endif
#ifndef WFO_T1_MT
#define WFO_T1_MT 3
#endif
#ifndef WFO_T1_MINB
#define WFO_T1_MINB 3
#endif
constexpr int T1_WARPS = WFO_T1_WARPS;
constexpr int T1_MT = WFO_T1_MT;                 // m-tiles (8 bra excitations) per warp
constexpr int T1_KTILE = T1_WARPS * T1_MT * 8;   // 128 bra excitations per CTA
#ifndef WFO_T1_LT
#define WFO_T1_LT 32
#endif
constexpr int T1_LT = WFO_T1_LT;                 // ket excitations per shared tile

struct KetMeta {   // one per ket excitation (16 bytes)
    int aoff;      // f' * n      (row offset into Ai)
    int tp;        // t'          (column of W)
    double y;      // Y[f', t']
};
struct BraMeta {   // one per bra excitation (16 bytes)
    int f;         // column of Ai
    int woff;      // t * v       (row offset into W)
    double x;      // X[t, f]
};

// ---- TMA (bulk async copy) + mbarrier helpers: one elected thread moves a whole ket tile
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WFO_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WFO_DONE_%=;\n"
        "bra WFO_WAIT_%=;\n"
        "WFO_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// NT = ldg / 8 column tiles of G.  Work item = (k tile, l split), fetched dynamically.
// gpart[(split*ksh + k)*ldg + j]
template <int NT>
__global__ void __launch_bounds__(T1_WARPS * 32, WFO_T1_MINB)
t1_fused_kernel(const double *__restrict__ Ai, const double *__restrict__ W,
                const BraMeta *__restrict__ bra, const double *__restrict__ bra_scale, int ksh,
                const KetMeta *__restrict__ ket, const double *__restrict__ ckts, int kk,
                int l_per_split, int n_ktiles, int n_items, int *__restrict__ counter,
                double *__restrict__ gpart) {
    constexpr int LDG = NT * 8;
    __shared__ __align__(128) double Bs[2][T1_LT * LDG];
    __shared__ __align__(16) KetMeta Ms[2][T1_LT];
    __shared__ __align__(8) unsigned long long full_bar[2];   // "tile landed" barriers of the two stages
    __shared__ int s_item;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    if (tid == 0) {
        mbar_init(&full_bar[0], 1);
        mbar_init(&full_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    unsigned phase0 = 0, phase1 = 0;   // parity of the next completion of each stage barrier
    __syncthreads();
  for (;;) {
    if (tid == 0) s_item = atomicAdd(counter, 1);
    __syncthreads();
    const int item = s_item;
    __syncthreads();
    if (item >= n_items) break;
    const int split = item / n_ktiles, ktile = item - split * n_ktiles;
    const int k0 = ktile * T1_KTILE + warp * (T1_MT * 8);
    const int l_begin = split * l_per_split;
    const int l_end = min(kk, l_begin + l_per_split);

    // per-thread bra data for its 4 rows (one per m-tile)
    int bf[T1_MT], bw[T1_MT];
    double bx[T1_MT];
#pragma unroll
    for (int mi = 0; mi < T1_MT; mi++) {
        int k = k0 + 8 * mi + g;
        if (k < ksh) {
            BraMeta m = bra[k];
            bf[mi] = m.f; bw[mi] = m.woff; bx[mi] = m.x;
        } else {
            bf[mi] = 0; bw[mi] = 0; bx[mi] = 0.0;   // harmless duplicates, never stored
        }
    }
    double acc[T1_MT][NT][2];
#pragma unroll
    for (int mi = 0; mi < T1_MT; mi++)
#pragma unroll
        for (int ni = 0; ni < NT; ni++) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

    // stage one tile of ket data with TMA: the coefficient rows of T1_LT consecutive ket
    // excitations are one contiguous block of CkTs (T1_LT * LDG * 8 bytes), their metadata another;
    // one thread issues two bulk copies that complete on the stage's mbarrier.  Rows beyond the
    // end of the split are zero-filled by the other threads (different addresses).
    auto stage = [&](int buf, int l0) {
        const int rows = min(T1_LT, l_end - l0);
        if (tid == 0) {
            mbar_expect_tx(&full_bar[buf], (unsigned)(rows * (LDG * 8 + 16)));
            tma_bulk_g2s(&Bs[buf][0], ckts + (size_t)l0 * LDG, (unsigned)(rows * LDG * 8), &full_bar[buf]);
            tma_bulk_g2s(&Ms[buf][0], ket + l0, (unsigned)(rows * 16), &full_bar[buf]);
        }
        if (rows < T1_LT) {
            for (int c = rows * LDG + tid; c < T1_LT * LDG; c += T1_WARPS * 32) Bs[buf][c] = 0.0;
            for (int r = rows + tid; r < T1_LT; r += T1_WARPS * 32) { Ms[buf][r].aoff = 0; Ms[buf][r].tp = 0; Ms[buf][r].y = 0.0; }
        }
    };

    int buf = 0;
    if (l_begin < l_end) stage(0, l_begin);
    for (int l0 = l_begin; l0 < l_end; l0 += T1_LT) {
        if (l0 + T1_LT < l_end) stage(buf ^ 1, l0 + T1_LT);
        if (buf == 0) { mbar_wait(&full_bar[0], phase0); phase0 ^= 1; }
        else { mbar_wait(&full_bar[1], phase1); phase1 ^= 1; }
        __syncthreads();   // also publishes the zero fill of a partial tile
        const double *bs = &Bs[buf][0];
        const KetMeta *ms = &Ms[buf][0];
        // software pipeline over the 8 k-steps (4 ket excitations each): gather one step ahead
        double a_cur[T1_MT], ai_n[T1_MT], w_n[T1_MT], y_n;
        {
            KetMeta m = ms[t];
            y_n = m.y;
#pragma unroll
            for (int mi = 0; mi < T1_MT; mi++) {
                ai_n[mi] = __ldg(Ai + m.aoff + bf[mi]);
                w_n[mi] = __ldg(W + bw[mi] + m.tp);
            }
        }
#pragma unroll
        for (int ks = 0; ks < T1_LT / 4; ks++) {
#pragma unroll
            for (int mi = 0; mi < T1_MT; mi++) a_cur[mi] = fma(bx[mi], y_n, ai_n[mi] * w_n[mi]);
            if (ks + 1 < T1_LT / 4) {
                KetMeta m = ms[4 * (ks + 1) + t];
                y_n = m.y;
#pragma unroll
                for (int mi = 0; mi < T1_MT; mi++) {
                    ai_n[mi] = __ldg(Ai + m.aoff + bf[mi]);
                    w_n[mi] = __ldg(W + bw[mi] + m.tp);
                }
            }
            double b[NT];
#pragma unroll
            for (int ni = 0; ni < NT; ni++) b[ni] = bs[(4 * ks + t) * LDG + 8 * ni + g];
#pragma unroll
            for (int mi = 0; mi < T1_MT; mi++)
#pragma unroll
                for (int ni = 0; ni < NT; ni++) dmma884(acc[mi][ni][0], acc[mi][ni][1], a_cur[mi], b[ni]);
        }
        __syncthreads();
        buf ^= 1;
    }

    // epilogue: row scale s_k * det(A) and store this split's partial
#pragma unroll
    for (int mi = 0; mi < T1_MT; mi++) {
        int k = k0 + 8 * mi + g;
        if (k >= ksh) continue;
        double sc = bra_scale[k];
        double *dst = gpart + ((size_t)split * ksh + k) * LDG + 2 * t;
#pragma unroll
        for (int ni = 0; ni < NT; ni++) {
            double2 v = make_double2(sc * acc[mi][ni][0], sc * acc[mi][ni][1]);
            *reinterpret_cast<double2 *>(dst + 8 * ni) = v;
        }
    }
  }
}

}  // namespace wfo
