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
// coefficient tiles (32 excitations x ldg) are staged in shared memory with
// cp.async double buffering and shared by the 4 warps.  Ai and W are gathered
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

__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
    unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

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
    __shared__ __align__(16) double Bs[2][T1_LT * LDG];
    __shared__ __align__(16) KetMeta Ms[2][T1_LT];
    __shared__ int s_item;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
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

    // stage one tile of ket data: T1_LT rows of LDG doubles + T1_LT metas (16 B chunks)
    auto stage = [&](int buf, int l0) {
        constexpr int CH_B = T1_LT * LDG / 2;   // 16-byte chunks of the coefficient tile
        for (int c = tid; c < CH_B + T1_LT; c += T1_WARPS * 32) {
            if (c < CH_B) {
                int row = c / (LDG / 2);
                double *dst = &Bs[buf][0] + 2 * c;
                if (l0 + row < l_end) cp_async16(dst, ckts + (size_t)l0 * LDG + 2 * c);
                else { dst[0] = 0.0; dst[1] = 0.0; }
            } else {
                int row = c - CH_B;
                KetMeta *dst = &Ms[buf][row];
                if (l0 + row < l_end) cp_async16(dst, ket + l0 + row);
                else { dst->aoff = 0; dst->tp = 0; dst->y = 0.0; }
            }
        }
        cp_async_commit();
    };

    int buf = 0;
    if (l_begin < l_end) stage(0, l_begin);
    for (int l0 = l_begin; l0 < l_end; l0 += T1_LT) {
        if (l0 + T1_LT < l_end) { stage(buf ^ 1, l0 + T1_LT); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
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
