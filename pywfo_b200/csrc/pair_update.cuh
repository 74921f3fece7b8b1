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
// ONE persistent CTA per SM: 8 warps x 32 bra excitations (4 m-tiles of 8 per warp) = 256 bra
// excitations per bra tile (12 warps x 24 rows run at the same speed at C4 but waste more of a
// small shard's ragged last tile).  Work decomposition is "stream-K": the iteration
// space (bra tile, ket tile of 32) is cut into gridDim.x equal contiguous ranges, one per CTA, so
// every SM executes the same number of inner iterations (+-1) whatever the shard size -- no wave
// quantisation, no atomic work counter, a static and therefore run-to-run deterministic summation
// order.  (Several CTAs per SM with equal static shares do NOT work: the warp schedulers favour
// the older CTA and the younger ones finish ~15 % later, measured.)  A CTA whose range crosses a bra-tile boundary flushes its accumulators into
// the partial-sum slot gpart[part] of that bra tile (part = rank of the CTA among the CTAs that
// touch the tile, see t1_part_range); the epilogue kernel adds the few parts of each tile.  The ket
// coefficient tiles (32 excitations x ldg, one contiguous block of CkTs) are staged in shared
// memory by TMA bulk copies (cp.async.bulk, SASS UBLKCP) through a 4-stage ring with full/empty
// mbarriers -- no block barrier in the main loop; the 8 warps share every tile and may drift by
// up to three tiles.  Ai and W are gathered
// through L1 (the host orders the ket excitations so that a CTA's working set
// of W is a few KB at a time, see pywfo_b200/excitations.py).
#pragma once
#include "common.cuh"

namespace wfo {

#ifndef WFO_T1_WARPS
#define WFO_T1_WARPS 8
#endif
#ifndef WFO_T1_STAGES
#define WFO_T1_STAGES 4
#endif
#ifndef WFO_T1_MT
#define WFO_T1_MT 4
#endif
constexpr int T1_WARPS = WFO_T1_WARPS;
constexpr int T1_MT = WFO_T1_MT;                 // m-tiles (8 bra excitations) per warp
constexpr int T1_KTILE = T1_WARPS * T1_MT * 8;   // 256 bra excitations per CTA
#ifndef WFO_T1_LT
#define WFO_T1_LT 32
#endif
constexpr int T1_LT = WFO_T1_LT;                 // ket excitations per shared tile
constexpr int T1_STAGES = WFO_T1_STAGES;         // depth of the TMA ring

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

__device__ __forceinline__ bool mbar_test(unsigned long long *bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

// How the partial sums of the fused kernels are laid out: gpart[(part*ksh + k)*ldg + j].
//   uniform (T0 kernels): every bra row has `uniform_parts` parts;
//   stream-K (T1): iteration x = ktile*n_ltiles + ltile, CTA c of G owns [c*I/G, (c+1)*I/G);
//   the CTA that contains x is ((x+1)*G - 1) / I, so bra tile `ktile` is touched by the CTAs
//   c_first..c_last of its first and last iteration and CTA c writes part c - c_first.
struct PartMap {
    long long I;         // total iterations (0 = uniform mode)
    int G;               // CTAs
    int n_ltiles;        // ket tiles per bra tile
    int ktile_rows;      // bra rows per tile
    int uniform_parts;
};
__host__ __device__ inline int t1_cta_of_iter(const PartMap &pm, long long x) {
    return (int)(((x + 1) * pm.G - 1) / pm.I);
}
__host__ __device__ inline void t1_part_range(const PartMap &pm, int ktile, int &c_first, int &n_parts) {
    const long long x0 = (long long)ktile * pm.n_ltiles;
    c_first = t1_cta_of_iter(pm, x0);
    n_parts = t1_cta_of_iter(pm, x0 + pm.n_ltiles - 1) - c_first + 1;
}
__host__ __device__ inline int parts_of_row(const PartMap &pm, int k) {
    if (pm.I == 0) return pm.uniform_parts;
    int c_first, n_parts;
    t1_part_range(pm, k / pm.ktile_rows, c_first, n_parts);
    return n_parts;
}

#ifdef WFO_T1_CLOCKS
__device__ long long g_t1_clk[2048][4];   // per CTA: elapsed clocks, SM id, start (globaltimer ns), end
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#endif

// NT = ldg / 8: row stride of CkTs / G in 8-column tiles (odd: conflict-free fragment loads).
// Of those columns the first 8*ND are contracted on the tensor cores (DMMA); ND = ceil(n_ket/8)
// can be smaller than NT (the D(k,0) column and the odd-stride padding are never computed here).
// REM further columns can be contracted with plain DFMA; measured on B200 that loses (a DFMA
// between DMMAs costs ~11 pipe cycles per SMSP, not 2), so the dispatch only uses REM = 0.
// One persistent CTA per SM; ckts and ket are padded with zero rows to a multiple of T1_LT
// excitations, so every tile is a full tile.
template <int NT>
constexpr size_t t1_smem_bytes() { return (size_t)T1_STAGES * (T1_LT * NT * 8 * sizeof(double) + T1_LT * sizeof(KetMeta)); }

template <int NT, int ND, int REM>
__global__ void __launch_bounds__(T1_WARPS * 32, 1)
t1_fused_kernel(const double *__restrict__ Ai, const double *__restrict__ W,
                const BraMeta *__restrict__ bra, const double *__restrict__ bra_scale, int ksh,
                const KetMeta *__restrict__ ket, const double *__restrict__ ckts, int kk,
                PartMap pm, double *__restrict__ gpart) {
    constexpr int LDG = NT * 8;
    constexpr unsigned TILE_BYTES = T1_LT * LDG * sizeof(double), META_BYTES = T1_LT * sizeof(KetMeta);
    extern __shared__ __align__(128) unsigned char t1_sm[];
    double *Bs = reinterpret_cast<double *>(t1_sm);                                       // [stage][T1_LT * LDG]
    KetMeta *Ms = reinterpret_cast<KetMeta *>(t1_sm + (size_t)T1_STAGES * TILE_BYTES);   // [stage][T1_LT]
    __shared__ __align__(8) unsigned long long full_bar[T1_STAGES];    // "tile landed" (TMA transaction count)
    __shared__ __align__(8) unsigned long long empty_bar[T1_STAGES];   // "all warps are done with the tile"

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < T1_STAGES; s++) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], T1_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
#ifdef WFO_T1_CLOCKS
    const long long dbg_c0 = clock64();
    const unsigned long long dbg_g0 = gtimer();
#endif
    const long long it0 = ((long long)blockIdx.x * pm.I) / pm.G;
    const int Q = (int)(((long long)(blockIdx.x + 1) * pm.I) / pm.G - it0);   // ket tiles this CTA consumes

    // ---- producer (warp 0; lane 0 issues): tile i of this CTA goes to stage i % T1_STAGES once every
    // warp has released the tile that used the stage before.  Non-blocking except when the tile is
    // needed right now, so the warp never waits for slower warps longer than the ring allows.
    int issued = 0;
    auto produce_upto = [&](int target, bool blocking) {
        if (target > Q) target = Q;
        while (issued < target) {
            const int s = issued % T1_STAGES;
            if (issued >= T1_STAGES) {
                const unsigned par = (unsigned)((issued / T1_STAGES) - 1) & 1u;
                if (blocking) mbar_wait(&empty_bar[s], par);
                else if (!mbar_test(&empty_bar[s], par)) break;
            }
            if (lane == 0) {
                const int l0 = (int)((it0 + issued) % pm.n_ltiles) * T1_LT;
                mbar_expect_tx(&full_bar[s], TILE_BYTES + META_BYTES);
                tma_bulk_g2s(Bs + (size_t)s * T1_LT * LDG, ckts + (size_t)l0 * LDG, TILE_BYTES, &full_bar[s]);
                tma_bulk_g2s(Ms + (size_t)s * T1_LT, ket + l0, META_BYTES, &full_bar[s]);
            }
            issued++;
        }
    };
    if (warp == 0) produce_upto(T1_STAGES, false);

    int q = 0;
    while (q < Q) {
        const long long it = it0 + q;
        const int ktile = (int)(it / pm.n_ltiles);
        const int lt0 = (int)(it - (long long)ktile * pm.n_ltiles);
        const int cnt = min(pm.n_ltiles - lt0, Q - q);
        const int split = (int)blockIdx.x - t1_cta_of_iter(pm, (long long)ktile * pm.n_ltiles);
        const int k0 = ktile * T1_KTILE + warp * (T1_MT * 8);

        // per-thread bra data for its rows (one per m-tile)
        int bf[T1_MT], bw[T1_MT];
        double bx[T1_MT];
#pragma unroll
        for (int mi = 0; mi < T1_MT; mi++) {
            const int k = k0 + 8 * mi + g;
            if (k < ksh) {
                const BraMeta m = bra[k];
                bf[mi] = m.f; bw[mi] = m.woff; bx[mi] = m.x;
            } else {
                bf[mi] = 0; bw[mi] = 0; bx[mi] = 0.0;   // harmless duplicates, never stored
            }
        }
        double acc[T1_MT][ND > 0 ? ND : 1][2];
        double acc2[T1_MT][REM > 0 ? REM : 1];   // DFMA columns: this thread's share (its l = 4 ks + t)
#pragma unroll
        for (int mi = 0; mi < T1_MT; mi++) {
#pragma unroll
            for (int ni = 0; ni < ND; ni++) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
#pragma unroll
            for (int jj = 0; jj < REM; jj++) acc2[mi][jj] = 0.0;
        }

        for (int j = 0; j < cnt; j++, q++) {
            const int s = q % T1_STAGES;
            if (warp == 0) {
                produce_upto(q + 1, true);
                produce_upto(q + T1_STAGES, false);
            }
            mbar_wait(&full_bar[s], (unsigned)(q / T1_STAGES) & 1u);
            const double *bs = Bs + (size_t)s * T1_LT * LDG;
            const KetMeta *ms = Ms + (size_t)s * T1_LT;
            // software pipeline over the 8 k-steps (4 ket excitations each): gather one step ahead
            double a_cur[T1_MT], ai_n[T1_MT], w_n[T1_MT], y_n;
            {
                const KetMeta m = ms[t];
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
                    const KetMeta m = ms[4 * (ks + 1) + t];
                    y_n = m.y;
#pragma unroll
                    for (int mi = 0; mi < T1_MT; mi++) {
                        ai_n[mi] = __ldg(Ai + m.aoff + bf[mi]);
                        w_n[mi] = __ldg(W + bw[mi] + m.tp);
                    }
                }
                // DFMA columns first, as one group (volatile: ptxas keeps them together in front of the
                // DMMAs of this k-step instead of sprinkling them between tensor instructions)
                if (REM > 0) {
                    double c2[REM > 0 ? REM : 1];
#pragma unroll
                    for (int jj = 0; jj < REM; jj += 2) {
                        const double2 v = *reinterpret_cast<const double2 *>(bs + (4 * ks + t) * LDG + 8 * ND + jj);
                        c2[jj] = v.x;
                        c2[jj + 1] = v.y;
                    }
#pragma unroll
                    for (int mi = 0; mi < T1_MT; mi++)
#pragma unroll
                        for (int jj = 0; jj < REM; jj++)
                            asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(acc2[mi][jj]) : "d"(a_cur[mi]), "d"(c2[jj]));
                }
                double b[ND > 0 ? ND : 1];
#pragma unroll
                for (int ni = 0; ni < ND; ni++) b[ni] = bs[(4 * ks + t) * LDG + 8 * ni + g];
#pragma unroll
                for (int mi = 0; mi < T1_MT; mi++)
#pragma unroll
                    for (int ni = 0; ni < ND; ni++) dmma884(acc[mi][ni][0], acc[mi][ni][1], a_cur[mi], b[ni]);
                if (ks == T1_LT / 8 - 1 && warp == 0) produce_upto(q + T1_STAGES, false);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);   // this warp is done with the stage
        }

        // flush: row scale s_k * det(A), store into this CTA's partial-sum slot of the bra tile
#pragma unroll
        for (int mi = 0; mi < T1_MT; mi++) {
            const int k = k0 + 8 * mi + g;
            // DFMA columns: add the four lanes (t = 0..3) that share the row, fixed order
            double r2[REM > 0 ? REM : 1];
#pragma unroll
            for (int jj = 0; jj < REM; jj++) {
                double v = acc2[mi][jj];
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                r2[jj] = v;
            }
            if (k >= ksh) continue;
            const double sc = bra_scale[k];
            double *row = gpart + ((size_t)split * ksh + k) * LDG;
#pragma unroll
            for (int ni = 0; ni < ND; ni++) {
                const double2 v = make_double2(sc * acc[mi][ni][0], sc * acc[mi][ni][1]);
                *reinterpret_cast<double2 *>(row + 8 * ni + 2 * t) = v;
            }
            // the rest of the row: the DFMA columns, then zeros up to the row stride
            constexpr int REST = LDG - 8 * ND;   // <= 16
#pragma unroll
            for (int c = 0; c < REST; c += 8) {
                const int j = c + 2 * t;
                if (j + 1 < REST || j < REST) {
                    double v0 = 0.0, v1 = 0.0;
#pragma unroll
                    for (int jj = 0; jj < REM; jj++) {
                        v0 = (jj == j) ? sc * r2[jj] : v0;
                        v1 = (jj == j + 1) ? sc * r2[jj] : v1;
                    }
                    if (j < REST) row[8 * ND + j] = v0;
                    if (j + 1 < REST) row[8 * ND + j + 1] = v1;
                }
            }
        }
    }
#ifdef WFO_T1_CLOCKS
    if (tid == 0 && blockIdx.x < 2048) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        g_t1_clk[blockIdx.x][0] = clock64() - dbg_c0;
        g_t1_clk[blockIdx.x][1] = smid;
        g_t1_clk[blockIdx.x][2] = (long long)dbg_g0;
        g_t1_clk[blockIdx.x][3] = (long long)gtimer();
    }
#endif
}

}  // namespace wfo
