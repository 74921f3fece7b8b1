// T0 (brute force), warp-per-matrix kernel for 32 < n <= 128: every warp factors its own
// determinant pair, with NO block barrier anywhere.
//
// Why: a pivoted LU has n strictly sequential column steps of ~200-300 cycles (pivot search,
// broadcast, reciprocal, update).  With the working matrix in shared memory only two 80 KB
// matrices fit per SM at n = 100, so those chains cannot overlap.  Here the working matrix
// lives in a per-warp GLOBAL scratch that stays L2 resident (8-16 warps x 83 KB per SM,
// ~100 MB on the chip), only the 16-column panel lives in registers, and 8-16 independent
// factorisations per SM hide each other's latency.
//
//   * nothing is gathered up front: the first panel and the first trailing update read the
//     minor straight from S_MO through the row/column index lists, the update writes the
//     Schur complement to the scratch;
//   * rows never move: a per-warp row map (shared memory) lists the still active rows in
//     order; pivot rows are dropped from it after every panel; the permutation sign is the
//     parity of the pivots' ranks in that list;
//   * panel: one row per lane and slot (SLOTS = ceil(n/32) rows per lane, 16 columns in
//     registers), REDUX.MAX pivot search over (|a| high word, slot, lane), shuffle broadcast
//     of the pivot row, branch-free elimination;
//   * the panel lanes turn their multipliers into M = L21 L11^-1, so that the Schur complement
//     is A22 - M A12 with the RAW pivot rows (no triangular solve, U12 is never needed);
//   * trailing update with DMMA.8x8x4 (rank 16): B fragments (pivot rows) are held in
//     registers for a chunk of 4 column tiles while the warp walks the active row tiles.
// The determinant is the signed product of the pivots = np.linalg.det(S_MO[rows][:, cols])
// (pywfo/main.py:583-584).
#pragma once
#include "common.cuh"

namespace wfo {

#ifdef WFO_WPM_CLOCKS
__device__ long long g_wpm_clk[8];
#define WPM_T(i) do { long long now_ = clock64(); t_acc[i] += now_ - t_prev; t_prev = now_; } while (0)
#else
#define WPM_T(i)
#endif

struct WpmShared {            // per warp
    double L11[16][16];       // unit-lower factor of the current pivot block (strict lower part)
    double urow[16];          // pivot row of the current column step (written by its lane, read by all)
    unsigned char rowmap[128];   // active rows in order
    unsigned char piv[16];    // pivot rows of the current panel, in step order
};

__host__ __device__ inline int wpm_ld(int n) { return (n + 7) & ~7; }
inline size_t wpm_mat_bytes(int n) { return (size_t)n * wpm_ld(n) * sizeof(double); }

// trailing update of the active rows, columns [j+16, n), with the 16 pivot rows of the panel at j
template <bool FIRST>
__device__ __forceinline__ void wpm_update(const double *__restrict__ S, int ld_s, const int *__restrict__ rows,
                                           const int *__restrict__ cols, double *W, int LD, const WpmShared *sh, int n,
                                           int j, int act, int lane) {
    const int g = lane >> 2, t = lane & 3;
    const int C0 = j + 16;
    const int TC = (n - C0 + 7) >> 3, TR = (act + 7) >> 3;
    for (int tc0 = 0; tc0 < TC; tc0 += 4) {
        // ---- B fragments: pivot row 4ks+t, column C0 + 8 jc + g
        double bf[4][4];
#pragma unroll
        for (int ks = 0; ks < 4; ks++) {
            const int pp = sh->piv[4 * ks + t];
            const double *prow = FIRST ? S + (size_t)rows[pp] * ld_s : W + (size_t)pp * LD;
#pragma unroll
            for (int jc = 0; jc < 4; jc++) {
                const int col = C0 + 8 * (tc0 + jc) + g;
                if (FIRST) bf[ks][jc] = (col < n) ? __ldg(prow + cols[col]) : 0.0;
                else bf[ks][jc] = (tc0 + jc < TC) ? prow[col] : 0.0;
            }
        }
        for (int tr = 0; tr < TR; tr++) {
            const int lp = 8 * tr + g;
            const bool rvalid = lp < act;
            const int row = sh->rowmap[rvalid ? lp : act - 1];
            double *wrow = W + (size_t)row * LD;
            double af[4];
#pragma unroll
            for (int ks = 0; ks < 4; ks++) af[ks] = -wrow[j + 4 * ks + t];
            double acc[4][2];
            const double *srow = FIRST ? S + (size_t)rows[row] * ld_s : nullptr;
#pragma unroll
            for (int jc = 0; jc < 4; jc++) {
                const int col0 = C0 + 8 * (tc0 + jc) + 2 * t;
                if (FIRST) {
                    acc[jc][0] = (col0 < n) ? __ldg(srow + cols[col0]) : 0.0;
                    acc[jc][1] = (col0 + 1 < n) ? __ldg(srow + cols[col0 + 1]) : 0.0;
                } else if (tc0 + jc < TC) {
                    const double2 v = *reinterpret_cast<const double2 *>(wrow + col0);
                    acc[jc][0] = v.x;
                    acc[jc][1] = v.y;
                } else {
                    acc[jc][0] = acc[jc][1] = 0.0;
                }
            }
#pragma unroll
            for (int ks = 0; ks < 4; ks++)
#pragma unroll
                for (int jc = 0; jc < 4; jc++) dmma884(acc[jc][0], acc[jc][1], af[ks], bf[ks][jc]);
            if (rvalid) {
#pragma unroll
                for (int jc = 0; jc < 4; jc++)
                    if (tc0 + jc < TC)
                        *reinterpret_cast<double2 *>(wrow + C0 + 8 * (tc0 + jc) + 2 * t) = make_double2(acc[jc][0], acc[jc][1]);
            }
        }
    }
}

// whole factorisation by one warp; returns det to every lane.  W: n x LD scratch (global).
// One 16-column panel with A active row slots (A = ceil(active rows / 32): the panel code is
// instantiated for every slot count and the factorisation picks the smallest that covers the rows
// still active, so the later, shorter panels do not pay for empty slots).  Factors the panel, turns
// the multipliers into M = L21 L11^-1, writes M and drops the pivot rows from the row map.
// Returns true when this was the last panel.
template <int A>
__device__ __forceinline__ bool wpm_panel(const double *__restrict__ S, int ld_s, const int *__restrict__ rows,
                                          const int *__restrict__ cols, int n, double *W, int LD, WpmShared *sh, int lane,
                                          int j, int &act, double &det, bool &zero, unsigned &par) {
    const unsigned FULL = 0xffffffffu;
    const int w = min(16, n - j);
    const bool first = (j == 0);
    // ---- load the panel: lane/slot <-> active position lane + 32 s
    int pr[A];
    bool done[A];
    int pstep[A];
    double a[A][16];
#pragma unroll
    for (int s = 0; s < A; s++) {
        const int L = lane + 32 * s;
        const bool valid = L < act;
        pr[s] = sh->rowmap[valid ? L : 0];
        done[s] = !valid;
        pstep[s] = -1;
        if (first) {
            const double *src = S + (size_t)rows[pr[s]] * ld_s;
#pragma unroll
            for (int c = 0; c < 16; c++) a[s][c] = (c < w) ? __ldg(src + cols[c]) : 0.0;
        } else {
            const double2 *src = reinterpret_cast<const double2 *>(W + (size_t)pr[s] * LD + j);
#pragma unroll
            for (int c = 0; c < 8; c++) {
                const double2 v = src[c];
                a[s][2 * c] = v.x;
                a[s][2 * c + 1] = v.y;
            }
        }
    }
    int ppos[16];
    // ---- 16 column steps
#pragma unroll
    for (int jj = 0; jj < 16; jj++) {
        ppos[jj] = 0;
        if (jj < w) {   // uniform
            unsigned key = 0;
#pragma unroll
            for (int s = 0; s < A; s++) {
                const unsigned k = (hi_abs(a[s][jj]) & 0xffffff00u) | 0x80u | (unsigned)(s << 5) | (unsigned)lane;
                key = max(key, done[s] ? 0u : k);
            }
            const unsigned kmax = __reduce_max_sync(FULL, key);
            const int wl = kmax & 31, ws = (kmax >> 5) & 3;
            ppos[jj] = wl + 32 * ws;
            // the pivot row goes through shared memory (one lane writes, all read it back): a
            // handful of STS/LDS instead of up to 34 shuffles, which 8 warps per SM saturate
            double u[16];
#pragma unroll
            for (int s = 0; s < A; s++) {
                if (ws == s && lane == wl) {
#pragma unroll
                    for (int c = jj; c < 16; c++) sh->urow[c] = a[s][c];
#pragma unroll
                    for (int c = 0; c < jj; c++) sh->L11[jj][c] = a[s][c];
                    sh->piv[jj] = (unsigned char)pr[s];
                }
            }
            __syncwarp();
#pragma unroll
            for (int c = jj; c < 16; c++) u[c] = sh->urow[c];
            __syncwarp();
            const double piv = u[jj];
            det *= piv;
            zero = zero || (piv == 0.0);
            const double rcp = fast_rcp(piv);
#pragma unroll
            for (int s = 0; s < A; s++) {
                const bool win = (lane == wl) && (ws == s) && (kmax != 0u);
                pstep[s] = win ? jj : pstep[s];
                done[s] = done[s] || win;
                const double l = done[s] ? 0.0 : a[s][jj] * rcp;
                a[s][jj] = done[s] ? a[s][jj] : l;
#pragma unroll
                for (int c = jj + 1; c < 16; c++) a[s][c] = fma(-l, u[c], a[s][c]);
            }
        }
    }
    // ---- sign: rank of every pivot among the rows still active when it is extracted
    {
        int myp = 0;
#pragma unroll
        for (int i = 0; i < 16; i++) myp = (lane == i) ? ppos[i] : myp;
        int cnt = 0;
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const int pi = __shfl_sync(FULL, myp, i);
            cnt += (i < lane && pi < myp) ? 1 : 0;
        }
        const bool odd = (lane < w) && (((myp - cnt) & 1) != 0);
        par ^= (unsigned)(__popc(__ballot_sync(FULL, odd)) & 1);
    }
    if (j + 16 >= n) return true;
    __syncwarp();   // L11[][] and piv[] complete
    // ---- M = L21 L11^-1 (axpy form of x L11 = l); pivot rows are modified too, harmlessly
#pragma unroll
    for (int q = 15; q >= 1; q--) {
#pragma unroll
        for (int c = 0; c < q; c++) {
            const double lqc = sh->L11[q][c];
#pragma unroll
            for (int s = 0; s < A; s++) a[s][c] = fma(-a[s][q], lqc, a[s][c]);
        }
    }
    // ---- write M of the remaining rows, drop the pivot rows from the row map
    int newL[A];
    int below = 0;
#pragma unroll
    for (int s = 0; s < A; s++) {
        const int L = lane + 32 * s;
        const bool valid = L < act;
        const bool isp = valid && pstep[s] >= 0;
        const unsigned bal = __ballot_sync(FULL, isp);
        newL[s] = L - (below + __popc(bal & ((1u << lane) - 1u)));
        below += __popc(bal);
        if (valid && !isp) {
            double2 *dst = reinterpret_cast<double2 *>(W + (size_t)pr[s] * LD + j);
#pragma unroll
            for (int c = 0; c < 8; c++) dst[c] = make_double2(a[s][2 * c], a[s][2 * c + 1]);
        }
    }
    __syncwarp();   // every lane has read its rowmap entries (pr[]) long ago; order the writes below
#pragma unroll
    for (int s = 0; s < A; s++) {
        const int L = lane + 32 * s;
        if (L < act && pstep[s] < 0) sh->rowmap[newL[s]] = (unsigned char)pr[s];
    }
    act -= 16;
    __syncwarp();   // M (global) and the row map are visible to the whole warp
    return false;
}

template <int SLOTS>
__device__ __forceinline__ double wpm_lu_det(const double *__restrict__ S, int ld_s, const int *__restrict__ rows,
                                             const int *__restrict__ cols, int n, double *W, int LD, WpmShared *sh,
                                             int lane) {
    for (int i = lane; i < n; i += 32) sh->rowmap[i] = (unsigned char)i;
    __syncwarp();
    int act = n;
    double det = 1.0;
    bool zero = false;
    unsigned par = 0;
    for (int j = 0; j < n; j += 16) {
        const bool first = (j == 0);
        const int need = (act + 31) >> 5;   // row slots this panel needs
        bool last;
        if (SLOTS >= 4 && need >= 4) last = wpm_panel<(SLOTS >= 4 ? 4 : 1)>(S, ld_s, rows, cols, n, W, LD, sh, lane, j, act, det, zero, par);
        else if (SLOTS >= 3 && need == 3) last = wpm_panel<(SLOTS >= 3 ? 3 : 1)>(S, ld_s, rows, cols, n, W, LD, sh, lane, j, act, det, zero, par);
        else if (SLOTS >= 2 && need == 2) last = wpm_panel<(SLOTS >= 2 ? 2 : 1)>(S, ld_s, rows, cols, n, W, LD, sh, lane, j, act, det, zero, par);
        else last = wpm_panel<1>(S, ld_s, rows, cols, n, W, LD, sh, lane, j, act, det, zero, par);
        if (last) break;
        if (first) wpm_update<true>(S, ld_s, rows, cols, W, LD, sh, n, j, act, lane);
        else wpm_update<false>(S, ld_s, rows, cols, W, LD, sh, n, j, act, lane);
        __syncwarp();
    }
    return zero ? 0.0 : (par ? -det : det);
}

#ifndef WFO_WPM_MINB
#define WFO_WPM_MINB 2
#endif
constexpr int WPM_WARPS = 4;   // warps per CTA (independent of each other)

template <int SLOTS>
__global__ void __launch_bounds__(WPM_WARPS * 32, SLOTS <= 2 ? 4 : WFO_WPM_MINB)
det_table_wpm_kernel(const double *__restrict__ S, int ld_s, int n, const int32_t *__restrict__ rows, int kb,
                     const int32_t *__restrict__ cols, int kk, double *__restrict__ D, double *gscratch) {
    __shared__ WpmShared shs[WPM_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int LD = wpm_ld(n);
    const long long gw = (long long)blockIdx.x * WPM_WARPS + warp, nw = (long long)gridDim.x * WPM_WARPS;
    double *W = gscratch + (size_t)gw * n * LD;
    const long long npairs = (long long)kb * kk;
    for (long long pair = gw; pair < npairs; pair += nw) {
        const int k = (int)(pair / kk), l = (int)(pair - (long long)k * kk);
        const double det = wpm_lu_det<SLOTS>(S, ld_s, rows + (size_t)k * n, cols + (size_t)l * n, n, W, LD, &shs[warp], lane);
        if (lane == 0) D[pair] = det;
        __syncwarp();
    }
}

// one work item per warp = (bra determinant k, chunk of ket determinants)
template <int SLOTS>
__global__ void __launch_bounds__(WPM_WARPS * 32, SLOTS <= 2 ? 4 : WFO_WPM_MINB)
t0_fused_wpm_kernel(const double *__restrict__ S, int ld_s, int n, const int32_t *__restrict__ rows, int ksh,
                    const int32_t *__restrict__ cols, int kk, const double *__restrict__ ckt, int ldg, int chunk_len,
                    int n_chunks, double *__restrict__ gpart, double *gscratch) {
    __shared__ WpmShared shs[WPM_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int LD = wpm_ld(n);
    const long long gw = (long long)blockIdx.x * WPM_WARPS + warp, nw = (long long)gridDim.x * WPM_WARPS;
    double *W = gscratch + (size_t)gw * n * LD;
    const long long nitems = (long long)ksh * n_chunks;
    for (long long item = gw; item < nitems; item += nw) {
        const int k = (int)(item / n_chunks), ch = (int)(item - (long long)k * n_chunks);
        const int l0 = ch * chunk_len, l1 = min(kk, l0 + chunk_len);
        double g0 = 0.0, g1 = 0.0;
        for (int l = l0; l < l1; l++) {
            const double det = wpm_lu_det<SLOTS>(S, ld_s, rows + (size_t)k * n, cols + (size_t)l * n, n, W, LD, &shs[warp], lane);
            if (lane < ldg) g0 = fma(det, __ldg(ckt + (size_t)l * ldg + lane), g0);
            if (lane + 32 < ldg) g1 = fma(det, __ldg(ckt + (size_t)l * ldg + lane + 32), g1);
            __syncwarp();
        }
        double *dst = gpart + ((size_t)ch * ksh + k) * ldg;
        if (lane < ldg) dst[lane] = g0;
        if (lane + 32 < ldg) dst[lane + 32] = g1;
    }
}

}  // namespace wfo
