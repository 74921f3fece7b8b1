// Base-block factorisation for the rank-update tiers: in-place Gauss-Jordan inverse with
// partial (row) pivoting of A = S_MO[:n,:n] (n <= 128) in ONE CTA, REGISTER resident (32 matrix
// elements per thread), blocked by panels of 8 columns and pipelined over the warps WITHOUT block
// barriers.  Pivoting is implicit (rows never move), the row/column permutation is undone when the
// inverse is written out.  Produces Ai = A^-1, det(A) and min/max |pivot| as a conditioning indicator.
//
// This is the "one base LU" that north_star's rank-1 (Sherman-Morrison / cofactor) updates
// reuse; upstream the same numbers would come from np.linalg.det / np.linalg.inv of
// S_MO[:occ,:occ] (pywfo/main.py:548).
#pragma once
#include "common.cuh"

namespace wfo {

struct BaseInfo {     // device-resident result block
    double det;       // det(A)
    double min_piv;   // min |pivot|
    double max_piv;   // max |pivot|
    int singular;     // 1 if an exactly zero / non-finite pivot was met
    int pad;
};

#ifdef WFO_LUB_CLOCKS
__device__ long long g_lub_clk[16];
#define LUB_T(i) do { long long now_ = clock64(); t_acc[i] += now_ - t_prev; t_prev = now_; } while (0)
#else
#define LUB_T(i)
#endif
constexpr int LUB_MAXN = 128;
constexpr int LUB_THREADS = 512;
constexpr int LUB_WARPS = LUB_THREADS / 32;   // 16
constexpr int LUB_CS = LUB_MAXN / LUB_WARPS;  // 8 columns per warp = one panel (columns 8 warp + c)
constexpr int LUB_RS = 4;                     // 4 row slots per thread (rows 4 lane + s)
constexpr int LUB_LDS = LUB_MAXN + 1;         // row stride of the staging tile
constexpr size_t LUB_EXCL_SMEM = 200 * 1024;  // dynamic shared memory request that leaves no room for another CTA on the SM
// staging tile (n x LUB_LDS) on entry / exit; in between the same memory holds the multiplier panels
// (ceil(n/8) panels x 8 columns x row stride 4 ceil(n/4))
inline size_t lub_smem_bytes(int n) {
    const size_t tile = (size_t)n * LUB_LDS, panels = (size_t)((n + 7) / 8) * 8 * ((n + 3) & ~3);
    return (tile > panels ? tile : panels) * sizeof(double);
}

// Ownership: lane l holds rows 4l..4l+3, warp w holds columns 8w..8w+7 (= panel w), so
//   * a panel is factored by ONE warp on its own: per column 4 compares per lane and one REDUX, the
//     pivot row is exchanged inside the warp, no block barrier; the reciprocal of the pivot is computed
//     speculatively by every lane for its own best candidate while the reduction is in flight;
//   * the update is branch- and select-free: the pivot row's own multiplier is published as d - 1,
//     so that a_p - (d-1) (a_p / d) = a_p / d comes out of the same multiply-add as every other row;
//   * the panel's multiplier columns M_k (n x 8), pivot rows and reciprocals are PUBLISHED to shared memory
//     and a flag is raised; every other warp applies the panel to its own 8 columns when it gets there:
//       raw_s = its part of pivot row s (as it was before the panel), s = 0..7
//       pr_s  = (raw_s - sum_{s'<s} M_k[p_s][s'] pr_s') / d_s          (8 x 8 triangular recurrence)
//       a    -= sum_s M_k[:, s] pr_s                                    (rank-8 update, 256 FMA per thread)
//     Each panel has its own shared-memory region (the staging tile is free once the registers are
//     loaded), so nothing is ever overwritten and no warp waits for a slower consumer: warp k+1 applies
//     panel k and factors panel k+1 while the other warps are still busy with older panels;
//   * det / min / max pivot are accumulated by the owning warps and combined at the end.
// The factorisation itself: lub_tile (n x n, row stride LUB_LDS, shared memory, lub_smem_bytes(n) bytes) holds A
// on entry and, unless the matrix is singular, A^-1 on return (all threads synchronised).  Every thread returns
// det(A), min/max |pivot| and the singular flag.  Shared by base_inverse_kernel (one matrix) and
// cofactor_rows_kernel (a batch).  (Round 1's version took one block barrier per COLUMN: 81 us at n = 100.)
__device__ __forceinline__ void lub_invert_tile(double *lub_tile, int n, double &det_out, double &min_out, double &max_out,
                                                int &sing_out) {
    __shared__ __align__(16) double rowbuf[LUB_WARPS][8][LUB_CS];   // warp-private: pivot rows in this warp's columns, then pr
    __shared__ __align__(16) double2 prec[LUB_MAXN];                // per column step: (1 / pivot, pivot row in the low word)
    __shared__ __align__(8) unsigned long long pbar[LUB_MAXN];       // mbarrier j completes when column step j has been published
    __shared__ int piv_of_col[LUB_MAXN];                            // p_j: pivot row chosen for column j
    __shared__ int col_of_piv[LUB_MAXN];                            // inverse map
    __shared__ double wdet[LUB_WARPS], wmin[LUB_WARPS], wmax[LUB_WARPS];
    __shared__ int wsing[LUB_WARPS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned FULL = 0xffffffffu;
    const int np = (n + 7) >> 3;          // panels = active warps
    const int rs = (n + 3) & ~3;          // row stride of one multiplier column
    double a[LUB_RS][LUB_CS];
    unsigned kmask[LUB_RS];   // 0 once row 4 lane + s is a used pivot row (or beyond n), else the key mask
#pragma unroll
    for (int s = 0; s < LUB_RS; s++) {
        const int r = 4 * lane + s;
        kmask[s] = (r < n) ? 0x7fffff00u : 0u;
#pragma unroll
        for (int c = 0; c < LUB_CS; c++) {
            const int col = LUB_CS * warp + c;
            a[s][c] = (r < n && col < n) ? lub_tile[r * LUB_LDS + col] : 0.0;
        }
    }
    if (tid < LUB_MAXN) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"((unsigned)__cvta_generic_to_shared(&pbar[tid])));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    double det = 1.0, minp = 1e300, maxp = 0.0;   // over the columns this warp owns
    int singular = 0;
    __syncthreads();   // the tile is free: it now holds M_k[s][i] at lub_tile[(8 k + s) rs + i]

#define LUB_STS_PRED(PTR_, X_, Y_, PRED_)                                                        \
    asm volatile("{\n.reg .pred q;\nsetp.ne.b32 q, %3, 0;\n@q st.shared.v2.f64 [%0], {%1, %2};\n}\n" ::     \
                 "r"((unsigned)__cvta_generic_to_shared(PTR_)), "d"(X_), "d"(Y_), "r"((int)(PRED_)) : "memory")
#define LUB_ROW_OUT(S_, DST_)                                                                   \
    _Pragma("unroll") for (int c = 0; c < LUB_CS; c += 2)                                       \
        *reinterpret_cast<double2 *>((DST_) + c) = make_double2(a[S_][c], a[S_][c + 1]);
#define LUB_ROW_SWITCH(PS_, DST_)                                                               \
    switch (PS_) {   /* static register indices */                                              \
        case 0: LUB_ROW_OUT(0, DST_) break;                                                     \
        case 1: LUB_ROW_OUT(1, DST_) break;                                                     \
        case 2: LUB_ROW_OUT(2, DST_) break;                                                     \
        default: LUB_ROW_OUT(3, DST_) break;                                                    \
    }

#ifdef WFO_LUB_CLOCKS
    long long t_acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long t_prev = clock64();
#endif
    if (warp < np) {
        for (int k = 0; k < np; k++) {
            double *Mk = lub_tile + (size_t)k * 8 * rs;
            const int ns = min(8, n - 8 * k);
            if (warp == k) {
                LUB_T(5);
                // ---- factor this warp's own panel: column slot s is column 8 k + s.
                // Software pipelined by hand: the dependent chain of a step is only
                //     keys of column s -> REDUX -> three shuffles (pivot, its reciprocal, the pivot row's entry in
                //     column s+1) -> multipliers -> update of column s+1 -> keys of column s+1;
                // the rest of step s (pivot row through shared memory, the other 28 multiply-adds per thread, the
                // multiplier column, bookkeeping) is issued behind the NEXT step's reduction, in its latency shadow.
                double m_prev[LUB_RS] = {0.0, 0.0, 0.0, 0.0};
                double rd_prev = 0.0;
                int p_prev = 0;
                unsigned kbad = 0;  // a step found no candidate
                // the deferred part of step S_ (static)
#define LUB_DEFERRED(S_)                                                                                      \
    do {                                                                                                      \
        __syncwarp();   /* the pivot row of step S_ is in rowbuf */                                           \
        double pr_[LUB_CS];                                                                                   \
        _Pragma("unroll") for (int c = 0; c < LUB_CS; c += 2) {                                               \
            const double2 v = *reinterpret_cast<const double2 *>(&rowbuf[warp][S_][c]);                       \
            pr_[c] = v.x * rd_prev;                                                                           \
            pr_[c + 1] = v.y * rd_prev;                                                                       \
        }                                                                                                     \
        pr_[S_] = rd_prev;   /* column S_ of the pivot row is the unit column's 1 */                          \
        _Pragma("unroll") for (int c = 0; c < LUB_CS; c++) {                                                  \
            if (c == (S_) + 1) continue;   /* done ahead of time */                                           \
            _Pragma("unroll") for (int r = 0; r < LUB_RS; r++) a[r][c] = fma(-m_prev[r], pr_[c], a[r][c]);    \
        }                                                                                                     \
        if (4 * lane < rs) {                                                                                  \
            *reinterpret_cast<double2 *>(Mk + (S_) * rs + 4 * lane) = make_double2(m_prev[0], m_prev[1]);     \
            *reinterpret_cast<double2 *>(Mk + (S_) * rs + 4 * lane + 2) = make_double2(m_prev[2], m_prev[3]); \
        }                                                                                                     \
        if (lane == 0) *reinterpret_cast<double2 *>(&prec[8 * k + (S_)]) = make_double2(rd_prev, __hiloint2double(0, p_prev)); \
        __syncwarp();   /* orders the other lanes' stores before lane 0's release */                          \
        if (lane == 0)   /* column step 8 k + S_ is published */                                              \
            asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];\n" ::"r"(                      \
                (unsigned)__cvta_generic_to_shared(&pbar[8 * k + (S_)])) : "memory");                         \
    } while (0)
#pragma unroll
                for (int s = 0; s < 8; s++) {
                    if (s < ns) {   // uniform
                        unsigned key = 0;
#pragma unroll
                        for (int r = 0; r < LUB_RS; r++)   // kmask: 0x7fffff00 for a free row, 0 for a used one
                            key = max(key, ((unsigned)__double2hiint(a[r][s]) & kmask[r]) | (0x80u + (unsigned)(4 * lane + r)));
                        const unsigned kmax = __reduce_max_sync(FULL, key);
                        // this lane's own best candidate and its reciprocal: if this lane wins, these ARE the pivot
                        // and 1/pivot (computed in the shadow of the reduction)
                        const int bs = key & 3;
                        double cand = a[0][s];
                        cand = (bs == 1) ? a[1][s] : cand;
                        cand = (bs == 2) ? a[2][s] : cand;
                        cand = (bs == 3) ? a[3][s] : cand;
                        const double rc_own = fast_rcp(cand);
                        if (s > 0) LUB_DEFERRED(s - 1);
                        // (after the deferred update of column s+1) the candidate's neighbour in the next column
                        double nxt = a[0][(s + 1) & 7];
                        nxt = (bs == 1) ? a[1][(s + 1) & 7] : nxt;
                        nxt = (bs == 2) ? a[2][(s + 1) & 7] : nxt;
                        nxt = (bs == 3) ? a[3][(s + 1) & 7] : nxt;
                        kbad |= (kmax < 0x100u) ? 1u : 0u;
                        const int p = (int)(kmax & 0x7fu);
                        const int pl = p >> 2;
                        const double d = __shfl_sync(FULL, cand, pl);
                        const double rd = __shfl_sync(FULL, rc_own, pl);
                        const double v1 = __shfl_sync(FULL, nxt, pl);
                        const double pr1 = v1 * rd;
                        // multipliers = the column itself; the slot restarts as the unit column e_p (in-place Gauss-Jordan);
                        // the pivot row's own multiplier is d - 1, so that a_p - (d-1) (a_p/d) = a_p/d
                        double m[LUB_RS];
#pragma unroll
                        for (int r = 0; r < LUB_RS; r++) m[r] = a[r][s];
                        {   // the winner's row goes out, its slot is marked used: predicated, no branch
                            const double dm1 = d - 1.0;
                            double *dst = &rowbuf[warp][s][0];
#pragma unroll
                            for (int r = 0; r < LUB_RS; r++) {
                                const bool isp = (4 * lane + r == p);
#pragma unroll
                                for (int c = 0; c < LUB_CS; c += 2) LUB_STS_PRED(dst + c, a[r][c], a[r][c + 1], isp);
                                m[r] = isp ? dm1 : m[r];
                                a[r][s] = isp ? 1.0 : 0.0;
                                kmask[r] = isp ? 0u : kmask[r];
                            }
                        }
#pragma unroll
                        for (int r = 0; r < LUB_RS; r++) {
                            if (s + 1 < 8) a[r][(s + 1) & 7] = fma(-m[r], pr1, a[r][(s + 1) & 7]);   // the next pivot column first
                            m_prev[r] = m[r];
                        }
                        rd_prev = rd;
                        p_prev = p;
                        {   // statistics and the permutation tables: nobody waits for these
                            const double ad = fabs(d);
                            singular |= (!(ad > 0.0) || !(ad < 1e300)) ? 1 : 0;
                            det *= d;
                            minp = fmin(minp, ad);
                            maxp = fmax(maxp, ad);
                            if (lane == 0) {
                                piv_of_col[8 * k + s] = p;
                                col_of_piv[p] = 8 * k + s;
                            }
                        }
                    }
                }
                // the last step's deferred part (ns is 8 except in a ragged last panel)
                switch (ns) {
                    case 1: LUB_DEFERRED(0); break;
                    case 2: LUB_DEFERRED(1); break;
                    case 3: LUB_DEFERRED(2); break;
                    case 4: LUB_DEFERRED(3); break;
                    case 5: LUB_DEFERRED(4); break;
                    case 6: LUB_DEFERRED(5); break;
                    case 7: LUB_DEFERRED(6); break;
                    default: LUB_DEFERRED(7); break;
                }
                singular |= (int)kbad;
#undef LUB_DEFERRED
                LUB_T(4);
            } else {
                // ---- apply the column steps of panel k to this warp's columns, one by one as they are published
#pragma unroll 1
                for (int sj = 0; sj < ns; sj++) {
                    // (a hardware-suspended wait: polling a flag from 15 warps starved the factoring warp of issue slots and LSU)
                    asm volatile(
                        "{\n.reg .pred p;\nLUB_WAIT:\n"
                        "mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%0], 0;\n"
                        "@p bra LUB_DONE;\nbra LUB_WAIT;\nLUB_DONE:\n}\n" ::"r"((unsigned)__cvta_generic_to_shared(&pbar[8 * k + sj])) : "memory");
                    const double2 rec = prec[8 * k + sj];
                    const double rdj = rec.x;
                    const int pj = __double2loint(rec.y);
                    double *rb = &rowbuf[warp][sj][0];
#pragma unroll
                    for (int r = 0; r < LUB_RS; r++) {   // predicated, no branch; a used pivot row leaves the search
                        const bool isp = (4 * lane + r == pj);
#pragma unroll
                        for (int c = 0; c < LUB_CS; c += 2) LUB_STS_PRED(rb + c, a[r][c], a[r][c + 1], isp);
                        kmask[r] = isp ? 0u : kmask[r];
                    }
                    double2 m01 = make_double2(0.0, 0.0), m23 = m01;
                    if (4 * lane < rs) {
                        m01 = *reinterpret_cast<const double2 *>(Mk + sj * rs + 4 * lane);
                        m23 = *reinterpret_cast<const double2 *>(Mk + sj * rs + 4 * lane + 2);
                    }
                    const double m[LUB_RS] = {m01.x, m01.y, m23.x, m23.y};
                    __syncwarp();
                    double pr[LUB_CS];
#pragma unroll
                    for (int c = 0; c < LUB_CS; c += 2) {
                        const double2 v = *reinterpret_cast<const double2 *>(rb + c);
                        pr[c] = v.x * rdj;
                        pr[c + 1] = v.y * rdj;
                    }
#pragma unroll
                    for (int c = 0; c < LUB_CS; c++)
#pragma unroll
                        for (int r = 0; r < LUB_RS; r++) a[r][c] = fma(-m[r], pr[c], a[r][c]);
                }
            }
        }
    }
#undef LUB_ROW_SWITCH
#undef LUB_ROW_OUT
#ifdef WFO_LUB_CLOCKS
    if (tid == 32 * (np > 1 ? np / 2 : 0)) for (int i = 0; i < 10; i++) g_lub_clk[i] = t_acc[i];
#endif
    // ---- combine the per-warp pivot statistics
    if (lane == 0) { wdet[warp] = det; wmin[warp] = minp; wmax[warp] = maxp; wsing[warp] = singular; }
    __syncthreads();   // also: every panel has been applied everywhere, the panel memory is free again
    if (tid < LUB_MAXN) asm volatile("mbarrier.inval.shared::cta.b64 [%0];\n" ::"r"((unsigned)__cvta_generic_to_shared(&pbar[tid])) : "memory");
    double tdet = 1.0, tmin = 1e300, tmax = 0.0;
    int tsing = 0;
#pragma unroll
    for (int w = 0; w < LUB_WARPS; w++) {
        tdet *= wdet[w];
        tmin = fmin(tmin, wmin[w]);
        tmax = fmax(tmax, wmax[w]);
        tsing |= wsing[w];
    }
    // ---- write out: stored[r][j'] = M[r][p_j'],  Ai[j][c] = M[p_j][c]  =>  Ai[col_of_piv[r]][piv_of_col[j']] = stored[r][j']
    // (permuted into the staging tile, then a coalesced copy)
    if (!tsing) {
#pragma unroll
        for (int s = 0; s < LUB_RS; s++) {
            const int r = 4 * lane + s;
            if (r < n) {
                const int orow = col_of_piv[r];
#pragma unroll
                for (int c = 0; c < LUB_CS; c++) {
                    const int col = LUB_CS * warp + c;
                    if (col < n) lub_tile[orow * LUB_LDS + piv_of_col[col]] = a[s][c];
                }
            }
        }
    }
    // sign of the permutation j -> piv_of_col[j]: parallel inversion count
    int inv = 0;
    if (!tsing && tid < n)
        for (int k = tid + 1; k < n; k++) inv += piv_of_col[k] < piv_of_col[tid];
    const int par = __syncthreads_count(inv & 1) & 1;   // also: the permuted inverse in lub_tile is complete
    det_out = tsing ? 0.0 : (par ? -tdet : tdet);
    min_out = tmin;
    max_out = tmax;
    sing_out = tsing;
}

// A_src: n x n block with leading dimension ld_src.  Ai_out: n x n, leading dimension ld_out.
// Dynamic shared memory: lub_smem_bytes(n) (staging tile for coalesced load / store).
__global__ void __launch_bounds__(LUB_THREADS)
base_inverse_kernel(const double *__restrict__ A_src, int ld_src, int n, double *__restrict__ Ai_out, int ld_out,
                    BaseInfo *__restrict__ info, int n_slabs, long long slab_stride, const int *wait_flag, int wait_value) {
    extern __shared__ __align__(16) double lub_tile[];    // n x LUB_LDS
    __shared__ int go_on;
    const int tid = threadIdx.x;
    if (wait_flag) {
        // launched AHEAD of its input (wfo_smo_inv_dev's early base block): the CTA sits on an SM of its own
        // (its shared-memory request keeps the GEMM CTAs away, whose FP64 traffic would slow this latency-bound
        // kernel by half) and waits until the producer stream raises the flag.  Bounded to a few ms: where kernels
        // do not run concurrently (profilers, CUDA_LAUNCH_BLOCKING, a device saturated by somebody else) the
        // flag cannot come; the kernel then leaves WITHOUT a result (info->singular = 2) and the consumer, which
        // looks at the verdict anyway, runs the inverse again behind the producer.
        if (tid == 0) {
            int seen = 0, ok = 0;
            for (int spin = 0; spin < 4096 && !ok; spin++) {
                asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(seen) : "l"(wait_flag) : "memory");
                ok = (seen == wait_value);
                if (!ok) __nanosleep(200);
            }
            go_on = ok;
            if (!ok) {
                info->det = 0.0;
                info->min_piv = 0.0;
                info->max_piv = 0.0;
                info->singular = 2;
            }
        }
        __syncthreads();
        if (!go_on) return;
    }
    // coalesced load through the staging tile; A may arrive as the K slabs of a split-K product (summed in fixed order)
    for (int idx = tid; idx < n * n; idx += LUB_THREADS) {
        const int r = idx / n, c = idx - r * n;
        double vz[8];   // at most 8 slabs, all loads in flight together
#pragma unroll
        for (int z = 0; z < 8; z++) vz[z] = (z < n_slabs) ? A_src[(size_t)z * slab_stride + (size_t)r * ld_src + c] : 0.0;
        double v = vz[0];
#pragma unroll
        for (int z = 1; z < 8; z++) v += vz[z];
        lub_tile[r * LUB_LDS + c] = v;
    }
    __syncthreads();
    double det, minp, maxp;
    int sing;
    lub_invert_tile(lub_tile, n, det, minp, maxp, sing);
    if (!sing)
        for (int idx = tid; idx < n * n; idx += LUB_THREADS) {
            const int r = idx / n, c = idx - r * n;
            Ai_out[(size_t)r * ld_out + c] = lub_tile[r * LUB_LDS + c];
        }
    if (tid == 0) {
        info->det = det;
        info->min_piv = minp;
        info->max_piv = maxp;
        info->singular = sing;
    }
}

// Super-blocks of a general determinant list (pywfo/dets.py:119-156: blocks that are equal up to the LAST
// orbital; the reference prepares `super_map` and counts the super-blocks, dets.py:233-236, but then takes one
// np.linalg.det per block pair): the n - 1 shared rows are factored ONCE per (super-block, ket block) and
// every member costs one dot product -- the Laplace expansion along the last row that laplace.py sketches,
//     det([A; r_m]) = sum_j r_m[j] C_j,   C_j = cofactor of (n, j) = det(M) (M^-1)[j][n]   for M = [A; r_rep],
// with the cofactors taken from the pivoted Gauss-Jordan inverse of the representative member's matrix
// (the cofactors of the last row do not depend on that row).
// One CTA per (super-block, ket block).  rows_rep: (n_super x n) index lists of the representatives;
// mem_off (n_super + 1), mem_block / mem_row: member block ids and their last-row indices; cols: (n_kblk x n).
// D (n_bblk x n_kblk, row stride ld_d) gets the members' determinants; flags (n_super x n_kblk) is set
// when the representative's matrix is singular or ill conditioned (pivot ratio below 1e-6): the caller
// recomputes those members by brute force.
__global__ void __launch_bounds__(LUB_THREADS)
cofactor_rows_kernel(const double *__restrict__ S, int ld_s, int n, const int32_t *__restrict__ rows_rep, int n_super,
                     const int32_t *__restrict__ mem_off, const int32_t *__restrict__ mem_block,
                     const int32_t *__restrict__ mem_row, const int32_t *__restrict__ cols, int n_kblk,
                     double *__restrict__ D, int ld_d, unsigned char *__restrict__ flags) {
    extern __shared__ __align__(16) double lub_tile[];    // n x LUB_LDS
    __shared__ double cof[LUB_MAXN];
    __shared__ int colv[LUB_MAXN];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (long long pair = blockIdx.x; pair < (long long)n_super * n_kblk; pair += gridDim.x) {
        const int sb = (int)(pair / n_kblk), kb = (int)(pair - (long long)sb * n_kblk);
        const int32_t *rl = rows_rep + (size_t)sb * n, *cl = cols + (size_t)kb * n;
        __syncthreads();   // the previous pair is done with the tile
        for (int idx = tid; idx < n * n; idx += LUB_THREADS) {
            const int r = idx / n, c = idx - r * n;
            lub_tile[r * LUB_LDS + c] = __ldg(S + (size_t)rl[r] * ld_s + cl[c]);
        }
        if (tid < n) colv[tid] = cl[tid];
        __syncthreads();
        double det, minp, maxp;
        int sing;
        lub_invert_tile(lub_tile, n, det, minp, maxp, sing);
        const bool bad = sing || !(minp > 1e-6 * maxp);
        if (tid == 0) flags[pair] = bad ? 1 : 0;
        if (bad) continue;   // uniform
        if (tid < n) cof[tid] = det * lub_tile[tid * LUB_LDS + (n - 1)];
        __syncthreads();
        // members: one warp per member, dot product of its last row with the cofactors
        for (int m = mem_off[sb] + warp; m < mem_off[sb + 1]; m += LUB_WARPS) {
            const double *srow = S + (size_t)mem_row[m] * ld_s;
            double acc = 0.0;
            for (int j = lane; j < n; j += 32) acc = fma(__ldg(srow + colv[j]), cof[j], acc);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) D[(size_t)mem_block[m] * ld_d + kb] = acc;
        }
    }
}

// one-time probe (wfo_smo_inv_dev): is a kernel that waits on the device for a flag reached by a kernel launched
// AFTER it on another stream?  (Not under profilers that serialise kernels, CUDA_LAUNCH_BLOCKING, ...)
__global__ void concurrency_probe_kernel(const int *flag, int value, int *result) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int seen = 0, ok = 0;
        for (int spin = 0; spin < 4096 && !ok; spin++) {
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(seen) : "l"(flag) : "memory");
            ok = (seen == value);
            if (!ok) __nanosleep(200);
        }
        *result = ok;
    }
}

__global__ void set_flag_kernel(int *flag, int value) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        __threadfence();
        asm volatile("st.release.gpu.global.s32 [%0], %1;\n" ::"l"(flag), "r"(value) : "memory");
    }
}

// verdict of a 2x2 block inverse: A = [[A11, A12], [A21, A22]], det(A) = det(A11) det(S), S the Schur complement
__global__ void combine_info_kernel(const BaseInfo *__restrict__ a, const BaseInfo *__restrict__ b, BaseInfo *__restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        BaseInfo r;
        r.det = a->det * b->det;
        r.min_piv = fmin(a->min_piv, b->min_piv);
        r.max_piv = fmax(a->max_piv, b->max_piv);
        r.singular = a->singular | b->singular;
        r.pad = 0;
        *out = r;
    }
}

// dst (rows x cols, ld_dst) = src (ld_src)
__global__ void copy2d_kernel(const double *__restrict__ src, int ld_src, double *__restrict__ dst, int ld_dst, int rows,
                              int cols) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)rows * cols) return;
    const int r = (int)(idx / cols), c = (int)(idx - (long long)r * cols);
    dst[(size_t)r * ld_dst + c] = src[(size_t)r * ld_src + c];
}

}  // namespace wfo
