// T0 (brute force), 32 < n <= 128: one determinant per warp, LEFT-LOOKING blocked LU with partial
// pivoting, everything on chip, no block barrier, no global scratch.
//
// Why left-looking.  A right-looking LU rewrites the whole trailing matrix after every panel
// (16 bytes of traffic per element per panel) and needs the full n x n working matrix on chip:
// 80 KB at n = 100, two matrices per SM, or an L2/HBM scratch (round 1: 316 MB of DRAM writes per
// 4096 pairs).  Only the determinant is wanted here, so U is never needed again and of L only the
// rows below each panel.  The left-looking order reads every element of the minor exactly ONCE,
// straight from S_MO (L2 resident) through the row/column index lists, and keeps on chip only
//      M_q = L21 L11^-1   of every finished 8-column panel q   (rows below the panel),
// i.e. n^2/2 doubles: 40 KB at n = 100, 14 KB at n = 64, 5 KB at n = 40.  Five (n = 100) to
// sixteen independent factorisations per SM hide each other's pivot-search latency.
//
// Per 8-column panel p of one matrix (one warp):
//   1. the panel's columns, all rows in their current order, arrive in DMMA accumulator layout
//      (they were prefetched into registers while panel p-1 was being factored);
//   2. for q = 0 .. p-1:  R_q = rows of tile q (the raw pivot rows of panel q, updated by the
//      panels before q), then   tile t -= M_q[t] R_q   for every row tile t > q: DMMA.8x8x4 with
//      A fragments from shared memory (one LDS.128 per two DMMAs), B fragments from R_q;
//   3. the live rows (tiles >= p) go through shared memory into a row-per-lane layout; 8 column
//      steps with REDUX.MAX pivot search over (|a| high word, slot, lane), the pivot row and its
//      speculatively computed reciprocal go through shared memory, branch-free elimination;
//      rows do not move in registers: every row tracks the position LAPACK-style swaps would
//      give it;
//   4. M_p = L21 L11^-1 in the panel lanes (no triangular solve: the Schur complement is
//      A22 - M A12 with the RAW pivot rows), written negated to its final row position in the
//      k-interleaved, swizzled layout the A fragments are read in; rows that changed position
//      are moved in the older M_q as well (rare) and re-fetched for the prefetched panel.
// The determinant is the signed product of the pivots = np.linalg.det(S_MO[rows][:, cols])
// (pywfo/main.py:583-584).  The matrix is padded to 8 NT rows/columns with an identity block.
//
// Measured and rejected in round 2 (standalone driver scratch/ll_test.cu, n = 100 / 64): the A fragments of panel
// q+1 requested into a second register set while panel q is applied (needs 8 warps at 255 registers: -10 %); four
// unrolled column steps run twice over swapped register halves to halve the step code (runtime record addresses and a
// uniform branch per step: -22 % / -5 %).  Kept: predicated winner records instead of branches (+11 %), negative
// multipliers throughout, one row-slot instantiation for NT >= 10 (instruction fetch, see ll_lu_det), work items from a
// global counter (t0_fused_ll_kernel).
#pragma once
#include "common.cuh"

namespace wfo {

#ifdef WFO_LL_CLOCKS
__device__ long long g_ll_clk[8];
struct LLClk { long long acc[8]; long long prev; };
#define LL_T(i) do { long long now_ = clock64(); clk.acc[i] += now_ - clk.prev; clk.prev = now_; } while (0)
#else
struct LLClk {};
#define LL_T(i)
#endif

// MG: the M panels live in a per-warp GLOBAL scratch that stays L2 resident (n > 64: more matrices in flight per SM
// than shared memory could hold); shared memory then holds only the staging area of one panel.
template <int NT, bool MG = false>
struct LLCfg {
    static constexpr int ROWS = 8 * NT;
    static constexpr int M_TILES = NT * (NT - 1) / 2;
    // byte offsets inside the per-warp shared block
    static constexpr int OFF_TOP = 0;                      // 8 x 8: staging of the top rows of the live block
    static constexpr int OFF_RBUF = 512;                   // 2 x (8 x 8): R_q, double buffered
    static constexpr int OFF_REC = 1536;                   // 8 x 12 doubles: pivot row, reciprocal, position, staging row
    static constexpr int OFF_MOVES = 2304;                 // 16 x (from, to) bytes: rows that changed position
    static constexpr int OFF_ROFF = 2336;                  // row offsets into S by current position
    static constexpr int OFF_ROFFB = OFF_ROFF + 4 * ROWS;  // ... as they were when the next panel was prefetched
    static constexpr int OFF_COFF = OFF_ROFFB + 4 * ROWS;  // column offsets
    static constexpr int OFF_M = (OFF_COFF + 4 * ROWS + 15) / 16 * 16;
    static constexpr int BYTES = OFF_M + 512 * (MG ? NT : M_TILES);
};

__device__ __forceinline__ double2 lds128(const double *p) { return *reinterpret_cast<const double2 *>(p); }
__device__ __forceinline__ void sts128(double *p, double x, double y) { *reinterpret_cast<double2 *>(p) = make_double2(x, y); }
// predicated shared-memory stores without a branch (a branch ends the basic block the scheduler works in)
__device__ __forceinline__ void sts128_if(double *p, double x, double y, bool pred) {
    asm volatile("{\n.reg .pred q;\nsetp.ne.b32 q, %3, 0;\n@q st.shared.v2.f64 [%0], {%1, %2};\n}\n" ::"r"(
                     (unsigned)__cvta_generic_to_shared(p)), "d"(x), "d"(y), "r"((int)pred) : "memory");
}
__device__ __forceinline__ void sts64_if(double *p, double x, bool pred) {
    asm volatile("{\n.reg .pred q;\nsetp.ne.b32 q, %2, 0;\n@q st.shared.f64 [%0], %1;\n}\n" ::"r"(
                     (unsigned)__cvta_generic_to_shared(p)), "d"(x), "r"((int)pred) : "memory");
}
__device__ __forceinline__ void sts32_if(int *p, int x, bool pred) {
    asm volatile("{\n.reg .pred q;\nsetp.ne.b32 q, %2, 0;\n@q st.shared.b32 [%0], %1;\n}\n" ::"r"(
                     (unsigned)__cvta_generic_to_shared(p)), "r"(x), "r"((int)pred) : "memory");
}
// a 16-byte chunk of an M panel: shared memory, or the L2-resident global scratch (L1 bypassed; prefetching the
// next panel into L1 while the current one is applied was measured: 5 % slower)
template <bool MG>
__device__ __forceinline__ double2 ldM(const double *p) {
    return MG ? __ldcg(reinterpret_cast<const double2 *>(p)) : *reinterpret_cast<const double2 *>(p);
}

// element (row position, column c) of the padded minor; roff < 0 marks pad row -(roff+1), coff < 0 a pad column
__device__ __forceinline__ double ll_elem(const double *__restrict__ S, int roff, int coff, int c) {
    if (roff >= 0 && coff >= 0) return __ldg(S + (size_t)roff + coff);
    return (roff < 0 && c == -(roff + 1)) ? 1.0 : 0.0;
}

// Columns [8P, 8P+8) of all rows in their current order, in DMMA accumulator layout (tile t: row 8t+g,
// columns 2tq, 2tq+1).  Pad rows exist in the last tile only and never move.
template <int NT>
__device__ __forceinline__ void ll_load_panel(const double *__restrict__ S, const int *roff, const int *coff, int P, int g,
                                              int tq, double (&out)[NT][2]) {
    const int c0 = 8 * P + 2 * tq;
    const int co0 = coff[c0], co1 = coff[c0 + 1];
#pragma unroll
    for (int t = 0; t < NT - 1; t++) {
        const double *src = S + (size_t)roff[8 * t + g];
        out[t][0] = (co0 >= 0) ? __ldg(src + co0) : 0.0;
        out[t][1] = (co1 >= 0) ? __ldg(src + co1) : 0.0;
    }
    const int r = roff[8 * (NT - 1) + g];
    out[NT - 1][0] = ll_elem(S, r, co0, c0);
    out[NT - 1][1] = ll_elem(S, r, co1, c0 + 1);
}

// Panel factorisation in the row-per-lane layout (SL row slots per lane), the 8 column steps unrolled with
// their exact widths; everything around it (left-looking update, row moves) is loop code, and only two
// slot counts are instantiated per matrix size, so that the kernel stays inside the 32 KB instruction
// cache level.  (Measured: with every loop unrolled and four slot counts, 125 KB of code, the warps
// spent 6 of 10 cycles per instruction in 'no instruction' stalls.)
// Returns 0 for a singular minor, 1 when no row changed position, 2 when rows moved.
template <int NT, int SL, bool MG>
__device__ __forceinline__ int ll_factor(char *sm, double *gM, int p, int n, int lane, double &det, unsigned &par, LLClk &clk) {
    using C = LLCfg<NT, MG>;
    const unsigned FULL = 0xffffffffu;
    const int L = 8 * (NT - p);   // live rows
    double *smd = reinterpret_cast<double *>(sm);
    double *recs = reinterpret_cast<double *>(sm + C::OFF_REC);
    unsigned char *moves = reinterpret_cast<unsigned char *>(sm + C::OFF_MOVES);
    int *roff = reinterpret_cast<int *>(sm + C::OFF_ROFF);
    // staging rows of the live block below the top tile: the future M_p (shared-memory variant) or the staging area
    const int mp_off = MG ? C::OFF_M / 8 : C::OFF_M / 8 + (p * (NT - 1) - p * (p - 1) / 2) * 64;
    double *Mp_dst = MG ? gM + (size_t)(p * (NT - 1) - p * (p - 1) / 2) * 64 : smd + mp_off;

    double a[SL][8];
    int rpos[SL], myroff[SL], rowoff[SL];
    bool alive[SL];
#pragma unroll
    for (int s = 0; s < SL; s++) {
        const int lam = lane + 32 * s;
        const bool valid = lam < L;
        rowoff[s] = (lam < 8 || !valid) ? C::OFF_TOP / 8 + (lam & 7) * 8 : mp_off + (lam - 8) * 8;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const double2 v = lds128(smd + rowoff[s] + 2 * c);
            a[s][2 * c] = v.x;
            a[s][2 * c + 1] = v.y;
        }
        rpos[s] = lam;
        alive[s] = valid;
        myroff[s] = roff[8 * p + (valid ? lam : 0)];
    }
    __syncwarp();   // the staging rows have been read; from here on a row's staging row holds its multipliers
    LL_T(3);

    const int nsteps = n - 8 * p;   // < 8 in a ragged last panel: the identity padding needs no elimination
    bool bad = false;
#pragma unroll
    for (int jj = 0; jj < 8; jj++) {
        if (jj >= nsteps) break;
        unsigned key = 0;
#pragma unroll
        for (int s = 0; s < SL; s++) {
            const unsigned k = (hi_abs(a[s][jj]) & 0xffffff00u) | 0x80u | (unsigned)(s << 5) | (unsigned)lane;
            key = max(key, alive[s] ? k : 0u);
        }
        const unsigned kmax = __reduce_max_sync(FULL, key);
        bad |= (kmax == 0u);   // no candidate left (checked after the panel: no branch inside the steps)
        // reciprocal of this lane's own best candidate, computed in the shadow of the reduction
        const int bs = (key >> 5) & 3;
        double cand = a[0][jj];
#pragma unroll
        for (int s = 1; s < SL; s++) cand = (bs == s) ? a[s][jj] : cand;
        const double rc_own = fast_rcp(cand);
        const int wl = kmax & 31, ws = (kmax >> 5) & 3;
        double *rec = recs + 12 * jj;   // [0, jj): multipliers of the pivot row = row jj of L11; [jj, 8): the pivot row
#pragma unroll
        for (int s = 0; s < SL; s++) {
            const bool win = (lane == wl) && (ws == s);
#pragma unroll
            for (int c = 0; c < 4; c++) sts128_if(rec + 2 * c, a[s][2 * c], a[s][2 * c + 1], win);
            sts64_if(rec + 8, rc_own, win);
            sts32_if(reinterpret_cast<int *>(rec + 9), rpos[s], win);
        }
        __syncwarp();
        double u[8];
#pragma unroll
        for (int c = jj; c < 8; c++) u[c] = rec[c];
        const double rcp = rec[8], nrcp = -rcp;
        const int pp = reinterpret_cast<const int *>(rec + 9)[0];
        bad |= (u[jj] == 0.0);   // the whole column is zero
        det *= u[jj];
        par ^= (pp != jj) ? 1u : 0u;
#pragma unroll
        for (int s = 0; s < SL; s++) {
            const bool win = (lane == wl) && (ws == s);
            rpos[s] = (alive[s] && !win && rpos[s] == jj) ? pp : rpos[s];   // the row the pivot displaces
            rpos[s] = win ? jj : rpos[s];
            alive[s] = alive[s] && !win;
            // the NEGATIVE multiplier -l is what every later use wants (the update, the recurrence for M below, the
            // stored -M): a negation would cost an FP64 issue slot each time
            const double nl = alive[s] ? a[s][jj] * nrcp : 0.0;
            a[s][jj] = alive[s] ? nl : a[s][jj];
#pragma unroll
            for (int c = jj + 1; c < 8; c++) a[s][c] = fma(nl, u[c], a[s][c]);
        }
    }
    LL_T(4);
    if (bad) return 0;           // uniform: singular minor
    if (p == NT - 1) return 1;   // last panel: nothing below it

    __syncwarp();   // the records are complete
    // ---- -M = -L21 L11^-1 (axpy form of x L11 = l, linear, so it holds for the negated rows as they are kept);
    // recs[q][c], c < q, is -L11[q][c] (the winner's row was recorded with its negative multipliers);
    // dead rows are modified too, harmlessly
#pragma unroll
    for (int q = 7; q >= 1; q--) {
#pragma unroll
        for (int c = 0; c < q; c++) {
            const double lqc = recs[12 * q + c];
#pragma unroll
            for (int s = 0; s < SL; s++) a[s][c] = fma(a[s][q], lqc, a[s][c]);
        }
    }
    __syncwarp();   // all staging rows consumed: M_p may now overwrite them
    // ---- -M to its final row, k-interleaved (chunk t = (k = t, k = t + 4)) and swizzled for the A-fragment reads
    bool moved[SL];
    bool any = false;
#pragma unroll
    for (int s = 0; s < SL; s++) {
        const int lam = lane + 32 * s;
        if (alive[s]) {
            double *dst = Mp_dst + (rpos[s] - 8) * 8;
            const int sw = (rpos[s] >> 1) & 3;
#pragma unroll
            for (int c = 0; c < 4; c++) sts128(dst + ((c ^ sw) << 1), a[s][c], a[s][c + 4]);
        }
        moved[s] = lam < L && rpos[s] != lam;
        any = any || moved[s];
    }
    LL_T(5);
    int ret = 1;
    if (__any_sync(FULL, any)) {
        // rows that changed position: new entry in the position table, and a (from, to) record for the
        // finished panels and the prefetched next panel (handled by the caller)
        int base = 0;
#pragma unroll
        for (int s = 0; s < SL; s++) {
            const unsigned bal = __ballot_sync(FULL, moved[s]);
            if (moved[s]) {
                const int idx = base + __popc(bal & ((1u << lane) - 1u));
                moves[2 * idx] = (unsigned char)(lane + 32 * s);
                moves[2 * idx + 1] = (unsigned char)rpos[s];
                roff[8 * p + rpos[s]] = myroff[s];
            }
            base += __popc(bal);
        }
        ret = 2 + base;
    }
    __syncwarp();   // M_p and the tables are visible to the whole warp
    return ret;
}

// whole factorisation by one warp; returns det to every lane.  sm: LLCfg<NT>::BYTES of shared memory
// owned by this warp.
template <int NT, bool MG>
__device__ __forceinline__ double ll_lu_det(const double *__restrict__ S, int ld_s, const int32_t *__restrict__ rows,
                                            const int32_t *__restrict__ cols, int n, char *sm, double *gM, int lane) {
    using C = LLCfg<NT, MG>;
    double *top = reinterpret_cast<double *>(sm + C::OFF_TOP);
    double *rbuf = reinterpret_cast<double *>(sm + C::OFF_RBUF);
    const unsigned char *moves = reinterpret_cast<const unsigned char *>(sm + C::OFF_MOVES);
    int *roff = reinterpret_cast<int *>(sm + C::OFF_ROFF);
    int *roffb = reinterpret_cast<int *>(sm + C::OFF_ROFFB);
    int *coff = reinterpret_cast<int *>(sm + C::OFF_COFF);
    double *Mall = MG ? gM : reinterpret_cast<double *>(sm + C::OFF_M);
    double *stage = reinterpret_cast<double *>(sm + C::OFF_M);   // MG only
    const int g = lane >> 2, tq = lane & 3;
    const int swz = (g >> 1) & 3;
    LLClk clk;
#ifdef WFO_LL_CLOCKS
    for (int i = 0; i < 8; i++) clk.acc[i] = 0;
    clk.prev = clock64();
#endif

    __syncwarp();
    for (int i = lane; i < C::ROWS; i += 32) {
        const int r = (i < n) ? rows[i] * ld_s : -(i + 1);
        roff[i] = r;
        roffb[i] = r;
        coff[i] = (i < n) ? cols[i] : -1;
    }
    __syncwarp();

    double acc[NT][2], nxt[NT][2];
    double det = 1.0;
    unsigned par = 0;
    int ok = 1;
#pragma unroll 1
    for (int p = -1; p < NT; p++) {
        if (p >= 0) {
#pragma unroll
            for (int t = 0; t < NT; t++) {
                acc[t][0] = nxt[t][0];
                acc[t][1] = nxt[t][1];
            }
        }
        if (p + 1 < NT) ll_load_panel<NT>(S, roff, coff, p + 1, g, tq, nxt);   // prefetch, row order as it is now
        if (p < 0) continue;
        LL_T(0);
        // ---- left-looking update with the finished panels
        // (unrolled over q with an early exit: every (q, t) block is three instructions with static
        // registers; as a loop over q with predicated tiles it cost 2.5x the issue slots)
#pragma unroll
        for (int q = 0; q < NT - 1; q++) {
            if (q >= p) break;
            double *rb = rbuf + 64 * (q & 1);
            sts128(rb + g * 8 + 2 * tq, acc[q][0], acc[q][1]);
            __syncwarp();
            const double b0 = rb[tq * 8 + g], b1 = rb[(tq + 4) * 8 + g];
            const double *Mq = Mall + (q * (NT - 1) - q * (q - 1) / 2) * 64 + g * 8 + ((tq ^ swz) << 1);
            double2 af[NT];
#pragma unroll
            for (int t = q + 1; t < NT; t++)
                af[t] = ldM<MG>(Mq + (t - q - 1) * 64);
#pragma unroll
            for (int t = q + 1; t < NT; t++) dmma884(acc[t][0], acc[t][1], af[t].x, b0);
#pragma unroll
            for (int t = q + 1; t < NT; t++) dmma884(acc[t][0], acc[t][1], af[t].y, b1);
        }
        LL_T(1);
        // ---- live rows to the row-per-lane layout through the staging area (top tile + future M_p)
        {
            double *Mp = MG ? stage - (size_t)(p + 1) * 64 : Mall + (size_t)(p * (NT - 1) - p * (p - 1) / 2 - p - 1) * 64;
#pragma unroll
            for (int t = 0; t < NT; t++) {
                if (t >= p) {
                    double *dst = (t == p) ? top + g * 8 : Mp + (size_t)(t * 8 + g) * 8;
                    sts128(dst + 2 * tq, acc[t][0], acc[t][1]);
                }
            }
        }
        __syncwarp();
        LL_T(2);
        const int L = 8 * (NT - p);
        // NT >= 10: ONE slot count.  The kernel is instruction-fetch bound there (ncu, n = 100: 39 % of the stall
        // samples are 'no instruction' with two instantiations, 92 KB of code); the late panels then carry idle row
        // slots, and it is still 5-12 % faster (n = 80 ... 128).  Below, two slot counts fit (n = 64: 52 KB, 5 %).
        constexpr int SLMAX = (8 * NT + 31) / 32, SLLO = (NT >= 10) ? SLMAX : (SLMAX + 1) / 2;
        if (L > 32 * SLLO) ok = ll_factor<NT, SLMAX, MG>(sm, gM, p, n, lane, det, par, clk);
        else ok = ll_factor<NT, SLLO, MG>(sm, gM, p, n, lane, det, par, clk);
        if (ok == 0) break;
        if (ok >= 2) {
            // rows moved: the same moves in the finished panels M_0 .. M_{p-1} ...
            const int nm = ok - 2;   // <= 16
            const int m0 = lane >> 2, c = lane & 3;
            for (int q = 0; q < p; q++) {
                double *Mq = Mall + (size_t)(q * (NT - 1) - q * (q - 1) / 2 + (p - q - 1)) * 64;   // row 0 = position 8p
                double2 v0 = make_double2(0.0, 0.0), v1 = v0;
                int d0 = 0, d1 = 0;
                if (m0 < nm) {
                    const int from = moves[2 * m0];
                    d0 = moves[2 * m0 + 1];
                    v0 = ldM<MG>(Mq + from * 8 + ((c ^ ((from >> 1) & 3)) << 1));
                }
                if (m0 + 8 < nm) {
                    const int from = moves[2 * (m0 + 8)];
                    d1 = moves[2 * (m0 + 8) + 1];
                    v1 = ldM<MG>(Mq + from * 8 + ((c ^ ((from >> 1) & 3)) << 1));
                }
                __syncwarp();
                if (m0 < nm) sts128(Mq + d0 * 8 + ((c ^ ((d0 >> 1) & 3)) << 1), v0.x, v0.y);
                if (m0 + 8 < nm) sts128(Mq + d1 * 8 + ((c ^ ((d1 >> 1) & 3)) << 1), v1.x, v1.y);
                __syncwarp();
            }
            // ... and in the prefetched next panel: re-fetch the positions whose row changed
            const int c0 = 8 * (p + 1) + 2 * tq;
            const int co0 = coff[c0], co1 = coff[c0 + 1];
#pragma unroll
            for (int t = 0; t < NT; t++) {
                if (t >= p) {
                    const int rn = roff[8 * t + g], ro = roffb[8 * t + g];
                    if (rn != ro) {
                        nxt[t][0] = ll_elem(S, rn, co0, c0);
                        nxt[t][1] = ll_elem(S, rn, co1, c0 + 1);
                        roffb[8 * t + g] = rn;
                    }
                }
            }
            __syncwarp();
        }
        LL_T(6);
    }
#ifdef WFO_LL_CLOCKS
    if (lane == 0)
        for (int i = 0; i < 8; i++) atomicAdd((unsigned long long *)&g_ll_clk[i], (unsigned long long)clk.acc[i]);
#endif
    return ok ? (par ? -det : det) : 0.0;
}

// Warps per CTA.  The warps are independent of each other, but the hardware maps warp w of a CTA to
// scheduler w % 4: single-warp CTAs would all sit on scheduler 0 (measured: 4-5x slower).  So ONE
// persistent CTA per SM with as many warps as shared memory (227 KB) and registers allow.
template <int NT, bool MG = false>
struct LLWarps {
    static constexpr int SMEM = (227 * 1024) / LLCfg<NT, MG>::BYTES;
    // register budget 65536 / (32 w).  L2-scratch variant: 12 warps (168 registers, a few spilled) beat 8 up to
    // n = 104 (n = 72: 5.6 -> 6.9 TFLOP/s, n = 88: 5.9 -> 7.2, n = 100: 5.0 -> 5.4), 8 warps win at n = 128
    static constexpr int REG = MG ? (NT <= 13 ? 12 : 8) : (NT >= 13 ? 8 : (NT >= 9 ? 10 : (NT >= 7 ? 12 : 16)));
#ifdef LL_WARPS_OVERRIDE
    static constexpr int W = LL_WARPS_OVERRIDE;
#else
    static constexpr int W = SMEM < REG ? SMEM : REG;
#endif
};
inline size_t ll_scratch_doubles(int nt) { return (size_t)nt * (nt - 1) / 2 * 64; }   // per warp, MG variant

template <int NT, bool MG>
__global__ void __launch_bounds__(LLWarps<NT, MG>::W * 32, 1)
det_table_ll_kernel(const double *__restrict__ S, int ld_s, int n, const int32_t *__restrict__ rows, int kb,
                    const int32_t *__restrict__ cols, int kk, double *__restrict__ D, double *gscratch) {
    extern __shared__ __align__(16) char ll_smem[];
    constexpr int W = LLWarps<NT, MG>::W;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    char *sm = ll_smem + (size_t)warp * LLCfg<NT, MG>::BYTES;
    const long long gw = (long long)blockIdx.x * W + warp, nw = (long long)gridDim.x * W;
    double *gM = MG ? gscratch + (size_t)gw * (NT * (NT - 1) / 2 * 64) : nullptr;
    const long long npairs = (long long)kb * kk;
    for (long long pair = gw; pair < npairs; pair += nw) {
        const int k = (int)(pair / kk), l = (int)(pair - (long long)k * kk);
        const double det = ll_lu_det<NT, MG>(S, ld_s, rows + (size_t)k * n, cols + (size_t)l * n, n, sm, gM, lane);
        if (lane == 0) D[pair] = det;
    }
}

// one work item per warp = (bra determinant k, chunk of ket determinants); g[j] += D(k,l) CkT[l][j]
template <int NT, bool MG>
__global__ void __launch_bounds__(LLWarps<NT, MG>::W * 32, 1)
t0_fused_ll_kernel(const double *__restrict__ S, int ld_s, int n, const int32_t *__restrict__ rows, int ksh,
                   const int32_t *__restrict__ cols, int kk, const double *__restrict__ ckt, int ldg, int chunk_len,
                   int n_chunks, double *__restrict__ gpart, double *gscratch, unsigned long long *next_item) {
    extern __shared__ __align__(16) char ll_smem[];
    constexpr int W = LLWarps<NT, MG>::W;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    char *sm = ll_smem + (size_t)warp * LLCfg<NT, MG>::BYTES;
    const long long gw = (long long)blockIdx.x * W + warp, nw = (long long)gridDim.x * W;
    double *gM = MG ? gscratch + (size_t)gw * (NT * (NT - 1) / 2 * 64) : nullptr;
    const long long nitems = (long long)ksh * n_chunks;
    // the first item of a warp is its index, further ones come from a global counter (starts at 0 = the warps'
    // first round): pivoting makes the determinants take different times, and a static round robin left the last
    // round a quarter empty on small tables
    auto fetch = [&]() -> long long {
        unsigned long long v = 0;
        if (lane == 0) v = atomicAdd(next_item, 1ull);
        return nw + (long long)__shfl_sync(0xffffffffu, v, 0);
    };
    for (long long item = gw; item < nitems; item = fetch()) {
        const int k = (int)(item / n_chunks), ch = (int)(item - (long long)k * n_chunks);
        const int l0 = ch * chunk_len, l1 = min(kk, l0 + chunk_len);
        double g0 = 0.0, g1 = 0.0;
        for (int l = l0; l < l1; l++) {
            const double det = ll_lu_det<NT, MG>(S, ld_s, rows + (size_t)k * n, cols + (size_t)l * n, n, sm, gM, lane);
            if (lane < ldg) g0 = fma(det, __ldg(ckt + (size_t)l * ldg + lane), g0);
            if (lane + 32 < ldg) g1 = fma(det, __ldg(ckt + (size_t)l * ldg + lane + 32), g1);
        }
        double *dst = gpart + ((size_t)ch * ksh + k) * ldg;
        if (lane < ldg) dst[lane] = g0;
        if (lane + 32 < ldg) dst[lane + 32] = g1;
    }
}

}  // namespace wfo
