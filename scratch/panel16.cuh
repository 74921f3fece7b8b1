// development copy of the 16-column panel factorisation (moved into det_lu_blocked.cuh when tuned)
#pragma once
#include "../pywfo_b200/csrc/common.cuh"
namespace wfo {
struct P16Shared {
    unsigned key[2][4];          // per-warp candidate keys, double buffered by column parity
    double cand[2][4][16];       // per-warp candidate rows
    int piv_row[16];
    int aff_dst[32], aff_src[32];
    int zero;
};
__device__ __forceinline__ void pbar(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }

// rows [j0, n), columns [j0, j0+16) (w valid columns), PW = ceil((n-j0)/32) <= 4 warps, one row per lane.
template <int PW>
__device__ __forceinline__ void factor_panel16(double *A, int LD, int n, int j0, int w, P16Shared *sh, int warp,
                                               int lane, double &det) {
    const unsigned FULL = 0xffffffffu;
    constexpr int NT = PW * 32;
    const int r = j0 + 32 * warp + lane;
    const bool valid = r < n;
    double a[16];
    {
        const double2 *src = reinterpret_cast<const double2 *>(A + (size_t)(valid ? r : j0) * LD + j0);
#pragma unroll
        for (int c = 0; c < 8; c++) { const double2 v = src[c]; a[2 * c] = v.x; a[2 * c + 1] = v.y; }
    }
#ifdef P16_TIMING
    long long tA = clock64();
#endif
    bool done = !valid;
    int pstep = -1;
    const unsigned tag = 0x80u | (unsigned)(warp << 5) | (unsigned)lane;
#pragma unroll
    for (int jj = 0; jj < 16; jj++) {
        if (jj < w) {
            const int b = jj & 1;
            const unsigned key = done ? 0u : ((hi_abs(a[jj]) & 0xffffff00u) | tag);
#ifdef NO_REDUX
            const unsigned kw = (hi_abs(a[jj]) & 0xffffff00u) | 0x80u | (unsigned)(jj & 31);
#else
            const unsigned kw = __reduce_max_sync(FULL, key);
#endif
            unsigned kmax = kw;
            double u[16];
            if (PW > 1) {
                if (key == kw && key != 0u) {            // this warp's candidate publishes its row
                    sh->key[b][warp] = kw;
#pragma unroll
                    for (int c = jj; c < 16; c++) sh->cand[b][warp][c] = a[c];
                }
                if (kw == 0u && lane == 0) sh->key[b][warp] = 0u;
                pbar(NT);
#pragma unroll
                for (int q = 0; q < PW; q++) kmax = max(kmax, sh->key[b][q]);
                const int ww = (kmax >> 5) & 3;
#pragma unroll
                for (int c = jj; c < 16; c++) u[c] = sh->cand[b][ww][c];
            } else {
                const int wl = kmax & 31;
#pragma unroll
#ifdef NO_SHFL
                for (int c = jj; c < 16; c++) u[c] = a[c] + wl;
#else
                for (int c = jj; c < 16; c++) u[c] = __shfl_sync(FULL, a[c], wl);
#endif
            }
            const double piv = u[jj];
#ifdef NO_RCP
            const double rcp = piv * 0.5;
#else
            const double rcp = fast_rcp(piv);
#endif
            det *= piv;
            const bool win = (key == kmax) && (key != 0u);
            if (win) sh->piv_row[jj] = r;
            pstep = win ? jj : pstep;
            done = done || win;
            // branch-free elimination: finished rows use a zero multiplier
            const double l = done ? 0.0 : a[jj] * rcp;
            a[jj] = done ? a[jj] : l;
#pragma unroll
            for (int c = jj + 1; c < 16; c++) a[c] = fma(-l, u[c], a[c]);
        }
    }
#ifdef P16_TIMING
    long long tB = clock64();
#endif
    if (PW > 1) pbar(NT); else __syncwarp();
    // final positions: pivot of step jj -> j0+jj; displaced rows (top w rows that are no pivots, all in
    // warp 0) -> positions vacated by pivots from below, in step order
    int newpos = (pstep >= 0) ? j0 + pstep : r;
    if (warp == 0) {
        const bool displaced = (lane < w) && pstep < 0 && valid;
        const unsigned dbal = __ballot_sync(FULL, displaced);
        const int drank = __popc(dbal & ((1u << lane) - 1));
        if (displaced) {
            int seen = 0;
#pragma unroll
            for (int q = 0; q < 16; q++) {
                const int pr = sh->piv_row[q];
                if (q < w && pr >= j0 + w) { if (seen == drank) newpos = pr; seen++; }
            }
        }
        if (lane < 16) {
            const int pr = (lane < w) ? sh->piv_row[lane] : -1;
            const bool moved = (lane < w) && pr != j0 + lane;
            sh->aff_dst[lane] = moved ? j0 + lane : -1;
            sh->aff_src[lane] = moved ? pr : -1;
            sh->aff_dst[16 + lane] = -1;
            sh->aff_src[16 + lane] = -1;
        }
        __syncwarp();
        if (displaced) { sh->aff_dst[16 + drank] = newpos; sh->aff_src[16 + drank] = r; }
    }
    if (valid) {
        double2 *dst = reinterpret_cast<double2 *>(A + (size_t)newpos * LD + j0);
#pragma unroll
        for (int c = 0; c < 8; c++) dst[c] = make_double2(a[2 * c], a[2 * c + 1]);
    }
#ifdef P16_TIMING
    if (threadIdx.x == 0) { sh->aff_dst[30] = (int)(tB - tA); sh->aff_dst[31] = (int)(clock64() - tB); }
#endif
}
}  // namespace wfo
