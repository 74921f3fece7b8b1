// Pivoted device inverse: the fallback behind wfo_inverse_dev for MO matrices Newton-Schulz cannot
// invert in FP64 (cond >~ 1e7), where the reference's np.linalg.inv (LAPACK getrf/getri, partial
// pivoting, pywfo/main.py:303) still returns a result.
//
// Gauss-Jordan elimination with partial (row) pivoting on the augmented pair (W | V) = (A | I), rows
// dealt round-robin over the CTAs of ONE cooperative grid, one grid-wide barrier per column:
//   * implicit pivoting: rows never move; pivot k is the unused row with the largest |W[i][k]|;
//   * every CTA eliminates column k from its own rows with the pivot row (read-only for everybody but
//     nobody: its owner skips it), then at once proposes its candidate for column k+1 from the rows it
//     has just updated -- so one barrier per step is enough;
//   * pivot rows are never normalised (a row scaled by its pivot stays consistent under later row
//     operations); the scale is divided out when the inverse is written:  Ainv[k][:] = V[p_k][:] / d_k.
// A zero / non-finite pivot column reports the matrix as singular (numpy raises LinAlgError there).
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"
#include "lu_base.cuh"   // BaseInfo

namespace wfo {

struct GjCand {
    double val;   // |W[i][k]| of the CTA's best unused row, -1 if it has none
    int row;
    int pad;
};

struct GjState {
    int singular;
    int pad;
};

constexpr int GJ_THREADS = 256;

// W: n x n copy of A (destroyed), V: n x n (overwritten), out: n x n.
// piv: n ints, dval: n doubles, cand: gridDim.x entries, used: n bytes (zeroed by the kernel).
__global__ void __launch_bounds__(GJ_THREADS)
gj_inverse_kernel(double *W, double *V, int n, double *__restrict__ out, int *piv, double *dval, GjCand *cand,
                  unsigned char *used, GjState *state) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    const int G = gridDim.x, b = blockIdx.x, tid = threadIdx.x;
    __shared__ double s_val[GJ_THREADS];
    __shared__ int s_row[GJ_THREADS];
    __shared__ int s_p;
    __shared__ double s_d;

    // V = I, used = 0
    for (int i = b; i < n; i += G) {
        for (int j = tid; j < n; j += GJ_THREADS) V[(size_t)i * n + j] = (i == j) ? 1.0 : 0.0;
        if (tid == 0) used[i] = 0;
    }
    if (b == 0 && tid == 0) state->singular = 0;
    __syncthreads();

    auto propose = [&](int k) {   // this CTA's best unused row for column k
        double bv = -1.0;
        int br = -1;
        for (int i = b + G * tid; i < n; i += G * GJ_THREADS) {
            if (!used[i]) {
                const double v = fabs(W[(size_t)i * n + k]);
                const double vv = (v == v) ? v : 1e308;   // a NaN wins: reported as singular below
                if (vv > bv) { bv = vv; br = i; }
            }
        }
        s_val[tid] = bv;
        s_row[tid] = br;
        __syncthreads();
        for (int s = GJ_THREADS / 2; s > 0; s >>= 1) {
            if (tid < s) {
                const double ov = s_val[tid + s];
                const int orow = s_row[tid + s];
                if (ov > s_val[tid] || (ov == s_val[tid] && orow >= 0 && (s_row[tid] < 0 || orow < s_row[tid]))) {
                    s_val[tid] = ov;
                    s_row[tid] = orow;
                }
            }
            __syncthreads();
        }
        if (tid == 0) {
            cand[b].val = s_val[0];
            cand[b].row = s_row[0];
        }
    };

    propose(0);
    for (int k = 0; k < n; k++) {
        grid.sync();
        // everybody picks the same winner: largest value, smallest row on ties
        {
            double bv = -1.0;
            int br = -1;
            for (int c = tid; c < G; c += GJ_THREADS) {
                const double v = cand[c].val;
                const int r = cand[c].row;
                if (r >= 0 && (v > bv || (v == bv && r < br))) { bv = v; br = r; }
            }
            s_val[tid] = bv;
            s_row[tid] = br;
            __syncthreads();
            for (int s = GJ_THREADS / 2; s > 0; s >>= 1) {
                if (tid < s) {
                    const double ov = s_val[tid + s];
                    const int orow = s_row[tid + s];
                    if (orow >= 0 && (ov > s_val[tid] || (ov == s_val[tid] && (s_row[tid] < 0 || orow < s_row[tid])))) {
                        s_val[tid] = ov;
                        s_row[tid] = orow;
                    }
                }
                __syncthreads();
            }
            if (tid == 0) {
                const int p = s_row[0];
                const double d = p >= 0 ? W[(size_t)p * n + k] : 0.0;
                s_p = (p >= 0 && d != 0.0 && fabs(d) < 1e300 && d == d) ? p : -1;
                s_d = d;
            }
            __syncthreads();
        }
        const int p = s_p;
        if (p < 0) {   // uniform over the grid: every CTA saw the same candidates
            if (b == 0 && tid == 0) state->singular = 1;
            return;
        }
        const double d = s_d, rd = 1.0 / d;
        if (b == 0 && tid == 0) { piv[k] = p; dval[k] = d; }
        if (p % G == b && tid == 0) used[p] = 1;
        const double *Wp = W + (size_t)p * n, *Vp = V + (size_t)p * n;
        // eliminate column k from this CTA's rows (all rows but the pivot row, used ones included)
        for (int i = b; i < n; i += G) {
            if (i == p) continue;
            double *Wi = W + (size_t)i * n, *Vi = V + (size_t)i * n;
            const double f = Wi[k] * rd;
            __syncthreads();   // everyone has read Wi[k] before it is zeroed
            if (f != 0.0) {
                for (int j = k + 1 + tid; j < n; j += GJ_THREADS) Wi[j] = fma(-f, Wp[j], Wi[j]);
                for (int j = tid; j < n; j += GJ_THREADS) Vi[j] = fma(-f, Vp[j], Vi[j]);
            }
            if (tid == 0) Wi[k] = 0.0;
        }
        __syncthreads();   // this CTA's rows (and its used[] flag) are up to date
        if (k + 1 < n) propose(k + 1);
    }
    grid.sync();
    for (int k = b; k < n; k += G) {
        const int p = piv[k];
        const double rd = 1.0 / dval[k];
        for (int j = tid; j < n; j += GJ_THREADS) out[(size_t)k * n + j] = V[(size_t)p * n + j] * rd;
    }
}

// verdict of a pivoted inverse in the form the rank-update tier wants (BaseInfo: det with the sign of the row
// permutation k -> piv[k], min / max |pivot|, singular flag).  One CTA.
__global__ void __launch_bounds__(256)
gj_info_kernel(const int *__restrict__ piv, const double *__restrict__ dval, int n, const GjState *__restrict__ state,
               BaseInfo *__restrict__ info) {
    __shared__ double s_det[256], s_min[256], s_max[256];
    __shared__ int s_inv[256];
    const int tid = threadIdx.x;
    const bool sing = state->singular != 0;
    double det = 1.0, mn = 1e300, mx = 0.0;
    int inv = 0;
    if (!sing) {
        for (int k = tid; k < n; k += 256) {
            const double d = dval[k];
            det *= d;
            mn = fmin(mn, fabs(d));
            mx = fmax(mx, fabs(d));
            const int pk = piv[k];
            for (int j = k + 1; j < n; j++) inv += piv[j] < pk;
        }
    }
    s_det[tid] = det; s_min[tid] = mn; s_max[tid] = mx; s_inv[tid] = inv & 1;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) {
            s_det[tid] *= s_det[tid + o];
            s_min[tid] = fmin(s_min[tid], s_min[tid + o]);
            s_max[tid] = fmax(s_max[tid], s_max[tid + o]);
            s_inv[tid] ^= s_inv[tid + o];
        }
        __syncthreads();
    }
    if (tid == 0) {
        info->det = sing ? 0.0 : (s_inv[0] ? -s_det[0] : s_det[0]);
        info->min_piv = s_min[0];
        info->max_piv = s_max[0];
        info->singular = sing ? 1 : 0;
        info->pad = 0;
    }
}

}  // namespace wfo