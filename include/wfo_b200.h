/*
 * wfo_b200.h -- C ABI of the B200-native CI wavefunction-overlap hot path.
 *
 * pywfo (the reference) has no FFI: its boundary is the Python signature
 *   overlaps | overlaps_naive | overlaps_most_naive | overlaps_cache
 *       (bra_mos, ket_mos, bra_ci, ket_ci, occ, ci_thresh, ao_ovlps)   pywfo/main.py:81,314,422,520
 *   moovlp | moovlp_expl | moovlp_dots (mos1, mos2, S_AO)              pywfo/main.py:27,46,72
 *   dets.wfoverlap(bra_mos, ket_mos, bra_ci, ket_ci, ci_thresh, S_AO)   pywfo/dets.py:201
 * pywfo_b200/main.py and pywfo_b200/dets.py keep those signatures and bind the
 * functions below through ctypes (INTEGRATION.md shows the stub a pywfo
 * maintainer would add).  Each entry point names the reference lines it replaces.
 *
 * Conventions
 *   - all matrices row-major float64, index tables int32
 *   - an excitation is the pair (from, to): from in [0, n_occ), to in [0, n_virt)
 *     relative to n_occ; its MO index list is [0..n_occ-1] minus from, then
 *     n_occ+to APPENDED (pywfo/main.py:565-576).  (-1,-1) is the ground state.
 *   - *_dev functions take DEVICE pointers and a cudaStream_t (as void*), run
 *     asynchronously on that stream and never copy to/from the host
 *   - *_host functions take HOST pointers, do the H2D/D2H copies themselves
 *     and return after the result is in host memory
 *   - return value: 0 on success, non-zero on error; wfo_last_error() then
 *     holds a message (thread-local).  There is no CPU fallback: without a
 *     CUDA device every compute entry point fails.
 */
#ifndef WFO_B200_H
#define WFO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WFO_OK 0
#define WFO_ERR_CUDA 1
#define WFO_ERR_ARG 2
#define WFO_ERR_SINGULAR 3 /* base block S_MO[:occ,:occ] singular: tier 1/2 not applicable */
#define WFO_ERR_EMPTY_STATE 4 /* pywfo.main.overlaps semantics: a state keeps no excitation (ValueError upstream) */
#define WFO_ERR_NOCONV 5 /* device inverse did not converge (singular / extremely ill-conditioned MO matrix; LinAlgError in Python) */
#define WFO_ERR_COMM 6 /* in-library collective: rendezvous failed, no peer access, or a rank did not take part */

/* where the inverse of the MO matrix in S_AO = C^-1 C^-T (pywfo/main.py:301-311) comes from */
#define WFO_AO_GIVEN 0 /* the caller supplies c_inv (e.g. np.linalg.inv, exactly as upstream) */
#define WFO_AO_BRA 1   /* ao_ovlps="bra": bra_mos is inverted on the device */
#define WFO_AO_KET 2   /* ao_ovlps="ket": ket_mos is inverted on the device */

/* algorithm tiers (identical results to ~1e-13).  n_occ <= 128 for the brute-force tier,
 * n_occ <= 4096 for the rank-update and closed-form tiers (base block inverted blockwise above 128; when that
 * fails on a block LAPACK would invert -- pivoting is needed across the 128-blocks, e.g. occupied orbitals reordered
 * between the geometries -- the rank-update tier falls back to a pivoted Gauss-Jordan inverse of the whole block). */
#define WFO_TIER_LU 0      /* T0: one partially pivoted LU per determinant pair          */
#define WFO_TIER_UPDATE 1  /* T1: one base LU + rank-1 row/column update per pair        */
#define WFO_TIER_CLOSED 2  /* T2: closed-form contraction (GEMMs only), pair determinants never formed; opt-in */
#define WFO_TIER_AUTO (-1) /* T1, falling back to T0 when the base block is singular      */

typedef struct wfo_ctx wfo_ctx;

/* ---- housekeeping -------------------------------------------------------- */
int wfo_version(void);
const char *wfo_last_error(void);
int wfo_device_count(void);
/* One context per (process, device): scratch memory, SM count, a private stream
 * for the *_host entry points. */
int wfo_create(int device, wfo_ctx **out);
int wfo_destroy(wfo_ctx *ctx);
/* kernels launched by this context since creation (bench.py's gpu_launches) */
long long wfo_launch_count(const wfo_ctx *ctx);

/* ---- stage 1: MO overlap  S_MO = C_bra S_AO C_ket^T ------------------------
 * replaces pywfo/main.py:532 (and moovlp main.py:27-43, moovlp_expl :46-68,
 * moovlp_dots :71-78, dets.py:264).  bra_mos (p x u), s_ao (u x v), ket_mos
 * (q x v) -> s_mo (p x q).  FP64 tensor-core (DMMA) GEMM chain; `work` is a
 * device scratch of p*v doubles (may be NULL: taken from the context). */
int wfo_smo_dev(wfo_ctx *ctx, const double *bra_mos, const double *s_ao, const double *ket_mos,
                int p, int u, int v, int q, double *s_mo, double *work, void *stream);
int wfo_smo_host(wfo_ctx *ctx, const double *bra_mos, const double *s_ao, const double *ket_mos,
                 int p, int u, int v, int q, double *s_mo);
/* S_AO = Cinv Cinv^T, replaces the product in pywfo/main.py:304 (the inverse itself:
 * wfo_inverse_dev below, or any inverse the caller already has). */
int wfo_sao_dev(wfo_ctx *ctx, const double *c_inv, int n, double *s_ao, void *stream);
/* S_MO from the inverse behind S_AO in two launches: S_MO = (bra Cinv)(ket Cinv)^T -- the same product as
 * pywfo/main.py:304 + :532, associated so that S_AO is never materialised (work: 2 n*n doubles or NULL). */
int wfo_smo_inv_dev(wfo_ctx *ctx, const double *bra_mos, const double *ket_mos, const double *c_inv, int n,
                    double *s_mo, double *work, void *stream);
/* Optional: tell the context how many occupied orbitals the next state-overlap call will use (0 = off, the default).
 * wfo_smo_inv_dev then forms the base block S_MO[:n_occ,:n_occ] (pywfo/main.py:548) first and inverts it on a side
 * stream while the rest of S_MO is still being multiplied; the next wfo_state_overlaps_dev call on THAT s_mo buffer
 * with that n_occ (same stream, s_mo untouched in between) picks the inverse up instead of computing it behind
 * S_MO.  Only n_occ <= 128 is started early; results agree with the unhinted path to rounding (the base block is
 * multiplied with another tile shape).  The raw-CI entry points do this on their own.  Not for streams under
 * CUDA-graph capture: the early inverse kernel is launched ahead of its input and waits on the device for a flag
 * (bounded), which relies on the two launches running concurrently. */
int wfo_set_occ_hint(wfo_ctx *ctx, int n_occ);
/* a_inv = a^-1 (n x n), replaces np.linalg.inv in pywfo/main.py:303 (98, 212).
 * n <= 128: one launch of the register-resident Gauss-Jordan inverse with partial pivoting (csrc/lu_base.cuh).
 * n > 128, fast path: Newton-Schulz iteration X <- X (2I - A X) from X0 = A^T / |A^T A|_inf, i.e. nothing but
 * FP64 tensor-core GEMMs; quadratically convergent, about 2 log2(cond) + 6 iterations (orthonormal
 * MOs: the first iterate is already the inverse); the stream is synchronised every 3 iterations to
 * read the residual max|A X - I|.
 * Fallback (residual stagnating above 1e-8 = rounding has taken over, cond >~ 1e7, or no convergence
 * in 51 iterations): Gauss-Jordan elimination with PARTIAL PIVOTING on the device (one cooperative
 * grid, csrc/inverse_gj.cuh), which is to this library what LAPACK getrf/getri is to the reference:
 * it returns an inverse as good as the conditioning allows instead of an error.
 * WFO_ERR_NOCONV only for a singular matrix (zero pivot column), where numpy raises LinAlgError.
 * `work` = 2 n*n doubles of device scratch (NULL: from the context). */
int wfo_inverse_dev(wfo_ctx *ctx, const double *a, int n, double *a_inv, double *work, void *stream);
int wfo_inverse_host(wfo_ctx *ctx, const double *a, int n, double *a_inv);
/* how the last inverse of this context was obtained: Newton-Schulz iterations run, and whether the
 * pivoted elimination produced the result */
int wfo_last_inverse_info(const wfo_ctx *ctx, int *iterations, int *pivoted);
/* generic C = alpha*op(A) op(B) + beta*C (row-major; ta/tb: 0 = N, 1 = T) */
int wfo_gemm_dev(wfo_ctx *ctx, int ta, int tb, int m, int n, int k, double alpha, const double *A,
                 int lda, const double *B, int ldb, double beta, double *C, int ldc, void *stream);

/* ---- stage 2: determinant table for explicit index lists --------------------
 * replaces pywfo/main.py:579-585 (sd_ovlp) and pywfo/dets.py:191-198
 * (precompute): D[k][l] = det(S_MO[rows[k]][:, cols[l]]), rows (kb x n),
 * cols (kk x n), partially pivoted LU per pair.  Test/bench entry: the product
 * path never materialises D (see wfo_state_overlaps_*). */
int wfo_det_table_dev(wfo_ctx *ctx, const double *s_mo, int n_mo, int ld_smo, int n,
                      const int32_t *rows, int kb, const int32_t *cols, int kk, double *D,
                      void *stream);
int wfo_det_table_host(wfo_ctx *ctx, const double *s_mo, int n_mo, int n, const int32_t *rows,
                       int kb, const int32_t *cols, int kk, double *D);
/* excitation tables -> appended-order index lists (pywfo/main.py:565-576) */
int wfo_exc_lists_dev(wfo_ctx *ctx, const int32_t *exc, int k, int n_occ, int32_t *lists,
                      void *stream);

/* ---- stages 2+3 fused: state-overlap matrix --------------------------------
 * replaces the loops pywfo/main.py:587-626 (sigma=+1) and main.py:170-197
 * (sigma=-1) and the contraction pywfo/dets.py:279-283:
 *   out[i][j] = sum_{k in [k_begin,k_end)} sum_l Cb[i][k] Ck[j][l]
 *               ( D(0,0) D(k,l) + sigma D(k,0) D(0,l) )
 * bra_exc (kb x 2), ket_exc (kk x 2) int32 excitation tables, Cb (n_bra x kb),
 * Ck (n_ket x kk) coefficient matrices (thresholding, 1/sqrt2 and the sign
 * conventions of each reference entry point are folded in by the caller, see
 * pywfo_b200/excitations.py).  [k_begin,k_end) is this rank's shard of the bra
 * excitations; partial results of all shards add up to the full matrix (one
 * all-reduce).  Per-pair determinants never leave registers/shared memory.
 * `out` (n_bra x n_ket) is overwritten.  tier: WFO_TIER_*. */
int wfo_state_overlaps_dev(wfo_ctx *ctx, const double *s_mo, int n_mo, int n_occ,
                           const int32_t *bra_exc, int kb, const double *Cb, int n_bra,
                           const int32_t *ket_exc, int kk, const double *Ck, int n_ket,
                           double sigma, int tier, int k_begin, int k_end, double *out,
                           void *stream);
int wfo_state_overlaps_host(wfo_ctx *ctx, const double *s_mo, int n_mo, int n_occ,
                            const int32_t *bra_exc, int kb, const double *Cb, int n_bra,
                            const int32_t *ket_exc, int kk, const double *Ck, int n_ket,
                            double sigma, int tier, int k_begin, int k_end, double *out);
/* The whole path from host buffers in one call (what pywfo_b200.main.overlaps_*
 * uses): S_AO = c_inv c_inv^T, S_MO = bra S_AO ket^T, then the fused stages. */
int wfo_overlaps_host(wfo_ctx *ctx, const double *bra_mos, const double *ket_mos,
                      const double *c_inv, int n_mo, int n_occ, const int32_t *bra_exc, int kb,
                      const double *Cb, int n_bra, const int32_t *ket_exc, int kk,
                      const double *Ck, int n_ket, double sigma, int tier, int k_begin,
                      int k_end, double *out);
/* The whole path from the RAW CI tensors (what pywfo_b200.main.overlaps* uses): the
 * selection of excitations -- pywfo/main.py:539-540 (|c| >= thr per state, semantics 0,
 * sigma=+1: overlaps_cache / overlaps_naive / overlaps_most_naive) or main.py:112,122-123
 * (|c| > thr, semantics 1, sigma=-1 and the sign^occ factor of main.py:154,176: overlaps)
 * -- runs on the device too, so the host only supplies bra_ci (n_bra, n_occ, n_virt),
 * ket_ci (n_ket, n_occ, n_virt) and the MO matrices.  ao_side (WFO_AO_*) says whether c_inv is
 * supplied (WFO_AO_GIVEN) or computed on the device from bra_mos / ket_mos (c_inv may be NULL).  The bra excitations are sharded
 * over `world` ranks ([rank] gets a contiguous, balanced slice); `out` is this rank's
 * partial matrix.  info (may be NULL) receives {K_bra, K_ket, empty_state, tier_used}.
 * Returns WFO_ERR_EMPTY_STATE for semantics 1 when a state keeps no excitation. */
int wfo_overlaps_ci_host(wfo_ctx *ctx, const double *bra_mos, const double *ket_mos,
                         const double *c_inv, int ao_side, int n_mo, int n_occ, const double *bra_ci, int n_bra,
                         const double *ket_ci, int n_ket, double ci_thresh, int semantics, int tier,
                         int rank, int world, double *out, int32_t *info);
/* The same with every buffer on the device (orbitals and CI vectors that already live in HBM).  Runs on
 * `stream` and a side stream of the context, synchronises `stream` once (table sizes) and returns with the
 * rest queued; `out` (n_bra x n_ket, device) is valid after the stream has drained. */
int wfo_overlaps_ci_dev(wfo_ctx *ctx, const double *bra_mos, const double *ket_mos, const double *c_inv,
                        int ao_side, int n_mo, int n_occ, const double *bra_ci, int n_bra,
                        const double *ket_ci, int n_ket, double ci_thresh, int semantics, int tier,
                        int rank, int world, double *out, int32_t *info, void *stream);
/* Trajectory driver: n_cycles geometries (mos: n_cycles x n_mo x n_mo, ci: n_cycles x n_states x n_occ x
 * n_virt, host), the state-overlap matrices of n_pairs (bra cycle, ket cycle) index pairs -> out
 * (n_pairs x n_states x n_states).  Replaces a loop of pywfo.main.overlaps_cache / overlaps calls over
 * the steps of a trajectory (how pysisyphus drives pywfo; pywfo/tests/ref_cytosin holds a 9-cycle
 * example): every cycle is uploaded once and stays resident, the inverse behind S_AO (ao_side =
 * WFO_AO_BRA or WFO_AO_KET) is formed once per cycle, per pair only S_AO, S_MO, the tables and the
 * fused stages run.  info: {max K_bra, max K_ket, empty_state, tier of the last pair}. */
int wfo_trajectory_ci_host(wfo_ctx *ctx, const double *mos, const double *ci, int n_cycles, int n_mo,
                           int n_occ, int n_states, double ci_thresh, int semantics, int tier,
                           int ao_side, const int32_t *pairs, int n_pairs, int rank, int world,
                           double *out, int32_t *info);
/* ---- general determinant lists ----------------------------------
 * replaces pywfo/dets.py:191-198 (precompute) + :279-283 (contraction) for ARBITRARY alpha/beta block lists
 * (the on-disk wfoverlap decks; the reference's CI-tensor entry dets.wfoverlap is wfo_state_overlaps_*):
 *   out[i][j] = sum_{d,e} bc[d][i] alpha[ba_of[d]][ka_of[e]] beta[bb_of[d]][kb_of[e]] kc[e][j]
 * alpha[x][y] = det(S_MO[ba_blocks[x]][:, ka_blocks[y]]) (n_alpha x n_alpha minors), beta likewise.
 * s_mo (n_rows x n_cols, host); *_blocks: unique occupation lists; *_of: block of every determinant;
 * bc (nd_b x ns_b), kc (nd_k x ns_k): coefficients with the string parity of dets.py:42-96 folded in.
 * Bra blocks that share all but their LAST orbital (dets.py:119-156 `super_map`) are served by one pivoted
 * factorisation per (super-block, ket block) and one dot product per member (Laplace expansion along the
 * last row, laplace.py:3-16); the block tables and the contraction stay on the device -- only `out`
 * comes back.  info (4 ints or NULL): super-blocks used, blocks served by them, blocks redone by
 * brute force because the shared factorisation was ill conditioned, 0. */
int wfo_overlap_from_lists_host(wfo_ctx *ctx, const double *s_mo, int n_rows, int n_cols, int n_alpha,
                                const int32_t *ba_blocks, int nba, const int32_t *ka_blocks, int nka, int n_beta,
                                const int32_t *bb_blocks, int nbb, const int32_t *kb_blocks, int nkb, int nd_b,
                                const int32_t *ba_of, const int32_t *bb_of, const double *bc, int ns_b, int nd_k,
                                const int32_t *ka_of, const int32_t *kb_of, const double *kc, int ns_k, double *out,
                                int32_t *info);

/* ---- K5: the one collective of the path, inside the library ------------------
 * (new: the reference is single-process).  One process per GPU of ONE node.
 * wfo_comm_init is collective: every rank calls it with the same `world` and the same job-unique
 * `tag`; the ranks exchange CUDA IPC handles of a per-rank buffer through /dev/shm/wfo_b200_<tag>_*
 * and map each other's buffers (NVLink peer access).  Afterwards
 *   - wfo_overlaps_ci_host(.., rank, world, ..) uploads 1/world of the ket CI tensor per rank and
 *     all-gathers it over NVLink, and returns the FULL n_bra x n_ket matrix on every rank;
 *   - wfo_allreduce_dev sums `count` doubles (<= 131072) over the ranks in place, in rank order
 *     (bitwise identical on all ranks), as two kernels on `stream` (one-shot peer stores + flags).
 * gather_bytes: capacity for the ket CI tensor (0 = 64 MB).  A collective in which a rank does not
 * take part times out after ~20 s of GPU time: wfo_comm_status / the *_host entry points then
 * return WFO_ERR_COMM instead of hanging. */
int wfo_comm_init(wfo_ctx *ctx, int rank, int world, const char *tag, long long gather_bytes);
int wfo_comm_destroy(wfo_ctx *ctx);
int wfo_comm_world(const wfo_ctx *ctx);   /* 0 = no communicator */
int wfo_allreduce_dev(wfo_ctx *ctx, double *buf, long long count, void *stream);
int wfo_comm_status(wfo_ctx *ctx);        /* after a stream sync: WFO_ERR_COMM if a wait timed out */

/* Brute-force tier, 32 < n_occ <= 128: which of the two warp-per-determinant LU kernels runs.
 * AUTO = LEFT = left-looking (csrc/det_lu_ll.cuh: panels in shared memory up to n_occ = 64, in an
 * L2-resident scratch above); RIGHT = right-looking with an L2-resident scratch (csrc/det_lu_wpm.cuh).  Both compute the same partially pivoted LU
 * (pywfo/main.py:583-584); the switch exists so that tests and benchmarks can run either one. */
#define WFO_LU_AUTO 0
#define WFO_LU_LEFT 1
#define WFO_LU_RIGHT 2
int wfo_set_lu_variant(wfo_ctx *ctx, int variant);
/* tier actually used by the last wfo_state_overlaps_* call of this context */
int wfo_last_tier(const wfo_ctx *ctx);
/* device time (ms, CUDA events on the launching stream) of the dominant
 * determinant kernel of the last wfo_state_overlaps_* call made with timing on */
int wfo_set_timing(wfo_ctx *ctx, int on);
double wfo_last_kernel_ms(const wfo_ctx *ctx);
/* timing level 2: CUDA events at the phase boundaries of the last wfo_state_overlaps_* call.
 * Fills ms[i] (up to 23 entries) with the device time of phase i and `names` with the
 * comma-separated phase labels; returns the number of phases (0 if timing < 2). */
int wfo_last_phases(wfo_ctx *ctx, double *ms, char *names, int names_cap);

/* ---- measurement helpers (bench.py roofline denominators) ------------------
 * MEASURED_PEAKS.json has no FP64 entry, so the harness measures it in-run:
 * kind 0 = register-resident DFMA loop, kind 1 = DMMA m8n8k4 loop,
 * kind 2 = DMMA m16n8k8 loop.  Returns TFLOP/s. */
int wfo_fp64_peak(wfo_ctx *ctx, int kind, int iters, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* WFO_B200_H */
