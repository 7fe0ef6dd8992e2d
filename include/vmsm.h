/* vmsm.h -- C ABI of the B200-native MSM / generator-fold engine (libvmsm.so).
 *
 * The reference (toonsegers/verifiable_mpc) is pure Python and has NO FFI layer for this path: the seam is Python
 * name binding.  Each entry point below therefore cites the reference *call site* whose group
 * arithmetic it replaces.  The Python host layer (verifiable_mpc_b200/) binds exactly these symbols with ctypes;
 * INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns int32 status: 0 = VMSM_OK, negative = error; vmsm_last_error() gives the
 *     thread-local message of the last failure.  No exception crosses the boundary.
 *   - host buffers are owned by the caller; device objects are opaque uint64 handles owned by a context and
 *     freed explicitly (or with the context).
 *   - a context is bound to ONE CUDA device and ONE stream; calls on one context must be serialised by the
 *     caller; different contexts may be driven from different threads/processes (ctypes releases the GIL).
 *   - wire formats (little-endian):   scalar = 32 B, already reduced below the group order;
 *       Ed25519 point = 64 B canonical affine x || y (identity = (0, 1));
 *       extended point (multi-GPU partials) = 128 B X || Y || Z || T, any representative.
 *   - there is no CPU fallback: without a CUDA device vmsm_ctx_create fails with VMSM_ERR_CUDA.
 */
#ifndef VMSM_H
#define VMSM_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VMSM_OK 0
#define VMSM_ERR_INVALID (-1)  /* bad argument / handle / range */
#define VMSM_ERR_CUDA (-2)     /* CUDA runtime error, no device */
#define VMSM_ERR_POINT (-3)    /* uploaded point not canonical or not on the curve */
#define VMSM_ERR_NOMEM (-4)
#define VMSM_ERR_UNSUPPORTED (-5)
#define VMSM_ERR_TIMEOUT (-6)  /* a multi-GPU partial did not arrive */

/* curves (mpyc.fingroups.EllipticCurve names used by the reference: demos/demo_zkp_ac20.py:46,
 * demos/demo_zkp_pynocchio.py:27-29) */
#define VMSM_CURVE_ED25519 0
#define VMSM_CURVE_BN256_G1 1
#define VMSM_CURVE_BN256_G2 2

/* vmsm_ctx_set_option keys */
#define VMSM_OPT_WINDOW_BITS 1  /* 0 = auto (default), else force the Pippenger window c in [2, 18] */
#define VMSM_OPT_PHASE_TIMING 2 /* 1 = bracket every MSM phase with CUDA events (see vmsm_phase_times) */
#define VMSM_OPT_SORT_BUCKETS 3 /* 1 = process buckets in order of decreasing population (default 1) */
#define VMSM_OPT_CHECK_POINTS 4 /* 1 = validate uploaded points (default 1) */
#define VMSM_OPT_REDUCE_RADIX 5 /* log2 of the bucket-tree radix (default 3) */
#define VMSM_OPT_ASYNC_TAIL 7 /* 1 (default) = run the latency-bound MSM tail on a side stream under the next MSM's head */
#define VMSM_OPT_CAP_FACTOR 8 /* per-thread bucket cap = max(64, factor * average bucket population) (default 8) */
#define VMSM_OPT_SHARD_SEQ 9 /* non-zero: the NEXT vmsm_msm_dev / vmsm_msm_async call is one shard of a multi-GPU MSM
                               with this sequence number (see the mailbox functions); clears itself */
#define VMSM_OPT_ASYNC_SORT 10 /* 1 (default) = run the counting sort of the next MSM on a side stream under the
                                 accumulate kernel of the previous one (CSR lists double-buffered) */
#define VMSM_OPT_SORT_BLOCKS 11 /* thread blocks of the digit-histogram and scatter kernels when they run on the sort
                                  stream (grid-stride); small = a thin slice of every SM for longer, so the sort
                                  shares the SMs with the accumulate kernel instead of displacing it (used only
                                  while the previous MSM is still accumulating); -1 (default) = four blocks per SM below
                                  2^18 terms, two up to 2^20, one from 2^21; 0 = always one thread per scalar */
#define VMSM_OPT_FOLD_QUAD_MAX 12 /* generator folds of at most this many outputs use the 4-lanes-per-element kernel
                                    (latency-bound rounds); 0 = never */
#define VMSM_OPT_BN_QUAD_ACC 13 /* BN256 accumulate kernel with four lanes per bucket: 0 never, 1 (default) for the sizes
                                  where it was measured faster, 2 always */
#define VMSM_OPT_PRE_SETS 14 /* MSMs over precomputed bases (vmsm_points_precompute): bucket sets shared by the windows;
                               0 (default) = chosen from the term count */
#define VMSM_OPT_PRE_MIN_TERMS 15 /* MSM calls with fewer terms ignore a precomputed table (default 256) */
#define VMSM_OPT_SEG_LEN 16 /* entries summed by one thread of the balanced (segmented) accumulate kernel; 0 (default) =
                              sized so that the launch is a whole number of waves of resident threads, at most 32 */
#define VMSM_OPT_SEG_MODE 17 /* accumulate kernel: 0 = one thread per bucket, 2 = equal segments of the sorted entries,
                               1 (default) = by geometry (segments when the windows of a table-based MSM share one
                               bucket set, below 2^19 terms) */
#define VMSM_OPT_HOST_NORMALIZE 18 /* 1 (default): Ed25519 results leave the device in extended coordinates and the
                                     fetching call inverts Z on the CPU (~15 us) instead of a lone GPU thread (0.19 ms at
                                     the end of every MSM); 0 = normalise on the device.  Same canonical output. */
#define VMSM_OPT_DUAL_HEAD 19 /* 1 (default): consecutive Ed25519 MSMs run their accumulate kernels on two alternating
                                streams (one launch below ~2^18 terms is one or two waves that end together: the next
                                MSM's kernel fills the ramp-down); 0 = all accumulate kernels on the context's stream */
#define VMSM_OPT_QUAD_THRESHOLD 6 /* bucket-tree levels with <= this many nodes use 4 lanes per node (0 = never) */
#define VMSM_OPT_BN_PRE_SETS 23 /* BN256 MSMs over key tables: bucket sets shared by the windows (0 = default: 2 with the
                                  balanced accumulate kernel, one per window with VMSM_OPT_SEG_MODE 0) */
#define VMSM_OPT_BN_SEG_LEN 24 /* BN256: entries per thread of the balanced accumulate kernel (0 = whole waves, >= 8) */
#define VMSM_OPT_BN_QUAD_FIX 25 /* BN256: 1 (default) = four lanes per bucket in the fix-up of buckets that straddle segments */
#define VMSM_OPT_BLOCK_SORT 20 /* 1: counting sort of the digits with per-block shared-memory counters (digits recoded
                                 once into 16-bit codes, no global atomics) for Ed25519 MSMs of at least
                                 VMSM_OPT_BLOCK_SORT_MIN terms and windows c <= 16; 0 (default) = two passes with global
                                 atomics, which measured faster on B200 (profiles/r02/block_sort_experiment.md) */
#define VMSM_OPT_BLOCK_SORT_MIN 21 /* smallest MSM that takes the block-privatised sort (default 2^15 terms) */
#define VMSM_OPT_ACC_CARVEOUT 22 /* experiment: preferred shared-memory carveout (percent, -1 = driver default, which is
                                   what the library uses) of the accumulate kernels; the blocks of VMSM_OPT_BLOCK_SORT
                                   need 128 KB of shared memory on the SMs those kernels occupy */

/* phases reported by vmsm_phase_times */
#define VMSM_PHASE_DIGITS 0
#define VMSM_PHASE_SCAN 1
#define VMSM_PHASE_SCATTER 2
#define VMSM_PHASE_ORDER 3
#define VMSM_PHASE_HANDOFF 4 /* sorted lists waiting for the main stream (previous MSM's accumulate kernel) */
#define VMSM_PHASE_ACCUMULATE 5
#define VMSM_PHASE_REDUCE 6
#define VMSM_PHASE_FINAL 7
#define VMSM_PHASE_COUNT 8

int32_t vmsm_version(void);
const char *vmsm_last_error(void);
int32_t vmsm_device_count(int32_t *count);

/* ---- contexts ------------------------------------------------------------------------------------------- */
int32_t vmsm_ctx_create(int32_t device, uint64_t *ctx);
int32_t vmsm_ctx_destroy(uint64_t ctx);
int32_t vmsm_ctx_set_option(uint64_t ctx, int32_t key, int64_t value);
int32_t vmsm_sync(uint64_t ctx);
/* CUDA-event timer on the context's stream (the stream every kernel of this context is launched on). */
int32_t vmsm_timer_start(uint64_t ctx);
int32_t vmsm_timer_stop(uint64_t ctx, float *ms); /* synchronises */
/* cumulative per-phase device time (ms) and number of MSM calls since the last query; resets the sums. */
int32_t vmsm_phase_times(uint64_t ctx, double *ms_out /* [VMSM_PHASE_COUNT] */, uint64_t *calls);
/* number of kernels this context has launched so far (bench.py's gpu_launches). */
int32_t vmsm_launch_count(uint64_t ctx, uint64_t *launches);

/* ---- device-resident point vectors (the generator lists g, g_hat of pivot.py:139 / compressed_pivot.py:29) */
int32_t vmsm_points_upload(uint64_t ctx, int32_t curve, const uint8_t *affine, uint64_t n, uint64_t *pts);
/* g_i = r_i * B, r_i given (32 B LE each, < order) or, when scalars == NULL, r_i = synth(seed, i) -- the batch
 * form of create_generators (verifiable_mpc/ac20/circuit_sat_r1cs.py:59-74: g.append(h ** r)). */
int32_t vmsm_points_fixed_base(uint64_t ctx, int32_t curve, const uint8_t *scalars, uint64_t seed, uint64_t n,
                               uint64_t *pts);
/* Table of 2^(c*w) * P_i (w < ceil(254 / c), niels form, c = window_bits or 0 = by size) for a vector of FIXED
 * generators: the g of create_generators (circuit_sat_r1cs.py:47-93) is made once and then used by every
 * pivot.vector_commitment of every proof (circuit_sat_cb.py:103, compressed_pivot.py:110).  Every later MSM call on the
 * vector (any sub-range; extra terms need a table of the same window on their own vector) then runs without a single
 * doubling: all windows share a few bucket sets, no Horner chain.  Costs W * 96 B per point of HBM and one kernel of
 * ~250 doublings + W inversions per point.  vmsm_fold drops the table (the folded generators are new points).
 * Ed25519 only. */
int32_t vmsm_points_precompute(uint64_t ctx, uint64_t pts, uint32_t window_bits);
int32_t vmsm_points_download(uint64_t ctx, uint64_t pts, uint64_t off, uint64_t n, uint8_t *affine_out);
/* new vector = a[a_off .. a_off+a_n) || b[b_off .. b_off+b_n)  (device-side copy; b_n may be 0: a clone).  The
 * g_hat = g + [h] of compressed_pivot.py:137 and the private copy a prover folds in place. */
int32_t vmsm_points_concat(uint64_t ctx, uint64_t a, uint64_t a_off, uint64_t a_n, uint64_t b, uint64_t b_off,
                           uint64_t b_n, uint64_t *out);
/* Decimal transcript text of a point range, formatted on the device: "[x0, y0, 1], [x1, y1, 1], ..." -- exactly what
 * repr() of the list of normalised group elements has between its outer brackets, which is what enters the
 * Fiat-Shamir pre-image of every folding round (compressed_pivot.py:51-59 via pivot.py:131-136).  `cap` >= 176 * n
 * always suffices; *len receives the number of bytes written (no terminator).  Ed25519 only. */
int32_t vmsm_points_text(uint64_t ctx, uint64_t pts, uint64_t off, uint64_t n, uint8_t *out, uint64_t cap,
                         uint64_t *len);
/* same, without the copy: *text points into a page-locked buffer owned by the context, valid until the next
 * vmsm_points_text* call on it. */
int32_t vmsm_points_text_ptr(uint64_t ctx, uint64_t pts, uint64_t off, uint64_t n, const uint8_t **text,
                             uint64_t *len);
/* canonical wire bytes of a range without a host copy: *affine points into the same context-owned page-locked buffer
 * as vmsm_points_text_ptr (valid until the next *_ptr call).  Feeds the opt-in binary Fiat-Shamir transcript, which
 * hashes 64 bytes per generator instead of ~157 characters of decimal text. */
int32_t vmsm_points_download_ptr(uint64_t ctx, uint64_t pts, uint64_t off, uint64_t n, const uint8_t **affine);
int32_t vmsm_points_count(uint64_t ctx, uint64_t pts, uint64_t *n);
int32_t vmsm_points_free(uint64_t ctx, uint64_t pts);

/* ---- device-resident scalar vectors ----------------------------------------------------------------------- */
int32_t vmsm_scalars_upload(uint64_t ctx, const uint8_t *le32, uint64_t n, uint64_t *sc);
int32_t vmsm_scalars_synth(uint64_t ctx, int32_t curve, uint64_t seed, uint64_t n, uint64_t *sc);
int32_t vmsm_scalars_download(uint64_t ctx, uint64_t sc, uint64_t off, uint64_t n, uint8_t *le32_out);
int32_t vmsm_scalars_free(uint64_t ctx, uint64_t sc);
int32_t vmsm_scalars_download_ptr(uint64_t ctx, uint64_t sc, uint64_t off, uint64_t n, const uint8_t **le32); /* as above */

/* Scalar vectors modulo the Ed25519 group order l: the halving of the witness and of the linear form that accompanies
 * every generator fold, without leaving HBM.  Entries must be reduced (< l).
 * vmsm_scalars_fold, in place on sc[0 .. 2*half), result in sc[0 .. half):
 *   VMSM_FOLD_WITNESS  z'_j = z_j + c * z_{half+j}      compressed_pivot.py:76  (z_prime)
 *   VMSM_FOLD_FORM     L'_j = c * L_j + L_{half+j}      compressed_pivot.py:68-73 (L_prime), :189-193 (verifier)
 * vmsm_scalars_dot: sum_i a[aoff+i] * b[boff+i] mod l -- L_tilde([0]*half + z_L), L_tilde(z_R + [0]*half), :41-42.
 * vmsm_scalars_text_ptr: "v0, v1, ..." as MPyC prints field elements (is_signed: representatives in (-l/2, l/2]),
 *   the coefficient text of L_tilde in the Fiat-Shamir pre-image, :51-54; *text points into a second page-locked
 *   buffer of the context (valid until the next vmsm_scalars_text_ptr call), so the text of the generators and of the
 *   form of one round can both be fetched while that round's commitments A_i, B_i are still being computed.
 * vmsm_scalars_axpy, the same element-wise operations on two vectors (ranges must not overlap):
 *   VMSM_AXPY_ADD_SCALED  dst_j = dst_j + c * src_j      z = r + c0 * x, phi = rho + c0 * gamma   :134-135
 *   VMSM_AXPY_SCALE_ADD   dst_j = c * dst_j + src_j
 *   VMSM_AXPY_SCALE       dst_j = c * dst_j              L_tilde = (L.coeffs + [0]) * c1           :141, :236 */
#define VMSM_FOLD_WITNESS 0
#define VMSM_FOLD_FORM 1
#define VMSM_AXPY_ADD_SCALED 0
#define VMSM_AXPY_SCALE_ADD 1
#define VMSM_AXPY_SCALE 2
int32_t vmsm_scalars_fold(uint64_t ctx, uint64_t sc, uint64_t half, const uint8_t *c_le32, int32_t mode);
int32_t vmsm_scalars_axpy(uint64_t ctx, uint64_t dst, uint64_t doff, uint64_t src, uint64_t soff, uint64_t n,
                          const uint8_t *c_le32, int32_t mode);
int32_t vmsm_scalars_dot(uint64_t ctx, uint64_t a, uint64_t aoff, uint64_t b, uint64_t boff, uint64_t n,
                         uint8_t *out_le32);
int32_t vmsm_scalars_text_ptr(uint64_t ctx, uint64_t sc, uint64_t off, uint64_t n, int32_t is_signed,
                              const uint8_t **text, uint64_t *len);

/* ---- multi-scalar multiplication -------------------------------------------------------------------------
 * out = sum_{i<n} s_i * P[off + i].  Replaces pivot.vector_commitment / list_mul
 * (verifiable_mpc/ac20/pivot.py:139-145, :26-28; call sites compressed_pivot.py:41-42,110,193,
 * circuit_sat_cb.py:103) and the per-key sums of pynocchio.compute_proof (trinocchio/pynocchio.py:229-246). */
/* end to end: host scalars in, host canonical-affine point out (H2D + kernels + D2H, synchronous) */
int32_t vmsm_msm(uint64_t ctx, uint64_t pts, uint64_t off, uint64_t n, const uint8_t *scalars_le32,
                 uint8_t *out_affine);
/* Pedersen form  out = sum_{i<n} s_i * P[off+i]  +  sum_{j<n_extra} s_{n+j} * E[extra_off+j]  in ONE pass: the
 * `(h ** gamma) * prod` of pivot.vector_commitment (pivot.py:143-144), also A_i/B_i of compressed_pivot.py:41-42 whose
 * blinding base k is not part of the generator vector.  `scalars_le32` holds n + n_extra scalars. */
int32_t vmsm_msm_ext(uint64_t ctx, uint64_t pts, uint64_t off, uint64_t n, uint64_t extra_pts, uint64_t extra_off,
                     uint64_t n_extra, const uint8_t *scalars_le32, uint8_t *out_affine);
/* asynchronous end to end: the H2D copy of the scalars (page-locked memory recommended, see vmsm_host_alloc) runs on
 * a copy stream and overlaps the previous MSM; fetch with vmsm_result_affine(slot), which waits for that result only.
 * `scalars_le32` must stay valid until the result has been fetched. */
int32_t vmsm_msm_async(uint64_t ctx, uint64_t pts, uint64_t off, uint64_t n, const uint8_t *scalars_le32,
                       uint32_t slot);
/* device resident, asynchronous on the context's streams: result goes to result slot `slot` (0..62).  A slot must not
 * be targeted again before its result has been fetched (vmsm_result_affine) or the context synchronised: consecutive
 * MSMs finish on different side streams. */
int32_t vmsm_msm_dev(uint64_t ctx, uint64_t pts, uint64_t poff, uint64_t n, uint64_t sc, uint64_t soff,
                     uint32_t slot);
/* As vmsm_msm_dev with n_extra more terms whose scalars come from the host (the k^{L(z)} factor of the cross terms):
 * A_i = g_R^{z_L} * k^{L_R(z_L)} with z resident in HBM -- verifiable_mpc/ac20/compressed_pivot.py:41-42; on the BN256
 * groups the seven mid-wire sums of pynocchio.py:248-262 share ONE device-resident witness and append their own
 * zero-knowledge delta terms this way.  Both vectors on the same curve; at most 64 extra terms. */
int32_t vmsm_msm_dev_ext(uint64_t ctx, uint64_t pts, uint64_t poff, uint64_t n, uint64_t sc, uint64_t soff,
                         uint64_t extra_pts, uint64_t extra_off, uint64_t n_extra, const uint8_t *extra_scalars_le32,
                         uint32_t slot);
/* The same commitment with the scalar of its ONE extra term computed on the device as the inner product
 * <dot_a[dot_aoff ..], dot_b[dot_boff ..]> of dot_n residues mod l: the cross term L_R(z_L) of
 * compressed_pivot.py:41-42 never visits the host (two host round trips less per folding round).  Ed25519. */
int32_t vmsm_msm_dev_ext_dot(uint64_t ctx, uint64_t pts, uint64_t poff, uint64_t n, uint64_t sc, uint64_t soff,
                             uint64_t extra_pts, uint64_t extra_off, uint64_t dot_a, uint64_t dot_aoff, uint64_t dot_b,
                             uint64_t dot_boff, uint64_t dot_n, uint32_t slot);
int32_t vmsm_result_affine(uint64_t ctx, uint32_t slot, uint8_t *out_affine);      /* synchronises */
int32_t vmsm_result_extended(uint64_t ctx, uint32_t slot, uint8_t *out_extended);  /* synchronises */

/* ---- multi-GPU: index-range split of ONE MSM over the GPUs of a box ---------------------------
 * Every GPU computes the partial sum of its slice of (bases, scalars) and its final kernel pushes the 128-byte
 * partial into a mailbox in the owner GPU's HBM (peer store over NVLink) followed by the sequence number `seq`;
 * on the owner a gather kernel waits for all `world` sequence numbers, adds the partials and normalises.  No host
 * hop and no NCCL.  Owner = rank 0: vmsm_mailbox_create (+ optional CUDA-IPC handle for other processes); other
 * ranks: vmsm_mailbox_open_ipc (another process) or vmsm_mailbox_open_local (another context of the same process).
 * vmsm_result_affine(slot) returns the combined result on the owner (VMSM_ERR_TIMEOUT if a partial never arrived)
 * and the rank's own partial elsewhere.  `seq` must be non-zero and strictly increase from call to call. */
int32_t vmsm_mailbox_create(uint64_t ctx, uint32_t world, uint8_t *ipc_handle_out /* 64 B or NULL */);
int32_t vmsm_mailbox_open_ipc(uint64_t ctx, const uint8_t *ipc_handle /* 64 B */, uint32_t rank, uint32_t world);
int32_t vmsm_mailbox_open_local(uint64_t ctx, uint64_t owner_ctx, uint32_t rank);
int32_t vmsm_msm_dev_shard(uint64_t ctx, uint64_t pts, uint64_t poff, uint64_t n, uint64_t sc, uint64_t soff,
                           uint32_t slot, uint32_t seq);

/* ---- generator fold ----------------------------------------------------------------------------------------
 * In place: P[j] = c * P[j] + P[half + j] for j < half, then the vector length becomes `half`.
 * Replaces the comprehension g_prime = [(g_hat_l[i] ** c) * g_hat_r[i] ...]
 * (verifiable_mpc/ac20/compressed_pivot.py:64 prover, :178 verifier; mpc_ac20.py:176). */
int32_t vmsm_fold(uint64_t ctx, uint64_t pts, uint64_t half, const uint8_t *c_le32);

/* ---- small group helper: out = sum_i s_i * P_i for a handful of host points (Q' = A * Q^c * B^(c^2),
 * compressed_pivot.py:66; Q = A * P^c0 * k^(...), :140).  n <= 64. */
int32_t vmsm_lincomb(uint64_t ctx, int32_t curve, const uint8_t *affine, const uint8_t *scalars_le32, uint64_t n,
                     uint8_t *out_affine);

/* Asynchronous form (Ed25519): the result is fetched with vmsm_result_affine(slot), so Q' of round i is computed
 * underneath the folds and the A/B commitments of round i+1 and only awaited by the next challenge hash.  An invalid
 * input point surfaces as VMSM_ERR_POINT from that fetch. */
int32_t vmsm_lincomb_async(uint64_t ctx, int32_t curve, const uint8_t *affine, const uint8_t *scalars_le32, uint64_t n,
                           uint32_t slot);

/* ---- pinned host memory for the end-to-end path (H2D of scalars straight from page-locked memory) -------- */
int32_t vmsm_host_alloc(uint64_t bytes, void **ptr);
int32_t vmsm_host_free(void *ptr);

/* ---- device self-test of the field arithmetic: out[i] = a[i] (op) b[i] on the GPU, n elements of 32 B.
 * op: 0 add, 1 sub, 2 mul, 3 inv(a), 4 canon(a), 5 sqr(a), 6 (a+b)(a+2b), 7 (a-b)*2(a+b), 8 (a+b)^2, 9 (2a+b)-3b,
 * 10 -a; results are canonical.  Used by tests/ to pin the device field arithmetic (PTX carry chains). */
int32_t vmsm_selftest_fe(uint64_t ctx, int32_t op, const uint8_t *a, const uint8_t *b, uint64_t n, uint8_t *out);

/* ---- integer-pipe peak: independent carry-chained IMAD.WIDE.U32 multiply-accumulate chains on every SM,
 * timed with CUDA events on the context's stream.  Returns tera limb-products (32x32+64->64) per second: the
 * roofline denominator bench.py reports next to the achieved figure ("measure on the box"). */
int32_t vmsm_microbench_imad(uint64_t ctx, double *tera_lp_per_s);

#ifdef __cplusplus
}
#endif
#endif /* VMSM_H */
