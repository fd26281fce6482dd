/* b200msm — C-ABI of the B200-native BLS12-381 multi-scalar-multiplication engine.
 *
 * This is the drop-in boundary behind ark-blst's arkworks surface.  It replaces, wholesale,
 *      crate::gpu::msm::<G>(bases, exponents)                reference src/gpu.rs:226-241
 *      SingleMultiexpKernel::{create, multiexp}              reference src/gpu.rs:101-119,126-210
 *      the ec-gpu-gen generated *_multiexp kernels           reference build.rs:9-12, gpu.rs:165-183
 * and is what the two trait impls call instead:
 *      <G1Projective as VariableBaseMSM>::msm / msm_bigint   reference src/g1.rs:621-632
 *      <G2Projective as VariableBaseMSM>::msm / msm_bigint   reference src/g2.rs:601-612
 * (INTEGRATION.md shows the Rust `extern "C"` block and the patched bodies.)
 *
 * Data layouts are the reference's own, passed as raw pointers with no conversion
 * (#[repr(transparent)] newtypes over blst types, src/g1.rs:54-56,435-437, src/g2.rs:66-68,415-417):
 *   G1 base   = blst_p1_affine : x, y          each 6×u64 LE Montgomery (R=2^384)   96 bytes
 *   G2 base   = blst_p2_affine : x.c0,x.c1,y.c0,y.c1                                192 bytes
 *   identity base = all-zero bytes (accepted; contributes nothing)
 *   scalar    = 4×u64 LE; Montgomery Fr (R=2^256) when scalars_are_montgomery != 0 — the
 *               &[Scalar] the trait's `msm` receives (src/scalar.rs:23-25) — else a plain
 *               integer < 2^256 (reduced mod r on the device) — the &[BigInt<4>] of `msm_bigint`
 *               and of the reference GPU arm (src/g1.rs:624-627)
 *   G1 result = blst_p1 : X, Y, Z Jacobian Montgomery, 18×u64 (identity ⇔ Z = 0)    144 bytes
 *   G2 result = blst_p2 : 36×u64                                                    288 bytes
 * The result is Σ sᵢ·Pᵢ (the definition pinned by reference src/tests.rs:58-67); it equals the
 * blst-backed value as a group element, i.e. after normalisation to affine.
 *
 * Error convention: every function returns 0 on success, a negative B200MSM_E* code otherwise;
 * nothing throws or aborts across the boundary.  The Rust shim maps non-zero → Err(0), the
 * reference GPU arm's convention (src/g1.rs:628-630).  b200msm_last_error() gives a thread-local
 * message.  There is no CPU fallback: without a usable sm_100 device every call fails.
 *
 * Threading: all entry points may be called concurrently from any host thread (arkworks provers
 * call msm from rayon workers); calls on the same device are serialised internally.
 */
#ifndef B200MSM_H
#define B200MSM_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200MSM_OK 0
#define B200MSM_ENODEV (-1)   /* no CUDA device / not sm_100 */
#define B200MSM_ECUDA (-2)    /* CUDA runtime error (message in b200msm_last_error) */
#define B200MSM_EINVAL (-3)   /* bad argument */
#define B200MSM_ENOMEM (-4)   /* device or host allocation failed */
#define B200MSM_ENCCL (-5)    /* NCCL error on the multi-GPU path */

#define B200MSM_G1 0
#define B200MSM_G2 1

/* Bind the engine to `n_devices` GPUs (0 = all visible) starting at `first_device`.  Optional:
 * the first MSM call initialises with (0, current device only = 1).  Replaces Device::all() +
 * program!(device) + SingleMultiexpKernel::create, which the reference repeats on every call
 * (src/gpu.rs:233-237); here it happens once per process. */
int b200msm_init(int first_device, int n_devices);
void b200msm_shutdown(void);
int b200msm_device_count(void);          /* devices bound by init */
const char *b200msm_last_error(void);
const char *b200msm_version(void);

/* One-shot MSM on HOST buffers: copies bases and scalars to the device(s), runs, returns the
 * Jacobian result in `out`.  Direct replacement of crate::gpu::msm (src/gpu.rs:226-241). */
int b200msm_g1(const uint64_t *bases, const uint64_t *scalars, size_t n,
               int scalars_are_montgomery, uint64_t out[18]);
int b200msm_g2(const uint64_t *bases, const uint64_t *scalars, size_t n,
               int scalars_are_montgomery, uint64_t out[36]);

/* Resident bases: upload a proving key's bases once — sharded evenly by index
 * range across the bound devices — then run many scalar vectors against them. */
typedef struct b200msm_bases b200msm_bases; /* opaque */
int b200msm_bases_upload(int group /*B200MSM_G1|G2*/, const uint64_t *bases, size_t n,
                         b200msm_bases **handle);
int b200msm_bases_free(b200msm_bases *handle);
/* scalars on the host; n must not exceed the uploaded count (prefix is used) */
int b200msm_run(const b200msm_bases *handle, const uint64_t *scalars, size_t n,
                int scalars_are_montgomery, uint64_t *out /*18 or 36 u64*/);

/* Fixed-base window table for resident bases ("upload a proving key once, run many
 * scalar vectors"): converts an uploaded handle in place into table[w][i] = 2^(c·w)·P_i (affine,
 * W = ceil(256/c) windows, W× the memory of the bases).  Every later b200msm_run on the handle
 * then accumulates all windows into ONE bucket set: the bucket reduction covers a single window
 * and the on-device Horner chain (255 dependent doublings) disappears; results are the same group
 * elements.  window_bits = 0 picks the width from the shard size.  One-time cost: ≈ 256 point
 * doublings + one batch normalisation per base.  The reference has no counterpart (it rebuilds
 * even its kernel per call, src/gpu.rs:233-237); provers that reuse a proving key are the use. */
int b200msm_bases_precompute(b200msm_bases *handle, int window_bits);
int b200msm_bases_table_info(const b200msm_bases *handle, int *window_bits, int *windows,
                             size_t *device_bytes);
/* Device-pointer forms of the same: *window_bits = 0 → automatic for n; returns the window count
 * so the caller can allocate d_table = windows × n affine points.  d_bases may alias d_table's
 * first window.  run_table: `stride` = points per window of the table (the n it was built for);
 * n ≤ stride uses the prefix. */
int b200msm_table_plan(int group, size_t n, int *window_bits, int *windows);
int b200msm_table_build_device(int group, const void *d_bases, size_t n, int window_bits,
                               void *d_table, void *stream);
int b200msm_run_table_device(int group, const void *d_table, size_t stride, int window_bits,
                             const void *d_scalars, size_t n, int scalars_are_montgomery,
                             void *d_out, void *stream);

/* Device-pointer form for callers that already hold the inputs in HBM on the CURRENT device
 * (bench.py's kernel-only figure; torch tensors via data_ptr()).  `d_out` is a device buffer of
 * 18/36 u64; the call is asynchronous on `stream` (a cudaStream_t, 0 = default). */
int b200msm_run_device(int group, const void *d_bases, const void *d_scalars, size_t n,
                       int scalars_are_montgomery, void *d_out, void *stream);
/* d_out = Σ of `count` Jacobian partials at d_partials (the final addition after the NCCL
 * gather of per-GPU partial sums); asynchronous on `stream`. */
int b200msm_sum_partials_device(int group, const void *d_partials, int count, void *d_out,
                                void *stream);

/* Batch normalisation of Jacobian points to affine: replaces the host-side
 * CurveGroup::normalize_batch → blstrs batch_normalize (reference src/g1.rs:536-543,
 * src/g2.rs:516-523), the step every caller runs right before msm (src/tests.rs:63).
 * proj: n × 18|36 u64 (blst_p1 / blst_p2), affine_out: n × 12|24 u64; identity → all-zero. */
int b200msm_normalize_batch(int group, const uint64_t *proj, size_t n, uint64_t *affine_out);
int b200msm_normalize_batch_device(int group, const void *d_proj, size_t n, void *d_affine, void *stream);

/* Point (de)serialisation: the ZCash / IETF encodings ↔ affine limbs, replacing the
 * host-side CanonicalSerialize / CanonicalDeserialize of G1Affine and G2Affine (reference
 * src/g1.rs:358-431, src/g2.rs:338-411 → blstrs to_/from_compressed, to_/from_uncompressed, and
 * Valid::check = on curve ∧ torsion free, src/g1.rs:386-396).  Element sizes: G1 48 (compressed) /
 * 96 bytes, G2 96 / 192 bytes.  status_out[i]: 0 ok; 1 malformed encoding — what blstrs'
 * from_*_unchecked rejects and the reference unwrap()s (src/g1.rs:411,419); 2 fails Valid::check
 * (only tested when validate != 0) — the reference's Err(InvalidData).  Entries with status 1 are
 * written as the identity. */
int b200msm_deserialize(int group, const uint8_t *in, size_t n, int compressed, int validate,
                        uint64_t *affine_out, uint8_t *status_out);
int b200msm_serialize(int group, const uint64_t *affine, size_t n, int compressed, uint8_t *out);

/* Cumulative count of this library's own kernel launches (no library kernel runs on the MSM path). */
unsigned long long b200msm_launch_count(void);

/* Tunables (SURVEY §5 "config/flags"): window width c; 0 = automatic from n. */
int b200msm_set_window_bits(int c);
/* GLV split of every scalar into two 128-bit halves k = k1 + k2·λ over (P, φ(P) = (β·x, y)) — half
 * the scalar bits, so half the windows to reduce and half the on-device Horner chain, at the same
 * number of bucket additions. -1 = automatic (time model; the default: on for n ≤ 2^22), 0 = never,
 * 1 = always (two parts). Results are the same group elements either way. */
int b200msm_set_glv(int mode);
/* (mode 2: on G2, FOUR parts instead of two — the base-|z| digits of the scalar over Q, −ψ(Q), ψ²(Q), −ψ³(Q) with ψ the
 * untwist-Frobenius-twist endomorphism, which acts on the subgroup as multiplication by the curve parameter z: a quarter
 * of the windows to reduce and of the dependent doublings of the Horner chain.  G1 has no such endomorphism: mode 2 means
 * two parts there.  The automatic mode chooses among none / two / four by its time model.) */
/* Batched-affine pairing rounds in front of the XYZZ bucket accumulation: the sorted entry array is halved
 * `rounds` times by affine additions that share one field inversion per ≈10^5 pairs (6 products per addition
 * instead of 10 for G1, 17 instead of 28 for G2); what is left of every bucket is accumulated in XYZZ form.
 * -1 = automatic (the default: 2 rounds from ≈40 entries per bucket, 1 from ≈14, else none), 0 = never, 1..3 forced.
 * Results are the same group elements either way. */
int b200msm_set_batch_affine(int rounds);
/* A pass called a second time with the same arguments (sizes, plan, device pointers) is recorded as two CUDA
 * graphs and replayed from then on: 2 launches instead of ≈45 per pass.  1 = on (default), 0 = always issue
 * kernel by kernel.  Never used while b200msm_set_profiling(1). */
int b200msm_set_graphs(int on);
/* Page-lock a caller-owned host buffer (a long-lived scalar or base array) so that the host-buffer
 * entry points copy from it asynchronously at full PCIe rate: ordinary pageable memory is staged
 * by the driver at about a fifth of that and blocks the calling thread (G1 2^20 one-shot: 8.3 ms
 * from pinned, 13.7 ms from pageable memory; with resident bases 8.0 vs 8.2 ms).  Thin wrappers of cudaHostRegister / Unregister. */
int b200msm_host_register(const void *ptr, size_t bytes);
int b200msm_host_unregister(const void *ptr);
/* Lanes: the device-pointer entry points called by THIS host thread use context `lane` (0..7,
 * default 0) of the current device — its own scratch arena, created on first use.  MSMs issued
 * on different lanes and different streams overlap: the latency-bound reduction/combination of
 * one (which runs on a high-priority stream) hides under the accumulation of the next — the
 * batched shape of a Groth16 prover (3×G1 + 1×G2 back to back).  Host-buffer calls use lane 0. */
int b200msm_set_lane(int lane);
/* One-shot host-buffer MSMs (b200msm_g1 / b200msm_g2) of at least `min_points` per device
 * (0 = default 2^18) upload their inputs in up to `slices` pieces of ≈2^17 points (default 8,
 * 1 = off): slice k is
 * grouped and accumulated into the shared buckets while slice k+1 crosses PCIe; reduction and
 * combination run once. */
int b200msm_set_stream_slices(int slices, size_t min_points);
/* Buckets holding more than max(32, factor × mean occupancy, entries/175000) entries leave the
 * one-thread-per-bucket kernel for the block-cooperative path (0 = automatic: 3, or 4 with a
 * fixed-base table). */
int b200msm_set_heavy_factor(int factor);
/* Large inputs are cut into passes automatically (sort arrays < 2^32 entries, scratch within the
 * free HBM) — the chunking the reference left as a TODO (src/gpu.rs:238-239). A non-zero value
 * forces at most that many points per pass (tests); 0 = automatic. */
int b200msm_set_max_chunk(size_t max_points_per_pass);
/* Per-phase device times (ms) of the most recent MSM on this thread's device:
 * [0] digits [1] sort [2] bucket bounds+order [3] accumulate [4] reduce [5] combine [6] total
 * [7] accumulate launches. Filled only when b200msm_set_profiling(1). */
int b200msm_set_profiling(int on);
/* Plan of the most recent pass on this thread's device: [0] window bits c, [1] windows,
 * [2] GLV (0 off; 1 / 2 two parts with a carry window / with the unsigned top digit; 3 / 4 the same with four parts, G2),
 * [3] fixed-base table used. */
int b200msm_last_plan(int out[4]);
/* The plan `auto_plan` would pick for an n-point MSM under the given GLV mode (-1 / 0 / 1), without
 * touching a device: [0] window bits, [1] windows, [2] GLV (0 / 1 / 2), [3] buckets in all. */
int b200msm_plan_query(int group, size_t n, int glv_mode, int out[4]);
int b200msm_last_phase_ms(double out[8]);

/* ---- synthetic data + measurement utilities (bench / tests; not on the reference's path) ---- */
/* Pᵢ = kᵢ·G with kᵢ the counter-based stream of oracle/bls12381.py::synth_dlog, written as
 * affine Montgomery limbs into device memory. */
int b200msm_synth_bases_device(int group, uint64_t seed, size_t n, void *d_out, void *stream);
/* sᵢ of oracle/bls12381.py::synth_scalar (canonical or Montgomery) into device memory */
int b200msm_synth_scalars_device(uint64_t seed, size_t n, int montgomery, void *d_out, void *stream);
/* Measured 32-bit IMAD issue rate of the current device, in IMAD/s (the roofline denominator;
 * MEASURED_PEAKS.json has no integer figure). out[0]=mad.lo rate, out[1]=fused lo/hi
 * (IMAD.WIDE) carry-chain rate ×2, out[2]=SM clock estimate in MHz during the run. */
int b200msm_imad_peak(double out[3]);

/* ---- unit hooks used by the parity tests (element-wise, device-side, host buffers) ---- */
/* op: 0 mul, 1 add, 2 sub, 3 sqr(b ignored), 4 neg(b ignored), 5 inv(b ignored; Fermat power),
 * 6 inv by divsteps (b ignored; csrc/modinv.cuh — the inversion the batched-affine rounds share).
 * a, b, out: n elements of 6 (Fp) or 12 (Fp2) u64 each. */
int b200msm_dbg_field_op(int fp2, int op, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n);
/* op: 0 madd (acc XYZZ + affine), 1 add (XYZZ + XYZZ), 2 dbl (XYZZ); results as XYZZ→Jacobian.
 * acc: n×4 field elems, q: n×(2|4) field elems, out: n×3 field elems (Jacobian). */
int b200msm_dbg_point_op(int group, int op, const uint64_t *acc, const uint64_t *q, uint64_t *out, size_t n);
/* signed window digits of host scalars exactly as the digit kernel emits them:
 * out[w*n + i] = digit (int32) for w < nwin. */
int b200msm_dbg_digits(const uint64_t *scalars, size_t n, int montgomery, int c, int32_t *out, int *nwin);

#ifdef __cplusplus
}
#endif
#endif
