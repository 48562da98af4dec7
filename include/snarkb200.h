/* snarkb200.h — C ABI of libsnarkb200.so, the B200 (sm_100a) backend for snarkjs' bulk curve operations.
 *
 * Every entry point replaces one async method of the ffjavascript `curve` object that snarkjs' provers call
 * (SURVEY.md §8b).  Citations are into /root/reference/build/snarkjs.js (first bundled copy of
 * ffjavascript@0.3.1) unless a src/ path is given.  INTEGRATION.md shows the N-API shim that binds them.
 *
 * Conventions
 *   - all pointers are HOST memory unless the name ends in _dev; buffers are little-endian;
 *   - field elements are 32 bytes (Fr, BN254 Fq) or 48 bytes (BLS12-381 Fq);
 *     "Montgomery" = x*2^(8*n8) mod p, fully reduced (reference 2873-2874, 3263-3272);
 *   - G1 affine = x||y (2*n8q), G2 affine = x.c0||x.c1||y.c0||y.c1 (4*n8q), infinity = all-zero bytes;
 *   - MSM output = Jacobian X||Y||Z Montgomery (3*n8q / 6*n8q), normalised to Z = 1 (infinity = (0,1,0));
 *     the reference returns an arbitrary projective representative, only its toAffine() is defined (§3.3);
 *   - return 0 on success, negative on error; sb_last_error(ctx) gives the message — the JS shim throws
 *     `new Error(msg)` so that error strings match the reference's;
 *   - the callee never retains caller memory (reference: inputs are sliced/copied, 14645-14646, 14739);
 *   - a context is bound to one CUDA device.  Every call locks its context for its duration, so overlapping calls on
 *     one context from several threads are safe and run one after the other (the reference awaits several bulk calls at
 *     once, 14653 / 14929-14932; an N-API shim runs them as AsyncWorkers on libuv threads).  For concurrency use one
 *     context per thread / per GPU.  sb_last_error returns the message of the calling thread's last failed call.
 */
#ifndef SNARKB200_H
#define SNARKB200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct sb_ctx sb_ctx;
enum { SB_BN254 = 0, SB_BLS12_381 = 1 };
enum { SB_G1 = 1, SB_G2 = 2 };
enum {
    SB_OK = 0,
    SB_ERR_ARG = -1,        /* bad argument (message mirrors the reference's Error text) */
    SB_ERR_CUDA = -2,       /* CUDA runtime error */
    SB_ERR_NOMEM = -3,
    SB_ERR_FORMAT = -4,     /* malformed zkey / wtns */
    SB_ERR_NODEVICE = -5    /* no CUDA device: the library never falls back to the CPU */
};

/* buildBn128 / buildBls12381 + buildEngine (16413-16523, 15433-15490): one context per curve and device. */
int  sb_create(int curve, int device_id, sb_ctx** out);
void sb_destroy(sb_ctx* ctx);                       /* curve.terminate() 14259-14264 */
const char* sb_last_error(sb_ctx* ctx);
const char* sb_version(void);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
uint64_t sb_launch_count(sb_ctx* ctx);

/* G1.multiExpAffine / G2.multiExpAffine (14666-14668 -> _multiExp 14605-14661).
 * n = number of points; scalar_bytes = bytes per scalar (the reference infers it as byteLength/n and throws
 * "Scalar size does not match" when not integral, 14562-14565 — the shim performs that check).
 * n == 0 -> zero point (14561, 14627). */
int sb_msm_g1_affine(sb_ctx* ctx, const uint8_t* bases, const uint8_t* scalars, uint32_t scalar_bytes, uint64_t n, uint8_t* out);
int sb_msm_g2_affine(sb_ctx* ctx, const uint8_t* bases, const uint8_t* scalars, uint32_t scalar_bytes, uint64_t n, uint8_t* out);

/* Read-only base sets (zkey sections 5-9, PLONK/fflonk PTau) registered once and kept in HBM across proofs. */
int sb_bases_register(sb_ctx* ctx, int group, const uint8_t* bases, uint64_t n, uint64_t* handle);
int sb_bases_release(sb_ctx* ctx, uint64_t handle);
/* MSM over registered bases [first, first+n) with host scalars. */
int sb_msm_registered(sb_ctx* ctx, uint64_t handle, uint64_t first, const uint8_t* scalars, uint32_t scalar_bytes, uint64_t n, uint8_t* out);
/* Same, additionally returning the un-normalised extended-Jacobian partial (X,Y,ZZ,ZZZ; 4 or 8 coordinates) that
 * sb_msm_sum_partials combines — the exchange unit of the multi-GPU MSM (one rank per GPU). */
int sb_msm_registered_partial(sb_ctx* ctx, uint64_t handle, uint64_t first, const uint8_t* scalars, uint32_t scalar_bytes, uint64_t n, uint8_t* partial_out);
int sb_msm_sum_partials(sb_ctx* ctx, int group, const uint8_t* partials, int count, uint8_t* out);
uint32_t sb_msm_partial_bytes(sb_ctx* ctx, int group);

/* Fr.fft / Fr.ifft (15101-15107 -> _fft 14675-14918).  n must be a power of two ("fft must be multiple of 2",
 * 14745-14747) with log2(n) <= Fr.s (28 / 32).  Natural order in and out; inverse != 0 scales by 1/n. */
int sb_ntt_fr(sb_ctx* ctx, const uint8_t* in, uint64_t n, int inverse, uint8_t* out);
/* Fr.batchApplyKey (14273-14384 / frm_batchApplyKey 9458): out[i] = in[i] * first * inc^i. */
int sb_fr_batch_apply_key(sb_ctx* ctx, const uint8_t* in, uint64_t n, const uint8_t first[32], const uint8_t inc[32], uint8_t* out);
/* Fr.batchToMontgomery / Fr.batchFromMontgomery (12895-12896 -> 12780-12830). */
int sb_fr_batch_to_montgomery(sb_ctx* ctx, const uint8_t* in, uint64_t n, uint8_t* out);
int sb_fr_batch_from_montgomery(sb_ctx* ctx, const uint8_t* in, uint64_t n, uint8_t* out);
/* tm.queueAction([qap_joinABC, frm_batchFromMontgomery]) as used by joinABC, src/groth16_prove.js:320-374:
 * out[i] = fromMontgomery(a[i]*b[i] - c[i]). */
int sb_qap_join_abc(sb_ctx* ctx, const uint8_t* a, const uint8_t* b, const uint8_t* c, uint64_t n, uint8_t* out_plain);
/* Fr constants the JS side reads from the curve object: what = -1 -> Fr.shift (nqr^2), -2 -> Fr.nqr,
 * 0..s -> Fr.w[what] (12866-12889).  Returns s. */
int sb_fr_root(sb_ctx* ctx, int what, uint8_t out[32]);

/* Fused Groth16 prover (src/groth16_prove.js:28-144) with every intermediate resident in HBM.
 * sb_groth16_load parses a Groth16 .zkey image (src/zkey_utils.js:229-259 + sections 4-9), uploads the five base
 * sets and a CSR form of the coefficient section once.  sb_groth16_prove takes the witness section payload
 * (n_witness * 32 bytes, plain LE, src/wtns_utils.js:25-37) and (r, s) as 32-byte Montgomery Fr elements (the
 * reference draws them with Fr.random(), :103-104), and writes the affine proof pi_a (2*n8q) || pi_b (4*n8q) ||
 * pi_c (2*n8q), Montgomery.  public signals are witness[1..nPublic]. */
int sb_groth16_load(sb_ctx* ctx, const uint8_t* zkey, uint64_t zkey_len, uint64_t* handle);
int sb_groth16_load_file(sb_ctx* ctx, const char* zkey_path, uint64_t* handle);
/* multi-GPU variant: this rank keeps (and builds window tables for) only its point-range shard of the five base sets;
 * the handle then serves sb_groth16_prove_shard(shard, n_shards) only. */
int sb_groth16_load_sharded(sb_ctx* ctx, const uint8_t* zkey, uint64_t zkey_len, int shard, int n_shards, uint64_t* handle);
int sb_groth16_info(sb_ctx* ctx, uint64_t handle, uint32_t* n_vars, uint32_t* n_public, uint32_t* domain_size);
int sb_groth16_prove(sb_ctx* ctx, uint64_t handle, const uint8_t* witness, uint64_t n_witness,
                     const uint8_t r[32], const uint8_t s[32], uint8_t* proof_affine_out);
/* full file-to-proof convenience: reads the .wtns container, checks curve and length like the reference (:44-50). */
int sb_groth16_prove_wtns(sb_ctx* ctx, uint64_t handle, const uint8_t* wtns, uint64_t wtns_len,
                          const uint8_t r[32], const uint8_t s[32], uint8_t* proof_affine_out);
/* same proof with the witness uploaded by the previous sb_groth16_prove on this handle still resident in HBM
 * (no host->device copy): the device-resident timing bench.py reports as `value`. */
int sb_groth16_prove_resident(sb_ctx* ctx, uint64_t handle, const uint8_t r[32], const uint8_t s[32], uint8_t* proof_affine_out);
int sb_groth16_release(sb_ctx* ctx, uint64_t handle);
/* ---- PLONK (src/plonk_prove.js:47-889), next-tier path per SURVEY §8f rank 3 ----------------------------------
 * sb_plonk_load: a PLONK zkey (protocol id 2, sections 2-14: src/zkey_utils.js:261-299, src/plonk_constants.js) goes to
 *   HBM once: selector / sigma / Lagrange polynomials in coefficient and 4n-evaluation form, wire maps, additions, and the
 *   PTau bases (section 14) with their MSM window tables.  Error "zkey file is not plonk" as plonk_prove.js:58-60.
 * sb_plonk_prove: witness = section 2 of the .wtns file (plain LE, nVars - nAdditions elements, plonk_prove.js:66-68);
 *   blinders = b_1..b_11 as 11 Montgomery field elements (the reference draws them with Fr.random(), :246-249);
 *   proof_out = A B C Z T1 T2 T3 Wxi Wxiw (affine Montgomery, 2*n8q bytes each) then eval_a eval_b eval_c eval_s1
 *   eval_s2 eval_zw (Montgomery, 32 bytes each): sb_plonk_proof_bytes().  Errors carry the reference's texts:
 *   "Invalid witness length. Circuit: N, witness: M, A", "Copy constraints does not match" (:436-438),
 *   "Polynomial is not divisible" (polynomial.js:608, 653), "T Polynomial is not well calculated" (:648-650). */
int sb_plonk_load(sb_ctx* ctx, const uint8_t* zkey, uint64_t zkey_len, uint64_t* handle);
/* same from a file: the .zkey is mapped read-only and streamed to HBM section by section */
int sb_plonk_load_file(sb_ctx* ctx, const char* zkey_path, uint64_t* handle);
int sb_plonk_info(sb_ctx* ctx, uint64_t handle, uint32_t* n_vars, uint32_t* n_public, uint32_t* domain_size, uint32_t* n_additions);
uint32_t sb_plonk_proof_bytes(sb_ctx* ctx);
/* the same proof from the witness the previous sb_plonk_prove on this key left in HBM (device-resident timing, and
 * re-proving with fresh blinders without a second upload) */
int sb_plonk_prove_resident(sb_ctx* ctx, uint64_t handle, const uint8_t* blinders, uint8_t* proof_out);
int sb_plonk_prove(sb_ctx* ctx, uint64_t handle, const uint8_t* witness, uint64_t n_witness, const uint8_t* blinders,
                   uint8_t* proof_out);
int sb_plonk_release(sb_ctx* ctx, uint64_t handle);
/* ---- fflonk (src/fflonk_prove.js:51-1286), BN254 only like the reference's setup constants ----------------------
 * sb_fflonk_load: an fflonk zkey (protocol id 10, sections 2-17: src/zkey_utils.js:301-339, src/fflonk_constants.js).
 * sb_fflonk_prove: witness as for sb_plonk_prove; blinders = b_1..b_9 as 9 Montgomery field elements (:321-324);
 *   proof_out = C1 C2 W1 W2 (affine Montgomery) then the 16 evaluations ql qr qm qo qc s1 s2 s3 a b c z zw t1w t2w inv
 *   (Montgomery, 32 bytes each): sb_fflonk_proof_bytes().  Errors: "zkey file is not fflonk" (:71-73), "Invalid witness
 *   length. Circuit: N, witness: M, A" (:79-81), "Copy constraints does not match" (:649-651), "Polynomial is not
 *   divisible", "T0/T1/T2 Polynomial is not well calculated". */
int sb_fflonk_load(sb_ctx* ctx, const uint8_t* zkey, uint64_t zkey_len, uint64_t* handle);
int sb_fflonk_load_file(sb_ctx* ctx, const char* zkey_path, uint64_t* handle);
int sb_fflonk_info(sb_ctx* ctx, uint64_t handle, uint32_t* n_vars, uint32_t* n_public, uint32_t* domain_size, uint32_t* n_additions);
uint32_t sb_fflonk_proof_bytes(sb_ctx* ctx);
int sb_fflonk_prove_resident(sb_ctx* ctx, uint64_t handle, const uint8_t* blinders, uint8_t* proof_out);
int sb_fflonk_prove(sb_ctx* ctx, uint64_t handle, const uint8_t* witness, uint64_t n_witness, const uint8_t* blinders,
                    uint8_t* proof_out);
int sb_fflonk_release(sb_ctx* ctx, uint64_t handle);
/* multi-GPU: this rank proves with its shard [shard, n_shards) of every MSM and returns the five un-normalised
 * MSM partials (A, B1, C, H in G1; B2 in G2) instead of a proof; the ranks exchange them (NCCL all-gather) and any
 * rank finishes with sb_groth16_finish. */
int sb_groth16_prove_shard(sb_ctx* ctx, uint64_t handle, const uint8_t* witness, uint64_t n_witness,
                           int shard, int n_shards, uint8_t* partials_out);
uint32_t sb_groth16_partials_bytes(sb_ctx* ctx);
int sb_groth16_finish(sb_ctx* ctx, uint64_t handle, const uint8_t* partials_all_ranks, int n_shards,
                      const uint8_t r[32], const uint8_t s[32], uint8_t* proof_affine_out);

/* ---- multi-GPU inside the library (SURVEY §8b "NCCL comm built here", §8e) -------------------------------------------
 * One context per GPU, one NCCL rank per context.  libnccl.so.2 is dlopen'ed on first use (a copy already loaded in the
 * process, e.g. torch's, is reused), so single-GPU hosts need no NCCL.
 *   multi-process (one process per GPU, torchrun / a Node cluster): rank 0 calls sb_comm_unique_id and hands the 128 bytes
 *     to the other processes over any channel; every rank calls sb_comm_init_rank, loads its key shard with
 *     sb_groth16_load_sharded(rank, world) and then sb_groth16_prove_dist per proof (a collective call).
 *   single process: sb_create_multi / sb_groth16_load_multi / sb_groth16_prove_multi do the same with one host thread per
 *     device — the _multiExp chunk fan-out (14636-14658) and the worker pool (14064-14232) replaced by GPUs.
 * What a distributed proof exchanges: each rank uploads 1/world of the witness and an all-gather over NVLink completes it;
 * the iNTT -> coset-NTT chains of A, B, C run on ranks sb_dist_chain_owner(0..2) and their evaluations are sent to the rank
 * that owns each H range; every rank multiplies its point range of the five base sets; one all-gather of the (A, C', B2)
 * partials (a few hundred bytes) ends the proof.  Proof bytes equal the single-GPU ones. */
int sb_comm_unique_id(uint8_t out[128]);
int sb_comm_init_rank(sb_ctx* ctx, int world, int rank, const uint8_t id[128]);
int sb_comm_info(sb_ctx* ctx, int* rank, int* world);      /* world = 0 when the context has no communicator */
int sb_comm_destroy(sb_ctx* ctx);
int sb_dist_chain_owner(int chain, int world);
int sb_groth16_prove_dist(sb_ctx* ctx, uint64_t handle, const uint8_t* witness, uint64_t n_witness,
                          const uint8_t r[32], const uint8_t s[32], uint8_t* proof_affine_out_or_null);
int sb_create_multi(int curve, const int* device_ids, int n_devices, sb_ctx** out_contexts);
int sb_groth16_load_multi(sb_ctx* const* ctxs, int n, const uint8_t* zkey, uint64_t zkey_len, uint64_t* handles_out);
int sb_groth16_prove_multi(sb_ctx* const* ctxs, const uint64_t* handles, int n, const uint8_t* witness, uint64_t n_witness,
                           const uint8_t r[32], const uint8_t s[32], uint8_t* proof_affine_out);

/* Host-only halves of the multi-GPU path (no context, no device needed): combine partials gathered from the ranks and
 * assemble the proof (src/groth16_prove.js:103-132).  partial = extended-Jacobian X,Y,ZZ,ZZZ Montgomery bytes;
 * Groth16 partial block = A | B1 | C | H (G1) | B2 (G2). */
int sb_host_sum_partials(int curve, int group, const uint8_t* partials, int count, uint8_t* out_jacobian);
int sb_host_partial_from_affine(int curve, int group, const uint8_t* affine, uint8_t* partial_out);
uint32_t sb_host_partial_bytes(int curve, int group);
int sb_host_groth16_finish(int curve, const uint8_t* vk_alpha1, const uint8_t* vk_beta1, const uint8_t* vk_beta2,
                           const uint8_t* vk_delta1, const uint8_t* vk_delta2, const uint8_t* partials_all_ranks, int n_shards,
                           const uint8_t r[32], const uint8_t s[32], uint8_t* proof_affine_out);
void sb_shard_range(uint64_t total, int shard, int n_shards, uint64_t* first, uint64_t* count);

/* Device-resident variants (inputs already in HBM): what bench.py's `value` times.  Pointers are device pointers
 * in this context's device; out is host memory. */
int sb_msm_dev(sb_ctx* ctx, int group, const void* bases_dev, const void* scalars_dev, uint32_t scalar_bytes, uint64_t n, uint8_t* out);
int sb_ntt_fr_dev(sb_ctx* ctx, void* data_dev, void* scratch_dev, uint64_t n, int inverse, void** result_dev);
void* sb_dev_alloc(sb_ctx* ctx, uint64_t bytes);
int sb_dev_free(sb_ctx* ctx, void* p);
int sb_dev_upload(sb_ctx* ctx, void* dst_dev, const uint8_t* src, uint64_t bytes);
int sb_dev_download(sb_ctx* ctx, uint8_t* dst, const void* src_dev, uint64_t bytes);
/* timing of the last call on this context, measured with CUDA events on the context's stream (ms):
 * which = 0 total device time of the call, 1.. = per-stage breakdown where the call defines one.
 * sb_plonk_prove / sb_fflonk_prove: 1..5 = host wall clock of rounds 1..5 (each round ends on a synchronising commit). */
float sb_last_ms(sb_ctx* ctx, int which);
/* counters of the last MSM / prove call: 0/1 = summed device time (ms) of the G1 / G2 bucket-accumulation kernel
 * launches, 2/3 = number of those launches, 4/5 = (scalar digit, point) entries they consumed; 8..15 = device time (ms) per
 * kernel class: digits + radix sort, G1 accumulation, G2 accumulation, head folding, bucket reduction + window sums,
 * QAP rows, NTT passes, joinABC (meaningful per class when the call ran serialised, sb_set_tuning(2, 1)). */
double sb_last_stat(sb_ctx* ctx, int which);
/* integer-pipe calibration on this device: what = 0 -> IMAD.WIDE.U32 per second, 1 -> register-resident BN254 Fq
 * Montgomery multiplies per second (the modmul-bound roofline denominators). */
double sb_calibrate(sb_ctx* ctx, int what);
/* kernel-variant selection for experiments and profiling (process-wide; every variant computes the same bytes):
 *   0  minBlocksPerSM of the extension-field (G2) bucket accumulation (2 default, 3, 4)
 *   1  bucket reduction: 0 axis sums + warp-shuffle weighted sums (default), 1 legacy running sums, 2 hierarchical running sums
 *   2  1 = run every stream of a prove call serialised on one stream (per-kernel-class timing, sb_last_stat 8..15)
 *   3  1 = ignore the precomputed window tables (plain windowed Pippenger on the raw bases)
 *   4  2 = batched-affine pairing rounds before the accumulation;  5 = cap on their number (-1: also disables the adaptive
 *      entries-per-thread choice)
 *   6  log2 of the points per MSM chunk (test hook)        7  log2 of the largest NTT tile (10..12)
 *   8  0 = no pinned staging of pageable host buffers       9  forced sorted entries per accumulation thread (0 = adaptive)
 *   10 minBlocksPerSM of the 12-limb base-field (BLS12-381 G1) accumulation (2 default, 3, 4)
 *   12 minBlocksPerSM of the 8-limb base-field (BN254 G1) accumulation (4 default, 3, 2)
 *   11 lane-pair G2 accumulation (each point spread over two lanes, ec.cuh Fp2L): 0 = off, 3 / 4 = its minBlocksPerSM
 * The Python mirror applies SB_TUNE="key=value,..." from the environment when it loads the library. */
int sb_set_tuning(int key, int value);
/* synthetic bases for tests/benchmarks: chunks of 4096 points P_{c,j} = (k0(seed, c) + j*kd(seed)) * G, affine Montgomery, computed
 * on the GPU; the same points as the CPU oracle's incremental generator (oracle/snark_oracle.cpp or_gen_points), see msm.cuh. */
int sb_gen_points(sb_ctx* ctx, int group, uint64_t seed, uint64_t n, uint8_t* out);
int sb_generator(sb_ctx* ctx, int group, uint8_t* out_affine);
int sb_sync(sb_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif
