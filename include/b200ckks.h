/* b200ckks.h - C ABI of the B200-native RNS-CKKS evaluation engine (libb200ckks.so).
 *
 * This is the drop-in boundary for the hot path of tleong073/FHE-GPT-2: every entry point
 * below replaces one member of the reference's modified SEAL 3.6.6 C++ API (paths relative to
 * cnn_ckks/cpu-ckks/single-key/seal-modified-3.6.6/native/src/seal/ in the reference).  The
 * C++ facade in fhe-gpt-2_b200/host/seal/ re-creates the `seal::` classes on top of this ABI;
 * INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * Conventions
 *  - plain C types only: opaque handles, pointers, sizes, doubles; no C++/torch types.
 *  - every function returns a bk_status; on failure bk_last_error() (thread local) holds the
 *    message the reference would have thrown, and the status maps to the exception type:
 *    BK_INVALID_ARGUMENT -> std::invalid_argument, BK_LOGIC_ERROR -> std::logic_error,
 *    BK_OUT_OF_RANGE -> std::out_of_range.
 *  - all arithmetic runs on the GPU (sm_100a).  There is NO CPU fallback: creating a context
 *    without a CUDA device fails with BK_NO_DEVICE.
 *  - ciphertext layout on the host side of upload/download is SEAL's: [poly][limb][coeff],
 *    coeff fastest (ciphertext.h:335-347); key-switch keys [digit][poly][key limb][coeff]
 *    (kswitchkeys.h:340, one PublicKey per digit).
 *  - "limbs" = coeff_modulus_size of the object (chain_index + 1, context.cpp:455-523).
 *  - work is enqueued on a per-host-thread CUDA stream owned by the context (the reference
 *    shares one Evaluator between up to 50 OpenMP threads, infer_seal.cpp:404); objects are
 *    confined to the thread that produced them until bk_sync().
 */
#ifndef B200CKKS_H
#define B200CKKS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int bk_status;
enum {
    BK_OK = 0,
    BK_INVALID_ARGUMENT = 1,
    BK_LOGIC_ERROR = 2,
    BK_OUT_OF_RANGE = 3,
    BK_CUDA_ERROR = 4,
    BK_NO_DEVICE = 5
};

typedef struct bk_context_s *bk_context_t;
typedef struct bk_ct_s *bk_ct_t;         /* seal::Ciphertext   (ciphertext.h:54)   */
typedef struct bk_pt_s *bk_pt_t;         /* seal::Plaintext, CKKS NTT form          */
typedef struct bk_kskey_s *bk_kskey_t;   /* one std::vector<PublicKey> of KSwitchKeys (kswitchkeys.h:340) */
typedef struct bk_gkeys_s *bk_gkeys_t;   /* seal::GaloisKeys   (galoiskeys.h:48-74) */

const char *bk_last_error(void);
const char *bk_version(void);

/* ---- host-only helpers (no GPU needed) -------------------------------------------------- */
/* CoeffModulus::Create (modulus.cpp:143-182) + get_primes (util/numth.cpp:277-320). */
bk_status bk_coeff_modulus_create(int log_n, const int *bit_sizes, int count, uint64_t *primes_out);
/* try_minimal_primitive_root (util/numth.cpp:398-425): minimal primitive 2N-th root mod q. */
bk_status bk_minimal_primitive_root(int log_n, uint64_t q, uint64_t *root_out);
/* GaloisTool::get_elt_from_step (util/galois.cpp:53-95); step 0 = conjugation (2N-1). */
bk_status bk_galois_elt_from_step(int log_n, int step, uint32_t *elt_out);
/* GaloisTool::generate_table_ntt (util/galois.cpp:18-51): table_out[N]. */
bk_status bk_galois_table_ntt(int log_n, uint32_t galois_elt, uint32_t *table_out);
/* NTTTables root powers in SEAL's bit-reversed order (util/ntt.cpp:58-77); operand words only. */
bk_status bk_ntt_root_powers(int log_n, uint64_t q, int inverse, uint64_t *out);

/* ---- context (SEALContext + Evaluator + CKKSEncoder state; context.cpp:455-523) -------- */
/* primes: the full key-level chain, special prime last.  device: CUDA ordinal. */
bk_status bk_context_create(int log_n, const uint64_t *primes, int n_primes, int device, bk_context_t *out);
bk_status bk_context_destroy(bk_context_t ctx);
bk_status bk_context_info(bk_context_t ctx, int *log_n, int *n_primes, int *device);
bk_status bk_context_get_primes(bk_context_t ctx, uint64_t *primes_out);
/* tuning: number of output moduli processed per key-switch chunk (intermediates stay in L2). */
bk_status bk_context_set_ks_chunk(bk_context_t ctx, int chunk);
/* block the calling host thread until its stream has drained. */
bk_status bk_sync(bk_context_t ctx);
/* CUDA stream of the calling host thread (cudaStream_t as void*), for event timing / interop. */
bk_status bk_stream(bk_context_t ctx, void **stream_out);
/* Level-aware hybrid key switching (tolerance mode, off by default; $B200CKKS_HYBRID_KS=1 turns it on at context
 * creation).  Keys generated while it is on are recipes: at each level l a rotation / relinearization is first used at,
 * a level-specific key with ceil(l / alpha) digits over l + alpha moduli is generated on the device, the alpha - 1 idle
 * primes above the level joining the special prime as temporary special moduli.  Decrypted results equal those of
 * Evaluator::switch_key_inplace (evaluator.cpp:2281-2525) up to key-switching noise; ciphertext limbs do not, and the
 * keys are not in SEAL's layout.  Uploaded keys (bk_kskey_upload) always take SEAL's path. */
/* Randomness (replaces randomgen.h: the Blake2xb/SHAKE PRNG factories): key generation and encryption draw from
 * ChaCha20 in counter mode on the device.  Each context owns a 256-bit master key taken from the operating system
 * (getrandom) when the context is created; $B200CKKS_SEED or this call replace it for REPRODUCIBLE runs (tests,
 * benchmarks) - never set either in production.  The `seed` argument of the sampling calls below is a nonce that
 * separates the streams of different calls; 0 is as good as any other value. */
bk_status bk_context_set_rng_key(bk_context_t ctx, const uint8_t key[32]);
bk_status bk_context_set_hybrid(bk_context_t ctx, int on);
/* seed-compressed evaluation keys (SEAL ships the uniform half of a key as a seed, keygenerator.cpp:384-417,
 * util/rlwe.cpp:294-409; here the compressed form is what stays RESIDENT): level keys generated after this call keep only
 * polynomial 0 of every digit in HBM and regenerate the uniform polynomial from its public 256-bit ChaCha8 key each time
 * the key is used - half the key bytes for one expansion kernel per use.  Hybrid mode only.  $B200CKKS_COMPRESS_KEYS=1. */
bk_status bk_context_set_key_compression(bk_context_t ctx, int on);
bk_status bk_context_hybrid(bk_context_t ctx, int *on, uint64_t *key_bytes, uint64_t *keys);
/* shape of the level-aware key switch at `limbs` limbs: alpha special moduli (alpha - 1 idle primes + the special
 * prime), digits of dsize primes */
bk_status bk_context_hybrid_shape(bk_context_t ctx, int limbs, int *alpha_out, int *dsize_out);
/* block until every stream of the device has drained (before an object other host threads may be reading is freed). */
bk_status bk_sync_device(bk_context_t ctx);
/* Hand-over points between host threads (the reference runs one image per OpenMP thread over shared keys,
 * infer_seal.cpp:404): bk_event_record marks the work enqueued so far on the calling thread's stream,
 * bk_stream_wait_event makes the calling thread's stream wait for such a mark without blocking the host. */
typedef struct bk_event_s *bk_event_t;
bk_status bk_event_record(bk_context_t ctx, bk_event_t *event_out);
bk_status bk_stream_wait_event(bk_context_t ctx, bk_event_t event);
bk_status bk_event_destroy(bk_event_t event);
/* number of kernels this context has launched so far (all threads). */
bk_status bk_launch_count(bk_context_t ctx, uint64_t *count_out);

/* ---- measurement helpers ------------------------------------------------------------------ */
/* CUDA-event stopwatch on the calling thread's stream (nestable); bk_timer_end synchronises. */
bk_status bk_timer_begin(bk_context_t ctx);
bk_status bk_timer_end(bk_context_t ctx, double *ms_out);
/* live per-kernel timing: event pairs around every launch of one kernel family until
 * bk_profile_end.  Tags: 0 fwd column pass, 1 fwd block pass, 2 inv block pass, 3 inv column
 * pass, 4 key-switch inner product (k_ks_mac), 5 element-wise. */
bk_status bk_profile_begin(bk_context_t ctx, int kernel_tag);
bk_status bk_profile_end(bk_context_t ctx, uint64_t *launches_out, double *total_ms_out);
/* kernel_tag = -2 in bk_profile_begin times every family at once; this returns launches and milliseconds per
 * family: 0-5 as above, 6 encoder/decoder FFT, 7 other. */
bk_status bk_profile_end_all(bk_context_t ctx, uint64_t launches_out[8], double total_ms_out[8]);
/* always-on counters per kernel family: launches and limb-polynomials processed since context creation */
bk_status bk_kernel_counters(bk_context_t ctx, uint64_t launches_out[8], uint64_t units_out[8]);
/* bytes the C-ABI calls have copied host->device / device->host so far (uploads, encode inputs, downloads, ...) */
bk_status bk_transfer_bytes(bk_context_t ctx, uint64_t *h2d_out, uint64_t *d2h_out);
/* overwrite a 192 MiB scratch buffer (> the 126 MB L2) on the caller's stream. */
bk_status bk_flush_l2(bk_context_t ctx);
/* measured peak of 32-bit integer multiply-add thread-instructions per second on this GPU (the pipe that bounds the
 * NTT butterflies): the denominator of the benchmark's integer roofline */
bk_status bk_measure_imad_peak(bk_context_t ctx, double *imad_per_second_out);
/* the same for the three multiply-add forms a 64-bit modular butterfly is made of: rates_out = thread-instructions per
 * second of mad.lo.u32, mad.wide.u32 and mad.hi.u32 */
bk_status bk_measure_int_pipe(bk_context_t ctx, double rates_out[3]);

/* ---- ciphertext container (ciphertext.h) -------------------------------------------------- */
bk_status bk_ct_create(bk_context_t ctx, bk_ct_t *out);
bk_status bk_ct_destroy(bk_ct_t ct);
bk_status bk_ct_copy(bk_ct_t dst, bk_ct_t src);                 /* deep copy (value semantics) */
bk_status bk_ct_resize(bk_ct_t ct, int size, int limbs);        /* Ciphertext::resize(context, parms_id, size) */
bk_status bk_ct_info(bk_ct_t ct, int *size, int *limbs, double *scale, int *is_ntt);
bk_status bk_ct_set_scale(bk_ct_t ct, double scale);            /* Ciphertext::scale() is assignable */
bk_status bk_ct_set_ntt_form(bk_ct_t ct, int is_ntt);           /* Ciphertext::is_ntt_form() is assignable */
bk_status bk_ct_upload(bk_ct_t ct, const uint64_t *host, int size, int limbs, double scale, int is_ntt);
bk_status bk_ct_download(bk_ct_t ct, uint64_t *host_out);
/* raw device pointer ([poly][limb][coeff] in HBM) - for zero-copy interop with torch / NCCL. */
bk_status bk_ct_device_ptr(bk_ct_t ct, void **dev_ptr_out);

/* ---- plaintext container -------------------------------------------------------------------- */
bk_status bk_pt_create(bk_context_t ctx, bk_pt_t *out);
bk_status bk_pt_destroy(bk_pt_t pt);
bk_status bk_pt_copy(bk_pt_t dst, bk_pt_t src);
bk_status bk_pt_info(bk_pt_t pt, int *limbs, double *scale);
bk_status bk_pt_set_scale(bk_pt_t pt, double scale);
bk_status bk_pt_upload(bk_pt_t pt, const uint64_t *host, int limbs, double scale);
bk_status bk_pt_download(bk_pt_t pt, uint64_t *host_out);
/* Evaluator::mod_switch_to_inplace(Plaintext&) (evaluator.cpp:1248-1281,1350-1376): drop limbs. */
bk_status bk_pt_mod_switch_to(bk_pt_t pt, int limbs);

/* ---- keys ------------------------------------------------------------------------------------ */
/* Uploads one key-switching key in SEAL's layout [digits][2][n_primes][N] (kswitchkeys.h:340).
 * max_limbs > 0 keeps only digits < max_limbs and limbs < max_limbs (+ the special prime):
 * level-pruned residency for keys that are only ever used at <= max_limbs limbs. */
bk_status bk_kskey_upload(bk_context_t ctx, const uint64_t *host, int digits, int max_limbs, bk_kskey_t *out);
bk_status bk_kskey_destroy(bk_kskey_t key);
bk_status bk_kskey_info(bk_kskey_t key, int *digits, int *limbs, uint64_t *device_bytes);
/* Key plans.  The reference materialises every evaluation key up front (KeyGenerator::create_galois_keys,
 * infer_seal.cpp:379) and evaluates without the secret key.  In the level-aware hybrid mode a generated key is a
 * recipe that makes one key per level on first use - which needs the secret key at evaluation time.  To evaluate
 * WITHOUT it: learn the levels a workload uses (bk_kskey_levels after a dry run under any throw-away key; the set
 * is data independent), generate exactly those keys up front (bk_kskey_prepare_level), then detach the secret
 * (bk_kskey_drop_secret): from then on a missing level is BK_INVALID_ARGUMENT instead of a key generation. */
bk_status bk_kskey_levels(bk_kskey_t key, int *levels_out, int cap, int *count_out);
bk_status bk_kskey_prepare_level(bk_kskey_t key, int limbs);
bk_status bk_kskey_drop_secret(bk_kskey_t key);
/* resident part in SEAL's order: [digits][2][limbs+1][N], the special prime's limb last. */
bk_status bk_kskey_download(bk_kskey_t key, uint64_t *host_out);
/* device-to-device export/import of the resident key ([digits][2][limbs+1][N] words, engine
 * layout) - the buffer an NCCL broadcast moves between GPUs (keys are generated once). */
bk_status bk_kskey_export_device(bk_kskey_t key, void *dev_dst);
bk_status bk_kskey_import_device(bk_context_t ctx, const void *dev_src, int digits, int limbs, bk_kskey_t *out);
bk_status bk_gkeys_create(bk_context_t ctx, bk_gkeys_t *out);
bk_status bk_gkeys_destroy(bk_gkeys_t gk);                      /* destroys the keys it owns */
bk_status bk_gkeys_set(bk_gkeys_t gk, uint32_t galois_elt, bk_kskey_t key); /* takes ownership */
bk_status bk_gkeys_has(bk_gkeys_t gk, uint32_t galois_elt, int *has_out);
/* the key of one element (still owned by the set), for the key-plan calls above */
bk_status bk_gkeys_get(bk_gkeys_t gk, uint32_t galois_elt, bk_kskey_t *key_out);

/* Native key generation (keygenerator.cpp:64-76,164-233,384-417; util/rlwe.cpp:21-70,294-409).
 * The secret key is a device-resident plaintext-shaped object [n_primes][N] in NTT form. */
typedef struct bk_sk_s *bk_sk_t;
bk_status bk_sk_generate(bk_context_t ctx, int hamming_weight, uint64_t seed, bk_sk_t *out);
bk_status bk_sk_upload(bk_context_t ctx, const uint64_t *host /*[n_primes][N] NTT form*/, bk_sk_t *out);
bk_status bk_sk_download(bk_sk_t sk, uint64_t *host_out);
bk_status bk_sk_destroy(bk_sk_t sk);
bk_status bk_pk_generate(bk_context_t ctx, bk_sk_t sk, uint64_t seed, bk_ct_t pk_out /* size 2, key level */);
bk_status bk_relin_key_generate(bk_context_t ctx, bk_sk_t sk, uint64_t seed, int max_limbs, bk_kskey_t *out);
bk_status bk_galois_key_generate(bk_context_t ctx, bk_sk_t sk, uint32_t galois_elt, uint64_t seed, int max_limbs,
                                 bk_kskey_t *out);

/* ---- Evaluator (evaluator.h / evaluator.cpp) ------------------------------------------------ */
bk_status bk_add_inplace(bk_context_t ctx, bk_ct_t a, bk_ct_t b);              /* :103-163  */
bk_status bk_sub_inplace(bk_context_t ctx, bk_ct_t a, bk_ct_t b);              /* :190-246  */
bk_status bk_negate_inplace(bk_context_t ctx, bk_ct_t a);                      /* :76-101   */
bk_status bk_multiply_inplace(bk_context_t ctx, bk_ct_t a, bk_ct_t b);         /* ckks_multiply :673-814 */
bk_status bk_square_inplace(bk_context_t ctx, bk_ct_t a);                      /* ckks_square :1000-1059 */
bk_status bk_relinearize_inplace(bk_context_t ctx, bk_ct_t a, bk_kskey_t relin_key); /* :1061-1116 + :2281-2525 */
bk_status bk_rescale_to_next_inplace(bk_context_t ctx, bk_ct_t a);             /* :1118-1181,1378-1414; rns.cpp:737-808 */
/* relinearize_inplace then rescale_to_next_inplace as one call (the pair every ciphertext product of the
 * reference's polynomial evaluations ends with: common/Polynomial.cpp:242-252, comp/SEALfunc.cpp:18-24,
 * ckks_bootstrapping/ModularReducer.cpp:43-46).  In tolerance mode (hybrid key switching, a level with idle primes)
 * the ModDown by P_S and the division by q_last are one division by q_last * P_S; otherwise exactly the two calls. */
bk_status bk_relinearize_rescale_inplace(bk_context_t ctx, bk_ct_t a, bk_kskey_t relin_key);
bk_status bk_mod_switch_to_next_inplace(bk_context_t ctx, bk_ct_t a);          /* :1183-1246 */
bk_status bk_mod_switch_to_inplace(bk_context_t ctx, bk_ct_t a, int limbs);    /* :1326-1348 */
bk_status bk_apply_galois_inplace(bk_context_t ctx, bk_ct_t a, uint32_t galois_elt, bk_gkeys_t gk);
/* out of place (Evaluator::apply_galois / rotate_vector / complex_conjugate with a destination, evaluator.h:1120-1180):
 * `a` is read, dst receives the result - the reference copies `a` into the destination and works in place
 * (evaluator.h:1131-1137); the key switch here writes a fresh buffer anyway, so the copy is not made */
bk_status bk_apply_galois(bk_context_t ctx, bk_ct_t a, uint32_t galois_elt, bk_gkeys_t gk, bk_ct_t dst); /* :2120-2222 */
bk_status bk_rotate_vector_inplace(bk_context_t ctx, bk_ct_t a, int steps, bk_gkeys_t gk);         /* :2224-2279 */
bk_status bk_complex_conjugate_inplace(bk_context_t ctx, bk_ct_t a, bk_gkeys_t gk);                /* evaluator.h:1321-1341 */
/* Hoisted automorphisms (engine extension; the reference carries the idea as dead code, Bootstrapper.cpp:2088-2230,
 * keygenerator.cpp:236-304): outs[k] = apply_galois(in, galois_elts[k]) for all k with ONE decomposition and digit NTT
 * of `in`.  Decrypts to the same values as bk_apply_galois_inplace up to key-switching noise, but the limbs are not
 * those of evaluator.cpp:2120-2222 (which permutes before decomposing) - a tolerance-mode fast path for the baby steps
 * of BSGS linear transforms.  outs[k] must be distinct ciphertext objects different from `in`. */
bk_status bk_apply_galois_hoisted(bk_context_t ctx, bk_ct_t in, const uint32_t *galois_elts, int count, bk_gkeys_t gk,
                                  bk_ct_t *outs);
/* Double-hoisted inner sums of a baby-step / giant-step linear transform (Bootstrapper::bsgs_linear_transform,
 * Bootstrapper.cpp:1952-2016, inner loop :1995-2012): outs[g] = sum_k rotate(ct, baby_k) (.) pts[g * n_baby + k].
 * The input is decomposed once, the baby rotations stay in the extended basis Q_l * P_S, the plaintexts (extended,
 * bk_encode_ext) are multiplied there, and the division by P_S is done once per giant step instead of once per baby
 * rotation.  elts[k] = 1 means "no rotation"; NULL plaintexts are skipped.  Level-aware hybrid mode only
 * (BK_LOGIC_ERROR otherwise); tolerance mode: decrypted values equal the rotation-by-rotation sequence up to
 * key-switching noise.  rescale != 0: outs[g] is rescale_to_next of that sum - the rescale the reference applies after
 * the whole transform (Bootstrapper.cpp:2018-2086 callers) commutes with the giant-step rotations and the final sum,
 * so it is taken here as one division by q_last * P_S with the ModDown, and the giant steps run one level lower. */
bk_status bk_bsgs_inner_sums(bk_context_t ctx, bk_ct_t in, const uint32_t *elts, int n_baby, bk_gkeys_t gk,
                             const bk_pt_t *pts, int n_giant, bk_ct_t *outs, int rescale);
/* CKKSEncoder::encode at `limbs` limbs plus the special moduli the level-aware key switch uses at that level (the
 * operand format of bk_bsgs_inner_sums); usable as an ordinary plaintext of that level as well. */
bk_status bk_encode_ext(bk_context_t ctx, const double *values, int n_values, int is_complex, int limbs, double scale,
                        bk_pt_t out);
bk_status bk_add_plain_inplace(bk_context_t ctx, bk_ct_t a, bk_pt_t p);        /* :1578-1650 */
bk_status bk_sub_plain_inplace(bk_context_t ctx, bk_ct_t a, bk_pt_t p);        /* :1652-1724 */
bk_status bk_multiply_plain_inplace(bk_context_t ctx, bk_ct_t a, bk_pt_t p);   /* :1726-1761,1891-1930 */
/* acc <- acc + a (*) p: multiply_plain (:1891-1930) and add_inplace (:103-163) in one pass over the data; an empty
 * acc (size 0) is initialised with the product.  acc takes the product's scale (the reduced-error add overwrites
 * scales the same way, :316-321).  Residues equal those of the two separate calls. */
bk_status bk_multiply_plain_accumulate(bk_context_t ctx, bk_ct_t acc, bk_ct_t a, bk_pt_t p);
/* dst = sum_t cts[t] (.) pts[t]: the inner sum of a BSGS group (Bootstrapper.cpp:1995-2012) or of the filter taps of a
 * convolution (cnn_seal.cpp:455-470) in one pass over the operands; same residues as multiply_plain + add_inplace term
 * by term.  All operands on one level with equal scales; dst must not be an operand. */
bk_status bk_multiply_plain_sum(bk_context_t ctx, bk_ct_t dst, const bk_ct_t *cts, const bk_pt_t *pts, int count);
/* dst = constant + sum_j values[j] * cts[j] (1 <= count <= 8) at the lowest level among the sources and at scale
 * target_scale: term j's scalar is encoded at target_scale / scale(cts[j]).  Tolerance-mode replacement for the
 * multiply_const + rescale_to_next + add_reduced_error chains of the reference's polynomial evaluation leaves
 * (cnn_ckks/common/Polynomial.cpp:438-456): one rescale per leaf is left to the caller instead of one per term. */
bk_status bk_scalar_linear_combination(bk_context_t ctx, bk_ct_t dst, const bk_ct_t *cts, const double *values, int count,
                                       double constant, double target_scale);
bk_status bk_transform_to_ntt_inplace(bk_context_t ctx, bk_ct_t a);            /* :2069-2118 */
bk_status bk_transform_from_ntt_inplace(bk_context_t ctx, bk_ct_t a);
/* fork: add_const / multiply_const (evaluator.cpp:287-302): scalar encode (ckks.cpp:77-153)
 * fused into the element-wise kernel - no plaintext is materialised. */
bk_status bk_add_const_inplace(bk_context_t ctx, bk_ct_t a, double value);
bk_status bk_multiply_const_inplace(bk_context_t ctx, bk_ct_t a, double value);
/* Bootstrapper::modraise_inplace (ckks_bootstrapping/Bootstrapper.cpp:2894-2948). */
bk_status bk_modraise_inplace(bk_context_t ctx, bk_ct_t a);

/* ---- raw kernels (ntt.h:235-264,336-358) on device-resident limbs: data = [count][N], limb k
 * uses prime prime_idx[k].  Exposed for parity tests and kernel benchmarks. ------------------ */
bk_status bk_ntt_limbs(bk_context_t ctx, uint64_t *dev_data, const int *prime_idx, int count, int inverse);
bk_status bk_ntt_limbs_host(bk_context_t ctx, uint64_t *host_data, const int *prime_idx, int count, int inverse);

/* ---- CKKSEncoder (ckks.h:457-761, ckks.cpp:13-273) ---------------------------------------- */
/* values: n_values doubles, or n_values (re,im) pairs when is_complex; encoded at `limbs`. */
bk_status bk_encode(bk_context_t ctx, const double *values, int n_values, int is_complex, int limbs, double scale,
                    bk_pt_t out);
/* encode(values, scale, pt) at the TOP level followed by mod_switch_to_inplace(pt, limbs)
 * (evaluator.h:1270-1278 does exactly this for every multiply_vector): same residues, same
 * range checks (against the top level), but only `limbs` limbs are ever computed. */
bk_status bk_encode_top_dropped(bk_context_t ctx, const double *values, int n_values, int is_complex, int limbs,
                                double scale, bk_pt_t out);
bk_status bk_encode_scalar(bk_context_t ctx, double value, int limbs, double scale, bk_pt_t out);
bk_status bk_decode(bk_context_t ctx, bk_pt_t pt, double *out_complex /* slot_count (re,im) pairs */);
bk_status bk_set_sparse_slots(bk_context_t ctx, int sparse_slots);   /* ckks.h:446-450,704-713 */

/* ---- Encryptor / Decryptor (encryptor.cpp:88-239, decryptor.cpp:150-183) ---------------- */
bk_status bk_encrypt(bk_context_t ctx, bk_ct_t pk, bk_pt_t pt, uint64_t seed, bk_ct_t out);
bk_status bk_encrypt_symmetric(bk_context_t ctx, bk_sk_t sk, bk_pt_t pt, uint64_t seed, bk_ct_t out);
bk_status bk_decrypt(bk_context_t ctx, bk_sk_t sk, bk_ct_t ct, bk_pt_t out);

#ifdef __cplusplus
}
#endif
#endif /* B200CKKS_H */
