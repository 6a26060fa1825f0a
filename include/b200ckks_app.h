/* b200ckks_app.h - C ABI of the application layers restated on top of the engine (libb200ckks_app.so):
 * bootstrapping, approximate ReLU, multiplexed-packing CNN operators and the ResNet driver of
 * tleong073/FHE-GPT-2's cnn_ckks.  It exists so that the Python harness (tests/, bench.py) and foreign-language
 * hosts can drive the C++ classes in fhe-gpt-2_b200/host/ - which mirror the reference's own classes - without a
 * C++ compiler.  A C++ caller uses those classes directly.
 *
 * Reference interfaces replaced (paths under cnn_ckks/ in the reference):
 *   bka_session_*        the set-up block of ResNet_cifar10_seal_sparse, cpu-ckks/single-key/cnn/infer_seal.cpp:288-342
 *                        (EncryptionParameters, SEALContext, KeyGenerator, pk / relin / Galois keys, encoder, ...)
 *   bka_bootstrapper_*   class Bootstrapper, cpu-ckks/single-key/ckks_bootstrapping/Bootstrapper.h:14-201
 *   bka_relu             minimax_ReLU_seal, cpu-ckks/single-key/comp/SEALcomp.cpp:3-60
 *   bka_conv / bka_bn / bka_downsample / bka_avgpool / bka_fc / bka_tensor_add
 *                        cpu-ckks/single-key/cnn/cnn_seal.cpp:284-787
 *   bka_resnet_*         ResNet_cifar10_seal_sparse, cpu-ckks/single-key/cnn/infer_seal.cpp:251-584
 *
 * Conventions: as include/b200ckks.h (int status, bka_last_error(), opaque handles).  Slot vectors cross the ABI
 * as doubles (real) or (re, im) pairs (complex).  All ciphertext arithmetic runs on the GPU.
 */
#ifndef B200CKKS_APP_H
#define B200CKKS_APP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bka_session_s *bka_session_t;
typedef struct bka_ct_s *bka_ct_t;               /* seal::Ciphertext */
typedef struct bka_bootstrapper_s *bka_bootstrapper_t;
typedef struct bka_resnet_s *bka_resnet_t;

const char *bka_last_error(void);
/* "engine" when linked against libb200ckks.so; "reference-seal" for the CPU oracle build (oracle/_ref). */
const char *bka_backend(void);

/* ---- session: parameters, keys, encoder, evaluator (infer_seal.cpp:288-342) ------------------------------------
 * bit_sizes: CoeffModulus::Create sizes, special prime last.  rotation_steps: the steps handed to
 * KeyGenerator::create_galois_keys (0 = conjugation); more can be added until the first rotation. */
int bka_session_create(int log_n, const int *bit_sizes, int n_bits, int hamming_weight, int device,
                       const int *rotation_steps, int n_steps, bka_session_t *out);
/* same, with a given secret key ([n_bits][N] NTT-form words from bka_session_secret_key of another session with the
 * same parameters; NULL = sample a fresh one): one secret for every GPU of a node (engine backend only) */
int bka_session_create_with_secret(int log_n, const int *bit_sizes, int n_bits, int hamming_weight, int device,
                                   const int *rotation_steps, int n_steps, const uint64_t *secret_key, bka_session_t *out);
int bka_session_secret_key(bka_session_t s, uint64_t *host_out /* [n_bits][N] */);
int bka_session_destroy(bka_session_t s);
/* the engine's bk_context_t behind this session (include/b200ckks.h), for its measurement helpers; NULL on the
 * reference-SEAL backend */
int bka_session_engine_context(bka_session_t s, void **bk_context_out);
/* operation counts by coeff_modulus_size of the ciphertext operand since the last reset.  which: 0 key switches
 * (rotate + relinearize), 1 rescales, 2 vector encode + multiply_plain, 3 ct x ct multiplications, 4 scalar ops,
 * 5 additions */
int bka_session_level_histogram(bka_session_t s, int which, uint64_t counts_out[64], int reset);
int bka_session_add_rotation_steps(bka_session_t s, const int *steps, int n_steps);
int bka_session_primes(bka_session_t s, uint64_t *primes_out /* n_bits */);
int bka_session_sync(bka_session_t s);
/* Evaluator operation counters since the last reset: rotate/conjugate key switches, relinearizations, rescales,
 * ct x ct multiplications, ct x pt multiplications, vector encodes, additions, mod switches, scalar ops. */
int bka_session_stats(bka_session_t s, uint64_t counts_out[9], int reset);
/* bytes of Galois keys resident in HBM and number of key generations so far (engine backend; 0 otherwise) */
int bka_session_key_residency(bka_session_t s, uint64_t *bytes_out, uint64_t *generated_out);
/* Key plans (host/seal/seal.h KeyPlan; include/b200ckks.h bk_kskey_levels ...): the (Galois element, level) pairs and
 * relinearization levels the session's keys cover so far, as text; and the reverse - generate exactly the keys of a
 * plan now and, with detach_secret != 0, drop every reference the evaluation keys hold on the secret key, after which
 * the Evaluator can no longer generate keys (a rotation outside the plan fails like a missing key in the reference).
 * A plan is data- and key-independent: capture it once from a dry run under a throw-away key. */
int bka_session_key_plan(bka_session_t s, char *text_out, int cap, int *length_out);
int bka_session_apply_key_plan(bka_session_t s, const char *text, int detach_secret);
/* encoded plaintext operands kept resident in HBM (bootstrapping diagonals, convolution weight masks; see
 * host/common/cached.h): bytes, and hits / misses of the cache since the session was created */
int bka_session_plain_cache(bka_session_t s, uint64_t *bytes_out, uint64_t *hits_out, uint64_t *misses_out);

/* ---- ciphertexts ------------------------------------------------------------------------------------------------ */
/* encode at the top level with `scale`, encrypt with the public key, mod_switch_to `limbs` (0 = stay on top) */
int bka_encrypt(bka_session_t s, const double *values, int n_values, int is_complex, double scale, int limbs,
                bka_ct_t *out);
int bka_decrypt(bka_session_t s, bka_ct_t ct, double *out_complex /* slot_count (re, im) pairs */);
int bka_ct_clone(bka_ct_t ct, bka_ct_t *out);
int bka_ct_free(bka_ct_t ct);
int bka_ct_info(bka_ct_t ct, int *size, int *limbs, double *scale);
int bka_ct_set_scale(bka_ct_t ct, double scale);
int bka_ct_mod_switch_to(bka_session_t s, bka_ct_t ct, int limbs);
/* raw limbs [size][limbs][N] (for limb-level comparisons) */
int bka_ct_download(bka_ct_t ct, uint64_t *host_out);

/* a few Evaluator members, enough to compose test programs through this ABI */
int bka_rotate(bka_session_t s, bka_ct_t ct, int steps);
int bka_multiply_relin_rescale(bka_session_t s, bka_ct_t a, bka_ct_t b); /* a <- rescale(relin(a * b)) */
int bka_add_reduced_error(bka_session_t s, bka_ct_t a, bka_ct_t b);      /* a <- a + b */
int bka_multiply_vector_rescale(bka_session_t s, bka_ct_t a, const double *values, int n_values, int is_complex);
/* SEAL 3.6's binary wire format (Ciphertext / RelinKeys / GaloisKeys / SecretKey / PublicKey ::save and ::load,
 * seal/serialization.h, compr_mode_type::none) through files, interchangeable with the reference's own SEAL.
 * what: 0 ciphertext (ct / ct_out), 2 relinearization keys, 3 Galois keys, 4 secret key, 5 public key of the session.
 * Keys are saved at full size in SEAL's layout (not available for level-aware hybrid keys); loaded keys are complete,
 * so nothing is generated afterwards. */
int bka_save(bka_session_t s, int what, bka_ct_t ct, const char *path);
int bka_load(bka_session_t s, int what, const char *path, bka_ct_t *ct_out);
/* Evaluator::{add,sub,multiply}_inplace_reduced_error (evaluator.cpp:312-486; which = 0, 1, 2): a <- a op b, operands
 * may sit at different levels (the higher one is walked down as the reference does); multiply relinearizes. */
int bka_reduced_error_op(bka_session_t s, int which, bka_ct_t a, bka_ct_t b);
/* a ciphertext from raw limbs in the reference's layout [size][limbs][N] (Ciphertext::data(), ciphertext.h:335-347) */
int bka_ct_upload(bka_session_t s, const uint64_t *host, int size, int limbs, double scale, int is_ntt, bka_ct_t *out);
/* replace the session's relinearization key by one in SEAL's layout [digits][2][n_primes][N] (kswitchkeys.h:340) */
int bka_session_import_relin_key(bka_session_t s, const uint64_t *host, int digits);

/* ---- bootstrapping (Bootstrapper.h) ----------------------------------------------------------------------------
 * create = constructor + prepare_mod_polynomial + slot_vec.push_back(logn) + generate_LT_coefficient_3; the
 * rotation steps it needs (addLeftRotKeys_Linear_to_vector_3 plus the power-of-two steps and conjugation) are
 * added to the session's Galois key set. */
int bka_bootstrapper_create(bka_session_t s, int loge, int logn, int total_level, double final_scale, int boundary_k,
                            int sin_cos_deg, int scale_factor, int inverse_deg, bka_bootstrapper_t *out);
int bka_bootstrapper_destroy(bka_bootstrapper_t b);
/* baby-step rotations through one shared decomposition (engine; default on unless $B200CKKS_NO_HOIST) or one by one
 * exactly as the reference issues them (0).  Returns the previous setting in *previous (may be NULL). */
int bka_bootstrapper_set_hoisting(bka_bootstrapper_t b, int on, int *previous);
/* `on` above: bit 0 = hoisted baby steps; bit 1 set = WITHOUT the double-hoisted inner sums (one ModDown per giant step,
 * Evaluator::bsgs_inner_sums_cached; level-aware hybrid mode), which are otherwise used together with hoisting unless
 * $B200CKKS_NO_DOUBLE_HOIST is set.  bka_session_double_hoisted_groups: giant steps served that way so far. */
int bka_session_double_hoisted_groups(bka_session_t s, uint64_t *count_out);
/* steps of addLeftRotKeys_Linear_to_vector_3 for this logn, appended to steps_out (capacity cap) */
int bka_bootstrapper_rotation_steps(bka_bootstrapper_t b, int *steps_out, int cap, int *count_out);
/* LT coefficients: which = 0..2 SlotToCoeff matrices 1..3, 3..5 CoeffToSlot matrices 1..3.
 * Returns the number of diagonals and their length; data_out (may be NULL) receives (re, im) pairs. */
int bka_bootstrapper_lt_coefficients(bka_bootstrapper_t b, int which, int *n_diagonals, int *length, double *data_out);
/* Replaces the baby-step/giant-step heap of the EvalMod polynomial (boot::Polynomial::generate_poly_heap) by given
 * coefficients: data = {heaplen, heap_k, heap_m, scale_inverse_coeff, then per heap node: degree (-1 = absent) followed
 * by degree + 1 Chebyshev coefficients}.  For callers that bring their own minimax polynomial (the tests feed the
 * coefficients of the reference's own Remez run, ModularReducer.cpp:37-51, to compare bootstraps limb by limb). */
int bka_bootstrapper_set_evalmod_heap(bka_bootstrapper_t b, const double *data, int count);
/* bootstrap_3 (real_message = 0) / bootstrap_real_3 (real_message = 1); ct is consumed like the reference's input */
int bka_bootstrap(bka_bootstrapper_t b, bka_ct_t ct, int real_message, bka_ct_t *out);
/* EvalMod alone: ModularReducer::modular_reduction */
int bka_modular_reduction(bka_bootstrapper_t b, bka_ct_t ct, bka_ct_t *out);

/* ---- GPT-2 operators (gpt2_ckks/gpt2-ckks/single-key/gpt2/approx.h: MatrixMul.cpp, PolyApprox.cpp, IterApprox.cpp,
 * Fold.cpp, pack.cpp, optimize.cpp, util.cpp) ---------------------------------------------------------------------
 * One entry point selected by operator name; ciphertext arguments in `in`, the operator's integer / double
 * arguments in iparams / dparams, results as new handles in `out` (count in *n_out).  Inputs are never modified
 * (operators the reference runs in place work on a copy that is returned).  `boot` may be NULL except for
 * "bootstrap", "quickMax" below 18 limbs and "softmax".
 *   quickSum(i: n)  sign_f  sign_g  gelu_p  gelu_q  cheby_basis(i: n)  sign(i: df, dg)  gelu  exp(i: r)
 *   inverse(i: iters)  taylor(i: iters; d: guess)  inv_sqrt(i: iters; d: guess)
 *   layernorm(i: row_size; d: gamma[row_size], beta[row_size])  max(2 cts)  quickMax(i: n)
 *   smax(i: r, gamma)  softmax(i: r, unused)  mask_out(i: start, length)  rotate_inplace(i: steps)
 *   surefire_rotate(i: shift)  fake_bootstrap  bootstrap  init_output(i: count)
 *   pack_from_row(i: rows, cols; d: matrix)  expand_bias(d: bias)  pack_tight(8 + 3 cts)  unpack_tight(3 + 8 cts)
 *   row_matmul / attn_proj_row / attn_proj_col(i: n_left, n_weights, n_outputs, A_rows, A_cols, W_rows, W_cols;
 *       cts: left.., weights.., bias, outputs..)   col_matmul(i: rows, cols; cts: left.., right..)
 *   qk_matmul / sv_matmul(i: heads, n_outputs; cts: first.., second.., outputs..)
 *   augment_row / augment_col(i: n, padded_row_size, idx; cts: A.., cached..) */
int bka_gpt2_call(bka_session_t s, bka_bootstrapper_t boot, const char *op, bka_ct_t *in, int n_in, const double *dparams, int n_d,
                  const int *iparams, int n_i, bka_ct_t *out, int out_cap, int *n_out);
/* the 37-prime chain {49, 46 x 21, 49 x 14, 60} and the rotation-step list of the reference's INIT macro (util.h:37-75) */
int bka_gpt2_init_chain(int *bits_out, int bits_cap, int *n_bits, int *steps_out, int steps_cap, int *n_steps);

/* ---- approximate ReLU (alpha = 13: comp_no 3, degrees {15,15,27}, scaled_val 1.7; infer_seal.cpp:255-262) ------ */
int bka_relu(bka_session_t s, bka_ct_t ct, bka_ct_t *out);
/* evaluation trees of upgrade_oddbaby(deg): heap array (capacity cap), m, l */
int bka_oddbaby_tree(int deg, int *tree_out, int cap, int *len_out, int *depth_out, int *m_out, int *l_out);

/* ---- multiplexed-packing tensors (cnn_seal.h:20-46): packing parameters travel as int[7] = k,h,w,c,t,p,logn ---- */
int bka_conv(bka_session_t s, bka_ct_t in, const int in_parms[7], int co, int st, int fh, int fw, const double *weight,
             const double *running_var, const double *constant_weight, double epsilon, int end, bka_ct_t *out,
             int out_parms[7]);
int bka_bn(bka_session_t s, bka_ct_t in, const int parms[7], const double *bias, const double *running_mean,
           const double *running_var, const double *weight, double epsilon, double B, bka_ct_t *out);
int bka_downsample(bka_session_t s, bka_ct_t in, const int in_parms[7], bka_ct_t *out, int out_parms[7]);
int bka_avgpool(bka_session_t s, bka_ct_t in, const int in_parms[7], double B, bka_ct_t *out, int out_parms[7]);
int bka_fc(bka_session_t s, bka_ct_t in, const int parms[7], const double *matrix, const double *bias, int q, int r,
           bka_ct_t *out);
int bka_tensor_add(bka_session_t s, bka_ct_t a, bka_ct_t b, bka_ct_t *out);

/* ---- ResNet on CIFAR-10 (infer_seal.cpp:251-584) ---------------------------------------------------------------
 * The session must have been created with the CNN chain.  Weights in the reference's file order:
 * conv_weight[layer_num-1][9*ci*co], bn_{bias,mean,var,weight}[layer_num-1][co], linear_weight[10*64],
 * linear_bias[10], each list flattened and concatenated.  create builds the three bootstrappers (logn 14/13/12),
 * the evaluation trees and registers every rotation step of infer_seal.cpp:345-378. */
int bka_resnet_create(bka_session_t s, int layer_num, const double *conv_weight, const double *bn_bias,
                      const double *bn_mean, const double *bn_var, const double *bn_weight, const double *linear_weight,
                      const double *linear_bias, bka_resnet_t *out);
/* The CIFAR-100 network the fork has commented out (ResNet_cifar100_seal_sparse, infer_seal.cpp:585-891): B = 65, 100
 * classes (linear_weight 100 x 64, linear_bias 100), and 1x1 stride-2 shortcut convolutions with batch norm at the two
 * down-sampling blocks: shortcut_weight = [16*32 | 32*64] values, shortcut_bn_* = [32 | 64] values. */
int bka_resnet_create_cifar100(bka_session_t s, int layer_num, const double *conv_weight, const double *bn_bias,
                               const double *bn_mean, const double *bn_var, const double *bn_weight, const double *linear_weight,
                               const double *linear_bias, const double *shortcut_weight, const double *shortcut_bn_bias,
                               const double *shortcut_bn_mean, const double *shortcut_bn_var, const double *shortcut_bn_weight,
                               bka_resnet_t *out);
/* number of logits every inference call of this network writes (10 or 100) */
int bka_resnet_classes(bka_resnet_t net, int *classes_out);
int bka_resnet_destroy(bka_resnet_t net);
/* image: 3*32*32 doubles (CHW).  logits_out: 10 doubles.  trace_out (may be NULL): per stage
 * {op code, remaining level, scale, milliseconds}, up to trace_cap rows; trace_rows receives the row count.
 * op codes: 0 conv, 1 bn, 2 relu, 3 bootstrap, 4 add, 5 downsample, 6 avgpool, 7 fc. */
int bka_resnet_infer(bka_resnet_t net, const double *image, double *logits_out, double *trace_out, int trace_cap,
                     int *trace_rows);
/* the three phases of bka_resnet_infer separately: client-side packing + encryption (infer_seal.cpp:434-453), the
 * encrypted network on a ciphertext resident in HBM (:455-537), client-side decryption of the logits (:543-551) */
int bka_resnet_encrypt_image(bka_resnet_t net, const double *image, bka_ct_t *out);
int bka_resnet_infer_encrypted(bka_resnet_t net, bka_ct_t image_ct, bka_ct_t *logits_ct, double *trace_out, int trace_cap,
                               int *trace_rows);
int bka_resnet_decrypt_logits(bka_resnet_t net, bka_ct_t logits_ct, double *logits_out);
/* The reference's image loop (`#pragma omp parallel for` over images sharing the keys, infer_seal.cpp:404-577):
 * n_images images, up to in_flight of them concurrently, one host thread and one CUDA stream per image in flight.
 * images: n_images x 3072 doubles, logits_out: n_images x 10.  The encrypted variant keeps inputs and outputs in
 * HBM.  Work of the batch is ordered after what the calling thread enqueued before and before what it enqueues next. */
int bka_resnet_infer_batch(bka_resnet_t net, const double *images, int n_images, int in_flight, double *logits_out);
int bka_resnet_infer_encrypted_batch(bka_resnet_t net, bka_ct_t *image_cts, int n_images, int in_flight, bka_ct_t *logits_cts);

#ifdef __cplusplus
}
#endif
#endif /* B200CKKS_APP_H */
