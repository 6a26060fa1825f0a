// gpt2/approx.h - the GPT-2 operators over packed ciphertexts: matrix products, polynomial and iterative
// approximations of the non-linear functions, folds, KV-cache helpers.
//
// Same names, argument order and operation sequences as gpt2_ckks/gpt2-ckks/single-key/gpt2/approx.h of the
// reference (MatrixMul.cpp, PolyApprox.cpp, IterApprox.cpp, Fold.cpp, optimize.cpp), so that decrypted results,
// levels and scales match the reference's on the same inputs; unused parameters are kept for source compatibility.
// The transformer block assembly of the reference (layers.cpp, ChebyPoly.cpp) does not compile there and is not
// restated.
#pragma once
#include "gpt2/pack.h"
#include "gpt2/util.h"

namespace gpt2
{
#define GPT2_CKKS_ARGS                                                                                                 \
    seal::CKKSEncoder &encoder, seal::Encryptor &encryptor, seal::Decryptor &decryptor, seal::Evaluator &evaluator,     \
        seal::GaloisKeys &gal_keys, seal::RelinKeys &relin_keys

    // ---- matrix products (MatrixMul.cpp)
    void col_matrix_multiplication_seal(std::vector<TensorCipher> &left_inputs, std::vector<TensorCipher> &right_inputs,
                                        std::vector<TensorCipher> &outputs, std::vector<double> bias, int rows, int cols, Config &config,
                                        GPT2_CKKS_ARGS);
    void row_matrix_multiplication_seal(vc &left_inputs, vc &weights, seal::Ciphertext bias, vc &outputs, int A_rows, int A_cols,
                                        int W_rows, int W_cols, GPT2_CKKS_ARGS);
    void diagonal_to_row_matrix_seal(std::vector<TensorCipher> &inputs, std::vector<TensorCipher> &outputs, int rows, int cols,
                                     Config &config, GPT2_CKKS_ARGS);
    void attn_proj_row_seal(vc &left_inputs, vc &weights, seal::Ciphertext bias, vc &outputs, int A_rows, int A_cols, int W_rows,
                            int W_cols, seal::KeyGenerator &keygen, GPT2_CKKS_ARGS);
    void attn_proj_col_seal(vc &left_inputs, vc &weights, seal::Ciphertext bias, vc &outputs, int A_rows, int A_cols, int W_rows,
                            int W_cols, seal::KeyGenerator &keygen, GPT2_CKKS_ARGS);
    void qk_matmul(vc &Q, vc &K, vc &outputs, int A_rows, int A_cols, int W_rows, int W_cols, seal::KeyGenerator &keygen,
                   GPT2_CKKS_ARGS);
    void sv_matmul(vc &S, vc &V, vc &outputs, int A_rows, int A_cols, int W_rows, int W_cols, seal::KeyGenerator &keygen,
                   GPT2_CKKS_ARGS);
    void cipher_plain_128_128(seal::Ciphertext &left_input, std::unordered_map<std::string, std::vector<double>> &weights,
                              seal::Ciphertext bias, vc &outputs, int A_rows, int A_cols, int W_rows, int W_cols,
                              seal::KeyGenerator &keygen, GPT2_CKKS_ARGS);
    void batch_matmul(vc &left_input, std::unordered_map<std::string, std::vector<double>> &weights, seal::Ciphertext bias, vc &outputs,
                      int A_rows, int A_cols, int W_rows, int W_cols, seal::KeyGenerator &keygen, GPT2_CKKS_ARGS);
    void qk_matmul_col(vc &left_input, vc &right_input, std::unordered_map<std::string, std::vector<double>> &weights,
                       seal::Ciphertext bias, vc &outputs, int A_rows, int A_cols, int W_rows, int W_cols, seal::KeyGenerator &keygen,
                       GPT2_CKKS_ARGS);

    // ---- polynomial approximations (PolyApprox.cpp)
    void build_cheby_basis(seal::Ciphertext &input, vc &chebyBasis, int n, GPT2_CKKS_ARGS);
    void compute_sign_f(seal::Ciphertext &input, seal::Ciphertext &output, GPT2_CKKS_ARGS);
    void compute_sign_g(seal::Ciphertext &input, seal::Ciphertext &output, GPT2_CKKS_ARGS);
    void sign_function(TensorCipher &inputs, TensorCipher &outputs, int df, int dg, Bootstrapper &bootstrapper, GPT2_CKKS_ARGS);
    void compute_gelu_p(seal::Ciphertext &input, seal::Ciphertext &output, GPT2_CKKS_ARGS);
    void compute_gelu_q(seal::Ciphertext &input, seal::Ciphertext &output, GPT2_CKKS_ARGS);
    void compute_gelu(seal::Ciphertext &inputs, seal::Ciphertext &outputs, Bootstrapper &bootstrapper, GPT2_CKKS_ARGS);
    void compute_exp(seal::Ciphertext &input, seal::Ciphertext &output, int r, GPT2_CKKS_ARGS);
    void compute_softmax(seal::Ciphertext &input, int r, Bootstrapper &bootstrapper, GPT2_CKKS_ARGS);
    void compute_smax(seal::Ciphertext &input, int r, int gamma, GPT2_CKKS_ARGS);

    // ---- iterative approximations (IterApprox.cpp)
    void compute_inverse(seal::Ciphertext &input, seal::Ciphertext &output, int iters, GPT2_CKKS_ARGS);
    void taylor_expand(seal::Ciphertext &input, seal::Ciphertext &output, int iters, double guess, GPT2_CKKS_ARGS);
    void compute_inv_sqrt(seal::Ciphertext &input, seal::Ciphertext &output, int iters, double guess, GPT2_CKKS_ARGS);
    void compute_layernorm(seal::Ciphertext &input, seal::Ciphertext &output, std::vector<double> gamma, std::vector<double> beta,
                           int row_size, GPT2_CKKS_ARGS);

    // ---- folds (Fold.cpp)
    void quickSum(seal::Ciphertext &input, seal::Ciphertext &output, int n, GPT2_CKKS_ARGS);
    void computeMax(seal::Ciphertext &input1, seal::Ciphertext &input2, seal::Ciphertext &output, Bootstrapper &bootstrapper,
                    GPT2_CKKS_ARGS);
    void quickMax(seal::Ciphertext &input, seal::Ciphertext &output, int n, Bootstrapper &bootstrapper, GPT2_CKKS_ARGS);

    // ---- KV-cache augmentation (optimize.cpp)
    void augment_value_row(vc &A, vc &cached_val, int padded_row_size, int idx, GPT2_CKKS_ARGS);
    void augment_value_col(vc &A, vc &cached_val, int padded_row_size, int idx, GPT2_CKKS_ARGS);
} // namespace gpt2
