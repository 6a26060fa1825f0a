// cnn/cnn_seal.h - CNN operators on multiplexed-packed ciphertexts (Lee et al., "Low-complexity deep convolutional
// neural networks on fully homomorphic encryption using multiplexed parallel convolutions").
//
// Restates the reference's cnn_ckks/cpu-ckks/single-key/cnn/cnn_seal.{h,cpp} with the same class and function names,
// argument order and operation order: TensorCipher (cnn_seal.h:20-46, .cpp:3-100), multiplexed_parallel_convolution_seal
// (:284-530), multiplexed_parallel_batch_norm_seal (:531-576), ReLU_seal (:577-592), cnn_add_seal (:593-609),
// multiplexed_parallel_downsampling_seal (:610-679), averagepooling_seal_scale (:680-746),
// matrix_multiplication_seal (:747-787), memory_save_rotate (:788-809).  The *_print wrappers of the reference
// (timing + decrypt_and_print) are replaced by the trace hook of cnn/infer_seal.h.
//
// Slot layout of a tensor (k gap, h x w image, c channels, t = ceil(c / k^2) channel groups, p copies, n = 2^logn
// slots): value (channel ch, row y, col x) sits at slot  k^2 h w * (ch / k^2) + k w * (k y + (ch % k^2) / k) +
// k x + ch % k, repeated p times with period n / p.
#pragma once
#include "ckks_bootstrapping/Bootstrapper.h"
#include "comp/SEALcomp.h"
#include "seal/seal.h"
#include <vector>

class TensorCipher
{
private:
    int k_ = 0; // gap
    int h_ = 0; // height
    int w_ = 0; // width
    int c_ = 0; // number of channels
    int t_ = 0; // ceil(c / k^2)
    int p_ = 0; // number of copies, 2^floor(log2(n / (k^2 h w t)))
    int logn_ = 0;
    seal::Ciphertext cipher_;

public:
    TensorCipher() = default;
    // data: h*w*c reals in channel-major order (k must be 1), zero padded to 2^logn slots, encoded at 2^logp
    TensorCipher(int logn, int k, int h, int w, int c, int t, int p, std::vector<double> data, seal::Encryptor &encryptor,
                 seal::CKKSEncoder &encoder, int logp);
    TensorCipher(int logn, int k, int h, int w, int c, int t, int p, seal::Ciphertext cipher);
    int k() const { return k_; }
    int h() const { return h_; }
    int w() const { return w_; }
    int c() const { return c_; }
    int t() const { return t_; }
    int p() const { return p_; }
    int logn() const { return logn_; }
    seal::Ciphertext cipher() const { return cipher_; }
    const seal::Ciphertext &cipher_ref() const { return cipher_; }
    void set_ciphertext(seal::Ciphertext cipher) { cipher_ = std::move(cipher); }
};

// Everything of a convolution that depends on the weights and the input tensor's packing only: the fh*fw*q shifted
// weight vectors and the co select-and-scale masks (cnn_seal.cpp:329-399 of the reference builds them inside every
// call).  Built once per layer, a plan also names those vectors for the engine's plaintext cache (common/cached.h).
struct ConvPlan
{
    int ki, hi, wi, ci, ti, pi, logn; // input packing
    int co, st, fh, fw;
    int ko, ho, wo, to, po;
    long q;
    std::vector<std::vector<double>> tap_weights;    // [(i1 * fw + i2) * q + g][slot]
    std::vector<std::vector<double>> select_one_vec; // [output channel][slot]
};
ConvPlan build_conv_plan(const TensorCipher &cnn_in, int co, int st, int fh, int fw, const std::vector<double> &data,
                         const std::vector<double> &running_var, const std::vector<double> &constant_weight, double epsilon);
// named: let the engine keep the plan's encoded vectors resident (the plan must then outlive the evaluator's use of
// it, or evaluator.forget_cached(&plan) must be called first)
void multiplexed_parallel_convolution_planned(const TensorCipher &cnn_in, TensorCipher &cnn_out, const ConvPlan &plan,
                                              seal::CKKSEncoder &encoder, seal::Encryptor &encryptor,
                                              seal::Evaluator &evaluator, seal::GaloisKeys &gal_keys, bool end = false,
                                              bool named = true);

void multiplexed_parallel_convolution_seal(const TensorCipher &cnn_in, TensorCipher &cnn_out, int co, int st, int fh, int fw,
                                           const std::vector<double> &data, std::vector<double> running_var,
                                           std::vector<double> constant_weight, double epsilon, seal::CKKSEncoder &encoder,
                                           seal::Encryptor &encryptor, seal::Evaluator &evaluator, seal::GaloisKeys &gal_keys,
                                           std::vector<seal::Ciphertext> &cipher_pool, bool end = false);
void multiplexed_parallel_batch_norm_named(const TensorCipher &cnn_in, TensorCipher &cnn_out, const std::vector<double> &bias,
                                           const std::vector<double> &running_mean, const std::vector<double> &running_var,
                                           const std::vector<double> &weight, double epsilon, seal::CKKSEncoder &encoder,
                                           seal::Encryptor &encryptor, seal::Evaluator &evaluator, double B, const void *owner,
                                           std::uint64_t index);
void multiplexed_parallel_batch_norm_seal(const TensorCipher &cnn_in, TensorCipher &cnn_out, std::vector<double> bias,
                                          std::vector<double> running_mean, std::vector<double> running_var,
                                          std::vector<double> weight, double epsilon, seal::CKKSEncoder &encoder,
                                          seal::Encryptor &encryptor, seal::Evaluator &evaluator, double B, bool end = false);
void ReLU_seal(const TensorCipher &cnn_in, TensorCipher &cnn_out, long comp_no, std::vector<int> deg, long alpha,
               std::vector<minicomp::Tree> &tree, double scaled_val, long scalingfactor, seal::Encryptor &encryptor,
               seal::Evaluator &evaluator, seal::Decryptor &decryptor, seal::CKKSEncoder &encoder, seal::PublicKey &public_key,
               seal::SecretKey &secret_key, seal::RelinKeys &relin_keys, double scale = 1.0);
void cnn_add_seal(const TensorCipher &cnn1, const TensorCipher &cnn2, TensorCipher &destination, seal::Evaluator &evaluator);
void multiplexed_parallel_downsampling_seal(const TensorCipher &cnn_in, TensorCipher &cnn_out, seal::Evaluator &evaluator,
                                            seal::GaloisKeys &gal_keys);
void averagepooling_seal_scale(const TensorCipher &cnn_in, TensorCipher &cnn_out, seal::Evaluator &evaluator,
                               seal::GaloisKeys &gal_keys, double B);
void matrix_multiplication_seal(const TensorCipher &cnn_in, TensorCipher &cnn_out, std::vector<double> matrix,
                                std::vector<double> bias, int q, int r, seal::Evaluator &evaluator, seal::GaloisKeys &gal_keys,
                                const void *owner = nullptr); // owner: name under which the diagonals are cached (engine)
// rotate by `steps` (any sign) using only keys the network generates: steps 34..55 and 57..61 go through 33 first
void memory_save_rotate(const seal::Ciphertext &cipher_in, seal::Ciphertext &cipher_out, int steps, seal::Evaluator &evaluator,
                        seal::GaloisKeys &gal_keys);
