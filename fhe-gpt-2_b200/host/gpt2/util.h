// gpt2/util.h - shared types and helpers of the GPT-2 operators.
//
// Restates gpt2_ckks/gpt2-ckks/single-key/gpt2/{tensor.h,util.h,util.cpp} of the reference: the TensorCipher and
// Config carriers, rotation helpers, the decrypt-and-re-encrypt stand-in for a bootstrap, zero-ciphertext
// initialisation, slot masking, row packing of plain matrices, the mod-switch-then-bootstrap wrapper and the
// generate-a-key-then-rotate helper.  Everything lives in namespace gpt2 because the CNN layer of this library has a
// TensorCipher of its own; a caller of the reference's functions adds `using namespace gpt2;`.
//
// The reference hard-codes 32768 slots (logN = 16).  Where the literal only means "all slots" this restatement asks
// the encoder, so the same code also runs on the small rings the CPU oracle tests use; the matrix layouts that depend
// on 128 x 768 activations keep the reference's literals.
#pragma once
#include "ckks_bootstrapping/Bootstrapper.h"
#include "seal/seal.h"
#include <cmath>
#include <string>
#include <unordered_map>
#include <vector>

namespace gpt2
{
    // util.h:21-26 of the reference
    constexpr int LOGP = 46, LOGQ = 49, BOOT_LEVEL = 14, TOTAL_LEVEL = 35, THREAD_NUM = 32;
    inline double encode_scale()
    {
        return std::pow(2.0, LOGP);
    }
    using vec = std::vector<double>;
    using vvec = std::vector<std::vector<double>>;
    using vc = std::vector<seal::Ciphertext>;

    // The reference fans its matrix operators out over THREAD_NUM = 32 OpenMP threads, one ciphertext operation per
    // core.  Here one ciphertext operation already fills the GPU, so the same loops run in order on the caller's
    // stream; nothing in the results depends on the interleaving (the shared accumulators only see modular additions).
    // progress lines of the reference's printf calls; off unless B200CKKS_GPT2_VERBOSE is set
    bool verbose();

    // tensor.h:23-52
    class TensorCipher
    {
    public:
        TensorCipher() = default;
        explicit TensorCipher(seal::Ciphertext cipher) : cipher_(std::move(cipher))
        {}
        TensorCipher(int logn, int k, int h, int w, int c, int t, int p, std::vector<double> data, seal::Encryptor &encryptor,
                     seal::CKKSEncoder &encoder, int logp);
        TensorCipher(int logn, int k, int h, int w, int c, int t, int p, seal::Ciphertext cipher)
            : k_(k), h_(h), w_(w), c_(c), t_(t), p_(p), logn_(logn), cipher_(std::move(cipher))
        {}
        int k() const { return k_; }
        int h() const { return h_; }
        int w() const { return w_; }
        int c() const { return c_; }
        int t() const { return t_; }
        int p() const { return p_; }
        int logn() const { return logn_; }
        seal::Ciphertext cipher() const { return cipher_; }
        void set_ciphertext(seal::Ciphertext cipher) { cipher_ = std::move(cipher); }
        void print_parms() const;

    private:
        int k_ = 0, h_ = 0, w_ = 0, c_ = 0, t_ = 0, p_ = 0, logn_ = 0;
        seal::Ciphertext cipher_;
    };

    // tensor.h:54-100, util.cpp:119-176 (defaults of the CNN parameter set; log_integer_part is always recomputed)
    struct Config
    {
        long boundary_K = 25, boot_deg = 59, scale_factor = 2, inverse_deg = 1, logN = 16, loge = 10, logn = 15, logn_1 = 14,
             logn_2 = 13, logn_3 = 12;
        int logp = 46, logq = 51, log_special_prime = 51, log_integer_part = 51 - 46 - 10 + 5, remaining_level = 16, boot_level = 14,
            total_level = 30;
        Config() = default;
        Config(long boundary_K, long boot_deg, long scale_factor, long inverse_deg, long logN, long loge, long logn, long logn_1,
               long logn_2, long logn_3, int logp, int logq, int log_special_prime, int log_integer_part, int remaining_level,
               int boot_level, int total_level);
    };

    // the 37-prime chain and rotation-key list of the INIT macro (util.h:37-75): {logq, logp x remaining_level,
    // logq x boot_level, log_special_prime} and powers of two + the listed steps + multiples of 2048
    std::vector<int> init_coeff_bit_vec(int logq = LOGQ, int logp = LOGP, int remaining_level = 21, int boot_level = BOOT_LEVEL,
                                        int log_special_prime = 60);
    std::vector<int> init_rotation_steps(int logN = 16);

    int round_to_2(double x);
    void rotate_inplace(seal::Ciphertext &cipher_in, int steps, seal::Evaluator &evaluator, seal::GaloisKeys &gal_keys);
    void rotate_vec(const seal::Ciphertext &cipher_in, seal::Ciphertext &cipher_out, int steps, seal::Evaluator &evaluator,
                    seal::GaloisKeys &gal_keys);
    void fakeBootstrap(seal::Ciphertext &input, seal::Ciphertext &output, seal::CKKSEncoder &encoder, seal::Encryptor &encryptor,
                       seal::Decryptor &decryptor, seal::Evaluator &evaluator, seal::GaloisKeys &gal_keys, seal::RelinKeys &relin_keys);
    void init_output(int num_ciphers, vc &output, seal::CKKSEncoder &encoder, seal::Encryptor &encryptor, seal::Decryptor &decryptor,
                     seal::Evaluator &evaluator, seal::GaloisKeys &gal_keys, seal::RelinKeys &relin_keys);
    void mask_out(seal::Ciphertext &cipher, seal::Ciphertext &out, int start, int length, seal::CKKSEncoder &encoder,
                  seal::Evaluator &evaluator, seal::RelinKeys &relin_keys);
    void pack_plain_row(vvec &v, int rows, int row_size, vvec &out);
    void add_galois_keys(std::vector<double> &gal_steps_vector);
    void init_bootstrap(Bootstrapper &bootstrapper, std::vector<int> &gal_steps_vector, int logn);
    void bootstrap(seal::Ciphertext &ctxt, seal::Ciphertext &rtn, Bootstrapper &bootstrapper, seal::Evaluator &evaluator);
    void surefire_rotate(seal::Ciphertext &cipher, int shift_amt, seal::KeyGenerator &keygen, seal::Evaluator &evaluator);
    // common/func.cpp:284-313; prints only when verbose()
    void decrypt_and_print_and_max_round(const seal::Ciphertext &cipher, seal::Decryptor &decryptor, seal::CKKSEncoder &encoder,
                                         double unit, long sparse_slots, std::size_t front = 5, std::size_t back = 5);
} // namespace gpt2
