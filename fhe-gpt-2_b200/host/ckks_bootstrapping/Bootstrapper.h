// ckks_bootstrapping/Bootstrapper.h - CKKS bootstrapping with three-level CoeffToSlot / SlotToCoeff.
//
// Restates the live part of the reference's cnn_ckks/cpu-ckks/single-key/ckks_bootstrapping/Bootstrapper.{h,cpp}
// with the same class name, constructor, members and method names, so that its callers (cnn/infer_seal.cpp,
// run/run_bootstrapping.cpp, gpt2/util.cpp) read the same: key list (:82-177), LT coefficient generation
// (:512-592,1116-1909), BSGS linear transforms (:1952-2086), sfl/sflinv (:2376-2575), CoeffToSlot/SlotToCoeff
// (:2675-2749), ModRaise (:2894-2948) and the bootstrap_*_3 entry points (:3074-3262,3409-3431).  The ~20 dead
// variants (one-depth, hoisting, 2-level) are out of scope.
#pragma once
#include "ckks_bootstrapping/LinearTransform.h"
#include "ckks_bootstrapping/ModularReducer.h"
#include "seal/seal.h"
#include <complex>
#include <vector>

class Bootstrapper
{
public:
    typedef std::vector<std::vector<std::complex<double>>> Diagonals; // [diagonal][slot]

    long loge;
    long logn;
    long n;
    long logNh;
    long Nh;
    long L;

    double initial_scale = 0;
    double final_scale;

    long boundary_K;
    long sin_cos_deg;
    long scale_factor;
    long inverse_deg;

    seal::SEALContext &context;
    seal::KeyGenerator &keygen;
    seal::CKKSEncoder &encoder;
    seal::Encryptor &encryptor;
    seal::Decryptor &decryptor;
    seal::Evaluator &evaluator;
    seal::RelinKeys &relin_keys;
    seal::GaloisKeys &gal_keys;

    std::vector<long> slot_vec;
    long slot_index = 0;
    // per entry of slot_vec: the three merged matrices of each direction as lists of diagonals
    std::vector<Diagonals> fftcoeff1, fftcoeff2, fftcoeff3;          // SlotToCoeff
    std::vector<Diagonals> invfftcoeff1, invfftcoeff2, invfftcoeff3; // CoeffToSlot

    ModularReducer *mod_reducer;

    // Baby-step rotations of a BSGS transform rotate one ciphertext by many steps.  On the engine they can share one
    // decomposition (Evaluator::rotate_vector_hoisted): same decrypted values up to key-switching noise, different
    // limbs than the reference's one-by-one rotations.  Off = the reference's exact operation sequence.  Default: on,
    // unless $B200CKKS_NO_HOIST is set; always off on stock SEAL.
    bool hoisting;
    // With hoisting on and the engine in level-aware hybrid mode, the inner sums of a transform are double-hoisted
    // (Evaluator::bsgs_inner_sums_cached): the baby rotations stay in the extended basis of the key switch and every
    // giant step pays one division by the special modulus instead of one per baby rotation.  Default: on, unless
    // $B200CKKS_NO_DOUBLE_HOIST is set.
    bool double_hoisting;
    // fixed at construction (it decides the rotation keys): double hoisting available -> transforms use bsgs_width()
    // baby steps instead of the reference's balanced giantstep() split
    bool wide_babies = false;
    int bsgs_width(int M, int limbs) const;

    Bootstrapper(long _loge, long _logn, long _logNh, long _L, double _final_scale, long _boundary_K, long _sin_cos_deg,
                 long _scale_factor, long _inverse_deg, seal::SEALContext &_context, seal::KeyGenerator &_keygen,
                 seal::CKKSEncoder &_encoder, seal::Encryptor &_encryptor, seal::Decryptor &_decryptor,
                 seal::Evaluator &_evaluator, seal::RelinKeys &_relin_keys, seal::GaloisKeys &_gal_keys);
    ~Bootstrapper();
    Bootstrapper(const Bootstrapper &) = delete;
    Bootstrapper &operator=(const Bootstrapper &) = delete;

    // rotation steps the three-level transforms need
    void addLeftRotKeys_Linear_to_vector_3(std::vector<int> &gal_steps_vector);
    void addBootKeys_3(seal::GaloisKeys &gal_keys);
    void addBootKeys_3_other_keys(seal::GaloisKeys &gal_keys, std::vector<int> &other_keys);
    void change_logn(long new_logn);

    void genfftcoeff_3();
    void geninvfftcoeff_3();
    void generate_LT_coefficient_3();
    void prepare_mod_polynomial();

    void subsum(double scale, seal::Ciphertext &cipher);
    // cache_owner / cache_variant name the matrix for the engine's plaintext cache (common/cached.h): pass the
    // address of a Diagonals object that outlives the call, and a variant if its values were rescaled
    void bsgs_linear_transform(seal::Ciphertext &rtncipher, seal::Ciphertext &cipher, int totlen, int basicstep,
                               int coeff_logn, const Diagonals &fftcoeff, const void *cache_owner = nullptr,
                               std::uint64_t cache_variant = 0, bool rescale = false);
    void rotated_bsgs_linear_transform(seal::Ciphertext &rtncipher, seal::Ciphertext &cipher, int totlen, int basicstep,
                                       int coeff_logn, const Diagonals &fftcoeff, const void *cache_owner = nullptr,
                                       std::uint64_t cache_variant = 0, bool rescale = false);

    void sfl_3(seal::Ciphertext &rtncipher, seal::Ciphertext &cipher);
    void sfl_full_3(seal::Ciphertext &rtncipher, seal::Ciphertext &cipher);
    void sfl_half_3(seal::Ciphertext &rtncipher, seal::Ciphertext &cipher);
    void sfl_full_half_3(seal::Ciphertext &rtncipher, seal::Ciphertext &cipher);
    void sflinv_3(seal::Ciphertext &rtncipher, seal::Ciphertext &cipher);
    void sflinv_full_3(seal::Ciphertext &rtncipher, seal::Ciphertext &cipher);

    void coefftoslot_3(seal::Ciphertext &rtncipher, seal::Ciphertext &cipher);
    void slottocoeff_3(seal::Ciphertext &rtncipher, seal::Ciphertext &cipher);
    void slottocoeff_half_3(seal::Ciphertext &rtncipher, seal::Ciphertext &cipher);
    void coefftoslot_full_3(seal::Ciphertext &rtncipher1, seal::Ciphertext &rtncipher2, seal::Ciphertext &cipher);
    void slottocoeff_full_3(seal::Ciphertext &rtncipher, seal::Ciphertext &cipher1, seal::Ciphertext &cipher2);
    void slottocoeff_full_half_3(seal::Ciphertext &rtncipher, seal::Ciphertext &cipher1, seal::Ciphertext &cipher2);

    void modraise_inplace(seal::Ciphertext &cipher);

    void bootstrap_sparse_3(seal::Ciphertext &rtncipher, seal::Ciphertext &cipher);
    void bootstrap_full_3(seal::Ciphertext &rtncipher, seal::Ciphertext &cipher);
    void bootstrap_sparse_real_3(seal::Ciphertext &rtncipher, seal::Ciphertext &cipher);
    void bootstrap_full_real_3(seal::Ciphertext &rtncipher, seal::Ciphertext &cipher);

    void bootstrap_3(seal::Ciphertext &rtncipher, seal::Ciphertext &cipher);
    void bootstrap_inplace_3(seal::Ciphertext &cipher);
    void bootstrap_real_3(seal::Ciphertext &rtncipher, seal::Ciphertext &cipher);
    void bootstrap_inplace_real_3(seal::Ciphertext &cipher);

private:
    struct Split
    {
        int part[3], totlen[3], basicstep[3];
    };
    Split split_decode(long logn_) const; // SlotToCoeff grouping (largest group first)
    Split split_encode(long logn_) const; // CoeffToSlot grouping (smallest group first)
    void find_slot_index();
    void sfl_common(seal::Ciphertext &rtncipher, seal::Ciphertext &cipher, bool full, double last_divisor);
    void sparse_head(seal::Ciphertext &rtn, seal::Ciphertext &cipher);
    void sparse_tail(seal::Ciphertext &rtncipher, seal::Ciphertext &modrtn, bool half);
};
