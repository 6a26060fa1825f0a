// comp/SEALfunc.h - homomorphic evaluation of one tree-decomposed odd Chebyshev polynomial.
//
// Restates (same names, arguments and operation order) the evaluation half of the reference's
// cnn_ckks/cpu-ckks/single-key/comp/SEALfunc.cpp: geneT0T1 (:33-50), evalT (:51-58),
// eval_polynomial_integrate, odd-baby branch (:59-193), coeff_number (:335-359), ShowFailure_ReLU (:360-375).
#pragma once
#include "comp/PolyUpdate.h"
#include "seal/seal.h"
#include <vector>

namespace seal
{
    using minicomp::Tree;

    // number of coefficients the decomposition of a degree-`deg` polynomial over `tree` stores
    long coeff_number(long deg, Tree &tree);

    // T_{m+n} = 2 T_m T_n - T_{m-n}
    void evalT(Evaluator &evaluator, PublicKey &public_key, SecretKey &secret_key, RelinKeys &relin_keys,
               Ciphertext &Tmplusn, const Ciphertext &Tm, const Ciphertext &Tn, const Ciphertext &Tmminusn);

    // T0 = fresh encryption of the all-ones vector at the ciphertext's scale, T1 = the input
    void geneT0T1(Encryptor &encryptor, Evaluator &evaluator, CKKSEncoder &encoder, PublicKey &public_key,
                  SecretKey &secret_key, RelinKeys &relin_keys, Ciphertext &T0, Ciphertext &T1, Ciphertext &cipher);

    void eval_polynomial_integrate(Encryptor &encryptor, Evaluator &evaluator, Decryptor &decryptor, CKKSEncoder &encoder,
                                   PublicKey &public_key, SecretKey &secret_key, RelinKeys &relin_keys, Ciphertext &res,
                                   Ciphertext &cipher, long deg, const std::vector<double> &decomp_coeff, Tree &tree);

    // number of slots i < n with |ReLU(x_i) - decrypted_i| > 2^-precision
    long ShowFailure_ReLU(Decryptor &decryptor, CKKSEncoder &encoder, Ciphertext &cipher, std::vector<double> &x,
                          long precision, long n);
} // namespace seal
