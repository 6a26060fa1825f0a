// ckks_bootstrapping/ModularReducer.h - EvalMod: homomorphic reduction modulo q0 by a scaled cosine and
// double-angle steps.
//
// Restates the reference's cnn_ckks/cpu-ckks/single-key/ckks_bootstrapping/ModularReducer.{h,cpp} (same class,
// members and call order).  The reference derives the cosine and arcsine minimax polynomials at start-up with a
// multi-interval Remez in 1000-bit NTL::RR (common/Remez.cpp, RemezCos.h, RemezArcsin.h); here they come from
// evalmod_table.inc, produced offline by tools/gen_evalmod_table.py from the same two minimax problems.
#pragma once
#include "common/Polynomial.h"
#include "seal/seal.h"

class ModularReducer
{
public:
    long boundary_K;
    double log_width;
    long deg;
    long num_double_formula;

    double inverse_log_width;
    long inverse_deg;

    double scale_inverse_coeff = 1.0;

    seal::SEALContext &context;
    seal::CKKSEncoder &encoder;
    seal::Encryptor &encryptor;
    seal::Evaluator &evaluator;
    seal::RelinKeys &relin_keys;
    seal::Decryptor &decryptor;

    boot::Polynomial sin_cos_polynomial;
    boot::Polynomial inverse_sin_polynomial;

    ModularReducer(long _boundary_K, double _log_width, long _deg, long _num_double_formula, long _inverse_deg,
                   seal::SEALContext &_context, seal::CKKSEncoder &_encoder, seal::Encryptor &_encryptor,
                   seal::Evaluator &_evaluator, seal::RelinKeys &_relin_keys, seal::Decryptor &_decryptor);

    void double_angle_formula(seal::Ciphertext &cipher);
    void double_angle_formula_scaled(seal::Ciphertext &cipher, double scale_coeff);
    void generate_sin_cos_polynomial();
    void generate_inverse_sine_polynomial();
    void modular_reduction(seal::Ciphertext &rtn, seal::Ciphertext &cipher);

private:
    long double arcsin_slope_ = 0; // degree-1 minimax of asin(y)/(2 pi) on |y| <= sin(2 pi 2^-log_width)
};
