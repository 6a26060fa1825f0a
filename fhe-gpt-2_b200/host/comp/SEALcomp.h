// comp/SEALcomp.h - ReLU(x) ~ x (1/2 + 1/2 sgn(x)) with sgn approximated by a composition of minimax polynomials.
//
// Restates cnn_ckks/cpu-ckks/single-key/comp/SEALcomp.cpp:3-60 (same name, arguments and operation order).
// The reference opens ../result/d<alpha>.txt on every call; here the alpha = 13 table is compiled in
// (comp/minimax_relu_alpha13.inc) and other tables can be registered with set_minimax_coefficients().
#pragma once
#include "comp/SEALfunc.h"
#include "comp/program.h"
#include <vector>

// tree-decomposed coefficients of all component polynomials, concatenated in file order
void set_minimax_coefficients(long alpha, const std::vector<double> &values);
const std::vector<double> &minimax_coefficients(long alpha);

void minimax_ReLU_seal(long comp_no, std::vector<int> deg, long alpha, std::vector<minicomp::Tree> &tree, double scaled_val,
                       long scalingfactor, seal::Encryptor &encryptor, seal::Evaluator &evaluator, seal::Decryptor &decryptor,
                       seal::CKKSEncoder &encoder, seal::PublicKey &public_key, seal::SecretKey &secret_key,
                       seal::RelinKeys &relin_keys, seal::Ciphertext &cipher_in, seal::Ciphertext &cipher_res);
