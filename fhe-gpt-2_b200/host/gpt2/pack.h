// gpt2/pack.h - packing of 128 x 768 activations into ciphertext slots (gpt2_ckks/.../gpt2/pack.{h,cpp}).
//
// "Row" (fold) format: matrix row i sits at slot offset i * 2 * round_to_2(cols) of the concatenated ciphertexts,
// followed by zero padding up to the next chunk, so that rotate-and-add folds never mix rows (16 rows of 768 per
// 32768-slot ciphertext, 8 ciphertexts for 128 rows).  "Tight" format: rows back to back, 768 slots each (3
// ciphertexts), the layout bootstrapping is applied to.
#pragma once
#include "gpt2/util.h"

namespace gpt2
{
#define GPT2_PACK_ARGS                                                                                                 \
    seal::CKKSEncoder &encoder, seal::Encryptor &encryptor, seal::Decryptor &decryptor, seal::Evaluator &evaluator,     \
        seal::GaloisKeys &gal_keys, seal::RelinKeys &relin_keys
    void pack_tight(vc &input, vc &output, GPT2_PACK_ARGS);
    void unpack_tight(vc &input, vc &output, GPT2_PACK_ARGS);
    void pack_from_row(vvec &input, vc &output, GPT2_PACK_ARGS);
    void expand_bias(std::vector<double> &input, seal::Ciphertext &output, GPT2_PACK_ARGS);
    void expand_bias_head_row(std::vector<double> &input, vc &output, int heads, GPT2_PACK_ARGS);
    void expand_bias_head_col(std::vector<double> &input, vc &output, int heads, int rows, int cols, GPT2_PACK_ARGS);
    std::vector<double> repeat(std::vector<double> &input, int times);
} // namespace gpt2
