// gpt2/MatrixMul.cpp - matrix products over packed ciphertexts: rotate, multiply, fold, mask, rotate, accumulate.
//
// Follows gpt2_ckks/gpt2-ckks/single-key/gpt2/MatrixMul.cpp of the reference.  Several of its operators are
// benchmark harnesses rather than finished linear algebra (the attention projections skip the final placement
// rotation "to save key memory", cipher_plain_128_128 / batch_matmul / qk_matmul_col multiply by one test vector,
// qk_matmul does not advance K between its 128 passes); they are restated as they are, with the sequence and count of
// evaluator calls of the reference, since those are what its timings and outputs come from.  The OpenMP fan-out of
// the reference is run in order (see util.h); loop-local ciphertexts that the reference shares between threads are
// private to an iteration here.
#include "gpt2/approx.h"
#include <cmath>

using namespace seal;
using std::vector;

namespace gpt2
{
    namespace
    {
        constexpr int kSlots = 32768;

        // the 16 one-hot masks at the head of each 2048-slot chunk (MatrixMul.cpp:253-260)
        vvec chunk_head_masks(std::size_t slots)
        {
            vvec masks;
            for (int k = 0; k < 16; k++)
            {
                vec v(slots, 0.0);
                v[(std::size_t)k * 2048] = 1.0;
                masks.push_back(std::move(v));
            }
            return masks;
        }
    } // namespace

    // MatrixMul.cpp:26-112.  A (rows ciphertexts, one per column vector) times B^T with B's columns stored twice in a
    // row so that a rotation by one acts cyclically: output i accumulates sum_j a_j * rot(b_j, i).  The accumulators
    // start as encryptions of zero multiplied by an encryption of one and are never relinearised, so they keep three
    // polynomials, as in the reference.
    void col_matrix_multiplication_seal(vector<TensorCipher> &left_inputs, vector<TensorCipher> &right_inputs,
                                        vector<TensorCipher> &outputs, vector<double>, int rows, int cols, Config &, CKKSEncoder &encoder,
                                        Encryptor &encryptor, Decryptor &, Evaluator &evaluator, GaloisKeys &gal_keys,
                                        RelinKeys &relin_keys)
    {
        const double input_scale = left_inputs[0].cipher().scale();
        const vector<double> zeros((std::size_t)cols, 0.0);
        vc accumulators;
        Plaintext plain, scaler;
        Ciphertext cipher, one;

        for (int i = 0; i < cols; i++)
        {
            encoder.encode(zeros, input_scale, plain);
            encryptor.encrypt(plain, cipher);
            encoder.encode(1, input_scale, scaler);
            encryptor.encrypt(scaler, one);
            evaluator.multiply_inplace(cipher, one);
            evaluator.rescale_to_next_inplace(cipher);
            accumulators.push_back(cipher);
        }

        for (int i = 0; i < cols; i++)
        {
            for (int j = 0; j < rows; j++)
            {
                evaluator.multiply(left_inputs[(std::size_t)j].cipher(), right_inputs[(std::size_t)j].cipher(), cipher);
                relinearize_then_rescale(evaluator, cipher, relin_keys);
                evaluator.mod_switch_to_inplace(accumulators[(std::size_t)i], cipher.parms_id());
                evaluator.add_inplace_reduced_error(accumulators[(std::size_t)i], cipher);
            }
            for (int j = 0; j < rows; j++)
            {
                cipher = right_inputs[(std::size_t)j].cipher();
                evaluator.rotate_vector_inplace(cipher, 1, gal_keys);
                right_inputs[(std::size_t)j].set_ciphertext(cipher);
            }
        }
        for (int i = 0; i < cols; i++)
        {
            TensorCipher tensor;
            tensor.set_ciphertext(accumulators[(std::size_t)i]);
            outputs.push_back(tensor);
        }
    }

    // MatrixMul.cpp:124-193.  A W^T for fold-format ciphertexts of A (rows of W_rows values) and of W (likewise): for
    // every pair of ciphertexts and every chunk rotation the Hadamard product is folded to one dot product per chunk,
    // and each dot product is masked out and moved to row-major position (row, col) of the pre-initialised outputs.
    void row_matrix_multiplication_seal(vc &left_inputs, vc &weights, Ciphertext bias, vc &outputs, int, int, int W_rows, int W_cols,
                                        CKKSEncoder &encoder, Encryptor &encryptor, Decryptor &decryptor, Evaluator &evaluator,
                                        GaloisKeys &gal_keys, RelinKeys &relin_keys)
    {
        const int slots = (int)encoder.slot_count();
        const int W_rows_rounded = round_to_2(W_rows), W_cols_rounded = round_to_2(W_cols);
        const int chunk_size = W_rows_rounded * 2, num_chunks = slots / chunk_size, out_chunk_size = W_cols_rounded * 2;
        Ciphertext rolled, folded, product, masked_out;

        for (std::size_t i = 0; i < left_inputs.size(); i++)
            for (std::size_t j = 0; j < weights.size(); j++)
                for (int rots = 0; rots < num_chunks; rots++)
                {
                    evaluator.rotate_vector(weights[j], rots * chunk_size, gal_keys, rolled);
                    evaluator.multiply_reduced_error(left_inputs[i], rolled, relin_keys, product);
                    evaluator.rescale_to_next_inplace(product);

                    evaluator.rotate_vector(product, slots - W_rows_rounded, gal_keys, rolled);
                    evaluator.add_inplace_reduced_error(product, rolled);
                    quickSum(product, folded, W_rows_rounded, encoder, encryptor, decryptor, evaluator, gal_keys, relin_keys);

                    for (int pos = 0; pos < num_chunks; ++pos)
                    {
                        const int row = (int)i * num_chunks + pos;
                        const int col = (int)j * num_chunks + ((rots + pos) % num_chunks);
                        mask_out(folded, masked_out, pos * chunk_size, 1, encoder, evaluator, relin_keys);

                        const int cipher_idx = (row * out_chunk_size) / slots;
                        const int cipher_chunk = ((row * out_chunk_size) % slots) / out_chunk_size;
                        const int desired_location = cipher_chunk * out_chunk_size + col;
                        const int shift_amt = desired_location - pos * chunk_size;
                        evaluator.rotate_vector_inplace(masked_out, -shift_amt, gal_keys);
                        evaluator.add_inplace_reduced_error(outputs[(std::size_t)cipher_idx], masked_out);
                    }
                }
        for (auto &out : outputs)
            evaluator.add_inplace_reduced_error(out, bias);
    }

    // MatrixMul.cpp:205-241: diagonal-packed inputs to one ciphertext per column
    void diagonal_to_row_matrix_seal(vector<TensorCipher> &inputs, vector<TensorCipher> &outputs, int rows, int cols, Config &,
                                     CKKSEncoder &encoder, Encryptor &, Decryptor &, Evaluator &evaluator, GaloisKeys &gal_keys,
                                     RelinKeys &)
    {
        const double scale = inputs[0].cipher().scale();
        Plaintext plain;
        Ciphertext cipher, piece;
        for (int i = 0; i < cols; i++)
        {
            vector<double> mask((std::size_t)cols, 0.0);
            mask[(std::size_t)i] = 1.0;
            encoder.encode(mask, scale, plain);
            evaluator.multiply_plain(inputs[0].cipher(), plain, cipher);
            evaluator.rescale_to_next_inplace(cipher);
            evaluator.rotate_vector_inplace(cipher, i, gal_keys);
            for (int j = 1; j < rows; j++)
            {
                evaluator.multiply_plain(inputs[(std::size_t)j].cipher(), plain, piece);
                evaluator.rescale_to_next_inplace(piece);
                evaluator.rotate_vector_inplace(piece, -(j - i), gal_keys);
                evaluator.add_inplace(cipher, piece);
            }
            outputs.push_back(TensorCipher(cipher));
        }
    }

    namespace
    {
        // Shared body of attn_proj_row_seal / attn_proj_col_seal (MatrixMul.cpp:243-346, 358-466).  Per pair of an
        // activation and a weight ciphertext: 16 working ciphertexts, each the weight ciphertext masked to the head of
        // chunk 0, rotated right by 1024 and fold-summed over 1024 slots (the reference indexes masks[0] for all 16
        // and never multiplies by the activations); then 16 x 16 masked pieces are accumulated into the per-head
        // outputs.  The placement rotation that would move a piece to (row, head column) is applied with step 0 -
        // "TEMPORARY MEMORY SAVING FEATURE" in the reference - so it is computed and not used here either.
        void attention_projection(bool column_layout, vc &left_inputs, vc &weights, const Ciphertext &bias, vc &outputs, int A_rows,
                                  int W_cols, CKKSEncoder &encoder, Encryptor &encryptor, Decryptor &decryptor, Evaluator &evaluator,
                                  GaloisKeys &gal_keys, RelinKeys &relin_keys)
        {
            const vvec masks = chunk_head_masks(encoder.slot_count());
            vc working(16);
            Ciphertext piece;
            for (std::size_t i = 0; i < left_inputs.size(); i++)
                for (std::size_t j = 0; j < weights.size(); j++)
                {
                    for (int k = 0; k < 16; k++)
                    {
                        evaluator.multiply_vector_reduced_error(weights[j], masks[0], working[(std::size_t)k]);
                        evaluator.rescale_to_next_inplace(working[(std::size_t)k]);
                        evaluator.rotate_vector_inplace(working[(std::size_t)k], -1024, gal_keys);
                        quickSum(working[(std::size_t)k], working[(std::size_t)k], 1024, encoder, encryptor, decryptor, evaluator,
                                 gal_keys, relin_keys);
                    }
                    for (int rots = 0; rots < 16; rots++)
                        for (int pos = 0; pos < 16; pos++)
                        {
                            const int row = (int)i * 16 + pos, col = (int)j * 16 + ((rots + pos) % 16);
                            const int abs_pos = row * (column_layout ? 768 : W_cols) + col;
                            const int head_col = abs_pos % 64, head = (abs_pos / 64) % 12;

                            evaluator.multiply_vector_reduced_error(working[(std::size_t)rots], masks[(std::size_t)pos], piece);
                            evaluator.rescale_to_next_inplace(piece);

                            const int desired_location = row * A_rows + head_col;
                            const int shift_amt = desired_location - pos * 2048;
                            (void)shift_amt;
                            evaluator.rotate_vector_inplace(piece, 0, gal_keys);
                            evaluator.add_inplace_reduced_error(outputs[(std::size_t)head], piece);
                        }
                }
            for (auto &out : outputs)
                evaluator.add_inplace_reduced_error(out, bias);
        }
    } // namespace

    void attn_proj_row_seal(vc &left_inputs, vc &weights, Ciphertext bias, vc &outputs, int A_rows, int, int, int W_cols, KeyGenerator &,
                            CKKSEncoder &encoder, Encryptor &encryptor, Decryptor &decryptor, Evaluator &evaluator, GaloisKeys &gal_keys,
                            RelinKeys &relin_keys)
    {
        attention_projection(false, left_inputs, weights, bias, outputs, A_rows, W_cols, encoder, encryptor, decryptor, evaluator,
                             gal_keys, relin_keys);
    }

    void attn_proj_col_seal(vc &left_inputs, vc &weights, Ciphertext bias, vc &outputs, int A_rows, int, int, int W_cols, KeyGenerator &,
                            CKKSEncoder &encoder, Encryptor &encryptor, Decryptor &decryptor, Evaluator &evaluator, GaloisKeys &gal_keys,
                            RelinKeys &relin_keys)
    {
        attention_projection(true, left_inputs, weights, bias, outputs, A_rows, W_cols, encoder, encryptor, decryptor, evaluator,
                             gal_keys, relin_keys);
    }

    // MatrixMul.cpp:478-521.  Per head: K is doubled (K + rot(K, 16384)); 128 passes of Q * K, a rotation by
    // 32768 - 64, a 64-slot fold, and 128 single-slot pieces moved to row * 256 + column with a Galois key generated
    // for that one shift (surefire_rotate).  K is not rotated between passes in the reference, so every pass
    // contributes the same products to different destinations; kept.
    void qk_matmul(vc &Q, vc &K, vc &outputs, int, int, int, int, KeyGenerator &keygen, CKKSEncoder &encoder, Encryptor &encryptor,
                   Decryptor &decryptor, Evaluator &evaluator, GaloisKeys &gal_keys, RelinKeys &relin_keys)
    {
        Ciphertext rolled;
        for (std::size_t i = 0; i < Q.size(); i++)
        {
            evaluator.rotate_vector(K[i], 16384, gal_keys, rolled);
            evaluator.add_inplace_reduced_error(K[i], rolled);
            for (int rots = 0; rots < 128; rots++)
            {
                Ciphertext product, shifted, folded, masked_out;
                evaluator.multiply_reduced_error(Q[i], K[i], relin_keys, product);
                evaluator.rescale_to_next_inplace(product);
                evaluator.rotate_vector(product, kSlots - 64, gal_keys, shifted);
                quickSum(shifted, folded, 64, encoder, encryptor, decryptor, evaluator, gal_keys, relin_keys);
                for (int pos = 0; pos < 128; pos++)
                {
                    const int row = (int)i * 128 + pos, col = (int)i * 128 + ((rots + pos) % 128);
                    const int abs_pos = row * 128 + col, head_col = abs_pos % 128;
                    mask_out(folded, masked_out, pos * 128, 1, encoder, evaluator, relin_keys);
                    const int desired_location = row * 256 + head_col;
                    const int shift_amt = desired_location - pos * 128;
                    surefire_rotate(masked_out, shift_amt, keygen, evaluator);
                    evaluator.add_inplace_reduced_error(outputs[i], masked_out);
                }
            }
        }
    }

    // MatrixMul.cpp:533-582.  Per head: 64 passes of rot(rot(V, 16384), 256 rots) * S, a rotation by 32768 - 128, a
    // 128-slot fold, and 128 single-slot pieces moved to (row % 16) * 2048 + head * 64 + column of outputs[row] (the
    // reference indexes outputs by the row itself, so 128 pre-initialised outputs are needed).
    void sv_matmul(vc &S, vc &V, vc &outputs, int, int, int, int, KeyGenerator &keygen, CKKSEncoder &encoder, Encryptor &encryptor,
                   Decryptor &decryptor, Evaluator &evaluator, GaloisKeys &gal_keys, RelinKeys &relin_keys)
    {
        Ciphertext cipher, rolled, folded, masked_out;
        for (std::size_t i = 0; i < S.size(); i++)
            for (int rots = 0; rots < 64; rots++)
            {
                evaluator.rotate_vector(V[i], kSlots - 16384, gal_keys, cipher);
                evaluator.rotate_vector_inplace(cipher, rots * 256, gal_keys);
                evaluator.multiply_inplace_reduced_error(cipher, S[i], relin_keys);
                evaluator.rescale_to_next_inplace(cipher);
                evaluator.rotate_vector(cipher, kSlots - 128, gal_keys, rolled);
                quickSum(rolled, folded, 128, encoder, encryptor, decryptor, evaluator, gal_keys, relin_keys);
                for (int pos = 0; pos < 128; pos++)
                {
                    const int row = pos, col = (rots + pos) % 64;
                    const int chunk_offset = (int)i * 64 + col, cipher_idx = row;
                    const int desired_location = (row % 16) * 2048 + chunk_offset;
                    mask_out(folded, masked_out, pos * 256, 1, encoder, evaluator, relin_keys);
                    const int shift_amt = desired_location - pos * 256;
                    surefire_rotate(masked_out, shift_amt, keygen, evaluator);
                    evaluator.add_inplace_reduced_error(outputs[(std::size_t)cipher_idx], masked_out);
                }
            }
    }

    // MatrixMul.cpp:584-626: the cost model of a 128 x 128 ciphertext-plaintext block product - 128 passes of four
    // masked products, two 128-slot folds and a two-term accumulation, all with weights["test"]; nothing is stored in
    // `outputs` (the reference keeps the per-thread results in locals).
    void cipher_plain_128_128(Ciphertext &left_input, std::unordered_map<std::string, vector<double>> &weights, Ciphertext, vc &, int,
                              int, int, int, KeyGenerator &, CKKSEncoder &encoder, Encryptor &encryptor, Decryptor &decryptor,
                              Evaluator &evaluator, GaloisKeys &gal_keys, RelinKeys &relin_keys)
    {
        const vector<double> &w = weights["test"];
        Ciphertext prod0_left, prod0_right, prod1_left, prod1_right, l_tmp, r_tmp, out0;
        for (int rots = 0; rots < 128; rots++)
        {
            for (Ciphertext *p : { &prod0_left, &prod0_right, &prod1_left, &prod1_right })
            {
                evaluator.multiply_vector_reduced_error(left_input, w, *p);
                evaluator.rescale_to_next_inplace(*p);
            }
            quickSum(prod0_left, prod0_left, 128, encoder, encryptor, decryptor, evaluator, gal_keys, relin_keys);
            quickSum(prod0_right, prod0_right, 128, encoder, encryptor, decryptor, evaluator, gal_keys, relin_keys);
            evaluator.multiply_vector_reduced_error(prod0_left, w, l_tmp);
            evaluator.multiply_vector_reduced_error(prod0_right, w, r_tmp);
            evaluator.add(l_tmp, r_tmp, out0);
            evaluator.rescale_to_next_inplace(out0);
        }
    }

    // MatrixMul.cpp:628-648: 64 outputs, each the sum of 128 ciphertext-plaintext products with weights["test"]
    void batch_matmul(vc &left_inputs, std::unordered_map<std::string, vector<double>> &weights, Ciphertext, vc &outputs, int, int, int,
                      int, KeyGenerator &, CKKSEncoder &, Encryptor &, Decryptor &, Evaluator &evaluator, GaloisKeys &, RelinKeys &)
    {
        const vector<double> &w = weights["test"];
        vc products(128);
        for (int i = 0; i < 64; i++)
        {
            for (int j = 0; j < 128; j++)
            {
                evaluator.multiply_vector_reduced_error(left_inputs[(std::size_t)j], w, products[(std::size_t)j]);
                evaluator.rescale_to_next_inplace(products[(std::size_t)j]);
            }
            evaluator.add_many(products, outputs[(std::size_t)i]);
        }
    }

    // MatrixMul.cpp:650-677: column-packed Q K^T - 128 outputs, output i = sum_j left_j * rot(right_j, i)
    void qk_matmul_col(vc &left_input, vc &right_input, std::unordered_map<std::string, vector<double>> &, Ciphertext, vc &outputs, int,
                       int, int, int, KeyGenerator &, CKKSEncoder &, Encryptor &, Decryptor &, Evaluator &evaluator, GaloisKeys &gal_keys,
                       RelinKeys &relin_keys)
    {
        vc products(64);
        for (int i = 0; i < 128; i++)
        {
            for (int j = 0; j < 64; j++)
            {
                evaluator.multiply_reduced_error(left_input[(std::size_t)j], right_input[(std::size_t)j], relin_keys,
                                                 products[(std::size_t)j]);
                evaluator.rescale_to_next_inplace(products[(std::size_t)j]);
            }
            evaluator.add_many(products, outputs[(std::size_t)i]);
            for (int j = 0; j < 64; j++)
                evaluator.rotate_vector_inplace(right_input[(std::size_t)j], 1, gal_keys);
        }
    }

    // ---- optimize.cpp: KV-cache augmentation, in place
    // row layout: clear row `idx` of the fresh projection and add the cached rows
    void augment_value_row(vc &A, vc &cached_val, int padded_row_size, int idx, CKKSEncoder &encoder, Encryptor &, Decryptor &,
                           Evaluator &evaluator, GaloisKeys &, RelinKeys &)
    {
        vec mask(encoder.slot_count(), 1.0);
        std::fill_n(mask.begin() + (std::ptrdiff_t)idx * padded_row_size, padded_row_size, 0.0);
        Plaintext plain;
        encoder.encode(mask, encode_scale(), plain);
        for (std::size_t i = 0; i < A.size(); i++)
        {
            evaluator.multiply_plain_inplace(A[i], plain);
            evaluator.rescale_to_next_inplace(A[i]);
            evaluator.add_inplace_reduced_error(A[i], cached_val[i]);
        }
    }

    // column layout: clear column `idx` of the cache (not rescaled, as in the reference), rotate the fresh
    // projection into that column and add
    void augment_value_col(vc &A, vc &cached_val, int padded_row_size, int idx, CKKSEncoder &encoder, Encryptor &, Decryptor &,
                           Evaluator &evaluator, GaloisKeys &gal_keys, RelinKeys &)
    {
        vec mask(encoder.slot_count(), 1.0);
        for (int i = 0; i < padded_row_size / 2; i++)
            mask[(std::size_t)(i * padded_row_size + idx)] = 0.0;
        Plaintext plain;
        encoder.encode(mask, encode_scale(), plain);
        for (std::size_t i = 0; i < A.size(); i++)
        {
            evaluator.multiply_plain_inplace(cached_val[i], plain);
            evaluator.rotate_vector_inplace(A[i], idx, gal_keys);
            evaluator.add_inplace_reduced_error(A[i], cached_val[i]);
        }
    }
} // namespace gpt2
