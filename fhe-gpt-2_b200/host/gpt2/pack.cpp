// gpt2/pack.cpp - see pack.h.  Follows gpt2_ckks/gpt2-ckks/single-key/gpt2/pack.cpp of the reference operation by
// operation, including the places where the reference departs from its own numpy model (noted inline), because the
// parity target is what the reference computes.
#include "gpt2/pack.h"
#include <algorithm>

using namespace seal;
using std::vector;

namespace gpt2
{
    namespace
    {
        constexpr int kRow = 768;      // activation width
        constexpr int kChunk = 2048;   // 2 * round_to_2(768): one row of the fold format
        constexpr int kSlots = 32768;  // the layouts below are those of logN = 16
        constexpr int kRows = 128;
    } // namespace

    // fold format (8 ciphertexts, pre-initialised `output` of 3) -> tight format.  pack.cpp:9-57: the row at the
    // front of input[i] is masked, moved to its tight position and accumulated, then input[i] is advanced by one
    // chunk; a row that straddles two output ciphertexts is moved in two pieces.
    void pack_tight(vc &input, vc &output, CKKSEncoder &encoder, Encryptor &, Decryptor &decryptor, Evaluator &evaluator,
                    GaloisKeys &gal_keys, RelinKeys &relin_keys)
    {
        Ciphertext piece;
        int written = 0;
        for (int i = 0; i < 8; i++)
        {
            bool straddled = false;
            for (int j = 0; j < 16; j++)
            {
                if ((j == 15 && straddled) || written == kRows * kRow)
                    break;
                mask_out(input[(std::size_t)i], piece, 0, kRow, encoder, evaluator, relin_keys);
                rotate_inplace(piece, -(written % kSlots), evaluator, gal_keys);
                decrypt_and_print_and_max_round(piece, decryptor, encoder, 1.0, 0);
                evaluator.add_inplace_reduced_error(output[(std::size_t)(written / kSlots)], piece);
                written += kRow;
                evaluator.rotate_vector_inplace(input[(std::size_t)i], kChunk, gal_keys);

                const int leftover = kSlots - (written % kSlots);
                if (leftover < kRow)
                {
                    mask_out(input[(std::size_t)i], piece, 0, leftover, encoder, evaluator, relin_keys);
                    rotate_inplace(piece, -(written % kSlots), evaluator, gal_keys);
                    evaluator.add_inplace_reduced_error(output[(std::size_t)(written / kSlots)], piece);
                    written += leftover;

                    mask_out(input[(std::size_t)i], piece, leftover, kRow - leftover, encoder, evaluator, relin_keys);
                    rotate_inplace(piece, leftover, evaluator, gal_keys);
                    evaluator.add_inplace_reduced_error(output[(std::size_t)(written / kSlots)], piece);
                    written += kRow - leftover;

                    rotate_inplace(input[(std::size_t)i], kChunk, evaluator, gal_keys);
                    straddled = true;
                }
            }
        }
    }

    // tight format (3 ciphertexts) -> fold format (pre-initialised `output` of 8).  pack.cpp:99-141.  In the
    // straddling case the reference masks 2048 - leftover slots for the second piece (its numpy model in the comment
    // above the function masks 768 - leftover); the reference's length is kept.
    void unpack_tight(vc &input, vc &output, CKKSEncoder &encoder, Encryptor &, Decryptor &, Evaluator &evaluator, GaloisKeys &gal_keys,
                      RelinKeys &relin_keys)
    {
        Ciphertext piece, moved;
        int src = 0, dst = 0;
        auto move_piece = [&](int length) {
            mask_out(input[(std::size_t)(src / kSlots)], piece, src % kSlots, length, encoder, evaluator, relin_keys);
            evaluator.rotate_vector(piece, (src % kSlots) - (dst % kSlots), gal_keys, moved);
            evaluator.add_inplace_reduced_error(output[(std::size_t)(dst / kSlots)], moved);
        };
        while (src < kSlots * 3)
        {
            move_piece(kRow);
            src += kRow;
            dst += kChunk;
            const int leftover = kSlots - (src % kSlots);
            if (leftover < kRow)
            {
                move_piece(leftover);
                src += leftover;
                dst += leftover;
                move_piece(kChunk - leftover);
                src += kRow - leftover;
                dst += kChunk - leftover;
            }
        }
    }

    // plain matrix -> fold-format ciphertexts appended to `output` (pack.cpp:144-175)
    void pack_from_row(vvec &input, vc &output, CKKSEncoder &encoder, Encryptor &encryptor, Decryptor &, Evaluator &, GaloisKeys &,
                       RelinKeys &)
    {
        const int rows = (int)input.size(), cols = (int)input[0].size();
        const int slots = (int)encoder.slot_count();
        const int chunk = round_to_2(cols) * 2;
        const int num_ciphers = std::max(1, (rows * chunk) / slots);
        vvec packed((std::size_t)num_ciphers, vector<double>((std::size_t)slots, 0.0));
        pack_plain_row(input, rows, cols, packed);
        Plaintext plain;
        for (int i = 0; i < num_ciphers; i++)
        {
            Ciphertext cipher;
            encoder.encode(packed[(std::size_t)i], encode_scale(), plain);
            encryptor.encrypt(plain, cipher);
            output.push_back(std::move(cipher));
        }
    }

    vector<double> repeat(vector<double> &input, int times)
    {
        vector<double> result(input.size() * (std::size_t)times);
        for (int rep = 0; rep < times; rep++)
            std::copy(input.begin(), input.end(), result.begin() + (std::ptrdiff_t)((std::size_t)rep * input.size()));
        return result;
    }

    // pack.cpp:188-205.  The reference resizes the tile to (chunk - size) entries rather than to the chunk, so the
    // period of the tiling is chunk - size; kept as is.
    void expand_bias(vector<double> &input, Ciphertext &output, CKKSEncoder &encoder, Encryptor &encryptor, Decryptor &, Evaluator &,
                     GaloisKeys &, RelinKeys &)
    {
        const int size = (int)input.size(), chunk = round_to_2(size) * 2;
        vector<double> tile(input.begin(), input.end());
        tile.resize((std::size_t)(chunk - size));
        vector<double> tiled = repeat(tile, (int)encoder.slot_count() / chunk);
        Plaintext plain;
        encoder.encode(tiled, encode_scale(), plain);
        encryptor.encrypt(plain, output);
    }

    // pack.cpp:208-225: the reference tiles each head's bias into a local vector and discards it; `output` is not
    // touched there either.
    void expand_bias_head_row(vector<double> &, vc &, int, CKKSEncoder &, Encryptor &, Decryptor &, Evaluator &, GaloisKeys &,
                              RelinKeys &)
    {}

    // pack.cpp:228-247: for head i and column j, `rows` copies of bias[i * cols + j] are written at offset
    // j * rows * 2 of a slot vector that is never cleared, and the whole vector is added to output[i] each time.
    void expand_bias_head_col(vector<double> &input, vc &output, int heads, int rows, int cols, CKKSEncoder &encoder, Encryptor &,
                              Decryptor &, Evaluator &evaluator, GaloisKeys &, RelinKeys &)
    {
        vector<double> slots_vec(encoder.slot_count(), 0.0);
        Plaintext plain;
        for (int i = 0; i < heads; i++)
            for (int j = 0; j < cols; j++)
            {
                std::fill_n(slots_vec.begin() + (std::ptrdiff_t)j * rows * 2, rows, input[(std::size_t)(i * cols + j)]);
                encoder.encode(slots_vec, encode_scale(), plain);
                evaluator.add_plain_inplace(output[(std::size_t)i], plain);
            }
    }
} // namespace gpt2
