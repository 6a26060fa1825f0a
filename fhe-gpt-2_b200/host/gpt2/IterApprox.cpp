// gpt2/IterApprox.cpp - iterative approximations: Goldschmidt division, Newton inverse square root, LayerNorm.
//
// Follows gpt2_ckks/gpt2-ckks/single-key/gpt2/IterApprox.cpp of the reference call for call.  Where the reference's
// C++ differs from the numpy model quoted in its own comments the C++ is what is restated (noted inline): the parity
// target is what the reference computes, not what it meant.
#include "gpt2/approx.h"
#include <cmath>

using namespace seal;
using std::vector;

namespace gpt2
{
    // IterApprox.cpp:16-57.  Goldschmidt: n_0 = d_0-normaliser = 0.001, d_0 = 0.001 x; f = 2 - d; n <- n f; d <- d f.
    // Converges to 1 / x for 0 < 0.001 x < 2.  Two levels per iteration on n, one on d.
    void compute_inverse(Ciphertext &input, Ciphertext &output, int iters, CKKSEncoder &encoder, Encryptor &encryptor, Decryptor &,
                         Evaluator &evaluator, GaloisKeys &, RelinKeys &relin_keys)
    {
        const double normalize_factor = 0.001;
        const std::size_t slots = encoder.slot_count();
        Ciphertext two, d, f;
        Plaintext plain;

        encoder.encode(vector<double>(slots, normalize_factor), encode_scale(), plain);
        evaluator.mod_switch_to_inplace(plain, input.parms_id());
        encryptor.encrypt(plain, output);

        encoder.encode(vector<double>(slots, 2.0), encode_scale(), plain);
        evaluator.mod_switch_to_inplace(plain, input.parms_id());
        encryptor.encrypt(plain, two);

        evaluator.multiply_const(input, normalize_factor, d);
        evaluator.rescale_to_next_inplace(d);

        for (int i = 0; i < iters; i++)
        {
            evaluator.sub_reduced_error(two, d, f);
            evaluator.multiply_inplace_reduced_error(output, f, relin_keys);
            evaluator.rescale_to_next_inplace(output);
            evaluator.multiply_inplace_reduced_error(d, f, relin_keys);
            evaluator.rescale_to_next_inplace(d);
        }
    }

    // IterApprox.cpp:70-122.  Starting value for the Newton iteration: sum over i = 1..3 of
    // c_i / i! * (guess^(p_i / i) x)^i with c = {-1/2, 3/4, -15/8}, p = {-3/2, -5/2, -7/2}.  (The numpy model in the
    // reference's comment expands around a - 1 with a constant term and powers of x - a; the C++ has neither.)
    void taylor_expand(Ciphertext &input, Ciphertext &output, int, double guess, CKKSEncoder &encoder, Encryptor &,
                       Decryptor &decryptor, Evaluator &evaluator, GaloisKeys &, RelinKeys &relin_keys)
    {
        const double coeffs[3] = { -0.5, -0.5 * -1.5, -2.5 * -1.5 * -0.5 };
        const double powers[3] = { -1.5, -2.5, -3.5 };
        int factorial = 1;
        Ciphertext sum, power, base;
        for (int i = 0; i < 3; i++)
        {
            const double coefficient = coeffs[i] * 1 / factorial;
            evaluator.multiply_const(input, std::pow(guess, powers[i] / (i + 1)), base);
            evaluator.rescale_to_next_inplace(base);
            power = base;
            for (int j = 0; j < i; j++)
            {
                evaluator.multiply_inplace_reduced_error(power, base, relin_keys);
                evaluator.rescale_to_next_inplace(power);
            }
            evaluator.multiply_const_inplace(power, coefficient);
            evaluator.rescale_to_next_inplace(power);
            if (i == 0)
                sum = power;
            else
                evaluator.add_inplace_reduced_error(sum, power);
            factorial *= (i + 2);
            decrypt_and_print_and_max_round(sum, decryptor, encoder, 1.0, 0, 5, 5);
        }
        output = sum;
    }

    // IterApprox.cpp:131-171.  y <- y (1.5 - 0.5 x y^2), `iters` times from the Taylor starting value; every iteration
    // ends with the reference's decrypt-and-re-encrypt refresh (fakeBootstrap), which needs the secret key.
    void compute_inv_sqrt(Ciphertext &input, Ciphertext &output, int iters, double guess, CKKSEncoder &encoder, Encryptor &encryptor,
                          Decryptor &decryptor, Evaluator &evaluator, GaloisKeys &gal_keys, RelinKeys &relin_keys)
    {
        Ciphertext square, minus_half_x;
        Plaintext plain;
        encoder.encode(guess, input.scale(), plain);
        encryptor.encrypt(plain, output);

        taylor_expand(input, output, 3, guess, encoder, encryptor, decryptor, evaluator, gal_keys, relin_keys);
        decrypt_and_print_and_max_round(output, decryptor, encoder, 1.0, 0);

        evaluator.multiply_const(input, -0.5, minus_half_x);
        evaluator.rescale_to_next_inplace(minus_half_x);

        for (int i = 0; i < iters; i++)
        {
            evaluator.square(output, square);
            relinearize_then_rescale(evaluator, square, relin_keys);

            evaluator.multiply_inplace_reduced_error(square, minus_half_x, relin_keys);
            evaluator.rescale_to_next_inplace(square);
            evaluator.add_const_inplace(square, 1.5);

            evaluator.multiply_inplace_reduced_error(output, square, relin_keys);
            evaluator.rescale_to_next_inplace(output);

            fakeBootstrap(output, output, encoder, encryptor, decryptor, evaluator, gal_keys, relin_keys);
            decrypt_and_print_and_max_round(output, decryptor, encoder, 1.0, 0, 5, 5);
        }
    }

    // IterApprox.cpp:173-252.  Row-wise LayerNorm of 16 rows of `row_size` values in 2 * round_to_2(row_size) slot
    // chunks: z = row_size x - sum(x); y = mask z^2; fold-sum of y; inverse square root (4 Newton steps from 323251);
    // then y z, times gamma sqrt(row_size), plus beta.  As in the reference the inverse square root is computed and
    // refreshed but the product that follows is y z, and beta is encoded untiled; the reference never assigns its
    // `output` parameter - here the final y is stored there so that the result can be read at all.
    void compute_layernorm(Ciphertext &input, Ciphertext &output, vector<double> gamma, vector<double> beta, int row_size,
                           CKKSEncoder &encoder, Encryptor &encryptor, Decryptor &decryptor, Evaluator &evaluator, GaloisKeys &gal_keys,
                           RelinKeys &relin_keys)
    {
        const int rounded_row_size = round_to_2(row_size);
        const std::size_t slots = encoder.slot_count();
        Plaintext plain_beta;
        Ciphertext rolled, folded, y, z, inv_sqrt;

        for (auto &g : gamma)
            g *= std::sqrt((double)row_size);

        vector<double> mask(slots, 0.0), mul_factor(slots, 0.0), beta_factor(slots, 0.0);
        for (int i = 0; i < 16; i++)
        {
            const std::ptrdiff_t at = (std::ptrdiff_t)i * rounded_row_size * 2;
            std::fill_n(mask.begin() + at, rounded_row_size, 1.0);
            std::copy(gamma.begin(), gamma.end(), mul_factor.begin() + at);
            std::copy(beta.begin(), beta.end(), beta_factor.begin() + at);
        }

        evaluator.rotate_vector(input, -rounded_row_size, gal_keys, rolled);
        evaluator.add_inplace_reduced_error(rolled, input);
        quickSum(rolled, folded, rounded_row_size, encoder, encryptor, decryptor, evaluator, gal_keys, relin_keys);

        evaluator.multiply_const(input, row_size, z);
        evaluator.rescale_to_next_inplace(z);
        decrypt_and_print_and_max_round(z, decryptor, encoder, 1.0, 0);
        evaluator.sub_inplace_reduced_error(z, folded);

        evaluator.square(z, y);
        relinearize_then_rescale(evaluator, y, relin_keys);
        decrypt_and_print_and_max_round(y, decryptor, encoder, 1.0, 0);

        evaluator.multiply_vector_inplace_reduced_error(y, mask);
        evaluator.rescale_to_next_inplace(y);

        evaluator.rotate_vector(y, (int)slots - rounded_row_size, gal_keys, rolled);
        evaluator.add_inplace_reduced_error(rolled, y);
        quickSum(rolled, folded, rounded_row_size, encoder, encryptor, decryptor, evaluator, gal_keys, relin_keys);

        compute_inv_sqrt(folded, inv_sqrt, 4, 323251, encoder, encryptor, decryptor, evaluator, gal_keys, relin_keys);
        fakeBootstrap(inv_sqrt, inv_sqrt, encoder, encryptor, decryptor, evaluator, gal_keys, relin_keys);

        evaluator.multiply_inplace_reduced_error(y, z, relin_keys);
        evaluator.rescale_to_next_inplace(y);

        evaluator.multiply_vector_inplace_reduced_error(y, mul_factor);
        evaluator.rescale_to_next_inplace(y);

        encoder.encode(beta, y.scale(), plain_beta);
        evaluator.mod_switch_to_inplace(plain_beta, y.parms_id());
        evaluator.add_plain_inplace(y, plain_beta);
        output = y;
    }
} // namespace gpt2
