// gpt2/Fold.cpp - rotate-and-combine folds over ciphertext slots (gpt2_ckks/gpt2-ckks/single-key/gpt2/Fold.cpp).
#include "gpt2/approx.h"
#include <cmath>

using namespace seal;

namespace gpt2
{
    // Fold.cpp:21-46.  log2(n) doubling steps: after the call slot i holds the sum of slots i .. i + n - 1.  One
    // plain add first, then log2(n) - 1 reduced-error adds (the loop bound is evaluated in floating point as in the
    // reference, so a non-power-of-two n runs ceil(log2 n) - 1 of them).  `input` and `output` may be one object.
    void quickSum(Ciphertext &input, Ciphertext &output, int n, CKKSEncoder &, Encryptor &, Decryptor &, Evaluator &evaluator,
                  GaloisKeys &gal_keys, RelinKeys &)
    {
        Ciphertext shifted;
        int stride = 1;
        evaluator.rotate_vector(input, stride, gal_keys, shifted);
        evaluator.add(input, shifted, output);
        stride *= 2;
        for (int i = 0; i < std::log2(n) - 1; i++)
        {
            evaluator.rotate_vector(output, stride, gal_keys, shifted);
            evaluator.add_inplace_reduced_error(output, shifted);
            stride *= 2;
        }
    }

    // Fold.cpp:49-86: max(a, b) = ((a + b) + (a - b) * sign((a - b) / 10)) / 2 with the composite sign g(g(f(f(.))))
    void computeMax(Ciphertext &input1, Ciphertext &input2, Ciphertext &output, Bootstrapper &bootstrapper, CKKSEncoder &encoder,
                    Encryptor &encryptor, Decryptor &decryptor, Evaluator &evaluator, GaloisKeys &gal_keys, RelinKeys &relin_keys)
    {
        Ciphertext difference, normalized;
        evaluator.sub(input1, input2, difference);
        evaluator.multiply_const(difference, 0.1, normalized);
        evaluator.rescale_to_next_inplace(normalized);

        TensorCipher sign_in(normalized), sign_out;
        sign_function(sign_in, sign_out, 2, 2, bootstrapper, encoder, encryptor, decryptor, evaluator, gal_keys, relin_keys);
        Ciphertext sign = sign_out.cipher();

        evaluator.multiply_inplace_reduced_error(difference, sign, relin_keys);
        evaluator.rescale_to_next_inplace(difference);
        evaluator.add_inplace_reduced_error(difference, input1);
        evaluator.add_inplace_reduced_error(difference, input2);
        evaluator.multiply_const(difference, 0.5, output);
        evaluator.rescale_to_next_inplace(output);
    }

    // Fold.cpp:89-107: log2(n) rounds of max(x, rot(x, 2^i)); a round that ends below 18 limbs is followed by a
    // bootstrap (there is not enough modulus left for another composite sign)
    void quickMax(Ciphertext &input, Ciphertext &output, int n, Bootstrapper &bootstrapper, CKKSEncoder &encoder, Encryptor &encryptor,
                  Decryptor &decryptor, Evaluator &evaluator, GaloisKeys &gal_keys, RelinKeys &relin_keys)
    {
        Ciphertext current = input, previous, shifted;
        int stride = 1;
        for (int i = 0; i < std::log2(n); i++)
        {
            previous = current;
            evaluator.rotate_vector(previous, stride, gal_keys, shifted);
            computeMax(previous, shifted, current, bootstrapper, encoder, encryptor, decryptor, evaluator, gal_keys, relin_keys);
            if (current.coeff_modulus_size() < 18)
                bootstrap(current, current, bootstrapper, evaluator);
            stride *= 2;
        }
        output = current;
    }
} // namespace gpt2
