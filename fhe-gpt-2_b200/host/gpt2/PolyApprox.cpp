// gpt2/PolyApprox.cpp - polynomial approximations of sign, GELU, exp and softmax over ciphertext slots.
//
// Follows gpt2_ckks/gpt2-ckks/single-key/gpt2/PolyApprox.cpp of the reference: a logarithmic Chebyshev basis
// {T0, T1, T2, T3, T4, T8, ...}, degree-9 sign polynomials f and g evaluated by Paterson-Stockmeyer style division by
// T2, T4 and T8, the piecewise GELU (three sign evaluations select between 0, p(x), q(x) and x), exp as
// (1 + x / 2^r)^(2^r), and softmax as exp / sum(exp) with a Goldschmidt inverse.  The order of the evaluator calls is
// the reference's, so levels, scales and - given the same keys - ciphertext limbs agree with it.
#include "gpt2/approx.h"
#include <cassert>

using namespace seal;
using std::vector;

namespace gpt2
{
    namespace
    {
        // cipher <- 2 * cipher^2 - 1 (one level); the doubling is an addition, not a constant multiplication
        void double_angle(Ciphertext &cipher, Evaluator &evaluator, RelinKeys &relin_keys)
        {
            evaluator.square_inplace(cipher);
            relinearize_then_rescale(evaluator, cipher, relin_keys);
            evaluator.add_inplace(cipher, cipher);
            evaluator.add_const_inplace(cipher, -1.0);
        }

        // destination <- rescale(constant * source)
        void scaled(const Ciphertext &source, double constant, Ciphertext &destination, Evaluator &evaluator)
        {
            evaluator.multiply_const(source, constant, destination);
            evaluator.rescale_to_next_inplace(destination);
        }

        // a <- rescale(relinearize(a * b)) with the reduced-error level alignment
        void times(Ciphertext &a, const Ciphertext &b, Evaluator &evaluator, RelinKeys &relin_keys)
        {
            evaluator.multiply_inplace_reduced_error(a, b, relin_keys);
            evaluator.rescale_to_next_inplace(a);
        }

        struct SignCoefficients
        {
            double q1, r1, q2_t3, q2_x, q3;
        };
        // f of PolyApprox.cpp:115-129,171 and g of :207-230,270 (odd degree-9 polynomials in the Chebyshev basis)
        constexpr SignCoefficients kSignF{ -0.6767578125, 1.563049316, -0.02685546875, 0.1384277344, 0.002136230469 };
        constexpr SignCoefficients kSignG{ -1.121704102, 1.978370667, -0.6178588867, 0.403533935, 0.3557052612 };

        // output = (q1 x) T2 + r1 x  +  (q2_t3 T3 + q2_x x) T4  +  (q3 x) T8       (PolyApprox.cpp:104-197)
        void sign_polynomial(const SignCoefficients &c, Ciphertext &input, Ciphertext &output, CKKSEncoder &encoder,
                             Encryptor &encryptor, Decryptor &decryptor, Evaluator &evaluator, GaloisKeys &gal_keys,
                             RelinKeys &relin_keys)
        {
            vc basis;
            build_cheby_basis(input, basis, 4, encoder, encryptor, decryptor, evaluator, gal_keys, relin_keys);
            Ciphertext term, partial;

            output = input;
            scaled(input, c.q1, output, evaluator);
            times(output, basis[2], evaluator, relin_keys);
            scaled(input, c.r1, term, evaluator);
            evaluator.add_inplace_reduced_error(output, term);

            scaled(basis[3], c.q2_t3, term, evaluator);
            scaled(input, c.q2_x, partial, evaluator);
            evaluator.add_inplace_reduced_error(term, partial);
            times(term, basis[4], evaluator, relin_keys);
            evaluator.add_inplace_reduced_error(output, term);

            scaled(input, c.q3, term, evaluator);
            times(term, basis[5], evaluator, relin_keys);
            evaluator.add_inplace_reduced_error(output, term);
        }
    } // namespace

    // PolyApprox.cpp:15-101.  chebyBasis (assumed empty) receives T0 (a fresh encryption of ones at the top level; no
    // caller reads it), T1 = input, T2, T3 = 2x T2 - x, then T4, T8, ... by n - 2 further double-angle steps.
    void build_cheby_basis(Ciphertext &input, vc &chebyBasis, int n, CKKSEncoder &encoder, Encryptor &encryptor, Decryptor &,
                           Evaluator &evaluator, GaloisKeys &, RelinKeys &relin_keys)
    {
        assert(n > 0);
        Plaintext plain;
        Ciphertext power, t3, twice;
        encoder.encode(vector<double>(encoder.slot_count(), 1.0), input.scale(), plain);
        encryptor.encrypt(plain, power);
        chebyBasis.push_back(power);
        chebyBasis.push_back(input);

        power = input;
        double_angle(power, evaluator, relin_keys);
        chebyBasis.push_back(power);

        evaluator.add(input, input, twice);
        evaluator.multiply_reduced_error(twice, power, relin_keys, t3);
        evaluator.rescale_to_next_inplace(t3);
        scaled(input, -1.0, twice, evaluator);
        evaluator.add_inplace_reduced_error(t3, twice);
        chebyBasis.push_back(t3);

        for (int i = 0; i < n - 2; i++)
        {
            double_angle(power, evaluator, relin_keys);
            chebyBasis.push_back(power);
        }
    }

    void compute_sign_f(Ciphertext &input, Ciphertext &output, CKKSEncoder &encoder, Encryptor &encryptor, Decryptor &decryptor,
                        Evaluator &evaluator, GaloisKeys &gal_keys, RelinKeys &relin_keys)
    {
        sign_polynomial(kSignF, input, output, encoder, encryptor, decryptor, evaluator, gal_keys, relin_keys);
    }

    void compute_sign_g(Ciphertext &input, Ciphertext &output, CKKSEncoder &encoder, Encryptor &encryptor, Decryptor &decryptor,
                        Evaluator &evaluator, GaloisKeys &gal_keys, RelinKeys &relin_keys)
    {
        sign_polynomial(kSignG, input, output, encoder, encryptor, decryptor, evaluator, gal_keys, relin_keys);
    }

    // PolyApprox.cpp:281-309: dg/2 pairs of g, then df/2 pairs of f (the bootstrapper is not used)
    void sign_function(TensorCipher &inputs, TensorCipher &outputs, int df, int dg, Bootstrapper &, CKKSEncoder &encoder,
                       Encryptor &encryptor, Decryptor &decryptor, Evaluator &evaluator, GaloisKeys &gal_keys, RelinKeys &relin_keys)
    {
        Ciphertext current = inputs.cipher(), next;
        for (int i = 0; i < dg / 2; i++)
        {
            compute_sign_g(current, next, encoder, encryptor, decryptor, evaluator, gal_keys, relin_keys);
            compute_sign_g(next, current, encoder, encryptor, decryptor, evaluator, gal_keys, relin_keys);
        }
        for (int i = 0; i < df / 2; i++)
        {
            compute_sign_f(current, next, encoder, encryptor, decryptor, evaluator, gal_keys, relin_keys);
            compute_sign_f(next, current, encoder, encryptor, decryptor, evaluator, gal_keys, relin_keys);
        }
        outputs = TensorCipher(current);
    }

    // PolyApprox.cpp:311-346: p(x) = (q1 x + q0) T2 + (r1 x + r0), the GELU piece on [-4, -1.95]
    void compute_gelu_p(Ciphertext &input, Ciphertext &output, CKKSEncoder &encoder, Encryptor &encryptor, Decryptor &decryptor,
                        Evaluator &evaluator, GaloisKeys &gal_keys, RelinKeys &relin_keys)
    {
        vc basis;
        build_cheby_basis(input, basis, 2, encoder, encryptor, decryptor, evaluator, gal_keys, relin_keys);
        const double q_0 = -0.05745879353, q_1 = -0.005337069175, r_0 = -0.55528939, r_1 = -0.4187418723;
        Ciphertext quotient;

        scaled(input, q_1, quotient, evaluator);
        evaluator.add_const_inplace(quotient, q_0);
        times(quotient, basis[2], evaluator, relin_keys);

        scaled(input, r_1, output, evaluator);
        evaluator.add_const_inplace(output, r_0);
        evaluator.add_inplace_reduced_error(output, quotient);
    }

    // PolyApprox.cpp:347-409: q(x), the degree-6 GELU piece on [-1.95, 3]:
    // (qq1_1 x + qq1_0) T2 + (qr1_1 x + qr1_0) + (qq_2 x^2 + qq2_1 x + qq2_0) T4.  The last product is added without a
    // rescale of its own, as in the reference (the reduced-error addition reconciles the scales).
    void compute_gelu_q(Ciphertext &input, Ciphertext &output, CKKSEncoder &encoder, Encryptor &encryptor, Decryptor &decryptor,
                        Evaluator &evaluator, GaloisKeys &gal_keys, RelinKeys &relin_keys)
    {
        vc basis;
        build_cheby_basis(input, basis, 4, encoder, encryptor, decryptor, evaluator, gal_keys, relin_keys);
        const double qq1_0 = 0.1634058825, qq1_1 = -0.00324699876, qr1_0 = 0.1750485092, qr1_1 = 0.5027208006;
        const double qq2_0 = -0.004401064777, qq2_1 = 0.0002609111473, qq_2 = 0.0001533078376;
        Ciphertext quotient, high;

        scaled(input, qq1_1, quotient, evaluator);
        evaluator.add_const_inplace(quotient, qq1_0);
        times(quotient, basis[2], evaluator, relin_keys);

        scaled(input, qr1_1, output, evaluator);
        evaluator.add_const_inplace(output, qr1_0);
        evaluator.add_inplace_reduced_error(output, quotient);

        evaluator.square(input, quotient);
        relinearize_then_rescale(evaluator, quotient, relin_keys);
        evaluator.multiply_const_inplace(quotient, qq_2);
        evaluator.rescale_to_next_inplace(quotient);

        scaled(input, qq2_1, high, evaluator);
        evaluator.add_inplace_reduced_error(high, quotient);
        evaluator.add_const_inplace(high, qq2_0);
        evaluator.multiply_inplace_reduced_error(high, basis[4], relin_keys);
        evaluator.add_inplace_reduced_error(output, high);
    }

    // PolyApprox.cpp:419-483: with s_t = sign(x - t) / 2 for t in {3, -1.95, -4}:
    // gelu(x) = (s_-4 - s_-1.95) p(x) + (s_-1.95 - s_3) q(x) + (s_3 / 2) x
    void compute_gelu(Ciphertext &inputs, Ciphertext &outputs, Bootstrapper &bootstrapper, CKKSEncoder &encoder, Encryptor &encryptor,
                      Decryptor &decryptor, Evaluator &evaluator, GaloisKeys &gal_keys, RelinKeys &relin_keys)
    {
        auto half_sign = [&](double shift, Ciphertext &destination) {
            Ciphertext shifted;
            evaluator.add_const(inputs, shift, shifted);
            TensorCipher tc(shifted);
            sign_function(tc, tc, 2, 2, bootstrapper, encoder, encryptor, decryptor, evaluator, gal_keys, relin_keys);
            scaled(tc.cipher(), 0.5, destination, evaluator);
        };
        Ciphertext s0, s1, s2, b1, b2, b3, p, q;
        half_sign(-3.0, s2);
        half_sign(1.95, s1);
        half_sign(4.0, s0);

        evaluator.sub_reduced_error(s0, s1, b1);
        evaluator.sub_reduced_error(s1, s2, b2);
        scaled(s2, 0.5, b3, evaluator);

        compute_gelu_p(inputs, p, encoder, encryptor, decryptor, evaluator, gal_keys, relin_keys);
        compute_gelu_q(inputs, q, encoder, encryptor, decryptor, evaluator, gal_keys, relin_keys);

        evaluator.multiply_reduced_error(b1, p, relin_keys, outputs);
        evaluator.rescale_to_next_inplace(outputs);
        times(b2, q, evaluator, relin_keys);
        times(b3, inputs, evaluator, relin_keys);
        evaluator.add_inplace_reduced_error(outputs, b2);
        evaluator.add_inplace_reduced_error(outputs, b3);
    }

    // PolyApprox.cpp:491-512: (1 + x / 2^r)^(2^r), r + 1 levels
    void compute_exp(Ciphertext &input, Ciphertext &output, int r, CKKSEncoder &, Encryptor &, Decryptor &, Evaluator &evaluator,
                     GaloisKeys &, RelinKeys &relin_keys)
    {
        scaled(input, 1.0 / std::pow(2.0, r), output, evaluator);
        evaluator.add_const_inplace(output, 1);
        for (int i = 0; i < r; i++)
        {
            evaluator.square_inplace(output);
            relinearize_then_rescale(evaluator, output, relin_keys);
        }
    }

    namespace
    {
        // 128 rows of 256 slots: the first 128 of each row hold scores, the second 128 are the fold padding
        vector<double> padding_mask(std::size_t slots, double on_scores, double on_padding)
        {
            vector<double> mask(slots, on_scores);
            for (int i = 0; i < 128; i++)
                std::fill_n(mask.begin() + i * 256 + 128, 128, on_padding);
            return mask;
        }
    } // namespace

    // PolyApprox.cpp:514-575.  Row-wise softmax of a 128 x 128 score matrix in fold format (256-slot chunks), in
    // place: subtract the row maximum (quickMax), exponentiate, clear the padding, bootstrap, fold-sum, Goldschmidt
    // inverse, multiply.  The literal `r` of the reference's exp call (6) is kept; the parameter is unused there too.
    void compute_softmax(Ciphertext &input, int, Bootstrapper &bootstrapper, CKKSEncoder &encoder, Encryptor &encryptor,
                         Decryptor &decryptor, Evaluator &evaluator, GaloisKeys &gal_keys, RelinKeys &relin_keys)
    {
        const vector<double> zeros_mask = padding_mask(encoder.slot_count(), 1.0, 0.0);
        Ciphertext rolled, maxes, exps, summed, inverses;

        evaluator.rotate_vector(input, 32640, gal_keys, rolled);
        evaluator.add_inplace_reduced_error(input, rolled);
        quickMax(input, maxes, 128, bootstrapper, encoder, encryptor, decryptor, evaluator, gal_keys, relin_keys);
        evaluator.sub_inplace_reduced_error(input, maxes);

        compute_exp(input, exps, 6, encoder, encryptor, decryptor, evaluator, gal_keys, relin_keys);
        evaluator.multiply_vector_inplace_reduced_error(exps, zeros_mask);
        evaluator.rescale_to_next_inplace(exps);

        bootstrap(exps, rolled, bootstrapper, evaluator);

        evaluator.rotate_vector_inplace(rolled, -128, gal_keys);
        evaluator.add_inplace_reduced_error(rolled, exps);
        quickSum(rolled, summed, 128, encoder, encryptor, decryptor, evaluator, gal_keys, relin_keys);
        decrypt_and_print_and_max_round(summed, decryptor, encoder, 1.0, 0);

        compute_inverse(summed, inverses, 4, encoder, encryptor, decryptor, evaluator, gal_keys, relin_keys);
        evaluator.multiply_reduced_error(exps, inverses, relin_keys, input);
        evaluator.rescale_to_next_inplace(input);
    }

    // PolyApprox.cpp:577-634.  The bootstrap-free variant: instead of the row maximum a constant gamma is subtracted
    // (the reference adds -gamma on the padding half of each chunk, PolyApprox.cpp:592-597, and takes gamma as int).
    void compute_smax(Ciphertext &input, int, int gamma, CKKSEncoder &encoder, Encryptor &encryptor, Decryptor &decryptor,
                      Evaluator &evaluator, GaloisKeys &gal_keys, RelinKeys &relin_keys)
    {
        const vector<double> zeros_mask = padding_mask(encoder.slot_count(), 1.0, 0.0);
        const vector<double> gamma_mask = padding_mask(encoder.slot_count(), 0.0, -(double)gamma);
        Plaintext plain_gamma;
        Ciphertext rolled, exps, summed, inverses;

        encoder.encode(gamma_mask, input.scale(), plain_gamma);
        evaluator.mod_switch_to_inplace(plain_gamma, input.parms_id());
        evaluator.add_plain_inplace(input, plain_gamma);

        compute_exp(input, exps, 6, encoder, encryptor, decryptor, evaluator, gal_keys, relin_keys);
        evaluator.multiply_vector_inplace_reduced_error(exps, zeros_mask);
        evaluator.rescale_to_next_inplace(exps);

        evaluator.rotate_vector(exps, 32768 - 128, gal_keys, rolled);
        evaluator.add_inplace_reduced_error(rolled, exps);
        quickSum(rolled, summed, 128, encoder, encryptor, decryptor, evaluator, gal_keys, relin_keys);
        decrypt_and_print_and_max_round(summed, decryptor, encoder, 1.0, 0);

        compute_inverse(summed, inverses, 4, encoder, encryptor, decryptor, evaluator, gal_keys, relin_keys);
        evaluator.multiply_reduced_error(exps, inverses, relin_keys, input);
        evaluator.rescale_to_next_inplace(input);
    }
} // namespace gpt2
