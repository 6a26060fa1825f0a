// ckks_bootstrapping/ModularReducer.cpp - see ModularReducer.h.
#include "ckks_bootstrapping/ModularReducer.h"
#include <cmath>
#include <stdexcept>

using namespace seal;

namespace
{
    struct EvalModEntry
    {
        long K, deg, logw, r;
        long double arcsin_slope;
        std::vector<long double> cheb;
    };
    const std::vector<EvalModEntry> &evalmod_tables()
    {
        static const std::vector<EvalModEntry> t = {
#include "ckks_bootstrapping/evalmod_table.inc"
        };
        return t;
    }
} // namespace

ModularReducer::ModularReducer(long _boundary_K, double _log_width, long _deg, long _num_double_formula, long _inverse_deg,
                               SEALContext &_context, CKKSEncoder &_encoder, Encryptor &_encryptor, Evaluator &_evaluator,
                               RelinKeys &_relin_keys, Decryptor &_decryptor)
    : boundary_K(_boundary_K), log_width(_log_width), deg(_deg), num_double_formula(_num_double_formula),
      inverse_deg(_inverse_deg), context(_context), encoder(_encoder), encryptor(_encryptor), evaluator(_evaluator),
      relin_keys(_relin_keys), decryptor(_decryptor)
{
    inverse_log_width = -std::log2(std::sin(2 * M_PI * std::pow(2.0, -log_width)));
}

// cos(2t) = 2 cos(t)^2 - 1
void ModularReducer::double_angle_formula(Ciphertext &cipher)
{
    double_angle_formula_scaled(cipher, 1.0);
}

// (s^2 cos 2t) = 2 (s cos t)^2 - s^2
void ModularReducer::double_angle_formula_scaled(Ciphertext &cipher, double scale_coeff)
{
    evaluator.square_inplace(cipher);
    relinearize_then_rescale(evaluator, cipher, relin_keys);
    evaluator.double_inplace(cipher);
    evaluator.add_const(cipher, -scale_coeff, cipher);
}

void ModularReducer::generate_sin_cos_polynomial()
{
    for (const auto &e : evalmod_tables())
        if (e.K == boundary_K && e.deg == deg && (double)e.logw == log_width && e.r == num_double_formula)
        {
            sin_cos_polynomial.set_zero_polynomial(deg);
            sin_cos_polynomial.chebcoeff = e.cheb;
            sin_cos_polynomial.cheb_to_power();
            sin_cos_polynomial.generate_poly_heap();
            arcsin_slope_ = e.arcsin_slope;
            return;
        }
    throw std::invalid_argument("no precomputed EvalMod minimax polynomial for these (K, degree, log width, double-angle "
                                "count); add one with tools/gen_evalmod_table.py");
}

void ModularReducer::generate_inverse_sine_polynomial()
{
    if (inverse_deg != 1)
        throw std::invalid_argument("only inverse_deg == 1 (arcsine folded into the cosine) is supported");
    if (arcsin_slope_ == 0)
        throw std::logic_error("generate_sin_cos_polynomial() must be called first");
    long double s = arcsin_slope_;
    inverse_sin_polynomial.set_zero_polynomial(1);
    inverse_sin_polynomial.coeff[1] = inverse_sin_polynomial.chebcoeff[1] = s;
    // fold the arcsine slope c into the cosine: (c^(1/2^r) cos)^(2^r double angles) = c cos(2^r t)
    scale_inverse_coeff = static_cast<double>(s);
    for (long i = 0; i < num_double_formula; i++)
        scale_inverse_coeff = std::sqrt(scale_inverse_coeff);
    sin_cos_polynomial.constmul((long double)scale_inverse_coeff);
    sin_cos_polynomial.generate_poly_heap();
}

void ModularReducer::modular_reduction(Ciphertext &rtn, Ciphertext &cipher)
{
    Ciphertext in = cipher, acc;
    sin_cos_polynomial.homomorphic_poly_evaluation(context, encoder, encryptor, evaluator, relin_keys, acc, in, decryptor);
    if (inverse_deg != 1)
        throw std::invalid_argument("only inverse_deg == 1 is supported");
    double curr_scale = scale_inverse_coeff;
    for (long i = 0; i < num_double_formula; i++)
    {
        curr_scale = curr_scale * curr_scale;
        double_angle_formula_scaled(acc, curr_scale);
    }
    rtn = acc;
}
