// comp/SEALcomp.cpp - see SEALcomp.h.
#include "comp/SEALcomp.h"
#include <map>
#include <mutex>
#include <stdexcept>

using namespace seal;

namespace
{
    std::mutex g_mu;
    std::map<long, std::vector<double>> &tables()
    {
        static std::map<long, std::vector<double>> t = { { 13L,
                                                            {
#include "comp/minimax_relu_alpha13.inc"
                                                            } } };
        return t;
    }
} // namespace

void set_minimax_coefficients(long alpha, const std::vector<double> &values)
{
    std::lock_guard<std::mutex> g(g_mu);
    tables()[alpha] = values;
}

const std::vector<double> &minimax_coefficients(long alpha)
{
    std::lock_guard<std::mutex> g(g_mu);
    auto it = tables().find(alpha);
    if (it == tables().end())
        throw std::invalid_argument("no minimax composite coefficients registered for this alpha");
    return it->second;
}

void minimax_ReLU_seal(long comp_no, std::vector<int> deg, long alpha, std::vector<minicomp::Tree> &tree, double scaled_val,
                       long, Encryptor &encryptor, Evaluator &evaluator, Decryptor &decryptor, CKKSEncoder &encoder,
                       PublicKey &public_key, SecretKey &secret_key, RelinKeys &relin_keys, Ciphertext &cipher_in,
                       Ciphertext &cipher_res)
{
    if ((long)deg.size() != comp_no || (long)tree.size() != comp_no)
        throw std::invalid_argument("deg and tree must have comp_no entries");
    const std::vector<double> &table = minimax_coefficients(alpha);

    // component i maps [-1,1] into the input range of component i+1, which is scaled to [-2,2] (the last one to
    // [-scaled_val, scaled_val]); dividing the coefficients by that range keeps every intermediate in [-1,1].
    // The last component is halved so that the chain ends in sgn(x)/2.
    std::vector<std::vector<double>> coeff((std::size_t)comp_no);
    std::size_t pos = 0;
    for (long i = 0; i < comp_no; i++)
    {
        long count = coeff_number(deg[(std::size_t)i], tree[(std::size_t)i]);
        if (pos + (std::size_t)count > table.size())
            throw std::invalid_argument("minimax coefficient table is too short for these degrees");
        double next_range = (i + 1 < comp_no) ? ((i + 1 == comp_no - 1) ? scaled_val : 2.0) : 2.0 /* => * 0.5 */;
        for (long j = 0; j < count; j++)
            coeff[(std::size_t)i].push_back(table[pos + (std::size_t)j] / next_range);
        pos += (std::size_t)count;
    }

    Ciphertext x = cipher_in;
    for (long i = 0; i < comp_no; i++)
        eval_polynomial_integrate(encryptor, evaluator, decryptor, encoder, public_key, secret_key, relin_keys, x, x,
                                  deg[(std::size_t)i], coeff[(std::size_t)i], tree[(std::size_t)i]);

    // x (1 + sgn x) / 2 from sgn(x) / 2: add an encryption of 1/2, multiply by the input
    std::vector<double> half(cipher_in.poly_modulus_degree() / 2, 0.5);
    Plaintext plain_half;
    Ciphertext cipher_half, sum;
    if (encrypt_constants())
    {
        encoder.encode(half, x.scale(), plain_half);
        encryptor.encrypt(plain_half, cipher_half);
        evaluator.add_reduced_error(x, cipher_half, sum);
    }
    else
        evaluator.add_const(x, 0.5, sum);
#ifdef B200CKKS_FACADE
    if (merged_rescale())
    {
        evaluator.multiply_reduced_error_unrelinearized(sum, cipher_in, cipher_res);
        evaluator.relinearize_rescale_inplace(cipher_res, relin_keys);
        return;
    }
#endif
    evaluator.multiply_reduced_error(sum, cipher_in, relin_keys, cipher_res);
    evaluator.rescale_to_next_inplace(cipher_res);
}
