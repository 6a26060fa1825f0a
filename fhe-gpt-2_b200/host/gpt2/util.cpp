// gpt2/util.cpp - see util.h.  Follows gpt2_ckks/gpt2-ckks/single-key/gpt2/util.cpp of the reference.
#include "gpt2/util.h"
#include "common/cached.h"
#include <algorithm>
#include <chrono>
#include <complex>
#include <cstdlib>
#include <iostream>
#include <stdexcept>

using namespace seal;
using std::vector;

namespace gpt2
{
    bool verbose()
    {
        static const bool on = std::getenv("B200CKKS_GPT2_VERBOSE") != nullptr;
        return on;
    }

    // util.cpp:27-63: zero-padded data vector of 2^logn entries, encoded at 2^logp and encrypted
    TensorCipher::TensorCipher(int logn, int k, int h, int w, int c, int t, int p, vector<double> data, Encryptor &encryptor,
                               CKKSEncoder &encoder, int logp)
        : k_(k), h_(h), w_(w), c_(c), t_(t), p_(p), logn_(logn)
    {
        if (k != 1)
            throw std::invalid_argument("supported k is only 1 right now");
        if (logn < 1 || logn > 16)
            throw std::out_of_range("the value of logn is out of range");
        if (data.size() > (std::size_t(1) << logn))
            throw std::out_of_range("the size of data is larger than n");
        data.resize(std::size_t(1) << logn, 0.0);
        Plaintext plain;
        encoder.encode(data, std::pow(2.0, logp), plain);
        encryptor.encrypt(plain, cipher_);
    }
    void TensorCipher::print_parms() const
    {
        std::cout << "k: " << k_ << "\nh: " << h_ << "\nw: " << w_ << "\nc: " << c_ << "\nt: " << t_ << "\np: " << p_ << std::endl;
    }

    Config::Config(long boundary_K_, long boot_deg_, long scale_factor_, long inverse_deg_, long logN_, long loge_, long logn_,
                   long logn_1_, long logn_2_, long logn_3_, int logp_, int logq_, int log_special_prime_, int, int remaining_level_,
                   int boot_level_, int)
        : boundary_K(boundary_K_), boot_deg(boot_deg_), scale_factor(scale_factor_), inverse_deg(inverse_deg_), logN(logN_),
          loge(loge_), logn(logn_), logn_1(logn_1_), logn_2(logn_2_), logn_3(logn_3_), logp(logp_), logq(logq_),
          log_special_prime(log_special_prime_), log_integer_part(logq_ - logp_ - (int)loge_ + 5), remaining_level(remaining_level_),
          boot_level(boot_level_), total_level(remaining_level_ + boot_level_)
    {}

    vector<int> init_coeff_bit_vec(int logq, int logp, int remaining_level, int boot_level, int log_special_prime)
    {
        vector<int> bits{ logq };
        bits.insert(bits.end(), (std::size_t)remaining_level, logp);
        bits.insert(bits.end(), (std::size_t)boot_level, logq);
        bits.push_back(log_special_prime);
        return bits;
    }
    vector<int> init_rotation_steps(int logN)
    {
        vector<int> steps;
        for (int i = 0; i < logN - 1; i++)
            steps.push_back(1 << i);
        vector<int> kinds = { 0,     1,     2,     3,     4,     5,     6,     7,     8,     9,     10,    32640, 31744, 12288,
                              16384, 20480, 24576, 28672, 32672, 32704, 32736, 32,    64,    96,    31872, 32096, 32320, 32544,
                              224,   448,   672,   896,   32765, 32766, 32767, 32740, 32747, 32754, 32761, 14,    21,    28 };
        for (int i = 0; i < 32768; i += 2048)
            kinds.push_back(i);
        for (int r : kinds)
            if (std::find(steps.begin(), steps.end(), r) == steps.end())
                steps.push_back(r);
        return steps;
    }

    int round_to_2(double x)
    {
        return (int)std::pow(2.0, std::ceil(std::log2(x)));
    }

    namespace
    {
        // util.cpp:206-222: negative steps wrap to slots + steps; the half-ring rotation is taken in two hops
        // (slots/2 - slots/16, then slots/16), the reference's 16384 = 14336 + 2048
        int wrapped(const Ciphertext &c, int steps)
        {
            const int slots = (int)(c.poly_modulus_degree() / 2);
            return steps < 0 ? slots + steps : steps;
        }
    } // namespace

    void rotate_inplace(Ciphertext &cipher_in, int steps, Evaluator &evaluator, GaloisKeys &gal_keys)
    {
        if (steps == 0)
            return;
        const int slots = (int)(cipher_in.poly_modulus_degree() / 2), amount = wrapped(cipher_in, steps);
        if (amount == slots / 2)
        {
            evaluator.rotate_vector_inplace(cipher_in, amount - slots / 16, gal_keys);
            evaluator.rotate_vector_inplace(cipher_in, slots / 16, gal_keys);
        }
        else
            evaluator.rotate_vector_inplace(cipher_in, amount, gal_keys);
    }

    // as in the reference, a zero step leaves cipher_out untouched (util.cpp:228)
    void rotate_vec(const Ciphertext &cipher_in, Ciphertext &cipher_out, int steps, Evaluator &evaluator, GaloisKeys &gal_keys)
    {
        if (steps == 0)
            return;
        const int slots = (int)(cipher_in.poly_modulus_degree() / 2), amount = wrapped(cipher_in, steps);
        if (amount == slots / 2)
        {
            evaluator.rotate_vector(cipher_in, amount - slots / 16, gal_keys, cipher_out);
            evaluator.rotate_vector_inplace(cipher_out, slots / 16, gal_keys);
        }
        else
            evaluator.rotate_vector(cipher_in, amount, gal_keys, cipher_out);
    }

    void fakeBootstrap(Ciphertext &input, Ciphertext &output, CKKSEncoder &encoder, Encryptor &encryptor, Decryptor &decryptor,
                       Evaluator &, GaloisKeys &, RelinKeys &)
    {
        Plaintext plain;
        vector<double> values;
        decryptor.decrypt(input, plain);
        encoder.decode(plain, values);
        encoder.encode(values, encode_scale(), plain);
        encryptor.encrypt(plain, output);
    }

    void init_output(int num_ciphers, vc &output, CKKSEncoder &encoder, Encryptor &encryptor, Decryptor &, Evaluator &, GaloisKeys &,
                     RelinKeys &)
    {
        const vector<double> zero(1, 0.0);
        Plaintext plain;
        for (int i = 0; i < num_ciphers; i++)
        {
            Ciphertext cipher;
            encoder.encode(zero, encode_scale(), plain);
            encryptor.encrypt(plain, cipher);
            output.push_back(std::move(cipher));
        }
    }

    // util.cpp:279-287 of the reference: multiply by a 0/1 slot vector, rescale.  The matrix operators ask for the same
    // few masks thousands of times; on the engine the encoded mask is kept per (start, length, level, scale)
    // (common/cached.h), on stock SEAL this is the reference's multiply_vector_reduced_error.
    void mask_out(Ciphertext &cipher, Ciphertext &out, int start, int length, CKKSEncoder &encoder, Evaluator &evaluator, RelinKeys &)
    {
        static const char mask_owner = 0;
        vector<double> mask;
        if (&out != &cipher)
            out = cipher;
        multiply_vector_named(evaluator, out, &mask_owner, ((std::uint64_t)(std::uint32_t)start << 32) | (std::uint32_t)length, 0,
                              [&]() -> const vector<double> & {
                                  mask.assign(encoder.slot_count(), 0.0);
                                  std::fill(mask.begin() + start, mask.begin() + start + length, 1.0);
                                  return mask;
                              });
        evaluator.rescale_to_next_inplace(out);
    }

    // row i of v goes to offset i * 2 * round_to_2(row_size) of the concatenated output ciphertext slots
    void pack_plain_row(vvec &v, int rows, int row_size, vvec &out)
    {
        const long stride = 2L * round_to_2(row_size), slots = (long)out[0].size();
        for (int i = 0; i < rows; i++)
            for (int j = 0; j < row_size; j++)
            {
                const long pos = i * stride + j;
                out[(std::size_t)(pos / slots)][(std::size_t)(pos % slots)] = v[(std::size_t)i][(std::size_t)j];
            }
    }

    void add_galois_keys(vector<double> &)
    {}

    void init_bootstrap(Bootstrapper &bootstrapper, vector<int> &gal_steps_vector, int logn)
    {
        bootstrapper.prepare_mod_polynomial();
        bootstrapper.addLeftRotKeys_Linear_to_vector_3(gal_steps_vector);
        bootstrapper.slot_vec.push_back(logn);
        bootstrapper.generate_LT_coefficient_3();
    }

    void bootstrap(Ciphertext &ctxt, Ciphertext &rtn, Bootstrapper &bootstrapper, Evaluator &evaluator)
    {
        while (ctxt.coeff_modulus_size() > 1)
            evaluator.mod_switch_to_next_inplace(ctxt);
        if (&ctxt == &rtn)
        {
            // the reference's callers pass one object for both (Fold.cpp:93); hand the pipeline a private copy of the
            // exhausted ciphertext so that it never reads an operand it has begun to overwrite
            Ciphertext exhausted = ctxt;
            bootstrapper.bootstrap_3(rtn, exhausted);
        }
        else
            bootstrapper.bootstrap_3(rtn, ctxt);
    }

    void surefire_rotate(Ciphertext &cipher, int shift_amt, KeyGenerator &keygen, Evaluator &evaluator)
    {
        auto start = std::chrono::system_clock::now();
        vector<int> steps{ -shift_amt };
        GaloisKeys tmp_keys;
        keygen.create_galois_keys(steps, tmp_keys);
        if (verbose())
            std::cout << "Keygen time : " << std::chrono::duration<double>(std::chrono::system_clock::now() - start).count() << "s"
                      << std::endl;
        evaluator.rotate_vector_inplace(cipher, -shift_amt, tmp_keys);
    }

    void decrypt_and_print_and_max_round(const Ciphertext &cipher, Decryptor &decryptor, CKKSEncoder &encoder, double unit,
                                         long sparse_slots, std::size_t front, std::size_t back)
    {
        if (!verbose())
            return;
        Plaintext plain;
        decryptor.decrypt(cipher, plain);
        vector<std::complex<double>> values;
        encoder.decode(plain, values);
        std::cout << "( ";
        for (std::size_t i = 0; i < front; i++)
            std::cout << values[i] << ", ";
        std::cout << "... ";
        const std::size_t slots = sparse_slots == 0 ? cipher.poly_modulus_degree() / 2 : (std::size_t)sparse_slots;
        for (std::size_t i = 0; i < back; i++)
            std::cout << values[slots - back + i] << (i + 1 != back ? ", " : "");
        std::cout << ")" << std::endl;
        long max_round = 0;
        for (const auto &z : values)
            max_round = std::max(max_round, std::labs((long)std::llround(z.real() / unit)));
        std::cout << "max_round = " << max_round << std::endl;
    }
} // namespace gpt2
