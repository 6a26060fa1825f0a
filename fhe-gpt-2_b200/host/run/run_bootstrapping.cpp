// run/run_bootstrapping.cpp - the reference's stand-alone bootstrapping driver (cnn_ckks/run/run_bootstrapping.cpp:50-226)
// on the engine: parameters of the CNN (logN 16, primes 51 | 46x16 | 51x14 | 51, Hamming weight 192), three sparse-slot
// bootstrappers (logn 14 / 13 / 12), a random complex message, modulus switched to the last level, bootstrap_3, mean
// absolute error.  Optional argument: logN (12..16; the slot counts shrink with it) for a quick run.
#include "ckks_bootstrapping/Bootstrapper.h"
#include <chrono>
#include <cstdlib>
#include <iostream>
#include <random>

using namespace seal;
using namespace std;

int main(int argc, char **argv)
{
    const long boundary_K = 25, deg = 59, scale_factor = 2, inverse_deg = 1, loge = 10;
    const long logN = argc > 1 ? std::atol(argv[1]) : 16;
    const long logns[3] = { logN - 2, logN - 3, logN - 4 };
    const int logp = 46, logq = 51, log_special_prime = 51, secret_key_hamming_weight = logN >= 15 ? 192 : 64;
    const int log_integer_part = logq - logp - (int)loge + 5;
    const int remaining_level = 16, boot_level = 14, total_level = remaining_level + boot_level;

    vector<int> coeff_bit_vec;
    coeff_bit_vec.push_back(logq);
    for (int i = 0; i < remaining_level; i++)
        coeff_bit_vec.push_back(logp);
    for (int i = 0; i < boot_level; i++)
        coeff_bit_vec.push_back(logq);
    coeff_bit_vec.push_back(log_special_prime);

    cout << "Setting Parameters" << endl;
    EncryptionParameters parms(scheme_type::ckks);
    const size_t poly_modulus_degree = (size_t)1 << logN;
    parms.set_poly_modulus_degree(poly_modulus_degree);
    parms.set_coeff_modulus(CoeffModulus::Create(poly_modulus_degree, coeff_bit_vec));
    const double scale = pow(2.0, logp);
    parms.set_secret_key_hamming_weight(secret_key_hamming_weight);

    SEALContext context(parms, true, sec_level_type::none);
    KeyGenerator keygen(context);
    PublicKey public_key;
    keygen.create_public_key(public_key);
    auto secret_key = keygen.secret_key();
    RelinKeys relin_keys;
    keygen.create_relin_keys(relin_keys);
    GaloisKeys gal_keys;
    CKKSEncoder encoder(context);
    Encryptor encryptor(context, public_key);
    Evaluator evaluator(context, encoder);
    Decryptor decryptor(context, secret_key);
    const size_t slot_count = encoder.slot_count();

    vector<unique_ptr<Bootstrapper>> boots;
    for (long logn : logns)
        boots.emplace_back(new Bootstrapper(loge, logn, logN - 1, total_level, scale, boundary_K, deg, scale_factor, inverse_deg,
                                            context, keygen, encoder, encryptor, decryptor, evaluator, relin_keys, gal_keys));
    cout << "Generating Optimal Minimax Polynomials..." << endl;
    for (auto &b : boots)
        b->prepare_mod_polynomial();
    cout << "Adding Bootstrapping Keys..." << endl;
    vector<int> gal_steps_vector;
    gal_steps_vector.push_back(0);
    for (int i = 0; i < logN - 1; i++)
        gal_steps_vector.push_back(1 << i);
    for (auto &b : boots)
        b->addLeftRotKeys_Linear_to_vector_3(gal_steps_vector);
    keygen.create_galois_keys(gal_steps_vector, gal_keys);
    for (auto &b : boots)
        b->slot_vec.push_back(b->logn);
    cout << "Generating Linear Transformation Coefficients..." << endl;
    for (auto &b : boots)
        b->generate_LT_coefficient_3();

    std::mt19937_64 rng(1);
    std::uniform_real_distribution<double> unif(-1.0, 1.0);
    vector<complex<double>> sparse((size_t)1 << logns[0]), input(slot_count), before, after;
    for (auto &v : sparse)
        v = { unif(rng), unif(rng) };

    double worst = 0;
    for (size_t it = 0; it < 3; it++)
    {
        const size_t sparse_slots = (size_t)1 << logns[it];
        cout << it << "-th iteration : sparse_slots = " << sparse_slots << endl;
        for (size_t i = 0; i < slot_count; i++)
            input[i] = static_cast<double>(1 << log_integer_part) * sparse[i % sparse_slots];
        Plaintext plain;
        Ciphertext cipher, rtn;
        encoder.encode(input, scale, plain);
        encryptor.encrypt(plain, cipher);
        for (int i = 0; i < total_level; i++)
            evaluator.mod_switch_to_next_inplace(cipher);
        decryptor.decrypt(cipher, plain);
        encoder.decode(plain, before);

        auto t0 = chrono::system_clock::now();
        boots[it]->bootstrap_3(rtn, cipher);
        decryptor.decrypt(rtn, plain); // (also drains the GPU stream)
        chrono::duration<double> sec = chrono::system_clock::now() - t0;
        cout << "bootstrapping time : " << sec.count() << "s" << endl;
        encoder.decode(plain, after);

        double mean_err = 0;
        for (size_t i = 0; i < sparse_slots; i++)
        {
            if (i < 4)
                cout << i << " " << before[i] << " " << after[i] << endl;
            mean_err += abs(before[i].real() - after[i].real()) + abs(before[i].imag() - after[i].imag());
        }
        mean_err /= 2.0 * (double)sparse_slots;
        cout << "Absolute mean of error: " << mean_err << endl;
        cout << "remaining level : " << context.get_context_data(rtn.parms_id())->chain_index() << ", scale: " << rtn.scale() << endl;
        worst = max(worst, mean_err);
    }
    return worst < 1e-4 ? 0 : 1;
}
