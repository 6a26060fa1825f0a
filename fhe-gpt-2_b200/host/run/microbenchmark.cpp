// run/microbenchmark.cpp - `./microbenchmark <entry>`: the reference's GPT-2 operator benchmarks
// (gpt2_ckks/run/microbenchmark.cpp) on the engine.
//
//   0 cipher-plain 128x128   1 attention projection   2 QK^T   3 SV   4 softmax   5 smax   6 gelu   7 layernorm
//   8 bootstrap
//
// Same set-up as the reference: logN = 16, the 37-prime chain {49, 46 x 21, 49 x 14, 60}, Hamming weight 192, the
// rotation-key list of the INIT macro, one full-slot Bootstrapper (logn = 15).  The reference wires entries 1 and 2 to
// the attention projection, 4 / 5 / 8 to softmax / smax / smax and 7 to LayerNorm and has no case for 0, 3 and 6; the
// missing ones run the operator their description names.  Each prints the reference's timing line.  Inputs are
// uniform in (-1, 1) (B200CKKS_SEED fixes them); LayerNorm inputs are scaled so that the hard-coded Newton starting
// point of compute_layernorm (323251) is in range - with (-1, 1) inputs the reference's own run overflows the encoder.
#include "gpt2/approx.h"
#include "gpt2/test_util.h"
#include <chrono>
#include <cstdlib>
#include <iostream>

using namespace seal;
using namespace gpt2;
using std::cout;
using std::endl;
using std::vector;

namespace
{
    struct Rig
    {
        SEALContext &context;
        KeyGenerator &keygen;
        CKKSEncoder &encoder;
        Encryptor &encryptor;
        Decryptor &decryptor;
        Evaluator &evaluator;
        GaloisKeys &gal_keys;
        RelinKeys &relin_keys;
    };
#define RIG r.encoder, r.encryptor, r.decryptor, r.evaluator, r.gal_keys, r.relin_keys

    double seconds_since(std::chrono::system_clock::time_point t0)
    {
        return std::chrono::duration<double>(std::chrono::system_clock::now() - t0).count();
    }

    void to_limbs(Rig &r, Ciphertext &c, std::size_t limbs)
    {
        while (c.coeff_modulus_size() > limbs)
            r.evaluator.mod_switch_to_next_inplace(c);
    }

    Ciphertext random_cipher(Rig &r, double magnitude = 1.0)
    {
        vvec v(1, vec(32768, 0.0));
        generate_random(v);
        for (auto &x : v[0])
            x *= magnitude;
        Plaintext plain;
        Ciphertext cipher;
        r.encoder.encode(v[0], encode_scale(), plain);
        r.encryptor.encrypt(plain, cipher);
        return cipher;
    }

    // microbenchmark.cpp:10-67 (and :69-125, which the reference also points at the projection)
    void benchmark_attn_proj(Rig &r)
    {
        vvec A1(128, vec(768)), A2(768, vec(768)), A_t(768, vec(768, 0.0));
        generate_random(A1);
        generate_random(A2);
        transpose(A2, A_t);
        vc A1_cipher, A2_cipher, output;
        init_output(12, output, RIG);
        pack_from_row(A1, A1_cipher, RIG);
        pack_from_row(A_t, A2_cipher, RIG);
        printf("Done packing into ciphertexts: A1: %zu  A2: %zu\n", A1_cipher.size(), A2_cipher.size());
        Plaintext plain;
        Ciphertext bias;
        r.encoder.encode(vec(32768, 0.0), encode_scale(), plain);
        r.encryptor.encrypt(plain, bias);
        auto start = std::chrono::system_clock::now();
        attn_proj_row_seal(A1_cipher, A2_cipher, bias, output, 128, 768, 768, 128, r.keygen, RIG);
        r.context.sync();
        cout << "attn_proj time : " << seconds_since(start) << "s" << endl;
    }

    void benchmark_cipher_plain(Rig &r)
    {
        Ciphertext input = random_cipher(r), bias;
        std::unordered_map<std::string, vec> weights;
        vvec w(1, vec(32768));
        generate_random(w);
        weights["test"] = w[0];
        vc outputs;
        auto start = std::chrono::system_clock::now();
        cipher_plain_128_128(input, weights, bias, outputs, 128, 128, 128, 128, r.keygen, RIG);
        r.context.sync();
        cout << "cipher_plain_128 time : " << seconds_since(start) << "s" << endl;
    }

    // one head, at the level the attention block reaches them (10 limbs): the reference's loops are 128 x 128 (QK^T)
    // and 64 x 128 (SV) single-slot moves per head, each with its own freshly generated Galois key
    void benchmark_head_product(Rig &r, bool qk)
    {
        vc first{ random_cipher(r) }, second{ random_cipher(r) }, outputs;
        to_limbs(r, first[0], 10);
        to_limbs(r, second[0], 10);
        init_output(qk ? 1 : 128, outputs, RIG);
        auto start = std::chrono::system_clock::now();
        if (qk)
            qk_matmul(first, second, outputs, 128, 768, 768, 128, r.keygen, RIG);
        else
            sv_matmul(first, second, outputs, 128, 128, 128, 768, r.keygen, RIG);
        r.context.sync();
        cout << (qk ? "qk_matmul" : "sv_matmul") << " time (one head) : " << seconds_since(start) << "s" << endl;
    }

    // microbenchmark.cpp:127-168
    void benchmark_softmax(Rig &r, Bootstrapper &bootstrapper, bool optimized)
    {
        Ciphertext cipher = random_cipher(r);
        to_limbs(r, cipher, TOTAL_LEVEL - BOOT_LEVEL - 1);
        auto start = std::chrono::system_clock::now();
        if (!optimized)
            compute_softmax(cipher, 6, bootstrapper, RIG);
        else
            compute_smax(cipher, 6, (int)0.1, RIG);
        r.context.sync();
        cout << (optimized ? "Smax time: " : "Softmax time: ") << seconds_since(start) << "s" << endl;
        decrypt_and_print_and_max_round(cipher, r.decryptor, r.encoder, 1.0, 0);
    }

    void benchmark_gelu(Rig &r, Bootstrapper &bootstrapper)
    {
        Ciphertext cipher = random_cipher(r), out;
        to_limbs(r, cipher, TOTAL_LEVEL - BOOT_LEVEL);
        auto start = std::chrono::system_clock::now();
        compute_gelu(cipher, out, bootstrapper, RIG);
        r.context.sync();
        cout << "gelu time: " << seconds_since(start) << "s" << endl;
    }

    // microbenchmark.cpp:170-204
    void benchmark_layernorm(Rig &r)
    {
        Ciphertext cipher = random_cipher(r, 0.046), out;
        to_limbs(r, cipher, TOTAL_LEVEL - BOOT_LEVEL);
        vec gamma(768, 0.8), beta(768, 0.9);
        auto start = std::chrono::system_clock::now();
        compute_layernorm(cipher, out, gamma, beta, 768, RIG);
        r.context.sync();
        cout << "lnorm_time: " << seconds_since(start) << "s" << endl;
        decrypt_and_print_and_max_round(out, r.decryptor, r.encoder, 1.0, 0);
    }

    void benchmark_bootstrap(Rig &r, Bootstrapper &bootstrapper)
    {
        Ciphertext cipher = random_cipher(r), out;
        auto start = std::chrono::system_clock::now();
        bootstrap(cipher, out, bootstrapper, r.evaluator);
        r.context.sync();
        cout << "bootstrap time: " << seconds_since(start) << "s" << endl;
    }
} // namespace

int main(int argc, char *argv[])
{
    std::map<int, test_entry_t> tests;
    populate_tests(tests);
    if (argc < 2)
    {
        cout << "Please specify which microbenchmark you want to run. " << endl;
        print_tests(tests);
        return 0;
    }
    const int entry = std::atoi(argv[1]);
    if (entry < CIPHER_PLAIN_128 || entry > BOOTSTRAP)
    {
        cout << "Invalid entry. Please input a valid test type: " << endl;
        print_tests(tests);
        return 0;
    }

    // microbenchmark.cpp:217-237 and the INIT macro (util.h:37-75)
    const long boundary_K = 25, deg = 59, scale_factor = 2, inverse_deg = 1, logN = 16, loge = 10, logn = logN - 1;
    const int remaining_level = 21, boot_level = 14, total_level = remaining_level + boot_level;
    cout << "Setting Parameters" << endl;
    EncryptionParameters params(scheme_type::ckks);
    const std::size_t poly_modulus_degree = std::size_t(1) << logN;
    params.set_poly_modulus_degree(poly_modulus_degree);
    params.set_coeff_modulus(CoeffModulus::Create(poly_modulus_degree, init_coeff_bit_vec(LOGQ, LOGP, remaining_level, boot_level, 60)));
    params.set_secret_key_hamming_weight(192);
    const double scale = std::pow(2.0, LOGP);
    SEALContext context(params);
    KeyGenerator keygen(context);
    PublicKey public_key;
    keygen.create_public_key(public_key);
    SecretKey secret_key = keygen.secret_key();
    RelinKeys relin_keys;
    keygen.create_relin_keys(relin_keys);
    GaloisKeys gal_keys;
    vector<int> gal_steps_vector = init_rotation_steps((int)logN);
    CKKSEncoder encoder(context);
    Encryptor encryptor(context, public_key);
    Evaluator evaluator(context, encoder);
    Decryptor decryptor(context, secret_key);

    Bootstrapper bootstrapper(loge, logn, logN - 1, total_level, scale, boundary_K, deg, scale_factor, inverse_deg, context, keygen,
                              encoder, encryptor, decryptor, evaluator, relin_keys, gal_keys);
    cout << "Generating Optimal Minimax Polynomials..." << endl;
    bootstrapper.prepare_mod_polynomial();
    cout << "Adding Bootstrapping Keys..." << endl;
    bootstrapper.addLeftRotKeys_Linear_to_vector_3(gal_steps_vector);
    // steps the operators issue that the INIT list reaches only through SEAL's power-of-two fallback
    for (int extra : { -1024, -128, 32768 - 128, 32768 - 64, 32768 - 1024 })
        gal_steps_vector.push_back(extra);
    for (int rots = 1; rots < 64; rots++)
        gal_steps_vector.push_back(rots * 256);
    keygen.create_galois_keys(gal_steps_vector, gal_keys);
    bootstrapper.slot_vec.push_back(logn);
    cout << "Generating Linear Transformation Coefficients..." << endl;
    bootstrapper.generate_LT_coefficient_3();

    Rig r{ context, keygen, encoder, encryptor, decryptor, evaluator, gal_keys, relin_keys };
    cout << "Executing: " << tests[entry].name << endl;
    switch (entry)
    {
    case CIPHER_PLAIN_128: benchmark_cipher_plain(r); break;
    case ATTN_PROJ_ROW: benchmark_attn_proj(r); break;
    case QK_MATMUL: benchmark_head_product(r, true); break;
    case SV_MATMUL: benchmark_head_product(r, false); break;
    case SOFTMAX: benchmark_softmax(r, bootstrapper, false); break;
    case SMAX: benchmark_softmax(r, bootstrapper, true); break;
    case GELU: benchmark_gelu(r, bootstrapper); break;
    case LAYERNORM: benchmark_layernorm(r); break;
    case BOOTSTRAP: benchmark_bootstrap(r, bootstrapper); break;
    }
    cout << "Done Executing " << tests[entry].name << endl;
    return 0;
}
