// oracle/cnn_ref_capi.cpp - TEST INFRASTRUCTURE ONLY (checker, never the product path).
//
// A C ABI over the REFERENCE's OWN application classes, compiled unmodified from /root/reference/cnn_ckks
// (common/*.cpp, cpu-ckks/single-key/{ckks_bootstrapping,comp,cnn}/*.cpp) with oracle/ntl_shim standing in for NTL:
//
//   oracle/_ref/libcnn_ref.so     those sources + this file against the reference's modified SEAL 3.6.6
//                                 -> the L2-L4 oracle: the reference's Bootstrapper / minimax ReLU / multiplexed
//                                    convolution code on the reference's library, on the CPU
//   oracle/_ref/libcnn_dropin.so  the same sources + this file against fhe-gpt-2_b200/host/seal/seal.h and
//                                 libb200ckks.so -> the drop-in proof: the reference's object code on the B200 engine
//
// It exports the subset of include/b200ckks_app.h (same names, same argument meaning) that the reference's classes
// can serve, so that tests drive all four combinations {restated, reference} x {reference SEAL, engine} through one
// Python binding (b200ckks.app.App(path)).  Each entry cites the reference function it calls.
#include "cnn_seal.h"      // /root/reference/cnn_ckks/cpu-ckks/single-key/cnn/cnn_seal.h (pulls in Bootstrapper.h, SEALcomp.h, ...)
#include "infer_seal.h"
#include <cstring>
#include <memory>
#include <string>
#include <unistd.h>

using namespace seal;
using std::vector;

namespace
{
    thread_local std::string g_err;
    int fail(const std::exception &e, int code)
    {
        g_err = e.what();
        return code;
    }
#define BKA_TRY try {
#define BKA_END                                                                                                        \
    }                                                                                                                  \
    catch (const std::invalid_argument &e) { return fail(e, 1); }                                                      \
    catch (const std::out_of_range &e) { return fail(e, 3); }                                                          \
    catch (const std::logic_error &e) { return fail(e, 2); }                                                           \
    catch (const std::exception &e) { return fail(e, 4); }                                                             \
    return 0;
} // namespace

struct bka_ct_s
{
    Ciphertext ct;
};
typedef bka_ct_s *bka_ct_t;

struct bka_session_s
{
    int log_n = 0;
    vector<int> bits;
    EncryptionParameters parms{ scheme_type::ckks };
    std::unique_ptr<SEALContext> context;
    std::unique_ptr<KeyGenerator> keygen;
    PublicKey public_key;
    SecretKey secret_key;
    RelinKeys relin_keys;
    GaloisKeys gal_keys;
    std::unique_ptr<CKKSEncoder> encoder;
    std::unique_ptr<Encryptor> encryptor;
    std::unique_ptr<Evaluator> evaluator;
    std::unique_ptr<Decryptor> decryptor;
    vector<int> steps;
    bool keys_ready = false;
    vector<Tree> relu_tree;

    void add_steps(const int *s, int n)
    {
        for (int i = 0; i < n; i++)
            if (std::find(steps.begin(), steps.end(), s[i]) == steps.end())
            {
                steps.push_back(s[i]);
                keys_ready = false;
            }
    }
    void ensure_keys() // KeyGenerator::create_galois_keys on the collected step list (infer_seal.cpp:379)
    {
        if (keys_ready)
            return;
        keygen->create_galois_keys(steps, gal_keys);
        keys_ready = true;
    }
};
typedef bka_session_s *bka_session_t;

struct bka_bootstrapper_s
{
    bka_session_t s;
    std::unique_ptr<Bootstrapper> b;
    bool coeffs = false;
    void ready()
    {
        s->ensure_keys();
        if (!coeffs)
        {
            b->slot_vec.push_back(b->logn);       // infer_seal.cpp:381-383
            b->generate_LT_coefficient_3();       // infer_seal.cpp:386-388
            coeffs = true;
        }
    }
};
typedef bka_bootstrapper_s *bka_bootstrapper_t;

static bka_ct_t wrap(Ciphertext &&c)
{
    auto h = new bka_ct_s();
    h->ct = std::move(c);
    return h;
}

extern "C"
{
    const char *bka_last_error(void)
    {
        return g_err.c_str();
    }
    const char *bka_backend(void)
    {
#ifdef B200CKKS_FACADE
        return "reference-app/engine";
#else
        return "reference-app/reference-seal";
#endif
    }

    // the set-up block of ResNet_cifar10_seal_sparse, infer_seal.cpp:306-337
    int bka_session_create_with_secret(int log_n, const int *bit_sizes, int n_bits, int hamming_weight, int device,
                                       const int *rotation_steps, int n_steps, const uint64_t *secret_key, bka_session_t *out)
    {
        BKA_TRY
        if (secret_key)
            throw std::logic_error("importing a secret key is not part of the reference's API");
        auto s = std::make_unique<bka_session_s>();
        s->log_n = log_n;
        s->bits.assign(bit_sizes, bit_sizes + n_bits);
        const std::size_t N = std::size_t(1) << log_n;
        s->parms.set_poly_modulus_degree(N);
        s->parms.set_coeff_modulus(CoeffModulus::Create(N, s->bits));
        s->parms.set_secret_key_hamming_weight((std::size_t)hamming_weight);
#ifdef B200CKKS_FACADE
        s->context = std::make_unique<SEALContext>(s->parms, true, sec_level_type::none, device);
#else
        (void)device;
        // $B200CKKS_REF_SEED: a seeded factory hands every encryption the same stream (randomgen.h:440-447), which makes
        // two sessions - this library's and libapp_ref.so's - produce identical keys and fresh ciphertexts
        if (const char *e = std::getenv("B200CKKS_REF_SEED"))
        {
            std::uint64_t v = std::strtoull(e, nullptr, 0);
            prng_seed_type seed = { v, v ^ 0x9E3779B97F4A7C15ull, v + 1, v + 2, v + 3, v + 4, v + 5, v + 6 };
            s->parms.set_random_generator(std::make_shared<Blake2xbPRNGFactory>(seed));
        }
        s->context = std::make_unique<SEALContext>(s->parms, true, sec_level_type::none);
#endif
        s->keygen = std::make_unique<KeyGenerator>(*s->context);
        s->keygen->create_public_key(s->public_key);
        s->secret_key = s->keygen->secret_key();
        s->keygen->create_relin_keys(s->relin_keys);
        s->encoder = std::make_unique<CKKSEncoder>(*s->context);
        s->encryptor = std::make_unique<Encryptor>(*s->context, s->public_key);
        s->evaluator = std::make_unique<Evaluator>(*s->context, *s->encoder);
        s->decryptor = std::make_unique<Decryptor>(*s->context, s->secret_key);
        s->add_steps(rotation_steps, n_steps);
        *out = s.release();
        BKA_END
    }
    int bka_session_create(int log_n, const int *bit_sizes, int n_bits, int hamming_weight, int device, const int *rotation_steps,
                           int n_steps, bka_session_t *out)
    {
        return bka_session_create_with_secret(log_n, bit_sizes, n_bits, hamming_weight, device, rotation_steps, n_steps, nullptr, out);
    }
    int bka_session_destroy(bka_session_t s)
    {
        BKA_TRY
        delete s;
        BKA_END
    }
    int bka_session_add_rotation_steps(bka_session_t s, const int *steps, int n_steps)
    {
        BKA_TRY
        s->add_steps(steps, n_steps);
        BKA_END
    }
    int bka_session_primes(bka_session_t s, uint64_t *primes_out)
    {
        BKA_TRY
        const auto &m = s->context->key_context_data()->parms().coeff_modulus();
        for (std::size_t i = 0; i < m.size(); i++)
            primes_out[i] = m[i].value();
        BKA_END
    }
    int bka_session_sync(bka_session_t s)
    {
        BKA_TRY
#ifdef B200CKKS_FACADE
        s->context->sync();
#else
        (void)s;
#endif
        BKA_END
    }
    int bka_session_engine_context(bka_session_t s, void **bk_context_out)
    {
        BKA_TRY
#ifdef B200CKKS_FACADE
        *bk_context_out = (void *)s->context->handle();
#else
        (void)s;
        *bk_context_out = nullptr;
#endif
        BKA_END
    }
    // Evaluator operation counters (engine backend): what the reference's code asked the Evaluator to do
    int bka_session_stats(bka_session_t s, uint64_t counts_out[9], int reset)
    {
        BKA_TRY
#ifdef B200CKKS_FACADE
        auto &st = s->evaluator->stats();
        std::atomic<std::uint64_t> *f[9] = { &st.key_switch_rotate, &st.key_switch_relin, &st.rescale,    &st.multiply, &st.multiply_plain,
                                             &st.encode_vector,     &st.add,              &st.mod_switch, &st.scalar_op };
        for (int i = 0; i < 9; i++)
        {
            counts_out[i] = f[i]->load();
            if (reset)
                f[i]->store(0);
        }
#else
        (void)s;
        (void)reset;
        std::memset(counts_out, 0, 9 * sizeof(uint64_t));
#endif
        BKA_END
    }

    // ---- ciphertexts (TensorCipher's constructor, cnn_seal.cpp:12-44, does the same encode + encrypt) ------------
    int bka_encrypt(bka_session_t s, const double *values, int n_values, int is_complex, double scale, int limbs, bka_ct_t *out)
    {
        BKA_TRY
        Plaintext plain;
        if (is_complex)
        {
            vector<std::complex<double>> v((std::size_t)n_values);
            for (int i = 0; i < n_values; i++)
                v[(std::size_t)i] = { values[2 * i], values[2 * i + 1] };
            s->encoder->encode(v, scale, plain);
        }
        else
            s->encoder->encode(vector<double>(values, values + n_values), scale, plain);
        Ciphertext ct;
        s->encryptor->encrypt(plain, ct);
        if (limbs > 0)
        {
            auto cd = s->context->first_context_data();
            while (cd && (int)cd->parms().coeff_modulus().size() > limbs)
                cd = cd->next_context_data();
            if (!cd || (int)cd->parms().coeff_modulus().size() != limbs)
                throw std::invalid_argument("limbs is out of range");
            s->evaluator->mod_switch_to_inplace(ct, cd->parms_id());
        }
        *out = wrap(std::move(ct));
        BKA_END
    }
    int bka_decrypt(bka_session_t s, bka_ct_t ct, double *out_complex)
    {
        BKA_TRY
        Plaintext plain;
        s->decryptor->decrypt(ct->ct, plain);
        vector<std::complex<double>> v;
        s->encoder->decode(plain, v);
        std::memcpy(out_complex, v.data(), v.size() * sizeof(std::complex<double>));
        BKA_END
    }
    int bka_ct_clone(bka_ct_t ct, bka_ct_t *out)
    {
        BKA_TRY
        Ciphertext c = ct->ct;
        *out = wrap(std::move(c));
        BKA_END
    }
    int bka_ct_free(bka_ct_t ct)
    {
        BKA_TRY
        delete ct;
        BKA_END
    }
    int bka_ct_info(bka_ct_t ct, int *size, int *limbs, double *scale)
    {
        BKA_TRY
        if (size)
            *size = (int)ct->ct.size();
        if (limbs)
            *limbs = (int)ct->ct.coeff_modulus_size();
        if (scale)
            *scale = ct->ct.scale();
        BKA_END
    }
    int bka_ct_set_scale(bka_ct_t ct, double scale)
    {
        BKA_TRY
        ct->ct.scale() = scale;
        BKA_END
    }
    int bka_ct_mod_switch_to(bka_session_t s, bka_ct_t ct, int limbs)
    {
        BKA_TRY
        auto cd = s->context->first_context_data();
        while (cd && (int)cd->parms().coeff_modulus().size() > limbs)
            cd = cd->next_context_data();
        if (!cd)
            throw std::invalid_argument("limbs is out of range");
        s->evaluator->mod_switch_to_inplace(ct->ct, cd->parms_id());
        BKA_END
    }
    int bka_ct_download(bka_ct_t ct, uint64_t *host_out)
    {
        BKA_TRY
#ifdef B200CKKS_FACADE
        ct->ct.download(host_out);
#else
        std::memcpy(host_out, ct->ct.data(),
                    ct->ct.size() * ct->ct.coeff_modulus_size() * ct->ct.poly_modulus_degree() * sizeof(uint64_t));
#endif
        BKA_END
    }
    int bka_rotate(bka_session_t s, bka_ct_t ct, int steps)
    {
        BKA_TRY
        s->ensure_keys();
        s->evaluator->rotate_vector_inplace(ct->ct, steps, s->gal_keys);
        BKA_END
    }

    // ---- bootstrapping: class Bootstrapper, ckks_bootstrapping/Bootstrapper.h:14-201 -----------------------------
    // create = constructor (Bootstrapper.cpp:3-27) + prepare_mod_polynomial (the reference's own multi-interval Remez,
    // common/Remez.cpp, in NTL::RR at 1000 bits) + addLeftRotKeys_Linear_to_vector_3 (Bootstrapper.cpp:82-177)
    int bka_bootstrapper_create(bka_session_t s, int loge, int logn, int total_level, double final_scale, int boundary_k,
                                int sin_cos_deg, int scale_factor, int inverse_deg, bka_bootstrapper_t *out)
    {
        BKA_TRY
        auto h = std::make_unique<bka_bootstrapper_s>();
        h->s = s;
        h->b = std::make_unique<Bootstrapper>(loge, logn, s->log_n - 1, total_level, final_scale, boundary_k, sin_cos_deg,
                                              scale_factor, inverse_deg, *s->context, *s->keygen, *s->encoder, *s->encryptor,
                                              *s->decryptor, *s->evaluator, s->relin_keys, s->gal_keys);
        h->b->prepare_mod_polynomial();
        vector<int> steps;
        steps.push_back(0);
        for (int i = 0; i < s->log_n - 1; i++)
            steps.push_back(1 << i);
        h->b->addLeftRotKeys_Linear_to_vector_3(steps);
        s->add_steps(steps.data(), (int)steps.size());
        *out = h.release();
        BKA_END
    }
    int bka_bootstrapper_destroy(bka_bootstrapper_t b)
    {
        BKA_TRY
        delete b;
        BKA_END
    }
    int bka_bootstrapper_set_hoisting(bka_bootstrapper_t, int, int *previous)
    {
        if (previous)
            *previous = 0; // the reference issues its baby-step rotations one by one
        return 0;
    }
    int bka_bootstrapper_rotation_steps(bka_bootstrapper_t b, int *steps_out, int cap, int *count_out)
    {
        BKA_TRY
        vector<int> steps;
        b->b->addLeftRotKeys_Linear_to_vector_3(steps);
        *count_out = (int)steps.size();
        for (int i = 0; i < (int)steps.size() && i < cap; i++)
            steps_out[i] = steps[(std::size_t)i];
        BKA_END
    }
    int bka_bootstrapper_lt_coefficients(bka_bootstrapper_t b, int which, int *n_diagonals, int *length, double *data_out)
    {
        BKA_TRY
        if (!b->coeffs)
        {
            b->b->slot_vec.push_back(b->b->logn);
            b->b->generate_LT_coefficient_3();
            b->coeffs = true;
        }
        const std::size_t u = (std::size_t)b->b->slot_index;
        const vector<vector<std::complex<double>>> *d = nullptr;
        switch (which)
        {
        case 0: d = &b->b->fftcoeff1[u]; break;
        case 1: d = &b->b->fftcoeff2[u]; break;
        case 2: d = &b->b->fftcoeff3[u]; break;
        case 3: d = &b->b->invfftcoeff1[u]; break;
        case 4: d = &b->b->invfftcoeff2[u]; break;
        case 5: d = &b->b->invfftcoeff3[u]; break;
        default: throw std::invalid_argument("which must be 0..5");
        }
        *n_diagonals = (int)d->size();
        *length = d->empty() ? 0 : (int)(*d)[0].size();
        if (data_out)
            for (std::size_t i = 0; i < d->size(); i++)
                std::memcpy(data_out + 2 * i * (*d)[0].size(), (*d)[i].data(), (*d)[i].size() * sizeof(std::complex<double>));
        BKA_END
    }
    // bootstrap_3 (Bootstrapper.cpp:3409-3419) / bootstrap_real_3 (:3421-3431)
    int bka_bootstrap(bka_bootstrapper_t b, bka_ct_t ct, int real_message, bka_ct_t *out)
    {
        BKA_TRY
        b->ready();
        Ciphertext rtn;
        if (real_message)
            b->b->bootstrap_real_3(rtn, ct->ct);
        else
            b->b->bootstrap_3(rtn, ct->ct);
        *out = wrap(std::move(rtn));
        BKA_END
    }
    // ModularReducer::modular_reduction (ModularReducer.cpp:61-80)
    int bka_modular_reduction(bka_bootstrapper_t b, bka_ct_t ct, bka_ct_t *out)
    {
        BKA_TRY
        Ciphertext rtn;
        b->b->mod_reducer->modular_reduction(rtn, ct->ct);
        *out = wrap(std::move(rtn));
        BKA_END
    }
    // The EvalMod polynomial the reference's Remez produced: Chebyshev coefficients of sin_cos_polynomial after
    // generate_sin_cos_polynomial + the inverse_deg == 1 folding (ModularReducer.cpp:37-51), as doubles, and
    // scale_inverse_coeff.  Returns deg + 1 values (cap permitting).
    int bkr_evalmod_chebyshev(bka_bootstrapper_t b, double *cheb_out, int cap, int *count_out, double *scale_inverse_coeff)
    {
        BKA_TRY
        auto &p = b->b->mod_reducer->sin_cos_polynomial;
        *count_out = (int)p.deg + 1;
        for (long i = 0; i <= p.deg && i < cap; i++)
            cheb_out[i] = to_double(p.chebcoeff[i]);
        if (scale_inverse_coeff)
            *scale_inverse_coeff = b->b->mod_reducer->scale_inverse_coeff;
        BKA_END
    }

    // The baby-step/giant-step heap of that polynomial (Polynomial::generate_poly_heap, common/Polynomial.cpp:169-215)
    // as the doubles homomorphic_poly_evaluation hands to the Evaluator (`to_double(poly_heap[i]->chebcoeff[j])`,
    // :437-456): data = {heaplen, heap_k, heap_m, scale_inverse_coeff, then per node: deg (-1 = absent), deg + 1 values}.
    int bkr_evalmod_heap(bka_bootstrapper_t b, double *data, int cap, int *count_out)
    {
        BKA_TRY
        auto &p = b->b->mod_reducer->sin_cos_polynomial;
        vector<double> v = { (double)p.heaplen, (double)p.heap_k, (double)p.heap_m, b->b->mod_reducer->scale_inverse_coeff };
        for (long i = 0; i < p.heaplen; i++)
        {
            auto *node = p.poly_heap[i];
            v.push_back(node ? (double)node->deg : -1.0);
            if (node)
                for (long j = 0; j <= node->deg; j++)
                    v.push_back(to_double(node->chebcoeff[j]));
        }
        *count_out = (int)v.size();
        for (std::size_t i = 0; i < v.size() && (int)i < cap; i++)
            data[i] = v[i];
        BKA_END
    }

    // ---- approximate ReLU: minimax_ReLU_seal, comp/SEALcomp.cpp:3-60 (alpha = 13, degrees {15,15,27}, 1.7;
    // infer_seal.cpp:255-262).  Reads ../result/d13.txt relative to the working directory, as the reference does. ----
    int bka_relu(bka_session_t s, bka_ct_t ct, bka_ct_t *out)
    {
        BKA_TRY
        vector<int> deg = { 15, 15, 27 };
        if (s->relu_tree.empty())
            for (int d : deg)
            {
                Tree t;
                upgrade_oddbaby(d, t); // comp/program.cpp:3
                s->relu_tree.push_back(t);
            }
        Ciphertext res;
        minimax_ReLU_seal(3, deg, 13, s->relu_tree, 1.7, 46, *s->encryptor, *s->evaluator, *s->decryptor, *s->encoder,
                          s->public_key, s->secret_key, s->relin_keys, ct->ct, res);
        *out = wrap(std::move(res));
        BKA_END
    }
    int bka_oddbaby_tree(int deg, int *tree_out, int cap, int *len_out, int *depth_out, int *m_out, int *l_out)
    {
        BKA_TRY
        Tree t;
        upgrade_oddbaby(deg, t);
        *len_out = (int)t.tree.size();
        for (int i = 0; i < (int)t.tree.size() && i < cap; i++)
            tree_out[i] = t.tree[(std::size_t)i];
        *depth_out = t.depth;
        *m_out = t.m;
        *l_out = t.l;
        BKA_END
    }

    // ---- multiplexed-packing tensors: cnn/cnn_seal.cpp:284-787 --------------------------------------------------
    static TensorCipher tensor_of(const int p[7], bka_ct_t ct)
    {
        return TensorCipher(p[6], p[0], p[1], p[2], p[3], p[4], p[5], ct->ct);
    }
    static void parms_of(const TensorCipher &t, int p[7])
    {
        p[0] = t.k();
        p[1] = t.h();
        p[2] = t.w();
        p[3] = t.c();
        p[4] = t.t();
        p[5] = t.p();
        p[6] = t.logn();
    }
    int bka_conv(bka_session_t s, bka_ct_t in, const int in_parms[7], int co, int st, int fh, int fw, const double *weight,
                 const double *running_var, const double *constant_weight, double epsilon, int end, bka_ct_t *out, int out_parms[7])
    {
        BKA_TRY
        s->ensure_keys();
        TensorCipher tin = tensor_of(in_parms, in), tout;
        vector<Ciphertext> pool(14); // infer_seal.cpp:456: cipher_pool(14)
        vector<double> data(weight, weight + (std::size_t)fh * fw * tin.c() * co);
        multiplexed_parallel_convolution_seal(tin, tout, co, st, fh, fw, data, vector<double>(running_var, running_var + co),
                                              vector<double>(constant_weight, constant_weight + co), epsilon, *s->encoder,
                                              *s->encryptor, *s->evaluator, s->gal_keys, pool, end != 0);
        parms_of(tout, out_parms);
        *out = wrap(tout.cipher());
        BKA_END
    }
    int bka_bn(bka_session_t s, bka_ct_t in, const int parms[7], const double *bias, const double *running_mean,
               const double *running_var, const double *weight, double epsilon, double B, bka_ct_t *out)
    {
        BKA_TRY
        TensorCipher tin = tensor_of(parms, in), tout;
        const int c = tin.c();
        multiplexed_parallel_batch_norm_seal(tin, tout, vector<double>(bias, bias + c), vector<double>(running_mean, running_mean + c),
                                             vector<double>(running_var, running_var + c), vector<double>(weight, weight + c),
                                             epsilon, *s->encoder, *s->encryptor, *s->evaluator, B);
        *out = wrap(tout.cipher());
        BKA_END
    }
    int bka_downsample(bka_session_t s, bka_ct_t in, const int in_parms[7], bka_ct_t *out, int out_parms[7])
    {
        BKA_TRY
        s->ensure_keys();
        TensorCipher tin = tensor_of(in_parms, in), tout;
        multiplexed_parallel_downsampling_seal(tin, tout, *s->evaluator, s->gal_keys);
        parms_of(tout, out_parms);
        *out = wrap(tout.cipher());
        BKA_END
    }
    int bka_avgpool(bka_session_t s, bka_ct_t in, const int in_parms[7], double B, bka_ct_t *out, int out_parms[7])
    {
        BKA_TRY
        s->ensure_keys();
        TensorCipher tin = tensor_of(in_parms, in), tout;
        std::ofstream sink("/dev/null");
        averagepooling_seal_scale(tin, tout, *s->evaluator, s->gal_keys, B, *s->encoder, *s->decryptor, sink);
        parms_of(tout, out_parms);
        *out = wrap(tout.cipher());
        BKA_END
    }
    int bka_fc(bka_session_t s, bka_ct_t in, const int parms[7], const double *matrix, const double *bias, int q, int r, bka_ct_t *out)
    {
        BKA_TRY
        s->ensure_keys();
        TensorCipher tin = tensor_of(parms, in), tout;
        matrix_multiplication_seal(tin, tout, vector<double>(matrix, matrix + (std::size_t)q * r), vector<double>(bias, bias + q), q, r,
                                   *s->evaluator, s->gal_keys);
        *out = wrap(tout.cipher());
        BKA_END
    }
    int bka_tensor_add(bka_session_t s, bka_ct_t a, bka_ct_t b, bka_ct_t *out)
    {
        BKA_TRY
        int p[7] = { 1, 1, 1, 1, 1, 1, s->log_n - 1 };
        TensorCipher ta = tensor_of(p, a), tb = tensor_of(p, b), tout;
        cnn_add_seal(ta, tb, tout, *s->evaluator); // cnn_seal.cpp:593-609
        *out = wrap(tout.cipher());
        BKA_END
    }

    // ---- the whole driver: ResNet_cifar10_seal_sparse(layer_num, start, end), infer_seal.cpp:251-584, exactly as
    // run/run_cnn.cpp calls it.  Paths are the reference's (relative to the working directory):
    // ../../pretrained_parameters/resnet<L>_new/*.txt, ../../../testFile/test_values.txt, ../result/*.
    int bkr_run_cnn(int layer_num, int start_image, int end_image)
    {
        BKA_TRY
        ResNet_cifar10_seal_sparse((std::size_t)layer_num, (std::size_t)start_image, (std::size_t)end_image);
        BKA_END
    }
    int bkr_chdir(const char *path)
    {
        return chdir(path) == 0 ? 0 : 4;
    }
}
