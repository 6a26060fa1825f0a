// capi/app_capi.cpp - the C ABI of include/b200ckks_app.h over the C++ application classes.
//
// The same file is compiled twice: against fhe-gpt-2_b200/host/seal/seal.h (the engine; product) and, as test
// infrastructure only, against the reference's own SEAL headers (oracle/Makefile, target `app_ref`), which yields a
// CPU oracle for the application layers that runs the identical host code on the reference library.
#include "../../../include/b200ckks_app.h"
#include "cnn/infer_seal.h"
#include "gpt2/approx.h"
#include "gpt2/test_util.h"
#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <fstream>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>

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
    vector<minicomp::Tree> relu_tree;

    void add_steps(const int *s, int n)
    {
        for (int i = 0; i < n; i++)
            if (std::find(steps.begin(), steps.end(), s[i]) == steps.end())
            {
                steps.push_back(s[i]);
                keys_ready = false;
            }
    }
    // KeyGenerator::create_galois_keys on the collected step list (infer_seal.cpp:379)
    void ensure_keys()
    {
        if (keys_ready)
            return;
        keygen->create_galois_keys(steps, gal_keys);
        keys_ready = true;
    }
};

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
            b->slot_vec.push_back(b->logn);
            b->generate_LT_coefficient_3();
            coeffs = true;
        }
    }
};

// Host threads that each run whole images on their own CUDA stream - the reference's `#pragma omp parallel for` over
// images (infer_seal.cpp:404) with a fixed set of threads, so that the engine's per-thread streams and staging rings
// are created once.
class ImageWorkers
{
public:
    ~ImageWorkers()
    {
        {
            std::lock_guard<std::mutex> g(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : threads_)
            t.join();
    }
    // runs job(i) for i in [0, n) on up to `width` workers; rethrows the first exception
    void run(int n, int width, const std::function<void(int)> &job)
    {
        width = std::max(1, std::min(width, n));
        while ((int)threads_.size() < width)
            threads_.emplace_back([this] { loop(); });
        std::unique_lock<std::mutex> g(mu_);
        job_ = &job;
        next_ = 0;
        total_ = n;
        width_ = width;
        pending_ = n;
        error_ = nullptr;
        ++generation_;
        cv_.notify_all();
        done_.wait(g, [this] { return pending_ == 0; });
        job_ = nullptr;
        if (error_)
            std::rethrow_exception(error_);
    }

private:
    void loop()
    {
        std::uint64_t seen = 0;
        std::unique_lock<std::mutex> g(mu_);
        for (;;)
        {
            cv_.wait(g, [&] { return stop_ || (generation_ != seen && next_ < total_ && active_ < width_); });
            if (stop_)
                return;
            ++active_;
            while (next_ < total_)
            {
                const int i = next_++;
                const auto *job = job_;
                g.unlock();
                std::exception_ptr err;
                try
                {
                    (*job)(i);
                }
                catch (...)
                {
                    err = std::current_exception();
                }
                g.lock();
                if (err && !error_)
                    error_ = err;
                if (--pending_ == 0)
                    done_.notify_all();
            }
            --active_;
            seen = generation_;
        }
    }
    std::vector<std::thread> threads_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    const std::function<void(int)> *job_ = nullptr;
    int next_ = 0, total_ = 0, width_ = 0, active_ = 0, pending_ = 0;
    std::uint64_t generation_ = 0;
    std::exception_ptr error_;
    bool stop_ = false;
};

struct bka_resnet_s
{
    bka_session_t s;
    std::unique_ptr<ResNetCifar10> net;
    bool prepared = false;
    ImageWorkers workers;
};

static bka_ct_t wrap(Ciphertext &&c)
{
    auto *h = new bka_ct_s();
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
        return "engine";
#else
        return "reference-seal";
#endif
    }

    int bka_session_create(int log_n, const int *bit_sizes, int n_bits, int hamming_weight, int device,
                           const int *rotation_steps, int n_steps, bka_session_t *out)
    {
        BKA_TRY
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
        // $B200CKKS_REF_SEED (reference-SEAL checker build only): a seeded factory hands every encryption the same
        // stream (randomgen.h:440-447), so this session and one of oracle/_ref/libcnn_ref.so - the reference's own
        // application code - produce identical keys and fresh ciphertexts, and their results can be compared limb by limb
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
    int bka_session_create_with_secret(int log_n, const int *bit_sizes, int n_bits, int hamming_weight, int device,
                                       const int *rotation_steps, int n_steps, const uint64_t *secret_key, bka_session_t *out)
    {
        if (!secret_key)
            return bka_session_create(log_n, bit_sizes, n_bits, hamming_weight, device, rotation_steps, n_steps, out);
        BKA_TRY
#ifdef B200CKKS_FACADE
        auto s = std::make_unique<bka_session_s>();
        s->log_n = log_n;
        s->bits.assign(bit_sizes, bit_sizes + n_bits);
        const std::size_t N = std::size_t(1) << log_n;
        s->parms.set_poly_modulus_degree(N);
        s->parms.set_coeff_modulus(CoeffModulus::Create(N, s->bits));
        s->parms.set_secret_key_hamming_weight((std::size_t)hamming_weight);
        s->context = std::make_unique<SEALContext>(s->parms, true, sec_level_type::none, device);
        s->secret_key = SecretKey::upload(*s->context, secret_key);
        s->keygen = std::make_unique<KeyGenerator>(*s->context, s->secret_key);
        s->keygen->create_public_key(s->public_key);
        s->keygen->create_relin_keys(s->relin_keys);
        s->encoder = std::make_unique<CKKSEncoder>(*s->context);
        s->encryptor = std::make_unique<Encryptor>(*s->context, s->public_key);
        s->evaluator = std::make_unique<Evaluator>(*s->context, *s->encoder);
        s->decryptor = std::make_unique<Decryptor>(*s->context, s->secret_key);
        s->add_steps(rotation_steps, n_steps);
        *out = s.release();
#else
        throw std::logic_error("importing a secret key is an engine-backend feature");
#endif
        BKA_END
    }
    int bka_session_secret_key(bka_session_t s, uint64_t *host_out)
    {
        BKA_TRY
#ifdef B200CKKS_FACADE
        s->secret_key.download(host_out);
#else
        (void)s;
        (void)host_out;
        throw std::logic_error("exporting the secret key is an engine-backend feature");
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
    int bka_session_level_histogram(bka_session_t s, int which, uint64_t counts_out[64], int reset)
    {
        BKA_TRY
        if (which < 0 || which > 5)
            throw std::invalid_argument("which must be 0..5");
#ifdef B200CKKS_FACADE
        auto &st = s->evaluator->stats();
        for (int i = 0; i < 64; i++)
        {
            counts_out[i] = st.by_limbs[which][i].load();
            if (reset)
                st.by_limbs[which][i].store(0);
        }
#else
        (void)s;
        (void)reset;
        std::memset(counts_out, 0, 64 * sizeof(uint64_t));
#endif
        BKA_END
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
    int bka_session_key_residency(bka_session_t s, uint64_t *bytes_out, uint64_t *generated_out)
    {
        BKA_TRY
#ifdef B200CKKS_FACADE
        *bytes_out = s->gal_keys.resident_bytes();
        *generated_out = s->gal_keys.generated();
#else
        (void)s;
        *bytes_out = 0;
        *generated_out = 0;
#endif
        BKA_END
    }

    int bka_session_double_hoisted_groups(bka_session_t s, uint64_t *count_out)
    {
        BKA_TRY
#ifdef B200CKKS_FACADE
        *count_out = s->evaluator->stats().double_hoisted_groups.load();
#else
        (void)s;
        *count_out = 0;
#endif
        BKA_END
    }
    int bka_session_key_plan(bka_session_t s, char *text_out, int cap, int *length_out)
    {
        BKA_TRY
#ifdef B200CKKS_FACADE
        s->ensure_keys();
        std::string t = KeyPlan::capture(*s->context, s->relin_keys, s->gal_keys).to_string();
        if (length_out)
            *length_out = (int)t.size();
        if (text_out && cap > 0)
        {
            std::size_t n = std::min<std::size_t>(t.size(), (std::size_t)cap - 1);
            std::memcpy(text_out, t.data(), n);
            text_out[n] = 0;
        }
#else
        (void)s, (void)text_out, (void)cap, (void)length_out;
        throw std::logic_error("key plans are an engine-backend feature (the reference generates every key in full)");
#endif
        BKA_END
    }
    int bka_session_apply_key_plan(bka_session_t s, const char *text, int detach_secret)
    {
        BKA_TRY
#ifdef B200CKKS_FACADE
        s->ensure_keys();
        KeyPlan::from_string(text).generate(*s->context, s->relin_keys, s->gal_keys);
        if (detach_secret)
            KeyPlan::detach_secret(s->relin_keys, s->gal_keys);
#else
        (void)s, (void)text, (void)detach_secret;
        throw std::logic_error("key plans are an engine-backend feature (the reference generates every key in full)");
#endif
        BKA_END
    }

    int bka_session_plain_cache(bka_session_t s, uint64_t *bytes_out, uint64_t *hits_out, uint64_t *misses_out)
    {
        BKA_TRY
#ifdef B200CKKS_FACADE
        *bytes_out = s->evaluator->cached_plaintext_bytes();
        *hits_out = s->evaluator->stats().cache_hits.load();
        *misses_out = s->evaluator->stats().cache_misses.load();
#else
        (void)s;
        *bytes_out = *hits_out = *misses_out = 0;
#endif
        BKA_END
    }

    // ---- ciphertexts ---------------------------------------------------------------------------------------------
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
    int bka_multiply_relin_rescale(bka_session_t s, bka_ct_t a, bka_ct_t b)
    {
        BKA_TRY
        s->evaluator->multiply_inplace_reduced_error(a->ct, b->ct, s->relin_keys);
        s->evaluator->rescale_to_next_inplace(a->ct);
        BKA_END
    }
    int bka_add_reduced_error(bka_session_t s, bka_ct_t a, bka_ct_t b)
    {
        BKA_TRY
        s->evaluator->add_inplace_reduced_error(a->ct, b->ct);
        BKA_END
    }
    int bka_reduced_error_op(bka_session_t s, int which, bka_ct_t a, bka_ct_t b)
    {
        BKA_TRY
        switch (which)
        {
        case 0: s->evaluator->add_inplace_reduced_error(a->ct, b->ct); break;
        case 1: s->evaluator->sub_inplace_reduced_error(a->ct, b->ct); break;
        case 2: s->evaluator->multiply_inplace_reduced_error(a->ct, b->ct, s->relin_keys); break;
        default: throw std::invalid_argument("which must be 0 (add), 1 (sub) or 2 (multiply)");
        }
        BKA_END
    }
    int bka_ct_upload(bka_session_t s, const uint64_t *host, int size, int limbs, double scale, int is_ntt, bka_ct_t *out)
    {
        BKA_TRY
#ifdef B200CKKS_FACADE
        Ciphertext c;
        c.upload(*s->context, host, size, limbs, scale, is_ntt != 0);
        *out = wrap(std::move(c));
#else
        (void)s, (void)host, (void)size, (void)limbs, (void)scale, (void)is_ntt, (void)out;
        throw std::logic_error("raw ciphertext upload is an engine-backend feature");
#endif
        BKA_END
    }
    int bka_session_import_relin_key(bka_session_t s, const uint64_t *host, int digits)
    {
        BKA_TRY
#ifdef B200CKKS_FACADE
        auto k = std::make_shared<RelinKeys::Holder>();
        seal::detail::check(bk_kskey_upload(s->context->handle(), host, digits, 0, &k->h));
        s->relin_keys.k_ = k;
        s->relin_keys.ctx_ = s->context->impl();
#else
        (void)s, (void)host, (void)digits;
        throw std::logic_error("raw key import is an engine-backend feature");
#endif
        BKA_END
    }
    // SEAL's wire format through files.  what: 0 ciphertext, 2 relinearization keys, 3 Galois keys, 4 secret key,
    // 5 public key
    int bka_save(bka_session_t s, int what, bka_ct_t ct, const char *path)
    {
        BKA_TRY
        std::ofstream f(path, std::ios::binary);
        if (!f)
            throw std::runtime_error("cannot open file");
        switch (what)
        {
        case 0: ct->ct.save(f, compr_mode_type::none); break;
        case 2: s->relin_keys.save(f, compr_mode_type::none); break;
        case 3:
            s->ensure_keys();
            s->gal_keys.save(f, compr_mode_type::none);
            break;
        case 4: s->secret_key.save(f, compr_mode_type::none); break;
        case 5: s->public_key.save(f, compr_mode_type::none); break;
        default: throw std::invalid_argument("what must be 0, 2, 3, 4 or 5");
        }
        BKA_END
    }
    int bka_load(bka_session_t s, int what, const char *path, bka_ct_t *ct_out)
    {
        BKA_TRY
        std::ifstream f(path, std::ios::binary);
        if (!f)
            throw std::runtime_error("cannot open file");
        switch (what)
        {
        case 0:
        {
            Ciphertext c;
            c.load(*s->context, f);
            *ct_out = wrap(std::move(c));
            break;
        }
        case 2: s->relin_keys.load(*s->context, f); break;
        case 3:
            s->gal_keys.load(*s->context, f);
            s->keys_ready = true; // loaded keys are complete: nothing is generated afterwards
            break;
        case 4:
            s->secret_key.load(*s->context, f);
            s->decryptor = std::make_unique<Decryptor>(*s->context, s->secret_key);
            break;
        case 5:
            s->public_key.load(*s->context, f);
            s->encryptor = std::make_unique<Encryptor>(*s->context, s->public_key);
            break;
        default: throw std::invalid_argument("what must be 0, 2, 3, 4 or 5");
        }
        BKA_END
    }
    int bka_multiply_vector_rescale(bka_session_t s, bka_ct_t a, const double *values, int n_values, int is_complex)
    {
        BKA_TRY
        if (is_complex)
        {
            vector<std::complex<double>> v((std::size_t)n_values);
            for (int i = 0; i < n_values; i++)
                v[(std::size_t)i] = { values[2 * i], values[2 * i + 1] };
            s->evaluator->multiply_vector_inplace_reduced_error(a->ct, v);
        }
        else
        {
            vector<double> v(values, values + n_values);
            s->evaluator->multiply_vector_inplace_reduced_error(a->ct, v);
        }
        s->evaluator->rescale_to_next_inplace(a->ct);
        BKA_END
    }

    // ---- bootstrapping -------------------------------------------------------------------------------------------
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
    int bka_bootstrapper_set_hoisting(bka_bootstrapper_t b, int on, int *previous)
    {
        BKA_TRY
        if (previous)
            *previous = (b->b->hoisting ? 1 : 0) | ((b->b->hoisting && !b->b->double_hoisting) ? 2 : 0);
#ifdef B200CKKS_FACADE
        b->b->hoisting = (on & 1) != 0;
        b->b->double_hoisting = (on & 2) == 0;
#else
        (void)on; // stock SEAL has no hoisted rotation
#endif
        BKA_END
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
        { // coefficient tables do not need keys
            b->b->slot_vec.push_back(b->b->logn);
            b->b->generate_LT_coefficient_3();
            b->coeffs = true;
        }
        const std::size_t u = (std::size_t)b->b->slot_index;
        const Bootstrapper::Diagonals *d = nullptr;
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
    int bka_modular_reduction(bka_bootstrapper_t b, bka_ct_t ct, bka_ct_t *out)
    {
        BKA_TRY
        Ciphertext rtn;
        b->b->mod_reducer->modular_reduction(rtn, ct->ct);
        *out = wrap(std::move(rtn));
        BKA_END
    }

    int bka_bootstrapper_set_evalmod_heap(bka_bootstrapper_t b, const double *data, int count)
    {
        BKA_TRY
        if (count < 4)
            throw std::invalid_argument("heap data is too short");
        auto &red = *b->b->mod_reducer;
        auto &p = red.sin_cos_polynomial;
        const long heaplen = (long)data[0];
        p.heap_k = (long)data[1];
        p.heap_m = (long)data[2];
        p.heaplen = heaplen;
        red.scale_inverse_coeff = data[3];
        p.poly_heap.clear();
        p.poly_heap.resize((std::size_t)heaplen);
        int pos = 4;
        for (long i = 0; i < heaplen; i++)
        {
            if (pos >= count)
                throw std::invalid_argument("heap data is truncated");
            const long deg = (long)data[pos++];
            if (deg < 0)
                continue;
            if (pos + deg + 1 > count)
                throw std::invalid_argument("heap data is truncated");
            auto node = std::make_unique<boot::Polynomial>();
            node->set_zero_polynomial(deg);
            for (long j = 0; j <= deg; j++)
                node->chebcoeff[(std::size_t)j] = data[pos++];
            if (deg <= 3)
                node->cheb_to_power();
            p.poly_heap[(std::size_t)i] = std::move(node);
        }
        BKA_END
    }

    // ---- ReLU ----------------------------------------------------------------------------------------------------
    int bka_relu(bka_session_t s, bka_ct_t ct, bka_ct_t *out)
    {
        BKA_TRY
        vector<int> deg = { 15, 15, 27 };
        if (s->relu_tree.empty())
            for (int d : deg)
            {
                minicomp::Tree t;
                upgrade_oddbaby(d, t);
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
        minicomp::Tree t;
        upgrade_oddbaby(deg, t);
        *len_out = (int)t.tree.size();
        *depth_out = t.depth;
        *m_out = t.m;
        *l_out = t.l;
        for (int i = 0; i < (int)t.tree.size() && i < cap; i++)
            tree_out[i] = t.tree[(std::size_t)i];
        BKA_END
    }

    // ---- tensors -------------------------------------------------------------------------------------------------
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
                 const double *running_var, const double *constant_weight, double epsilon, int end, bka_ct_t *out,
                 int out_parms[7])
    {
        BKA_TRY
        s->ensure_keys();
        TensorCipher tin = tensor_of(in_parms, in), tout;
        vector<Ciphertext> pool;
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
        averagepooling_seal_scale(tin, tout, *s->evaluator, s->gal_keys, B);
        parms_of(tout, out_parms);
        *out = wrap(tout.cipher());
        BKA_END
    }
    int bka_fc(bka_session_t s, bka_ct_t in, const int parms[7], const double *matrix, const double *bias, int q, int r,
               bka_ct_t *out)
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
        Ciphertext c = a->ct;
        s->evaluator->add_inplace_reduced_error(c, b->ct);
        *out = wrap(std::move(c));
        BKA_END
    }

    // ---- ResNet --------------------------------------------------------------------------------------------------
    static void resnet_create(bka_session_t s, int layer_num, const double *conv_weight, const double *bn_bias, const double *bn_mean,
                              const double *bn_var, const double *bn_weight, const double *linear_weight, const double *linear_bias,
                              const ResNetVariant &variant, const double *sc_weight, const double *sc_bias, const double *sc_mean,
                              const double *sc_var, const double *sc_gamma, bka_resnet_t *out);
    int bka_resnet_create(bka_session_t s, int layer_num, const double *conv_weight, const double *bn_bias, const double *bn_mean,
                          const double *bn_var, const double *bn_weight, const double *linear_weight, const double *linear_bias,
                          bka_resnet_t *out)
    {
        BKA_TRY
        resnet_create(s, layer_num, conv_weight, bn_bias, bn_mean, bn_var, bn_weight, linear_weight, linear_bias,
                      ResNetVariant::cifar10(), nullptr, nullptr, nullptr, nullptr, nullptr, out);
        BKA_END
    }
    int bka_resnet_create_cifar100(bka_session_t s, int layer_num, const double *conv_weight, const double *bn_bias,
                                   const double *bn_mean, const double *bn_var, const double *bn_weight, const double *linear_weight,
                                   const double *linear_bias, const double *shortcut_weight, const double *shortcut_bn_bias,
                                   const double *shortcut_bn_mean, const double *shortcut_bn_var, const double *shortcut_bn_weight,
                                   bka_resnet_t *out)
    {
        BKA_TRY
        resnet_create(s, layer_num, conv_weight, bn_bias, bn_mean, bn_var, bn_weight, linear_weight, linear_bias,
                      ResNetVariant::cifar100(), shortcut_weight, shortcut_bn_bias, shortcut_bn_mean, shortcut_bn_var,
                      shortcut_bn_weight, out);
        BKA_END
    }
    int bka_resnet_classes(bka_resnet_t net, int *classes_out)
    {
        BKA_TRY
        *classes_out = net->net->classes();
        BKA_END
    }
    static void resnet_create(bka_session_t s, int layer_num, const double *conv_weight, const double *bn_bias, const double *bn_mean,
                              const double *bn_var, const double *bn_weight, const double *linear_weight, const double *linear_bias,
                              const ResNetVariant &variant, const double *sc_weight, const double *sc_bias, const double *sc_mean,
                              const double *sc_var, const double *sc_gamma, bka_resnet_t *out)
    {
        {
        ResNetParameters p;
        const std::size_t layers = (std::size_t)layer_num - 1;
        std::size_t wpos = 0, cpos = 0;
        for (std::size_t i = 0; i < layers; i++)
        {
            int ci, co;
            resnet_conv_shape((std::size_t)layer_num, i, ci, co);
            p.conv_weight.emplace_back(conv_weight + wpos, conv_weight + wpos + (std::size_t)9 * ci * co);
            wpos += (std::size_t)9 * ci * co;
            p.bn_bias.emplace_back(bn_bias + cpos, bn_bias + cpos + co);
            p.bn_running_mean.emplace_back(bn_mean + cpos, bn_mean + cpos + co);
            p.bn_running_var.emplace_back(bn_var + cpos, bn_var + cpos + co);
            p.bn_weight.emplace_back(bn_weight + cpos, bn_weight + cpos + co);
            cpos += (std::size_t)co;
        }
        p.linear_weight.assign(linear_weight, linear_weight + (std::size_t)variant.classes * 64);
        p.linear_bias.assign(linear_bias, linear_bias + variant.classes);
        if (variant.shortcut_conv)
        {
            std::size_t wp = 0, cp = 0;
            for (int j = 0; j < 2; j++)
            {
                const std::size_t ci = 16u << j, co = 32u << j;
                p.shortcut_weight.emplace_back(sc_weight + wp, sc_weight + wp + ci * co);
                wp += ci * co;
                p.shortcut_bn_bias.emplace_back(sc_bias + cp, sc_bias + cp + co);
                p.shortcut_bn_mean.emplace_back(sc_mean + cp, sc_mean + cp + co);
                p.shortcut_bn_var.emplace_back(sc_var + cp, sc_var + cp + co);
                p.shortcut_bn_weight.emplace_back(sc_gamma + cp, sc_gamma + cp + co);
                cp += co;
            }
        }
        auto h = std::make_unique<bka_resnet_s>();
        h->s = s;
        h->net = std::make_unique<ResNetCifar10>((std::size_t)layer_num, std::move(p), *s->context, *s->keygen, *s->encoder,
                                                 *s->encryptor, *s->decryptor, *s->evaluator, s->public_key, s->secret_key,
                                                 s->relin_keys, s->gal_keys, variant);
        vector<int> steps = h->net->galois_steps();
        s->add_steps(steps.data(), (int)steps.size());
        *out = h.release();
        }
    }
    int bka_resnet_destroy(bka_resnet_t net)
    {
        BKA_TRY
        delete net;
        BKA_END
    }
    static void resnet_ready(bka_resnet_t net)
    {
        net->s->ensure_keys();
        if (!net->prepared)
        {
            net->net->prepare();
            net->prepared = true;
        }
    }
    static void copy_trace(const vector<ResNetTraceRow> &trace, double *trace_out, int trace_cap, int *trace_rows)
    {
        if (trace_rows)
            *trace_rows = (int)trace.size();
        if (trace_out)
            for (int i = 0; i < (int)trace.size() && i < trace_cap; i++)
            {
                trace_out[4 * i + 0] = trace[(std::size_t)i].op;
                trace_out[4 * i + 1] = trace[(std::size_t)i].remaining_level;
                trace_out[4 * i + 2] = trace[(std::size_t)i].scale;
                trace_out[4 * i + 3] = trace[(std::size_t)i].milliseconds;
            }
    }
    int bka_resnet_encrypt_image(bka_resnet_t net, const double *image, bka_ct_t *out)
    {
        BKA_TRY
        TensorCipher t = net->net->encrypt_image(vector<double>(image, image + 3072));
        *out = wrap(t.cipher());
        BKA_END
    }
    int bka_resnet_infer_encrypted(bka_resnet_t net, bka_ct_t image_ct, bka_ct_t *logits_ct, double *trace_out, int trace_cap,
                                   int *trace_rows)
    {
        BKA_TRY
        resnet_ready(net);
        vector<ResNetTraceRow> trace;
        TensorCipher in(ResNetCifar10::logn, 1, 32, 32, 3, 3, 8, image_ct->ct);
        TensorCipher outt = net->net->infer_encrypted(in, trace_out ? &trace : nullptr);
        *logits_ct = wrap(outt.cipher());
        copy_trace(trace, trace_out, trace_cap, trace_rows);
        BKA_END
    }
    // n images, up to `in_flight` of them at a time, one host thread and CUDA stream each.  The workers' streams
    // start behind everything the caller has enqueued and the caller's stream continues behind all of them, so CUDA
    // events recorded by the caller around this call bracket the whole batch on the device.
    static void run_images(bka_resnet_t net, int n, int in_flight, const std::function<void(int)> &job)
    {
#ifdef B200CKKS_FACADE
        SEALContext &context = *net->s->context;
        SEALContext::Mark begin = context.mark();
        std::vector<SEALContext::Mark> ends((std::size_t)n);
        net->workers.run(n, in_flight, [&](int i) {
            context.after(begin);
            job(i);
            ends[(std::size_t)i] = context.mark();
        });
        for (auto &m : ends)
            context.after(m);
#else
        net->workers.run(n, in_flight, [&](int i) { job(i); });
#endif
    }
    int bka_resnet_infer_encrypted_batch(bka_resnet_t net, bka_ct_t *image_cts, int n_images, int in_flight, bka_ct_t *logits_cts)
    {
        BKA_TRY
        resnet_ready(net);
        std::vector<Ciphertext> results((std::size_t)n_images);
        run_images(net, n_images, in_flight, [&](int i) {
            TensorCipher in(ResNetCifar10::logn, 1, 32, 32, 3, 3, 8, image_cts[i]->ct);
            results[(std::size_t)i] = net->net->infer_encrypted(in, nullptr).cipher();
        });
        for (int i = 0; i < n_images; i++)
            logits_cts[i] = wrap(std::move(results[(std::size_t)i]));
        BKA_END
    }
    int bka_resnet_infer_batch(bka_resnet_t net, const double *images, int n_images, int in_flight, double *logits_out)
    {
        BKA_TRY
        resnet_ready(net);
        run_images(net, n_images, in_flight, [&](int i) {
            vector<double> logits = net->net->infer(vector<double>(images + (std::size_t)i * 3072, images + (std::size_t)(i + 1) * 3072));
            std::memcpy(logits_out + (std::size_t)i * logits.size(), logits.data(), logits.size() * sizeof(double));
        });
        BKA_END
    }
    int bka_resnet_decrypt_logits(bka_resnet_t net, bka_ct_t logits_ct, double *logits_out)
    {
        BKA_TRY
        TensorCipher t(ResNetCifar10::logn, 1, 1, 1, 64, 4, 1, logits_ct->ct);
        vector<double> logits = net->net->decrypt_logits(t);
        std::memcpy(logits_out, logits.data(), logits.size() * sizeof(double));
        BKA_END
    }
    int bka_resnet_infer(bka_resnet_t net, const double *image, double *logits_out, double *trace_out, int trace_cap,
                         int *trace_rows)
    {
        BKA_TRY
        resnet_ready(net);
        vector<ResNetTraceRow> trace;
        vector<double> logits = net->net->infer(vector<double>(image, image + 3072), trace_out ? &trace : nullptr);
        std::memcpy(logits_out, logits.data(), logits.size() * sizeof(double));
        if (trace_rows)
            *trace_rows = (int)trace.size();
        if (trace_out)
            for (int i = 0; i < (int)trace.size() && i < trace_cap; i++)
            {
                trace_out[4 * i + 0] = trace[(std::size_t)i].op;
                trace_out[4 * i + 1] = trace[(std::size_t)i].remaining_level;
                trace_out[4 * i + 2] = trace[(std::size_t)i].scale;
                trace_out[4 * i + 3] = trace[(std::size_t)i].milliseconds;
            }
        BKA_END
    }

    // ---- GPT-2 operators (gpt2/approx.h) -----------------------------------------------------------------------------
    int bka_gpt2_call(bka_session_t s, bka_bootstrapper_t boot, const char *op_name, bka_ct_t *in, int n_in, const double *dparams,
                      int n_d, const int *iparams, int n_i, bka_ct_t *out, int out_cap, int *n_out)
    {
        BKA_TRY
        const std::string op(op_name);
        s->ensure_keys();
        CKKSEncoder &enc = *s->encoder;
        Encryptor &cr = *s->encryptor;
        Decryptor &de = *s->decryptor;
        Evaluator &ev = *s->evaluator;
        GaloisKeys &gk = s->gal_keys;
        RelinKeys &rk = s->relin_keys;
        KeyGenerator &kg = *s->keygen;

        auto need = [&](int cts, int ints, int doubles) {
            if (n_in < cts || n_i < ints || n_d < doubles)
                throw std::invalid_argument("bka_gpt2_call(" + op + "): too few arguments");
        };
        // the operators that take a Bootstrapper& only to pass it on (sign_function never touches it) get a bare one
        std::unique_ptr<Bootstrapper> bare;
        auto bootstrapper = [&]() -> Bootstrapper & {
            if (boot)
            {
                boot->ready();
                return *boot->b;
            }
            if (!bare)
                bare = std::make_unique<Bootstrapper>(10, s->log_n - 1, s->log_n - 1, (int)s->bits.size() - 2, gpt2::encode_scale(), 25,
                                                      59, 2, 1, *s->context, kg, enc, cr, de, ev, rk, gk);
            return *bare;
        };
        vector<Ciphertext> results;
        auto range = [&](int from, int count) {
            gpt2::vc v;
            for (int i = 0; i < count; i++)
                v.push_back(in[from + i]->ct);
            return v;
        };
        Ciphertext r;
        if (op == "quickSum")
        {
            need(1, 1, 0);
            Ciphertext x = in[0]->ct;
            gpt2::quickSum(x, r, iparams[0], enc, cr, de, ev, gk, rk);
            results.push_back(r);
        }
        else if (op == "sign_f" || op == "sign_g" || op == "gelu_p" || op == "gelu_q")
        {
            need(1, 0, 0);
            Ciphertext x = in[0]->ct;
            (op == "sign_f"   ? gpt2::compute_sign_f
             : op == "sign_g" ? gpt2::compute_sign_g
             : op == "gelu_p" ? gpt2::compute_gelu_p
                              : gpt2::compute_gelu_q)(x, r, enc, cr, de, ev, gk, rk);
            results.push_back(r);
        }
        else if (op == "cheby_basis")
        {
            need(1, 1, 0);
            Ciphertext x = in[0]->ct;
            gpt2::build_cheby_basis(x, results, iparams[0], enc, cr, de, ev, gk, rk);
        }
        else if (op == "sign")
        {
            need(1, 2, 0);
            gpt2::TensorCipher t(in[0]->ct), o;
            gpt2::sign_function(t, o, iparams[0], iparams[1], bootstrapper(), enc, cr, de, ev, gk, rk);
            results.push_back(o.cipher());
        }
        else if (op == "gelu")
        {
            need(1, 0, 0);
            Ciphertext x = in[0]->ct;
            gpt2::compute_gelu(x, r, bootstrapper(), enc, cr, de, ev, gk, rk);
            results.push_back(r);
        }
        else if (op == "exp")
        {
            need(1, 1, 0);
            Ciphertext x = in[0]->ct;
            gpt2::compute_exp(x, r, iparams[0], enc, cr, de, ev, gk, rk);
            results.push_back(r);
        }
        else if (op == "inverse")
        {
            need(1, 1, 0);
            Ciphertext x = in[0]->ct;
            gpt2::compute_inverse(x, r, iparams[0], enc, cr, de, ev, gk, rk);
            results.push_back(r);
        }
        else if (op == "taylor" || op == "inv_sqrt")
        {
            need(1, 1, 1);
            Ciphertext x = in[0]->ct;
            if (op == "taylor")
                gpt2::taylor_expand(x, r, iparams[0], dparams[0], enc, cr, de, ev, gk, rk);
            else
                gpt2::compute_inv_sqrt(x, r, iparams[0], dparams[0], enc, cr, de, ev, gk, rk);
            results.push_back(r);
        }
        else if (op == "layernorm")
        {
            need(1, 1, 2 * iparams[0]);
            Ciphertext x = in[0]->ct;
            const int rs = iparams[0];
            gpt2::compute_layernorm(x, r, vector<double>(dparams, dparams + rs), vector<double>(dparams + rs, dparams + 2 * rs), rs, enc,
                                    cr, de, ev, gk, rk);
            results.push_back(r);
        }
        else if (op == "max")
        {
            need(2, 0, 0);
            Ciphertext a = in[0]->ct, b = in[1]->ct;
            gpt2::computeMax(a, b, r, bootstrapper(), enc, cr, de, ev, gk, rk);
            results.push_back(r);
        }
        else if (op == "quickMax")
        {
            need(1, 1, 0);
            Ciphertext x = in[0]->ct;
            gpt2::quickMax(x, r, iparams[0], bootstrapper(), enc, cr, de, ev, gk, rk);
            results.push_back(r);
        }
        else if (op == "smax" || op == "softmax")
        {
            need(1, 2, 0);
            Ciphertext x = in[0]->ct;
            if (op == "smax")
                gpt2::compute_smax(x, iparams[0], iparams[1], enc, cr, de, ev, gk, rk);
            else
                gpt2::compute_softmax(x, iparams[0], bootstrapper(), enc, cr, de, ev, gk, rk);
            results.push_back(x);
        }
        else if (op == "mask_out")
        {
            need(1, 2, 0);
            Ciphertext x = in[0]->ct;
            gpt2::mask_out(x, r, iparams[0], iparams[1], enc, ev, rk);
            results.push_back(r);
        }
        else if (op == "rotate_inplace" || op == "surefire_rotate")
        {
            need(1, 1, 0);
            Ciphertext x = in[0]->ct;
            if (op == "rotate_inplace")
                gpt2::rotate_inplace(x, iparams[0], ev, gk);
            else
                gpt2::surefire_rotate(x, iparams[0], kg, ev);
            results.push_back(x);
        }
        else if (op == "fake_bootstrap")
        {
            need(1, 0, 0);
            Ciphertext x = in[0]->ct;
            gpt2::fakeBootstrap(x, r, enc, cr, de, ev, gk, rk);
            results.push_back(r);
        }
        else if (op == "bootstrap")
        {
            need(1, 0, 0);
            if (!boot)
                throw std::invalid_argument("bka_gpt2_call(bootstrap): a bootstrapper is required");
            Ciphertext x = in[0]->ct;
            gpt2::bootstrap(x, r, bootstrapper(), ev);
            results.push_back(r);
        }
        else if (op == "init_output")
        {
            need(0, 1, 0);
            gpt2::init_output(iparams[0], results, enc, cr, de, ev, gk, rk);
        }
        else if (op == "pack_from_row")
        {
            need(0, 2, iparams[0] * iparams[1]);
            gpt2::vvec m((std::size_t)iparams[0]);
            for (int i = 0; i < iparams[0]; i++)
                m[(std::size_t)i].assign(dparams + (std::size_t)i * iparams[1], dparams + (std::size_t)(i + 1) * iparams[1]);
            gpt2::pack_from_row(m, results, enc, cr, de, ev, gk, rk);
        }
        else if (op == "expand_bias")
        {
            need(0, 0, 1);
            vector<double> b(dparams, dparams + n_d);
            gpt2::expand_bias(b, r, enc, cr, de, ev, gk, rk);
            results.push_back(r);
        }
        else if (op == "pack_tight" || op == "unpack_tight")
        {
            const int ni = op == "pack_tight" ? 8 : 3, no = op == "pack_tight" ? 3 : 8;
            need(ni + no, 0, 0);
            gpt2::vc a = range(0, ni);
            results = range(ni, no);
            (op == "pack_tight" ? gpt2::pack_tight : gpt2::unpack_tight)(a, results, enc, cr, de, ev, gk, rk);
        }
        else if (op == "row_matmul" || op == "attn_proj_row" || op == "attn_proj_col")
        {
            // iparams: n_left, n_weights, n_outputs, A_rows, A_cols, W_rows, W_cols; in: left.., weights.., bias, outputs..
            need(0, 7, 0);
            const int nl = iparams[0], nw = iparams[1], no = iparams[2];
            need(nl + nw + 1 + no, 7, 0);
            gpt2::vc left = range(0, nl), weights = range(nl, nw);
            Ciphertext bias = in[nl + nw]->ct;
            results = range(nl + nw + 1, no);
            if (op == "row_matmul")
                gpt2::row_matrix_multiplication_seal(left, weights, bias, results, iparams[3], iparams[4], iparams[5], iparams[6], enc, cr,
                                                     de, ev, gk, rk);
            else
                (op == "attn_proj_row" ? gpt2::attn_proj_row_seal : gpt2::attn_proj_col_seal)(
                    left, weights, bias, results, iparams[3], iparams[4], iparams[5], iparams[6], kg, enc, cr, de, ev, gk, rk);
        }
        else if (op == "col_matmul")
        {
            need(0, 2, 0);
            const int rows = iparams[0], cols = iparams[1];
            need(2 * rows, 2, 0);
            vector<gpt2::TensorCipher> left, right, outs;
            for (int i = 0; i < rows; i++)
            {
                left.emplace_back(in[i]->ct);
                right.emplace_back(in[rows + i]->ct);
            }
            gpt2::Config config;
            gpt2::col_matrix_multiplication_seal(left, right, outs, {}, rows, cols, config, enc, cr, de, ev, gk, rk);
            for (auto &t : outs)
                results.push_back(t.cipher());
        }
        else if (op == "qk_matmul" || op == "sv_matmul")
        {
            // iparams: heads, n_outputs; in: first.., second.., outputs..
            need(0, 2, 0);
            const int h = iparams[0], no = iparams[1];
            need(2 * h + no, 2, 0);
            gpt2::vc a = range(0, h), b = range(h, h);
            results = range(2 * h, no);
            (op == "qk_matmul" ? gpt2::qk_matmul : gpt2::sv_matmul)(a, b, results, 128, 768, 768, 128, kg, enc, cr, de, ev, gk, rk);
        }
        else if (op == "augment_row" || op == "augment_col")
        {
            // iparams: n, padded_row_size, idx; in: A.., cached..; out: A..
            need(0, 3, 0);
            const int n = iparams[0];
            need(2 * n, 3, 0);
            results = range(0, n);
            gpt2::vc cached = range(n, n);
            (op == "augment_row" ? gpt2::augment_value_row : gpt2::augment_value_col)(results, cached, iparams[1], iparams[2], enc, cr, de,
                                                                                     ev, gk, rk);
        }
        else
            throw std::invalid_argument("bka_gpt2_call: unknown operator " + op);

        if ((int)results.size() > out_cap)
            throw std::out_of_range("bka_gpt2_call(" + op + "): out_cap too small");
        for (std::size_t i = 0; i < results.size(); i++)
            out[i] = wrap(std::move(results[i]));
        if (n_out)
            *n_out = (int)results.size();
        BKA_END
    }
    int bka_gpt2_init_chain(int *bits_out, int bits_cap, int *n_bits, int *steps_out, int steps_cap, int *n_steps)
    {
        BKA_TRY
        const vector<int> bits = gpt2::init_coeff_bit_vec(), steps = gpt2::init_rotation_steps();
        if ((int)bits.size() > bits_cap || (int)steps.size() > steps_cap)
            throw std::out_of_range("bka_gpt2_init_chain: capacity too small");
        std::copy(bits.begin(), bits.end(), bits_out);
        std::copy(steps.begin(), steps.end(), steps_out);
        *n_bits = (int)bits.size();
        *n_steps = (int)steps.size();
        BKA_END
    }
}
