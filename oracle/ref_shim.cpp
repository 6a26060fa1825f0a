// oracle/ref_shim.cpp - TEST INFRASTRUCTURE ONLY.
//
// A C-ABI shim (written for this repo) that links the reference's own modified SEAL 3.6.6
// (compiled in place by oracle/Makefile into oracle/_ref/libseal_ref.so) and exposes the
// operations of the hot path on RAW limb arrays so that the tests, the golden-vector
// generator and bench.py's cpu_baseline / --impl reference leg can drive the unmodified
// reference through its public API (seal::Evaluator, CKKSEncoder, KeyGenerator, Encryptor,
// Decryptor).  Nothing in the product path (fhe-gpt-2_b200/) may load this library.
//
// Layouts are SEAL's own: ciphertext = [poly][limb][coeff] (ciphertext.h:335-347),
// kswitch key = [digit][poly][key-level limb][coeff] (kswitchkeys.h:340).
#include "seal/seal.h"
#include <chrono>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <omp.h>
#include <string>
#include <vector>

using namespace seal;
using namespace std;

namespace
{
    struct Ref
    {
        unique_ptr<SEALContext> ctx;
        unique_ptr<KeyGenerator> keygen;
        unique_ptr<CKKSEncoder> encoder;
        unique_ptr<Evaluator> evaluator;
        unique_ptr<Encryptor> encryptor;
        unique_ptr<Decryptor> decryptor;
        PublicKey pk;
        RelinKeys rk;
        GaloisKeys gk;
        bool have_rk = false;
        vector<uint32_t> gk_elts;
        vector<parms_id_type> chain; // chain[l-1] = parms_id with l limbs
        map<int, Ciphertext> cts;
        map<int, Plaintext> pts;
        int next_id = 1;
        string err;
    };

    thread_local string g_err;

    parms_id_type pid_for_limbs(Ref *r, int limbs)
    {
        if (limbs < 1 || limbs > (int)r->chain.size())
            throw invalid_argument("limbs out of range");
        return r->chain[limbs - 1];
    }
} // namespace

#define REF_TRY try {
#define REF_CATCH                                                                                                      \
    }                                                                                                                  \
    catch (const exception &e)                                                                                         \
    {                                                                                                                  \
        g_err = e.what();                                                                                              \
        return -1;                                                                                                     \
    }                                                                                                                  \
    return 0;

extern "C"
{
    const char *ref_last_error()
    {
        return g_err.c_str();
    }

    // bits: prime bit sizes incl. the special prime (last). seed==0 -> SEAL's default random source.
    void *ref_create(int log_n, const int *bits, int nbits, int hamming_weight, int sparse_slots, uint64_t seed)
    {
        try
        {
            auto r = make_unique<Ref>();
            EncryptionParameters parms(scheme_type::ckks);
            size_t n = size_t(1) << log_n;
            parms.set_poly_modulus_degree(n);
            parms.set_coeff_modulus(CoeffModulus::Create(n, vector<int>(bits, bits + nbits)));
            if (hamming_weight)
                parms.set_secret_key_hamming_weight(size_t(hamming_weight));
            if (sparse_slots)
                parms.set_sparse_slots(size_t(sparse_slots));
            if (seed)
            {
                prng_seed_type s{};
                for (size_t i = 0; i < s.size(); i++)
                    s[i] = seed * 0x9E3779B97F4A7C15ULL + i;
                parms.set_random_generator(make_shared<Blake2xbPRNGFactory>(s));
            }
            r->ctx = make_unique<SEALContext>(parms, true, sec_level_type::none);
            if (!r->ctx->parameters_set())
                throw invalid_argument(string("parameters not set: ") + r->ctx->parameter_error_message());
            r->keygen = make_unique<KeyGenerator>(*r->ctx);
            r->keygen->create_public_key(r->pk);
            r->encoder = make_unique<CKKSEncoder>(*r->ctx);
            r->evaluator = make_unique<Evaluator>(*r->ctx, *r->encoder);
            r->encryptor = make_unique<Encryptor>(*r->ctx, r->pk);
            r->decryptor = make_unique<Decryptor>(*r->ctx, r->keygen->secret_key());
            // chain
            size_t top = r->ctx->first_context_data()->parms().coeff_modulus().size();
            r->chain.resize(top);
            for (auto cd = r->ctx->first_context_data(); cd; cd = cd->next_context_data())
                r->chain[cd->parms().coeff_modulus().size() - 1] = cd->parms_id();
            return r.release();
        }
        catch (const exception &e)
        {
            g_err = e.what();
            return nullptr;
        }
    }

    void ref_destroy(void *h)
    {
        delete static_cast<Ref *>(h);
    }

    int ref_n_primes(void *h)
    {
        auto r = static_cast<Ref *>(h);
        return (int)r->ctx->key_context_data()->parms().coeff_modulus().size();
    }

    int ref_get_primes(void *h, uint64_t *out)
    {
        auto r = static_cast<Ref *>(h);
        auto &m = r->ctx->key_context_data()->parms().coeff_modulus();
        for (size_t i = 0; i < m.size(); i++)
            out[i] = m[i].value();
        return 0;
    }

    // raw NTT of one limb with the key-level tables (ntt.h:235-264 / :336-358)
    int ref_ntt(void *h, int prime_idx, uint64_t *data, int inverse)
    {
        REF_TRY
        auto r = static_cast<Ref *>(h);
        auto tables = r->ctx->key_context_data()->small_ntt_tables();
        if (inverse)
            util::inverse_ntt_negacyclic_harvey(data, tables[prime_idx]);
        else
            util::ntt_negacyclic_harvey(data, tables[prime_idx]);
        REF_CATCH
    }

    // root_powers (operand only) of prime idx, SEAL's order (ntt.cpp:58-67); inverse!=0 -> inv_root_powers
    int ref_root_powers(void *h, int prime_idx, uint64_t *out, int inverse)
    {
        REF_TRY
        auto r = static_cast<Ref *>(h);
        auto &t = r->ctx->key_context_data()->small_ntt_tables()[prime_idx];
        size_t n = t.coeff_count();
        for (size_t i = 0; i < n; i++)
            out[i] = inverse ? t.get_from_inv_root_powers(i).operand : t.get_from_root_powers(i).operand;
        REF_CATCH
    }

    int ref_galois_elt_from_step(void *h, int step, uint32_t *elt)
    {
        REF_TRY
        auto r = static_cast<Ref *>(h);
        *elt = r->ctx->key_context_data()->galois_tool()->get_elt_from_step(step);
        REF_CATCH
    }

    // out[limb][coeff]: result of apply_galois_ntt's permutation table (galois.cpp:18-51)
    int ref_apply_galois_ntt(void *h, uint32_t elt, const uint64_t *in, uint64_t *out)
    {
        REF_TRY
        auto r = static_cast<Ref *>(h);
        auto gt = r->ctx->key_context_data()->galois_tool();
        size_t n = r->ctx->key_context_data()->parms().poly_modulus_degree();
        int log_n = util::get_power_of_two(n);
        gt->apply_galois_ntt(util::ConstCoeffIter(in), elt, util::CoeffIter(out));
        (void)log_n;
        REF_CATCH
    }

    // ---- keys (raw) ----
    int ref_get_secret_key(void *h, uint64_t *out) // [key limbs][N], NTT form
    {
        REF_TRY
        auto r = static_cast<Ref *>(h);
        auto &sk = r->keygen->secret_key();
        memcpy(out, sk.data().data(), sk.data().coeff_count() * sizeof(uint64_t));
        REF_CATCH
    }

    int ref_get_public_key(void *h, uint64_t *out) // [2][key limbs][N]
    {
        REF_TRY
        auto r = static_cast<Ref *>(h);
        auto &c = r->pk.data();
        memcpy(out, c.data(), c.size() * c.coeff_modulus_size() * c.poly_modulus_degree() * sizeof(uint64_t));
        REF_CATCH
    }

    static void dump_kswitch(const vector<PublicKey> &kv, uint64_t *out)
    {
        size_t off = 0;
        for (auto &pk : kv)
        {
            auto &c = pk.data();
            size_t cnt = c.size() * c.coeff_modulus_size() * c.poly_modulus_degree();
            memcpy(out + off, c.data(), cnt * sizeof(uint64_t));
            off += cnt;
        }
    }

    int ref_make_relin_key(void *h, uint64_t *out) // [digits][2][key limbs][N]; out may be NULL
    {
        REF_TRY
        auto r = static_cast<Ref *>(h);
        if (!r->have_rk)
        {
            r->keygen->create_relin_keys(r->rk);
            r->have_rk = true;
        }
        if (out)
            dump_kswitch(r->rk.data()[0], out);
        REF_CATCH
    }

    // (Re)creates the Galois key set for exactly these steps (step 0 = conjugation, as in
    // galois.cpp:53-95) - previous Galois keys are discarded.
    int ref_make_galois_keys(void *h, const int *steps, int nsteps)
    {
        REF_TRY
        auto r = static_cast<Ref *>(h);
        auto gt = r->ctx->key_context_data()->galois_tool();
        r->gk_elts.clear();
        for (int i = 0; i < nsteps; i++)
            r->gk_elts.push_back(gt->get_elt_from_step(steps[i]));
        r->keygen->create_galois_keys(r->gk_elts, r->gk);
        REF_CATCH
    }

    int ref_get_galois_key(void *h, uint32_t elt, uint64_t *out)
    {
        REF_TRY
        auto r = static_cast<Ref *>(h);
        if (!r->gk.has_key(elt))
            throw invalid_argument("Galois key not present");
        dump_kswitch(r->gk.key(elt), out);
        REF_CATCH
    }

    // ---- ciphertext / plaintext registry ----
    int ref_ct_new(void *h)
    {
        auto r = static_cast<Ref *>(h);
        int id = r->next_id++;
        r->cts[id];
        return id;
    }
    int ref_ct_free(void *h, int id)
    {
        static_cast<Ref *>(h)->cts.erase(id);
        return 0;
    }
    int ref_ct_info(void *h, int id, int *size, int *limbs, double *scale, int *is_ntt)
    {
        REF_TRY
        auto &c = static_cast<Ref *>(h)->cts.at(id);
        *size = (int)c.size();
        *limbs = (int)c.coeff_modulus_size();
        *scale = c.scale();
        *is_ntt = c.is_ntt_form();
        REF_CATCH
    }
    int ref_ct_get(void *h, int id, uint64_t *out)
    {
        REF_TRY
        auto &c = static_cast<Ref *>(h)->cts.at(id);
        memcpy(out, c.data(), c.size() * c.coeff_modulus_size() * c.poly_modulus_degree() * sizeof(uint64_t));
        REF_CATCH
    }
    int ref_ct_set(void *h, int id, const uint64_t *data, int size, int limbs, double scale, int is_ntt)
    {
        REF_TRY
        auto r = static_cast<Ref *>(h);
        auto &c = r->cts.at(id);
        c.resize(*r->ctx, pid_for_limbs(r, limbs), size_t(size));
        c.is_ntt_form() = is_ntt != 0;
        c.scale() = scale;
        memcpy(c.data(), data, size_t(size) * limbs * c.poly_modulus_degree() * sizeof(uint64_t));
        REF_CATCH
    }
    int ref_ct_set_scale(void *h, int id, double scale)
    {
        REF_TRY
        static_cast<Ref *>(h)->cts.at(id).scale() = scale;
        REF_CATCH
    }
    int ref_ct_copy(void *h, int dst, int src)
    {
        REF_TRY
        auto r = static_cast<Ref *>(h);
        r->cts.at(dst) = r->cts.at(src);
        REF_CATCH
    }

    int ref_pt_new(void *h)
    {
        auto r = static_cast<Ref *>(h);
        int id = r->next_id++;
        r->pts[id];
        return id;
    }
    int ref_pt_free(void *h, int id)
    {
        static_cast<Ref *>(h)->pts.erase(id);
        return 0;
    }
    int ref_pt_info(void *h, int id, int *limbs, double *scale)
    {
        REF_TRY
        auto r = static_cast<Ref *>(h);
        auto &p = r->pts.at(id);
        *limbs = (int)r->ctx->get_context_data(p.parms_id())->parms().coeff_modulus().size();
        *scale = p.scale();
        REF_CATCH
    }
    int ref_pt_get(void *h, int id, uint64_t *out)
    {
        REF_TRY
        auto &p = static_cast<Ref *>(h)->pts.at(id);
        memcpy(out, p.data(), p.coeff_count() * sizeof(uint64_t));
        REF_CATCH
    }

    // ---- encoder / encryptor / decryptor ----
    // values: n complex (re,im interleaved) if is_complex else n doubles. Encoded at `limbs` limbs
    // via encode(values, parms_id, scale, pt) (ckks.h:179-184 -> :457-638).
    int ref_encode(void *h, int pt, const double *values, int n, int is_complex, int limbs, double scale)
    {
        REF_TRY
        auto r = static_cast<Ref *>(h);
        auto pid = pid_for_limbs(r, limbs);
        if (is_complex)
        {
            vector<complex<double>> v(n);
            for (int i = 0; i < n; i++)
                v[i] = complex<double>(values[2 * i], values[2 * i + 1]);
            r->encoder->encode(v, pid, scale, r->pts.at(pt));
        }
        else
        {
            vector<double> v(values, values + n);
            r->encoder->encode(v, pid, scale, r->pts.at(pt));
        }
        REF_CATCH
    }
    int ref_encode_scalar(void *h, int pt, double value, int limbs, double scale)
    {
        REF_TRY
        auto r = static_cast<Ref *>(h);
        r->encoder->encode(value, pid_for_limbs(r, limbs), scale, r->pts.at(pt));
        REF_CATCH
    }
    int ref_decode(void *h, int pt, double *out) // slot_count complex values
    {
        REF_TRY
        auto r = static_cast<Ref *>(h);
        vector<complex<double>> v;
        r->encoder->decode(r->pts.at(pt), v);
        for (size_t i = 0; i < v.size(); i++)
        {
            out[2 * i] = v[i].real();
            out[2 * i + 1] = v[i].imag();
        }
        REF_CATCH
    }
    int ref_encrypt(void *h, int pt, int ct)
    {
        REF_TRY
        auto r = static_cast<Ref *>(h);
        r->encryptor->encrypt(r->pts.at(pt), r->cts.at(ct));
        REF_CATCH
    }
    int ref_decrypt(void *h, int ct, int pt)
    {
        REF_TRY
        auto r = static_cast<Ref *>(h);
        r->decryptor->decrypt(r->cts.at(ct), r->pts.at(pt));
        REF_CATCH
    }

    // ---- evaluator ops (in place on ct `a`; `b` second operand where needed) ----
    enum
    {
        OP_ADD = 1,
        OP_SUB = 2,
        OP_MULTIPLY = 3,
        OP_SQUARE = 4,
        OP_RELINEARIZE = 5,
        OP_RESCALE = 6,
        OP_MOD_SWITCH_NEXT = 7,
        OP_ROTATE = 8,
        OP_CONJUGATE = 9,
        OP_NEGATE = 10,
        OP_ADD_PLAIN = 11,
        OP_SUB_PLAIN = 12,
        OP_MULTIPLY_PLAIN = 13,
        OP_ADD_CONST = 14,
        OP_MULTIPLY_CONST = 15,
        OP_ADD_REDUCED_ERROR = 16,
        OP_SUB_REDUCED_ERROR = 17,
        OP_MULTIPLY_REDUCED_ERROR = 18,
        OP_NTT_FWD = 19,
        OP_NTT_INV = 20,
        OP_MOD_SWITCH_TO = 21,
        OP_MULTIPLY_VECTOR = 22  // darg = the constant every slot holds
    };

    static void do_op(Ref *r, int op, Ciphertext &a, int b, int iarg, double darg)
    {
        auto &ev = *r->evaluator;
        switch (op)
        {
        case OP_ADD:
            ev.add_inplace(a, r->cts.at(b));
            break;
        case OP_SUB:
            ev.sub_inplace(a, r->cts.at(b));
            break;
        case OP_MULTIPLY:
            ev.multiply_inplace(a, r->cts.at(b));
            break;
        case OP_SQUARE:
            ev.square_inplace(a);
            break;
        case OP_RELINEARIZE:
            ev.relinearize_inplace(a, r->rk);
            break;
        case OP_RESCALE:
            ev.rescale_to_next_inplace(a);
            break;
        case OP_MOD_SWITCH_NEXT:
            ev.mod_switch_to_next_inplace(a);
            break;
        case OP_ROTATE:
            ev.rotate_vector_inplace(a, iarg, r->gk);
            break;
        case OP_CONJUGATE:
            ev.complex_conjugate_inplace(a, r->gk);
            break;
        case OP_NEGATE:
            ev.negate_inplace(a);
            break;
        case OP_ADD_PLAIN:
            ev.add_plain_inplace(a, r->pts.at(b));
            break;
        case OP_SUB_PLAIN:
            ev.sub_plain_inplace(a, r->pts.at(b));
            break;
        case OP_MULTIPLY_PLAIN:
            ev.multiply_plain_inplace(a, r->pts.at(b));
            break;
        case OP_ADD_CONST:
            ev.add_const_inplace(a, darg);
            break;
        case OP_MULTIPLY_CONST:
            ev.multiply_const_inplace(a, darg);
            break;
        case OP_ADD_REDUCED_ERROR:
            ev.add_inplace_reduced_error(a, r->cts.at(b));
            break;
        case OP_SUB_REDUCED_ERROR:
            ev.sub_inplace_reduced_error(a, r->cts.at(b));
            break;
        case OP_MULTIPLY_REDUCED_ERROR:
            ev.multiply_inplace_reduced_error(a, r->cts.at(b), r->rk);
            break;
        case OP_NTT_FWD:
            ev.transform_to_ntt_inplace(a);
            break;
        case OP_NTT_INV:
            ev.transform_from_ntt_inplace(a);
            break;
        case OP_MOD_SWITCH_TO:
            ev.mod_switch_to_inplace(a, pid_for_limbs(r, iarg));
            break;
        case OP_MULTIPLY_VECTOR:
        {
            // evaluator.h:1270-1278 with a constant slot vector: encode at the top level, drop, multiply_plain
            vector<double> v(r->encoder->slot_count(), darg);
            r->evaluator->multiply_vector_inplace_reduced_error(a, v);
            break;
        }
        default:
            throw invalid_argument("unknown op");
        }
    }

    int ref_op(void *h, int op, int a, int b, int iarg, double darg)
    {
        REF_TRY
        auto r = static_cast<Ref *>(h);
        do_op(r, op, r->cts.at(a), b, iarg, darg);
        REF_CATCH
    }

    // multiply_vector_inplace_reduced_error (evaluator.h:1270-1278): encode at top level, drop, multiply
    int ref_multiply_vector_reduced_error(void *h, int a, const double *values, int n, int is_complex)
    {
        REF_TRY
        auto r = static_cast<Ref *>(h);
        if (is_complex)
        {
            vector<complex<double>> v(n);
            for (int i = 0; i < n; i++)
                v[i] = complex<double>(values[2 * i], values[2 * i + 1]);
            r->evaluator->multiply_vector_inplace_reduced_error(r->cts.at(a), v);
        }
        else
        {
            vector<double> v(values, values + n);
            r->evaluator->multiply_vector_inplace_reduced_error(r->cts.at(a), v);
        }
        REF_CATCH
    }

    // CPU baseline: time `reps` applications of `op` on `threads` independent copies of ct `a`
    // (the reference's own parallelisation style: one ciphertext per OpenMP thread,
    // infer_seal.cpp:404). Every application starts from a fresh copy of `a` (copy excluded from
    // the per-thread timer). Returns seconds of wall time for the whole batch in *wall_s and the
    // mean per-op single-thread latency in *mean_op_s.
    int ref_time_op(void *h, int op, int a, int b, int iarg, double darg, int threads, int reps, double *wall_s,
                    double *mean_op_s)
    {
        REF_TRY
        auto r = static_cast<Ref *>(h);
        const Ciphertext &src = r->cts.at(a);
        vector<Ciphertext> work(threads);
        vector<double> acc(threads, 0.0);
        auto t0 = chrono::steady_clock::now();
#pragma omp parallel for num_threads(threads) schedule(static)
        for (int t = 0; t < threads; t++)
        {
            for (int k = 0; k < reps; k++)
            {
                work[t] = src;
                auto s = chrono::steady_clock::now();
                do_op(r, op, work[t], b, iarg, darg);
                acc[t] += chrono::duration<double>(chrono::steady_clock::now() - s).count();
            }
        }
        *wall_s = chrono::duration<double>(chrono::steady_clock::now() - t0).count();
        double s = 0;
        for (double x : acc)
            s += x;
        *mean_op_s = s / (double(threads) * reps);
        REF_CATCH
    }

    // ---- SEAL's own serialization (save / load members of the reference's classes), through files.
    // what: 0 ciphertext `id`, 1 plaintext `id`, 2 relinearization keys, 3 Galois keys, 4 secret key, 5 public key
    int ref_save(void *h, int what, int id, const char *path)
    {
        REF_TRY
        Ref *r = (Ref *)h;
        std::ofstream f(path, std::ios::binary);
        if (!f)
            throw runtime_error("cannot open file");
        switch (what)
        {
        case 0: r->cts.at(id).save(f, compr_mode_type::none); break;
        case 1: r->pts.at(id).save(f, compr_mode_type::none); break;
        case 2:
            if (!r->have_rk)
            {
                r->keygen->create_relin_keys(r->rk);
                r->have_rk = true;
            }
            r->rk.save(f, compr_mode_type::none);
            break;
        case 3: r->gk.save(f, compr_mode_type::none); break;
        case 4: r->keygen->secret_key().save(f, compr_mode_type::none); break;
        case 5: r->pk.save(f, compr_mode_type::none); break;
        default: throw invalid_argument("what");
        }
        REF_CATCH
    }
    int ref_load(void *h, int what, int id, const char *path)
    {
        REF_TRY
        Ref *r = (Ref *)h;
        std::ifstream f(path, std::ios::binary);
        if (!f)
            throw runtime_error("cannot open file");
        switch (what)
        {
        case 0: r->cts[id].load(*r->ctx, f); break;
        case 1: r->pts[id].load(*r->ctx, f); break;
        case 2:
            r->rk.load(*r->ctx, f);
            r->have_rk = true;
            break;
        case 3: r->gk.load(*r->ctx, f); break;
        default: throw invalid_argument("what");
        }
        REF_CATCH
    }

    int ref_max_threads()
    {
        return omp_get_max_threads();
    }
}
