// seal/serialization.h - SEAL 3.6's binary wire format for the facade's objects (included by seal/seal.h).
//
// Replaces, for the objects the reference's applications exchange (gpt2_ckks/network: parameters, keys, ciphertexts):
//   serialization.{h,cpp}   Serialization::SEALHeader + Save / Load (compr_mode_type::none only - the reference's
//                           SEAL is built without zlib / zstd here)
//   ciphertext.cpp:183-239  Ciphertext::save_members,  :241-360 load_members
//   plaintext.cpp:204-229   Plaintext::save_members,   :231-300 load_members
//   kswitchkeys.cpp:42-84   KSwitchKeys::save_members, :86-145  load_members   (RelinKeys, GaloisKeys)
//   publickey.h:106-151, secretkey.h:131-190, util/dynarray.h:652-720 (DynArray), encryptionparams.cpp:124-158
//   (parms_id = BLAKE2b-256 over {scheme, N, q_0 .. q_k-1, plain_modulus})
//
// Layout (little endian):  SEALHeader{magic 0xA15E, header size 0x10, version 3.6, compr_mode, reserved, total bytes}
// followed by the object's members; a DynArray is again a SEALHeader + element count + raw words.  Objects written
// here load in the reference's SEAL and vice versa (tests/test_serialization_gpu.py, byte-identical round trips).
// Seed-compressed objects (a half replaced by a PRNG seed, ciphertext.cpp:202-221) are not written, and reading one
// throws std::logic_error: expanding it needs the reference's Blake2xb / SHAKE generators.
#pragma once
#include <istream>
#include <ostream>

namespace seal
{
    namespace detail
    {
        // ---- BLAKE2b (RFC 7693), unkeyed, digest length <= 64 -------------------------------------------------
        inline void blake2b(const void *in, std::size_t inlen, void *out, std::size_t outlen)
        {
            static const std::uint64_t iv[8] = { 0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull,
                                                 0xa54ff53a5f1d36f1ull, 0x510e527fade682d1ull, 0x9b05688c2b3e6c1full,
                                                 0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull };
            static const unsigned char sigma[12][16] = {
                { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15 }, { 14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3 },
                { 11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4 }, { 7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8 },
                { 9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13 }, { 2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9 },
                { 12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11 }, { 13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10 },
                { 6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5 }, { 10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0 },
                { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15 }, { 14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3 }
            };
            std::uint64_t h[8];
            for (int i = 0; i < 8; i++)
                h[i] = iv[i];
            h[0] ^= 0x01010000ull ^ (std::uint64_t)outlen;
            auto rotr = [](std::uint64_t x, int r) { return (x >> r) | (x << (64 - r)); };
            auto compress = [&](const unsigned char *block, std::uint64_t t, bool last) {
                std::uint64_t m[16], v[16];
                for (int i = 0; i < 16; i++)
                    std::memcpy(&m[i], block + 8 * i, 8);
                for (int i = 0; i < 8; i++)
                {
                    v[i] = h[i];
                    v[i + 8] = iv[i];
                }
                v[12] ^= t;
                if (last)
                    v[14] = ~v[14];
                for (int r = 0; r < 12; r++)
                {
                    const unsigned char *s = sigma[r];
                    auto G = [&](int a, int b, int c, int d, std::uint64_t x, std::uint64_t y) {
                        v[a] = v[a] + v[b] + x;
                        v[d] = rotr(v[d] ^ v[a], 32);
                        v[c] = v[c] + v[d];
                        v[b] = rotr(v[b] ^ v[c], 24);
                        v[a] = v[a] + v[b] + y;
                        v[d] = rotr(v[d] ^ v[a], 16);
                        v[c] = v[c] + v[d];
                        v[b] = rotr(v[b] ^ v[c], 63);
                    };
                    G(0, 4, 8, 12, m[s[0]], m[s[1]]);
                    G(1, 5, 9, 13, m[s[2]], m[s[3]]);
                    G(2, 6, 10, 14, m[s[4]], m[s[5]]);
                    G(3, 7, 11, 15, m[s[6]], m[s[7]]);
                    G(0, 5, 10, 15, m[s[8]], m[s[9]]);
                    G(1, 6, 11, 12, m[s[10]], m[s[11]]);
                    G(2, 7, 8, 13, m[s[12]], m[s[13]]);
                    G(3, 4, 9, 14, m[s[14]], m[s[15]]);
                }
                for (int i = 0; i < 8; i++)
                    h[i] ^= v[i] ^ v[i + 8];
            };
            const unsigned char *p = static_cast<const unsigned char *>(in);
            std::uint64_t t = 0;
            while (inlen > 128)
            {
                t += 128;
                compress(p, t, false);
                p += 128;
                inlen -= 128;
            }
            unsigned char last[128] = { 0 };
            std::memcpy(last, p, inlen);
            t += inlen;
            compress(last, t, true);
            std::memcpy(out, h, outlen);
        }

        // EncryptionParameters::compute_parms_id (encryptionparams.cpp:124-158) of the chain element with `limbs`
        // primes (limbs == n_primes: the key level)
        inline std::array<std::uint64_t, 4> seal_parms_id(const CtxImpl &ctx, int limbs)
        {
            std::vector<std::uint64_t> words = { (std::uint64_t)scheme_type::ckks, std::uint64_t(1) << ctx->log_n };
            const auto &q = ctx->parms.coeff_modulus();
            for (int i = 0; i < limbs; i++)
                words.push_back(q[(std::size_t)i].value());
            words.push_back(0); // plain_modulus of CKKS
            std::array<std::uint64_t, 4> id{};
            blake2b(words.data(), words.size() * 8, id.data(), 32);
            return id;
        }
        // level (limb count) of a stored parms_id; throws like is_metadata_valid_for (valcheck.cpp) when it is foreign
        inline int limbs_of_seal_id(const CtxImpl &ctx, const std::array<std::uint64_t, 4> &id, bool allow_key_level)
        {
            for (int l = 1; l <= ctx->n_primes - (allow_key_level ? 0 : 1); l++)
                if (seal_parms_id(ctx, l) == id)
                    return l;
            throw std::logic_error("data is not valid for the encryption parameters (unknown parms_id)");
        }

        struct SEALHeader
        {
            std::uint16_t magic = 0xA15E;
            std::uint8_t header_size = 0x10;
            std::uint8_t version_major = 3;
            std::uint8_t version_minor = 6;
            std::uint8_t compr_mode = 0;
            std::uint16_t reserved = 0;
            std::uint64_t size = 0;
        };
        static_assert(sizeof(SEALHeader) == 16, "SEALHeader is 16 bytes");

        template <class T>
        inline void put(std::ostream &s, const T &v)
        {
            s.write(reinterpret_cast<const char *>(&v), sizeof(T));
            if (!s)
                throw std::runtime_error("I/O error");
        }
        template <class T>
        inline T get(std::istream &s)
        {
            T v{};
            s.read(reinterpret_cast<char *>(&v), sizeof(T));
            if (!s)
                throw std::runtime_error("I/O error");
            return v;
        }
        inline void put_header(std::ostream &s, std::uint64_t total)
        {
            SEALHeader h;
            h.size = total;
            put(s, h);
        }
        // returns the object's total size (header included)
        inline std::uint64_t get_header(std::istream &s)
        {
            SEALHeader h = get<SEALHeader>(s);
            if (h.magic != 0xA15E || h.header_size != 0x10)
                throw std::logic_error("loaded SEALHeader is invalid");
            if (h.version_major != 3 || h.version_minor < 4)
                throw std::logic_error("incompatible version");
            if (h.compr_mode != 0)
                throw std::logic_error("unsupported compression mode (this build reads compr_mode_type::none only)");
            return h.size;
        }
        // DynArray<uint64_t>::save / load (util/dynarray.h:652-720)
        inline std::uint64_t dynarray_bytes(std::uint64_t words)
        {
            return 16 + 8 + 8 * words;
        }
        inline void put_dynarray(std::ostream &s, const std::uint64_t *data, std::uint64_t words)
        {
            put_header(s, dynarray_bytes(words));
            put(s, words);
            s.write(reinterpret_cast<const char *>(data), (std::streamsize)(8 * words));
            if (!s)
                throw std::runtime_error("I/O error");
        }
        inline std::vector<std::uint64_t> get_dynarray(std::istream &s, std::uint64_t bound)
        {
            std::uint64_t total = get_header(s);
            std::uint64_t words = get<std::uint64_t>(s);
            if (words > bound || total != dynarray_bytes(words))
                throw std::logic_error("unexpected size");
            std::vector<std::uint64_t> v((std::size_t)words);
            s.read(reinterpret_cast<char *>(v.data()), (std::streamsize)(8 * words));
            if (!s)
                throw std::runtime_error("I/O error");
            return v;
        }
        inline void need_none(compr_mode_type m)
        {
            if (m != compr_mode_type::none)
                throw std::invalid_argument("unsupported compression mode");
        }

        // Ciphertext members (ciphertext.cpp:183-239) for raw words in SEAL's layout
        inline std::uint64_t ct_members_bytes(std::uint64_t words)
        {
            return 32 + 1 + 8 + 8 + 8 + 8 + dynarray_bytes(words);
        }
        inline void put_ct_members(std::ostream &s, const std::array<std::uint64_t, 4> &id, bool ntt, std::uint64_t size, std::uint64_t n,
                                   std::uint64_t limbs, double scale, const std::uint64_t *data)
        {
            put(s, id);
            put(s, (std::uint8_t)(ntt ? 1 : 0));
            put(s, size);
            put(s, n);
            put(s, limbs);
            put(s, scale);
            put_dynarray(s, data, size * n * limbs);
        }
        struct CtMembers
        {
            std::array<std::uint64_t, 4> id;
            bool ntt;
            std::uint64_t size, n, limbs;
            double scale;
            std::vector<std::uint64_t> data;
        };
        inline CtMembers get_ct_members(std::istream &s, const CtxImpl &ctx)
        {
            CtMembers m;
            m.id = get<std::array<std::uint64_t, 4>>(s);
            m.ntt = get<std::uint8_t>(s) != 0;
            m.size = get<std::uint64_t>(s);
            m.n = get<std::uint64_t>(s);
            m.limbs = get<std::uint64_t>(s);
            m.scale = get<double>(s);
            if (m.n != (std::uint64_t(1) << ctx->log_n) || m.limbs < 1 || m.limbs > (std::uint64_t)ctx->n_primes || m.size > 16)
                throw std::logic_error("ciphertext data is invalid");
            if (seal_parms_id(ctx, (int)m.limbs) != m.id)
                throw std::logic_error("ciphertext data is invalid");
            m.data = get_dynarray(s, m.size * m.n * m.limbs);
            if (m.data.size() != m.size * m.n * m.limbs)
            {
                if (m.size >= 2 && m.data.size() == m.n * m.limbs)
                    throw std::logic_error("seed-compressed ciphertexts are not supported (expanding the seed needs the "
                                           "reference's Blake2xb / SHAKE generator)");
                throw std::logic_error("ciphertext data is invalid");
            }
            return m;
        }
    } // namespace detail

    // ------------------------------------------------------------------------------------------------ Ciphertext
    inline std::streamoff Ciphertext::save_size(compr_mode_type m) const
    {
        detail::need_none(m);
        return (std::streamoff)(16 + detail::ct_members_bytes((std::uint64_t)size_ * limbs_ * poly_modulus_degree()));
    }
    inline std::streamoff Ciphertext::save(std::ostream &stream, compr_mode_type m) const
    {
        detail::need_none(m);
        if (!h_ || !size_)
        { // an empty ciphertext: parms_id_zero, no data (ciphertext.h default state)
            detail::put_header(stream, 16 + detail::ct_members_bytes(0));
            detail::put_ct_members(stream, parms_id_zero, ntt_, 0, 0, 0, scale_, nullptr);
            return (std::streamoff)(16 + detail::ct_members_bytes(0));
        }
        std::vector<std::uint64_t> words((std::size_t)size_ * limbs_ * poly_modulus_degree());
        download(words.data());
        const std::uint64_t total = 16 + detail::ct_members_bytes(words.size());
        detail::put_header(stream, total);
        detail::put_ct_members(stream, detail::seal_parms_id(ctx_, limbs_), ntt_, (std::uint64_t)size_, poly_modulus_degree(),
                               (std::uint64_t)limbs_, scale_, words.data());
        return (std::streamoff)total;
    }
    inline std::streamoff Ciphertext::unsafe_load(const SEALContext &context, std::istream &stream)
    {
        const std::uint64_t total = detail::get_header(stream);
        detail::CtMembers m = detail::get_ct_members(stream, context.impl());
        if (total != 16 + detail::ct_members_bytes(m.data.size()))
            throw std::logic_error("unexpected size");
        if (m.size == 0)
            throw std::logic_error("ciphertext data is invalid");
        upload(context, m.data.data(), (int)m.size, (int)m.limbs, m.scale, m.ntt);
        return (std::streamoff)total;
    }
    inline std::streamoff Ciphertext::load(const SEALContext &context, std::istream &stream)
    {
        Ciphertext fresh;
        std::streamoff n = fresh.unsafe_load(context, stream);
        if (fresh.coeff_modulus_size() >= (std::size_t)context.impl()->n_primes)
            throw std::logic_error("ciphertext data is invalid"); // key level is for keys only (is_valid_for)
        *this = std::move(fresh);
        return n;
    }

    // ------------------------------------------------------------------------------------------------- Plaintext
    // CKKS plaintexts are in NTT form: parms_id of their level, coeff_count = N * limbs (plaintext.cpp:204-229)
    inline std::streamoff Plaintext::save(std::ostream &stream, compr_mode_type m) const
    {
        detail::need_none(m);
        if (!h_ || !limbs_)
            throw std::logic_error("plaintext is empty");
        std::vector<std::uint64_t> words(coeff_count());
        detail::check(bk_pt_download(h_, words.data()));
        const std::uint64_t total = 16 + 32 + 8 + 8 + detail::dynarray_bytes(words.size());
        detail::put_header(stream, total);
        detail::put(stream, detail::seal_parms_id(ctx_, limbs_));
        detail::put(stream, (std::uint64_t)words.size());
        detail::put(stream, scale_);
        detail::put_dynarray(stream, words.data(), words.size());
        return (std::streamoff)total;
    }
    inline std::streamoff Plaintext::load(const SEALContext &context, std::istream &stream)
    {
        const std::uint64_t total = detail::get_header(stream);
        auto id = detail::get<std::array<std::uint64_t, 4>>(stream);
        std::uint64_t count = detail::get<std::uint64_t>(stream);
        double scale = detail::get<double>(stream);
        if (id == parms_id_zero)
            throw std::logic_error("plaintexts that are not in NTT form belong to BFV; this engine evaluates CKKS only");
        int limbs = detail::limbs_of_seal_id(context.impl(), id, false);
        std::vector<std::uint64_t> words = detail::get_dynarray(stream, count);
        if (count != ((std::uint64_t)limbs << context.impl()->log_n) || words.size() != count ||
            total != 16 + 32 + 8 + 8 + detail::dynarray_bytes(count))
            throw std::logic_error("plaintext data is invalid");
        bind(context.impl());
        detail::check(bk_pt_upload(h_, words.data(), limbs, scale));
        pull();
        return (std::streamoff)total;
    }

    // ------------------------------------------------------------------------------------------- PublicKey, SecretKey
    inline std::streamoff PublicKey::save(std::ostream &stream, compr_mode_type m) const
    {
        return ct_.save(stream, m); // publickey.h:106-110
    }
    inline std::streamoff PublicKey::load(const SEALContext &context, std::istream &stream)
    {
        Ciphertext fresh;
        std::streamoff n = fresh.unsafe_load(context, stream);
        if (fresh.coeff_modulus_size() != (std::size_t)context.impl()->n_primes || fresh.size() != 2)
            throw std::logic_error("PublicKey data is invalid");
        ct_ = std::move(fresh);
        return n;
    }
    // secretkey.h:131-190: a Plaintext at the key level holding s in NTT form
    inline std::streamoff SecretKey::save(std::ostream &stream, compr_mode_type m) const
    {
        detail::need_none(m);
        if (!sk_)
            throw std::logic_error("secret key is empty");
        const CtxImpl &ctx = sk_->ctx;
        std::vector<std::uint64_t> words((std::size_t)ctx->n_primes << ctx->log_n);
        download(words.data());
        const std::uint64_t total = 16 + 32 + 8 + 8 + detail::dynarray_bytes(words.size());
        detail::put_header(stream, total);
        detail::put(stream, detail::seal_parms_id(ctx, ctx->n_primes));
        detail::put(stream, (std::uint64_t)words.size());
        detail::put(stream, 1.0);
        detail::put_dynarray(stream, words.data(), words.size());
        return (std::streamoff)total;
    }
    inline std::streamoff SecretKey::load(const SEALContext &context, std::istream &stream)
    {
        const CtxImpl &ctx = context.impl();
        const std::uint64_t total = detail::get_header(stream);
        auto id = detail::get<std::array<std::uint64_t, 4>>(stream);
        std::uint64_t count = detail::get<std::uint64_t>(stream);
        (void)detail::get<double>(stream);
        std::vector<std::uint64_t> words = detail::get_dynarray(stream, count);
        if (id != detail::seal_parms_id(ctx, ctx->n_primes) || count != ((std::uint64_t)ctx->n_primes << ctx->log_n) ||
            words.size() != count || total != 16 + 32 + 8 + 8 + detail::dynarray_bytes(count))
            throw std::logic_error("SecretKey data is invalid");
        *this = SecretKey::upload(context, words.data());
        return (std::streamoff)total;
    }

    // ------------------------------------------------------------------------------ KSwitchKeys (RelinKeys, GaloisKeys)
    namespace detail
    {
        // one key_vector: `digits` PublicKeys, each a size-2 ciphertext at the key level.  words: SEAL's layout
        // [digits][2][n_primes][N]
        inline std::uint64_t kswitch_entry_bytes(const CtxImpl &ctx, int digits)
        {
            const std::uint64_t per_key = 16 + ct_members_bytes((std::uint64_t)2 * ctx->n_primes << ctx->log_n);
            return 8 + (std::uint64_t)digits * per_key;
        }
        inline void put_kswitch_entry(std::ostream &s, const CtxImpl &ctx, const std::uint64_t *words, int digits)
        {
            const std::uint64_t n = std::uint64_t(1) << ctx->log_n, per = (std::uint64_t)2 * ctx->n_primes * n;
            const auto id = seal_parms_id(ctx, ctx->n_primes);
            put(s, (std::uint64_t)digits);
            for (int j = 0; j < digits; j++)
            {
                put_header(s, 16 + ct_members_bytes(per));
                put_ct_members(s, id, true, 2, n, (std::uint64_t)ctx->n_primes, 1.0, words + (std::size_t)j * per);
            }
        }
        // reads one key_vector; returns its digit count (0 = no key at this index)
        inline int get_kswitch_entry(std::istream &s, const CtxImpl &ctx, std::vector<std::uint64_t> &words)
        {
            const std::uint64_t n = std::uint64_t(1) << ctx->log_n, per = (std::uint64_t)2 * ctx->n_primes * n;
            std::uint64_t digits = get<std::uint64_t>(s);
            if (digits > (std::uint64_t)ctx->n_primes)
                throw std::logic_error("KSwitchKeys data is invalid");
            words.resize((std::size_t)(digits * per));
            for (std::uint64_t j = 0; j < digits; j++)
            {
                (void)get_header(s);
                CtMembers m = get_ct_members(s, ctx);
                if (m.size != 2 || m.limbs != (std::uint64_t)ctx->n_primes || !m.ntt)
                    throw std::logic_error("KSwitchKeys data is invalid");
                std::memcpy(words.data() + (std::size_t)(j * per), m.data.data(), (std::size_t)per * 8);
            }
            return (int)digits;
        }
        inline std::vector<std::uint64_t> download_full_key(const CtxImpl &ctx, bk_kskey_t key, int &digits)
        {
            int limbs = 0;
            check(bk_kskey_info(key, &digits, &limbs, nullptr));
            if (limbs != ctx->n_primes - 1 || digits != ctx->n_primes - 1)
                throw std::logic_error("only full-size keys in SEAL's layout can be saved (level-pruned and hybrid-mode keys "
                                       "have no SEAL representation)");
            std::vector<std::uint64_t> words((std::size_t)digits * 2 * ctx->n_primes << ctx->log_n);
            check(bk_kskey_download(key, words.data()));
            return words;
        }
    } // namespace detail

    inline std::streamoff RelinKeys::save(std::ostream &stream, compr_mode_type m) const
    {
        detail::need_none(m);
        if (!k_ || !ctx_)
            throw std::logic_error("relinearization keys are empty");
        int digits = 0;
        std::vector<std::uint64_t> words = detail::download_full_key(ctx_, k_->h, digits);
        const std::uint64_t total = 16 + 32 + 8 + detail::kswitch_entry_bytes(ctx_, digits);
        detail::put_header(stream, total);
        detail::put(stream, detail::seal_parms_id(ctx_, ctx_->n_primes));
        detail::put(stream, (std::uint64_t)1); // keys_dim1: RelinKeys of size-3 ciphertexts hold one key_vector
        detail::put_kswitch_entry(stream, ctx_, words.data(), digits);
        return (std::streamoff)total;
    }
    inline std::streamoff RelinKeys::load(const SEALContext &context, std::istream &stream)
    {
        const CtxImpl &ctx = context.impl();
        const std::uint64_t total = detail::get_header(stream);
        auto id = detail::get<std::array<std::uint64_t, 4>>(stream);
        if (id != detail::seal_parms_id(ctx, ctx->n_primes))
            throw std::logic_error("RelinKeys data is invalid");
        std::uint64_t dim1 = detail::get<std::uint64_t>(stream);
        if (dim1 != 1)
            throw std::logic_error("RelinKeys for ciphertexts larger than size 3 are not supported");
        std::vector<std::uint64_t> words;
        int digits = detail::get_kswitch_entry(stream, ctx, words);
        if (digits < 1)
            throw std::logic_error("RelinKeys data is invalid");
        auto k = std::make_shared<Holder>();
        detail::check(bk_kskey_upload(ctx->h, words.data(), digits, 0, &k->h));
        k_ = k;
        ctx_ = ctx;
        return (std::streamoff)total;
    }

    inline std::streamoff GaloisKeys::save(std::ostream &stream, compr_mode_type m) const
    {
        detail::need_none(m);
        if (!st_)
            throw std::logic_error("Galois keys are empty");
        State &s = *st_;
        const CtxImpl &ctx = s.ctx;
        // galoiskeys.h: one slot per odd element below 2N (index (elt - 1) / 2), empty where no key was created
        const std::uint64_t slots = std::uint64_t(1) << ctx->log_n;
        for (std::uint32_t elt : s.declared)
            ensure(elt, ctx->n_primes - 1); // a saved key must cover every level (needs the secret key, SEAL layout)
        std::uint64_t total = 16 + 32 + 8 + (slots - s.declared.size()) * 8;
        total += s.declared.size() * detail::kswitch_entry_bytes(ctx, ctx->n_primes - 1);
        detail::put_header(stream, total);
        detail::put(stream, detail::seal_parms_id(ctx, ctx->n_primes));
        detail::put(stream, slots);
        for (std::uint64_t index = 0; index < slots; index++)
        {
            std::uint32_t elt = (std::uint32_t)(2 * index + 1);
            if (!s.declared.count(elt))
            {
                detail::put(stream, (std::uint64_t)0);
                continue;
            }
            bk_kskey_t key = nullptr;
            detail::check(bk_gkeys_get(s.h, elt, &key));
            int digits = 0;
            std::vector<std::uint64_t> words = detail::download_full_key(ctx, key, digits);
            detail::put_kswitch_entry(stream, ctx, words.data(), digits);
        }
        return (std::streamoff)total;
    }
    // Loaded keys are complete (SEAL's layout, every level): the object needs no secret key afterwards.
    inline std::streamoff GaloisKeys::load(const SEALContext &context, std::istream &stream)
    {
        const CtxImpl &ctx = context.impl();
        const std::uint64_t total = detail::get_header(stream);
        auto id = detail::get<std::array<std::uint64_t, 4>>(stream);
        if (id != detail::seal_parms_id(ctx, ctx->n_primes))
            throw std::logic_error("GaloisKeys data is invalid");
        std::uint64_t dim1 = detail::get<std::uint64_t>(stream);
        if (dim1 > (std::uint64_t(1) << ctx->log_n))
            throw std::logic_error("GaloisKeys data is invalid");
        auto st = std::make_shared<State>();
        st->ctx = ctx;
        st->sealed = true;
        detail::check(bk_gkeys_create(ctx->h, &st->h));
        std::vector<std::uint64_t> words;
        for (std::uint64_t index = 0; index < dim1; index++)
        {
            int digits = detail::get_kswitch_entry(stream, ctx, words);
            if (!digits)
                continue;
            std::uint32_t elt = (std::uint32_t)(2 * index + 1);
            bk_kskey_t key = nullptr;
            detail::check(bk_kskey_upload(ctx->h, words.data(), digits, 0, &key));
            std::uint64_t nb = 0;
            detail::check(bk_kskey_info(key, nullptr, nullptr, &nb));
            detail::check(bk_gkeys_set(st->h, elt, key));
            st->declared.insert(elt);
            st->resident[elt] = ctx->n_primes - 1;
            st->bytes += nb;
        }
        detail::check(bk_sync(ctx->h));
        st_ = st;
        return (std::streamoff)total;
    }
} // namespace seal
