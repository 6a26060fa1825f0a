// keygen.cu - KeyGenerator / Encryptor / Decryptor on the device.
//
// Replaces, in the reference's modified SEAL 3.6.6:
//   keygenerator.cpp:50-97    secret key (fork: Hamming-weight secret :64-76)
//   keygenerator.cpp:99-131   public key = encrypt_zero_symmetric at the key level
//   keygenerator.cpp:133-162  relin keys (new key = s^2), :164-233 Galois keys (new key = s∘galois)
//   keygenerator.cpp:384-417  generate_one_kswitch_key: per digit j a fresh symmetric encryption
//                             of zero at the key level with (P mod q_j) * new_key added to limb j
//   util/rlwe.cpp:21-70       ternary / sparse-ternary samplers, :96-135 centred binomial noise
//   util/rlwe.cpp:221-292     encrypt_zero_asymmetric, :294-409 encrypt_zero_symmetric
//   encryptor.cpp:88-239      public-key encryption = encrypt zero one level up, divide by the last prime
//   decryptor.cpp:150-183     ckks_decrypt = dot product of the ciphertext with powers of s
//
// The arithmetic (NTT, products, rounding division) is identical to the reference's.  The
// RANDOMNESS is not: the reference draws from a Blake2xb stream on the CPU, this engine draws
// from ChaCha20 in counter mode on the GPU, keyed per context with 256 bits from the operating
// system (rng.cuh; 275 GiB of key material cannot be sampled on the host in reasonable time),
// so freshly generated keys / ciphertexts have the same distribution but different bits.  Bit-exact parity tests therefore upload the
// reference's keys and ciphertexts (bk_kskey_upload / bk_ct_upload) instead.
#include "engine.h"
#include "rng.cuh"
#include <algorithm>
#include <cstring>
#include <random>

namespace bk
{
    void ntt_fwd_small(Context &c, cudaStream_t s, const int *small, u64 *out, int polys, int limbs, JobMap map);
}
using namespace bk;

// ---- ChaCha20 block -> 128 bits (rng.cuh) --------------------------------------------------------------------
__device__ __forceinline__ uint4 rand128(const RngKey &key, unsigned c0, unsigned c1, unsigned stream, unsigned attempt)
{
    uint32_t w[4];
    chacha20_block<4>(key, c0, c1, stream, attempt, w);
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// uniform residue in [0,q) by rejection (sample_poly_uniform, util/rlwe.cpp:137-178)
__device__ __forceinline__ u64 uniform_mod(const RngKey &key, unsigned stream, u64 index, const PrimeDev &pd)
{
    const u64 max_multiple = 0xFFFFFFFFFFFFFFFFull - barrett64(0xFFFFFFFFFFFFFFFFull, pd) - 1;
    for (unsigned attempt = 0;; attempt++)
    {
        uint4 r = rand128(key, (unsigned)index, (unsigned)(index >> 32), stream, attempt);
        u64 v = ((u64)r.x << 32) | r.y;
        if (v < max_multiple)
            return barrett64(v, pd);
        v = ((u64)r.z << 32) | r.w;
        if (v < max_multiple)
            return barrett64(v, pd);
    }
}

// ---- the uniform halves `a` of level keys: PUBLIC randomness -----------------------------------------------------
// `a` is published with the key (SEAL ships it as a seed, keygenerator.cpp:384-417 / rlwe.cpp:294-409), so it comes
// from its own 256-bit key `akey` - never from the key that samples the errors - through ChaCha8, eight 64-bit words
// per block: word i of a digit's [ne][N] index space is word i % 8 of block i / 8 (nonce = (astream, 0)).  A word at or
// above the largest multiple of q (probability < 2^-13) is replaced by rejection sampling on blocks keyed by the word
// index (nonce = (astream | 2^31, attempt)).  Storing (akey, astream) instead of `a` is the seed-compressed form of a
// key: k_expand_public_a regenerates exactly the words keygen used.
__device__ __forceinline__ u64 public_uniform_retry(const RngKey &akey, unsigned astream, u64 index, u64 max_multiple, const PrimeDev &pd)
{
    for (unsigned attempt = 1;; attempt++)
    {
        uint32_t w[4];
        chacha_block<4, 4>(akey, (unsigned)index, (unsigned)(index >> 32), astream | 0x80000000u, attempt, w);
        u64 v = ((u64)w[0] << 32) | w[1];
        if (v < max_multiple)
            return barrett64(v, pd);
        v = ((u64)w[2] << 32) | w[3];
        if (v < max_multiple)
            return barrett64(v, pd);
    }
}
__device__ __forceinline__ void public_uniform8(const RngKey &akey, unsigned astream, u64 block, const PrimeDev &pd, u64 *out)
{
    const u64 max_multiple = 0xFFFFFFFFFFFFFFFFull - barrett64(0xFFFFFFFFFFFFFFFFull, pd) - 1;
    uint32_t w[16];
    chacha_block<16, 4>(akey, (unsigned)block, (unsigned)(block >> 32), astream, 0u, w);
#pragma unroll
    for (int k = 0; k < 8; k++)
    {
        const u64 v = ((u64)w[2 * k] << 32) | w[2 * k + 1];
        out[k] = v < max_multiple ? barrett64(v, pd) : public_uniform_retry(akey, astream, block * 8 + k, max_multiple, pd);
    }
}

// out[d][j][N] for digits d < dnum and limbs limb0 + j, j < nlimbs, of a key over ne limbs; one thread per block of 8 words
__global__ void __launch_bounds__(256) k_expand_public_a(u64 *__restrict__ out, JobMap map, const PrimeDev *primes, int log_n,
                                                         int ne, int limb0, int nlimbs, int dnum, RngKey akey, unsigned astream0)
{
    pdl_prologue();
    const size_t n = size_t(1) << log_n;
    const size_t per_limb = n / 8, total = (size_t)dnum * nlimbs * per_limb;
    for (size_t b = blockIdx.x * (size_t)blockDim.x + threadIdx.x; b < total; b += (size_t)gridDim.x * blockDim.x)
    {
        const int d = (int)(b / ((size_t)nlimbs * per_limb));
        const size_t r = b % ((size_t)nlimbs * per_limb);
        const int j = (int)(r / per_limb);
        const size_t within = r % per_limb;
        const int limb = limb0 + j;
        const PrimeDev pd = primes[map.prime(limb)];
        u64 v[8];
        public_uniform8(akey, astream0 + (unsigned)d, (size_t)limb * per_limb + within, pd, v);
        ulonglong2 *o = reinterpret_cast<ulonglong2 *>(out + ((size_t)d * nlimbs + j) * n + within * 8);
#pragma unroll
        for (int k = 0; k < 4; k++)
            o[k] = make_ulonglong2(v[2 * k], v[2 * k + 1]);
    }
}

// small signed polynomials: mode 0 = uniform ternary (sample_poly_ternary), 1 = centred binomial
// with sigma 3.2 (sample_poly_cbd: 21 bits minus 21 bits)
__global__ void __launch_bounds__(256) k_sample_small(int *__restrict__ out, size_t count, RngKey key, unsigned stream,
                                                      int mode)
{
    pdl_prologue();
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= count)
        return;
    if (mode == 0)
    {
        for (unsigned attempt = 0;; attempt++)
        {
            uint4 r = rand128(key, (unsigned)i, (unsigned)(i >> 32), stream, attempt);
            // uniform in {0,1,2} by rejection on 32-bit words
            unsigned w[4] = { r.x, r.y, r.z, r.w };
            bool done = false;
            for (int k = 0; k < 4 && !done; k++)
                if (w[k] < 0xFFFFFFFFu - (0xFFFFFFFFu % 3u) - 0u)
                {
                    out[i] = (int)(w[k] % 3u) - 1;
                    done = true;
                }
            if (done)
                return;
        }
    }
    else
    {
        uint4 r = rand128(key, (unsigned)i, (unsigned)(i >> 32), stream, 0u);
        out[i] = __popc(r.x & 0x1FFFFFu) - __popc(r.y & 0x1FFFFFu);
    }
}

// forward column-pass loader for small signed polynomials: job = p * limbs + l
struct LdSmall
{
    const int *src; // [polys][N]
    JobMap map;
    size_t n;
    int limbs;
    __device__ __forceinline__ bool skip(int) const { return false; }
    __device__ __forceinline__ int prime(int job) const { return map.prime(job); }
    __device__ __forceinline__ u64 load(int job, int idx, const PrimeDev &pd) const
    {
        int v = src[(size_t)(job / limbs) * n + idx];
        return v < 0 ? pd.q - (u64)(-v) : (u64)v;
    }
};

namespace bk
{
    // out[p][l][.] = NTT_{prime(l)}(small[p])
    void ntt_fwd_small(Context &c, cudaStream_t s, const int *small, u64 *out, int polys, int limbs, JobMap map)
    {
        const int jobs = polys * limbs;
        Scratch tmp(s, (size_t)jobs * c.n);
        LdSmall ld{ small, map, c.n, limbs };
        dim3 grid(16, jobs);
        switch (c.log_n)
        {
        case 12: launch_pdl(k_fwd_cols<4, LdSmall>, grid, 16, 0, s, ld, tmp.p, c.tables); break;
        case 13: launch_pdl(k_fwd_cols<5, LdSmall>, grid, 32, 0, s, ld, tmp.p, c.tables); break;
        case 14: launch_pdl(k_fwd_cols<6, LdSmall>, grid, 64, 0, s, ld, tmp.p, c.tables); break;
        case 15: launch_pdl(k_fwd_cols<7, LdSmall>, grid, 128, 0, s, ld, tmp.p, c.tables); break;
        default: launch_pdl(k_fwd_cols<8, LdSmall>, grid, 256, 0, s, ld, tmp.p, c.tables); break;
        }
        c.count();
        StPlain st{ out, map, c.n };
        dim3 grid2((unsigned)(c.n >> 12), jobs);
        launch_pdl(k_fwd_blocks<StPlain>, grid2, 256, 256 * 128, s, tmp.p, st, c.tables);
        c.count();
    }
}

// symmetric encryption of zero, NTT form, over `limbs` limbs whose prime is map.prime(l):
//   c1 = a (uniform, sampled directly in NTT form), c0 = -(a*s + e)      (rlwe.cpp:340-371)
// optional: c0[limb == add_limb] += factor * newkey      (keygenerator.cpp:406-416)
// TRANSPOSED: write in the key's transposed-block layout.
template <bool TRANSPOSED, bool PUBLIC_A = false>
__global__ void __launch_bounds__(256) k_sym_zero(u64 *__restrict__ c0, u64 *__restrict__ c1,
                                                  const u64 *__restrict__ sk /*[n_primes][N]*/,
                                                  const u64 *__restrict__ e_ntt /*[limbs][N]*/,
                                                  const u64 *__restrict__ newkey /*[n_primes][N] or null*/,
                                                  int add_limb, u64 factor, JobMap map, const PrimeDev *primes,
                                                  int log_n, int limbs, RngKey rkey, unsigned stream,
                                                  const u64 *__restrict__ factors = nullptr /*[limbs], 0 = none*/)
{
    pdl_prologue();
    const size_t n = size_t(1) << log_n;
    const size_t total = (size_t)limbs * n;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
    {
        int l = (int)(i >> log_n);
        size_t idx = i & (n - 1);
        int pi = map.prime(l);
        const PrimeDev pd = primes[pi];
        u64 a;
        if (PUBLIC_A)
        { // (akey, astream) ride in (rkey, stream): see hybrid_key
            u64 blockw[8];
            public_uniform8(rkey, stream, i >> 3, pd, blockw);
            a = blockw[i & 7];
        }
        else
            a = uniform_mod(rkey, stream, i, pd);
        u64 s = sk[(size_t)pi * n + idx];
        u64 v = addmod(mulmod(a, s, pd), e_ntt[i], pd.q);
        v = v ? pd.q - v : 0ull;
        if (newkey && factors)
        {
            u64 f = factors[l];
            if (f)
                v = addmod(v, mulmod(f, newkey[(size_t)pi * n + idx], pd), pd.q);
        }
        else if (newkey && l == add_limb)
            v = addmod(v, mulmod(factor, newkey[(size_t)pi * n + idx], pd), pd.q);
        size_t o = i;
        if (TRANSPOSED)
        {
            size_t blk = idx >> 8;
            int e = (int)(idx & 255);
            int t = e >> 4, k = e & 15;
            o = ((size_t)l << log_n) + (blk << 8) + (size_t)(k * 16 + t);
        }
        c0[o] = v;
        if (c1)
            c1[o] = a;
    }
}

// c_j = pk_j * u + e_j over limbs 0..limbs-1 (rlwe.cpp:253-291); pk is [2][n_primes][N]
__global__ void __launch_bounds__(256) k_asym_zero(u64 *__restrict__ out /*[2][limbs][N]*/,
                                                   const u64 *__restrict__ pk, const u64 *__restrict__ u_ntt,
                                                   const u64 *__restrict__ e_ntt /*[2][limbs][N]*/,
                                                   const PrimeDev *primes, int log_n, int limbs, int n_primes)
{
    pdl_prologue();
    const size_t n = size_t(1) << log_n;
    const size_t per = (size_t)limbs * n;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < 2 * per; i += (size_t)gridDim.x * blockDim.x)
    {
        int p = (int)(i / per);
        size_t r = i % per;
        int l = (int)(r >> log_n);
        size_t idx = r & (n - 1);
        const PrimeDev pd = primes[l];
        u64 k = pk[((size_t)p * n_primes + l) * n + idx];
        out[i] = addmod(mulmod(k, u_ntt[r], pd), e_ntt[i], pd.q);
    }
}

// dot_product_ct_sk_array (decryptor.cpp:185-260): m = c0 + c1 s + c2 s^2 ...
__global__ void __launch_bounds__(256) k_decrypt(const u64 *__restrict__ ct, const u64 *__restrict__ sk,
                                                 u64 *__restrict__ out, const PrimeDev *primes, int log_n, int limbs,
                                                 int size)
{
    pdl_prologue();
    const size_t n = size_t(1) << log_n;
    const size_t per = (size_t)limbs * n;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < per; i += (size_t)gridDim.x * blockDim.x)
    {
        int l = (int)(i >> log_n);
        const PrimeDev pd = primes[l];
        u64 s = sk[i];
        u64 acc = ct[(size_t)(size - 1) * per + i];
        for (int p = size - 2; p >= 0; p--)
            acc = addmod(mulmod(acc, s, pd), ct[(size_t)p * per + i], pd.q);
        out[i] = acc;
    }
}

__global__ void __launch_bounds__(256) k_permute_limbs(const u64 *__restrict__ src, u64 *__restrict__ dst,
                                                       const uint32_t *__restrict__ perm, int log_n, int jobs)
{
    pdl_prologue();
    const size_t n = size_t(1) << log_n;
    const size_t total = (size_t)jobs * n;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
    {
        size_t j = i >> log_n, k = i & (n - 1);
        dst[i] = src[(j << log_n) + perm[k]];
    }
}

namespace bk
{
    static std::atomic<unsigned> g_stream_counter{ 1 };

    // one kswitch key for `newkey` ([n_primes][N] NTT form), pruned to max_limbs
    static bk_kskey_t make_kskey(Context &c, bk_sk_t sk, const u64 *newkey, u64 seed, int max_limbs, int kind, uint32_t elt)
    {
        const int top = c.top_limbs();
        const int sp = c.n_primes - 1;
        int kl = (max_limbs > 0 && max_limbs < top) ? max_limbs : top;
        cudaStream_t s = c.stream();
        const size_t n = c.n;
        auto key = new bk_kskey_s();
        key->ctx = static_cast<bk_context_t>(&c);
        key->sk = sk;
        key->kind = kind;
        key->elt = elt;
        key->seed = seed;
        const RngKey rk = derive_call_key(c.rng_master, seed);
        if (c.hybrid)
        {
            // a recipe only: the level-specific keys are generated by hybrid_key() at the levels that use them
            key->recipe = true;
            key->digits = top;
            key->klimbs = top;
            return key;
        }
        key->digits = kl;
        key->klimbs = kl;
        key->words = (size_t)kl * 2 * (kl + 1) * n;
        BK_CUDA(cudaMalloc((void **)&key->d, key->words * sizeof(u64)));
        JobMap map = limb_map(kl + 1);
        map.special_pos = kl;
        map.special_prime = sp;
        Scratch small(s, (n + 1) / 2);
        Scratch e_ntt(s, (size_t)(kl + 1) * n);
        for (int j = 0; j < kl; j++)
        {
            unsigned st_e = g_stream_counter.fetch_add(2);
            launch_pdl(k_sample_small, (unsigned)((n + 255) / 256), 256, 0, s, (int *)small.p, n, rk, st_e, 1);
            c.count();
            ntt_fwd_small(c, s, (const int *)small.p, e_ntt.p, 1, kl + 1, map);
            u64 factor = c.primes[sp] % c.primes[j];
            u64 *c0 = key->d + ((size_t)j * 2) * (kl + 1) * n;
            u64 *c1 = c0 + (size_t)(kl + 1) * n;
            launch_pdl(k_sym_zero<false>, c.ew_grid((size_t)(kl + 1) * n), 256, 0, s, c0, c1, sk->d, e_ntt.p, newkey, j, factor,
                                                                            map, c.d_primes, c.log_n, kl + 1, rk,
                                                                            st_e + 1, (const u64 *)nullptr);
            c.count();
        }
        BK_CUDA(cudaStreamSynchronize(s));
        return key;
    }

    // what the key switches from, [n_primes][N] NTT form: s^2 (keygenerator.cpp:133-162) or sigma_elt(s) (:199-227)
    static void compute_newkey(Context &c, cudaStream_t s, bk_sk_t sk, int kind, uint32_t elt, u64 *out)
    {
        const size_t words = (size_t)c.n_primes * c.n;
        if (kind == 1)
        {
            BK_CUDA(cudaMemcpyAsync(out, sk->d, words * sizeof(u64), cudaMemcpyDeviceToDevice, s));
            launch_pdl(k_ew<EW_MUL>, c.ew_grid(words / 2), 256, 0, s, out, sk->d, c.d_primes, c.log_n, c.n_primes, 1, 1);
        }
        else
        {
            const uint32_t *perm = c.galois_table(elt);
            launch_pdl(k_permute_limbs, c.ew_grid(words), 256, 0, s, sk->d, out, perm, c.log_n, c.n_primes);
        }
        c.count();
    }

    // The level-l key of a recipe: dnum digits, digit d = (b_d, a_d) over the l + alpha extended moduli with
    // b_d = -(a_d s + e_d) + [P_S mod q_i] s' on the limbs of digit d (HybridPlan::d_keyfactor), where s' is the key
    // being switched from and P_S the product of the alpha special moduli of this level.
    bk_hybkey_s *hybrid_key(Context &c, bk_kskey_s *key, int l)
    {
        std::lock_guard<std::mutex> guard(key->hmu);
        auto it = key->hyb.find(l);
        if (it != key->hyb.end())
            return it->second;
        if (!key->sk || !key->kind)
            throw std::invalid_argument(key->kind ? "key switching key not present for this level (the keys were generated from a "
                                                    "plan and the secret key has been detached)"
                                                  : "kswitch_keys is not valid for encryption parameters (no recipe for a level key)");
        const HybridPlan &P = hybrid_plan(c, l);
        cudaStream_t s = c.stream();
        const size_t n = c.n;
        const int ne = P.ne;
        auto hk = std::make_unique<bk_hybkey_s>();
        hk->l = l;
        hk->alpha = P.alpha;
        hk->dsize = P.dsize;
        hk->dnum = P.dnum;
        // The uniform halves come from a PUBLIC 256-bit key of their own (never the key the errors are sampled with):
        // with key compression on, only (akey, astream0) is kept and the halves are regenerated when the key is used
        // (hyb_extend_and_mac, engine.cu) - SEAL's seeded keys (keygenerator.cpp:384-417), resident in half the bytes.
        hk->compressed = c.compress_keys;
        hk->akey = derive_call_key(c.rng_master, key->seed + 0xA5A5A5A5DEADBEEFull + 0x9E3779B97F4A7C15ull * (u64)(l + 1));
        hk->astream0 = g_stream_counter.fetch_add((unsigned)P.dnum);
        const int polys = hk->compressed ? 1 : 2;
        hk->words = (size_t)P.dnum * polys * ne * n;
        BK_CUDA(cudaMalloc((void **)&hk->d, hk->words * sizeof(u64)));
        JobMap map = limb_map(ne);
        map.special_pos = ne - 1;
        map.special_prime = c.n_primes - 1;
        Scratch newkey(s, (size_t)c.n_primes * n);
        compute_newkey(c, s, key->sk, key->kind, key->elt, newkey.p);
        Scratch small(s, (n + 1) / 2);
        Scratch e_ntt(s, (size_t)ne * n);
        const RngKey rk = derive_call_key(c.rng_master, key->seed + 0x9E3779B97F4A7C15ull * (u64)(l + 1));
        for (int d = 0; d < P.dnum; d++)
        {
            unsigned st_e = g_stream_counter.fetch_add(2);
            launch_pdl(k_sample_small, (unsigned)((n + 255) / 256), 256, 0, s, (int *)small.p, n, rk, st_e, 1);
            c.count();
            ntt_fwd_small(c, s, (const int *)small.p, e_ntt.p, 1, ne, map);
            u64 *c0 = hk->d + ((size_t)d * polys) * ne * n;
            u64 *c1 = hk->compressed ? nullptr : c0 + (size_t)ne * n;
            launch_pdl(k_sym_zero<false, true>, c.ew_grid((size_t)ne * n), 256, 0, s, c0, c1, key->sk->d, e_ntt.p, newkey.p, -1, 0, map,
                       c.d_primes, c.log_n, ne, hk->akey, hk->astream0 + (unsigned)d, P.d_keyfactor + (size_t)d * ne);
            c.count();
        }
        BK_CUDA(cudaStreamSynchronize(s)); // complete before other host threads (streams) can pick it up
        c.hybrid_key_bytes.fetch_add(hk->words * sizeof(u64));
        c.hybrid_keys.fetch_add(1);
        bk_hybkey_s *raw = hk.release();
        key->hyb[l] = raw;
        return raw;
    }
}

namespace bk
{
    // the uniform halves of a seed-compressed level key for limbs [limb0, limb0 + nlimbs): out[dnum][nlimbs][N]
    void expand_public_halves(Context &c, cudaStream_t s, const bk_hybkey_s *hk, int limb0, int nlimbs, u64 *out)
    {
        const int ne = hk->l + hk->alpha;
        JobMap map = limb_map(ne);
        map.special_pos = ne - 1;
        map.special_prime = c.n_primes - 1;
        ProfScope ps(c, s, TAG_OTHER, hk->dnum * nlimbs);
        launch_pdl(k_expand_public_a, c.ew_grid((size_t)hk->dnum * nlimbs * c.n / 8), 256, 0, s, out, map, c.d_primes, c.log_n, ne, limb0,
                   nlimbs, hk->dnum, hk->akey, hk->astream0);
        c.count();
    }
}

extern "C"
{
    bk_status bk_sk_generate(bk_context_t ctx, int hamming_weight, uint64_t seed, bk_sk_t *out)
    {
        BK_TRY
        Context &c = *ctx;
        const size_t n = c.n;
        if (hamming_weight < 0 || (size_t)hamming_weight > n)
            throw std::invalid_argument("hamming_weight is invalid");
        std::vector<int> small(n, 0);
        // host-side words of ChaCha20(call key, counter = block number, nonce = "sk")
        struct HostStream
        {
            RngKey key;
            uint32_t buf[16];
            uint64_t block = 0;
            int pos = 16;
            uint32_t next()
            {
                if (pos == 16)
                {
                    chacha20_block<16>(key, (uint32_t)block, (uint32_t)(block >> 32), 0x6b73u, 0u, buf);
                    block++;
                    pos = 0;
                }
                return buf[pos++];
            }
            // uniform in [0, bound) by rejection
            uint32_t below(uint32_t bound)
            {
                const uint32_t limit = 0xFFFFFFFFu - (0xFFFFFFFFu % bound + 1u) % bound;
                for (;;)
                {
                    uint32_t v = next();
                    if (v <= limit)
                        return v % bound;
                }
            }
        } rng{ derive_call_key(c.rng_master, seed) };
        if (hamming_weight == 0)
        {
            for (auto &v : small)
                v = (int)rng.below(3) - 1;
        }
        else
        {
            // sample_poly_sparse_ternary (rlwe.cpp:40-70): h distinct positions, each +-1
            int w = 0;
            while (w < hamming_weight)
            {
                size_t i = rng.below((uint32_t)n);
                if (small[i])
                    continue;
                small[i] = (rng.next() & 1u) ? 1 : -1;
                w++;
            }
        }
        cudaStream_t s = c.stream();
        Scratch d_small(s, (n + 1) / 2);
        BK_CUDA(cudaMemcpyAsync(d_small.p, small.data(), n * sizeof(int), cudaMemcpyHostToDevice, s));
        auto sk = new bk_sk_s();
        sk->ctx = ctx;
        BK_CUDA(cudaMalloc((void **)&sk->d, (size_t)c.n_primes * n * sizeof(u64)));
        ntt_fwd_small(c, s, (const int *)d_small.p, sk->d, 1, c.n_primes, limb_map(c.n_primes));
        BK_CUDA(cudaStreamSynchronize(s));
        *out = sk;
        BK_END
    }
    bk_status bk_sk_upload(bk_context_t ctx, const uint64_t *host, bk_sk_t *out)
    {
        BK_TRY
        Context &c = *ctx;
        c.activate();
        auto sk = new bk_sk_s();
        sk->ctx = ctx;
        BK_CUDA(cudaMalloc((void **)&sk->d, (size_t)c.n_primes * c.n * sizeof(u64)));
        BK_CUDA(cudaMemcpy(sk->d, host, (size_t)c.n_primes * c.n * sizeof(u64), cudaMemcpyHostToDevice));
        BK_CUDA(cudaDeviceSynchronize());
        *out = sk;
        BK_END
    }
    bk_status bk_sk_download(bk_sk_t sk, uint64_t *host_out)
    {
        BK_TRY
        Context &c = *sk->ctx;
        BK_CUDA(cudaStreamSynchronize(c.stream()));
        BK_CUDA(cudaMemcpy(host_out, sk->d, (size_t)c.n_primes * c.n * sizeof(u64), cudaMemcpyDeviceToHost));
        BK_END
    }
    bk_status bk_sk_destroy(bk_sk_t sk)
    {
        BK_TRY
        if (sk)
        {
            sk->ctx->activate();
            cudaStreamSynchronize(sk->ctx->stream());
            cudaFree(sk->d);
            delete sk;
        }
        BK_END
    }

    bk_status bk_pk_generate(bk_context_t ctx, bk_sk_t sk, uint64_t seed, bk_ct_t pk_out)
    {
        BK_TRY
        // keygenerator.cpp:99-131: encrypt_zero_symmetric at the key level, NTT form
        Context &c = *ctx;
        const RngKey rk = derive_call_key(c.rng_master, seed);
        if (!sk || sk->ctx != ctx)
            throw std::logic_error("cannot generate public key for unspecified secret key");
        cudaStream_t s = c.stream();
        const size_t n = c.n;
        const int L = c.n_primes;
        ensure_ct(pk_out, 2, L, false);
        Scratch small(s, (n + 1) / 2);
        Scratch e_ntt(s, (size_t)L * n);
        unsigned st = g_stream_counter.fetch_add(2);
        launch_pdl(k_sample_small, (unsigned)((n + 255) / 256), 256, 0, s, (int *)small.p, n, rk, st, 1);
        c.count();
        ntt_fwd_small(c, s, (const int *)small.p, e_ntt.p, 1, L, limb_map(L));
        launch_pdl(k_sym_zero<false>, c.ew_grid((size_t)L * n), 256, 0, s, pk_out->d, pk_out->d + (size_t)L * n, sk->d,
                                                                   e_ntt.p, (const u64 *)nullptr, -1, 0, limb_map(L), c.d_primes,
                                                                   c.log_n, L, rk, st + 1, (const u64 *)nullptr);
        c.count();
        pk_out->scale = 1.0;
        pk_out->ntt = true;
        BK_END
    }

    bk_status bk_relin_key_generate(bk_context_t ctx, bk_sk_t sk, uint64_t seed, int max_limbs, bk_kskey_t *out)
    {
        BK_TRY
        // keygenerator.cpp:133-162: new key = s^2 (compute_secret_key_array)
        Context &c = *ctx;
        if (!sk || sk->ctx != ctx)
            throw std::logic_error("cannot generate relinearization keys for unspecified secret key");
        cudaStream_t s = c.stream();
        size_t words = (size_t)c.n_primes * c.n;
        Scratch sq(s, words);
        BK_CUDA(cudaMemcpyAsync(sq.p, sk->d, words * sizeof(u64), cudaMemcpyDeviceToDevice, s));
        launch_pdl(k_ew<EW_MUL>, c.ew_grid(words / 2), 256, 0, s, sq.p, sk->d, c.d_primes, c.log_n, c.n_primes, 1, 1);
        c.count();
        *out = make_kskey(c, sk, sq.p, seed, max_limbs, 1, 0);
        BK_END
    }

    bk_status bk_galois_key_generate(bk_context_t ctx, bk_sk_t sk, uint32_t galois_elt, uint64_t seed, int max_limbs,
                                     bk_kskey_t *out)
    {
        BK_TRY
        // keygenerator.cpp:199-227: new key = apply_galois_ntt(s, elt)
        Context &c = *ctx;
        if (!sk || sk->ctx != ctx)
            throw std::logic_error("cannot generate Galois keys for unspecified secret key");
        if (!(galois_elt & 1) || galois_elt >= 2 * c.n)
            throw std::invalid_argument("Galois element is not valid");
        cudaStream_t s = c.stream();
        size_t words = (size_t)c.n_primes * c.n;
        Scratch rot(s, words);
        const uint32_t *perm = c.galois_table(galois_elt);
        launch_pdl(k_permute_limbs, c.ew_grid(words), 256, 0, s, sk->d, rot.p, perm, c.log_n, c.n_primes);
        c.count();
        *out = make_kskey(c, sk, rot.p, seed, max_limbs, 2, galois_elt);
        BK_END
    }

    bk_status bk_encrypt(bk_context_t ctx, bk_ct_t pk, bk_pt_t pt, uint64_t seed, bk_ct_t out)
    {
        BK_TRY
        // Encryptor::encrypt_internal (encryptor.cpp:165-239) -> encrypt_zero_internal (:88-163):
        // encrypt zero with l+1 limbs (the next prime up, or the special prime at the top level),
        // divide by that prime with rounding, add the plaintext to c0.
        Context &c = *ctx;
        const RngKey rk = derive_call_key(c.rng_master, seed);
        if (!pk || pk->ctx != ctx || pk->size != 2 || pk->limbs != c.n_primes)
            throw std::logic_error("public key is not set");
        if (!pt || pt->ctx != ctx || !pt->d)
            throw std::invalid_argument("plain is not valid for encryption parameters");
        cudaStream_t s = c.stream();
        const size_t n = c.n;
        const int l = pt->limbs, l1 = l + 1;
        if (l > c.top_limbs())
            throw std::invalid_argument("plain is not valid for encryption parameters");
        Scratch small(s, (3 * n + 1) / 2);
        int *d_small = (int *)small.p;
        unsigned st = g_stream_counter.fetch_add(2);
        launch_pdl(k_sample_small, (unsigned)((n + 255) / 256), 256, 0, s, d_small, n, rk, st, 0);
        launch_pdl(k_sample_small, (unsigned)((2 * n + 255) / 256), 256, 0, s, d_small + n, 2 * n, rk, st + 1, 1);
        c.count(2);
        Scratch u_ntt(s, (size_t)l1 * n);
        Scratch e_ntt(s, (size_t)2 * l1 * n);
        ntt_fwd_small(c, s, d_small, u_ntt.p, 1, l1, limb_map(l1));
        ntt_fwd_small(c, s, d_small + n, e_ntt.p, 2, l1, limb_map(l1));
        bk_ct_s tmp;
        tmp.ctx = ctx;
        ensure_ct(&tmp, 2, l1, false);
        launch_pdl(k_asym_zero, c.ew_grid((size_t)2 * l1 * n), 256, 0, s, tmp.d, pk->d, u_ntt.p, e_ntt.p, c.d_primes, c.log_n, l1,
                                                                  c.n_primes);
        c.count();
        tmp.ntt = true;
        tmp.scale = 1.0;
        // divide_and_round_q_last_ntt_inplace by prime index l (the special prime when l is the top level:
        // there prime index l == n_primes - 1)
        rescale_core(c, &tmp);
        launch_pdl(k_ew<EW_ADD>, c.ew_grid((size_t)l * n / 2), 256, 0, s, tmp.d, pt->d, c.d_primes, c.log_n, l, 1, 1);
        c.count();
        c.release_words(out->d, out->owner);
        out->d = tmp.d;
        out->owner = tmp.owner ? tmp.owner : s;
        out->cap = tmp.cap;
        out->size = 2;
        out->limbs = l;
        out->scale = pt->scale;
        out->ntt = true;
        BK_END
    }

    bk_status bk_encrypt_symmetric(bk_context_t ctx, bk_sk_t sk, bk_pt_t pt, uint64_t seed, bk_ct_t out)
    {
        BK_TRY
        Context &c = *ctx;
        const RngKey rk = derive_call_key(c.rng_master, seed);
        if (!sk || sk->ctx != ctx)
            throw std::logic_error("secret key is not set");
        if (!pt || pt->ctx != ctx || !pt->d)
            throw std::invalid_argument("plain is not valid for encryption parameters");
        cudaStream_t s = c.stream();
        const size_t n = c.n;
        const int l = pt->limbs;
        ensure_ct(out, 2, l, false);
        Scratch small(s, (n + 1) / 2);
        Scratch e_ntt(s, (size_t)l * n);
        unsigned st = g_stream_counter.fetch_add(2);
        launch_pdl(k_sample_small, (unsigned)((n + 255) / 256), 256, 0, s, (int *)small.p, n, rk, st, 1);
        c.count();
        ntt_fwd_small(c, s, (const int *)small.p, e_ntt.p, 1, l, limb_map(l));
        launch_pdl(k_sym_zero<false>, c.ew_grid((size_t)l * n), 256, 0, s, out->d, out->d + (size_t)l * n, sk->d, e_ntt.p,
                                                                   (const u64 *)nullptr, -1, 0, limb_map(l), c.d_primes, c.log_n, l,
                                                                   rk, st + 1, (const u64 *)nullptr);
        launch_pdl(k_ew<EW_ADD>, c.ew_grid((size_t)l * n / 2), 256, 0, s, out->d, pt->d, c.d_primes, c.log_n, l, 1, 1);
        c.count(2);
        out->scale = pt->scale;
        out->ntt = true;
        BK_END
    }

    bk_status bk_decrypt(bk_context_t ctx, bk_sk_t sk, bk_ct_t ct, bk_pt_t out)
    {
        BK_TRY
        Context &c = *ctx;
        if (!sk || sk->ctx != ctx)
            throw std::invalid_argument("secret key is not valid for encryption parameters");
        if (!ct || ct->ctx != ctx || !ct->d || ct->size < 2)
            throw std::invalid_argument("encrypted is not valid for encryption parameters");
        if (!ct->ntt)
            throw std::invalid_argument("encrypted must be in NTT form");
        const int l = ct->limbs;
        ensure_pt(out, l);
        launch_pdl(k_decrypt, c.ew_grid((size_t)l * c.n), 256, 0, c.stream(), ct->d, sk->d, out->d, c.d_primes, c.log_n, l,
                                                                      ct->size);
        c.count();
        out->scale = ct->scale;
        BK_END
    }
}
