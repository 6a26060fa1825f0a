// engine.cu - context, containers and the Evaluator half of the C ABI (include/b200ckks.h).
//
// Host orchestration of the sm_100a kernels in ntt.cuh / kernels.cuh.  Each entry point
// replaces one member of the reference's seal::Evaluator (evaluator.cpp, cited per function);
// argument checks and their messages follow the reference so the C++ facade can rethrow the
// same exception types.  There is no CPU arithmetic path: without a CUDA device
// bk_context_create fails with BK_NO_DEVICE.
#include "engine.h"
#include <sys/random.h>
#include <execinfo.h>
#include <cstdio>
#include <cmath>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <limits>

// 8 independent chains of 32-bit multiply-adds per thread (bk_measure_imad_peak).  KIND 0: mad.lo.u32 (32-bit result),
// 1: mad.wide.u32 (64-bit result), 2: mad.hi.u32
template <int KIND>
static __global__ void __launch_bounds__(256) k_imad_peak(unsigned long long *sink, int iters, unsigned m)
{
    pdl_prologue();
    unsigned a[8];
    unsigned long long w[8];
#pragma unroll
    for (int j = 0; j < 8; j++)
    {
        a[j] = threadIdx.x + j;
        w[j] = threadIdx.x + j;
    }
#pragma unroll 4
    for (int i = 0; i < iters; i++)
    {
#pragma unroll
        for (int j = 0; j < 8; j++)
        {
            if (KIND == 0)
                asm volatile("mad.lo.u32 %0, %0, %1, %0;" : "+r"(a[j]) : "r"(m));
            else if (KIND == 1)
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[j]) : "r"((unsigned)w[j]), "r"(m)); // operand from the chain: not loop-invariant
            else
                asm volatile("mad.hi.u32 %0, %0, %1, %0;" : "+r"(a[j]) : "r"(m));
        }
    }
    unsigned long long r = 0;
#pragma unroll
    for (int j = 0; j < 8; j++)
        r ^= a[j] ^ w[j];
    if (r == 0x12345678ull && iters < 0)
        *sink = r;
}

namespace bk
{
    static thread_local std::string g_err;

    void set_error(const char *msg)
    {
        g_err = msg;
    }

    bk_status fence(const std::exception_ptr &e)
    {
        try
        {
            std::rethrow_exception(e);
        }
        catch (const NoDevice &x)
        {
            g_err = x.what();
            return BK_NO_DEVICE;
        }
        catch (const CudaError &x)
        {
            g_err = x.what();
            return BK_CUDA_ERROR;
        }
        catch (const std::invalid_argument &x)
        {
            g_err = x.what();
            return BK_INVALID_ARGUMENT;
        }
        catch (const std::out_of_range &x)
        {
            g_err = x.what();
            return BK_OUT_OF_RANGE;
        }
        catch (const std::logic_error &x)
        {
            g_err = x.what();
            return BK_LOGIC_ERROR;
        }
        catch (const std::exception &x)
        {
            g_err = x.what();
            return BK_LOGIC_ERROR;
        }
        catch (...)
        {
            g_err = "unknown error";
            return BK_LOGIC_ERROR;
        }
    }

    static int bit_count_u64(uint64_t v)
    {
        int b = 0;
        while (v)
        {
            b++;
            v >>= 1;
        }
        return b;
    }

    // ---------------------------------------------------------------------------------- Context
    // ---- scratch arenas, one per stream a context created
    static std::mutex g_arena_mu;
    static std::unordered_map<cudaStream_t, ScratchArena *> g_arenas;
    static std::atomic<uint64_t> g_arena_generation{ 1 }; // bumped whenever an arena is retired (stream handles get reused)

    Context::Context(int log_n_, const uint64_t *primes_, int n_primes_, int device_)
        : log_n(log_n_), n(size_t(1) << log_n_), n_primes(n_primes_), device(device_)
    {
        if (log_n < 12 || log_n > 16)
            throw std::invalid_argument("poly_modulus_degree must be 2^12 .. 2^16");
        if (n_primes < 2 || n_primes > 62)
            throw std::invalid_argument("coeff_modulus size is invalid");
        if (const char *e = std::getenv("B200CKKS_HYBRID_KS"))
            hybrid = std::atoi(e) != 0;
        if (const char *e = std::getenv("B200CKKS_COMPRESS_KEYS"))
            compress_keys = std::atoi(e) != 0;
        if (const char *e = std::getenv("B200CKKS_DEBUG_SYNC"))
            debug_sync = std::atoi(e) != 0;
        // random generator master key (rng.cuh): the operating system's entropy unless a reproducible run is asked for
        if (const char *e = std::getenv("B200CKKS_SEED"))
        {
            uint64_t v = std::strtoull(e, nullptr, 0);
            RngKey seed_key = { { (uint32_t)v, (uint32_t)(v >> 32), 0x62323030u, 0x636b6b73u, 0, 0, 0, 0 } };
            rng_master = derive_call_key(seed_key, 0);
        }
        else
        {
            size_t got = 0;
            while (got < sizeof rng_master.k)
            {
                ssize_t r = getrandom((char *)rng_master.k + got, sizeof rng_master.k - got, 0);
                if (r < 0)
                    throw std::runtime_error("getrandom failed: no entropy source for key and encryption randomness");
                got += (size_t)r;
            }
        }
        int dev_count = 0;
        if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count == 0)
            throw NoDevice("no CUDA device: this engine has no CPU path");
        if (device < 0 || device >= dev_count)
            throw std::invalid_argument("device ordinal out of range");
        BK_CUDA(cudaSetDevice(device));
        cudaDeviceProp prop;
        BK_CUDA(cudaGetDeviceProperties(&prop, device));
        sm_count = prop.multiProcessorCount;

        primes.assign(primes_, primes_ + n_primes);
        for (int i = 0; i < n_primes; i++)
        {
            uint64_t q = primes[i];
            // SEAL's own user limit (SEAL_USER_MOD_BIT_COUNT_MAX = 60, util/defines.h:33-40).  It also bounds the unreduced
            // 128-bit sums of k_ks_mac / k_mul_plain_sum: at most 61 digits in [0, 4q) times keys in [0, q) stay below
            // 61 * 4 * 2^120 < 2^128.
            if (q >> 60 || !is_prime_u64(q) || (q - 1) % (2 * n) != 0)
                throw std::invalid_argument("coeff_modulus is not valid (need NTT-friendly primes of at most 60 bits)");
            for (int j = 0; j < i; j++)
                if (primes[j] == q)
                    throw std::invalid_argument("coeff_modulus primes must be distinct");
        }
        // total_coeff_modulus_bit_count per level (context.cpp:455-523): exact product bit length
        total_bits.assign(n_primes + 1, 0);
        {
            std::vector<uint64_t> prod(1, 1); // little-endian multiword product
            for (int l = 1; l <= n_primes; l++)
            {
                uint64_t carry = 0;
                for (auto &w : prod)
                {
                    u128 t = (u128)w * primes[l - 1] + carry;
                    w = (uint64_t)t;
                    carry = (uint64_t)(t >> 64);
                }
                if (carry)
                    prod.push_back(carry);
                total_bits[l] = (int)(prod.size() - 1) * 64 + bit_count_u64(prod.back());
            }
        }

        // memory pool: never hand memory back to the driver between ops
        cudaMemPool_t pool;
        BK_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
        uint64_t thresh = UINT64_MAX;
        BK_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh));
        // $B200CKKS_POOL_NODEPS=1: never let the allocator make one stream wait for another to reuse a freed block.
        // Measured with images in flight on their own streams (tools/in_flight_sweep.py): slower at 4 in flight (0.76 vs
        // 0.86 images/s, fresh blocks cost more than the waits), faster at 6 (0.84 vs 0.65); the default stays.
        if (std::getenv("B200CKKS_POOL_NODEPS"))
        {
            int off = 0;
            BK_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolReuseAllowInternalDependencies, &off));
        }

        // per-prime constants + twiddle tables
        h_primes.resize(n_primes);
        std::vector<ulonglong2> tw((size_t)n_primes * n), itw((size_t)n_primes * n);
        for (int i = 0; i < n_primes; i++)
        {
            uint64_t q = primes[i];
            std::vector<uint64_t> w, iw;
            ntt_tables(log_n, q, w, iw);
            for (size_t k = 0; k < n; k++)
            {
                tw[(size_t)i * n + k] = make_ulonglong2(w[k], shoup(w[k], q));
                itw[(size_t)i * n + k] = make_ulonglong2(iw[k], shoup(iw[k], q));
            }
            PrimeDev &pd = h_primes[i];
            pd.q = q;
            pd.two_q = 2 * q;
            // floor(2^128 / q) as two words (Modulus::const_ratio, modulus.cpp)
            u128 hi = (~(u128)0) / q; // floor((2^128-1)/q) == floor(2^128/q) because q is odd > 1
            pd.r0 = (uint64_t)hi;
            pd.r1 = (uint64_t)(hi >> 64);
            uint64_t ninv = invmod((uint64_t)n % q, q);
            pd.ninv = ninv;
            pd.ninv_s = shoup(ninv, q);
            uint64_t nw = mulmod(ninv, iw[1], q);
            pd.ninvw = nw;
            pd.ninvw_s = shoup(nw, q);
        }
        std::vector<ulonglong2> inv((size_t)n_primes * n_primes);
        for (int last = 0; last < n_primes; last++)
            for (int i = 0; i < n_primes; i++)
            {
                if (i == last)
                {
                    inv[(size_t)last * n_primes + i] = make_ulonglong2(0, 0);
                    continue;
                }
                uint64_t v = invmod(primes[last] % primes[i], primes[i]);
                inv[(size_t)last * n_primes + i] = make_ulonglong2(v, shoup(v, primes[i]));
            }
        BK_CUDA(cudaMalloc((void **)&d_primes, sizeof(PrimeDev) * n_primes));
        BK_CUDA(cudaMalloc((void **)&d_tw, sizeof(ulonglong2) * tw.size()));
        BK_CUDA(cudaMalloc((void **)&d_itw, sizeof(ulonglong2) * itw.size()));
        BK_CUDA(cudaMalloc((void **)&d_inv, sizeof(ulonglong2) * inv.size()));
        BK_CUDA(cudaMemcpy(d_primes, h_primes.data(), sizeof(PrimeDev) * n_primes, cudaMemcpyHostToDevice));
        BK_CUDA(cudaMemcpy(d_tw, tw.data(), sizeof(ulonglong2) * tw.size(), cudaMemcpyHostToDevice));
        BK_CUDA(cudaMemcpy(d_itw, itw.data(), sizeof(ulonglong2) * itw.size(), cudaMemcpyHostToDevice));
        BK_CUDA(cudaMemcpy(d_inv, inv.data(), sizeof(ulonglong2) * inv.size(), cudaMemcpyHostToDevice));
        // a pageable host-to-device cudaMemcpy may return before the DMA has landed, and the engine's streams are
        // non-blocking (no implicit ordering against the legacy stream): drain before any kernel can read the tables
        BK_CUDA(cudaDeviceSynchronize());
        tables.tw = d_tw;
        tables.itw = d_itw;
        tables.primes = d_primes;
        tables.log_n = log_n;
        {
            size_t ok = 0;
            for (uint64_t q : primes)
                ok += !(q >> 57) && (q >> 32); // the unreduced forward butterflies need 66 q < 2^64 (and barrett64_r32)
            tables.wide = ok == primes.size() ? 1 : ok ? 2 : 0;
        }
        if (std::getenv("B200CKKS_CLASSIC_NTT"))
            tables.wide = 0;
    }

    bool pdl_enabled()
    {
        static const bool on = std::getenv("B200CKKS_NO_PDL") == nullptr;
        return on;
    }

    Context::~Context()
    {
        cudaSetDevice(device);
        cudaDeviceSynchronize();
        destroy_encoder(*this);
        for (auto &kv : staging_rings)
        {
            for (auto &e : kv.second->done)
                if (e)
                    cudaEventDestroy(e);
            cudaFreeHost(kv.second->host);
            delete kv.second;
        }
        for (auto &kv : streams)
        {
            ScratchArena *arena = nullptr;
            {
                std::lock_guard<std::mutex> ga(g_arena_mu);
                auto it = g_arenas.find(kv.second);
                if (it != g_arenas.end())
                {
                    arena = it->second;
                    g_arenas.erase(it);
                    g_arena_generation.fetch_add(1, std::memory_order_release);
                }
            }
            if (arena)
            {
                for (auto &ch : arena->chunks)
                    cudaFree(ch.first);
                delete arena;
            }
            cudaStreamDestroy(kv.second);
        }
        for (auto &kv : galois_tables)
            cudaFree(kv.second);
        for (auto &kv : hplans)
        {
            delete kv.second;
        }
        cudaFree(d_primes);
        cudaFree(d_tw);
        cudaFree(d_itw);
        cudaFree(d_inv);
    }

    ScratchArena *arena_of(cudaStream_t s)
    {
        static thread_local cudaStream_t last_stream = nullptr;
        static thread_local ScratchArena *last_arena = nullptr;
        static thread_local uint64_t last_generation = 0;
        const uint64_t generation = g_arena_generation.load(std::memory_order_acquire);
        if (last_arena && last_stream == s && last_generation == generation)
            return last_arena;
        std::lock_guard<std::mutex> g(g_arena_mu);
        auto it = g_arenas.find(s);
        if (it == g_arenas.end())
            return nullptr;
        last_stream = s;
        last_arena = it->second;
        last_generation = generation;
        return it->second;
    }

    void Context::activate() const
    {
        int cur = -1;
        cudaGetDevice(&cur);
        if (cur != device)
            BK_CUDA(cudaSetDevice(device));
    }

    cudaStream_t Context::stream()
    {
        activate();
        std::lock_guard<std::mutex> g(mu);
        auto id = std::this_thread::get_id();
        auto it = streams.find(id);
        if (it != streams.end())
            return it->second;
        cudaStream_t s;
        BK_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        streams[id] = s;
        {
            std::lock_guard<std::mutex> ga(g_arena_mu);
            g_arenas[s] = new ScratchArena();
        }
        return s;
    }

    void Context::debug_check()
    {
        cudaError_t e = cudaDeviceSynchronize();
        if (e == cudaSuccess)
            return;
        std::fprintf(stderr, "b200ckks: launch #%llu failed: %s\n", (unsigned long long)launches.load(), cudaGetErrorString(e));
        void *frames[24];
        int n = backtrace(frames, 24);
        backtrace_symbols_fd(frames, n, 2);
        std::abort();
    }

    cudaStream_t Context::stream_if_any()
    {
        std::lock_guard<std::mutex> g(mu);
        auto it = streams.find(std::this_thread::get_id());
        return it == streams.end() ? nullptr : it->second;
    }

    void Context::release_words(void *d, cudaStream_t owner)
    {
        if (!d)
            return;
        activate();
        cudaStream_t cur = stream_if_any();
        if (cur && (cur == owner || !owner))
        {
            cudaFreeAsync(d, cur);
            return;
        }
        cudaDeviceSynchronize();
        if (owner)
            cudaFreeAsync(d, owner);
        else
            cudaFree(d);
    }

    char *Context::staging(size_t bytes, cudaEvent_t *done_out)
    {
        activate();
        StagingRing *r;
        {
            std::lock_guard<std::mutex> g(mu);
            auto &slot = staging_rings[std::this_thread::get_id()];
            if (!slot)
                slot = new StagingRing();
            r = slot;
        }
        const size_t need = std::max(bytes, n * sizeof(u64)); // a full slot vector of complex doubles is N * 8 bytes
        if (r->slot_bytes < need)
        {
            if (r->host)
            {
                BK_CUDA(cudaStreamSynchronize(stream()));
                cudaFreeHost(r->host);
            }
            BK_CUDA(cudaMallocHost((void **)&r->host, need * StagingRing::SLOTS));
            r->slot_bytes = need;
            for (auto &e : r->done)
                if (!e)
                    BK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        }
        const int i = r->next;
        r->next = (i + 1) % StagingRing::SLOTS;
        BK_CUDA(cudaEventSynchronize(r->done[i])); // returns at once unless the ring has wrapped onto an in-flight copy
        *done_out = r->done[i];
        return r->host + (size_t)i * r->slot_bytes;
    }

    const uint32_t *Context::galois_table(uint32_t elt)
    {
        // GaloisTool caches one permutation table per element (galois.cpp:26-50)
        std::lock_guard<std::mutex> g(mu);
        auto it = galois_tables.find(elt);
        if (it != galois_tables.end())
            return it->second;
        std::vector<uint32_t> t(n);
        galois_table_ntt(log_n, elt, t.data());
        uint32_t *d;
        BK_CUDA(cudaMalloc((void **)&d, n * sizeof(uint32_t)));
        BK_CUDA(cudaMemcpy(d, t.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice));
        BK_CUDA(cudaDeviceSynchronize()); // see Context::Context: the copy must have landed before a kernel gathers through it
        galois_tables[elt] = d;
        return d;
    }

    bool Context::scale_in_bounds(double scale, int limbs) const
    {
        // is_scale_within_bounds (evaluator.cpp:29-50), CKKS branch
        return !(scale <= 0 || ((int)std::log2(scale) >= total_bits[limbs]));
    }

    int Context::ew_grid(size_t work_items) const
    {
        size_t blocks = (work_items + 255) / 256;
        size_t cap = (size_t)sm_count * 8;
        return (int)std::max<size_t>(1, std::min(blocks, cap));
    }

    JobMap limb_map(int limbs)
    {
        JobMap m;
        m.limbs = limbs;
        m.special_pos = -1;
        m.special_prime = 0;
        m.explicit_primes = nullptr;
        return m;
    }

    // ------------------------------------------------------------------------------ NTT drivers
#define BK_DISPATCH_LOGR(logn, ...)                                                                                  \
    switch (logn)                                                                                                      \
    {                                                                                                                  \
    case 12: { constexpr int LOGR = 4; __VA_ARGS__; } break;                                                                  \
    case 13: { constexpr int LOGR = 5; __VA_ARGS__; } break;                                                                  \
    case 14: { constexpr int LOGR = 6; __VA_ARGS__; } break;                                                                  \
    case 15: { constexpr int LOGR = 7; __VA_ARGS__; } break;                                                                  \
    default: { constexpr int LOGR = 8; __VA_ARGS__; } break;                                                                  \
    }

    // Launch geometry (measured with tools/lab/ntt_lab.cu, profiles/r2_ntt_lab.md): the passes are bound by the
    // integer multiplier and by how evenly the CTAs tile the 148 SMs, so small CTAs win at every size - column tiles
    // of 8 columns (128 threads, 32 CTAs per limb-polynomial) and block-pass CTAs of 128 threads (8 blocks of 256
    // coefficients), 64 threads when the whole launch would otherwise leave SMs empty.
    static unsigned block_pass_threads(const Context &c, int jobs)
    {
        return (long)jobs * 32 < 2L * c.sm_count ? 64u : 128u;
    }
    template <class Load>
    static void launch_fwd_cols(Context &c, cudaStream_t s, const Load &ld, u64 *out, int jobs)
    {
        if (jobs <= 0)
            return;
        ProfScope ps(c, s, TAG_FWD_COLS, jobs);
        dim3 grid(32, jobs);
        if (c.tables.wide == 1)
        {
            BK_DISPATCH_LOGR(c.log_n, launch_pdl(k_fwd_cols<LOGR, Load, 1, 8>, grid, 8 * ((1 << LOGR) / 16), 0, s, ld, out, c.tables));
        }
        else if (c.tables.wide == 2)
        {
            BK_DISPATCH_LOGR(c.log_n, launch_pdl(k_fwd_cols<LOGR, Load, 2, 8>, grid, 8 * ((1 << LOGR) / 16), 0, s, ld, out, c.tables));
        }
        else
        {
            BK_DISPATCH_LOGR(c.log_n, launch_pdl(k_fwd_cols<LOGR, Load, 0, 8>, grid, 8 * ((1 << LOGR) / 16), 0, s, ld, out, c.tables));
        }
        c.count();
    }
    template <class Store>
    static void launch_fwd_blocks(Context &c, cudaStream_t s, const u64 *in, const Store &st, int jobs)
    {
        if (jobs <= 0)
            return;
        const unsigned threads = block_pass_threads(c, jobs);
        dim3 grid((unsigned)(c.n / (16 * threads)), jobs);
        ProfScope ps(c, s, TAG_FWD_BLOCKS, jobs);
        if (c.tables.wide == 1)
            launch_pdl(k_fwd_blocks<Store, 1>, grid, threads, threads * 128, s, in, st, c.tables);
        else if (c.tables.wide == 2)
            launch_pdl(k_fwd_blocks<Store, 2>, grid, threads, threads * 128, s, in, st, c.tables);
        else
            launch_pdl(k_fwd_blocks<Store, 0>, grid, threads, threads * 128, s, in, st, c.tables);
        c.count();
    }
    template <class Load>
    static void launch_inv_blocks(Context &c, cudaStream_t s, const Load &ld, u64 *out, int jobs)
    {
        if (jobs <= 0)
            return;
        const unsigned threads = block_pass_threads(c, jobs);
        dim3 grid((unsigned)(c.n / (16 * threads)), jobs);
        ProfScope ps(c, s, TAG_INV_BLOCKS, jobs);
        launch_pdl(k_inv_blocks<Load>, grid, threads, threads * 128, s, ld, out, c.tables);
        c.count();
    }
    template <class Store>
    static void launch_inv_cols(Context &c, cudaStream_t s, const u64 *in, const Store &st, int jobs)
    {
        if (jobs <= 0)
            return;
        ProfScope ps(c, s, TAG_INV_COLS, jobs);
        dim3 grid(32, jobs);
        BK_DISPATCH_LOGR(c.log_n, launch_pdl(k_inv_cols<LOGR, Store, 8>, grid, 8 * ((1 << LOGR) / 16), 0, s, in, st, c.tables));
        c.count();
    }

    void ntt_fwd(Context &c, cudaStream_t s, u64 *data, int jobs, JobMap map)
    {
        Scratch tmp(s, (size_t)jobs * c.n);
        LdPlain ld{ data, map, c.n };
        launch_fwd_cols(c, s, ld, tmp.p, jobs);
        StPlain st{ data, map, c.n };
        launch_fwd_blocks(c, s, tmp.p, st, jobs);
    }

    void ntt_inv(Context &c, cudaStream_t s, u64 *data, int jobs, JobMap map)
    {
        Scratch tmp(s, (size_t)jobs * c.n);
        LdInvPlain ld{ data, map, c.n, nullptr };
        launch_inv_blocks(c, s, ld, tmp.p, jobs);
        StInvPlain st{ data, map, c.n };
        launch_inv_cols(c, s, tmp.p, st, jobs);
    }

    // ------------------------------------------------------------------------------- containers
    static void realloc_words(Context &c, u64 *&d, cudaStream_t &owner, size_t &cap, size_t words, bool keep, size_t keep_words)
    {
        if (words <= cap)
            return;
        cudaStream_t s = c.stream();
        u64 *nd;
        BK_CUDA(cudaMallocAsync((void **)&nd, words * sizeof(u64), s));
        if (d)
        {
            if (keep && keep_words)
                BK_CUDA(cudaMemcpyAsync(nd, d, keep_words * sizeof(u64), cudaMemcpyDeviceToDevice, s));
            c.release_words(d, owner);
        }
        d = nd;
        owner = s;
        cap = words;
    }

    void ensure_ct(bk_ct_t ct, int size, int limbs, bool keep)
    {
        Context &c = *ct->ctx;
        size_t words = (size_t)size * limbs * c.n;
        realloc_words(c, ct->d, ct->owner, ct->cap, words, keep, (size_t)ct->size * ct->limbs * c.n);
        ct->size = size;
        ct->limbs = limbs;
    }

    void ensure_pt(bk_pt_t pt, int limbs)
    {
        Context &c = *pt->ctx;
        realloc_words(c, pt->d, pt->owner, pt->cap, (size_t)limbs * c.n, false, 0);
        pt->limbs = limbs;
    }

    // replace ct's buffer by a freshly produced one
    static void adopt(bk_ct_t ct, u64 *nd, size_t cap, int size, int limbs)
    {
        cudaStream_t s = ct->ctx->stream();
        ct->ctx->release_words(ct->d, ct->owner);
        ct->d = nd;
        ct->owner = s; // alloc_words() below allocates on the calling thread's stream
        ct->cap = cap;
        ct->size = size;
        ct->limbs = limbs;
    }

    static u64 *alloc_words(Context &c, size_t words)
    {
        u64 *p;
        BK_CUDA(cudaMallocAsync((void **)&p, words * sizeof(u64), c.stream()));
        return p;
    }

    static void check_ct(const Context *ctx, bk_ct_t a, const char *name)
    {
        if (!a || a->ctx != ctx || !a->d || a->size < 2 || a->limbs < 1 || a->limbs > ctx->top_limbs())
            throw std::invalid_argument(std::string(name) + " is not valid for encryption parameters");
    }

    static bool close_scale(double a, double b)
    {
        // util::are_close<double> (util/common.h:569-573)
        double f = std::max({ std::fabs(a), std::fabs(b), 1.0 });
        return std::fabs(a - b) < std::numeric_limits<double>::epsilon() * f;
    }

    // threads per CTA of k_ks_mac: 128-thread CTAs (two coefficients per thread) make a 4-modulus chunk 1024 CTAs, all
    // resident at once on 148 SMs (7 per SM at 66 registers) - 256-thread CTAs left a nearly empty second wave
    static int ks_mac_threads()
    {
        static const int t = [] {
            const char *e = std::getenv("B200CKKS_KS_MAC_THREADS");
            int v = e ? std::atoi(e) : 128;
            return (v == 64 || v == 128 || v == 256) ? v : 128;
        }();
        return t;
    }
#define KS_MAC_THREADS ks_mac_threads()

    // ------------------------------------------------------------------- hybrid key switch (tolerance mode)
    // Shape of the key switch at level l of a chain with `top` data primes: alpha special moduli (the special prime plus
    // alpha - 1 idle primes above the level; at most top - l + 1 exist) and digits of dsize = alpha - 1 primes, so that
    // P_S exceeds every digit by a whole prime and the key-switching noise stays at the level of SEAL's 46-bit digits
    // over a 51-bit special prime.  The pair minimises the NTT-equivalent work: l inverse NTTs, dnum (l + alpha) - l digit
    // NTTs, the basis conversions (dsize multiply-adds per converted coefficient, ~1/8 of an NTT each) and the ModDown
    // over alpha moduli.  alpha = 1 is SEAL's own scheme (one digit per prime); it is kept at the top level, where no
    // prime is idle, and at l <= 5, where a key switch is a handful of latency-bound launches either way.
    void hybrid_shape(int l, int top, int &alpha, int &dsize)
    {
        alpha = 1;
        dsize = 1;
        if (l <= 5)
            return;
        double best_cost = (double)l + (double)l * (l + 1) - l + l * (double)l / 8.0 + 2.0 + 2.0 * l / 8.0 + 2.0 * l;
        // At the two levels below the top only one or two primes are idle; there a digit may be as large as P_S
        // (dsize = alpha), which makes the noise that of SEAL's own top-level key switch (51-bit digits over a 51-bit
        // special prime) instead of a prime below it, and halves / thirds the digit count.  Bootstrapping error with it:
        // rms 4.3e-6 against 1.6e-5 on the reference-exact path and 3.0e-6 with narrow digits everywhere
        // (tools/boot_precision.py).  $B200CKKS_HYBRID_NARROW_TOP=1 keeps dsize = alpha - 1 at every level.
        static const bool wide_top = std::getenv("B200CKKS_HYBRID_NARROW_TOP") == nullptr;
        for (int a = 2; a <= std::min(top - l + 1, 17); a++)
        {
            for (int ds = a - 1; ds <= ((wide_top && l >= top - 2) ? a : a - 1); ds++)
            {
                const int d = (l + ds - 1) / ds;
                double cost = l + (double)d * (l + a) - l + d * ds * (double)l / 8.0 + 2.0 * a + 2.0 * a * l / 8.0 + 2.0 * l;
                if (cost < best_cost - 1e-9)
                {
                    best_cost = cost;
                    alpha = a;
                    dsize = ds;
                }
            }
        }
    }

    template <class T> static T *upload_sync(const std::vector<T> &v)
    {
        T *d = nullptr;
        BK_CUDA(cudaMalloc((void **)&d, std::max<size_t>(1, v.size()) * sizeof(T)));
        if (!v.empty())
            BK_CUDA(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
        return d;
    }

    const HybridPlan &hybrid_plan(Context &c, int l)
    {
        {
            std::lock_guard<std::mutex> g(c.mu);
            auto it = c.hplans.find(l);
            if (it != c.hplans.end())
                return *it->second;
        }
        c.activate();
        const int top = c.top_limbs(), sp = c.n_primes - 1;
        auto P = std::make_unique<HybridPlan>();
        P->l = l;
        hybrid_shape(l, top, P->alpha, P->dsize);
        P->dnum = (l + P->dsize - 1) / P->dsize;
        P->ne = l + P->alpha;
        const int alpha = P->alpha, dsize = P->dsize, dnum = P->dnum, ne = P->ne;
        auto eprime = [&](int e) { return e < l + alpha - 1 ? e : sp; };
        auto q = [&](int prime_index) { return c.primes[(size_t)prime_index]; };
        // product of the digit's primes except `skip`, modulo m
        auto digit_hat = [&](int d, int skip, uint64_t m) {
            uint64_t r = 1 % m;
            for (int j = d * dsize; j < std::min((d + 1) * dsize, l); j++)
                if (j != skip)
                    r = bk::mulmod(r, q(j) % m, m);
            return r;
        };
        auto special_hat = [&](int skip, uint64_t m) {
            uint64_t r = 1 % m;
            for (int b = 0; b < alpha; b++)
                if (b != skip)
                    r = bk::mulmod(r, q(eprime(l + b)) % m, m);
            return r;
        };
        std::vector<ulonglong2> prescale((size_t)l), sprescale((size_t)2 * alpha), psinv((size_t)l);
        std::vector<int> limb_primes((size_t)l), sprimes((size_t)2 * alpha);
        std::vector<u64> w((size_t)ne * dnum * dsize, 0), ws((size_t)l * alpha), keyfactor((size_t)dnum * ne, 0);
        for (int i = 0; i < l; i++)
        {
            uint64_t v = bk::invmod(digit_hat(i / dsize, i, q(i)), q(i));
            prescale[(size_t)i] = make_ulonglong2(v, bk::shoup(v, q(i)));
            limb_primes[(size_t)i] = i;
            uint64_t ps = special_hat(-1, q(i));
            uint64_t pi = bk::invmod(ps, q(i));
            psinv[(size_t)i] = make_ulonglong2(pi, bk::shoup(pi, q(i)));
            for (int a = 0; a < alpha; a++)
                ws[(size_t)i * alpha + a] = special_hat(a, q(i));
            keyfactor[(size_t)(i / dsize) * ne + i] = ps;
        }
        for (int e = 0; e < ne; e++)
            for (int d = 0; d < dnum; d++)
                for (int a = 0; a < dsize && d * dsize + a < l; a++)
                    w[((size_t)e * dnum + d) * dsize + a] = digit_hat(d, d * dsize + a, q(eprime(e)));
        for (int p = 0; p < 2; p++)
            for (int a = 0; a < alpha; a++)
            {
                uint64_t m = q(eprime(l + a));
                uint64_t v = bk::invmod(special_hat(a, m), m);
                sprescale[(size_t)p * alpha + a] = make_ulonglong2(v, bk::shoup(v, m));
                sprimes[(size_t)p * alpha + a] = eprime(l + a);
            }
        P->d_prescale = upload_sync(prescale);
        P->d_limb_primes = upload_sync(limb_primes);
        P->d_w = upload_sync(w);
        P->d_sprescale = upload_sync(sprescale);
        P->d_sprimes = upload_sync(sprimes);
        P->d_ws = upload_sync(ws);
        P->d_psinv = upload_sync(psinv);
        P->d_keyfactor = upload_sync(keyfactor);
        // rounding and exact-conversion constants of a division by D = the product of `dropped` onto `targets` limbs
        auto rounding_tables = [&](const std::vector<int> &dropped, int targets, u64 *&d_half, double *&d_pinv, u64 *&d_negd, u64 *&d_addc) {
            const int cnt = (int)dropped.size();
            auto dmod = [&](uint64_t m) {
                uint64_t r = 1 % m;
                for (int b = 0; b < cnt; b++)
                    r = bk::mulmod(r, q(dropped[(size_t)b]) % m, m);
                return r;
            };
            // floor(D / 2) = (D - 1) / 2 mod an odd m
            auto half_mod = [&](uint64_t m) { return bk::mulmod((dmod(m) + m - 1) % m, (m + 1) / 2, m); };
            std::vector<u64> half((size_t)2 * cnt), negd((size_t)targets), addc((size_t)targets);
            std::vector<double> pinv((size_t)cnt);
            for (int a = 0; a < cnt; a++)
            {
                half[(size_t)a] = half[(size_t)cnt + a] = half_mod(q(dropped[(size_t)a]));
                pinv[(size_t)a] = 1.0 / (double)q(dropped[(size_t)a]);
            }
            for (int i = 0; i < targets; i++)
            {
                negd[(size_t)i] = (q(i) - dmod(q(i))) % q(i);
                addc[(size_t)i] = (q(i) - half_mod(q(i))) % q(i);
            }
            d_half = upload_sync(half);
            d_pinv = upload_sync(pinv);
            d_negd = upload_sync(negd);
            d_addc = upload_sync(addc);
        };
        {
            std::vector<int> dropped;
            for (int a = 0; a < alpha; a++)
                dropped.push_back(eprime(l + a));
            rounding_tables(dropped, l, P->d_shalf, P->d_spinv, P->d_negd, P->d_addc);
        }
        if (alpha + 1 <= 17 && l >= 2)
        {
            // dropped basis of the merged ModDown + rescale: D_0 = q_{l-1}, D_{1+a} = special modulus a
            const int da = alpha + 1, lo = l - 1;
            auto dprime = [&](int a) { return a == 0 ? lo : eprime(l + a - 1); };
            auto dropped_hat = [&](int skip, uint64_t m) {
                uint64_t r = 1 % m;
                for (int b = 0; b < da; b++)
                    if (b != skip)
                        r = bk::mulmod(r, q(dprime(b)) % m, m);
                return r;
            };
            std::vector<ulonglong2> rpre((size_t)2 * da), dinv((size_t)lo), qlinv((size_t)lo);
            std::vector<int> rprimes((size_t)2 * da);
            std::vector<u64> rws((size_t)lo * da);
            for (int p = 0; p < 2; p++)
                for (int a = 0; a < da; a++)
                {
                    uint64_t m = q(dprime(a));
                    uint64_t v = bk::invmod(dropped_hat(a, m), m);
                    rpre[(size_t)p * da + a] = make_ulonglong2(v, bk::shoup(v, m));
                    rprimes[(size_t)p * da + a] = dprime(a);
                }
            for (int i = 0; i < lo; i++)
            {
                for (int a = 0; a < da; a++)
                    rws[(size_t)i * da + a] = dropped_hat(a, q(i));
                uint64_t v = bk::invmod(dropped_hat(-1, q(i)), q(i));
                dinv[(size_t)i] = make_ulonglong2(v, bk::shoup(v, q(i)));
                uint64_t w = bk::invmod(q(lo) % q(i), q(i));
                qlinv[(size_t)i] = make_ulonglong2(w, bk::shoup(w, q(i)));
            }
            uint64_t pm = special_hat(-1, q(lo));
            P->r_pmod = make_ulonglong2(pm, bk::shoup(pm, q(lo)));
            P->d_r_sprescale = upload_sync(rpre);
            P->d_r_sprimes = upload_sync(rprimes);
            P->d_r_ws = upload_sync(rws);
            P->d_r_dinv = upload_sync(dinv);
            P->d_r_qlinv = upload_sync(qlinv);
            std::vector<int> dropped;
            for (int a = 0; a < da; a++)
                dropped.push_back(dprime(a));
            rounding_tables(dropped, lo, P->d_r_shalf, P->d_r_spinv, P->d_r_negd, P->d_r_addc);
            P->rescale_tables = true;
        }
        BK_CUDA(cudaDeviceSynchronize()); // pageable uploads must have landed before any stream reads them
        std::lock_guard<std::mutex> g(c.mu);
        auto ins = c.hplans.emplace(l, P.get());
        if (ins.second)
            P.release();
        return *ins.first->second;
    }

    static void launch_hyb_conv(Context &c, cudaStream_t s, const HybConvArgs &a, int groups, int ds)
    {
        dim3 grid((unsigned)((c.n / 2 + 127) / 128), (unsigned)groups, (unsigned)((a.nT + HYB_CONV_TARGETS - 1) / HYB_CONV_TARGETS));
        ProfScope ps(c, s, TAG_OTHER, groups * a.nT);
        switch (ds)
        {
#define BK_HYB_CONV_CASE(DS)                                                                                           \
    case DS:                                                                                                           \
        if (a.down)                                                                                                    \
            launch_pdl(k_hyb_conv<DS, true>, grid, 128, 0, s, a, c.tables);                                            \
        else                                                                                                           \
            launch_pdl(k_hyb_conv<DS, false>, grid, 128, 0, s, a, c.tables);                                           \
        break;
            BK_HYB_CONV_CASE(1) BK_HYB_CONV_CASE(2) BK_HYB_CONV_CASE(3) BK_HYB_CONV_CASE(4) BK_HYB_CONV_CASE(5)
            BK_HYB_CONV_CASE(6) BK_HYB_CONV_CASE(7) BK_HYB_CONV_CASE(8) BK_HYB_CONV_CASE(9) BK_HYB_CONV_CASE(10)
            BK_HYB_CONV_CASE(11) BK_HYB_CONV_CASE(12) BK_HYB_CONV_CASE(13) BK_HYB_CONV_CASE(14) BK_HYB_CONV_CASE(15)
            BK_HYB_CONV_CASE(16) BK_HYB_CONV_CASE(17)
#undef BK_HYB_CONV_CASE
        default: throw std::logic_error("hybrid key switch: digit size out of range");
        }
        c.count();
    }

    static int hyb_chunk(const HybridPlan &P)
    {
        // digit NTT outputs of one chunk stay in L2: at most ~96 limb-polynomials (48 MiB at N = 2^16)
        return std::max(1, std::min(P.ne, 96 / P.dnum));
    }

    // steps B and C for one accumulator: digits of `y` extended and multiplied into `hk`, then ModDown into out
    static void hyb_extend_and_mac(Context &c, cudaStream_t s, const HybridPlan &P, const HybDims &h, const u64 *y, u64 *conv, u64 *inter,
                                   const u64 *target_ntt, int count, const uint32_t *const *perms, bk_hybkey_s *const *keys,
                                   u64 *acc /*[count][2][ne][N]*/, int gather)
    {
        const size_t n = c.n;
        const int chunk = hyb_chunk(P);
        for (int e0 = 0; e0 < P.ne; e0 += chunk)
        {
            const int nE = std::min(chunk, P.ne - e0);
            HybConvArgs cv{ y, P.d_w, conv, n, h, e0, nE, 0, nullptr, nullptr, nullptr };
            launch_hyb_conv(c, s, cv, P.dnum, P.dsize);
            LdHybPlain ld{ conv, n, h, e0 };
            launch_fwd_cols(c, s, ld, inter, nE * P.dnum);
            StHybDigit st{ inter, n, h, e0 };
            launch_fwd_blocks(c, s, inter, st, nE * P.dnum);
            // seed-compressed keys (all keys of a context are, or none): the uniform halves of this chunk's limbs are
            // regenerated into scratch first, a few rotations at a time to bound the scratch
            const bool compressed = keys[0]->compressed;
            const int batch = compressed ? 4 : HYB_MAC_BATCH;
            const size_t kstride = (size_t)P.ne * n;
            Scratch halves(s, compressed ? (size_t)std::min(batch, count) * P.dnum * nE * n : 0);
            for (int k0 = 0; k0 < count; k0 += batch)
            {
                const int nk = std::min(batch, count - k0);
                HybMacArgs a{};
                a.digits = inter;
                a.target_ntt = target_ntt;
                a.dstride0 = compressed ? kstride : 2 * kstride;
                a.dstride1 = compressed ? (size_t)nE * n : 2 * kstride;
                for (int k = 0; k < nk; k++)
                {
                    a.perm[k] = perms[k0 + k];
                    a.key[k] = keys[k0 + k]->d;
                    if (compressed)
                    {
                        u64 *buf = halves.p + (size_t)k * P.dnum * nE * n;
                        expand_public_halves(c, s, keys[k0 + k], e0, nE, buf);
                        a.key1[k] = buf - (size_t)e0 * n; // the kernels add e * N with e = e0 + local limb
                    }
                    else
                        a.key1[k] = keys[k0 + k]->d + kstride;
                }
                a.acc = acc + (size_t)k0 * 2 * P.ne * n;
                a.n = n;
                a.h = h;
                a.e0 = e0;
                a.gather_digits = gather;
                // key stream through the bulk-copy engine (cp.async.bulk + mbarrier ring) unless $B200CKKS_MAC_BULK=0
                static const bool bulk = [] {
                    const char *e = std::getenv("B200CKKS_MAC_BULK");
                    return !e || std::atoi(e) != 0;
                }();
                {
                    ProfScope ps(c, s, TAG_KS_MAC, nE * 2 * P.dnum * nk);
                    if (bulk && n % KS_BULK_TILE == 0)
                    {
                        dim3 grid((unsigned)(n / KS_BULK_TILE), nE, nk);
                        launch_pdl(k_ks_mac_hyb_bulk, grid, 128, 0, s, a, c.tables);
                    }
                    else
                    {
                        dim3 grid((unsigned)((n / 2 + KS_MAC_THREADS - 1) / KS_MAC_THREADS), nE, nk);
                        launch_pdl(k_ks_mac_hyb, grid, KS_MAC_THREADS, 0, s, a, c.tables);
                    }
                }
                c.count();
            }
        }
    }

    static void hyb_mod_down(Context &c, cudaStream_t s, const HybridPlan &P, const HybDims &h, const u64 *acc, u64 *conv, u64 *inter, u64 *tl,
                             u64 *out, const u64 *base0, const u64 *base1, const uint32_t *perm)
    {
        const size_t n = c.n;
        LdInvSpecials ld{ acc, n, h };
        launch_inv_blocks(c, s, ld, inter, 2 * P.alpha);
        StInvScaledAdd st{ tl, n, P.d_sprescale, P.d_sprimes, P.d_shalf };
        launch_inv_cols(c, s, inter, st, 2 * P.alpha);
        HybConvArgs cv{ tl, P.d_ws, conv, n, h, 0, P.l, 1, P.d_spinv, P.d_negd, P.d_addc };
        launch_hyb_conv(c, s, cv, 2, P.alpha);
        launch_fwd_cols(c, s, LdPlain{ conv, limb_map(P.l), n }, inter, 2 * P.l);
        StModDown st2{ acc, out, base0, base1, perm, P.d_psinv, n, P.l, P.ne };
        launch_fwd_blocks(c, s, inter, st2, 2 * P.l);
    }

    // The same ModDown with the rescale that follows a relinearization folded in: one division by D = q_{l-1} P_S
    // instead of one by P_S and one by q_{l-1}.  With x = acc + P_S base over the extended basis (P_S base vanishes on
    // the special limbs, so only limb l-1 of the dropped ones sees the base),
    //   out_i = (x_i - [x]_D) D^-1 = (acc_i - conv_i) D^-1 + base_i q_{l-1}^-1   (mod q_i, i < l-1).
    // Saves the 2l forward transforms of the intermediate result and the rescale's two inverse ones; the quotient is
    // rounded to nearest and the basis conversion exact (HybridPlan::d_shalf), as in hyb_mod_down.
    static void hyb_mod_down_rescale(Context &c, cudaStream_t s, const HybridPlan &P, const HybDims &h, const u64 *acc, u64 *conv, u64 *inter,
                                     u64 *tl, u64 *out, const u64 *base0, const u64 *base1)
    {
        const size_t n = c.n;
        const int da = P.alpha + 1, lo = P.l - 1;
        LdInvDropped ld{ acc, base0, base1, n, h, P.r_pmod };
        launch_inv_blocks(c, s, ld, inter, 2 * da);
        StInvScaledAdd st{ tl, n, P.d_r_sprescale, P.d_r_sprimes, P.d_r_shalf };
        launch_inv_cols(c, s, inter, st, 2 * da);
        const HybDims hd{ lo, da, h.dsize, h.dnum, h.special_prime };
        HybConvArgs cv{ tl, P.d_r_ws, conv, n, hd, 0, lo, 1, P.d_r_spinv, P.d_r_negd, P.d_r_addc };
        launch_hyb_conv(c, s, cv, 2, da);
        launch_fwd_cols(c, s, LdPlain{ conv, limb_map(lo), n }, inter, 2 * lo);
        StModDownRescale st2{ acc, out, base0, base1, nullptr, P.d_r_dinv, n, lo, P.ne, P.d_r_qlinv };
        launch_fwd_blocks(c, s, inter, st2, 2 * lo);
    }

    // polynomial-1 pointer and digit strides of a SEAL-shaped key for k_ks_mac; a seed-compressed level key gets its
    // uniform halves for output moduli [I0, I0 + nI) expanded into `halves` first
    static void classic_key_halves(Context &c, cudaStream_t s, const bk_kskey_s *key, int I0, int nI, u64 *halves, KsMacArgs &a)
    {
        const size_t kstride = (size_t)(key->klimbs + 1) * c.n;
        if (key->view_of && key->view_of->compressed)
        {
            expand_public_halves(c, s, key->view_of, I0, nI, halves);
            a.key1 = halves - (size_t)I0 * c.n;
            a.dstride0 = kstride;
            a.dstride1 = (size_t)nI * c.n;
        }
        else
        {
            a.key1 = key->d + kstride;
            a.dstride0 = a.dstride1 = 2 * kstride;
        }
    }

    // Same contract as key_switch below; the key is the level-l hybrid key of `key`'s recipe.
    // rescale: out is [2][l-1][N] and receives the result divided by q_{l-1} as well (perm must be null).
    static void key_switch_hybrid(Context &c, cudaStream_t s, const u64 *target, const uint32_t *perm, const u64 *base0,
                                  const u64 *base1, u64 *out, int l, bk_kskey_t key, bool rescale = false)
    {
        const HybridPlan &P = hybrid_plan(c, l);
        bk_hybkey_s *hk = hybrid_key(c, key, l);
        const size_t n = c.n;
        const HybDims h{ l, P.alpha, P.dsize, P.dnum, c.n_primes - 1 };
        Scratch y(s, (size_t)l * n);
        Scratch inter(s, (size_t)std::max({ hyb_chunk(P) * P.dnum, 2 * l, 2 * P.alpha + 2 }) * n);
        Scratch conv(s, (size_t)std::max(hyb_chunk(P) * P.dnum, 2 * l) * n);
        Scratch acc(s, (size_t)2 * P.ne * n);
        Scratch tl(s, (size_t)(2 * P.alpha + 2) * n);
        {
            LdInvPlain ld{ target, limb_map(l), n, perm };
            launch_inv_blocks(c, s, ld, inter.p, l);
            StInvScaled st{ y.p, n, P.d_prescale, P.d_limb_primes };
            launch_inv_cols(c, s, inter.p, st, l);
        }
        hyb_extend_and_mac(c, s, P, h, y.p, conv.p, inter.p, target, 1, &perm, &hk, acc.p, 0);
        if (rescale)
        {
            if (perm || !base0 || !P.rescale_tables)
                throw std::logic_error("merged ModDown and rescale: relinearization at a hybrid level only");
            hyb_mod_down_rescale(c, s, P, h, acc.p, conv.p, inter.p, tl.p, out, base0, base1);
        }
        else
            hyb_mod_down(c, s, P, h, acc.p, conv.p, inter.p, tl.p, out, base0, base1, perm);
    }

    static void key_switch_hoisted_hybrid(Context &c, cudaStream_t s, const bk_ct_s *in, int count, const uint32_t *const *perms,
                                          const bk_kskey_t *keys, u64 *const *outs)
    {
        const int l = in->limbs;
        const HybridPlan &P = hybrid_plan(c, l);
        std::vector<bk_hybkey_s *> hks((size_t)count);
        for (int k = 0; k < count; k++)
            hks[(size_t)k] = hybrid_key(c, keys[k], l);
        const size_t n = c.n;
        const HybDims h{ l, P.alpha, P.dsize, P.dnum, c.n_primes - 1 };
        const u64 *c0 = in->d, *c1 = in->d + (size_t)l * n;
        Scratch y(s, (size_t)l * n);
        Scratch inter(s, (size_t)std::max({ hyb_chunk(P) * P.dnum, 2 * l, 2 * P.alpha }) * n);
        Scratch conv(s, (size_t)std::max(hyb_chunk(P) * P.dnum, 2 * l) * n);
        Scratch acc(s, (size_t)count * 2 * P.ne * n);
        Scratch tl(s, (size_t)2 * P.alpha * n);
        {
            LdInvPlain ld{ c1, limb_map(l), n, nullptr };
            launch_inv_blocks(c, s, ld, inter.p, l);
            StInvScaled st{ y.p, n, P.d_prescale, P.d_limb_primes };
            launch_inv_cols(c, s, inter.p, st, l);
        }
        hyb_extend_and_mac(c, s, P, h, y.p, conv.p, inter.p, c1, count, perms, hks.data(), acc.p, 1);
        for (int k = 0; k < count; k++)
            hyb_mod_down(c, s, P, h, acc.p + (size_t)k * 2 * P.ne * n, conv.p, inter.p, tl.p, outs[k], c0, nullptr, perms[k]);
    }

    // ------------------------------------------------------------------------------- key switch
    // Evaluator::switch_key_inplace (evaluator.cpp:2281-2525) with the Galois permutation of
    // apply_galois_inplace (:2191-2207) fused in.  target: [l][N] NTT form.  If perm != null the
    // target is gathered through it and base0 (= c0) likewise (rotation); out[2][l][N] receives
    //   out0 = perm(base0) + ModDown(sum_J d_J * key[J][0]),  out1 = base1 + ModDown(... key[J][1]).
    static void key_switch(Context &c, cudaStream_t s, const u64 *target, const uint32_t *perm, const u64 *base0,
                           const u64 *base1, u64 *out, int l, bk_kskey_t key)
    {
        bk_kskey_s level_view; // hybrid mode at a level whose shape is SEAL's own: its level key is a pruned SEAL key
        if (key->recipe) // a recipe generated in hybrid mode has no SEAL-layout data
        {
            if (hybrid_plan(c, l).alpha > 1)
            {
                key_switch_hybrid(c, s, target, perm, base0, base1, out, l, key);
                return;
            }
            bk_hybkey_s *hk = hybrid_key(c, key, l);
            level_view.ctx = key->ctx;
            level_view.d = hk->d;
            level_view.view_of = hk;
            level_view.digits = level_view.klimbs = l;
            key = &level_view;
        }
        if (key->digits < l || key->klimbs < l)
            throw std::invalid_argument("kswitch_keys is not valid for encryption parameters (key pruned below "
                                        "this level)");
        const size_t n = c.n;
        const int sp = c.n_primes - 1;
        const int chunk = std::max(1, std::min(c.ks_chunk, l + 1));
        Scratch ttarget(s, (size_t)l * n);
        Scratch inter(s, (size_t)std::max(chunk * l, 2 * l) * n);
        Scratch acc(s, (size_t)2 * (l + 1) * n);
        Scratch tlast(s, 2 * n);

        // A. t_target = INTT(perm(target))   (:2358-2365)
        {
            LdInvPlain ld{ target, limb_map(l), n, perm };
            launch_inv_blocks(c, s, ld, inter.p, l);
            StInvPlain st{ ttarget.p, limb_map(l), n };
            launch_inv_cols(c, s, inter.p, st, l);
        }
        // B. per output modulus: decompose + NTT every digit (two passes, intermediates stay in L2),
        //    then stream the key through the multiply-accumulate (:2368-2463)
        for (int I0 = 0; I0 <= l; I0 += chunk)
        {
            int nI = std::min(chunk, l + 1 - I0);
            LdKsDigit ld{ ttarget.p, c.d_primes, n, l, I0, sp };
            launch_fwd_cols(c, s, ld, inter.p, nI * l);
            StKsDigit st{ inter.p, n, l, I0, sp };
            launch_fwd_blocks(c, s, inter.p, st, nI * l);
            KsMacArgs a{ inter.p, target, perm, key->d, acc.p, n, l, I0, sp, key->klimbs, 0 };
            const bool comp = key->view_of && key->view_of->compressed;
            Scratch halves(s, comp ? (size_t)l * nI * n : 0);
            classic_key_halves(c, s, key, I0, nI, halves.p, a);
            dim3 grid((unsigned)((n / 2 + KS_MAC_THREADS - 1) / KS_MAC_THREADS), nI);
            {
                ProfScope ps(c, s, TAG_KS_MAC, nI * (2 * l + 2));
                launch_pdl(k_ks_mac, grid, KS_MAC_THREADS, 0, s, a, c.tables);
            }
            c.count();
        }
        // C. ModDown by the special prime (:2465-2523)
        {
            LdInvLimbOf ld{ acc.p, n, l + 1, l, sp };
            launch_inv_blocks(c, s, ld, inter.p, 2);
            StInvAddHalf st{ tlast.p, n, sp };
            launch_inv_cols(c, s, inter.p, st, 2);
            LdDivRound ld2{ tlast.p, n, l, c.primes[sp] };
            launch_fwd_cols(c, s, ld2, inter.p, 2 * l);
            StModDown st2{ acc.p, out, base0, base1, perm, c.inv_last(sp), n, l };
            launch_fwd_blocks(c, s, inter.p, st2, 2 * l);
        }
    }

    // Hoisted rotations (Halevi-Shoup): one decomposition + digit NTT of c1 serves every automorphism of the set; per
    // element only the inner product (digits read through the Galois table), ModDown and the gathered c0 remain.
    // The digits are those of c1, not of sigma(c1): a valid decomposition of sigma(c1) with the same bounds, but NOT
    // the residues Evaluator::rotate_vector produces (evaluator.cpp:2191-2214 permutes first) - results decrypt to the
    // same values up to key-switching noise, limbs differ.  Opt-in through bk_rotate_hoisted only.
    static void key_switch_hoisted(Context &c, cudaStream_t s, const bk_ct_s *in, int count, const uint32_t *const *perms,
                                   const bk_kskey_t *keys, u64 *const *outs)
    {
        const int l = in->limbs;
        std::vector<bk_kskey_s> level_views;
        std::vector<bk_kskey_t> level_keys;
        if (count > 0 && keys[0]->recipe)
        {
            if (hybrid_plan(c, l).alpha > 1)
            {
                key_switch_hoisted_hybrid(c, s, in, count, perms, keys, outs);
                return;
            }
            level_views = std::vector<bk_kskey_s>((size_t)count);
            for (int k = 0; k < count; k++)
            {
                bk_hybkey_s *hk = hybrid_key(c, keys[k], l);
                level_views[(size_t)k].ctx = keys[k]->ctx;
                level_views[(size_t)k].d = hk->d;
                level_views[(size_t)k].view_of = hk;
                level_views[(size_t)k].digits = level_views[(size_t)k].klimbs = l;
                level_keys.push_back(&level_views[(size_t)k]);
            }
            keys = level_keys.data();
        }
        for (int k = 0; k < count; k++)
            if (keys[k]->digits < l || keys[k]->klimbs < l)
                throw std::invalid_argument("kswitch_keys is not valid for encryption parameters (key pruned below "
                                            "this level)");
        const size_t n = c.n;
        const int sp = c.n_primes - 1;
        const int chunk = std::max(1, std::min(c.ks_chunk, l + 1));
        const u64 *c0 = in->d, *c1 = in->d + (size_t)l * n;
        Scratch ttarget(s, (size_t)l * n);
        Scratch inter(s, (size_t)std::max(chunk * l, 2 * l) * n);
        Scratch acc(s, (size_t)count * 2 * (l + 1) * n);
        Scratch tlast(s, 2 * n);
        {
            LdInvPlain ld{ c1, limb_map(l), n, nullptr };
            launch_inv_blocks(c, s, ld, inter.p, l);
            StInvPlain st{ ttarget.p, limb_map(l), n };
            launch_inv_cols(c, s, inter.p, st, l);
        }
        for (int I0 = 0; I0 <= l; I0 += chunk)
        {
            int nI = std::min(chunk, l + 1 - I0);
            LdKsDigit ld{ ttarget.p, c.d_primes, n, l, I0, sp };
            launch_fwd_cols(c, s, ld, inter.p, nI * l);
            StKsDigit st{ inter.p, n, l, I0, sp };
            launch_fwd_blocks(c, s, inter.p, st, nI * l);
            for (int k = 0; k < count; k++)
            {
                KsMacArgs a{ inter.p, c1, perms[k], keys[k]->d, acc.p + (size_t)k * 2 * (l + 1) * n, n, l, I0, sp,
                             keys[k]->klimbs, 1 };
                const bool comp = keys[k]->view_of && keys[k]->view_of->compressed;
                Scratch halves(s, comp ? (size_t)l * nI * n : 0);
                classic_key_halves(c, s, keys[k], I0, nI, halves.p, a);
                dim3 grid((unsigned)((n / 2 + KS_MAC_THREADS - 1) / KS_MAC_THREADS), nI);
                {
                    ProfScope ps(c, s, TAG_KS_MAC, nI * (2 * l + 2));
                    launch_pdl(k_ks_mac, grid, KS_MAC_THREADS, 0, s, a, c.tables);
                }
                c.count();
            }
        }
        for (int k = 0; k < count; k++)
        {
            u64 *ak = acc.p + (size_t)k * 2 * (l + 1) * n;
            LdInvLimbOf ld{ ak, n, l + 1, l, sp };
            launch_inv_blocks(c, s, ld, inter.p, 2);
            StInvAddHalf st{ tlast.p, n, sp };
            launch_inv_cols(c, s, inter.p, st, 2);
            LdDivRound ld2{ tlast.p, n, l, c.primes[sp] };
            launch_fwd_cols(c, s, ld2, inter.p, 2 * l);
            StModDown st2{ ak, outs[k], c0, nullptr, perms[k], c.inv_last(sp), n, l };
            launch_fwd_blocks(c, s, inter.p, st2, 2 * l);
        }
    }

    static bk_kskey_t find_gkey(bk_gkeys_t gk, uint32_t elt)
    {
        std::lock_guard<std::mutex> g(gk->mu);
        auto it = gk->keys.find(elt);
        return it == gk->keys.end() ? nullptr : it->second;
    }

    // dst == nullptr or dst == a: in place; otherwise `a` is left untouched and dst receives the result (no copy of `a`)
    static void apply_galois(Context &c, bk_ct_t a, uint32_t elt, bk_gkeys_t gk, bk_ct_t dst = nullptr)
    {
        check_ct(&c, a, "encrypted");
        if (!gk || gk->ctx != &c)
            throw std::invalid_argument("galois_keys is not valid for encryption parameters");
        bk_kskey_t key = find_gkey(gk, elt);
        if (!key)
            throw std::invalid_argument("Galois key not present");
        if (!(elt & 1) || elt >= 2 * c.n)
            throw std::invalid_argument("Galois element is not valid");
        if (a->size > 2)
            throw std::invalid_argument("encrypted size must be 2");
        if (!a->ntt)
            throw std::invalid_argument("CKKS encrypted must be in NTT form");
        cudaStream_t s = c.stream();
        const uint32_t *perm = c.galois_table(elt);
        const int l = a->limbs;
        size_t words = (size_t)2 * l * c.n;
        u64 *out = alloc_words(c, words);
        key_switch(c, s, a->d + (size_t)l * c.n, perm, a->d, nullptr, out, l, key);
        if (dst && dst != a)
        {
            if (dst->ctx != &c)
                throw std::invalid_argument("destination is not valid for encryption parameters");
            adopt(dst, out, words, 2, l);
            dst->scale = a->scale;
            dst->ntt = true;
        }
        else
            adopt(a, out, words, 2, l);
    }

    static void rotate_internal(Context &c, bk_ct_t a, int steps, bk_gkeys_t gk)
    {
        // evaluator.cpp:2224-2279
        check_ct(&c, a, "encrypted");
        if (!gk || gk->ctx != &c)
            throw std::invalid_argument("galois_keys is not valid for encryption parameters");
        if (steps == 0)
            return;
        uint32_t elt = galois_elt_from_step(c.log_n, steps);
        if (find_gkey(gk, elt))
        {
            apply_galois(c, a, elt, gk);
            return;
        }
        // NAF fallback (util/numth.h:22-42)
        std::vector<int> naf;
        {
            bool sign = steps < 0;
            int v = std::abs(steps);
            for (int i = 0; v; i++)
            {
                int zi = (v & 1) ? 2 - (v & 3) : 0;
                v = (v - zi) >> 1;
                if (zi)
                    naf.push_back((sign ? -zi : zi) * (1 << i));
            }
        }
        if (naf.size() == 1)
            throw std::invalid_argument("Galois key not present");
        for (int st : naf)
            if ((size_t)std::abs(st) != (c.n >> 1))
                rotate_internal(c, a, st, gk);
    }

    // -------------------------------------------------------------------------- scalar encoding
    // CKKSEncoder::encode_internal(double, ...) (ckks.cpp:77-216): residues of round(value*scale)
    // for the first `limbs` primes.  The reference encodes at the top level and drops limbs;
    // the range checks therefore use the top-level bit count.
    static void scalar_residues(const Context &c, double value, double scale, int limbs, ulonglong2 *out)
    {
        int top = c.top_limbs();
        if (scale <= 0 || ((int)std::log2(scale) >= c.total_bits[top]))
            throw std::invalid_argument("scale out of bounds");
        value *= scale;
        int coeff_bit_count = value == 0 ? 0 : (int)std::log2(std::fabs(value)) + 2;
        if (coeff_bit_count >= c.total_bits[top])
            throw std::invalid_argument("encoded value is too large");
        double coeffd = std::round(value);
        bool neg = std::signbit(coeffd);
        coeffd = std::fabs(coeffd);
        // coeffd is an exact integer m * 2^e; reduce it exactly
        int e = 0;
        uint64_t m = 0;
        if (coeffd != 0)
        {
            double fr = std::frexp(coeffd, &e); // coeffd = fr * 2^e, fr in [0.5,1)
            m = (uint64_t)std::ldexp(fr, 53);
            e -= 53;
        }
        for (int j = 0; j < limbs; j++)
        {
            uint64_t q = c.primes[j];
            uint64_t r;
            if (e >= 0)
                r = mulmod(m % q, powmod(2, (uint64_t)e, q), q);
            else
                r = (e <= -64 ? 0 : (m >> (-e))) % q;
            if (neg && r)
                r = q - r;
            out[j] = make_ulonglong2(r, shoup(r, q));
        }
    }
} // namespace bk

namespace bk
{
    // divide_and_round_q_last_ntt_inplace (util/rns.cpp:737-808) on every polynomial of `a`;
    // the last limb's prime index is a->limbs - 1 (at the key level that is the special prime).
    void rescale_core(Context &c, bk_ct_t a)
    {
        cudaStream_t s = c.stream();
        const int l = a->limbs, lo = l - 1, k = a->size;
        const size_t n = c.n;
        Scratch tmp(s, (size_t)k * std::max(lo, 1) * n);
        Scratch tlast(s, (size_t)k * n);
        LdInvLimbOf ld{ a->d, n, l, lo, lo };
        launch_inv_blocks(c, s, ld, tmp.p, k);
        StInvAddHalf st{ tlast.p, n, lo };
        launch_inv_cols(c, s, tmp.p, st, k);
        LdDivRound ld2{ tlast.p, n, lo, c.primes[lo] };
        launch_fwd_cols(c, s, ld2, tmp.p, k * lo);
        size_t words = (size_t)k * lo * n;
        u64 *out = alloc_words(c, words);
        StRescale st2{ a->d, out, c.inv_last(lo), n, l, lo };
        launch_fwd_blocks(c, s, tmp.p, st2, k * lo);
        adopt(a, out, words, k, lo);
        a->scale = a->scale / (double)c.primes[lo];
    }

} // namespace bk

using namespace bk;

struct ScalarPack
{
    ulonglong2 c[62];
};

template <bool MUL>
__global__ void __launch_bounds__(256) k_scalar_pack(u64 *__restrict__ a, ScalarPack sp, const PrimeDev *primes,
                                                     int log_n, int limbs, int polys)
{
    pdl_prologue();
    const size_t n = size_t(1) << log_n;
    const size_t per_poly = (size_t)limbs * n;
    const size_t total = (size_t)polys * per_poly / 2;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
    {
        size_t e = i * 2;
        int limb = (int)((e % per_poly) >> log_n);
        const u64 q = primes[limb].q;
        ulonglong2 f = sp.c[limb];
        ulonglong2 va = *reinterpret_cast<ulonglong2 *>(a + e);
        if (MUL)
        {
            va.x = csub(mul_shoup_lazy(va.x, f.x, f.y, q), q);
            va.y = csub(mul_shoup_lazy(va.y, f.x, f.y, q), q);
        }
        else
        {
            va.x = addmod(va.x, f.x, q);
            va.y = addmod(va.y, f.x, q);
        }
        *reinterpret_cast<ulonglong2 *>(a + e) = va;
    }
}

// dst = sum_j scalar_j * ct_j (+ a constant on c0): the leaves of a Chebyshev evaluation tree in one pass, see
// bk_scalar_linear_combination.  Sources may carry more limbs than dst (only the first `limbs` are read).
constexpr int LINCOMB_TERMS = 8;
struct LinCombArgs
{
    const u64 *src[LINCOMB_TERMS];
    int limbs_in[LINCOMB_TERMS];
    int size_in[LINCOMB_TERMS]; // polynomials of each source (2, or 3 for an unrelinearized product)
    int terms;
    ulonglong2 c[LINCOMB_TERMS][62];
};
__global__ void __launch_bounds__(256) k_scalar_lincomb(u64 *__restrict__ dst, const __grid_constant__ LinCombArgs a,
                                                        const __grid_constant__ ScalarPack addc, const PrimeDev *primes,
                                                        int log_n, int limbs, int polys)
{
    pdl_prologue();
    const size_t n = size_t(1) << log_n;
    const size_t per_poly = (size_t)limbs * n;
    const size_t total = per_poly * polys / 2; // `polys` polynomials, two words per step
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
    {
        const size_t e = i * 2;
        const int p = (int)(e / per_poly);
        const size_t r = e - (size_t)p * per_poly;
        const int limb = (int)(r >> log_n);
        const size_t x = r & (n - 1);
        const PrimeDev pd = primes[limb];
        u64 lo0 = 0, hi0 = 0, lo1 = 0, hi1 = 0;
#pragma unroll
        for (int j = 0; j < LINCOMB_TERMS; j++)
        {
            if (j < a.terms && p < a.size_in[j])
            {
                const ulonglong2 v = __ldcs(reinterpret_cast<const ulonglong2 *>(a.src[j] + ((size_t)p * a.limbs_in[j] + limb) * n + x));
                const u64 f = a.c[j][limb].x;
                mac128(lo0, hi0, v.x, f);
                mac128(lo1, hi1, v.y, f);
            }
        }
        ulonglong2 o;
        o.x = barrett128(lo0, hi0, pd);
        o.y = barrett128(lo1, hi1, pd);
        if (p == 0)
        {
            o.x = addmod(o.x, addc.c[limb].x, pd.q);
            o.y = addmod(o.y, addc.c[limb].x, pd.q);
        }
        *reinterpret_cast<ulonglong2 *>(dst + e) = o;
    }
}

extern "C"
{
    const char *bk_last_error(void)
    {
        return g_err.c_str();
    }
    const char *bk_version(void)
    {
        return "b200ckks 0.1 (sm_100a)";
    }

    // ---- host-only helpers ---------------------------------------------------------------------
    bk_status bk_coeff_modulus_create(int log_n, const int *bit_sizes, int count, uint64_t *primes_out)
    {
        BK_TRY
        if (!bit_sizes || !primes_out || count < 1)
            throw std::invalid_argument("bit_sizes is invalid");
        auto p = coeff_modulus_create(log_n, std::vector<int>(bit_sizes, bit_sizes + count));
        std::memcpy(primes_out, p.data(), sizeof(uint64_t) * p.size());
        BK_END
    }
    bk_status bk_minimal_primitive_root(int log_n, uint64_t q, uint64_t *root_out)
    {
        BK_TRY
        *root_out = minimal_primitive_root(uint64_t(2) << log_n, q);
        BK_END
    }
    bk_status bk_galois_elt_from_step(int log_n, int step, uint32_t *elt_out)
    {
        BK_TRY
        *elt_out = galois_elt_from_step(log_n, step);
        BK_END
    }
    bk_status bk_galois_table_ntt(int log_n, uint32_t galois_elt, uint32_t *table_out)
    {
        BK_TRY
        galois_table_ntt(log_n, galois_elt, table_out);
        BK_END
    }
    bk_status bk_ntt_root_powers(int log_n, uint64_t q, int inverse, uint64_t *out)
    {
        BK_TRY
        if (inverse)
        {
            auto v = seal_inv_root_powers(log_n, q);
            std::memcpy(out, v.data(), v.size() * sizeof(uint64_t));
        }
        else
        {
            std::vector<uint64_t> w, iw;
            ntt_tables(log_n, q, w, iw);
            std::memcpy(out, w.data(), w.size() * sizeof(uint64_t));
        }
        BK_END
    }

    // ---- context -------------------------------------------------------------------------------
    bk_status bk_context_create(int log_n, const uint64_t *primes, int n_primes, int device, bk_context_t *out)
    {
        BK_TRY
        if (!primes || !out)
            throw std::invalid_argument("null argument");
        *out = new bk_context_s(log_n, primes, n_primes, device);
        BK_END
    }
    bk_status bk_context_destroy(bk_context_t ctx)
    {
        BK_TRY
        delete ctx;
        BK_END
    }
    bk_status bk_context_info(bk_context_t ctx, int *log_n, int *n_primes, int *device)
    {
        BK_TRY
        if (log_n)
            *log_n = ctx->log_n;
        if (n_primes)
            *n_primes = ctx->n_primes;
        if (device)
            *device = ctx->device;
        BK_END
    }
    bk_status bk_context_get_primes(bk_context_t ctx, uint64_t *primes_out)
    {
        BK_TRY
        std::memcpy(primes_out, ctx->primes.data(), sizeof(uint64_t) * ctx->n_primes);
        BK_END
    }
    bk_status bk_context_set_ks_chunk(bk_context_t ctx, int chunk)
    {
        BK_TRY
        if (chunk < 1)
            throw std::invalid_argument("chunk must be positive");
        ctx->ks_chunk = chunk;
        BK_END
    }
    bk_status bk_sync(bk_context_t ctx)
    {
        BK_TRY
        BK_CUDA(cudaStreamSynchronize(ctx->stream()));
        BK_END
    }
    bk_status bk_context_set_rng_key(bk_context_t ctx, const uint8_t key[32])
    {
        BK_TRY
        if (!key)
            throw std::invalid_argument("key is null");
        std::memcpy(ctx->rng_master.k, key, 32);
        BK_END
    }
    bk_status bk_context_set_hybrid(bk_context_t ctx, int on)
    {
        BK_TRY
        ctx->hybrid = on != 0;
        BK_END
    }
    bk_status bk_context_set_key_compression(bk_context_t ctx, int on)
    {
        BK_TRY
        ctx->compress_keys = on != 0;
        BK_END
    }
    bk_status bk_context_hybrid_shape(bk_context_t ctx, int limbs, int *alpha_out, int *dsize_out)
    {
        BK_TRY
        if (limbs < 1 || limbs > ctx->top_limbs())
            throw std::invalid_argument("limbs is out of range");
        int a = 1, d = 1;
        hybrid_shape(limbs, ctx->top_limbs(), a, d);
        if (alpha_out)
            *alpha_out = a;
        if (dsize_out)
            *dsize_out = d;
        BK_END
    }
    bk_status bk_context_hybrid(bk_context_t ctx, int *on, uint64_t *key_bytes, uint64_t *keys)
    {
        BK_TRY
        if (on)
            *on = ctx->hybrid ? 1 : 0;
        if (key_bytes)
            *key_bytes = ctx->hybrid_key_bytes.load();
        if (keys)
            *keys = ctx->hybrid_keys.load();
        BK_END
    }
    bk_status bk_sync_device(bk_context_t ctx)
    {
        BK_TRY
        ctx->activate();
        BK_CUDA(cudaDeviceSynchronize());
        BK_END
    }
    bk_status bk_event_record(bk_context_t ctx, bk_event_t *event_out)
    {
        BK_TRY
        cudaStream_t s = ctx->stream();
        cudaEvent_t e;
        BK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        BK_CUDA(cudaEventRecord(e, s));
        *event_out = reinterpret_cast<bk_event_t>(e);
        BK_END
    }
    bk_status bk_stream_wait_event(bk_context_t ctx, bk_event_t event)
    {
        BK_TRY
        BK_CUDA(cudaStreamWaitEvent(ctx->stream(), reinterpret_cast<cudaEvent_t>(event), 0));
        BK_END
    }
    bk_status bk_event_destroy(bk_event_t event)
    {
        BK_TRY
        if (event)
            BK_CUDA(cudaEventDestroy(reinterpret_cast<cudaEvent_t>(event)));
        BK_END
    }
    bk_status bk_stream(bk_context_t ctx, void **stream_out)
    {
        BK_TRY
        *stream_out = (void *)ctx->stream();
        BK_END
    }
    bk_status bk_launch_count(bk_context_t ctx, uint64_t *count_out)
    {
        BK_TRY
        *count_out = ctx->launches.load();
        BK_END
    }

    // ---- timing helpers (CUDA events on the calling thread's stream) ------------------------------
    bk_status bk_timer_begin(bk_context_t ctx)
    {
        BK_TRY
        cudaStream_t s = ctx->stream();
        cudaEvent_t e;
        BK_CUDA(cudaEventCreate(&e));
        BK_CUDA(cudaEventRecord(e, s));
        ctx->timer_stack.push_back(e);
        BK_END
    }
    bk_status bk_timer_end(bk_context_t ctx, double *ms_out)
    {
        BK_TRY
        if (ctx->timer_stack.empty())
            throw std::logic_error("bk_timer_end without bk_timer_begin");
        cudaStream_t s = ctx->stream();
        cudaEvent_t e0 = ctx->timer_stack.back(), e1;
        ctx->timer_stack.pop_back();
        BK_CUDA(cudaEventCreate(&e1));
        BK_CUDA(cudaEventRecord(e1, s));
        BK_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        BK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        *ms_out = ms;
        BK_END
    }
    bk_status bk_profile_begin(bk_context_t ctx, int kernel_tag)
    {
        BK_TRY
        if (kernel_tag != -2 && (kernel_tag < 0 || kernel_tag >= TAG_COUNT))
            throw std::invalid_argument("unknown kernel tag");
        BK_CUDA(cudaStreamSynchronize(ctx->stream()));
        ctx->prof_events.clear();
        ctx->prof_event_tags.clear();
        ctx->prof_tag = kernel_tag;
        BK_END
    }
    bk_status bk_profile_end(bk_context_t ctx, uint64_t *launches_out, double *total_ms_out)
    {
        BK_TRY
        ctx->prof_tag = -1;
        BK_CUDA(cudaStreamSynchronize(ctx->stream()));
        double total = 0;
        for (auto &pr : ctx->prof_events)
        {
            float ms = 0;
            BK_CUDA(cudaEventElapsedTime(&ms, pr.first, pr.second));
            total += ms;
            cudaEventDestroy(pr.first);
            cudaEventDestroy(pr.second);
        }
        *launches_out = ctx->prof_events.size();
        *total_ms_out = total;
        ctx->prof_events.clear();
        ctx->prof_event_tags.clear();
        BK_END
    }
    bk_status bk_profile_end_all(bk_context_t ctx, uint64_t launches_out[8], double total_ms_out[8])
    {
        BK_TRY
        ctx->prof_tag = -1;
        BK_CUDA(cudaStreamSynchronize(ctx->stream()));
        for (int t = 0; t < TAG_COUNT; t++)
        {
            launches_out[t] = 0;
            total_ms_out[t] = 0;
        }
        for (size_t i = 0; i < ctx->prof_events.size(); i++)
        {
            float ms = 0;
            auto &pr = ctx->prof_events[i];
            BK_CUDA(cudaEventElapsedTime(&ms, pr.first, pr.second));
            int t = ctx->prof_event_tags[i];
            launches_out[t]++;
            total_ms_out[t] += ms;
            cudaEventDestroy(pr.first);
            cudaEventDestroy(pr.second);
        }
        ctx->prof_events.clear();
        ctx->prof_event_tags.clear();
        BK_END
    }
    bk_status bk_kernel_counters(bk_context_t ctx, uint64_t launches_out[8], uint64_t units_out[8])
    {
        BK_TRY
        for (int t = 0; t < TAG_COUNT; t++)
        {
            launches_out[t] = ctx->tag_launches[t].load();
            units_out[t] = ctx->tag_units[t].load();
        }
        BK_END
    }
    bk_status bk_transfer_bytes(bk_context_t ctx, uint64_t *h2d_out, uint64_t *d2h_out)
    {
        BK_TRY
        *h2d_out = ctx->h2d_bytes.load();
        *d2h_out = ctx->d2h_bytes.load();
        BK_END
    }
    // write `bytes` bytes of junk (> L2) so the next timed region starts with a cold L2
    bk_status bk_flush_l2(bk_context_t ctx)
    {
        BK_TRY
        static const size_t words = (size_t)192 << 17; // 192 MiB
        Scratch junk(ctx->stream(), words);
        BK_CUDA(cudaMemsetAsync(junk.p, 0x5a, words * sizeof(u64), ctx->stream()));
        BK_END
    }

    // Peak of the pipe that bounds the NTT passes, measured on this GPU at its current clocks: 32-bit integer
    // multiply-add (IMAD) thread-instructions per second from a kernel of independent dependent-chains - the
    // denominator of bench.py's integer roofline (a 64-bit Shoup butterfly needs 9 of them, ntt.cuh ct_bfly_wide).
    static double imad_rate(Context &c, int kind)
    {
        cudaStream_t s = c.stream();
        Scratch sink(s, 1);
        const int iters = 4096, chains = 8;
        dim3 grid((unsigned)c.sm_count * 8);
        cudaEvent_t e0, e1;
        BK_CUDA(cudaEventCreate(&e0));
        BK_CUDA(cudaEventCreate(&e1));
        double best = 0;
        for (int rep = 0; rep < 4; rep++)
        {
            BK_CUDA(cudaEventRecord(e0, s));
            if (kind == 0)
                launch_pdl(k_imad_peak<0>, grid, 256, 0, s, (unsigned long long *)sink.p, iters, 0x9E3779B9u);
            else if (kind == 1)
                launch_pdl(k_imad_peak<1>, grid, 256, 0, s, (unsigned long long *)sink.p, iters, 0x9E3779B9u);
            else
                launch_pdl(k_imad_peak<2>, grid, 256, 0, s, (unsigned long long *)sink.p, iters, 0x9E3779B9u);
            BK_CUDA(cudaEventRecord(e1, s));
            BK_CUDA(cudaEventSynchronize(e1));
            float ms = 0;
            BK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            double rate = (double)grid.x * 256.0 * iters * chains / (ms * 1e-3);
            if (rep > 0 && rate > best)
                best = rate;
        }
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        c.count(4);
        return best;
    }
    bk_status bk_measure_imad_peak(bk_context_t ctx, double *imad_per_second_out)
    {
        BK_TRY
        *imad_per_second_out = imad_rate(*ctx, 0);
        BK_END
    }
    bk_status bk_measure_int_pipe(bk_context_t ctx, double rates_out[3])
    {
        BK_TRY
        for (int k = 0; k < 3; k++)
            rates_out[k] = imad_rate(*ctx, k);
        BK_END
    }

    // ---- ciphertext ----------------------------------------------------------------------------
    bk_status bk_ct_create(bk_context_t ctx, bk_ct_t *out)
    {
        BK_TRY
        auto ct = new bk_ct_s();
        ct->ctx = ctx;
        *out = ct;
        BK_END
    }
    bk_status bk_ct_destroy(bk_ct_t ct)
    {
        BK_TRY
        if (ct)
        {
            ct->ctx->release_words(ct->d, ct->owner);
            delete ct;
        }
        BK_END
    }
    bk_status bk_ct_copy(bk_ct_t dst, bk_ct_t src)
    {
        BK_TRY
        if (dst == src)
            return BK_OK;
        if (dst->ctx != src->ctx)
            throw std::invalid_argument("context mismatch");
        Context &c = *src->ctx;
        ensure_ct(dst, src->size, src->limbs, false);
        size_t words = (size_t)src->size * src->limbs * c.n;
        if (words)
            BK_CUDA(cudaMemcpyAsync(dst->d, src->d, words * sizeof(u64), cudaMemcpyDeviceToDevice, c.stream()));
        dst->scale = src->scale;
        dst->ntt = src->ntt;
        BK_END
    }
    bk_status bk_ct_resize(bk_ct_t ct, int size, int limbs)
    {
        BK_TRY
        Context &c = *ct->ctx;
        if (size < 2 || limbs < 1 || limbs > c.n_primes)
            throw std::invalid_argument("invalid size");
        // Ciphertext::resize keeps the leading words of the flat buffer (dynarray.h) - same here
        int old_size = ct->size, old_limbs = ct->limbs;
        size_t old_words = (size_t)old_size * old_limbs * c.n;
        size_t words = (size_t)size * limbs * c.n;
        ensure_ct(ct, size, limbs, true);
        if (words > old_words)
            BK_CUDA(cudaMemsetAsync(ct->d + old_words, 0, (words - old_words) * sizeof(u64), c.stream()));
        BK_END
    }
    bk_status bk_ct_info(bk_ct_t ct, int *size, int *limbs, double *scale, int *is_ntt)
    {
        BK_TRY
        if (size)
            *size = ct->size;
        if (limbs)
            *limbs = ct->limbs;
        if (scale)
            *scale = ct->scale;
        if (is_ntt)
            *is_ntt = ct->ntt ? 1 : 0;
        BK_END
    }
    bk_status bk_ct_set_scale(bk_ct_t ct, double scale)
    {
        BK_TRY
        ct->scale = scale;
        BK_END
    }
    bk_status bk_ct_set_ntt_form(bk_ct_t ct, int is_ntt)
    {
        BK_TRY
        ct->ntt = is_ntt != 0;
        BK_END
    }
    bk_status bk_ct_upload(bk_ct_t ct, const uint64_t *host, int size, int limbs, double scale, int is_ntt)
    {
        BK_TRY
        Context &c = *ct->ctx;
        if (size < 1 || limbs < 1 || limbs > c.n_primes)
            throw std::invalid_argument("invalid size");
        ensure_ct(ct, size, limbs, false);
        cudaStream_t s = c.stream();
        BK_CUDA(cudaMemcpyAsync(ct->d, host, (size_t)size * limbs * c.n * sizeof(u64), cudaMemcpyHostToDevice, s));
        c.h2d_bytes += (size_t)size * limbs * c.n * sizeof(u64);
        BK_CUDA(cudaStreamSynchronize(s));
        ct->scale = scale;
        ct->ntt = is_ntt != 0;
        BK_END
    }
    bk_status bk_ct_download(bk_ct_t ct, uint64_t *host_out)
    {
        BK_TRY
        Context &c = *ct->ctx;
        cudaStream_t s = c.stream();
        BK_CUDA(cudaMemcpyAsync(host_out, ct->d, (size_t)ct->size * ct->limbs * c.n * sizeof(u64),
                                cudaMemcpyDeviceToHost, s));
        c.d2h_bytes += (size_t)ct->size * ct->limbs * c.n * sizeof(u64);
        BK_CUDA(cudaStreamSynchronize(s));
        BK_END
    }
    bk_status bk_ct_device_ptr(bk_ct_t ct, void **dev_ptr_out)
    {
        BK_TRY
        *dev_ptr_out = ct->d;
        BK_END
    }

    // ---- plaintext -----------------------------------------------------------------------------
    bk_status bk_pt_create(bk_context_t ctx, bk_pt_t *out)
    {
        BK_TRY
        auto pt = new bk_pt_s();
        pt->ctx = ctx;
        *out = pt;
        BK_END
    }
    bk_status bk_pt_destroy(bk_pt_t pt)
    {
        BK_TRY
        if (pt)
        {
            pt->ctx->release_words(pt->d, pt->owner);
            delete pt;
        }
        BK_END
    }
    bk_status bk_pt_copy(bk_pt_t dst, bk_pt_t src)
    {
        BK_TRY
        if (dst == src)
            return BK_OK;
        Context &c = *src->ctx;
        ensure_pt(dst, src->limbs);
        BK_CUDA(cudaMemcpyAsync(dst->d, src->d, (size_t)src->limbs * c.n * sizeof(u64), cudaMemcpyDeviceToDevice,
                                c.stream()));
        dst->scale = src->scale;
        BK_END
    }
    bk_status bk_pt_info(bk_pt_t pt, int *limbs, double *scale)
    {
        BK_TRY
        if (limbs)
            *limbs = pt->limbs;
        if (scale)
            *scale = pt->scale;
        BK_END
    }
    bk_status bk_pt_set_scale(bk_pt_t pt, double scale)
    {
        BK_TRY
        pt->scale = scale;
        BK_END
    }
    bk_status bk_pt_upload(bk_pt_t pt, const uint64_t *host, int limbs, double scale)
    {
        BK_TRY
        Context &c = *pt->ctx;
        if (limbs < 1 || limbs > c.n_primes)
            throw std::invalid_argument("invalid size");
        ensure_pt(pt, limbs);
        cudaStream_t s = c.stream();
        BK_CUDA(cudaMemcpyAsync(pt->d, host, (size_t)limbs * c.n * sizeof(u64), cudaMemcpyHostToDevice, s));
        c.h2d_bytes += (size_t)limbs * c.n * sizeof(u64);
        BK_CUDA(cudaStreamSynchronize(s));
        pt->scale = scale;
        BK_END
    }
    bk_status bk_pt_download(bk_pt_t pt, uint64_t *host_out)
    {
        BK_TRY
        Context &c = *pt->ctx;
        cudaStream_t s = c.stream();
        BK_CUDA(cudaMemcpyAsync(host_out, pt->d, (size_t)pt->limbs * c.n * sizeof(u64), cudaMemcpyDeviceToHost, s));
        c.d2h_bytes += (size_t)pt->limbs * c.n * sizeof(u64);
        BK_CUDA(cudaStreamSynchronize(s));
        BK_END
    }
    bk_status bk_pt_mod_switch_to(bk_pt_t pt, int limbs)
    {
        BK_TRY
        // mod_switch_drop_to_next(Plaintext) (evaluator.cpp:1248-1281): limbs are contiguous, the
        // drop is a truncation
        if (limbs < 1 || limbs > pt->limbs)
            throw std::invalid_argument("cannot switch to higher level modulus");
        pt->limbs = limbs;
        BK_END
    }

    // ---- keys ----------------------------------------------------------------------------------
    bk_status bk_kskey_upload(bk_context_t ctx, const uint64_t *host, int digits, int max_limbs, bk_kskey_t *out)
    {
        BK_TRY
        Context &c = *ctx;
        const int top = c.top_limbs();
        if (digits < 1 || digits > top)
            throw std::invalid_argument("kswitch key digit count is invalid");
        int kl = (max_limbs > 0 && max_limbs < top) ? max_limbs : top;
        int kd = std::min(digits, kl);
        cudaStream_t s = c.stream();
        const size_t n = c.n;
        auto key = new bk_kskey_s();
        key->ctx = ctx;
        key->digits = kd;
        key->klimbs = kl;
        key->words = (size_t)kd * 2 * (kl + 1) * n;
        BK_CUDA(cudaMalloc((void **)&key->d, key->words * sizeof(u64)));
        // one (digit, poly) at a time: limbs 0..kl-1, then the special prime's limb
        for (int j = 0; j < kd; j++)
            for (int p = 0; p < 2; p++)
            {
                const uint64_t *src = host + ((size_t)j * 2 + p) * c.n_primes * n;
                u64 *dst = key->d + ((size_t)j * 2 + p) * (kl + 1) * n;
                BK_CUDA(cudaMemcpyAsync(dst, src, (size_t)kl * n * sizeof(u64), cudaMemcpyHostToDevice, s));
                BK_CUDA(cudaMemcpyAsync(dst + (size_t)kl * n, src + (size_t)(c.n_primes - 1) * n, n * sizeof(u64),
                                        cudaMemcpyHostToDevice, s));
            }
        BK_CUDA(cudaStreamSynchronize(s));
        *out = key;
        BK_END
    }
    bk_status bk_kskey_destroy(bk_kskey_t key)
    {
        BK_TRY
        if (key)
        {
            key->ctx->activate();
            cudaStreamSynchronize(key->ctx->stream());
            cudaFree(key->d);
            for (auto &kv : key->hyb)
            {
                key->ctx->hybrid_key_bytes.fetch_sub(kv.second->words * sizeof(u64));
                cudaFree(kv.second->d);
                delete kv.second;
            }
            delete key;
        }
        BK_END
    }
    bk_status bk_kskey_info(bk_kskey_t key, int *digits, int *limbs, uint64_t *device_bytes)
    {
        BK_TRY
        if (digits)
            *digits = key->digits;
        if (limbs)
            *limbs = key->klimbs;
        if (device_bytes)
            *device_bytes = key->words * sizeof(u64);
        BK_END
    }
    // ---- key plans: evaluation keys generated up front, evaluation without the secret key ------------------------
    bk_status bk_kskey_levels(bk_kskey_t key, int *levels_out, int cap, int *count_out)
    {
        BK_TRY
        std::lock_guard<std::mutex> g(key->hmu);
        int n = 0;
        if (key->d)
        {
            if (n < cap && levels_out)
                levels_out[n] = key->klimbs;
            n++;
        }
        for (auto &kv : key->hyb)
        {
            if (n < cap && levels_out)
                levels_out[n] = kv.first;
            n++;
        }
        if (count_out)
            *count_out = n;
        BK_END
    }
    bk_status bk_kskey_prepare_level(bk_kskey_t key, int limbs)
    {
        BK_TRY
        Context &c = *key->ctx;
        if (limbs < 1 || limbs > c.top_limbs())
            throw std::invalid_argument("limbs is out of range");
        if (!key->recipe)
            throw std::logic_error("only recipes of the level-aware hybrid mode have level-specific keys");
        hybrid_key(c, key, limbs);
        BK_END
    }
    bk_status bk_kskey_drop_secret(bk_kskey_t key)
    {
        BK_TRY
        std::lock_guard<std::mutex> g(key->hmu);
        key->sk = nullptr;
        BK_END
    }
    bk_status bk_kskey_download(bk_kskey_t key, uint64_t *host_out)
    {
        // SEAL layout restricted to the resident part: [digits][2][klimbs+1][N], special last
        BK_TRY
        Context &c = *key->ctx;
        if (!key->d)
            throw std::logic_error("this key is a hybrid-mode recipe: it has no SEAL-layout data to download");
        cudaStream_t s = c.stream();
        BK_CUDA(cudaMemcpyAsync(host_out, key->d, key->words * sizeof(u64), cudaMemcpyDeviceToHost, s));
        BK_CUDA(cudaStreamSynchronize(s));
        BK_END
    }
    bk_status bk_kskey_export_device(bk_kskey_t key, void *dev_dst)
    {
        BK_TRY
        Context &c = *key->ctx;
        if (!key->d)
            throw std::logic_error("this key is a hybrid-mode recipe: it has no SEAL-layout data to export");
        BK_CUDA(cudaMemcpyAsync(dev_dst, key->d, key->words * sizeof(u64), cudaMemcpyDeviceToDevice, c.stream()));
        BK_CUDA(cudaStreamSynchronize(c.stream()));
        BK_END
    }
    bk_status bk_kskey_import_device(bk_context_t ctx, const void *dev_src, int digits, int limbs, bk_kskey_t *out)
    {
        BK_TRY
        Context &c = *ctx;
        if (digits < 1 || limbs < digits || limbs > c.top_limbs())
            throw std::invalid_argument("kswitch key shape is invalid");
        auto key = new bk_kskey_s();
        key->ctx = ctx;
        key->digits = digits;
        key->klimbs = limbs;
        key->words = (size_t)digits * 2 * (limbs + 1) * c.n;
        c.activate();
        BK_CUDA(cudaMalloc((void **)&key->d, key->words * sizeof(u64)));
        BK_CUDA(cudaMemcpyAsync(key->d, dev_src, key->words * sizeof(u64), cudaMemcpyDeviceToDevice, c.stream()));
        BK_CUDA(cudaStreamSynchronize(c.stream()));
        *out = key;
        BK_END
    }
    bk_status bk_gkeys_create(bk_context_t ctx, bk_gkeys_t *out)
    {
        BK_TRY
        auto g = new bk_gkeys_s();
        g->ctx = ctx;
        *out = g;
        BK_END
    }
    bk_status bk_gkeys_destroy(bk_gkeys_t gk)
    {
        BK_TRY
        if (gk)
        {
            for (auto &kv : gk->keys)
                bk_kskey_destroy(kv.second);
            delete gk;
        }
        BK_END
    }
    bk_status bk_gkeys_set(bk_gkeys_t gk, uint32_t galois_elt, bk_kskey_t key)
    {
        BK_TRY
        std::lock_guard<std::mutex> g(gk->mu);
        auto it = gk->keys.find(galois_elt);
        if (it != gk->keys.end())
        {
            bk_kskey_destroy(it->second);
            it->second = key;
        }
        else
            gk->keys[galois_elt] = key;
        BK_END
    }
    bk_status bk_gkeys_get(bk_gkeys_t gk, uint32_t galois_elt, bk_kskey_t *key_out)
    {
        BK_TRY
        *key_out = find_gkey(gk, galois_elt);
        if (!*key_out)
            throw std::invalid_argument("Galois key not present");
        BK_END
    }
    bk_status bk_gkeys_has(bk_gkeys_t gk, uint32_t galois_elt, int *has_out)
    {
        BK_TRY
        *has_out = find_gkey(gk, galois_elt) ? 1 : 0;
        BK_END
    }

    // ---- Evaluator -----------------------------------------------------------------------------
    static void add_sub(bk_context_t ctx, bk_ct_t a, bk_ct_t b, bool sub)
    {
        Context &c = *ctx;
        check_ct(ctx, a, "encrypted1");
        check_ct(ctx, b, "encrypted2");
        if (a->limbs != b->limbs)
            throw std::invalid_argument("encrypted1 and encrypted2 parameter mismatch");
        if (a->ntt != b->ntt)
            throw std::invalid_argument("NTT form mismatch");
        if (!close_scale(a->scale, b->scale))
            throw std::invalid_argument("scale mismatch");
        cudaStream_t s = c.stream();
        const int l = a->limbs;
        int mn = std::min(a->size, b->size), mx = std::max(a->size, b->size);
        int a_size = a->size;
        if (mx > a->size)
            ensure_ct(a, mx, l, true);
        size_t per_poly = (size_t)l * c.n;
        int grid = c.ew_grid((size_t)mn * per_poly / 2);
        {
            ProfScope ps_ew(c, s, TAG_ELEMENTWISE, mn * l);
            if (sub)
                launch_pdl(k_ew<EW_SUB>, grid, 256, 0, s, a->d, b->d, c.d_primes, c.log_n, l, mn, mn);
            else
                launch_pdl(k_ew<EW_ADD>, grid, 256, 0, s, a->d, b->d, c.d_primes, c.log_n, l, mn, mn);
        }
        c.count();
        if (a_size < b->size)
        {
            size_t words = (size_t)(b->size - a_size) * per_poly;
            BK_CUDA(cudaMemcpyAsync(a->d + (size_t)a_size * per_poly, b->d + (size_t)a_size * per_poly,
                                    words * sizeof(u64), cudaMemcpyDeviceToDevice, s));
            if (sub)
            {
                ProfScope ps_ew(c, c.stream(), TAG_ELEMENTWISE);
                launch_pdl(k_ew<EW_NEG>, c.ew_grid(words / 2), 256, 0, s, a->d + (size_t)a_size * per_poly, nullptr,
                                                                  c.d_primes, c.log_n, l, b->size - a_size, 0);
                c.count();
            }
        }
    }

    bk_status bk_add_inplace(bk_context_t ctx, bk_ct_t a, bk_ct_t b)
    {
        BK_TRY
        add_sub(ctx, a, b, false);
        BK_END
    }
    bk_status bk_sub_inplace(bk_context_t ctx, bk_ct_t a, bk_ct_t b)
    {
        BK_TRY
        add_sub(ctx, a, b, true);
        BK_END
    }
    bk_status bk_negate_inplace(bk_context_t ctx, bk_ct_t a)
    {
        BK_TRY
        Context &c = *ctx;
        check_ct(ctx, a, "encrypted");
        size_t words = (size_t)a->size * a->limbs * c.n;
        ProfScope ps_ew(c, c.stream(), TAG_ELEMENTWISE);
        launch_pdl(k_ew<EW_NEG>, c.ew_grid(words / 2), 256, 0, c.stream(), a->d, nullptr, c.d_primes, c.log_n, a->limbs,
                                                                   a->size, 0);
        c.count();
        BK_END
    }

    bk_status bk_multiply_inplace(bk_context_t ctx, bk_ct_t a, bk_ct_t b)
    {
        BK_TRY
        // ckks_multiply (evaluator.cpp:673-814), size 2 x size 2 -> size 3
        Context &c = *ctx;
        check_ct(ctx, a, "encrypted1");
        check_ct(ctx, b, "encrypted2");
        if (a->limbs != b->limbs)
            throw std::invalid_argument("encrypted1 and encrypted2 parameter mismatch");
        if (!a->ntt || !b->ntt)
            throw std::invalid_argument("encrypted1 or encrypted2 must be in NTT form");
        if (a->size != 2 || b->size != 2)
            throw std::invalid_argument("only size-2 ciphertexts are supported by multiply");
        double new_scale = a->scale * b->scale;
        if (!c.scale_in_bounds(new_scale, a->limbs))
            throw std::invalid_argument("scale out of bounds");
        const int l = a->limbs;
        size_t words = (size_t)3 * l * c.n;
        u64 *out = alloc_words(c, words);
        {
            ProfScope ps_ew(c, c.stream(), TAG_ELEMENTWISE, 3 * l);
            if (a == b)
                launch_pdl(k_square, c.ew_grid((size_t)l * c.n / 2), 256, 0, c.stream(), a->d, out, c.d_primes, c.log_n, l);
            else
                launch_pdl(k_tensor, c.ew_grid((size_t)l * c.n / 2), 256, 0, c.stream(), a->d, b->d, out, c.d_primes, c.log_n, l);
        }
        c.count();
        adopt(a, out, words, 3, l);
        a->scale = new_scale;
        BK_END
    }
    bk_status bk_square_inplace(bk_context_t ctx, bk_ct_t a)
    {
        BK_TRY
        // ckks_square (evaluator.cpp:1000-1059)
        Context &c = *ctx;
        check_ct(ctx, a, "encrypted");
        if (!a->ntt)
            throw std::invalid_argument("encrypted must be in NTT form");
        if (a->size != 2)
            throw std::invalid_argument("only size-2 ciphertexts are supported by square");
        double new_scale = a->scale * a->scale;
        if (!c.scale_in_bounds(new_scale, a->limbs))
            throw std::invalid_argument("scale out of bounds");
        const int l = a->limbs;
        size_t words = (size_t)3 * l * c.n;
        u64 *out = alloc_words(c, words);
        ProfScope ps_ew(c, c.stream(), TAG_ELEMENTWISE);
        launch_pdl(k_square, c.ew_grid((size_t)l * c.n / 2), 256, 0, c.stream(), a->d, out, c.d_primes, c.log_n, l);
        c.count();
        adopt(a, out, words, 3, l);
        a->scale = new_scale;
        BK_END
    }

    bk_status bk_relinearize_inplace(bk_context_t ctx, bk_ct_t a, bk_kskey_t relin_key)
    {
        BK_TRY
        // relinearize_internal (evaluator.cpp:1061-1116) for size 3 -> 2
        Context &c = *ctx;
        check_ct(ctx, a, "encrypted");
        if (!relin_key || relin_key->ctx != ctx)
            throw std::invalid_argument("relin_keys is not valid for encryption parameters");
        if (a->size == 2)
            return BK_OK;
        if (a->size != 3)
            throw std::invalid_argument("not enough relinearization keys");
        if (!a->ntt)
            throw std::invalid_argument("CKKS encrypted must be in NTT form");
        const int l = a->limbs;
        size_t per_poly = (size_t)l * c.n;
        size_t words = 2 * per_poly;
        u64 *out = alloc_words(c, words);
        key_switch(c, c.stream(), a->d + 2 * per_poly, nullptr, a->d, a->d + per_poly, out, l, relin_key);
        adopt(a, out, words, 2, l);
        BK_END
    }

    bk_status bk_relinearize_rescale_inplace(bk_context_t ctx, bk_ct_t a, bk_kskey_t relin_key)
    {
        BK_TRY
        // relinearize_inplace followed by rescale_to_next_inplace (evaluator.cpp:1061-1116, 1378-1414).  With hybrid
        // key switching at a level that has idle primes the two divisions (ModDown by P_S, rescale by q_{l-1}) are
        // one (hyb_mod_down_rescale): same value up to one rounding instead of two, limbs
        // differ - tolerance mode, like hybrid key switching itself.  Everywhere else: the two calls in sequence.
        Context &c = *ctx;
        check_ct(ctx, a, "encrypted");
        if (!relin_key || relin_key->ctx != ctx)
            throw std::invalid_argument("relin_keys is not valid for encryption parameters");
        if (a->size != 2 && a->size != 3)
            throw std::invalid_argument("not enough relinearization keys");
        if (!a->ntt)
            throw std::invalid_argument("CKKS encrypted must be in NTT form");
        if (a->limbs < 2)
            throw std::invalid_argument("end of modulus switching chain reached");
        const int l = a->limbs;
        static const bool merged = [] {
            const char *e = std::getenv("B200CKKS_MERGED_RESCALE");
            return !e || std::atoi(e) != 0;
        }();
        if (merged && a->size == 3 && relin_key->recipe && hybrid_plan(c, l).alpha > 1 && hybrid_plan(c, l).rescale_tables)
        {
            const size_t per_poly = (size_t)l * c.n;
            const size_t words = (size_t)2 * (l - 1) * c.n;
            u64 *out = alloc_words(c, words);
            key_switch_hybrid(c, c.stream(), a->d + 2 * per_poly, nullptr, a->d, a->d + per_poly, out, l, relin_key, true);
            adopt(a, out, words, 2, l - 1);
            a->scale = a->scale / (double)c.primes[(size_t)(l - 1)];
            return BK_OK;
        }
        if (a->size == 3)
        {
            const size_t per_poly = (size_t)l * c.n;
            const size_t words = 2 * per_poly;
            u64 *out = alloc_words(c, words);
            key_switch(c, c.stream(), a->d + 2 * per_poly, nullptr, a->d, a->d + per_poly, out, l, relin_key);
            adopt(a, out, words, 2, l);
        }
        rescale_core(c, a);
        BK_END
    }

    bk_status bk_rescale_to_next_inplace(bk_context_t ctx, bk_ct_t a)
    {
        BK_TRY
        // rescale_to_next (evaluator.cpp:1378-1414) -> mod_switch_scale_to_next (:1118-1181)
        check_ct(ctx, a, "encrypted");
        if (a->limbs < 2)
            throw std::invalid_argument("end of modulus switching chain reached");
        if (!a->ntt)
            throw std::invalid_argument("CKKS encrypted must be in NTT form");
        rescale_core(*ctx, a);
        BK_END
    }

    static void drop_to(Context &c, bk_ct_t a, int limbs)
    {
        if (limbs == a->limbs)
            return;
        size_t words = (size_t)a->size * limbs * c.n;
        u64 *out = alloc_words(c, words);
        ProfScope ps_ew(c, c.stream(), TAG_ELEMENTWISE, a->size * limbs);
        launch_pdl(k_drop_limbs, c.ew_grid(words / 2), 256, 0, c.stream(), a->d, out, c.log_n, a->limbs, limbs, a->size);
        c.count();
        adopt(a, out, words, a->size, limbs);
    }

    bk_status bk_mod_switch_to_next_inplace(bk_context_t ctx, bk_ct_t a)
    {
        BK_TRY
        // mod_switch_drop_to_next (evaluator.cpp:1183-1246)
        check_ct(ctx, a, "encrypted");
        if (a->limbs < 2)
            throw std::invalid_argument("end of modulus switching chain reached");
        if (!a->ntt)
            throw std::invalid_argument("CKKS encrypted must be in NTT form");
        if (!ctx->scale_in_bounds(a->scale, a->limbs - 1))
            throw std::invalid_argument("scale out of bounds");
        drop_to(*ctx, a, a->limbs - 1);
        BK_END
    }
    bk_status bk_mod_switch_to_inplace(bk_context_t ctx, bk_ct_t a, int limbs)
    {
        BK_TRY
        // mod_switch_to_inplace (evaluator.cpp:1326-1348): repeated drop == one strided copy
        check_ct(ctx, a, "encrypted");
        if (limbs < 1 || limbs > ctx->top_limbs())
            throw std::invalid_argument("parms_id is not valid for encryption parameters");
        if (limbs > a->limbs)
            throw std::invalid_argument("cannot switch to higher level modulus");
        if (limbs < a->limbs)
        {
            if (!a->ntt)
                throw std::invalid_argument("CKKS encrypted must be in NTT form");
            if (!ctx->scale_in_bounds(a->scale, limbs))
                throw std::invalid_argument("scale out of bounds");
        }
        drop_to(*ctx, a, limbs);
        BK_END
    }

    bk_status bk_apply_galois_inplace(bk_context_t ctx, bk_ct_t a, uint32_t galois_elt, bk_gkeys_t gk)
    {
        BK_TRY
        apply_galois(*ctx, a, galois_elt, gk);
        BK_END
    }
    bk_status bk_apply_galois(bk_context_t ctx, bk_ct_t a, uint32_t galois_elt, bk_gkeys_t gk, bk_ct_t dst)
    {
        BK_TRY
        if (!dst)
            throw std::invalid_argument("destination is null");
        apply_galois(*ctx, a, galois_elt, gk, dst);
        BK_END
    }
    bk_status bk_rotate_vector_inplace(bk_context_t ctx, bk_ct_t a, int steps, bk_gkeys_t gk)
    {
        BK_TRY
        rotate_internal(*ctx, a, steps, gk);
        BK_END
    }
    bk_status bk_apply_galois_hoisted(bk_context_t ctx, bk_ct_t in, const uint32_t *galois_elts, int count, bk_gkeys_t gk,
                                      bk_ct_t *outs)
    {
        BK_TRY
        Context &c = *ctx;
        check_ct(ctx, in, "encrypted");
        if (!gk || gk->ctx != ctx)
            throw std::invalid_argument("galois_keys is not valid for encryption parameters");
        if (in->size > 2)
            throw std::invalid_argument("encrypted size must be 2");
        if (!in->ntt)
            throw std::invalid_argument("CKKS encrypted must be in NTT form");
        if (count < 1 || count > 256)
            throw std::invalid_argument("count is out of range");
        std::vector<const uint32_t *> perms(count);
        std::vector<bk_kskey_t> keys(count);
        std::vector<u64 *> bufs(count);
        const int l = in->limbs;
        const size_t words = (size_t)2 * l * c.n;
        for (int k = 0; k < count; k++)
        {
            if (!(galois_elts[k] & 1) || galois_elts[k] >= 2 * c.n)
                throw std::invalid_argument("Galois element is not valid");
            keys[k] = find_gkey(gk, galois_elts[k]);
            if (!keys[k])
                throw std::invalid_argument("Galois key not present");
            if (!outs[k] || outs[k]->ctx != ctx || outs[k] == in)
                throw std::invalid_argument("destination is not valid");
            perms[k] = c.galois_table(galois_elts[k]);
        }
        for (int k = 0; k < count; k++)
            bufs[k] = alloc_words(c, words);
        key_switch_hoisted(c, c.stream(), in, count, perms.data(), keys.data(), bufs.data());
        for (int k = 0; k < count; k++)
        {
            adopt(outs[k], bufs[k], words, 2, l);
            outs[k]->scale = in->scale;
            outs[k]->ntt = true;
        }
        BK_END
    }
    // Double-hoisted inner sums of a baby-step / giant-step linear transform (Bossuat, Mouchet, Troncoso-Pastoriza,
    // Hubaux: "Efficient bootstrapping for approximate homomorphic encryption with non-sparse keys", Alg. 6), for the
    // level-aware hybrid key switching of this engine.  The reference computes, per giant step g,
    //     sum_k  rotate(ct, baby_k) (.) pt[g][k]                      (Bootstrapper.cpp:1969-2012)
    // with one full key switch (incl. its division by the special modulus) per baby rotation.  Here the input is
    // decomposed once, every baby rotation stops after the inner product with its key - an accumulator over the
    // EXTENDED basis Q_l * P_S - the plaintexts are multiplied in that basis, and the division by P_S (ModDown) is done
    // once per giant step on the sum.  The c0 halves need no key switch at all: rotate(ct)_0 = perm(c0) + ModDown(acc_0),
    // and ModDown(P_S x) = x, so sum_k perm_k(c0) pt[g][k] is added after the ModDown.  Decrypted values equal the
    // reference's up to key-switching noise (one rounding per giant step instead of one per baby step).
    //   elts[k]: Galois element of baby step k, 1 = no rotation.  pts[g * n_baby + k]: extended plaintext
    //   (bk_encode_ext at the ciphertext's level) or NULL where the group has no such term.  outs[g]: the giant step's
    //   ciphertext at the input's level, scale = ct scale * plaintext scale.
    bk_status bk_bsgs_inner_sums(bk_context_t ctx, bk_ct_t in, const uint32_t *elts, int n_baby, bk_gkeys_t gk,
                                 const bk_pt_t *pts, int n_giant, bk_ct_t *outs, int rescale)
    {
        BK_TRY
        Context &c = *ctx;
        check_ct(ctx, in, "encrypted");
        if (!gk || gk->ctx != ctx)
            throw std::invalid_argument("galois_keys is not valid for encryption parameters");
        if (in->size != 2 || !in->ntt)
            throw std::invalid_argument("encrypted must be a size-2 ciphertext in NTT form");
        if (n_baby < 1 || n_baby > 64 || n_giant < 1 || n_giant > 64)
            throw std::invalid_argument("count is out of range");
        if (!c.hybrid)
            throw std::logic_error("double-hoisted inner sums need the level-aware hybrid key switching mode");
        const int l = in->limbs;
        const size_t n = c.n;
        const HybridPlan &P = hybrid_plan(c, l);
        const HybDims h{ l, P.alpha, P.dsize, P.dnum, c.n_primes - 1 };
        const int ne = P.ne;
        cudaStream_t s = c.stream();

        // rotating babies: keys, tables
        std::vector<int> rot_of((size_t)n_baby, -1);
        std::vector<const uint32_t *> perms;
        std::vector<bk_hybkey_s *> hks;
        for (int k = 0; k < n_baby; k++)
        {
            if (elts[k] == 1)
                continue;
            if (!(elts[k] & 1) || elts[k] >= 2 * n)
                throw std::invalid_argument("Galois element is not valid");
            bk_kskey_t key = find_gkey(gk, elts[k]);
            if (!key)
                throw std::invalid_argument("Galois key not present");
            if (!key->recipe)
                throw std::logic_error("double-hoisted inner sums need keys generated in hybrid mode");
            rot_of[(size_t)k] = (int)perms.size();
            perms.push_back(c.galois_table(elts[k]));
            hks.push_back(hybrid_key(c, key, l));
        }
        double pt_scale = 0;
        for (int i = 0; i < n_giant * n_baby; i++)
        {
            bk_pt_t p = pts[i];
            if (!p)
                continue;
            if (p->ctx != ctx || !p->d || p->limbs != l || p->ext != P.alpha)
                throw std::invalid_argument("plain is not an extended plaintext of the ciphertext's level");
            if (pt_scale == 0)
                pt_scale = p->scale;
            else if (!close_scale(p->scale, pt_scale))
                throw std::invalid_argument("scale mismatch");
        }
        if (pt_scale == 0)
            throw std::invalid_argument("no plaintext operands");
        const double new_scale = in->scale * pt_scale;
        if (!c.scale_in_bounds(new_scale, l))
            throw std::invalid_argument("scale out of bounds");
        for (int g = 0; g < n_giant; g++)
            if (!outs[g] || outs[g]->ctx != ctx || outs[g] == in)
                throw std::invalid_argument("destination is not valid");

        const u64 *c0 = in->d, *c1 = in->d + (size_t)l * n;
        const int count = (int)perms.size();
        Scratch y(s, (size_t)l * n);
        Scratch inter(s, (size_t)std::max({ hyb_chunk(P) * P.dnum, 2 * l, 2 * P.alpha + 2 }) * n);
        Scratch conv(s, (size_t)std::max(hyb_chunk(P) * P.dnum, 2 * l) * n);
        Scratch acc(s, (size_t)std::max(count, 1) * 2 * ne * n);
        Scratch tl(s, (size_t)(2 * P.alpha + 2) * n);
        // rescale != 0: every inner sum leaves divided by q_{l-1} as well (l - 1 limbs) - the rescale that follows the
        // transform moves in front of the giant-step rotations, which then run one level lower, and is one division
        // with the ModDown (hyb_mod_down_rescale)
        if (rescale && l < 2)
            throw std::invalid_argument("end of modulus switching chain reached");
        const bool merged = rescale && P.rescale_tables;
        const int lo = rescale ? l - 1 : l;
        Scratch base0(s, (size_t)l * n), base1(s, (size_t)l * n);
        if (count > 0)
        {
            LdInvPlain ld{ c1, limb_map(l), n, nullptr };
            launch_inv_blocks(c, s, ld, inter.p, l);
            StInvScaled st{ y.p, n, P.d_prescale, P.d_limb_primes };
            launch_inv_cols(c, s, inter.p, st, l);
            hyb_extend_and_mac(c, s, P, h, y.p, conv.p, inter.p, c1, count, perms.data(), hks.data(), acc.p, 1);
        }
        // (1) the key-switched halves, summed in the extended basis: all giant steps multiply the same rotated
        //     ciphertexts, so up to MUL_SUM_GROUPS of them share one pass over those (k_mul_plain_sum_multi)
        const size_t sum_words = (size_t)2 * ne * n;
        Scratch sums(s, (size_t)n_giant * sum_words);
        std::vector<int> ext_terms_of((size_t)n_giant, 0);
        for (int g0 = 0; g0 < n_giant; g0 += MUL_SUM_GROUPS)
        {
            const int ng = std::min(MUL_SUM_GROUPS, n_giant - g0);
            bool first = true;
            for (int k0 = 0; k0 < n_baby;)
            {
                MulSumMultiArgs a{};
                a.count = 0;
                int k = k0, ops = 0;
                for (; k < n_baby && a.count < MUL_SUM_TERMS; k++)
                {
                    if (rot_of[(size_t)k] < 0)
                        continue;
                    bool used = false;
                    for (int g = 0; g < ng; g++)
                        if (pts[(size_t)(g0 + g) * n_baby + k])
                        {
                            a.pt[g][a.count] = pts[(size_t)(g0 + g) * n_baby + k]->d;
                            ext_terms_of[(size_t)(g0 + g)]++;
                            used = true;
                            ops++;
                        }
                    if (used)
                    {
                        a.ct[a.count] = acc.p + (size_t)rot_of[(size_t)k] * 2 * ne * n;
                        a.count++;
                    }
                }
                k0 = k;
                if (!a.count)
                    continue;
                const size_t total2 = sum_words / 2;
                u64 *dst = sums.p + (size_t)g0 * sum_words;
                ProfScope ps_ew(c, s, TAG_ELEMENTWISE, 2 * ne * ops);
#define BK_MULTI(ACC, GG)                                                                                                  \
    launch_pdl(k_mul_plain_sum_multi<ACC, GG>, c.ew_grid(total2), 256, 0, s, dst, sum_words, a, c.d_primes, c.log_n, ne, 2, ne - 1,  \
               c.n_primes - 1)
                switch (ng * 2 + (first ? 0 : 1))
                {
                case 2: BK_MULTI(false, 1); break;
                case 3: BK_MULTI(true, 1); break;
                case 4: BK_MULTI(false, 2); break;
                case 5: BK_MULTI(true, 2); break;
                case 6: BK_MULTI(false, 3); break;
                case 7: BK_MULTI(true, 3); break;
                case 8: BK_MULTI(false, 4); break;
                default: BK_MULTI(true, 4); break;
                }
#undef BK_MULTI
                c.count();
                first = false;
            }
        }
        for (int g = 0; g < n_giant; g++)
        {
            const bk_pt_t *row = pts + (size_t)g * n_baby;
            ensure_ct(outs[g], 2, merged ? lo : l, false);
            const int ext_terms = ext_terms_of[(size_t)g];
            u64 *const sum_g = sums.p + (size_t)g * sum_words;
            // (2) the c0 halves: permutations of the input's c0 (and c1 for the unrotated term), in the ordinary basis
            int plain_terms = 0;
            const bk_pt_s *identity = nullptr;
            for (int k0 = 0; k0 < n_baby;)
            {
                GatherSumArgs a{};
                a.count = 0;
                int k = k0;
                for (; k < n_baby && a.count < GATHER_SUM_TERMS; k++)
                    if (row[k])
                    {
                        a.perm[a.count] = rot_of[(size_t)k] >= 0 ? perms[(size_t)rot_of[(size_t)k]] : nullptr;
                        a.pt[a.count] = row[k]->d;
                        a.count++;
                        if (rot_of[(size_t)k] < 0)
                            identity = row[k];
                    }
                k0 = k;
                if (!a.count)
                    continue;
                ProfScope ps_ew(c, s, TAG_ELEMENTWISE, l * a.count);
                if (plain_terms == 0)
                    launch_pdl(k_gather_mul_sum<false>, c.ew_grid((size_t)l * n / 2), 256, 0, s, base0.p, c0, a, c.d_primes, c.log_n, l);
                else
                    launch_pdl(k_gather_mul_sum<true>, c.ew_grid((size_t)l * n / 2), 256, 0, s, base0.p, c0, a, c.d_primes, c.log_n, l);
                c.count();
                plain_terms += a.count;
            }
            if (!plain_terms)
                throw std::invalid_argument("a giant step has no terms");
            if (identity)
            {
                MulSumArgs a{};
                a.count = 1;
                a.ct[0] = c1;
                a.pt[0] = identity->d;
                ProfScope ps_ew(c, s, TAG_ELEMENTWISE, l);
                launch_pdl(k_mul_plain_sum<false>, c.ew_grid((size_t)l * n / 2), 256, 0, s, base1.p, a, c.d_primes, c.log_n, l, 1, -1, 0);
                c.count();
            }
            bool divided = false;
            if (ext_terms && merged)
            {
                hyb_mod_down_rescale(c, s, P, h, sum_g, conv.p, inter.p, tl.p, outs[g]->d, base0.p, identity ? base1.p : nullptr);
                divided = true;
            }
            else
            {
                if (merged)
                    ensure_ct(outs[g], 2, l, false);
                if (ext_terms)
                    hyb_mod_down(c, s, P, h, sum_g, conv.p, inter.p, tl.p, outs[g]->d, base0.p, identity ? base1.p : nullptr, nullptr);
                else
                { // only the unrotated term: (c0 pt, c1 pt)
                    BK_CUDA(cudaMemcpyAsync(outs[g]->d, base0.p, (size_t)l * n * sizeof(u64), cudaMemcpyDeviceToDevice, s));
                    BK_CUDA(cudaMemcpyAsync(outs[g]->d + (size_t)l * n, base1.p, (size_t)l * n * sizeof(u64), cudaMemcpyDeviceToDevice, s));
                }
            }
            outs[g]->scale = new_scale;
            outs[g]->ntt = true;
            if (divided)
                outs[g]->scale = new_scale / (double)c.primes[(size_t)lo];
            else if (rescale)
                rescale_core(c, outs[g]);
        }
        BK_END
    }
    bk_status bk_complex_conjugate_inplace(bk_context_t ctx, bk_ct_t a, bk_gkeys_t gk)
    {
        BK_TRY
        apply_galois(*ctx, a, (uint32_t)(2 * ctx->n - 1), gk);
        BK_END
    }

    static void plain_check(bk_context_t ctx, bk_ct_t a, bk_pt_t p)
    {
        check_ct(ctx, a, "encrypted");
        if (!p || p->ctx != ctx || !p->d)
            throw std::invalid_argument("plain is not valid for encryption parameters");
        if (!a->ntt)
            throw std::invalid_argument("encrypted is not in NTT form");
        if (a->limbs != p->limbs)
            throw std::invalid_argument("encrypted and plain parameter mismatch");
    }
    bk_status bk_add_plain_inplace(bk_context_t ctx, bk_ct_t a, bk_pt_t p)
    {
        BK_TRY
        Context &c = *ctx;
        plain_check(ctx, a, p);
        if (!close_scale(a->scale, p->scale))
            throw std::invalid_argument("scale mismatch");
        ProfScope ps_ew(c, c.stream(), TAG_ELEMENTWISE);
        launch_pdl(k_ew<EW_ADD>, c.ew_grid((size_t)a->limbs * c.n / 2), 256, 0, c.stream(), a->d, p->d, c.d_primes, c.log_n,
                                                                                    a->limbs, 1, 1);
        c.count();
        BK_END
    }
    bk_status bk_sub_plain_inplace(bk_context_t ctx, bk_ct_t a, bk_pt_t p)
    {
        BK_TRY
        Context &c = *ctx;
        plain_check(ctx, a, p);
        if (!close_scale(a->scale, p->scale))
            throw std::invalid_argument("scale mismatch");
        ProfScope ps_ew(c, c.stream(), TAG_ELEMENTWISE);
        launch_pdl(k_ew<EW_SUB>, c.ew_grid((size_t)a->limbs * c.n / 2), 256, 0, c.stream(), a->d, p->d, c.d_primes, c.log_n,
                                                                                    a->limbs, 1, 1);
        c.count();
        BK_END
    }
    bk_status bk_multiply_plain_inplace(bk_context_t ctx, bk_ct_t a, bk_pt_t p)
    {
        BK_TRY
        Context &c = *ctx;
        plain_check(ctx, a, p);
        double new_scale = a->scale * p->scale;
        if (!c.scale_in_bounds(new_scale, a->limbs))
            throw std::invalid_argument("scale out of bounds");
        ProfScope ps_ew(c, c.stream(), TAG_ELEMENTWISE, a->size * a->limbs);
        launch_pdl(k_ew<EW_MUL>, c.ew_grid((size_t)a->size * a->limbs * c.n / 2), 256, 0, c.stream(), 
            a->d, p->d, c.d_primes, c.log_n, a->limbs, a->size, 1);
        c.count();
        a->scale = new_scale;
        BK_END
    }

    // acc <- acc + a (*) p in one pass (multiply_plain followed by add_inplace: evaluator.cpp:1891-1930, :103-163).
    // An empty acc (size 0) is initialised with the product.  Residues equal those of the two separate calls.
    bk_status bk_multiply_plain_accumulate(bk_context_t ctx, bk_ct_t acc, bk_ct_t a, bk_pt_t p)
    {
        BK_TRY
        Context &c = *ctx;
        plain_check(ctx, a, p);
        if (!acc || acc->ctx != ctx || acc == a)
            throw std::invalid_argument("destination is not valid for encryption parameters");
        double new_scale = a->scale * p->scale;
        if (!c.scale_in_bounds(new_scale, a->limbs))
            throw std::invalid_argument("scale out of bounds");
        const bool first = acc->size == 0 || !acc->d;
        if (!first)
        {
            if (acc->size != a->size || acc->limbs != a->limbs)
                throw std::invalid_argument("encrypted1 and encrypted2 parameter mismatch");
            if (!acc->ntt)
                throw std::invalid_argument("NTT form mismatch");
        }
        else
            ensure_ct(acc, a->size, a->limbs, false);
        const size_t total2 = (size_t)a->size * a->limbs * c.n / 2;
        ProfScope ps_ew(c, c.stream(), TAG_ELEMENTWISE, a->size * a->limbs);
        if (first)
            launch_pdl(k_mul_plain_acc<true>, c.ew_grid(total2), 256, 0, c.stream(), acc->d, a->d, p->d, c.d_primes, c.log_n,
                                                                                  a->limbs, a->size);
        else
            launch_pdl(k_mul_plain_acc<false>, c.ew_grid(total2), 256, 0, c.stream(), acc->d, a->d, p->d, c.d_primes, c.log_n,
                                                                                   a->limbs, a->size);
        c.count();
        acc->scale = new_scale;
        acc->ntt = true;
        BK_END
    }

    bk_status bk_multiply_plain_sum(bk_context_t ctx, bk_ct_t dst, const bk_ct_t *cts, const bk_pt_t *pts, int count)
    {
        BK_TRY
        Context &c = *ctx;
        if (count < 1)
            throw std::invalid_argument("count must be positive");
        if (!dst || dst->ctx != ctx)
            throw std::invalid_argument("destination is not valid for encryption parameters");
        for (int t = 0; t < count; t++)
        {
            plain_check(ctx, cts[t], pts[t]);
            if (cts[t] == dst)
                throw std::invalid_argument("destination must not be one of the operands");
            if (cts[t]->size != cts[0]->size || cts[t]->limbs != cts[0]->limbs || !close_scale(cts[t]->scale, cts[0]->scale) ||
                !close_scale(pts[t]->scale, pts[0]->scale))
                throw std::invalid_argument("encrypted1 and encrypted2 parameter mismatch");
        }
        const double new_scale = cts[0]->scale * pts[0]->scale;
        if (!c.scale_in_bounds(new_scale, cts[0]->limbs))
            throw std::invalid_argument("scale out of bounds");
        const int size = cts[0]->size, limbs = cts[0]->limbs;
        ensure_ct(dst, size, limbs, false);
        const size_t total2 = (size_t)size * limbs * c.n / 2;
        for (int t0 = 0; t0 < count; t0 += MUL_SUM_TERMS)
        {
            MulSumArgs a{};
            a.count = std::min(MUL_SUM_TERMS, count - t0);
            for (int t = 0; t < a.count; t++)
            {
                a.ct[t] = cts[t0 + t]->d;
                a.pt[t] = pts[t0 + t]->d;
            }
            ProfScope ps_ew(c, c.stream(), TAG_ELEMENTWISE, size * limbs * a.count);
            if (t0 == 0)
                launch_pdl(k_mul_plain_sum<false>, c.ew_grid(total2), 256, 0, c.stream(), dst->d, a, c.d_primes, c.log_n, limbs, size, -1, 0);
            else
                launch_pdl(k_mul_plain_sum<true>, c.ew_grid(total2), 256, 0, c.stream(), dst->d, a, c.d_primes, c.log_n, limbs, size, -1, 0);
            c.count();
        }
        dst->scale = new_scale;
        dst->ntt = true;
        BK_END
    }

    bk_status bk_transform_to_ntt_inplace(bk_context_t ctx, bk_ct_t a)
    {
        BK_TRY
        check_ct(ctx, a, "encrypted");
        if (a->ntt)
            throw std::invalid_argument("encrypted is already in NTT form");
        ntt_fwd(*ctx, ctx->stream(), a->d, a->size * a->limbs, limb_map(a->limbs));
        a->ntt = true;
        BK_END
    }
    bk_status bk_transform_from_ntt_inplace(bk_context_t ctx, bk_ct_t a)
    {
        BK_TRY
        check_ct(ctx, a, "encrypted");
        if (!a->ntt)
            throw std::invalid_argument("encrypted is not in NTT form");
        ntt_inv(*ctx, ctx->stream(), a->d, a->size * a->limbs, limb_map(a->limbs));
        a->ntt = false;
        BK_END
    }

    bk_status bk_add_const_inplace(bk_context_t ctx, bk_ct_t a, double value)
    {
        BK_TRY
        // add_const_inplace (evaluator.cpp:287-293): scalar encode at scale = ct.scale, add to c0
        Context &c = *ctx;
        check_ct(ctx, a, "encrypted");
        if (!a->ntt)
            throw std::invalid_argument("encrypted is not in NTT form");
        ScalarPack sp;
        scalar_residues(c, value, a->scale, a->limbs, sp.c);
        ProfScope ps_ew(c, c.stream(), TAG_ELEMENTWISE);
        launch_pdl(k_scalar_pack<false>, c.ew_grid((size_t)a->limbs * c.n / 2), 256, 0, c.stream(), a->d, sp, c.d_primes,
                                                                                            c.log_n, a->limbs, 1);
        c.count();
        BK_END
    }
    bk_status bk_multiply_const_inplace(bk_context_t ctx, bk_ct_t a, double value)
    {
        BK_TRY
        // multiply_const_inplace (evaluator.cpp:295-301): scalar encode at scale = ct.scale
        Context &c = *ctx;
        check_ct(ctx, a, "encrypted");
        if (!a->ntt)
            throw std::invalid_argument("encrypted is not in NTT form");
        ScalarPack sp;
        scalar_residues(c, value, a->scale, a->limbs, sp.c);
        double new_scale = a->scale * a->scale;
        if (!c.scale_in_bounds(new_scale, a->limbs))
            throw std::invalid_argument("scale out of bounds");
        ProfScope ps_ew(c, c.stream(), TAG_ELEMENTWISE, a->size * a->limbs);
        launch_pdl(k_scalar_pack<true>, c.ew_grid((size_t)a->size * a->limbs * c.n / 2), 256, 0, c.stream(), 
            a->d, sp, c.d_primes, c.log_n, a->limbs, a->size);
        c.count();
        a->scale = new_scale;
        BK_END
    }

    bk_status bk_scalar_linear_combination(bk_context_t ctx, bk_ct_t dst, const bk_ct_t *cts, const double *values,
                                           int count, double constant, double target_scale)
    {
        BK_TRY
        // tolerance-mode replacement for the multiply_const + rescale + add_reduced_error chains with which the
        // reference builds each leaf of a polynomial evaluation tree (common/Polynomial.cpp:438-456,
        // comp/SEALfunc.cpp): dst = constant + sum_j values[j] * cts[j] at the lowest level among the sources and at
        // scale target_scale - term j's scalar is encoded at target_scale / cts[j]->scale, so sources at different
        // scales line up exactly - leaving ONE rescale for the caller instead of one per term.
        Context &c = *ctx;
        if (!cts || !values || count < 1 || count > LINCOMB_TERMS)
            throw std::invalid_argument("between 1 and 8 terms are supported");
        int limbs = 1 << 30, polys = 2;
        for (int j = 0; j < count; j++)
        {
            check_ct(ctx, cts[j], "encrypted");
            if (!cts[j]->ntt || (cts[j]->size != 2 && cts[j]->size != 3))
                throw std::invalid_argument("encrypted must be of size 2 or 3 and in NTT form");
            if (cts[j] == dst)
                throw std::invalid_argument("destination must not be one of the sources");
            limbs = std::min(limbs, cts[j]->limbs);
            polys = std::max(polys, cts[j]->size);
        }
        if (!c.scale_in_bounds(target_scale, limbs))
            throw std::invalid_argument("scale out of bounds");
        LinCombArgs a{};
        a.terms = count;
        for (int j = 0; j < count; j++)
        {
            a.src[j] = cts[j]->d;
            a.limbs_in[j] = cts[j]->limbs;
            a.size_in[j] = cts[j]->size;
            scalar_residues(c, values[j], target_scale / cts[j]->scale, limbs, a.c[j]);
        }
        ScalarPack addc{};
        scalar_residues(c, constant, target_scale, limbs, addc.c);
        ensure_ct(dst, polys, limbs, false);
        ProfScope ps_ew(c, c.stream(), TAG_ELEMENTWISE, 2 * limbs * count);
        launch_pdl(k_scalar_lincomb, c.ew_grid((size_t)limbs * c.n * polys / 2), 256, 0, c.stream(), dst->d, a, addc, c.d_primes, c.log_n,
                   limbs, polys);
        c.count();
        dst->scale = target_scale;
        dst->ntt = true;
        BK_END
    }

    bk_status bk_modraise_inplace(bk_context_t ctx, bk_ct_t a)
    {
        BK_TRY
        // Bootstrapper::modraise_inplace (ckks_bootstrapping/Bootstrapper.cpp:2894-2948)
        Context &c = *ctx;
        check_ct(ctx, a, "encrypted");
        if (a->size != 2)
            throw std::invalid_argument("Ciphertexts of size 2 are supported only!");
        if (a->limbs != 1)
            throw std::invalid_argument("Ciphertexts in the lowest level are supported only!");
        cudaStream_t s = c.stream();
        const size_t n = c.n;
        const int L = c.top_limbs();
        if (a->ntt)
        {
            ntt_inv(c, s, a->d, 2, limb_map(1));
            a->ntt = false;
        }
        size_t words = (size_t)2 * L * n;
        Scratch tmp(s, words);
        u64 *out = alloc_words(c, words);
        LdModRaise ld{ a->d, n, L, c.primes[0] };
        launch_fwd_cols(c, s, ld, tmp.p, 2 * L);
        StPlain st{ out, limb_map(L), n };
        launch_fwd_blocks(c, s, tmp.p, st, 2 * L);
        adopt(a, out, words, 2, L);
        a->ntt = true;
        BK_END
    }

    // ---- raw NTT -------------------------------------------------------------------------------
    bk_status bk_ntt_limbs(bk_context_t ctx, uint64_t *dev_data, const int *prime_idx, int count, int inverse)
    {
        BK_TRY
        Context &c = *ctx;
        if (count < 1)
            return BK_OK;
        for (int i = 0; i < count; i++)
            if (prime_idx[i] < 0 || prime_idx[i] >= c.n_primes)
                throw std::invalid_argument("prime index out of range");
        cudaStream_t s = c.stream();
        int *d_idx;
        BK_CUDA(cudaMallocAsync((void **)&d_idx, sizeof(int) * count, s));
        BK_CUDA(cudaMemcpyAsync(d_idx, prime_idx, sizeof(int) * count, cudaMemcpyHostToDevice, s));
        JobMap m = limb_map(count);
        m.explicit_primes = d_idx;
        if (inverse)
            ntt_inv(c, s, (u64 *)dev_data, count, m);
        else
            ntt_fwd(c, s, (u64 *)dev_data, count, m);
        BK_CUDA(cudaFreeAsync(d_idx, s));
        BK_END
    }
    bk_status bk_ntt_limbs_host(bk_context_t ctx, uint64_t *host_data, const int *prime_idx, int count, int inverse)
    {
        BK_TRY
        Context &c = *ctx;
        cudaStream_t s = c.stream();
        Scratch buf(s, (size_t)count * c.n);
        BK_CUDA(cudaMemcpyAsync(buf.p, host_data, (size_t)count * c.n * sizeof(u64), cudaMemcpyHostToDevice, s));
        bk_status st = bk_ntt_limbs(ctx, (uint64_t *)buf.p, prime_idx, count, inverse);
        if (st != BK_OK)
            return st;
        BK_CUDA(cudaMemcpyAsync(host_data, buf.p, (size_t)count * c.n * sizeof(u64), cudaMemcpyDeviceToHost, s));
        BK_CUDA(cudaStreamSynchronize(s));
        BK_END
    }
}
