// engine.h - host-side objects behind the C ABI (include/b200ckks.h).
#pragma once
#include "rng.cuh"
#include "../../include/b200ckks.h"
#include "hostmath.h"
#include "kernels.cuh"
#include <algorithm>
#include <atomic>
#include <cuda_runtime.h>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace bk
{
    struct CudaError : std::runtime_error
    {
        using std::runtime_error::runtime_error;
    };
    struct NoDevice : std::runtime_error
    {
        using std::runtime_error::runtime_error;
    };

#define BK_CUDA(expr)                                                                                                  \
    do                                                                                                                 \
    {                                                                                                                  \
        cudaError_t _e = (expr);                                                                                       \
        if (_e != cudaSuccess)                                                                                         \
            throw bk::CudaError(std::string(#expr) + ": " + cudaGetErrorString(_e));                                   \
    } while (0)

    // Every kernel launch of the engine: cudaLaunchKernelEx with programmatic stream serialization, so that the next
    // kernel of a stream is set up (and its CTAs scheduled, parked at griddepcontrol.wait - pdl_prologue() in
    // modarith.cuh) while the previous one still runs.  Between two dependent launches of a few limb-polynomials this
    // hides 4-5 of the ~12 microseconds (tools/lab/ntt_lab.cu, profiles/r2_ntt_lab.md).  $B200CKKS_NO_PDL=1 launches
    // without the attribute.
    bool pdl_enabled();
    template <class... KArgs, class... Args>
    inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args &&...args)
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid;
        cfg.blockDim = block;
        cfg.dynamicSmemBytes = smem;
        cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = pdl_enabled() ? 1 : 0;
        BK_CUDA(cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...));
    }

    // Scratch memory of one CUDA stream: a stack of device chunks.  Scratch objects are automatic variables, so their
    // lifetimes nest; every kernel that touches a scratch buffer is enqueued on the stream the arena belongs to, hence a
    // block popped by one operation can be handed to the next operation on that stream without any synchronisation -
    // and, unlike cudaMallocAsync / cudaFreeAsync per call (5 per key switch), without the allocator ordering one
    // host thread's stream behind another's to reuse a freed block.
    struct ScratchArena
    {
        std::vector<std::pair<char *, size_t>> chunks;
        size_t cur = 0, off = 0;
    };
    ScratchArena *arena_of(cudaStream_t s); // engine.cu: arenas of the streams the contexts created; null for foreign streams

    struct Scratch
    {
        u64 *p = nullptr;
        cudaStream_t s;
        ScratchArena *a = nullptr;
        size_t save_cur = 0, save_off = 0;
        Scratch(cudaStream_t stream, size_t words) : s(stream)
        {
            a = arena_of(stream);
            if (!a)
            {
                BK_CUDA(cudaMallocAsync((void **)&p, words * sizeof(u64), s));
                return;
            }
            const size_t bytes = (words * sizeof(u64) + 255) & ~size_t(255);
            save_cur = a->cur;
            save_off = a->off;
            for (;;)
            {
                if (a->cur < a->chunks.size() && a->off + bytes <= a->chunks[a->cur].second)
                {
                    p = reinterpret_cast<u64 *>(a->chunks[a->cur].first + a->off);
                    a->off += bytes;
                    return;
                }
                if (a->cur + 1 < a->chunks.size())
                {
                    a->cur++;
                    a->off = 0;
                    continue;
                }
                const size_t last = a->chunks.empty() ? 0 : a->chunks.back().second;
                const size_t cap = std::max(bytes, std::max<size_t>(size_t(64) << 20, 2 * last));
                char *chunk = nullptr;
                BK_CUDA(cudaMalloc((void **)&chunk, cap));
                a->chunks.emplace_back(chunk, cap);
                a->cur = a->chunks.size() - 1;
                a->off = 0;
            }
        }
        ~Scratch()
        {
            if (a)
            {
                a->cur = save_cur;
                a->off = save_off;
            }
            else if (p)
                cudaFreeAsync(p, s);
        }
        Scratch(const Scratch &) = delete;
        Scratch &operator=(const Scratch &) = delete;
    };

    struct EncoderState; // encoder.cu

    // Pinned host staging for slot vectors on their way to the encoder: the caller's (pageable) buffer is copied
    // into the next slot of a small ring and leaves with cudaMemcpyAsync, so an encode never has to drain the
    // stream before it returns.  A slot is reused once the event recorded after its copy has completed.
    struct StagingRing
    {
        static constexpr int SLOTS = 8;
        char *host = nullptr;
        size_t slot_bytes = 0;
        cudaEvent_t done[SLOTS] = {};
        int next = 0;
    };

    // Level-aware hybrid key switching (tolerance mode): constants of one level, see kernels.cuh (HybDims) and
    // engine.cu (hybrid_plan / key_switch_hybrid).  All pointers are device memory owned by the context.
    struct HybridPlan
    {
        int l = 0, alpha = 1, dsize = 1, dnum = 0, ne = 0; // alpha special moduli, digits of dsize primes
        ulonglong2 *d_prescale = nullptr;  // [l] {(Q_d / q_i)^-1 mod q_i, shoup}
        int *d_limb_primes = nullptr;      // [l] identity map for the scaled INTT store
        u64 *d_w = nullptr;                // [ne][dnum][dsize] (Q_d / q_i) mod p_e
        ulonglong2 *d_sprescale = nullptr; // [2 alpha] {(P_S / p_a)^-1 mod p_a, shoup}, repeated for both polynomials
        int *d_sprimes = nullptr;          // [2 alpha] prime index of each special limb
        u64 *d_ws = nullptr;               // [l][alpha] (P_S / p_a) mod q_i
        ulonglong2 *d_psinv = nullptr;     // [l] {P_S^-1 mod q_i, shoup}
        u64 *d_keyfactor = nullptr;        // [dnum][ne] P_S mod q_e on the digit's own limbs, 0 elsewhere
        // the division by D (= P_S here) rounds to nearest and the basis conversion is exact: floor(D/2) joins the
        // dropped residues before the conversion, the multiple u D by which the fast conversion overshoots is found
        // from the fractional parts sum_a y_a / p_a in double precision and taken off again (k_hyb_conv<DS, true>).
        // Without it every coefficient carries the same negative offset of a few units, and a constant polynomial
        // times s(X) is ~2N/pi ~ 4e4 times larger in the slots whose root of unity lies next to 1 than random
        // rounding noise of the same size (measured: 1e-6 instead of 3e-9 on a rescaled product).
        u64 *d_shalf = nullptr;            // [2 alpha] floor(D/2) mod p_a, both polynomials
        double *d_spinv = nullptr;         // [alpha] 1 / p_a
        u64 *d_negd = nullptr;             // [l] -D mod q_i
        u64 *d_addc = nullptr;             // [l] -floor(D/2) mod q_i
        // ModDown and rescale as ONE division by D = q_{l-1} * P_S (relinearization followed by a rescale): the dropped
        // basis is {q_{l-1}} + the alpha special moduli, the target limbs are 0 .. l-2.  Present when alpha + 1 <= 17.
        bool rescale_tables = false;
        ulonglong2 *d_r_sprescale = nullptr; // [2 (alpha+1)] {(D / p_a)^-1 mod p_a, shoup}, both polynomials
        int *d_r_sprimes = nullptr;          // [2 (alpha+1)] prime index of each dropped limb (q_{l-1} first)
        u64 *d_r_ws = nullptr;               // [l-1][alpha+1] (D / p_a) mod q_i
        ulonglong2 *d_r_dinv = nullptr;      // [l-1] {D^-1 mod q_i, shoup}
        ulonglong2 *d_r_qlinv = nullptr;     // [l-1] {q_{l-1}^-1 mod q_i, shoup}: the factor of the base ciphertext
        ulonglong2 r_pmod{ 0, 0 };           // {P_S mod q_{l-1}, shoup}: the base joins the dropped limb l-1 times P_S
        u64 *d_r_shalf = nullptr;            // as d_shalf .. d_addc for D = q_{l-1} P_S
        double *d_r_spinv = nullptr;
        u64 *d_r_negd = nullptr;
        u64 *d_r_addc = nullptr;
        HybridPlan() = default;
        HybridPlan(const HybridPlan &) = delete;
        HybridPlan &operator=(const HybridPlan &) = delete;
        ~HybridPlan() // owns its tables (also the copy that loses the publication race in hybrid_plan())
        {
            cudaFree(d_prescale);
            cudaFree(d_limb_primes);
            cudaFree(d_w);
            cudaFree(d_sprescale);
            cudaFree(d_sprimes);
            cudaFree(d_ws);
            cudaFree(d_psinv);
            cudaFree(d_keyfactor);
            cudaFree(d_r_sprescale);
            cudaFree(d_r_sprimes);
            cudaFree(d_r_ws);
            cudaFree(d_r_dinv);
            cudaFree(d_r_qlinv);
            cudaFree(d_shalf);
            cudaFree(d_spinv);
            cudaFree(d_negd);
            cudaFree(d_addc);
            cudaFree(d_r_shalf);
            cudaFree(d_r_spinv);
            cudaFree(d_r_negd);
            cudaFree(d_r_addc);
        }
    };

    struct Context
    {
        int log_n = 0;
        size_t n = 0;
        int n_primes = 0; // key-level chain, special prime last
        int device = 0;
        int sm_count = 148;
        std::vector<uint64_t> primes;
        std::vector<int> total_bits; // total_bits[l] = bit count of q_0 * ... * q_{l-1}
        PrimeDev *d_primes = nullptr;
        std::vector<PrimeDev> h_primes;
        ulonglong2 *d_tw = nullptr, *d_itw = nullptr;
        ulonglong2 *d_inv = nullptr; // [n_primes(last)][n_primes(i)] {q_last^-1 mod q_i, shoup}
        NttTables tables{};
        int ks_chunk = 4;
        int sparse_slots = 0;
        // 256-bit master key of the random generator (rng.cuh): getrandom(2), or expanded from $B200CKKS_SEED /
        // bk_context_set_rng_key for reproducible runs
        RngKey rng_master{};
        // hybrid key switching: generated keys are level-specific and not in SEAL's layout ($B200CKKS_HYBRID_KS=1 or
        // bk_context_set_hybrid before any key is generated)
        bool hybrid = false;
        // seed-compressed level keys ($B200CKKS_COMPRESS_KEYS=1 or bk_context_set_key_compression before keys are
        // generated): only polynomial 0 of every digit is resident, the uniform polynomial is regenerated from its
        // public ChaCha8 key when the key is used
        bool compress_keys = false;
        std::map<int, HybridPlan *> hplans;
        std::atomic<uint64_t> hybrid_key_bytes{ 0 }, hybrid_keys{ 0 };
        std::atomic<uint64_t> launches{ 0 };
        // bytes moved between host and device by C-ABI calls (uploads, downloads, encode inputs, decode outputs)
        std::atomic<uint64_t> h2d_bytes{ 0 }, d2h_bytes{ 0 };
        // per kernel family (KernelTag): launches and limb-polynomials processed, always counted
        std::atomic<uint64_t> tag_launches[8] = {}, tag_units[8] = {};
        // per-kernel live timing (bk_profile_begin/end): events around every launch of one tag
        int prof_tag = -1; // -1 off, -2 every family, else one KernelTag
        std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
        std::vector<int> prof_event_tags;
        std::vector<cudaEvent_t> timer_stack;

        std::mutex mu;
        std::unordered_map<std::thread::id, cudaStream_t> streams;
        std::unordered_map<std::thread::id, StagingRing *> staging_rings;
        std::unordered_map<uint32_t, uint32_t *> galois_tables; // device tables

        EncoderState *enc = nullptr; // built lazily on first encode/decode (encoder.cu)
        std::mutex enc_mu;

        Context(int log_n, const uint64_t *primes, int n_primes, int device);
        ~Context();
        cudaStream_t stream();
        // the calling thread's stream if it has one (does not create it: a thread that only destroys objects must not
        // acquire a stream and a scratch arena)
        cudaStream_t stream_if_any();
        // returns `d` to the stream-ordered pool.  `owner` is the stream the buffer was allocated on; a release from
        // another thread cannot know who else still reads it, so it waits for the whole device first (rare: hand-over
        // of results between threads, teardown of shared caches)
        void release_words(void *d, cudaStream_t owner);
        // next free pinned slot of the calling thread's ring (>= bytes); *done_out must be recorded on the stream
        // after the copy that reads the slot
        char *staging(size_t bytes, cudaEvent_t *done_out);
        const uint32_t *galois_table(uint32_t elt);
        void count(int k = 1)
        {
            launches.fetch_add((uint64_t)k, std::memory_order_relaxed);
            if (debug_sync)
                debug_check();
        }
        // $B200CKKS_DEBUG_SYNC=1: wait for the device after every launch and report the first failing one with a host
        // backtrace (compute-sanitizer is not available on the GPU pool)
        bool debug_sync = false;
        void debug_check();
        int top_limbs() const
        {
            return n_primes - 1;
        }
        bool scale_in_bounds(double scale, int limbs) const;
        const ulonglong2 *inv_last(int last_prime) const
        {
            return d_inv + (size_t)last_prime * n_primes;
        }
        int ew_grid(size_t work_items) const;
        void activate() const;
    };

    enum KernelTag
    {
        TAG_FWD_COLS = 0,
        TAG_FWD_BLOCKS = 1,
        TAG_INV_BLOCKS = 2,
        TAG_INV_COLS = 3,
        TAG_KS_MAC = 4,
        TAG_ELEMENTWISE = 5,
        TAG_FFT = 6,   // encoder / decoder complex FFT passes
        TAG_OTHER = 7, // basis conversions of hybrid key switching (k_hyb_conv); samplers, gathers, CRT composition
        TAG_COUNT = 8
    };
    // RAII: records an event pair around one launch when its tag is being profiled
    struct ProfScope
    {
        Context &c;
        cudaStream_t s;
        cudaEvent_t stop = nullptr;
        ProfScope(Context &ctx, cudaStream_t stream, int tag, int units = 1) : c(ctx), s(stream)
        {
            c.tag_launches[tag].fetch_add(1, std::memory_order_relaxed);
            c.tag_units[tag].fetch_add((uint64_t)units, std::memory_order_relaxed);
            if (c.prof_tag != tag && c.prof_tag != -2)
                return;
            cudaEvent_t start;
            cudaEventCreate(&start);
            cudaEventCreate(&stop);
            cudaEventRecord(start, s);
            c.prof_events.emplace_back(start, stop);
            c.prof_event_tags.push_back(tag);
        }
        ~ProfScope()
        {
            if (stop)
                cudaEventRecord(stop, s);
        }
    };

    // ---- internal entry points shared between the .cu files ------------------------------------
    // NTT of `jobs` limb-polynomials (natural layout) in place; prime of job j = map.prime(j)
    void ntt_fwd(Context &c, cudaStream_t s, u64 *data, int jobs, JobMap map);
    void ntt_inv(Context &c, cudaStream_t s, u64 *data, int jobs, JobMap map);
    JobMap limb_map(int limbs);
    JobMap key_map(const Context &c); // limbs 0..n_primes-1 incl. special
    void ensure_ct(bk_ct_t ct, int size, int limbs, bool keep);
    void ensure_pt(bk_pt_t pt, int limbs);
    void destroy_encoder(Context &c);
    void rescale_core(Context &c, bk_ct_t a);
    // hybrid key switching: constants of level l (engine.cu) and the level-l key of a key recipe (keygen.cu)
    void hybrid_shape(int l, int top, int &alpha, int &dsize);
    const HybridPlan &hybrid_plan(Context &c, int l);
} // namespace bk
struct bk_kskey_s;
struct bk_hybkey_s;
namespace bk
{
    bk_hybkey_s *hybrid_key(Context &c, bk_kskey_s *key, int l);
}

struct bk_context_s : bk::Context
{
    using bk::Context::Context;
};

struct bk_ct_s
{
    bk::Context *ctx;
    u64 *d = nullptr;
    cudaStream_t owner = nullptr; // stream `d` was allocated on
    size_t cap = 0; // words
    int size = 0, limbs = 0;
    double scale = 1.0;
    bool ntt = true;
};

struct bk_pt_s
{
    bk::Context *ctx;
    u64 *d = nullptr;
    cudaStream_t owner = nullptr;
    size_t cap = 0;
    int limbs = 0;
    int ext = 0; // > 0: `ext` further limbs follow the first `limbs` (bk_encode_ext): the special moduli of that level
    double scale = 1.0;
};

struct bk_sk_s;
struct bk_hybkey_s
{
    u64 *d = nullptr; // [dnum][2][l + alpha][N]: limbs 0 .. l + alpha - 2, then the special prime
                      // compressed: [dnum][l + alpha][N], polynomial 0 only - polynomial 1 is (akey, astream0 + digit)
    int l = 0, alpha = 0, dsize = 0, dnum = 0;
    size_t words = 0;
    bool compressed = false;
    bk::RngKey akey{};     // PUBLIC key of the uniform halves (keygen.cu: public_uniform8)
    unsigned astream0 = 0;
};
namespace bk
{
    void expand_public_halves(Context &c, cudaStream_t s, const bk_hybkey_s *hk, int limb0, int nlimbs, u64 *out);
}
struct bk_kskey_s
{
    bk::Context *ctx;
    u64 *d = nullptr; // [digits][2][klimbs+1][N], the special prime's limb at index klimbs
    int digits = 0, klimbs = 0;
    size_t words = 0;
    // hybrid mode: the key is a recipe (secret key, what it switches from, seed) and one level-specific key per level
    // it has been used at, generated on first use
    bk_sk_s *sk = nullptr; // null once detached (bk_kskey_drop_secret): only the level keys made so far remain usable
    bool recipe = false;   // generated in hybrid mode: no SEAL-layout data, level keys in `hyb`
    int kind = 0;       // 0 uploaded / SEAL layout only, 1 relinearization (s^2), 2 Galois (elt)
    uint32_t elt = 0;
    uint64_t seed = 0;
    std::mutex hmu;
    std::map<int, bk_hybkey_s *> hyb;
    const bk_hybkey_s *view_of = nullptr; // a transient SEAL-shaped view of this level key (key_switch): may be compressed
};

struct bk_gkeys_s
{
    bk::Context *ctx;
    std::mutex mu;
    std::map<uint32_t, bk_kskey_t> keys;
};

struct bk_sk_s
{
    bk::Context *ctx;
    u64 *d = nullptr; // [n_primes][N] NTT form
};

// ---- C-ABI exception fence ---------------------------------------------------------------------
namespace bk
{
    void set_error(const char *msg);
    bk_status fence(const std::exception_ptr &e);
}
#define BK_TRY try {
#define BK_END                                                                                                         \
    }                                                                                                                  \
    catch (...)                                                                                                        \
    {                                                                                                                  \
        return bk::fence(std::current_exception());                                                                    \
    }                                                                                                                  \
    return BK_OK;
