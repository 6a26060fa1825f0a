// kernels.cuh - load/store functors fused into the NTT passes, the key-switch inner-product
// kernel and the element-wise limb kernels.  Reference loops replaced (all under
// seal-modified-3.6.6/native/src/seal/):
//   util/polyarithsmallmod.cpp:18-169   add/sub/negate/dyadic_product/multiply_poly_scalar
//   util/rns.cpp:737-808                divide_and_round_q_last_ntt_inplace (rescale)
//   evaluator.cpp:2281-2525             switch_key_inplace (decompose, NTT, MAC, ModDown)
//   util/galois.cpp:192-218             apply_galois_ntt (gather)
#pragma once
#include "ntt.cuh"

// job -> prime index.  A "job" is one limb-polynomial; job % limbs is the limb index which is
// also the prime index, except for one optional position mapped to the special prime.
struct JobMap
{
    int limbs;
    int special_pos;   // limb index that maps to special_prime, or -1
    int special_prime;
    const int *explicit_primes; // optional [jobs] table (raw NTT API)
    __device__ __forceinline__ int prime(int job) const
    {
        if (explicit_primes)
            return explicit_primes[job];
        int l = job % limbs;
        return l == special_pos ? special_prime : l;
    }
};

__device__ __forceinline__ u64 reduce4q(u64 v, const PrimeDev &pd)
{
    v = csub(v, pd.two_q);
    return csub(v, pd.q);
}

// ---------------------------------------------------------------- forward column-pass loaders
struct LdPlain
{
    const u64 *src;
    JobMap map;
    size_t n;
    __device__ __forceinline__ bool skip(int) const { return false; }
    __device__ __forceinline__ int prime(int job) const { return map.prime(job); }
    __device__ __forceinline__ u64 load(int job, int idx, const PrimeDev &) const
    {
        return src[(size_t)job * n + idx];
    }
};

// Key-switch digit decomposition (evaluator.cpp:2386-2408): job = iloc * l + J; output modulus
// I = I0 + iloc (I == l -> special prime); digit J of the INTT'd target is reduced mod q_I only
// when q_J > q_I; the I == J product uses the NTT-form input directly, so that job is skipped.
struct LdKsDigit
{
    const u64 *ttarget; // [l][N] coefficient form, canonical
    const PrimeDev *primes;
    size_t n;
    int l, I0, special_prime;
    __device__ __forceinline__ bool skip(int job) const { return (I0 + job / l) == (job % l); }
    __device__ __forceinline__ int prime(int job) const
    {
        int I = I0 + job / l;
        return I == l ? special_prime : I;
    }
    __device__ __forceinline__ u64 load(int job, int idx, const PrimeDev &pd) const
    {
        int J = job % l;
        u64 v = ttarget[(size_t)J * n + idx];
        if (primes[J].q > pd.q)
            v = barrett64(v, pd);
        return v;
    }
};

// Divide-and-round source (rns.cpp:766-784, evaluator.cpp:2480-2497): job = p * limbs_out + i;
// value = (tlast[p] mod q_i) + (q_i - (half mod q_i)), half = q_last >> 1.
struct LdDivRound
{
    const u64 *tlast; // [polys][N] coefficient form, canonical mod q_last, half already added
    size_t n;
    int limbs_out;
    u64 qlast;
    __device__ __forceinline__ bool skip(int) const { return false; }
    __device__ __forceinline__ int prime(int job) const { return job % limbs_out; }
    __device__ __forceinline__ u64 load(int job, int idx, const PrimeDev &pd) const
    {
        int p = job / limbs_out;
        u64 v = tlast[(size_t)p * n + idx];
        if (qlast > pd.q)
            v = barrett64(v, pd);
        u64 fix = pd.q - barrett64(qlast >> 1, pd);
        return v + fix;
    }
};

// ModRaise (Bootstrapper.cpp:2894-2948): job = p * limbs_out + i; source = coefficient-form
// 1-limb ciphertext; centred lift of the q0 residue to q_i.
struct LdModRaise
{
    const u64 *c; // [polys][N] coefficient form mod q0
    size_t n;
    int limbs_out;
    u64 q0;
    __device__ __forceinline__ bool skip(int) const { return false; }
    __device__ __forceinline__ int prime(int job) const { return job % limbs_out; }
    __device__ __forceinline__ u64 load(int job, int idx, const PrimeDev &pd) const
    {
        int p = job / limbs_out;
        u64 v = c[(size_t)p * n + idx];
        // value in (-q0/2, q0/2]: v > q0/2 represents v - q0
        if (v > (q0 >> 1))
        {
            u64 neg = q0 - v; // magnitude
            u64 r = barrett64(neg, pd);
            return r ? pd.q - r : 0ull;
        }
        return barrett64(v, pd);
    }
};

// ---------------------------------------------------------------- forward block-pass stores
struct StPlain
{
    static constexpr bool RAW = false;
    static constexpr bool BATCH = false;
    __device__ __forceinline__ bool skip(int) const { return false; }
    u64 *dst;
    JobMap map;
    size_t n;
    __device__ __forceinline__ int prime(int job) const { return map.prime(job); }
    __device__ __forceinline__ u64 pre(int, int, int, int, u64 v, const PrimeDev &) const { return v; }
    __device__ __forceinline__ void post(int job, int idx, u64 v, const PrimeDev &pd) const
    {
        dst[(size_t)job * n + idx] = reduce4q(v, pd);
    }
};

// Key-switch digits: keep the unreduced NTT output in place (the reference multiplies its lazy
// values straight into the key, evaluator.cpp:2407-2436).  Range: [0,4q) from the lazy butterflies
// (primes >= 2^57), [0,70q) from the unreduced "wide" butterflies (primes < 2^57, ntt.cuh).  Either way
// the 128-bit sums of k_ks_mac hold: 61 * 4 * 2^120 < 2^128 and 61 * 70 * 2^114 < 2^128 (primes are
// limited to 60 bits at context creation).  job = iloc * l + J as in
// LdKsDigit; the I == J job is skipped (its operand is the NTT-form input itself).
struct StKsDigit
{
    static constexpr bool RAW = true; // k_ks_mac multiplies whatever 64-bit representative it is given
    static constexpr bool BATCH = false;
    u64 *dst;
    size_t n;
    int l, I0, special_prime;
    __device__ __forceinline__ bool skip(int job) const { return (I0 + job / l) == (job % l); }
    __device__ __forceinline__ int prime(int job) const
    {
        int I = I0 + job / l;
        return I == l ? special_prime : I;
    }
    __device__ __forceinline__ u64 pre(int, int, int, int, u64 v, const PrimeDev &) const { return v; }
    __device__ __forceinline__ void post(int job, int idx, u64 v, const PrimeDev &) const
    {
        dst[(size_t)job * n + idx] = v;
    }
};

// Rescale tail (rns.cpp:786-806): dst[p][i] = (x[p][i] - NTT_i(t)) * q_last^-1 mod q_i.
struct StRescale
{
    static constexpr bool RAW = false;
    static constexpr bool BATCH = true; // post_all: the 16 loads of a thread in flight together
    __device__ __forceinline__ bool skip(int) const { return false; }
    const u64 *x;  // [polys][limbs_in][N]
    u64 *dst;      // [polys][limbs_out][N]
    const ulonglong2 *inv; // [n_primes] {q_last^-1 mod q_i, shoup}
    size_t n;
    int limbs_in, limbs_out;
    __device__ __forceinline__ int prime(int job) const { return job % limbs_out; }
    __device__ __forceinline__ u64 pre(int, int, int, int, u64 v, const PrimeDev &) const { return v; }
    // coefficients base + t + 16 k, k < 16, their transform values at s[swz(t + 16 k)] (in [0,4q))
    __device__ __forceinline__ void post_all(int job, int base, int t, const u64 *s, const PrimeDev &pd) const
    {
        const int p = job / limbs_out, i = job % limbs_out;
        const u64 *src = x + ((size_t)p * limbs_in + i) * n + base + t;
        u64 *out = dst + (size_t)job * n + base + t;
        const ulonglong2 f = __ldg(inv + i);
        u64 r[16];
#pragma unroll
        for (int k = 0; k < 16; k++)
            r[k] = __ldg(src + 16 * k);
#pragma unroll
        for (int k = 0; k < 16; k++)
        {
            u64 d = r[k] + 2 * pd.two_q - s[swz(t + 16 * k)];
            out[16 * k] = csub(mul_shoup_lazy(d, f.x, f.y, pd.q), pd.q);
        }
    }
};

// ModDown tail of the key switch (evaluator.cpp:2499-2522):
//   dst[p][i] = base_p[i] + (acc[p][i] - NTT_i(t_p)) * P^-1 mod q_i
// acc is in natural layout (written by k_ks_mac); register element (t,k) of a block is
// coefficient 16t + k.  base_0 = c0 (optionally gathered through the Galois table), base_1 = c1
// or nothing.
template <bool SCALE_BASE> struct StModDownT
{
    static constexpr bool RAW = false;
    static constexpr bool BATCH = true; // post_all: three dependent loads per coefficient, 16 coefficients in flight
    __device__ __forceinline__ bool skip(int) const { return false; }
    const u64 *acc;  // [2][l+1][N]
    u64 *dst;        // [2][l][N]
    const u64 *base0; // [l][N] or null
    const u64 *base1; // [l][N] or null
    const uint32_t *perm; // Galois table or null (applies to base0 only)
    const ulonglong2 *inv; // [n_primes] {P^-1 mod q_i, shoup}
    size_t n;
    int l;
    int acc_limbs = 0;     // limbs per polynomial of acc; 0 = l + 1 (one special prime)
    const ulonglong2 *base_scale = nullptr; // SCALE_BASE: [l] per-limb factor of the base (merged rescale: q_last^-1)
    __device__ __forceinline__ int prime(int job) const { return job % l; }
    __device__ __forceinline__ u64 pre(int, int, int, int, u64 v, const PrimeDev &) const { return v; }
    // Coefficients base + t + 16 k, k < 16, in the coalesced order of the store (acc, base and dst as full lines),
    // their transform values at s[swz(t + 16 k)].  The loads are issued in three batches - Galois indices,
    // accumulator, base - so that a thread waits for three DRAM latencies, not for 48 (the per-coefficient form did:
    // 100 us for 58 limb-polynomials at 18 % of DRAM bandwidth, profiles/r2_ncu_full_moddown.md).  acc / base never
    // overlap dst (the callers of a rotation write to a fresh buffer), hence the read-only loads.
    __device__ __forceinline__ void post_all(int job, int base, int t, const u64 *s, const PrimeDev &pd) const
    {
        const int p = job / l, i = job % l;
        const int al = acc_limbs ? acc_limbs : l + 1;
        const u64 *a = acc + ((size_t)p * al + i) * n + base + t;
        const u64 *b = p == 0 ? base0 : base1;
        u64 *out = dst + (size_t)job * n + base + t;
        const ulonglong2 f = __ldg(inv + i);
        ulonglong2 g = make_ulonglong2(0, 0);
        if (SCALE_BASE)
            g = __ldg(base_scale + i);
        if (b)
            b += (size_t)i * n;
        // two batches of 8 coefficients: enough loads in flight to cover the latency, few enough registers not to spill
#pragma unroll
        for (int h = 0; h < 16; h += 8)
        {
            int src[8];
            if (p == 0 && perm)
            {
#pragma unroll
                for (int k = 0; k < 8; k++)
                    src[k] = (int)__ldg(perm + base + t + 16 * (h + k));
            }
            else
            {
#pragma unroll
                for (int k = 0; k < 8; k++)
                    src[k] = base + t + 16 * (h + k);
            }
            u64 r[8], add[8];
#pragma unroll
            for (int k = 0; k < 8; k++)
                r[k] = __ldg(a + 16 * (h + k));
            if (b)
            {
#pragma unroll
                for (int k = 0; k < 8; k++)
                    add[k] = __ldg(b + src[k]);
            }
            else
            {
#pragma unroll
                for (int k = 0; k < 8; k++)
                    add[k] = 0;
            }
#pragma unroll
            for (int k = 0; k < 8; k++)
            {
                u64 d = r[k] + 2 * pd.two_q - s[swz(t + 16 * (h + k))]; // transform value in [0,4q)
                u64 v = csub(mul_shoup_lazy(d, f.x, f.y, pd.q), pd.q);
                if (SCALE_BASE)
                    add[k] = csub(mul_shoup_lazy(add[k], g.x, g.y, pd.q), pd.q);
                out[16 * (h + k)] = addmod(v, add[k], pd.q);
            }
        }
    }
};
using StModDown = StModDownT<false>;
using StModDownRescale = StModDownT<true>; // hyb_mod_down_rescale (engine.cu)

// ---------------------------------------------------------------- inverse block-pass loaders
struct LdInvPlain
{
    static constexpr bool TLAYOUT = false;
    const u64 *src;
    JobMap map;
    size_t n;
    const uint32_t *perm; // optional Galois gather (apply_galois_ntt fused into the INTT load)
    __device__ __forceinline__ int prime(int job) const { return map.prime(job); }
    __device__ __forceinline__ u64 load(int job, int idx, const PrimeDev &) const
    {
        int s = perm ? (int)perm[idx] : idx;
        return src[(size_t)job * n + s];
    }
    __device__ __forceinline__ u64 load_t(int, int, int) const { return 0; }
};

// strided source: job p reads limb `limb` of poly p in a [polys][limbs][N] array
struct LdInvLimbOf
{
    static constexpr bool TLAYOUT = false;
    const u64 *src;
    size_t n;
    int limbs, limb, prime_idx;
    __device__ __forceinline__ int prime(int) const { return prime_idx; }
    __device__ __forceinline__ u64 load(int job, int idx, const PrimeDev &) const
    {
        return src[((size_t)job * limbs + limb) * n + idx];
    }
    __device__ __forceinline__ u64 load_t(int, int, int) const { return 0; }
};

// ---------------------------------------------------------------- inverse column-pass stores
struct StInvPlain
{
    u64 *dst;
    JobMap map;
    size_t n;
    __device__ __forceinline__ int prime(int job) const { return map.prime(job); }
    __device__ __forceinline__ void store(int job, int idx, u64 v, const PrimeDev &pd) const
    {
        dst[(size_t)job * n + idx] = csub(v, pd.q);
    }
};

// canonical INTT output + floor(q/2), reduced (rns.cpp:759-764, evaluator.cpp:2471-2478)
struct StInvAddHalf
{
    u64 *dst;
    size_t n;
    int prime_idx;
    __device__ __forceinline__ int prime(int) const { return prime_idx; }
    __device__ __forceinline__ void store(int job, int idx, u64 v, const PrimeDev &pd) const
    {
        v = csub(v, pd.q);
        dst[(size_t)job * n + idx] = csub(v + (pd.q >> 1), pd.q);
    }
};

// ============================================================================================
// Key-switch inner product (evaluator.cpp:2368-2463).  The digit NTTs are complete when this
// runs (column pass LdKsDigit + block pass StKsDigit, unreduced values); this kernel is the
// pure stream  acc_p[I][i] = sum_J digit[I][J][i] * key[J][p][I][i]  with 128-bit accumulators
// and one final barrett_reduce_128.  The key (2 l (l+1) limb-polys, up to 1 GiB) is the
// dominant HBM stream of the whole key switch: 16-byte loads, two coefficients per thread,
// J loop unrolled so several independent 128-byte lines are in flight per warp.
//   operand for J == I: target_ntt[J] gathered through the Galois table (if any).
// grid = (ceil(N/2/256), nI), block = 256.
// ============================================================================================
struct KsMacArgs
{
    const u64 *digits;     // [nI][l][N] NTT form, unreduced (see StKsDigit); slot J == I unused
    const u64 *target_ntt; // [l][N] NTT form
    const uint32_t *perm;  // Galois table or null
    const u64 *key;        // [digits][2][klimbs+1][N], special prime's limb at index klimbs
    u64 *acc;              // [2][l+1][N]
    size_t n;
    int l, I0, special_prime, klimbs;
    int gather_digits;     // hoisted rotations: the digits were computed before the automorphism, read them through perm
    // polynomial 1 of digit J, limb kl at key1 + J * dstride1 + kl * N (polynomial 0: key + J * dstride0 + kl * N): the
    // same array, or a buffer expanded from the key's seed just before (seed-compressed level keys)
    const u64 *key1;
    size_t dstride0, dstride1;
};

__device__ __forceinline__ ulonglong2 ldg_stream2(const u64 *p)
{
    ulonglong2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0,%1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p));
    return v;
}
// Evaluation keys are read exactly once per key switch while the digits they are multiplied with are re-read for every
// key (every baby rotation of a hoisted group) and every output modulus: key loads carry an L2 evict-first policy so
// that hundreds of megabytes of key do not push the digit buffer out of the 126 MB L2.
__device__ __forceinline__ u64 l2_evict_first_policy()
{
    u64 p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ ulonglong2 ldg_stream2(const u64 *p, u64 policy)
{
    ulonglong2 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u64 {%0,%1}, [%2], %3;"
                 : "=l"(v.x), "=l"(v.y)
                 : "l"(p), "l"(policy));
    return v;
}

static __global__ void __launch_bounds__(256) k_ks_mac(KsMacArgs a, NttTables T)
{
    pdl_prologue();
    const int iloc = blockIdx.y;
    const int I = a.I0 + iloc;
    const int pi = I == a.l ? a.special_prime : I;
    const int kl = I == a.l ? a.klimbs : I;
    const size_t n = a.n;
    const size_t e = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (e >= n)
        return;
    const PrimeDev pd = T.primes[pi];
    const u64 *dig = a.digits + (size_t)iloc * a.l * n + e;
    const u64 *k0 = a.key + (size_t)kl * n + e;
    const u64 *k1 = a.key1 + (size_t)kl * n + e;

    u64 l0x = 0, h0x = 0, l0y = 0, h0y = 0, l1x = 0, h1x = 0, l1y = 0, h1y = 0;
    const u64 key_policy = l2_evict_first_policy();
#pragma unroll 4
    for (int J = 0; J < a.l; J++)
    {
        ulonglong2 x;
        if (J == I)
        {
            const u64 *src = a.target_ntt + (size_t)J * n;
            if (a.perm)
            {
                x.x = src[a.perm[e]];
                x.y = src[a.perm[e + 1]];
            }
            else
                x = *reinterpret_cast<const ulonglong2 *>(src + e);
        }
        else if (a.gather_digits)
        {
            const u64 *src = a.digits + ((size_t)iloc * a.l + J) * n;
            x.x = src[a.perm[e]];
            x.y = src[a.perm[e + 1]];
        }
        else
            x = *reinterpret_cast<const ulonglong2 *>(dig + (size_t)J * n);
        ulonglong2 w0 = ldg_stream2(k0 + (size_t)J * a.dstride0, key_policy);
        ulonglong2 w1 = ldg_stream2(k1 + (size_t)J * a.dstride1, key_policy);
        mac128(l0x, h0x, x.x, w0.x);
        mac128(l0y, h0y, x.y, w0.y);
        mac128(l1x, h1x, x.x, w1.x);
        mac128(l1y, h1y, x.y, w1.y);
    }
    ulonglong2 r0, r1;
    r0.x = barrett128(l0x, h0x, pd);
    r0.y = barrett128(l0y, h0y, pd);
    r1.x = barrett128(l1x, h1x, pd);
    r1.y = barrett128(l1y, h1y, pd);
    *reinterpret_cast<ulonglong2 *>(a.acc + (size_t)I * n + e) = r0;
    *reinterpret_cast<ulonglong2 *>(a.acc + ((size_t)(a.l + 1) + I) * n + e) = r1;
}

// ============================================================================================
// Level-aware hybrid key switching (tolerance mode; engine.cu: key_switch_hybrid).  At level l the primes above
// the level are idle, so alpha - 1 of them join the special prime as temporary special moduli P_S; the l limbs are
// grouped into dnum = ceil(l / dsize) digits of dsize primes (dsize = alpha - 1, or 1 when alpha = 1).  Extended limb e in [0, l + alpha): prime e for
// e < l + alpha - 1, the special prime for the last one.  All operands below are coefficient-form residues that were
// pre-multiplied by (Q_d / q_i)^-1 mod q_i on their way out of the inverse NTT, so a basis conversion is the plain
// inner product  x mod p = sum_i y_i * ((Q_d / q_i) mod p)  with a 128-bit accumulator and one Barrett reduction.
// ============================================================================================
struct HybDims
{
    int l, alpha, dsize, dnum, special_prime; // alpha special moduli, digits of dsize primes (dsize < alpha keeps the
                                              // key-switching noise a full prime below the digit size)
    __device__ __forceinline__ int ne() const { return l + alpha; }
    __device__ __forceinline__ int eprime(int e) const { return e < l + alpha - 1 ? e : special_prime; }
    __device__ __forceinline__ bool own(int e, int d) const { return e < l && e / dsize == d; }
};

// canonical INTT output times a per-limb constant (Shoup), canonical
struct StInvScaled
{
    u64 *dst;
    size_t n;
    const ulonglong2 *scale; // [jobs] {c, shoup(c)}
    const int *primes;       // [jobs] prime index of each job
    __device__ __forceinline__ int prime(int job) const { return primes[job]; }
    __device__ __forceinline__ void store(int job, int idx, u64 v, const PrimeDev &pd) const
    {
        v = csub(v, pd.q);
        ulonglong2 f = scale[job];
        dst[(size_t)job * n + idx] = csub(mul_shoup_lazy(v, f.x, f.y, pd.q), pd.q);
    }
};

// the same after adding a per-job constant (floor(D/2) of a rounding division by D, HybridPlan::d_shalf)
struct StInvScaledAdd
{
    u64 *dst;
    size_t n;
    const ulonglong2 *scale; // [jobs] {c, shoup(c)}
    const int *primes;       // [jobs] prime index of each job
    const u64 *add;          // [jobs]
    __device__ __forceinline__ int prime(int job) const { return primes[job]; }
    __device__ __forceinline__ void store(int job, int idx, u64 v, const PrimeDev &pd) const
    {
        v = addmod(csub(v, pd.q), add[job], pd.q);
        ulonglong2 f = scale[job];
        dst[(size_t)job * n + idx] = csub(mul_shoup_lazy(v, f.x, f.y, pd.q), pd.q);
    }
};

// NTT block-pass store of the extended digits: job = eloc * dnum + d (extended limb e0 + eloc, digit d); the digit's own
// limbs are skipped (their NTT form is the input itself)
struct StHybDigit
{
    static constexpr bool RAW = true;
    static constexpr bool BATCH = false;
    u64 *dst;
    size_t n;
    HybDims h;
    int e0;
    __device__ __forceinline__ bool skip(int job) const { return h.own(e0 + job / h.dnum, job % h.dnum); }
    __device__ __forceinline__ int prime(int job) const { return h.eprime(e0 + job / h.dnum); }
    __device__ __forceinline__ u64 pre(int, int, int, int, u64 v, const PrimeDev &) const { return v; }
    __device__ __forceinline__ void post(int job, int idx, u64 v, const PrimeDev &) const
    {
        dst[(size_t)job * n + idx] = v;
    }
};

// the alpha special limbs of both accumulator polynomials: job = p * alpha + a -> acc[p][l + a]
struct LdInvSpecials
{
    static constexpr bool TLAYOUT = false;
    const u64 *acc; // [2][ne][N]
    size_t n;
    HybDims h;
    __device__ __forceinline__ int prime(int job) const { return h.eprime(h.l + job % h.alpha); }
    __device__ __forceinline__ u64 load(int job, int idx, const PrimeDev &) const
    {
        return acc[((size_t)(job / h.alpha) * h.ne() + h.l + job % h.alpha) * n + idx];
    }
    __device__ __forceinline__ u64 load_t(int, int, int) const { return 0; }
};

// the dropped limbs of a merged ModDown + rescale: job = p * (alpha + 1) + a; a = 0 is limb l-1 of the data basis, where
// the base ciphertext joins the accumulator times P_S (P_S base is 0 on the special limbs), a >= 1 special limb a - 1
struct LdInvDropped
{
    static constexpr bool TLAYOUT = false;
    const u64 *acc;   // [2][ne][N]
    const u64 *base0; // [l][N]
    const u64 *base1; // [l][N] or null
    size_t n;
    HybDims h;
    ulonglong2 pmod; // {P_S mod q_{l-1}, shoup}
    __device__ __forceinline__ int prime(int job) const
    {
        const int a = job % (h.alpha + 1);
        return a == 0 ? h.l - 1 : h.eprime(h.l + a - 1);
    }
    __device__ __forceinline__ u64 load(int job, int idx, const PrimeDev &pd) const
    {
        const int p = job / (h.alpha + 1), a = job % (h.alpha + 1);
        if (a != 0)
            return acc[((size_t)p * h.ne() + h.l + a - 1) * n + idx];
        const u64 x = acc[((size_t)p * h.ne() + h.l - 1) * n + idx];
        const u64 *bp = p == 0 ? base0 : base1;
        if (!bp)
            return x;
        const u64 b = bp[(size_t)(h.l - 1) * n + idx];
        return addmod(x, csub(mul_shoup_lazy(b, pmod.x, pmod.y, pd.q), pd.q), pd.q);
    }
    __device__ __forceinline__ u64 load_t(int, int, int) const { return 0; }
};

// Basis conversion as its own kernel (the NTT column pass then loads plain words and keeps its register budget):
// every thread keeps the DS source residues of two coefficients in registers and produces all targets of the chunk.
//   digits  (down == 0): source group dd = digit d, targets t = extended limbs e0 .. e0 + nT - 1, out job = t * dnum + d
//   ModDown (down == 1): source group dd = polynomial p (alpha special limbs), targets t = limbs 0 .. l - 1, job = p * l + t
struct HybConvArgs
{
    const u64 *src; // digits: y [l][N]; ModDown: t [2][alpha][N]
    const u64 *w;   // digits: [ne][dnum][dsize]; ModDown: [l][alpha]
    u64 *out;
    size_t n;
    HybDims h;
    int e0, nT, down;
    // ModDown only (k_hyb_conv<DS, true>): the conversion is made exact and the division rounds to nearest
    const double *pinv; // [alpha] 1 / p_a
    const u64 *negd;    // [l] -D mod q_i, D = product of the source moduli
    const u64 *addc;    // [l] -floor(D/2) mod q_i (the source residues carry +floor(D/2), StInvScaledAdd)
};

constexpr int HYB_CONV_TARGETS = 8;
template <int DS, bool DOWN>
static __global__ void __launch_bounds__(128) k_hyb_conv(HybConvArgs a, NttTables T)
{
    pdl_prologue();
    const size_t n = a.n;
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (i >= n)
        return;
    const int dd = blockIdx.y;
    const int first = DOWN ? dd * a.h.alpha : dd * a.h.dsize;
    const int cnt = DOWN ? a.h.alpha : min(a.h.dsize, a.h.l - first);
    ulonglong2 y[DS];
#pragma unroll
    for (int k = 0; k < DS; k++)
        y[k] = k < cnt ? *reinterpret_cast<const ulonglong2 *>(a.src + (size_t)(first + k) * n + i) : make_ulonglong2(0, 0);
    // ModDown: sum_a y_a (D / p_a) = [x]_D + u D with u = floor(sum_a y_a / p_a) < alpha; u from the fractional parts
    // in double precision (off by one only if the sum is within ~alpha 2^-52 of an integer: a one-unit error of the
    // quotient with probability ~2^-48 per coefficient), u D is taken off again below
    u64 ux = 0, uy = 0;
    if (DOWN)
    {
        double fx = 0.0, fy = 0.0;
#pragma unroll
        for (int k = 0; k < DS; k++)
            if (k < cnt)
            {
                const double pi = a.pinv[k];
                fx = fma((double)y[k].x, pi, fx);
                fy = fma((double)y[k].y, pi, fy);
            }
        ux = (u64)fx;
        uy = (u64)fy;
    }
    // blockIdx.z selects a slice of HYB_CONV_TARGETS targets, so that small digit counts still fill the machine
    const int t_begin = blockIdx.z * HYB_CONV_TARGETS, t_end = min(a.nT, t_begin + HYB_CONV_TARGETS);
    for (int t = t_begin; t < t_end; t++)
    {
        const int e = a.e0 + t;
        if (!DOWN && a.h.own(e, dd))
            continue;
        const PrimeDev pd = T.primes[DOWN ? e : a.h.eprime(e)];
        const u64 *wv = DOWN ? a.w + (size_t)e * a.h.alpha : a.w + ((size_t)e * a.h.dnum + dd) * a.h.dsize;
        u64 lx = 0, hx = 0, ly = 0, hy = 0;
        if (DOWN)
        {
            // - u D - floor(D/2): the converted value is [x + floor(D/2)]_D - floor(D/2) exactly (mod q_e)
            const u64 nd = a.negd[e], ac = a.addc[e];
            lx = ly = ac;
            mac128(lx, hx, ux, nd);
            mac128(ly, hy, uy, nd);
        }
#pragma unroll
        for (int k = 0; k < DS; k++)
        {
            const u64 wk = k < cnt ? wv[k] : 0;
            mac128(lx, hx, y[k].x, wk);
            mac128(ly, hy, y[k].y, wk);
        }
        // the output feeds a forward NTT: where that transform runs the unreduced butterflies (wide modulus) it takes
        // any representative below 6q (6q + 16 stages x 4q < 2^64 for q < 2^57) and the three conditional
        // subtractions of the canonical reduction are skipped; elsewhere the value is brought to [0, q)
        ulonglong2 r;
        if (T.wide && wide_modulus(pd.q))
        {
            r.x = barrett128_lazy6(lx, hx, pd);
            r.y = barrett128_lazy6(ly, hy, pd);
        }
        else
        {
            r.x = barrett128(lx, hx, pd);
            r.y = barrett128(ly, hy, pd);
        }
        const size_t job = DOWN ? (size_t)dd * a.h.l + t : (size_t)t * a.h.dnum + dd;
        *reinterpret_cast<ulonglong2 *>(a.out + job * n + i) = r;
    }
}

// column-pass loader over the converted buffer; own-limb jobs are skipped as in StHybDigit
struct LdHybPlain
{
    const u64 *src; // [nE][dnum][N]
    size_t n;
    HybDims h;
    int e0;
    __device__ __forceinline__ bool skip(int job) const { return h.own(e0 + job / h.dnum, job % h.dnum); }
    __device__ __forceinline__ int prime(int job) const { return h.eprime(e0 + job / h.dnum); }
    __device__ __forceinline__ u64 load(int job, int idx, const PrimeDev &) const { return src[(size_t)job * n + idx]; }
};

constexpr int HYB_MAC_BATCH = 32;
struct HybMacArgs
{
    const u64 *digits;     // [nE][dnum][N] NTT form, lazy; own-limb slots unused
    const u64 *target_ntt; // [l][N] NTT form
    const uint32_t *perm[HYB_MAC_BATCH]; // per rotation of a hoisted group (blockIdx.z): Galois table or null
    const u64 *key[HYB_MAC_BATCH];       // per rotation: polynomial 0 of digit d, limb e at key[rot] + d * dstride0 + e * N
    const u64 *key1[HYB_MAC_BATCH];      // polynomial 1 (the uniform half) at key1[rot] + d * dstride1 + e * N: inside the
                                         // same [dnum][2][ne][N] array, or a buffer expanded from its seed just before
                                         // (seed-compressed keys, keygen.cu)
    size_t dstride0, dstride1;
    u64 *acc;              // [rotations][2][ne][N]
    size_t n;
    HybDims h;
    int e0;
    int gather_digits;
};

// acc_p[e] = sum_d ext_d[e] * key[d][p][e]; grid = (ceil(N / 2 / threads), nE, rotations): all rotations of a hoisted
// group in one launch - the digits of a chunk are shared by them and stay in L2, and one launch of thousands of CTAs
// fills the machine where one launch per rotation left a ramp and a tail every 30 microseconds.
static __global__ void __launch_bounds__(256) k_ks_mac_hyb(HybMacArgs a, NttTables T)
{
    pdl_prologue();
    const int eloc = blockIdx.y;
    const int e = a.e0 + eloc;
    const size_t n = a.n;
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (i >= n)
        return;
    const PrimeDev pd = T.primes[a.h.eprime(e)];
    const int ne = a.h.ne();
    const int rot = blockIdx.z;
    const uint32_t *perm = a.perm[rot];
    u64 *acc = a.acc + (size_t)rot * 2 * ne * n;
    const u64 *k0 = a.key[rot] + (size_t)e * n + i;
    const u64 *k1 = a.key1[rot] + (size_t)e * n + i;
    u64 l0x = 0, h0x = 0, l0y = 0, h0y = 0, l1x = 0, h1x = 0, l1y = 0, h1y = 0;
    const u64 key_policy = l2_evict_first_policy();
    // gather indices are the same for every digit
    size_t gx = i, gy = i + 1;
    if (perm)
    {
        gx = perm[i];
        gy = perm[i + 1];
    }
    auto load_digit = [&](int d) -> ulonglong2 {
        ulonglong2 x;
        if (a.h.own(e, d))
        {
            const u64 *src = a.target_ntt + (size_t)e * n;
            if (perm)
            {
                x.x = src[gx];
                x.y = src[gy];
            }
            else
                x = *reinterpret_cast<const ulonglong2 *>(src + i);
        }
        else
        {
            const u64 *src = a.digits + ((size_t)eloc * a.h.dnum + d) * n;
            if (a.gather_digits)
            {
                x.x = src[gx];
                x.y = src[gy];
            }
            else
                x = *reinterpret_cast<const ulonglong2 *>(src + i);
        }
        return x;
    };
    // two digits per iteration: six independent loads in flight per thread before the first multiply
    int d = 0;
    for (; d + 1 < a.h.dnum; d += 2)
    {
        ulonglong2 xa = load_digit(d), xb = load_digit(d + 1);
        ulonglong2 wa0 = ldg_stream2(k0 + (size_t)d * a.dstride0, key_policy);
        ulonglong2 wa1 = ldg_stream2(k1 + (size_t)d * a.dstride1, key_policy);
        ulonglong2 wb0 = ldg_stream2(k0 + (size_t)(d + 1) * a.dstride0, key_policy);
        ulonglong2 wb1 = ldg_stream2(k1 + (size_t)(d + 1) * a.dstride1, key_policy);
        mac128(l0x, h0x, xa.x, wa0.x);
        mac128(l0y, h0y, xa.y, wa0.y);
        mac128(l1x, h1x, xa.x, wa1.x);
        mac128(l1y, h1y, xa.y, wa1.y);
        mac128(l0x, h0x, xb.x, wb0.x);
        mac128(l0y, h0y, xb.y, wb0.y);
        mac128(l1x, h1x, xb.x, wb1.x);
        mac128(l1y, h1y, xb.y, wb1.y);
    }
    if (d < a.h.dnum)
    {
        ulonglong2 x = load_digit(d);
        ulonglong2 w0 = ldg_stream2(k0 + (size_t)d * a.dstride0, key_policy);
        ulonglong2 w1 = ldg_stream2(k1 + (size_t)d * a.dstride1, key_policy);
        mac128(l0x, h0x, x.x, w0.x);
        mac128(l0y, h0y, x.y, w0.y);
        mac128(l1x, h1x, x.x, w1.x);
        mac128(l1y, h1y, x.y, w1.y);
    }
    ulonglong2 r0, r1;
    r0.x = barrett128(l0x, h0x, pd);
    r0.y = barrett128(l0y, h0y, pd);
    r1.x = barrett128(l1x, h1x, pd);
    r1.y = barrett128(l1y, h1y, pd);
    *reinterpret_cast<ulonglong2 *>(acc + (size_t)e * n + i) = r0;
    *reinterpret_cast<ulonglong2 *>(acc + ((size_t)ne + e) * n + i) = r1;
}

// ---- the same inner product with every operand streamed by the bulk-copy engine (TMA 1-D: cp.async.bulk) ---------
// The key is the HBM stream of a key switch (2 dnum ne limb-polynomials, read once); the threads of the kernel above
// keep at most two digits' worth of it in flight in registers and fetch the digits of a hoisted rotation through the
// Galois table with scattered 8-byte loads.  Here one elected thread hands whole tiles to the copy engine - a ring of
// KS_BULK_STAGES digits, per digit 4 KB of each key polynomial and the two 2 KB blocks of the digit, completion
// signalled on one mbarrier per stage with complete_tx byte counting, L2 evict-first policy on the key - while all
// threads multiply the stage that has landed.
// The digit needs no scattered global loads: in NTT (bit-reversed) order a Galois automorphism maps every aligned
// block of 256 coefficients ONTO one aligned block of 256 (index i <-> exponent 2 brev(i) + 1; multiplying the exponent
// by the Galois element acts on its low bits, i.e. on the HIGH bits of i, independently of the rest).  So the source
// of a 256-coefficient half tile is one contiguous 2 KB block, copied in bulk; the permutation inside the block is
// applied when the threads read shared memory.
// grid = (N / 512, nE, rotations), block = 128: thread t owns coefficients 2t, 2t+1 and 256 + 2t, 256 + 2t + 1 of the
// tile (two conflict-free 16-byte shared-memory reads per key polynomial).
constexpr int KS_BULK_TILE = 512;
constexpr int KS_BULK_STAGES = 3;

__device__ __forceinline__ unsigned smem_u32(const void *p)
{
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(u64 *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, u64 *bar, u64 policy)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s_plain(void *dst, const void *src, unsigned bytes, u64 *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// bounded: a barrier that never completes traps (kernel error) instead of hanging the device
__device__ __forceinline__ void mbar_wait(u64 *bar, unsigned parity)
{
    for (unsigned spin = 0; spin < (1u << 26); spin++)
    {
        unsigned ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
        if (ok)
            return;
    }
    __trap();
}

static __global__ void __launch_bounds__(128) k_ks_mac_hyb_bulk(HybMacArgs a, NttTables T)
{
    pdl_prologue();
    __shared__ __align__(128) u64 sk[KS_BULK_STAGES][2][KS_BULK_TILE];
    __shared__ __align__(128) u64 sd[KS_BULK_STAGES][KS_BULK_TILE];
    __shared__ __align__(8) u64 full[KS_BULK_STAGES];
    const int eloc = blockIdx.y;
    const int e = a.e0 + eloc;
    const int rot = blockIdx.z;
    const size_t n = a.n;
    const size_t base = (size_t)blockIdx.x * KS_BULK_TILE;
    const int dnum = a.h.dnum, ne = a.h.ne();
    const PrimeDev pd = T.primes[a.h.eprime(e)];
    const u64 *k0 = a.key[rot] + (size_t)e * n + base;
    const u64 *k1 = a.key1[rot] + (size_t)e * n + base;
    const uint32_t *perm = a.perm[rot];
    u64 *acc = a.acc + (size_t)rot * 2 * ne * n;
    const unsigned t = threadIdx.x;
    constexpr unsigned TILE_BYTES = KS_BULK_TILE * (unsigned)sizeof(u64), HALF_BYTES = TILE_BYTES / 2;

    if (t == 0)
    {
        for (int s = 0; s < KS_BULK_STAGES; s++)
            mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const u64 policy = l2_evict_first_policy();
    // source blocks of the two half tiles under this rotation's automorphism, and of the unpermuted tile
    const size_t blk_a = perm ? ((size_t)perm[base] & ~(size_t)255) : base;
    const size_t blk_b = perm ? ((size_t)perm[base + 256] & ~(size_t)255) : base + 256;
    auto issue = [&](int d) {
        const int s = d % KS_BULK_STAGES;
        const bool own = a.h.own(e, d);
        const u64 *src = own ? a.target_ntt + (size_t)e * n : a.digits + ((size_t)eloc * dnum + d) * n;
        const bool gather = own ? perm != nullptr : a.gather_digits != 0;
        mbar_expect_tx(&full[s], 3 * TILE_BYTES);
        bulk_g2s(&sk[s][0][0], k0 + (size_t)d * a.dstride0, TILE_BYTES, &full[s], policy);
        bulk_g2s(&sk[s][1][0], k1 + (size_t)d * a.dstride1, TILE_BYTES, &full[s], policy);
        bulk_g2s_plain(&sd[s][0], src + (gather ? blk_a : base), HALF_BYTES, &full[s]);
        bulk_g2s_plain(&sd[s][256], src + (gather ? blk_b : base + 256), HALF_BYTES, &full[s]);
    };
    if (t == 0)
        for (int d = 0; d < dnum && d < KS_BULK_STAGES; d++)
            issue(d);

    // coefficients of this thread: two pairs, 256 apart; where they sit inside the source blocks
    const size_t ia = base + 2 * t, ib = base + 256 + 2 * t;
    unsigned pa0 = 2 * t, pa1 = 2 * t + 1, pb0 = 256 + 2 * t, pb1 = 256 + 2 * t + 1;
    unsigned qa0 = pa0, qa1 = pa1, qb0 = pb0, qb1 = pb1;
    if (perm)
    {
        qa0 = perm[ia] & 255u;
        qa1 = perm[ia + 1] & 255u;
        qb0 = 256u + (perm[ib] & 255u);
        qb1 = 256u + (perm[ib + 1] & 255u);
    }
    u64 lo[2][4] = {}, hi[2][4] = {}; // [key polynomial][coefficient]
    for (int d = 0; d < dnum; d++)
    {
        const int s = d % KS_BULK_STAGES;
        const bool own = a.h.own(e, d);
        const bool gather = own ? perm != nullptr : a.gather_digits != 0;
        mbar_wait(&full[s], (unsigned)(d / KS_BULK_STAGES) & 1u);
        u64 x[4];
        x[0] = sd[s][gather ? qa0 : pa0];
        x[1] = sd[s][gather ? qa1 : pa1];
        x[2] = sd[s][gather ? qb0 : pb0];
        x[3] = sd[s][gather ? qb1 : pb1];
#pragma unroll
        for (int p = 0; p < 2; p++)
        {
            const ulonglong2 wa = *reinterpret_cast<const ulonglong2 *>(&sk[s][p][2 * t]);
            const ulonglong2 wb = *reinterpret_cast<const ulonglong2 *>(&sk[s][p][256 + 2 * t]);
            mac128(lo[p][0], hi[p][0], x[0], wa.x);
            mac128(lo[p][1], hi[p][1], x[1], wa.y);
            mac128(lo[p][2], hi[p][2], x[2], wb.x);
            mac128(lo[p][3], hi[p][3], x[3], wb.y);
        }
        __syncthreads(); // every thread is done with stage s
        if (t == 0 && d + KS_BULK_STAGES < dnum)
        {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic-proxy reads before the async-proxy refill
            issue(d + KS_BULK_STAGES);
        }
    }
#pragma unroll
    for (int p = 0; p < 2; p++)
    {
        ulonglong2 ra, rb;
        ra.x = barrett128(lo[p][0], hi[p][0], pd);
        ra.y = barrett128(lo[p][1], hi[p][1], pd);
        rb.x = barrett128(lo[p][2], hi[p][2], pd);
        rb.y = barrett128(lo[p][3], hi[p][3], pd);
        u64 *dst = acc + ((size_t)p * ne + e) * n;
        *reinterpret_cast<ulonglong2 *>(dst + ia) = ra;
        *reinterpret_cast<ulonglong2 *>(dst + ib) = rb;
    }
}

// ============================================================================================
// Element-wise limb kernels.  Data = [polys][limbs][N]; grid-stride over polys*limbs*N with the
// prime taken from the limb index.  2 coefficients (16 bytes) per thread per step.
// ============================================================================================
enum EwOp
{
    EW_ADD = 0,
    EW_SUB = 1,
    EW_NEG = 2,
    EW_MUL = 3,       // a * b (b broadcast over polys if b_polys == 1: plaintext operand)
    EW_ADD_P0 = 4,    // a[0] += b (plaintext add touches poly 0 only)
    EW_SUB_P0 = 5,
    EW_MULADD = 6     // a += b * c
};

template <int OP>
__global__ void __launch_bounds__(256) k_ew(u64 *__restrict__ a, const u64 *__restrict__ b, const PrimeDev *primes,
                                            int log_n, int limbs, int a_polys, int b_polys)
{
    pdl_prologue();
    const size_t n = size_t(1) << log_n;
    const size_t per_poly = (size_t)limbs * n;
    const size_t total = (size_t)a_polys * per_poly / 2;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
    {
        size_t e = i * 2;
        int limb = (int)((e % per_poly) >> log_n);
        const PrimeDev pd = primes[limb];
        ulonglong2 va = *reinterpret_cast<ulonglong2 *>(a + e);
        ulonglong2 vb = make_ulonglong2(0, 0);
        if (OP != EW_NEG)
        {
            size_t eb = b_polys == 1 ? e % per_poly : e;
            vb = *reinterpret_cast<const ulonglong2 *>(b + eb);
        }
        if (OP == EW_ADD || OP == EW_ADD_P0)
        {
            va.x = addmod(va.x, vb.x, pd.q);
            va.y = addmod(va.y, vb.y, pd.q);
        }
        else if (OP == EW_SUB || OP == EW_SUB_P0)
        {
            va.x = submod(va.x, vb.x, pd.q);
            va.y = submod(va.y, vb.y, pd.q);
        }
        else if (OP == EW_NEG)
        {
            va.x = va.x ? pd.q - va.x : 0ull;
            va.y = va.y ? pd.q - va.y : 0ull;
        }
        else if (OP == EW_MUL)
        {
            va.x = mulmod(va.x, vb.x, pd);
            va.y = mulmod(va.y, vb.y, pd);
        }
        *reinterpret_cast<ulonglong2 *>(a + e) = va;
    }
}

// acc[p][l][i] (= or +=) a[p][l][i] * pt[l][i]: the "multiply by a plaintext diagonal and accumulate" step of the
// BSGS linear transforms and of the convolution taps in one pass over the data (dyadic_product_coeffmod +
// add_poly_coeffmod, polyarithsmallmod.cpp:111-169,18-43).
template <bool FIRST>
__global__ void __launch_bounds__(256) k_mul_plain_acc(u64 *__restrict__ acc, const u64 *__restrict__ a,
                                                       const u64 *__restrict__ pt, const PrimeDev *primes, int log_n,
                                                       int limbs, int polys)
{
    pdl_prologue();
    const size_t n = size_t(1) << log_n;
    const size_t per_poly = (size_t)limbs * n;
    const size_t total = (size_t)polys * per_poly / 2;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
    {
        size_t e = i * 2, ep = e % per_poly;
        const PrimeDev pd = primes[(int)(ep >> log_n)];
        ulonglong2 va = *reinterpret_cast<const ulonglong2 *>(a + e);
        ulonglong2 vp = *reinterpret_cast<const ulonglong2 *>(pt + ep);
        ulonglong2 r;
        r.x = mulmod(va.x, vp.x, pd);
        r.y = mulmod(va.y, vp.y, pd);
        if (!FIRST)
        {
            ulonglong2 vc = *reinterpret_cast<const ulonglong2 *>(acc + e);
            r.x = addmod(vc.x, r.x, pd.q);
            r.y = addmod(vc.y, r.y, pd.q);
        }
        *reinterpret_cast<ulonglong2 *>(acc + e) = r;
    }
}

// ct x ct tensor product (ckks_multiply, evaluator.cpp:744-772): (a0,a1) x (b0,b1) ->
// (a0 b0, a0 b1 + a1 b0, a1 b1).  out may alias neither input.
// dst = sum_t a_t (.) pt_t over up to MUL_SUM_TERMS terms in one pass (the inner sum of a BSGS group / the taps of a
// convolution): every operand is read once, the products are accumulated in 128 bits and reduced once - the residues
// equal those of the term-by-term multiply_plain + add sequence.  ACCUMULATE adds the previous content of dst.
constexpr int MUL_SUM_TERMS = 16;
struct MulSumArgs
{
    const u64 *ct[MUL_SUM_TERMS];
    const u64 *pt[MUL_SUM_TERMS];
    int count;
};
// special_pos >= 0: limb `special_pos` of every polynomial belongs to prime `special_prime` (operands in the extended
// basis of a key switch, whose last limb is the special prime).
template <bool ACCUMULATE>
__global__ void __launch_bounds__(256) k_mul_plain_sum(u64 *__restrict__ dst, MulSumArgs a, const PrimeDev *primes, int log_n,
                                                       int limbs, int polys, int special_pos = -1, int special_prime = 0)
{
    pdl_prologue();
    const size_t n = size_t(1) << log_n;
    const size_t per_poly = (size_t)limbs * n;
    const size_t total = (size_t)polys * per_poly / 2;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
    {
        const size_t e = i * 2, ep = e % per_poly;
        const int limb = (int)(ep >> log_n);
        const PrimeDev pd = primes[limb == special_pos ? special_prime : limb];
        u64 lx = 0, hx = 0, ly = 0, hy = 0;
        if (ACCUMULATE)
        {
            ulonglong2 vc = *reinterpret_cast<const ulonglong2 *>(dst + e);
            lx = vc.x;
            ly = vc.y;
        }
        for (int t = 0; t < a.count; t++)
        {
            ulonglong2 va = *reinterpret_cast<const ulonglong2 *>(a.ct[t] + e);
            ulonglong2 vp = *reinterpret_cast<const ulonglong2 *>(a.pt[t] + ep);
            mac128(lx, hx, va.x, vp.x);
            mac128(ly, hy, va.y, vp.y);
        }
        ulonglong2 r;
        r.x = barrett128(lx, hx, pd);
        r.y = barrett128(ly, hy, pd);
        *reinterpret_cast<ulonglong2 *>(dst + e) = r;
    }
}

// The same for up to MUL_SUM_GROUPS sums over the SAME ciphertext operands with different plaintexts - the giant steps
// of a double-hoisted BSGS transform, which all multiply the same rotated (extended-basis) ciphertexts: every
// ciphertext word is read once for the whole group instead of once per giant step (k babies, G giants: (2 + G) k + 2 G
// limb-polynomials per extended limb instead of G (3 k + 2)).  dst of sum g at dst + g * dst_stride; a null plaintext
// means "no such term in that sum".
constexpr int MUL_SUM_GROUPS = 4;
struct MulSumMultiArgs
{
    const u64 *ct[MUL_SUM_TERMS];
    const u64 *pt[MUL_SUM_GROUPS][MUL_SUM_TERMS];
    int count;
};
template <bool ACCUMULATE, int G>
__global__ void __launch_bounds__(256) k_mul_plain_sum_multi(u64 *__restrict__ dst, size_t dst_stride, const __grid_constant__ MulSumMultiArgs a,
                                                             const PrimeDev *primes, int log_n, int limbs, int polys, int special_pos,
                                                             int special_prime)
{
    pdl_prologue();
    const size_t n = size_t(1) << log_n;
    const size_t per_poly = (size_t)limbs * n;
    const size_t total = (size_t)polys * per_poly / 2;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
    {
        const size_t e = i * 2, ep = e % per_poly;
        const int limb = (int)(ep >> log_n);
        const PrimeDev pd = primes[limb == special_pos ? special_prime : limb];
        u64 lx[G], hx[G], ly[G], hy[G];
#pragma unroll
        for (int g = 0; g < G; g++)
        {
            lx[g] = hx[g] = ly[g] = hy[g] = 0;
            if (ACCUMULATE)
            {
                ulonglong2 vc = *reinterpret_cast<const ulonglong2 *>(dst + (size_t)g * dst_stride + e);
                lx[g] = vc.x;
                ly[g] = vc.y;
            }
        }
        for (int t = 0; t < a.count; t++)
        {
            const ulonglong2 va = __ldcs(reinterpret_cast<const ulonglong2 *>(a.ct[t] + e));
#pragma unroll
            for (int g = 0; g < G; g++)
            {
                const u64 *p = a.pt[g][t];
                if (p)
                {
                    const ulonglong2 vp = *reinterpret_cast<const ulonglong2 *>(p + ep);
                    mac128(lx[g], hx[g], va.x, vp.x);
                    mac128(ly[g], hy[g], va.y, vp.y);
                }
            }
        }
#pragma unroll
        for (int g = 0; g < G; g++)
        {
            ulonglong2 r;
            r.x = barrett128(lx[g], hx[g], pd);
            r.y = barrett128(ly[g], hy[g], pd);
            *reinterpret_cast<ulonglong2 *>(dst + (size_t)g * dst_stride + e) = r;
        }
    }
}

// dst[limbs][N] (+)= sum_t src[.][perm_t[x]] * pt_t[.][x]: the part of a double-hoisted BSGS inner sum that needs no key
// switch (the c0 halves of the rotated ciphertexts are plain permutations of the input's c0).  perm_t == null: identity.
constexpr int GATHER_SUM_TERMS = 16;
struct GatherSumArgs
{
    const uint32_t *perm[GATHER_SUM_TERMS];
    const u64 *pt[GATHER_SUM_TERMS];
    int count;
};
template <bool ACCUMULATE>
__global__ void __launch_bounds__(256) k_gather_mul_sum(u64 *__restrict__ dst, const u64 *__restrict__ src,
                                                        const __grid_constant__ GatherSumArgs a, const PrimeDev *primes,
                                                        int log_n, int limbs)
{
    pdl_prologue();
    const size_t n = size_t(1) << log_n;
    const size_t total = (size_t)limbs * n / 2; // two coefficients per step
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
    {
        const size_t e = i * 2;
        const int limb = (int)(e >> log_n);
        const size_t x = e & (n - 1);
        const PrimeDev pd = primes[limb];
        const u64 *row = src + (size_t)limb * n;
        u64 lo0 = 0, hi0 = 0, lo1 = 0, hi1 = 0;
        if (ACCUMULATE)
        {
            const ulonglong2 d = *reinterpret_cast<const ulonglong2 *>(dst + e);
            lo0 = d.x;
            lo1 = d.y;
        }
        // four terms at a time: their Galois indices and plaintext words are requested together, then the gathers
        // (the input's c0 limb stays in L2), then the multiplies - one DRAM latency per four terms instead of two per term
        for (int t0 = 0; t0 < a.count; t0 += 4)
        {
            uint2 sx[4];
            ulonglong2 pw[4];
#pragma unroll
            for (int u = 0; u < 4; u++)
            {
                const int t = min(t0 + u, a.count - 1);
                sx[u] = a.perm[t] ? __ldg(reinterpret_cast<const uint2 *>(a.perm[t] + x)) : make_uint2((unsigned)x, (unsigned)x + 1);
                pw[u] = __ldcs(reinterpret_cast<const ulonglong2 *>(a.pt[t] + e));
            }
            u64 v0[4], v1[4];
#pragma unroll
            for (int u = 0; u < 4; u++)
            {
                v0[u] = __ldg(row + sx[u].x);
                v1[u] = __ldg(row + sx[u].y);
            }
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (t0 + u < a.count)
                {
                    mac128(lo0, hi0, v0[u], pw[u].x);
                    mac128(lo1, hi1, v1[u], pw[u].y);
                }
        }
        ulonglong2 o;
        o.x = barrett128(lo0, hi0, pd);
        o.y = barrett128(lo1, hi1, pd);
        *reinterpret_cast<ulonglong2 *>(dst + e) = o;
    }
}

static __global__ void __launch_bounds__(256) k_tensor(const u64 *__restrict__ a, const u64 *__restrict__ b,
                                                u64 *__restrict__ out, const PrimeDev *primes, int log_n, int limbs)
{
    pdl_prologue();
    const size_t n = size_t(1) << log_n;
    const size_t per_poly = (size_t)limbs * n;
    const size_t total = per_poly / 2;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
    {
        size_t e = i * 2;
        const PrimeDev pd = primes[(int)(e >> log_n)];
        ulonglong2 a0 = *reinterpret_cast<const ulonglong2 *>(a + e);
        ulonglong2 a1 = *reinterpret_cast<const ulonglong2 *>(a + per_poly + e);
        ulonglong2 b0 = *reinterpret_cast<const ulonglong2 *>(b + e);
        ulonglong2 b1 = *reinterpret_cast<const ulonglong2 *>(b + per_poly + e);
        ulonglong2 d0, d1, d2;
        d0.x = mulmod(a0.x, b0.x, pd);
        d0.y = mulmod(a0.y, b0.y, pd);
        d2.x = mulmod(a1.x, b1.x, pd);
        d2.y = mulmod(a1.y, b1.y, pd);
        {
            u64 lo = 0, hi = 0;
            mac128(lo, hi, a0.x, b1.x);
            mac128(lo, hi, a1.x, b0.x);
            d1.x = barrett128(lo, hi, pd);
            lo = hi = 0;
            mac128(lo, hi, a0.y, b1.y);
            mac128(lo, hi, a1.y, b0.y);
            d1.y = barrett128(lo, hi, pd);
        }
        *reinterpret_cast<ulonglong2 *>(out + e) = d0;
        *reinterpret_cast<ulonglong2 *>(out + per_poly + e) = d1;
        *reinterpret_cast<ulonglong2 *>(out + 2 * per_poly + e) = d2;
    }
}

// ckks_square (evaluator.cpp:1000-1059): (a0,a1) -> (a0^2, 2 a0 a1, a1^2)
static __global__ void __launch_bounds__(256) k_square(const u64 *__restrict__ a, u64 *__restrict__ out,
                                                const PrimeDev *primes, int log_n, int limbs)
{
    pdl_prologue();
    const size_t n = size_t(1) << log_n;
    const size_t per_poly = (size_t)limbs * n;
    const size_t total = per_poly / 2;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
    {
        size_t e = i * 2;
        const PrimeDev pd = primes[(int)(e >> log_n)];
        ulonglong2 a0 = *reinterpret_cast<const ulonglong2 *>(a + e);
        ulonglong2 a1 = *reinterpret_cast<const ulonglong2 *>(a + per_poly + e);
        ulonglong2 d0, d1, d2;
        d0.x = mulmod(a0.x, a0.x, pd);
        d0.y = mulmod(a0.y, a0.y, pd);
        d2.x = mulmod(a1.x, a1.x, pd);
        d2.y = mulmod(a1.y, a1.y, pd);
        u64 m = mulmod(a0.x, a1.x, pd);
        d1.x = addmod(m, m, pd.q);
        m = mulmod(a0.y, a1.y, pd);
        d1.y = addmod(m, m, pd.q);
        *reinterpret_cast<ulonglong2 *>(out + e) = d0;
        *reinterpret_cast<ulonglong2 *>(out + per_poly + e) = d1;
        *reinterpret_cast<ulonglong2 *>(out + 2 * per_poly + e) = d2;
    }
}

// Galois gather on NTT-form limbs (galois.cpp:192-218): dst[j][i] = src[j][perm[i]].
static __global__ void __launch_bounds__(256) k_permute(const u64 *__restrict__ src, u64 *__restrict__ dst,
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

// strided limb copy: dst[p][i][.] = src[p][i][.] for i < limbs_out (mod_switch_drop_to_next,
// evaluator.cpp:1183-1246) - a pure re-pack, no arithmetic.
static __global__ void __launch_bounds__(256) k_drop_limbs(const u64 *__restrict__ src, u64 *__restrict__ dst, int log_n,
                                                    int limbs_in, int limbs_out, int polys)
{
    pdl_prologue();
    const size_t n = size_t(1) << log_n;
    const size_t per_out = (size_t)limbs_out * n;
    const size_t total = (size_t)polys * per_out / 2;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
    {
        size_t e = i * 2;
        size_t p = e / per_out, r = e % per_out;
        *reinterpret_cast<ulonglong2 *>(dst + e) =
            *reinterpret_cast<const ulonglong2 *>(src + p * (size_t)limbs_in * n + r);
    }
}
