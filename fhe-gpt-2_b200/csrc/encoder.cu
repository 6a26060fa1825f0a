// encoder.cu - CKKSEncoder on the device (C ABI: bk_encode / bk_encode_scalar / bk_decode).
//
// Replaces the reference's CKKSEncoder (seal/ckks.cpp:13-75 constructor tables,
// seal/ckks.h:457-638 encode_internal, :644-761 decode_internal, ckks.cpp:77-216 scalar
// encode) and the complex DWT it runs through DWTHandler (util/dwthandler.h:94-191,202-356).
//
// Bit-exactness: the canonical-embedding FFT is FP64.  Every butterfly below performs the same
// IEEE operations in the same order as the reference's scalar loop (complex add/sub, the
// four-multiply complex product re = ac - bd, im = ad + bc, scalar fix-up folded into the last
// inverse stage) using the never-contracted __dadd_rn/__dsub_rn/__dmul_rn intrinsics, and the
// root tables are generated on the host with the same libm calls as util/croots.cpp:17-73, so
// the rounded integer coefficients - and therefore every residue - equal the reference's
// (built without FMA contraction, which is what an x86-64 baseline build of it is).
//
// B200 mapping: the N-point complex transform (N = 2^16 -> 1 MiB of double2) is split into a
// shared-memory pass over contiguous 2048-point blocks (11 stages) and a register radix-2^k
// pass over stride-2048 columns (the remaining <= 5 stages); the real parts are rounded and
// RNS-decomposed inside the load functor of the first NTT pass, so the coefficient vector is
// never materialised per limb.
#include "engine.h"
#include <cmath>
#include <complex>
#include <cstring>

namespace bk
{
    typedef double2 cplx;

    struct EncoderState
    {
        uint32_t *d_index_map = nullptr; // matrix_reps_index_map_ [N]
        cplx *d_roots = nullptr;         // root_powers_ [N]
        cplx *d_inv_roots = nullptr;     // inv_root_powers_ [N]
        // CRT composition tables for decode (per level l: Garner constants)
        u64 *d_garner_inv = nullptr;  // [n_primes] (q_0..q_{j-1})^-1 mod q_j
        u64 *d_prodmod = nullptr;     // [n_primes][n_primes] (q_0..q_{i-1}) mod q_j
        u64 *d_prodwords = nullptr;   // [n_primes][n_primes] multiword q_0..q_{j-1}
        u64 *d_total = nullptr;       // [n_primes+1][n_primes] total modulus per level
        u64 *d_half = nullptr;        // [n_primes+1][n_primes] upper half threshold per level
    };

    static const double PI_ = 3.1415926535897932384626433832795028842;

    // util/croots.cpp:17-73
    struct HostRoots
    {
        size_t degree;
        std::vector<std::complex<double>> roots;
        explicit HostRoots(size_t m) : degree(m), roots(m / 8 + 1)
        {
            for (size_t i = 0; i <= m / 8; i++)
                roots[i] = std::polar<double>(1.0, 2 * PI_ * static_cast<double>(i) / static_cast<double>(m));
        }
        std::complex<double> get(size_t index) const
        {
            index &= degree - 1;
            if (index <= degree / 8)
                return roots[index];
            else if (index <= degree / 4)
            {
                auto a = roots[degree / 4 - index];
                return { a.imag(), a.real() };
            }
            else if (index <= degree / 2)
                return -std::conj(get(degree / 2 - index));
            else if (index <= 3 * degree / 4)
                return -get(index - degree / 2);
            else
                return std::conj(get(degree - index));
        }
    };

    static void mw_mul_word(std::vector<uint64_t> &a, uint64_t w)
    {
        uint64_t carry = 0;
        for (auto &x : a)
        {
            u128 t = (u128)x * w + carry;
            x = (uint64_t)t;
            carry = (uint64_t)(t >> 64);
        }
        // caller sizes `a` so the product fits
    }

    static EncoderState &encoder(Context &c)
    {
        std::lock_guard<std::mutex> g(c.enc_mu);
        if (c.enc)
            return *c.enc;
        c.activate();
        auto st = new EncoderState();
        const size_t n = c.n;
        const size_t slots = n >> 1;
        const int logn = c.log_n;
        const uint64_t m = (uint64_t)n << 1;
        std::vector<uint32_t> map(n);
        uint64_t pos = 1;
        for (size_t i = 0; i < slots; i++)
        {
            uint64_t index1 = (pos - 1) >> 1;
            uint64_t index2 = (m - pos - 1) >> 1;
            map[i] = bitrev((uint32_t)index1, logn);
            map[slots | i] = bitrev((uint32_t)index2, logn);
            pos *= 5;
            pos &= (m - 1);
        }
        HostRoots hr((size_t)m);
        std::vector<cplx> roots(n), iroots(n);
        roots[0] = make_double2(0, 0);
        iroots[0] = make_double2(0, 0);
        for (size_t i = 1; i < n; i++)
        {
            auto r = hr.get(bitrev((uint32_t)i, logn));
            auto ir = std::conj(hr.get((size_t)bitrev((uint32_t)(i - 1), logn) + 1));
            roots[i] = make_double2(r.real(), r.imag());
            iroots[i] = make_double2(ir.real(), ir.imag());
        }
        BK_CUDA(cudaMalloc((void **)&st->d_index_map, n * sizeof(uint32_t)));
        BK_CUDA(cudaMalloc((void **)&st->d_roots, n * sizeof(cplx)));
        BK_CUDA(cudaMalloc((void **)&st->d_inv_roots, n * sizeof(cplx)));
        BK_CUDA(cudaMemcpy(st->d_index_map, map.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice));
        BK_CUDA(cudaMemcpy(st->d_roots, roots.data(), n * sizeof(cplx), cudaMemcpyHostToDevice));
        BK_CUDA(cudaMemcpy(st->d_inv_roots, iroots.data(), n * sizeof(cplx), cudaMemcpyHostToDevice));
        BK_CUDA(cudaDeviceSynchronize());

        // Garner / CRT tables (rns.cpp RNSBase::compose semantics, restated as mixed radix)
        const int P = c.n_primes;
        std::vector<u64> ginv(P, 0), prodmod((size_t)P * P, 0), prodwords((size_t)P * P, 0);
        std::vector<u64> total((size_t)(P + 1) * P, 0), half((size_t)(P + 1) * P, 0);
        std::vector<uint64_t> prod(P, 0);
        prod[0] = 1;
        for (int j = 0; j < P; j++)
        {
            // prod = q_0..q_{j-1}
            for (int w = 0; w < P; w++)
                prodwords[(size_t)j * P + w] = prod[w];
            uint64_t pm = 1 % c.primes[j];
            for (int i = 0; i < j; i++)
            {
                prodmod[(size_t)j * P + i] = pm; // (q_0..q_{i-1}) mod q_j
                pm = mulmod(pm, c.primes[i] % c.primes[j], c.primes[j]);
            }
            ginv[j] = j == 0 ? 1 : invmod(pm, c.primes[j]);
            mw_mul_word(prod, c.primes[j]);
            // level l = j+1: total modulus = prod, threshold = (prod + 1) >> 1
            for (int w = 0; w < P; w++)
                total[(size_t)(j + 1) * P + w] = prod[w];
            std::vector<uint64_t> h(prod);
            // +1 then >>1 (context.cpp:377-383)
            for (int w = 0; w < P; w++)
            {
                if (++h[w] != 0)
                    break;
            }
            for (int w = 0; w < P; w++)
            {
                uint64_t nxt = w + 1 < P ? h[w + 1] : 0;
                h[w] = (h[w] >> 1) | (nxt << 63);
            }
            for (int w = 0; w < P; w++)
                half[(size_t)(j + 1) * P + w] = h[w];
        }
        auto up = [&](u64 **d, const std::vector<u64> &h) {
            BK_CUDA(cudaMalloc((void **)d, h.size() * sizeof(u64)));
            BK_CUDA(cudaMemcpy(*d, h.data(), h.size() * sizeof(u64), cudaMemcpyHostToDevice));
        };
        up(&st->d_garner_inv, ginv);
        up(&st->d_prodmod, prodmod);
        up(&st->d_prodwords, prodwords);
        up(&st->d_total, total);
        up(&st->d_half, half);
        BK_CUDA(cudaDeviceSynchronize());
        c.enc = st;
        return *st;
    }

    void destroy_encoder(Context &c)
    {
        if (!c.enc)
            return;
        cudaFree(c.enc->d_index_map);
        cudaFree(c.enc->d_roots);
        cudaFree(c.enc->d_inv_roots);
        cudaFree(c.enc->d_garner_inv);
        cudaFree(c.enc->d_prodmod);
        cudaFree(c.enc->d_prodwords);
        cudaFree(c.enc->d_total);
        cudaFree(c.enc->d_half);
        delete c.enc;
        c.enc = nullptr;
    }
} // namespace bk

using namespace bk;

// ---- complex arithmetic, operation-for-operation as std::complex<double> without contraction ----
__device__ __forceinline__ cplx c_add(cplx a, cplx b)
{
    return make_double2(__dadd_rn(a.x, b.x), __dadd_rn(a.y, b.y));
}
__device__ __forceinline__ cplx c_sub(cplx a, cplx b)
{
    return make_double2(__dsub_rn(a.x, b.x), __dsub_rn(a.y, b.y));
}
__device__ __forceinline__ cplx c_mul(cplx a, cplx b)
{
    double ac = __dmul_rn(a.x, b.x), bd = __dmul_rn(a.y, b.y);
    double ad = __dmul_rn(a.x, b.y), bc = __dmul_rn(a.y, b.x);
    return make_double2(__dsub_rn(ac, bd), __dadd_rn(ad, bc));
}
__device__ __forceinline__ cplx c_scale(cplx a, double s)
{
    return make_double2(__dmul_rn(a.x, s), __dmul_rn(a.y, s));
}

constexpr int FFT_LB = 11; // shared-memory block = 2048 points
constexpr int FFT_B = 1 << FFT_LB;

// scatter slot values and their conjugates (ckks.h:499-508); vals = n_values complex
__global__ void k_enc_scatter(const cplx *__restrict__ vals, int n_values, const uint32_t *__restrict__ map,
                              cplx *__restrict__ out, int slots)
{
    pdl_prologue();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_values)
        return;
    cplx v = vals[i];
    out[map[i]] = v;
    out[map[i + slots]] = make_double2(v.x, -v.y);
}

// Inverse DWT (Gentleman-Sande), stages with gap 1 .. 1024 inside one 2048-point block.
// Stage with gap g has m = n/(2g) groups; group i uses roots[n - 2m + 1 + i].
__global__ void __launch_bounds__(256) k_ifft_block(cplx *__restrict__ data, const cplx *__restrict__ roots, int log_n)
{
    pdl_prologue();
    __shared__ cplx sm[FFT_B];
    const size_t n = size_t(1) << log_n;
    cplx *base = data + (size_t)blockIdx.x * FFT_B;
    for (int i = threadIdx.x; i < FFT_B; i += 256)
        sm[i] = base[i];
    __syncthreads();
    for (int lg = 0; lg < FFT_LB; lg++)
    {
        const int gap = 1 << lg;
        const size_t m = n >> (lg + 1);
        for (int b = threadIdx.x; b < FFT_B / 2; b += 256)
        {
            int grp = b >> lg, j = b & (gap - 1);
            int xi = (grp << (lg + 1)) + j;
            size_t gi = ((size_t)blockIdx.x * FFT_B + xi) >> (lg + 1);
            cplx r = roots[n - 2 * m + 1 + gi];
            cplx u = sm[xi], v = sm[xi + gap];
            sm[xi] = c_add(u, v);
            sm[xi + gap] = c_mul(c_sub(u, v), r);
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < FFT_B; i += 256)
        base[i] = sm[i];
}

// Remaining LOGS inverse stages over stride-2048 columns; the last one applies `fix`
// (dwthandler.h:273-314).  Writes only the real part (ckks.h:513-517 uses .real()).
template <int LOGS>
__global__ void __launch_bounds__(256) k_ifft_cols(const cplx *__restrict__ data, const cplx *__restrict__ roots,
                                                   double fix, double *__restrict__ re_out,
                                                   unsigned long long *__restrict__ max_bits, int log_n)
{
    pdl_prologue();
    constexpr int S = 1 << LOGS;
    const size_t n = size_t(1) << log_n;
    const int c = blockIdx.x * blockDim.x + threadIdx.x; // column 0..2047
    cplx x[S];
#pragma unroll
    for (int k = 0; k < S; k++)
        x[k] = data[(size_t)k * FFT_B + c];
#pragma unroll
    for (int j = 0; j < LOGS; j++)
    {
        const int half = 1 << j;           // pair distance in k
        const size_t m = (size_t)S >> (j + 1); // groups in this stage = n / (2 * gap)
        const bool last = (j == LOGS - 1);
#pragma unroll
        for (int k = 0; k < S; k++)
        {
            if (!(k & half))
            {
                cplx r = roots[n - 2 * m + 1 + (size_t)(k >> (j + 1))];
                cplx u = x[k], v = x[k + half];
                if (last)
                {
                    cplx sr = c_scale(r, fix);
                    x[k] = c_scale(c_add(u, v), fix);
                    x[k + half] = c_mul(c_sub(u, v), sr);
                }
                else
                {
                    x[k] = c_add(u, v);
                    x[k + half] = c_mul(c_sub(u, v), r);
                }
            }
        }
    }
    double mx = 0;
#pragma unroll
    for (int k = 0; k < S; k++)
    {
        double re = x[k].x;
        re_out[(size_t)k * FFT_B + c] = re;
        mx = fmax(mx, fabs(re));
    }
    // non-negative doubles order like their bit patterns
    unsigned long long b = (unsigned long long)__double_as_longlong(mx);
    for (int o = 16; o; o >>= 1)
    {
        unsigned long long t = __shfl_xor_sync(0xffffffffu, b, o);
        b = t > b ? t : b;
    }
    if ((threadIdx.x & 31) == 0)
        atomicMax(max_bits, b);
}

// Forward DWT (Cooley-Tukey) for decode: first LOGS stages over stride-2048 columns
// (dwthandler.h:94-191; stage with m groups uses roots[m + i]).
template <int LOGS>
__global__ void __launch_bounds__(256) k_fft_cols(cplx *__restrict__ data, const cplx *__restrict__ roots)
{
    pdl_prologue();
    constexpr int S = 1 << LOGS;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    cplx x[S];
#pragma unroll
    for (int k = 0; k < S; k++)
        x[k] = data[(size_t)k * FFT_B + c];
#pragma unroll
    for (int j = 0; j < LOGS; j++)
    {
        const int half = S >> (j + 1);
        const int m = 1 << j;
#pragma unroll
        for (int k = 0; k < S; k++)
        {
            if (!(k & half))
            {
                cplx r = roots[m + (k >> (LOGS - j))];
                cplx u = x[k], v = c_mul(x[k + half], r);
                x[k] = c_add(u, v);
                x[k + half] = c_sub(u, v);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < S; k++)
        data[(size_t)k * FFT_B + c] = x[k];
}

__global__ void __launch_bounds__(256) k_fft_block(cplx *__restrict__ data, const cplx *__restrict__ roots, int log_n)
{
    pdl_prologue();
    __shared__ cplx sm[FFT_B];
    const size_t n = size_t(1) << log_n;
    cplx *base = data + (size_t)blockIdx.x * FFT_B;
    for (int i = threadIdx.x; i < FFT_B; i += 256)
        sm[i] = base[i];
    __syncthreads();
    for (int lg = FFT_LB - 1; lg >= 0; lg--)
    {
        const int gap = 1 << lg;
        const size_t m = n >> (lg + 1);
        for (int b = threadIdx.x; b < FFT_B / 2; b += 256)
        {
            int grp = b >> lg, j = b & (gap - 1);
            int xi = (grp << (lg + 1)) + j;
            size_t gi = ((size_t)blockIdx.x * FFT_B + xi) >> (lg + 1);
            cplx r = roots[m + gi];
            cplx u = sm[xi], v = c_mul(sm[xi + gap], r);
            sm[xi] = c_add(u, v);
            sm[xi + gap] = c_sub(u, v);
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < FFT_B; i += 256)
        base[i] = sm[i];
}

// Round + RNS-decompose a real coefficient (ckks.h:536-628).  All three reference branches
// compute |round(x)| mod q exactly and negate for negative x; we do the same from the exact
// binary expansion m * 2^e of the rounded double.
__device__ __forceinline__ u64 real_to_residue(double v, const PrimeDev &pd)
{
    double coeffd = round(v);
    bool neg = signbit(coeffd);
    coeffd = fabs(coeffd);
    u64 r;
    if (coeffd < 18446744073709551616.0)
    {
        r = barrett64((u64)coeffd, pd);
    }
    else
    {
        int e;
        double fr = frexp(coeffd, &e);
        u64 m = (u64)ldexp(fr, 53);
        e -= 53; // > 0 here
        r = barrett64(m, pd);
        for (int i = 0; i < e; i++)
            r = csub(r << 1, pd.q);
    }
    return (neg && r) ? pd.q - r : r;
}

struct LdEncode
{
    const double *re; // [N]
    size_t n;
    int limbs;
    int special_pos = -1, special_prime = 0; // extended encodes: the last job is the special prime
    __device__ __forceinline__ bool skip(int) const { return false; }
    __device__ __forceinline__ int prime(int job) const { return job == special_pos ? special_prime : job; }
    __device__ __forceinline__ u64 load(int, int idx, const PrimeDev &pd) const
    {
        return real_to_residue(re[idx], pd);
    }
};

// ---- decode: CRT composition + centred conversion to double (ckks.h:690-742) --------------------
constexpr int MAXP = 62;
__global__ void __launch_bounds__(128) k_dec_compose(const u64 *__restrict__ coeffs /*[l][N] coefficient form*/,
                                                     cplx *__restrict__ out, const PrimeDev *primes,
                                                     const u64 *__restrict__ ginv, const u64 *__restrict__ prodmod,
                                                     const u64 *__restrict__ prodwords,
                                                     const u64 *__restrict__ total, const u64 *__restrict__ half,
                                                     int log_n, int l, int P, double inv_scale, int sparsity)
{
    pdl_prologue();
    const size_t n = size_t(1) << log_n;
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    u64 X[MAXP];
    u64 v[MAXP];
    for (int w = 0; w < l; w++)
        X[w] = 0;
    bool zeroed = sparsity > 1 && (((i - 1) & (size_t)(sparsity - 1)) != (size_t)(sparsity - 1));
    if (!zeroed)
    {
        // Garner mixed-radix digits
        for (int j = 0; j < l; j++)
        {
            const PrimeDev pd = primes[j];
            u64 xj = coeffs[(size_t)j * n + i];
            u64 lo = 0, hi = 0;
            for (int k = 0; k < j; k++)
                mac128(lo, hi, barrett64(v[k], pd), prodmod[(size_t)j * P + k]);
            u64 s = barrett128(lo, hi, pd);
            u64 d = submod(xj, s, pd.q);
            v[j] = mulmod(d, ginv[j], pd);
            // X += v[j] * (q_0..q_{j-1})
            u64 carry = 0;
            for (int w = 0; w < l; w++)
            {
                u64 pw = prodwords[(size_t)j * P + w];
                u64 plo = v[j] * pw, phi = __umul64hi(v[j], pw);
                u64 t = X[w] + plo;
                u64 c1 = t < plo;
                u64 t2 = t + carry;
                u64 c2 = t2 < carry;
                X[w] = t2;
                carry = phi + c1 + c2;
            }
        }
    }
    const u64 *Q = total + (size_t)l * P;
    const u64 *H = half + (size_t)l * P;
    bool ge = true; // X >= H ?
    for (int w = l - 1; w >= 0; w--)
    {
        if (X[w] != H[w])
        {
            ge = X[w] > H[w];
            break;
        }
    }
    double res = 0.0;
    double sc = inv_scale;
    const double two_pow_64 = 18446744073709551616.0;
    if (ge)
    {
        for (int w = 0; w < l; w++, sc = __dmul_rn(sc, two_pow_64))
        {
            if (X[w] > Q[w])
            {
                u64 diff = X[w] - Q[w];
                res = __dadd_rn(res, diff ? __dmul_rn((double)diff, sc) : 0.0);
            }
            else
            {
                u64 diff = Q[w] - X[w];
                res = __dsub_rn(res, diff ? __dmul_rn((double)diff, sc) : 0.0);
            }
        }
    }
    else
    {
        for (int w = 0; w < l; w++, sc = __dmul_rn(sc, two_pow_64))
        {
            u64 cc = X[w];
            res = __dadd_rn(res, cc ? __dmul_rn((double)cc, sc) : 0.0);
        }
    }
    out[i] = make_double2(res, 0.0);
}

__global__ void k_dec_gather(const cplx *__restrict__ res, const uint32_t *__restrict__ map, cplx *__restrict__ out,
                             int count)
{
    pdl_prologue();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count)
        out[i] = res[map[i]];
}

namespace bk
{
    template <class F>
    static void dispatch_cols(int logs, F &&f)
    {
        switch (logs)
        {
        case 1: f(std::integral_constant<int, 1>()); break;
        case 2: f(std::integral_constant<int, 2>()); break;
        case 3: f(std::integral_constant<int, 3>()); break;
        case 4: f(std::integral_constant<int, 4>()); break;
        default: f(std::integral_constant<int, 5>()); break;
        }
    }

    // device-resident values -> plaintext.  d_vals: n_values complex on the device.
    // check_limbs: level whose bit count bounds the scale / coefficient checks (the reference's
    // 2-argument encode works at the top level and the caller then drops limbs).
    // ext > 0: `ext` further limbs after the first `limbs` - primes limbs .. limbs + ext - 2 and the special prime (the
    // special moduli of the level-aware key switch at that level), for operands multiplied in the extended basis.
    void encode_device(Context &c, cudaStream_t s, const cplx *d_vals, int n_values, int limbs, int check_limbs,
                       double scale, bk_pt_t out, int ext = 0)
    {
        EncoderState &e = encoder(c);
        const size_t n = c.n;
        const int slots = (int)(n >> 1);
        if (n_values > slots)
            throw std::invalid_argument("values_size is too large");
        if (limbs < 1 || limbs > c.top_limbs())
            throw std::invalid_argument("parms_id is not valid for encryption parameters");
        if (check_limbs < limbs || check_limbs > c.top_limbs())
            throw std::invalid_argument("parms_id is not valid for encryption parameters");
        if (scale <= 0 || ((int)std::log2(scale) + 1 >= c.total_bits[check_limbs]))
            throw std::invalid_argument("scale out of bounds");
        Scratch buf(s, 2 * n);      // N complex
        Scratch re(s, n + 1);       // N doubles + max slot
        cplx *cv = (cplx *)buf.p;
        unsigned long long *d_max = (unsigned long long *)(re.p + n);
        BK_CUDA(cudaMemsetAsync(cv, 0, n * sizeof(cplx), s));
        BK_CUDA(cudaMemsetAsync(d_max, 0, sizeof(unsigned long long), s));
        if (n_values > 0)
        {
            launch_pdl(k_enc_scatter, (n_values + 255) / 256, 256, 0, s, d_vals, n_values, e.d_index_map, cv, slots);
            c.count();
        }
        double fix = scale / static_cast<double>(n);
        const int logs = c.log_n - FFT_LB;
        {
            ProfScope ps(c, s, TAG_FFT, 2);
            launch_pdl(k_ifft_block, (unsigned)(n >> FFT_LB), 256, 0, s, cv, e.d_inv_roots, c.log_n);
            dispatch_cols(logs, [&](auto L) {
                launch_pdl(k_ifft_cols<decltype(L)::value>, FFT_B / 256, 256, 0, s, cv, e.d_inv_roots, fix, (double *)re.p, d_max,
                                                                            c.log_n);
            });
        }
        c.count(2);
        // "encoded values are too large" check (ckks.h:519-527): ceil(log2(max |coeff|)) + 1 >= total bit count.
        // A finite double is below 2^1024, so for moduli of more than 1025 bits the condition cannot hold and the
        // device-to-host read of the maximum (a full drain of the stream) is skipped - that is every run-time encode
        // of the reference's apps, which encode at the top level (ckks.h:179-184).
        if (c.total_bits[check_limbs] <= 1025)
        {
            unsigned long long h_max = 0;
            BK_CUDA(cudaMemcpyAsync(&h_max, d_max, sizeof(h_max), cudaMemcpyDeviceToHost, s));
            c.d2h_bytes += sizeof(h_max);
            BK_CUDA(cudaStreamSynchronize(s));
            double max_coeff;
            std::memcpy(&max_coeff, &h_max, sizeof(double));
            int max_coeff_bit_count = static_cast<int>(std::ceil(std::log2(std::max<>(max_coeff, 1.0)))) + 1;
            if (max_coeff_bit_count >= c.total_bits[check_limbs])
                throw std::invalid_argument("encoded values are too large");
        }
        const int jobs = limbs + ext;
        if (ext < 0 || limbs + std::max(ext - 1, 0) > c.top_limbs())
            throw std::invalid_argument("extended limbs are out of range");
        ensure_pt(out, jobs);
        out->limbs = limbs;
        out->ext = ext;
        // RNS decomposition fused into the first NTT pass (ckks.h:536-634)
        {
            Scratch tmp(s, (size_t)jobs * n);
            LdEncode ld{ (const double *)re.p, n, jobs };
            JobMap map = limb_map(jobs);
            if (ext > 0)
            {
                ld.special_pos = jobs - 1;
                ld.special_prime = c.n_primes - 1;
                map.special_pos = jobs - 1;
                map.special_prime = c.n_primes - 1;
            }
            dim3 grid(16, jobs);
            {
            ProfScope ps(c, s, TAG_FWD_COLS, jobs);
            switch (c.log_n)
            {
            case 12: launch_pdl(k_fwd_cols<4, LdEncode>, grid, 16, 0, s, ld, tmp.p, c.tables); break;
            case 13: launch_pdl(k_fwd_cols<5, LdEncode>, grid, 32, 0, s, ld, tmp.p, c.tables); break;
            case 14: launch_pdl(k_fwd_cols<6, LdEncode>, grid, 64, 0, s, ld, tmp.p, c.tables); break;
            case 15: launch_pdl(k_fwd_cols<7, LdEncode>, grid, 128, 0, s, ld, tmp.p, c.tables); break;
            default: launch_pdl(k_fwd_cols<8, LdEncode>, grid, 256, 0, s, ld, tmp.p, c.tables); break;
            }
            }
            c.count();
            StPlain st{ out->d, map, n };
            dim3 grid2((unsigned)(n >> 12), jobs);
            {
                ProfScope ps(c, s, TAG_FWD_BLOCKS, jobs);
                launch_pdl(k_fwd_blocks<StPlain>, grid2, 256, 256 * 128, s, tmp.p, st, c.tables);
            }
            c.count();
        }
        out->scale = scale;
    }
} // namespace bk

extern "C"
{
    static void encode_host(bk_context_t ctx, const double *values, int n_values, int is_complex, int limbs,
                            int check_limbs, double scale, bk_pt_t out, int ext = 0)
    {
        Context &c = *ctx;
        if (!values && n_values > 0)
            throw std::invalid_argument("values cannot be null");
        if (n_values < 0 || (size_t)n_values > (c.n >> 1))
            throw std::invalid_argument("values_size is too large");
        cudaStream_t s = c.stream();
        Scratch dv(s, (size_t)2 * std::max(n_values, 1));
        if (n_values > 0)
        {
            // through the pinned ring: the call returns without waiting for the copy, the caller's buffer is free
            cudaEvent_t done;
            cplx *stage = (cplx *)c.staging((size_t)n_values * sizeof(cplx), &done);
            if (is_complex)
                std::memcpy(stage, values, (size_t)n_values * sizeof(cplx));
            else
                for (int i = 0; i < n_values; i++)
                    stage[i] = make_double2(values[i], 0.0);
            BK_CUDA(cudaMemcpyAsync(dv.p, stage, (size_t)n_values * sizeof(cplx), cudaMemcpyHostToDevice, s));
            BK_CUDA(cudaEventRecord(done, s));
            c.h2d_bytes += (size_t)n_values * sizeof(cplx);
        }
        encode_device(c, s, (const cplx *)dv.p, n_values, limbs, check_limbs, scale, out, ext);
    }

    bk_status bk_encode(bk_context_t ctx, const double *values, int n_values, int is_complex, int limbs, double scale,
                        bk_pt_t out)
    {
        BK_TRY
        encode_host(ctx, values, n_values, is_complex, limbs, limbs, scale, out);
        BK_END
    }

    bk_status bk_encode_top_dropped(bk_context_t ctx, const double *values, int n_values, int is_complex, int limbs,
                                    double scale, bk_pt_t out)
    {
        BK_TRY
        encode_host(ctx, values, n_values, is_complex, limbs, ctx->top_limbs(), scale, out);
        BK_END
    }

    bk_status bk_encode_ext(bk_context_t ctx, const double *values, int n_values, int is_complex, int limbs, double scale,
                            bk_pt_t out)
    {
        BK_TRY
        Context &c = *ctx;
        if (limbs < 1 || limbs > c.top_limbs())
            throw std::invalid_argument("parms_id is not valid for encryption parameters");
        encode_host(ctx, values, n_values, is_complex, limbs, c.top_limbs(), scale, out, hybrid_plan(c, limbs).alpha);
        BK_END
    }

    bk_status bk_encode_scalar(bk_context_t ctx, double value, int limbs, double scale, bk_pt_t out)
    {
        BK_TRY
        // ckks.cpp:77-216: constant polynomial, already "NTT form" (a constant in every slot)
        Context &c = *ctx;
        if (limbs < 1 || limbs > c.top_limbs())
            throw std::invalid_argument("parms_id is not valid for encryption parameters");
        if (scale <= 0 || ((int)std::log2(scale) >= c.total_bits[limbs]))
            throw std::invalid_argument("scale out of bounds");
        double v = value * scale;
        int coeff_bit_count = (int)std::log2(std::fabs(v)) + 2;
        if (coeff_bit_count >= c.total_bits[limbs])
            throw std::invalid_argument("encoded value is too large");
        double coeffd = std::round(v);
        bool neg = std::signbit(coeffd);
        coeffd = std::fabs(coeffd);
        int e = 0;
        uint64_t m = 0;
        if (coeffd != 0)
        {
            double fr = std::frexp(coeffd, &e);
            m = (uint64_t)std::ldexp(fr, 53);
            e -= 53;
        }
        ensure_pt(out, limbs);
        cudaStream_t s = c.stream();
        std::vector<u64> host((size_t)limbs * c.n);
        for (int j = 0; j < limbs; j++)
        {
            uint64_t q = c.primes[j], r;
            if (e >= 0)
                r = mulmod(m % q, powmod(2, (uint64_t)e, q), q);
            else
                r = (e <= -64 ? 0 : (m >> (-e))) % q;
            if (neg && r)
                r = q - r;
            std::fill(host.begin() + (size_t)j * c.n, host.begin() + (size_t)(j + 1) * c.n, r);
        }
        BK_CUDA(cudaMemcpyAsync(out->d, host.data(), host.size() * sizeof(u64), cudaMemcpyHostToDevice, s));
        BK_CUDA(cudaStreamSynchronize(s));
        out->scale = scale;
        BK_END
    }

    bk_status bk_decode(bk_context_t ctx, bk_pt_t pt, double *out_complex)
    {
        BK_TRY
        Context &c = *ctx;
        if (!pt || pt->ctx != ctx || !pt->d)
            throw std::invalid_argument("plain is not valid for encryption parameters");
        if (!out_complex)
            throw std::invalid_argument("destination cannot be null");
        const int l = pt->limbs;
        if (pt->scale <= 0 || ((int)std::log2(pt->scale) >= c.total_bits[l]))
            throw std::invalid_argument("scale out of bounds");
        EncoderState &e = encoder(c);
        cudaStream_t s = c.stream();
        const size_t n = c.n;
        const int slots = (int)(n >> 1);
        int sparse = c.sparse_slots ? c.sparse_slots : slots;
        Scratch copy(s, (size_t)l * n);
        BK_CUDA(cudaMemcpyAsync(copy.p, pt->d, (size_t)l * n * sizeof(u64), cudaMemcpyDeviceToDevice, s));
        ntt_inv(c, s, copy.p, l, limb_map(l));
        Scratch res(s, 2 * n);
        cplx *rv = (cplx *)res.p;
        double inv_scale = double(1.0) / pt->scale;
        launch_pdl(k_dec_compose, (unsigned)((n + 127) / 128), 128, 0, s, copy.p, rv, c.d_primes, e.d_garner_inv, e.d_prodmod,
                                                                  e.d_prodwords, e.d_total, e.d_half, c.log_n, l,
                                                                  c.n_primes, inv_scale, slots / sparse);
        c.count();
        const int logs = c.log_n - FFT_LB;
        dispatch_cols(logs, [&](auto L) {
            launch_pdl(k_fft_cols<decltype(L)::value>, FFT_B / 256, 256, 0, s, rv, e.d_roots);
        });
        c.count();
        launch_pdl(k_fft_block, (unsigned)(n >> FFT_LB), 256, 0, s, rv, e.d_roots, c.log_n);
        c.count();
        Scratch outv(s, (size_t)2 * slots);
        BK_CUDA(cudaMemsetAsync(outv.p, 0, (size_t)slots * sizeof(cplx), s));
        launch_pdl(k_dec_gather, (sparse + 255) / 256, 256, 0, s, rv, e.d_index_map, (cplx *)outv.p, sparse);
        c.count();
        BK_CUDA(cudaMemcpyAsync(out_complex, outv.p, (size_t)slots * sizeof(cplx), cudaMemcpyDeviceToHost, s));
        c.d2h_bytes += (size_t)slots * sizeof(cplx);
        BK_CUDA(cudaStreamSynchronize(s));
        BK_END
    }

    bk_status bk_set_sparse_slots(bk_context_t ctx, int sparse_slots)
    {
        BK_TRY
        if (sparse_slots < 0 || (size_t)sparse_slots > (ctx->n >> 1) || (sparse_slots & (sparse_slots - 1)))
            throw std::invalid_argument("sparse_slots must be a power of two <= slot_count");
        ctx->sparse_slots = sparse_slots;
        BK_END
    }
}
