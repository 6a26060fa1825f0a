// ntt.cuh - negacyclic NTT / INTT over RNS limbs for sm_100a, N = 2^12 .. 2^16.
//
// Replaces the reference's scalar loops DWTHandler::transform_to_rev / transform_from_rev
// (seal/util/dwthandler.h:94-191, :202-356) as driven by ntt_negacyclic_harvey[_lazy] and
// inverse_ntt_negacyclic_harvey[_lazy] (seal/util/ntt.h:235-264, :336-358).
//
// Math (identical to the reference): forward = Cooley-Tukey, natural -> bit-reversed order,
// stage with m groups uses twiddle tw[m+i] = psi^bitrev(m+i) for group i (ntt.cpp:58-67),
// Harvey lazy butterflies with Shoup multiplication, values kept in [0,4q).  Inverse =
// Gentleman-Sande with itw[m+i] = tw[m+i]^-1 and N^-1 folded into the last stage.
//
// B200 mapping (this is not how the reference does it): the N-point transform is split
//   N = R x 256:  "column pass"  = the first log2(R) stages, stride-256 butterflies,
//                 "block pass"   = the last 8 stages, inside contiguous 256-element blocks.
// Each pass is a radix-16 x radix-16 register kernel: a thread owns 16 coefficients, does 4
// stages in registers, exchanges through (XOR-swizzled, conflict-free) shared memory, does 4
// more.  Every global access is a full 128-byte line per half-warp.  The intermediate between
// the passes (<= 31 MiB for a whole ciphertext) stays in B200's 126 MB L2, so HBM sees one
// read and one write per limb.  Load / store functors fuse the surrounding RNS arithmetic
// (base conversion, rounding fix-ups, scalar multiplication, Galois gather) into the passes.
#pragma once
#include "modarith.cuh"


struct NttTables
{
    const ulonglong2 *tw;  // [n_primes][N] {w, shoup(w)} forward, index m+i
    const ulonglong2 *itw; // [n_primes][N] inverse, same indexing
    const PrimeDev *primes;
    int log_n;
    int wide; // forward transforms use the unreduced butterflies below: 1 = for every modulus (all in (2^32, 2^57)),
              // 2 = for the moduli in that range, lazy butterflies for the others, 0 = never
};

// ---- butterflies ---------------------------------------------------------------------------
__device__ __forceinline__ void ct_bfly(u64 &X, u64 &Y, ulonglong2 w, u64 q, u64 two_q)
{
    u64 x = X - (X >= two_q ? two_q : 0ull);
    u64 t = mul_shoup_lazy(Y, w.x, w.y, q);
    X = x + t;
    Y = x + two_q - t;
}

__device__ __forceinline__ void gs_bfly(u64 &X, u64 &Y, ulonglong2 w, u64 q, u64 two_q)
{
    u64 s = X + Y;
    u64 d = X + two_q - Y;
    X = s - (s >= two_q ? two_q : 0ull);
    Y = mul_shoup_lazy(d, w.x, w.y, q);
}

// moduli the unreduced butterflies (and barrett64_r32) apply to
__device__ __forceinline__ bool wide_modulus(u64 q)
{
    return (q >> 57) == 0 && (q >> 32) != 0;
}

// Unreduced forward butterfly for moduli below 2^57.  Nothing is reduced between stages: with t in [0,4q)
//   X' = X + t,   Y' = X + 4q - t,
// both outputs grow by at most 4q per stage, so 16 stages starting below 6q (loaders return values below 4q, the
// basis conversion of hybrid key switching below 6q) stay below 70q < 2^64.  The Shoup
// quotient is taken from three of the four 32x32 partial products (x1 w1 + hi(x1 w0) + hi(x0 w1)); it undershoots
// floor(x ws / 2^64) by at most 2, hence t = x w - quot q lies in [0,4q) instead of [0,2q).  Residues mod q are
// unaffected, and every value is brought back to [0,q) before it leaves the transform.
__device__ __forceinline__ void ct_bfly_wide(u64 &X, u64 &Y, ulonglong2 w, u64 neg_q, u64 four_q)
{
    const unsigned y0 = (unsigned)Y, y1 = (unsigned)(Y >> 32);
    const unsigned s0 = (unsigned)w.y, s1 = (unsigned)(w.y >> 32);
    const u64 quot = (u64)y1 * s1 + __umulhi(y1, s0) + __umulhi(y0, s1);
    const u64 t = Y * w.x + quot * neg_q;
    const u64 x = X;
    X = x + t;
    Y = x + four_q - t;
}

template <int LOGS>
__device__ __forceinline__ void fwd_radix_wide(u64 *x, const ulonglong2 *__restrict__ tw, unsigned idx0, u64 neg_q, u64 four_q)
{
    constexpr int S = 1 << LOGS;
#pragma unroll
    for (int j = 0; j < LOGS; j++)
    {
        const int half = S >> (j + 1);
#pragma unroll
        for (int k = 0; k < S; k++)
        {
            if (!(k & half))
            {
                ulonglong2 w = __ldg(tw + ((idx0 << j) + (unsigned)(k >> (LOGS - j))));
                ct_bfly_wide(x[k], x[k + half], w, neg_q, four_q);
            }
        }
    }
}

// LOGS forward stages on S = 2^LOGS registers.  Stage j pairs (k, k + S/2^(j+1)); its twiddle
// index is (idx0 << j) + (k >> (LOGS - j)) where idx0 = m_loc * B + T identifies the first stage.
template <int LOGS>
__device__ __forceinline__ void fwd_radix(u64 *x, const ulonglong2 *__restrict__ tw, unsigned idx0, u64 q, u64 two_q)
{
    constexpr int S = 1 << LOGS;
#pragma unroll
    for (int j = 0; j < LOGS; j++)
    {
        const int half = S >> (j + 1);
#pragma unroll
        for (int k = 0; k < S; k++)
        {
            if (!(k & half))
            {
                ulonglong2 w = __ldg(tw + ((idx0 << j) + (unsigned)(k >> (LOGS - j))));
                ct_bfly(x[k], x[k + half], w, q, two_q);
            }
        }
    }
}

// Inverse of fwd_radix (stages in reverse order).  If FOLD, stage j == 0 (the last one, twiddle
// index idx0 == 1) multiplies by N^-1:  X' = (X+Y) N^-1,  Y' = (X-Y) N^-1 itw[1].
template <int LOGS, bool FOLD>
__device__ __forceinline__ void inv_radix(u64 *x, const ulonglong2 *__restrict__ itw, unsigned idx0,
                                          const PrimeDev &pd)
{
    constexpr int S = 1 << LOGS;
    const u64 q = pd.q, two_q = pd.two_q;
#pragma unroll
    for (int j = LOGS - 1; j >= 0; j--)
    {
        const int half = S >> (j + 1);
#pragma unroll
        for (int k = 0; k < S; k++)
        {
            if (!(k & half))
            {
                if (FOLD && j == 0)
                {
                    u64 s = x[k] + x[k + half];
                    u64 d = x[k] + two_q - x[k + half];
                    x[k] = mul_shoup_lazy(s, pd.ninv, pd.ninv_s, q);
                    x[k + half] = mul_shoup_lazy(d, pd.ninvw, pd.ninvw_s, q);
                }
                else
                {
                    ulonglong2 w = __ldg(itw + ((idx0 << j) + (unsigned)(k >> (LOGS - j))));
                    gs_bfly(x[k], x[k + half], w, q, two_q);
                }
            }
        }
    }
}

// XOR swizzle for a 256-element block held in shared memory as 8-byte words: both access
// patterns used by the block pass (e = t + 16k and e = 16t + k, t = lane within the
// half-warp) hit 16 distinct 8-byte banks.
__device__ __forceinline__ int swz(int e)
{
    return e ^ ((e >> 4) & 15);
}

// shared-memory offset of a row of the column-pass tile.  With 8 columns the four row groups a warp touches in the
// transposed phase (rows 16 apart) would fall on the same 16 banks; every group of 16 rows is shifted by a further
// 64 bytes (tile size R * COLS + R / 16 * 8 words).
template <int COLS>
__device__ __forceinline__ int cols_idx(int row)
{
    return row * COLS + (COLS == 8 ? (row >> 4) * 8 : 0);
}

// ============================================================================================
// Column pass, forward (first LOGR stages).  grid = (256 / COLS, jobs), block = COLS * (R/16) threads.
// CTA tile: all R rows x COLS consecutive columns: COLS = 16 is one 128-byte line per row; launches of a few
// limb-polynomials (the special limbs of a ModDown, the last limb of a rescale, key switches at 2-3 limbs) use
// COLS = 8 so that twice as many CTAs spread over the 148 SMs.
//   Load:  int prime(int job);  bool skip(int job);  u64 load(int job, int idx, const PrimeDev&)
//          (must return a value < 4q of prime(job)).
// Output: lazy values in [0,4q) at out[job*N + idx].
// ============================================================================================
template <int LOGR, class Load, int WIDE = 0, int COLS = 16>
__global__ void __launch_bounds__(256) k_fwd_cols(Load ld, u64 *__restrict__ out, NttTables T)
{
    pdl_prologue();
    constexpr int R = 1 << LOGR;
    constexpr int TR = R / 16;     // threads along rows
    constexpr int LOG2 = LOGR - 4; // stages in the second phase
    constexpr int S2 = 1 << LOG2;  // = TR
    constexpr int G = 16 / S2;     // groups per thread in the second phase
    __shared__ u64 sm[R * COLS + (R / 16 + 1) * 8];

    const int job = blockIdx.y;
    if (ld.skip(job))
        return;
    const int c = threadIdx.x % COLS;
    const int t = threadIdx.x / COLS;
    const int col = blockIdx.x * COLS + c;
    const int pi = ld.prime(job);
    const PrimeDev pd = T.primes[pi];
    const size_t n = size_t(1) << T.log_n;
    const ulonglong2 *tw = T.tw + (size_t)pi * n;

    u64 x[16];
    // phase 1: rows t + TR*k, radix-16 over the top 4 row bits (stages m = 1,2,4,8)
#pragma unroll
    for (int k = 0; k < 16; k++)
        x[k] = ld.load(job, (t + TR * k) * 256 + col, pd);
    // host dispatch on T.wide: 1 = every modulus lies in (2^32, 2^57); 2 = some do (the GPT-2 chain ends in a 60-bit
    // special prime) and a launch may mix them, so the job's own modulus decides - uniform over the CTA
    const bool wide = WIDE == 1 || (WIDE == 2 && wide_modulus(pd.q));
    const u64 neg_q = 0ull - pd.q, four_q = 2 * pd.two_q;
    if (wide)
        fwd_radix_wide<4>(x, tw, 1u, neg_q, four_q);
    else
        fwd_radix<4>(x, tw, 1u, pd.q, pd.two_q);
    if (LOG2 > 0)
    {
#pragma unroll
        for (int k = 0; k < 16; k++)
            sm[cols_idx<COLS>(t + TR * k) + c] = x[k];
        __syncthreads();
        // phase 2: G groups; group g covers rows u*S2 + k', u = t*G + g (top 4 bits), k' < S2
#pragma unroll
        for (int g = 0; g < G; g++)
        {
            const int u = t * G + g;
#pragma unroll
            for (int k = 0; k < S2; k++)
                x[g * S2 + k] = sm[cols_idx<COLS>(u * S2 + k) + c];
        }
#pragma unroll
        for (int g = 0; g < G; g++)
        {
            if (wide)
                fwd_radix_wide<LOG2>(x + g * S2, tw, 16u + (unsigned)(t * G + g), neg_q, four_q);
            else
                fwd_radix<LOG2>(x + g * S2, tw, 16u + (unsigned)(t * G + g), pd.q, pd.two_q);
        }
        u64 *o = out + (size_t)job * n;
#pragma unroll
        for (int g = 0; g < G; g++)
        {
            const int u = t * G + g;
#pragma unroll
            for (int k = 0; k < S2; k++)
                o[(size_t)(u * S2 + k) * 256 + col] = x[g * S2 + k];
        }
    }
    else
    {
        u64 *o = out + (size_t)job * n;
#pragma unroll
        for (int k = 0; k < 16; k++)
            o[(size_t)(t + TR * k) * 256 + col] = x[k];
    }
}

// ============================================================================================
// Block pass, forward (last 8 stages).  grid = (N / (16 blockDim.x), jobs), block = 256 (or 64) threads:
// blockDim.x / 16 blocks of 256 coefficients per CTA, one half-warp per block (only __syncwarp needed); small
// launches use 64-thread CTAs so that four times as many CTAs spread over the SMs.
//   Store: int prime(int job);  bool skip(int job);
//          u64  pre (int job, int blk, int t, int k, u64 v, const PrimeDev&)   - register layout
//               e = 16t + k, v in [0,4q); returns the value to be staged;
//          void post(int job, int idx, u64 v, const PrimeDev&)                  - coalesced order;
//          or, with BATCH: void post_all(int job, int base, int t, const u64 *s, const PrimeDev&) for all 16
//          coefficients base + t + 16 k of the thread at once (values at s[swz(t + 16 k)]).
// ============================================================================================
template <class Store, int WIDE = 0>
__global__ void __launch_bounds__(256, 2) k_fwd_blocks(const u64 *__restrict__ in, Store st, NttTables T)
{
    pdl_prologue();
    extern __shared__ __align__(16) u64 sm[]; // 256 words per half-warp: blockDim.x * 16 words
    const int job = blockIdx.y;
    if (st.skip(job))
        return;
    const int t = threadIdx.x & 15;
    const int lb = threadIdx.x >> 4;         // local block 0..15
    const int blk = blockIdx.x * (int)(blockDim.x >> 4) + lb; // 256-block index within the limb
    const int pi = st.prime(job);
    const PrimeDev pd = T.primes[pi];
    const size_t n = size_t(1) << T.log_n;
    const unsigned B = (unsigned)(n >> 8) + (unsigned)blk;
    const ulonglong2 *tw = T.tw + (size_t)pi * n;
    u64 *s = sm + lb * 256;

    const u64 *src = in + (size_t)job * n + (size_t)blk * 256;
    u64 x[16];
#pragma unroll
    for (int k = 0; k < 16; k++)
        x[k] = src[t + 16 * k];
    const bool wide = WIDE == 1 || (WIDE == 2 && wide_modulus(pd.q));
    const u64 neg_q = 0ull - pd.q, four_q = 2 * pd.two_q;
    if (wide)
        fwd_radix_wide<4>(x, tw, B, neg_q, four_q);
    else
        fwd_radix<4>(x, tw, B, pd.q, pd.two_q);
#pragma unroll
    for (int k = 0; k < 16; k++)
        s[swz(t + 16 * k)] = x[k];
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 16; k++)
        x[k] = s[swz(16 * t + k)];
    if (wide)
    {
        fwd_radix_wide<4>(x, tw, 16u * B + (unsigned)t, neg_q, four_q);
        // back to [0,q) unless the consumer takes unreduced words (Store::RAW: the key-switch inner product)
        if (!Store::RAW)
        {
#pragma unroll
            for (int k = 0; k < 16; k++)
                x[k] = barrett64_r32(x[k], pd);
        }
    }
    else
        fwd_radix<4>(x, tw, 16u * B + (unsigned)t, pd.q, pd.two_q);
#pragma unroll
    for (int k = 0; k < 16; k++)
        x[k] = st.pre(job, blk, t, k, x[k], pd);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 16; k++)
        s[swz(16 * t + k)] = x[k];
    __syncwarp();
    if constexpr (Store::BATCH)
        st.post_all(job, blk * 256, t, s, pd);
    else
    {
#pragma unroll
        for (int k = 0; k < 16; k++)
            st.post(job, blk * 256 + t + 16 * k, s[swz(t + 16 * k)], pd);
    }
}

// ============================================================================================
// Block pass, inverse (first 8 GS stages, gaps 1..128).  Same geometry as k_fwd_blocks.
//   Load: int prime(int job); u64 load(int job, int idx, const PrimeDev&) (< 2q, coalesced order)
//         static constexpr bool TLAYOUT: source block is stored transposed (pos = k*16 + t holds
//         e = 16t + k) and is read straight into registers via u64 load_t(job, blk, pos).
// Output: values in [0,2q) at out[job*N + idx].
// ============================================================================================
template <class Load>
__global__ void __launch_bounds__(256) k_inv_blocks(Load ld, u64 *__restrict__ out, NttTables T)
{
    pdl_prologue();
    extern __shared__ __align__(16) u64 sm[]; // 256 words per half-warp: blockDim.x * 16 words
    const int job = blockIdx.y;
    const int t = threadIdx.x & 15;
    const int lb = threadIdx.x >> 4;
    const int blk = blockIdx.x * (int)(blockDim.x >> 4) + lb;
    const int pi = ld.prime(job);
    const PrimeDev pd = T.primes[pi];
    const size_t n = size_t(1) << T.log_n;
    const unsigned B = (unsigned)(n >> 8) + (unsigned)blk;
    const ulonglong2 *itw = T.itw + (size_t)pi * n;
    u64 *s = sm + lb * 256;

    u64 x[16];
    if (Load::TLAYOUT)
    {
#pragma unroll
        for (int k = 0; k < 16; k++)
            x[k] = ld.load_t(job, blk, k * 16 + t);
    }
    else
    {
#pragma unroll
        for (int k = 0; k < 16; k++)
            s[swz(t + 16 * k)] = ld.load(job, blk * 256 + t + 16 * k, pd);
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 16; k++)
            x[k] = s[swz(16 * t + k)];
        __syncwarp();
    }
    inv_radix<4, false>(x, itw, 16u * B + (unsigned)t, pd);
#pragma unroll
    for (int k = 0; k < 16; k++)
        s[swz(16 * t + k)] = x[k];
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 16; k++)
        x[k] = s[swz(t + 16 * k)];
    inv_radix<4, false>(x, itw, B, pd);
    u64 *dst = out + (size_t)job * n + (size_t)blk * 256;
#pragma unroll
    for (int k = 0; k < 16; k++)
        dst[t + 16 * k] = x[k];
}

// ============================================================================================
// Column pass, inverse (last LOGR GS stages, N^-1 folded into the final one).
//   Store: int prime(int job); void store(int job, int idx, u64 v /* [0,2q) */, const PrimeDev&)
// ============================================================================================
template <int LOGR, class Store, int COLS = 16>
__global__ void __launch_bounds__(256) k_inv_cols(const u64 *__restrict__ in, Store st, NttTables T)
{
    pdl_prologue();
    constexpr int R = 1 << LOGR;
    constexpr int TR = R / 16;
    constexpr int LOG2 = LOGR - 4;
    constexpr int S2 = 1 << LOG2;
    constexpr int G = 16 / S2;
    __shared__ u64 sm[R * COLS + (R / 16 + 1) * 8];

    const int job = blockIdx.y;
    const int c = threadIdx.x % COLS;
    const int t = threadIdx.x / COLS;
    const int col = blockIdx.x * COLS + c;
    const int pi = st.prime(job);
    const PrimeDev pd = T.primes[pi];
    const size_t n = size_t(1) << T.log_n;
    const ulonglong2 *itw = T.itw + (size_t)pi * n;
    const u64 *src = in + (size_t)job * n;

    u64 x[16];
    if (LOG2 > 0)
    {
#pragma unroll
        for (int g = 0; g < G; g++)
        {
            const int u = t * G + g;
#pragma unroll
            for (int k = 0; k < S2; k++)
                x[g * S2 + k] = src[(size_t)(u * S2 + k) * 256 + col];
        }
#pragma unroll
        for (int g = 0; g < G; g++)
            inv_radix<LOG2, false>(x + g * S2, itw, 16u + (unsigned)(t * G + g), pd);
#pragma unroll
        for (int g = 0; g < G; g++)
        {
            const int u = t * G + g;
#pragma unroll
            for (int k = 0; k < S2; k++)
                sm[cols_idx<COLS>(u * S2 + k) + c] = x[g * S2 + k];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; k++)
            x[k] = sm[cols_idx<COLS>(t + TR * k) + c];
    }
    else
    {
#pragma unroll
        for (int k = 0; k < 16; k++)
            x[k] = src[(size_t)(t + TR * k) * 256 + col];
    }
    inv_radix<4, true>(x, itw, 1u, pd);
#pragma unroll
    for (int k = 0; k < 16; k++)
        st.store(job, (t + TR * k) * 256 + col, x[k], pd);
}

