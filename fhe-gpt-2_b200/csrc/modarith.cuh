// modarith.cuh - 64-bit modular arithmetic on the sm_100a integer pipes.
//
// Device restatement of the primitives in the reference's seal/util/uintarithsmallmod.h
// (:167-204 barrett_reduce_128, :211-240 barrett_reduce_64, :255-326 MultiplyUIntModOperand
// + multiply_uint_mod_lazy).  All results that leave a kernel are canonical residues in
// [0,q), so they are bit-identical to the reference's regardless of the lazy ranges used
// inside a kernel.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

typedef unsigned long long u64;

// Programmatic dependent launch: every kernel of the engine begins with this.  `wait` blocks until the grids this
// launch depends on have completed and flushed (a no-op when the launch carries no programmatic dependency);
// `launch_dependents` lets the next kernel of the stream be set up and its CTAs be scheduled while this one still
// runs - they stop at their own `wait`, so no kernel touches memory before its predecessors are done, but the launch
// latency between two dependent kernels is hidden (see launch_pdl in engine.h).
__device__ __forceinline__ void pdl_prologue()
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// Per-prime constants resident in HBM (one entry per modulus of the key-level chain).
struct PrimeDev
{
    u64 q;       // modulus (< 2^61)
    u64 two_q;   // 2q
    u64 r0, r1;  // floor(2^128 / q), low / high word (Modulus::const_ratio)
    u64 ninv, ninv_s;   // N^-1 mod q and its Shoup quotient
    u64 ninvw, ninvw_s; // N^-1 * itw[1] mod q (last inverse stage, dwthandler.h:273-314)
};

__device__ __forceinline__ u64 mul_shoup_lazy(u64 x, u64 w, u64 ws, u64 q)
{
    // x*w - floor(x*ws/2^64)*q  in [0, 2q)   (uintarithsmallmod.h:313-326)
    u64 hi = __umul64hi(x, ws);
    return x * w - hi * q;
}

__device__ __forceinline__ u64 csub(u64 x, u64 q)
{
    return x >= q ? x - q : x;
}

__device__ __forceinline__ u64 barrett64(u64 x, const PrimeDev &p)
{
    u64 t = __umul64hi(x, p.r1);
    u64 r = x - t * p.q;
    return csub(r, p.q);
}

// The same for moduli above 2^32: the high word of floor(2^128 / q) then fits in 32 bits, and two 32-bit multiplies
// give the same floor(x r1 / 2^64) as the four of __umul64hi.
__device__ __forceinline__ u64 barrett64_r32(u64 x, const PrimeDev &p)
{
    const unsigned r1 = (unsigned)p.r1;
    u64 t = ((u64)(unsigned)(x >> 32) * r1 + __umulhi((unsigned)x, r1)) >> 32;
    u64 r = x - t * p.q;
    return csub(r, p.q);
}

// (hi:lo) mod q, hi:lo < 2^128 arbitrary as long as q < 2^63 (here q < 2^61): the reference's two-round quotient.
__device__ __forceinline__ u64 barrett128_exact(u64 lo, u64 hi, const PrimeDev &p)
{
    // floor(input * ratio / 2^128), two rounds as in uintarithsmallmod.h:167-204
    u64 carry = __umul64hi(lo, p.r0);
    u64 t2lo = lo * p.r1;
    u64 t2hi = __umul64hi(lo, p.r1);
    u64 t1 = t2lo + carry;
    u64 t3 = t2hi + (t1 < t2lo ? 1ull : 0ull);
    u64 u2lo = hi * p.r0;
    u64 u2hi = __umul64hi(hi, p.r0);
    u64 s = t1 + u2lo;
    u64 c2 = u2hi + (s < t1 ? 1ull : 0ull);
    u64 quot = hi * p.r1 + t3 + c2;
    u64 r = lo - quot * p.q;
    return csub(r, p.q);
}

// The same residue with 17 multiplier-pipe units instead of ~39 (a unit = one IMAD; IMAD.WIDE / IMAD.HI count two):
// for q > 2^32 the high word r1 of floor(2^128 / q) fits in 32 bits, and the quotient only has to be close -
//   quot = hi r1 + floor(lo r1 / 2^64) + (floor(hi r0 / 2^64) from three of its four partial products)
// undershoots floor(x / q) by at most 5 (one from Barrett, two from dropped fractions, two from the dropped partial
// product, none from dropping lo r0 / 2^128 beyond those), so x - quot q lies in [0, 6q) and three conditional
// subtractions finish.  One of these per output coefficient is half the work of the basis conversion of hybrid key
// switching (k_hyb_conv) and the tail of every inner product / plaintext-product sum.
// x - quot q in [0, 6q) for q > 2^32 (see barrett128 below): for consumers that take a lazy representative
__device__ __forceinline__ u64 barrett128_lazy6(u64 lo, u64 hi, const PrimeDev &p)
{
    const unsigned r1 = (unsigned)p.r1, r00 = (unsigned)p.r0, r01 = (unsigned)(p.r0 >> 32);
    const unsigned l0 = (unsigned)lo, l1 = (unsigned)(lo >> 32), h0 = (unsigned)hi, h1 = (unsigned)(hi >> 32);
    u64 quot = (u64)h0 * r1 + ((u64)(h1 * r1) << 32);                   // hi r1 mod 2^64
    quot += ((u64)l1 * r1 + __umulhi(l0, r1)) >> 32;                    // floor(lo r1 / 2^64), exact
    quot += (u64)h1 * r01 + __umulhi(h1, r00) + __umulhi(h0, r01);      // floor(hi r0 / 2^64) - {0, 1, 2}
    return lo - quot * p.q;
}

__device__ __forceinline__ u64 barrett128(u64 lo, u64 hi, const PrimeDev &p)
{
    if (p.r1 >> 32) // moduli below 2^32 (test chains): uniform over the launch
        return barrett128_exact(lo, hi, p);
    u64 r = barrett128_lazy6(lo, hi, p);
    const u64 four_q = 2 * p.two_q;
    r = r >= four_q ? r - four_q : r;
    r = r >= p.two_q ? r - p.two_q : r;
    return csub(r, p.q);
}



__device__ __forceinline__ u64 addmod(u64 a, u64 b, u64 q)
{
    return csub(a + b, q);
}

__device__ __forceinline__ u64 submod(u64 a, u64 b, u64 q)
{
    return a >= b ? a - b : a + q - b;
}

// 128-bit accumulate of a 64x64 product: (hi:lo) += a*b, for (a >> 32) + (b >> 32) < 2^32 - true for every use in
// the engine: b is a residue or a key / plaintext / table word below its modulus (< 2^60), a is a residue or an
// unreduced NTT output (< 70 q with q < 2^57, or < 4 q with q < 2^60).  Written out in 32-bit halves because
// `a * b` next to `__umul64hi(a, b)` compiles to FIVE wide multiplies and two narrow ones (the low product twice):
// here it is the four wide multiplies of the schoolbook product - the two cross products share one 64-bit sum, which
// cannot overflow under the precondition - and the multiplier pipe, which bounds the key-switch inner product and the
// basis conversion, does a third less work per term (IMAD.WIDE issues at half the rate of IMAD, profiles/r2_ntt_lab.md).
__device__ __forceinline__ void mac128(u64 &lo, u64 &hi, u64 a, u64 b)
{
    asm("{\n\t"
        ".reg .u32 a0, a1, b0, b1, l0, l1, m0, m1;\n\t"
        ".reg .u64 p00, mid, t, plo, phi;\n\t"
        "mov.b64 {a0, a1}, %2;\n\t"
        "mov.b64 {b0, b1}, %3;\n\t"
        "mul.wide.u32 p00, a0, b0;\n\t"
        "mul.wide.u32 mid, a0, b1;\n\t"
        "mad.wide.u32 mid, a1, b0, mid;\n\t"
        "mov.b64 {l0, l1}, p00;\n\t"
        "mov.b64 {m0, m1}, mid;\n\t"
        "add.cc.u32 l1, l1, m0;\n\t"
        "addc.u32 m1, m1, 0;\n\t"
        "mov.b64 plo, {l0, l1};\n\t"
        "mov.b64 t, {m1, 0};\n\t"
        "mad.wide.u32 phi, a1, b1, t;\n\t"
        "add.cc.u64 %0, %0, plo;\n\t"
        "addc.u64 %1, %1, phi;\n\t"
        "}"
        : "+l"(lo), "+l"(hi)
        : "l"(a), "l"(b));
}

// a b mod q for residues a, b (mac128's precondition holds)
__device__ __forceinline__ u64 mulmod(u64 a, u64 b, const PrimeDev &p)
{
    u64 lo = 0, hi = 0;
    mac128(lo, hi, a, b);
    return barrett128(lo, hi, p);
}
