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

// (hi:lo) mod q, hi:lo < 2^128 arbitrary as long as q < 2^63 (here q < 2^61).
__device__ __forceinline__ u64 barrett128(u64 lo, u64 hi, const PrimeDev &p)
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

__device__ __forceinline__ u64 mulmod(u64 a, u64 b, const PrimeDev &p)
{
    return barrett128(a * b, __umul64hi(a, b), p);
}

__device__ __forceinline__ u64 addmod(u64 a, u64 b, u64 q)
{
    return csub(a + b, q);
}

__device__ __forceinline__ u64 submod(u64 a, u64 b, u64 q)
{
    return a >= b ? a - b : a + q - b;
}

// 128-bit accumulate of a 64x64 product: (hi:lo) += a*b
__device__ __forceinline__ void mac128(u64 &lo, u64 &hi, u64 a, u64 b)
{
    u64 pl = a * b;
    u64 ph = __umul64hi(a, b);
    asm("add.cc.u64 %0, %0, %2;\n\t"
        "addc.u64 %1, %1, %3;"
        : "+l"(lo), "+l"(hi)
        : "l"(pl), "l"(ph));
}
