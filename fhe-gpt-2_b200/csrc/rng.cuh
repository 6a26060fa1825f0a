// rng.cuh - the engine's random generator: ChaCha20 (RFC 8439 block function) in counter mode.
//
// Replaces the reference's Blake2xb/SHAKE stream (randomgen.h, util/blake2xb.c) as the source of key and encryption
// randomness.  Every context holds a 256-bit master key drawn from the operating system (getrandom(2)) - or expanded
// from $B200CKKS_SEED / bk_context_set_rng_key for reproducible runs.  Every C-ABI call that samples takes a 64-bit
// `seed` argument, which is only a nonce: the call's 256-bit key is ChaCha20(master, counter = seed), and a sample is
// word(s) of ChaCha20(call key, counter = element index, nonce = (stream, attempt)).  Distinct (seed, stream, index)
// give independent blocks; nothing is derivable without the master key.
#pragma once
#include <cstdint>

namespace bk
{
    struct RngKey
    {
        uint32_t k[8];
    };

    __host__ __device__ __forceinline__ uint32_t rotl32(uint32_t x, int r)
    {
        return (x << r) | (x >> (32 - r));
    }
#define BK_CHACHA_QR(a, b, c, d)                                                                                       \
    a += b; d ^= a; d = rotl32(d, 16);                                                                                 \
    c += d; b ^= c; b = rotl32(b, 12);                                                                                 \
    a += b; d ^= a; d = rotl32(d, 8);                                                                                  \
    c += d; b ^= c; b = rotl32(b, 7);

    // out[0..words) of the ChaCha20 block with this key, 64-bit block counter (c0, c1) and 64-bit nonce (n0, n1)
    // DOUBLE_ROUNDS = 10 is ChaCha20; 4 (ChaCha8) is used only for PUBLIC values - the uniform halves `a` of
    // evaluation keys, which are published with the key and have to look uniform, not stay secret (keygen.cu)
    template <int WORDS, int DOUBLE_ROUNDS = 10>
    __host__ __device__ __forceinline__ void chacha_block(const RngKey &key, uint32_t c0, uint32_t c1, uint32_t n0, uint32_t n1,
                                                          uint32_t *out)
    {
        uint32_t x0 = 0x61707865u, x1 = 0x3320646eu, x2 = 0x79622d32u, x3 = 0x6b206574u;
        uint32_t x4 = key.k[0], x5 = key.k[1], x6 = key.k[2], x7 = key.k[3], x8 = key.k[4], x9 = key.k[5], x10 = key.k[6],
                 x11 = key.k[7];
        uint32_t x12 = c0, x13 = c1, x14 = n0, x15 = n1;
#pragma unroll
        for (int r = 0; r < DOUBLE_ROUNDS; r++)
        {
            BK_CHACHA_QR(x0, x4, x8, x12)
            BK_CHACHA_QR(x1, x5, x9, x13)
            BK_CHACHA_QR(x2, x6, x10, x14)
            BK_CHACHA_QR(x3, x7, x11, x15)
            BK_CHACHA_QR(x0, x5, x10, x15)
            BK_CHACHA_QR(x1, x6, x11, x12)
            BK_CHACHA_QR(x2, x7, x8, x13)
            BK_CHACHA_QR(x3, x4, x9, x14)
        }
        const uint32_t v[16] = { x0 + 0x61707865u, x1 + 0x3320646eu, x2 + 0x79622d32u, x3 + 0x6b206574u,
                                 x4 + key.k[0],    x5 + key.k[1],    x6 + key.k[2],    x7 + key.k[3],
                                 x8 + key.k[4],    x9 + key.k[5],    x10 + key.k[6],   x11 + key.k[7],
                                 x12 + c0,         x13 + c1,         x14 + n0,         x15 + n1 };
#pragma unroll
        for (int i = 0; i < WORDS; i++)
            out[i] = v[i];
    }
#undef BK_CHACHA_QR
    template <int WORDS>
    __host__ __device__ __forceinline__ void chacha20_block(const RngKey &key, uint32_t c0, uint32_t c1, uint32_t n0, uint32_t n1,
                                                            uint32_t *out)
    {
        chacha_block<WORDS, 10>(key, c0, c1, n0, n1, out);
    }

    // the 256-bit key of one sampling call: ChaCha20(master, counter = seed)
    inline RngKey derive_call_key(const RngKey &master, uint64_t seed)
    {
        RngKey k;
        chacha20_block<8>(master, (uint32_t)seed, (uint32_t)(seed >> 32), 0x6b63326bu /* "b2ck" */, 0u, k.k);
        return k;
    }
} // namespace bk
