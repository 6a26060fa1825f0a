// common/func.h - small integer/slot helpers shared by the restated application layers.
//
// Same names and results as the helpers the reference's app code calls (cnn_ckks/common/func.cpp:117-140
// babycount, :203-213 giantstep, :215-224 rotation; common/MinicompFunc.cpp:11-62 pmod/pow2/ceil_to_int/
// floor_to_int/log2_long/num_one), written without NTL.  Header-only.
#pragma once
#include "seal/seal.h"
#include <cmath>
#include <cstdlib>
#include <complex>
#include <stdexcept>
#include <vector>

namespace minicomp
{
    inline long pmod(long a, long b)
    {
        long r = a % b;
        return r < 0 ? r + b : r;
    }
    inline long pow2(long n)
    {
        return n <= 0 ? 1L : (1L << n);
    }
    inline std::size_t ceil_to_int(double x)
    {
        return static_cast<std::size_t>(std::ceil(x) + 0.5);
    }
    inline int floor_to_int(double x)
    {
        return static_cast<int>(std::floor(x) + 0.5);
    }
    // exponent of a power of two <= 65536, -1 for anything else
    inline long log2_long(long n)
    {
        if (n > 65536 || n <= 0)
            throw std::out_of_range("n is too large.");
        for (int i = 0; i <= 16; i++)
            if ((1L << i) == n)
                return i;
        return -1;
    }
    // population count
    inline long num_one(long n)
    {
        long c = 0;
        for (; n > 0; n >>= 1)
            c += n & 1;
        return c;
    }
} // namespace minicomp

// The reference encrypts a handful of constants inside its hot paths (the zero accumulator of a convolution,
// cnn_seal.cpp:433-436; the batch-norm shift, :568-569; the 1/2 of the ReLU, SEALcomp.cpp:54-55; T0 = 1 of the Chebyshev
// basis, SEALfunc.cpp:43-45) and lets the reduced-error add walk those fresh top-level ciphertexts down to the working
// level.  By default they are used as the plaintext constants they are (same levels, scales and values, less noise,
// no encryption and no walk-down: SURVEY.md 8(f) rank 1); $B200CKKS_ENCRYPT_CONSTANTS=1 restores the reference's
// call sequence.
inline bool encrypt_constants()
{
    static const bool on = std::getenv("B200CKKS_ENCRYPT_CONSTANTS") != nullptr;
    return on;
}

// The reference builds every leaf of a polynomial evaluation tree term by term: multiply_const, rescale_to_next and a
// reduced-error add that walks the higher-level operand down (common/Polynomial.cpp:438-456, comp/SEALfunc.cpp) -
// ten rescales per degree-7 leaf.  By default (engine only) a leaf is one pass over its terms with the scalars encoded
// so that all terms meet at one scale, and ONE rescale (Evaluator::scalar_linear_combination); same level out, same
// value up to rounding.  $B200CKKS_TERMWISE_LEAVES=1 (or the reference's constants, above) restores the reference's
// sequence.
inline bool fused_leaves()
{
#ifdef B200CKKS_FACADE
    static const bool off = std::getenv("B200CKKS_TERMWISE_LEAVES") != nullptr || encrypt_constants();
    return !off;
#else
    return false;
#endif
}

// A ciphertext product of the reference ends with relinearize_inplace and, after whatever is added to it at the
// product's scale, rescale_to_next_inplace (common/Polynomial.cpp:242-252, comp/SEALfunc.cpp:18-24,
// ckks_bootstrapping/ModularReducer.cpp:43-46).  On the engine the pair is one call
// (Evaluator::relinearize_rescale_inplace: one division by q_last P_S in tolerance mode, the two calls otherwise), and
// in tolerance mode the additions move in front of the relinearization, which is linear: the product stays of size 3,
// the linear combination runs on three polynomials, and the sum is relinearized and rescaled once.
// $B200CKKS_MERGED_RESCALE=0 keeps the separate calls in the reference's order.
inline bool merged_rescale()
{
#ifdef B200CKKS_FACADE
    static const bool on = [] {
        const char *e = std::getenv("B200CKKS_MERGED_RESCALE");
        return fused_leaves() && (!e || std::atoi(e) != 0);
    }();
    return on;
#else
    return false;
#endif
}
// The rescale after a double-hoisted linear transform can move in front of the giant-step rotations (it commutes with
// them and with the sum), where it is part of each inner sum's division by the special modulus
// (Evaluator::bsgs_inner_sums_cached, rescale = true); the giant steps then run one level lower.  Measured at logn = 14:
// forward transforms per bootstrap 13054 -> 11742, 30.5 -> 29.3 ms - but the key-switching noise of the giant steps is
// then added at the rescaled scale instead of being divided by q_last with everything else, and at the two levels
// below the top that noise is SEAL's own (digits as wide as the special modulus): logit error of ResNet-20 3e-4 ->
// 1e-3, bootstrap error max 1.7e-5 -> 2.9e-5.  Off by default; $B200CKKS_EARLY_RESCALE=1 turns it on (the committed key
// plan lists the giant-step keys one level higher: use lazy keys or a plan from a dry run with the switch set).
inline bool early_rescale()
{
#ifdef B200CKKS_FACADE
    static const bool on = [] {
        const char *e = std::getenv("B200CKKS_EARLY_RESCALE");
        return merged_rescale() && e && std::atoi(e) != 0;
    }();
    return on;
#else
    return false;
#endif
}
template <class Evaluator, class Ciphertext, class RelinKeys>
inline void relinearize_then_rescale(Evaluator &evaluator, Ciphertext &cipher, RelinKeys &relin_keys)
{
#ifdef B200CKKS_FACADE
    evaluator.relinearize_rescale_inplace(cipher, relin_keys);
#else
    evaluator.relinearize_inplace(cipher, relin_keys);
    evaluator.rescale_to_next_inplace(cipher);
#endif
}

// baby-step size minimising ceil(M/k) + k - 1 (first minimiser, k <= 3 sqrt(M))
inline int giantstep(int M)
{
    int best = M, arg = 1;
    for (int k = 1; k <= 3 * std::sqrt((double)M); k++)
    {
        int cost = (M + k - 1) / k + k - 1;
        if (cost < best)
        {
            best = cost;
            arg = k;
        }
    }
    return arg;
}

// out = the 2^logslot-periodic vector `vec` rotated left by shiftcount, tiled to Nh slots
inline void rotation(int logslot, int Nh, int shiftcount, const std::vector<std::complex<double>> &vec,
                     std::vector<std::complex<double>> &rtnvec)
{
    const int slotlen = 1 << logslot;
    const int repeat = Nh / slotlen;
    rtnvec.resize((std::size_t)repeat * slotlen);
    int s = shiftcount % slotlen;
    if (s < 0)
        s += slotlen;
    for (int i = 0; i < slotlen; i++)
    {
        std::complex<double> v = vec[(std::size_t)((i + s) % slotlen)];
        for (int j = 0; j < repeat; j++)
            rtnvec[(std::size_t)j * slotlen + i] = v;
    }
}

// Paterson-Stockmeyer split for a degree-`deg` Chebyshev series: k babies, 2^m giants
inline void babycount(long &mink, long &minm, long deg)
{
    auto cost = [deg](long k, long &m) {
        double dk = static_cast<double>(deg) / k;
        m = (long)std::ceil(std::log2(dk));
        return m + k + (long)std::ceil(dk) - 3;
    };
    mink = 2;
    long best = cost(2, minm);
    for (long k = 3; k < 2 * std::sqrt((double)deg); k++)
    {
        long m, c = cost(k, m);
        if (c < best)
        {
            best = c;
            mink = k;
            minm = m;
        }
    }
}
