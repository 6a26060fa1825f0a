// ckks_bootstrapping/LinearTransform.h - the special-FFT factor matrices behind CoeffToSlot / SlotToCoeff.
//
// The CKKS decoding map on n slots factors into log2(n) butterfly stages; each stage is a matrix with three
// non-zero (cyclic) diagonals.  The reference tabulates the stages (Bootstrapper::genorigcoeff,
// ckks_bootstrapping/Bootstrapper.cpp:512-592) and then merges groups of stages into three matrices per direction
// by enumerating all 3^k diagonal choices (genfftcoeff_3 :1116-1385, geninvfftcoeff_3 :1516-1787).  Here the same
// matrices are obtained by multiplying the stages as sparse-diagonal matrices:
//     (A B)_d [k] = sum_{a + b = d} A_a[k] * B_b[(k + a) mod n],     (M x)[k] = sum_d M_d[k] x[(k + d) mod n],
// which is the product the reference's enumeration spells out term by term.  Offsets are kept signed; where the
// reference indexes a merged matrix modulo its period ("rotated" transforms) use folded().
#pragma once
#include <cmath>
#include <complex>
#include <map>
#include <stdexcept>
#include <vector>

namespace boot
{
    using cplx = std::complex<double>;

    struct DiagMatrix
    {
        long n = 0;
        // signed offset d -> n entries; entry k is non-zero only when 0 <= k + d < n (no wrap-around), so the same
        // table also describes the operator on 2n-periodic slot vectors that acts on both halves alike
        std::map<long, std::vector<cplx>> diag;

        explicit DiagMatrix(long n_ = 0) : n(n_)
        {}
        static DiagMatrix identity(long n)
        {
            DiagMatrix m(n);
            m.diag[0].assign((std::size_t)n, cplx(1.0, 0.0));
            return m;
        }
        std::vector<cplx> &at(long offset)
        {
            auto &v = diag[offset];
            if (v.empty())
                v.assign((std::size_t)n, cplx(0.0, 0.0));
            return v;
        }
        // diagonal at a (signed) offset; all-zero if the matrix has none there
        std::vector<cplx> get(long offset) const
        {
            auto it = diag.find(offset);
            return it == diag.end() ? std::vector<cplx>((std::size_t)n, cplx(0.0, 0.0)) : it->second;
        }
        // this * first  (apply `first`, then this)
        DiagMatrix after(const DiagMatrix &first) const
        {
            if (first.n != n)
                throw std::invalid_argument("dimension mismatch");
            DiagMatrix out(n);
            for (const auto &a : diag)
                for (const auto &b : first.diag)
                {
                    auto &d = out.at(a.first + b.first);
                    for (long k = 0; k < n; k++)
                        d[(std::size_t)k] += a.second[(std::size_t)k] * b.second[(std::size_t)((((k + a.first) % n) + n) % n)];
                }
            return out;
        }
        // offsets taken modulo `period` into [0, period): diagonals that coincide as cyclic rotations are added
        DiagMatrix folded(long period) const
        {
            DiagMatrix out(n);
            for (const auto &d : diag)
            {
                auto &v = out.at(((d.first % period) + period) % period);
                for (long k = 0; k < n; k++)
                    v[(std::size_t)k] += d.second[(std::size_t)k];
            }
            return out;
        }
        void scale(double f)
        {
            for (auto &d : diag)
                for (auto &v : d.second)
                    v *= f;
        }
    };

    // 5^j mod 2^bits, the order in which the special FFT visits the roots of unity
    inline std::vector<long> five_powers(long count, long bits)
    {
        std::vector<long> p((std::size_t)count);
        long v = 1;
        for (long j = 0; j < count; j++)
        {
            p[(std::size_t)j] = v;
            v = (5 * v) % (1L << bits);
        }
        return p;
    }

    // stage i (0-based, block length 2^(i+1)) of the slot -> coefficient direction:
    //   out[j] = in[j] + z in[j + h],   out[j + h] = in[j] - z in[j + h],   h = 2^i, z = exp(i pi 5^j 2^(logn-1-i) / (2n))
    inline DiagMatrix decode_stage(long logn, long i)
    {
        const long n = 1L << logn, block = 2L << i, half = block / 2;
        const double theta = M_PI / (2.0 * n) * (double)(1L << (logn - 1 - i));
        auto pw = five_powers(half, i + 3);
        DiagMatrix m(n);
        auto &lo = m.at(-half), &mid = m.at(0), &hi = m.at(half);
        for (long j = 0; j < half; j++)
        {
            cplx z = std::polar(1.0, theta * (double)pw[(std::size_t)j]);
            for (long b = 0; b < n; b += block)
            {
                mid[(std::size_t)(b + j)] = 1.0;
                mid[(std::size_t)(b + j + half)] = -z;
                lo[(std::size_t)(b + j + half)] = 1.0;
                hi[(std::size_t)(b + j)] = z;
            }
        }
        return m;
    }

    // stage i of the coefficient -> slot direction (block length n / 2^i), each stage carrying a factor 1/2:
    //   out[j] = (in[j] + in[j + h]) / 2,   out[j + h] = z (in[j] - in[j + h]) / 2,   z = exp(-i pi 5^j 2^i / (2n))
    inline DiagMatrix encode_stage(long logn, long i)
    {
        const long n = 1L << logn, block = n >> i, half = block / 2;
        const double theta = -M_PI / (2.0 * n) * (double)(1L << i);
        auto pw = five_powers(half, (logn - 1 - i) + 3);
        DiagMatrix m(n);
        auto &lo = m.at(-half), &mid = m.at(0), &hi = m.at(half);
        for (long j = 0; j < half; j++)
        {
            cplx z = std::polar(1.0, theta * (double)pw[(std::size_t)j]);
            for (long b = 0; b < n; b += block)
            {
                mid[(std::size_t)(b + j)] = 0.5;
                mid[(std::size_t)(b + j + half)] = -0.5 * z;
                lo[(std::size_t)(b + j + half)] = 0.5 * z;
                hi[(std::size_t)(b + j)] = 0.5;
            }
        }
        return m;
    }
} // namespace boot
