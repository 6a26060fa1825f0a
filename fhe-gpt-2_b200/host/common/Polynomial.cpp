// common/Polynomial.cpp - see Polynomial.h.
#include "common/Polynomial.h"
#include <algorithm>
#include <cmath>
#include <stdexcept>

using namespace seal;

namespace boot
{
    Polynomial::Polynomial(long _deg)
    {
        set_zero_polynomial(_deg);
    }
    Polynomial::Polynomial(long _deg, const long double *_coeff, const std::string &tag)
    {
        set_polynomial(_deg, _coeff, tag);
    }
    Polynomial::Polynomial(const Polynomial &o)
    {
        copy(o);
    }
    Polynomial &Polynomial::operator=(const Polynomial &o)
    {
        if (this != &o)
            copy(o);
        return *this;
    }
    void Polynomial::set_zero_polynomial(long _deg)
    {
        deg = _deg;
        coeff.assign((std::size_t)deg + 1, 0.0L);
        chebcoeff.assign((std::size_t)deg + 1, 0.0L);
    }
    void Polynomial::set_polynomial(long _deg, const long double *_coeff, const std::string &tag)
    {
        set_zero_polynomial(_deg);
        if (tag == "power")
        {
            coeff.assign(_coeff, _coeff + deg + 1);
            power_to_cheb();
        }
        else if (tag == "cheb")
        {
            chebcoeff.assign(_coeff, _coeff + deg + 1);
            cheb_to_power();
        }
        else
            throw std::invalid_argument("tag must be \"power\" or \"cheb\"");
    }
    void Polynomial::copy(const Polynomial &poly)
    {
        deg = poly.deg;
        coeff = poly.coeff;
        chebcoeff = poly.chebcoeff;
        heap_k = heap_m = heaplen = 0;
        poly_heap.clear();
    }

    // monomial <-> Chebyshev by the three-term recurrence.  Numerically meaningful for small degrees only; the
    // evaluation path reads `coeff` for deg <= 3 and `chebcoeff` otherwise.
    static std::vector<std::vector<long double>> cheb_monomials(long deg)
    {
        std::vector<std::vector<long double>> T((std::size_t)deg + 1, std::vector<long double>((std::size_t)deg + 1, 0.0L));
        T[0][0] = 1;
        if (deg >= 1)
            T[1][1] = 1;
        for (long i = 2; i <= deg; i++)
            for (long j = 0; j <= i; j++)
                T[(std::size_t)i][(std::size_t)j] =
                    (j ? 2 * T[(std::size_t)i - 1][(std::size_t)j - 1] : 0.0L) - T[(std::size_t)i - 2][(std::size_t)j];
        return T;
    }
    void Polynomial::cheb_to_power()
    {
        if (deg > 24)
        { // not representable; nobody reads it
            coeff.assign((std::size_t)deg + 1, 0.0L);
            return;
        }
        auto T = cheb_monomials(deg);
        coeff.assign((std::size_t)deg + 1, 0.0L);
        for (long i = 0; i <= deg; i++)
            for (long j = 0; j <= i; j++)
                coeff[(std::size_t)j] += chebcoeff[(std::size_t)i] * T[(std::size_t)i][(std::size_t)j];
    }
    void Polynomial::power_to_cheb()
    {
        if (deg > 24)
            throw std::invalid_argument("power-basis input is supported up to degree 24; pass Chebyshev coefficients");
        auto T = cheb_monomials(deg);
        std::vector<long double> rest = coeff;
        chebcoeff.assign((std::size_t)deg + 1, 0.0L);
        for (long i = deg; i >= 0; i--)
        {
            long double c = rest[(std::size_t)i] / T[(std::size_t)i][(std::size_t)i];
            chebcoeff[(std::size_t)i] = c;
            for (long j = 0; j <= i; j++)
                rest[(std::size_t)j] -= c * T[(std::size_t)i][(std::size_t)j];
        }
    }
    long double Polynomial::evaluate_cheb(long double x) const
    {
        long double b1 = 0, b2 = 0; // Clenshaw
        for (long j = deg; j >= 1; j--)
        {
            long double b0 = 2 * x * b1 - b2 + chebcoeff[(std::size_t)j];
            b2 = b1;
            b1 = b0;
        }
        return x * b1 - b2 + chebcoeff[0];
    }
    void Polynomial::constmul(long double constant)
    {
        for (auto &c : coeff)
            c *= constant;
        for (auto &c : chebcoeff)
            c *= constant;
    }

    void divide_by_chebyshev(Polynomial &quotient, Polynomial &remainder, const Polynomial &target, long g)
    {
        if (target.deg < g)
        {
            quotient.set_zero_polynomial(0);
            remainder.copy(target);
            return;
        }
        std::vector<long double> c = target.chebcoeff;
        quotient.set_zero_polynomial(target.deg - g);
        for (long j = target.deg; j >= g; j--)
        {
            long double cj = c[(std::size_t)j];
            c[(std::size_t)j] = 0;
            if (j == g)
                quotient.chebcoeff[0] += cj;
            else
            {
                quotient.chebcoeff[(std::size_t)(j - g)] += 2 * cj;
                long r = j - 2 * g;
                c[(std::size_t)(r < 0 ? -r : r)] -= cj;
            }
        }
        remainder.set_zero_polynomial(g - 1);
        for (long j = 0; j < g; j++)
            remainder.chebcoeff[(std::size_t)j] = c[(std::size_t)j];
        quotient.cheb_to_power();
        remainder.cheb_to_power();
    }

    void Polynomial::generate_poly_heap_manual(long k, long m)
    {
        heap_k = k;
        heap_m = m;
        heaplen = (1L << (m + 1)) - 1;
        poly_heap.clear();
        poly_heap.resize((std::size_t)heaplen);
        poly_heap[0] = std::make_unique<Polynomial>();
        poly_heap[0]->copy(*this);
        long g = k << m;
        for (long level = 0; level < m; level++)
        {
            g >>= 1; // this level divides by T_g
            for (long j = (1L << level) - 1; j < (1L << (level + 1)) - 1; j++)
            {
                if (!poly_heap[(std::size_t)j])
                    continue;
                std::size_t quo = (std::size_t)(2 * (j + 1) - 1), rem = (std::size_t)(2 * (j + 1));
                if (poly_heap[(std::size_t)j]->deg < g)
                { // nothing to divide: the node passes down unchanged on the remainder side
                    poly_heap[rem] = std::make_unique<Polynomial>();
                    poly_heap[rem]->copy(*poly_heap[(std::size_t)j]);
                }
                else
                {
                    poly_heap[quo] = std::make_unique<Polynomial>();
                    poly_heap[rem] = std::make_unique<Polynomial>();
                    divide_by_chebyshev(*poly_heap[quo], *poly_heap[rem], *poly_heap[(std::size_t)j], g);
                }
            }
        }
    }
    void Polynomial::generate_poly_heap()
    {
        babycount(heap_k, heap_m, deg);
        generate_poly_heap_manual(heap_k, heap_m);
    }

    void Polynomial::homomorphic_poly_evaluation(SEALContext &, CKKSEncoder &, Encryptor &, Evaluator &evaluator,
                                                 RelinKeys &relin_keys, Ciphertext &rtn, Ciphertext &cipher, Decryptor &)
    {
        const double zero = 1. / cipher.scale(); // coefficients below one unit of the scale encode to 0
        auto c = [this](long i) { return static_cast<double>(coeff[(std::size_t)i]); };
        auto live = [zero](double v) { return std::fabs(v) >= zero; };

        if (deg == 1)
        {
            evaluator.multiply_const(cipher, c(1), rtn);
            evaluator.rescale_to_next_inplace(rtn);
            evaluator.add_const(rtn, c(0), rtn);
            return;
        }
        if (deg == 2 || deg == 3)
        {
            Ciphertext squared, top;
            evaluator.square(cipher, squared);
            relinearize_then_rescale(evaluator, squared, relin_keys);
            if (deg == 2)
            {
                evaluator.multiply_const_inplace(squared, c(2));
                evaluator.rescale_to_next_inplace(squared);
                top = squared;
            }
            else
            {
                evaluator.multiply_const(cipher, c(3), top);
                evaluator.rescale_to_next_inplace(top);
                evaluator.multiply_inplace_reduced_error(top, squared, relin_keys);
                evaluator.rescale_to_next_inplace(top);
            }
            if (live(c(1)))
            {
                evaluator.multiply_const(cipher, c(1), rtn);
                evaluator.rescale_to_next_inplace(rtn);
                evaluator.add_reduced_error(rtn, top, rtn);
            }
            else
                rtn = top;
            if (deg == 3 && live(c(2)))
            {
                evaluator.multiply_const_inplace(squared, c(2));
                evaluator.rescale_to_next_inplace(squared);
                evaluator.add_reduced_error(rtn, squared, rtn);
            }
            evaluator.add_const_inplace(rtn, c(0));
            return;
        }

        if (poly_heap.empty())
            throw std::logic_error("generate_poly_heap() must be called before evaluating the polynomial");
        const long k = heap_k, m = heap_m;

        // T_{a+b} = 2 T_a T_b - T_{|a-b|}; a == b uses the cheaper square and the constant T_0 = 1
        auto chebyshev_sum = [&](const Ciphertext &Ta, const Ciphertext &Tb, const Ciphertext *Tdiff, Ciphertext &out) {
            if (!Tdiff)
            {
                evaluator.square(Ta, out);
                relinearize_then_rescale(evaluator, out, relin_keys);
                evaluator.double_inplace(out);
                evaluator.add_const(out, -1.0, out);
            }
            else
            {
#ifdef B200CKKS_FACADE
                if (merged_rescale() && Tdiff->coeff_modulus_size() >= std::min(Ta.coeff_modulus_size(), Tb.coeff_modulus_size()))
                {
                    // 2 Ta Tb - T|a-b| on the unrelinearized product, then one relinearization and rescale
                    Ciphertext product, sum;
                    evaluator.multiply_reduced_error_unrelinearized(Ta, Tb, product);
                    evaluator.scalar_linear_combination({ &product, Tdiff }, { 2.0, -1.0 }, 0.0, product.scale(), sum);
                    evaluator.relinearize_rescale_inplace(sum, relin_keys);
                    out = std::move(sum);
                    return;
                }
#endif
                evaluator.multiply_reduced_error(Ta, Tb, relin_keys, out);
#ifdef B200CKKS_FACADE
                if (fused_leaves() && Tdiff->coeff_modulus_size() >= out.coeff_modulus_size())
                {
                    // 2 Ta Tb - T|a-b| before the rescale (the subtrahend joins at the product's scale)
                    Ciphertext sum;
                    evaluator.scalar_linear_combination({ &out, Tdiff }, { 2.0, -1.0 }, 0.0, out.scale(), sum);
                    evaluator.rescale_to_next_inplace(sum);
                    out = std::move(sum);
                    return;
                }
#endif
                evaluator.rescale_to_next_inplace(out);
                evaluator.double_inplace(out);
                evaluator.sub_reduced_error(out, *Tdiff, out);
            }
        };

        // babies T_1 .. T_{k-1}: powers of two by squaring first, then the rest from the largest power of two below
        std::vector<Ciphertext> baby((std::size_t)k);
        std::vector<bool> have((std::size_t)k, false);
        baby[1] = cipher;
        have[1] = true;
        for (long i = 2; i < k; i *= 2)
        {
            chebyshev_sum(baby[(std::size_t)(i / 2)], baby[(std::size_t)(i / 2)], nullptr, baby[(std::size_t)i]);
            have[(std::size_t)i] = true;
        }
        for (long i = 1; i < k; i++)
        {
            if (have[(std::size_t)i])
                continue;
            long p = 1L << (long)std::floor(std::log((double)i) / std::log(2.0));
            long r = i - p, d = std::labs(p - r);
            chebyshev_sum(baby[(std::size_t)p], baby[(std::size_t)r], &baby[(std::size_t)d], baby[(std::size_t)i]);
            have[(std::size_t)i] = true;
        }

        // giants T_k, T_2k, ..., T_{2^(m-1) k}
        std::vector<Ciphertext> giant((std::size_t)m);
        {
            long p = 1L << ((long)std::ceil(std::log((double)k) / std::log(2.0)) - 1);
            long r = k - p, d = std::labs(p - r);
            if (r == 0)
                giant[0] = baby[(std::size_t)p];
            else if (d == 0)
                chebyshev_sum(baby[(std::size_t)p], baby[(std::size_t)p], nullptr, giant[0]);
            else
                chebyshev_sum(baby[(std::size_t)p], baby[(std::size_t)r], &baby[(std::size_t)d], giant[0]);
        }
        for (long i = 1; i < m; i++)
        {
            evaluator.square(giant[(std::size_t)(i - 1)], giant[(std::size_t)i]);
            relinearize_then_rescale(evaluator, giant[(std::size_t)i], relin_keys);
            evaluator.double_inplace(giant[(std::size_t)i]);
            evaluator.add_const_inplace(giant[(std::size_t)i], -1.0);
        }

        // leaves of the heap: linear combinations of the babies
        std::vector<Ciphertext> node((std::size_t)heaplen);
        std::vector<bool> set((std::size_t)heaplen, false);
        Ciphertext term;
        for (long i = (1L << m) - 1; i < heaplen; i++)
        {
            const Polynomial *p = poly_heap[(std::size_t)i].get();
            if (!p)
                continue;
            auto cc = [p](long j) { return static_cast<double>(p->chebcoeff[(std::size_t)j]); };
            set[(std::size_t)i] = true;
#ifdef B200CKKS_FACADE
            if (fused_leaves())
            {
                // all terms in one pass at the level of the lowest baby, one rescale for the leaf
                std::vector<const Ciphertext *> terms;
                std::vector<double> values;
                for (long j = 1; j <= p->deg; j++)
                    if (j == 1 || !(std::fabs(cc(j)) <= zero))
                    {
                        terms.push_back(j < k ? &baby[(std::size_t)j] : &giant[0]);
                        values.push_back(cc(j));
                    }
                const Ciphertext *lowest = terms[0];
                for (const Ciphertext *t : terms)
                    if (t->coeff_modulus_size() < lowest->coeff_modulus_size())
                        lowest = t;
                evaluator.scalar_linear_combination(terms, values, std::fabs(cc(1)) <= zero ? 0.0 : cc(0),
                                                    lowest->scale() * lowest->scale(), node[(std::size_t)i]);
                evaluator.rescale_to_next_inplace(node[(std::size_t)i]);
                continue;
            }
#endif
            evaluator.multiply_const(baby[1], cc(1), node[(std::size_t)i]);
            evaluator.rescale_to_next_inplace(node[(std::size_t)i]);
            if (!(std::fabs(cc(1)) <= zero))
                evaluator.add_const_inplace(node[(std::size_t)i], cc(0));
            for (long j = 2; j <= p->deg; j++)
            {
                if (std::fabs(cc(j)) <= zero)
                    continue;
                evaluator.multiply_const(j < k ? baby[(std::size_t)j] : giant[0], cc(j), term);
                evaluator.rescale_to_next_inplace(term);
                evaluator.add_reduced_error(node[(std::size_t)i], term, node[(std::size_t)i]);
            }
        }
        // fold the heap upwards: node = quotient * giant + remainder
        long gindex = 0;
        for (long depth = m - 1; depth >= 0; depth--, gindex++)
            for (long i = (1L << depth) - 1; i < (1L << (depth + 1)) - 1; i++)
            {
                if (!poly_heap[(std::size_t)i])
                    continue;
                std::size_t quo = (std::size_t)(2 * (i + 1) - 1), rem = (std::size_t)(2 * (i + 1));
                set[(std::size_t)i] = true;
                if (!set[quo])
                    node[(std::size_t)i] = node[rem];
                else
                {
#ifdef B200CKKS_FACADE
                    if (merged_rescale() && node[rem].coeff_modulus_size() >=
                                                std::min(node[quo].coeff_modulus_size(), giant[(std::size_t)gindex].coeff_modulus_size()))
                    {
                        // quotient * giant + remainder on the unrelinearized product, one relinearization and rescale
                        Ciphertext product, sum;
                        evaluator.multiply_reduced_error_unrelinearized(node[quo], giant[(std::size_t)gindex], product);
                        evaluator.scalar_linear_combination({ &product, &node[rem] }, { 1.0, 1.0 }, 0.0, product.scale(), sum);
                        evaluator.relinearize_rescale_inplace(sum, relin_keys);
                        node[(std::size_t)i] = std::move(sum);
                        continue;
                    }
#endif
                    evaluator.multiply_reduced_error(node[quo], giant[(std::size_t)gindex], relin_keys, node[(std::size_t)i]);
#ifdef B200CKKS_FACADE
                    if (fused_leaves() && node[rem].coeff_modulus_size() >= node[(std::size_t)i].coeff_modulus_size())
                    {
                        // quotient * giant + remainder before the rescale
                        Ciphertext sum;
                        evaluator.scalar_linear_combination({ &node[(std::size_t)i], &node[rem] }, { 1.0, 1.0 }, 0.0,
                                                            node[(std::size_t)i].scale(), sum);
                        evaluator.rescale_to_next_inplace(sum);
                        node[(std::size_t)i] = std::move(sum);
                        continue;
                    }
#endif
                    evaluator.rescale_to_next_inplace(node[(std::size_t)i]);
                    evaluator.add_reduced_error(node[(std::size_t)i], node[rem], node[(std::size_t)i]);
                }
            }
        rtn = node[0];
    }
} // namespace boot
