// common/Polynomial.h - boot::Polynomial: a Chebyshev series and its baby-step/giant-step evaluation on a ciphertext.
//
// Restates the evaluation half of the reference's cnn_ckks/common/Polynomial.{h,cpp}: generate_poly_heap[_manual]
// (:169-215), homomorphic_poly_evaluation (:256-509), constmul (:127-132).  The reference keeps coefficients as
// 1000-bit NTL::RR and splits the heap by power-basis long division; NTL is not available here and power-basis
// coefficients of a degree-59 Chebyshev series do not fit a double, so coefficients are `long double` and the
// heap is split directly in the Chebyshev basis (T_j = 2 T_{j-g} T_g - T_{|j-2g|}), which yields the same unique
// quotient and remainder.  Only double-rounded coefficients ever reach the evaluator, as in the reference
// (`to_double(...)`, Polynomial.cpp:260-456).
#pragma once
#include "common/func.h"
#include "seal/seal.h"
#include <memory>
#include <string>
#include <vector>

namespace boot
{
    class Polynomial
    {
    public:
        long deg = -1;
        long heap_k = 0, heap_m = 0, heaplen = 0;
        std::vector<long double> coeff;     // power basis (maintained for deg <= 3 only)
        std::vector<long double> chebcoeff; // Chebyshev basis
        std::vector<std::unique_ptr<Polynomial>> poly_heap;

        Polynomial() = default;
        explicit Polynomial(long _deg);
        Polynomial(long _deg, const long double *_coeff, const std::string &tag);
        Polynomial(const Polynomial &o);
        Polynomial &operator=(const Polynomial &o);

        void set_polynomial(long _deg, const long double *_coeff, const std::string &tag);
        void set_zero_polynomial(long _deg);
        void copy(const Polynomial &poly);
        void power_to_cheb();
        void cheb_to_power();
        long double evaluate_cheb(long double x) const;

        void constmul(long double constant);
        void generate_poly_heap_manual(long k, long m);
        void generate_poly_heap();

        void homomorphic_poly_evaluation(seal::SEALContext &context, seal::CKKSEncoder &encoder, seal::Encryptor &encryptor,
                                         seal::Evaluator &evaluator, seal::RelinKeys &relin_keys, seal::Ciphertext &rtn,
                                         seal::Ciphertext &cipher, seal::Decryptor &decryptor);
    };

    // target = quotient * T_g + remainder in the Chebyshev basis (deg remainder = g - 1)
    void divide_by_chebyshev(Polynomial &quotient, Polynomial &remainder, const Polynomial &target, long g);
} // namespace boot
