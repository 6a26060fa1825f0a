// comp/SEALfunc.cpp - see SEALfunc.h.  Every ciphertext operation goes through seal::Evaluator, i.e. through
// the C ABI into the CUDA engine; the code below only decides which operation comes next.
#include "comp/SEALfunc.h"
#include <algorithm>
#include <cmath>
#include <map>
#include <memory>

namespace seal
{
    using minicomp::ceil_to_int;
    using minicomp::num_one;
    using minicomp::pow2;

    namespace
    {
        // degree of the polynomial sitting at every heap node: the remainder at an even node has degree g - 1,
        // the quotient at an odd node has degree (parent degree - g)
        std::vector<long> node_degrees(long deg, const Tree &tree)
        {
            std::vector<long> d((std::size_t)pow2(tree.depth + 1), -1);
            d[1] = deg;
            for (long j = 2; j < (long)d.size(); j++)
            {
                long g = tree.tree[(std::size_t)(j / 2)];
                d[(std::size_t)j] = (j % 2 == 0) ? g - 1 : d[(std::size_t)(j / 2)] - g;
            }
            return d;
        }
    } // namespace

    long coeff_number(long deg, Tree &tree)
    {
        auto d = node_degrees(deg, tree);
        long num = 0;
        for (std::size_t i = 0; i < d.size(); i++)
            if (tree.tree[i] == 0)
                num += d[i] + 1;
        return num;
    }

    void evalT(Evaluator &evaluator, PublicKey &, SecretKey &, RelinKeys &relin_keys, Ciphertext &Tmplusn,
               const Ciphertext &Tm, const Ciphertext &Tn, const Ciphertext &Tmminusn)
    {
        Ciphertext twice;
#ifdef B200CKKS_FACADE
        if (merged_rescale() && (Tmminusn.size() == 0 ||
                                 Tmminusn.coeff_modulus_size() >= std::min(Tm.coeff_modulus_size(), Tn.coeff_modulus_size())))
        {
            // 2 Tm Tn - T(m-n) on the unrelinearized product, then one relinearization and rescale (func.h); T0 = 1 is
            // the constant it is (geneT0T1) and joins at the product's scale
            evaluator.multiply_reduced_error_unrelinearized(Tm, Tn, twice);
            if (Tmminusn.size() == 0)
                evaluator.scalar_linear_combination({ &twice }, { 2.0 }, -1.0, twice.scale(), Tmplusn);
            else
                evaluator.scalar_linear_combination({ &twice, &Tmminusn }, { 2.0, -1.0 }, 0.0, twice.scale(), Tmplusn);
            evaluator.relinearize_rescale_inplace(Tmplusn, relin_keys);
            return;
        }
#endif
        evaluator.multiply_reduced_error(Tm, Tn, relin_keys, twice);
#ifdef B200CKKS_FACADE
        if (fused_leaves() && Tmminusn.size() != 0 && Tmminusn.coeff_modulus_size() >= twice.coeff_modulus_size())
        {
            // 2 Tm Tn - T(m-n) before the rescale: the subtrahend joins at the product's scale (one pass) instead
            // of being walked down with its own multiply_const + rescale afterwards
            evaluator.scalar_linear_combination({ &twice, &Tmminusn }, { 2.0, -1.0 }, 0.0, twice.scale(), Tmplusn);
            evaluator.rescale_to_next_inplace(Tmplusn);
            return;
        }
#endif
        evaluator.add_inplace_reduced_error(twice, twice);
        evaluator.rescale_to_next_inplace(twice);
        if (Tmminusn.size() == 0) // T0 = 1 kept as the constant it is (geneT0T1)
            evaluator.add_const(twice, -1.0, Tmplusn);
        else
            evaluator.sub_reduced_error(twice, Tmminusn, Tmplusn);
    }

    void geneT0T1(Encryptor &encryptor, Evaluator &, CKKSEncoder &encoder, PublicKey &, SecretKey &, RelinKeys &,
                  Ciphertext &T0, Ciphertext &T1, Ciphertext &cipher)
    {
        if (encrypt_constants())
        {
            std::vector<double> ones(cipher.poly_modulus_degree() / 2, 1.0);
            Plaintext plain;
            encoder.encode(ones, cipher.scale(), plain);
            encryptor.encrypt(plain, T0);
        }
        else
            T0 = Ciphertext(); // empty: evalT subtracts the constant 1 instead
        T1 = cipher;
    }

    void eval_polynomial_integrate(Encryptor &encryptor, Evaluator &evaluator, Decryptor &, CKKSEncoder &encoder,
                                   PublicKey &public_key, SecretKey &secret_key, RelinKeys &relin_keys, Ciphertext &res,
                                   Ciphertext &cipher, long deg, const std::vector<double> &decomp_coeff, Tree &tree)
    {
        if (tree.type != evaltype::oddbaby)
            throw std::invalid_argument("only odd-baby evaluation trees are supported");
        const long depth_needed = (long)ceil_to_int(std::log(static_cast<double>(deg + 1)) / std::log(2.0));
        const long nodes = pow2(tree.depth + 1);
        const std::vector<long> degree = node_degrees(deg, tree);

        // coefficient offset of every leaf (leaves in heap order; slot 0 of the table is unused for odd trees)
        std::vector<long> first_coeff((std::size_t)nodes, -1);
        long cursor = 1;
        for (long j = 1; j < nodes; j++)
            if (tree.tree[(std::size_t)j] == 0)
            {
                first_coeff[(std::size_t)j] = cursor;
                cursor += degree[(std::size_t)j] + 1;
            }
        // (the reference also encrypts an all-zero vector at scale^2 here, SEALfunc.cpp:76-80; it is only read by
        //  the plain-baby branch, so the odd-baby evaluation never touches it)

        std::map<long, Ciphertext> T, part; // Chebyshev basis T_k(x) and partial polynomials per node
        geneT0T1(encryptor, evaluator, encoder, public_key, secret_key, relin_keys, T[0], T[1], cipher);

        auto stage_of = [&](long j) { return depth_needed + 1 - num_one(j); };
        auto basis = [&](long k) -> Ciphertext & {
            auto it = T.find(k);
            if (it == T.end())
                throw std::runtime_error("Chebyshev basis element is not available at this stage");
            return it->second;
        };
        Ciphertext term;

        for (long stage = 1; stage <= depth_needed; stage++)
        {
            // leaves whose result is due at this depth: sum_k c_k T_k over odd k, one (lazy) rescale
            for (long j = 1; j < nodes; j++)
            {
                if (tree.tree[(std::size_t)j] != 0 || stage_of(j) != stage)
                    continue;
                long idx = first_coeff[(std::size_t)j];
                Ciphertext &acc = part[j];
#ifdef B200CKKS_FACADE
                if (fused_leaves())
                {
                    // one pass over the odd basis elements at the level of the lowest one (the reference walks the
                    // accumulator down with one multiply_const + rescale per level it meets)
                    std::vector<const Ciphertext *> terms;
                    std::vector<double> values;
                    for (long k = 1; k <= degree[(std::size_t)j]; k += 2, idx += 2)
                    {
                        terms.push_back(&basis(k));
                        values.push_back(decomp_coeff[(std::size_t)idx]);
                    }
                    const Ciphertext *lowest = terms[0];
                    for (const Ciphertext *t : terms)
                        if (t->coeff_modulus_size() < lowest->coeff_modulus_size())
                            lowest = t;
                    evaluator.scalar_linear_combination(terms, values, 0.0, lowest->scale() * lowest->scale(), acc);
                    evaluator.rescale_to_next_inplace(acc);
                    continue;
                }
#endif
                evaluator.multiply_const(basis(1), decomp_coeff[(std::size_t)idx], acc);
                idx += 2;
                for (long k = 3; k <= degree[(std::size_t)j]; k += 2, idx += 2)
                {
                    evaluator.multiply_const(basis(k), decomp_coeff[(std::size_t)idx], term);
                    evaluator.add_inplace_reduced_error(acc, term);
                }
                evaluator.rescale_to_next_inplace(acc);
            }
            // inner nodes due at this depth: walk down the chain of remainders,
            //   p_j = T_g1 q_1 + T_g2 q_2 + ... + r   with a single rescale for the whole sum of products
            for (long j = 1; j < nodes; j += 2)
            {
                if (tree.tree[(std::size_t)j] <= 0 || stage_of(j) != stage)
                    continue;
                long k = j;
                Ciphertext &acc = part[j];
                bool folded = false;
#ifdef B200CKKS_FACADE
                if (merged_rescale())
                {
                    // every product of the chain stays of size 3; if they (and the remainder) meet at one level, the
                    // whole sum is ONE pass over three polynomials, ONE relinearization and ONE rescale
                    std::vector<Ciphertext> products;
                    products.emplace_back();
                    evaluator.multiply_reduced_error_unrelinearized(basis(tree.tree[(std::size_t)k]), part.at(2 * k + 1), products.back());
                    for (k *= 2; tree.tree[(std::size_t)k] != 0; k *= 2)
                    {
                        products.emplace_back();
                        evaluator.multiply_reduced_error_unrelinearized(basis(tree.tree[(std::size_t)k]), part.at(2 * k + 1),
                                                                        products.back());
                    }
                    bool one_level = products.size() + 1 <= 8 &&
                                     part.at(k).coeff_modulus_size() >= products[0].coeff_modulus_size();
                    for (const Ciphertext &p : products)
                        one_level = one_level && p.coeff_modulus_size() == products[0].coeff_modulus_size();
                    if (one_level)
                    {
                        std::vector<const Ciphertext *> terms;
                        for (const Ciphertext &p : products)
                            terms.push_back(&p);
                        terms.push_back(&part.at(k));
                        Ciphertext sum;
                        evaluator.scalar_linear_combination(terms, std::vector<double>(terms.size(), 1.0), 0.0,
                                                            products[0].scale(), sum);
                        evaluator.relinearize_rescale_inplace(sum, relin_keys);
                        acc = std::move(sum);
                        continue;
                    }
                    // levels differ: each product is relinearized and they are folded as the reference folds them
                    for (Ciphertext &p : products)
                        evaluator.relinearize_inplace(p, relin_keys);
                    acc = std::move(products[0]);
                    for (std::size_t i = 1; i < products.size(); i++)
                        evaluator.add_inplace_reduced_error(acc, products[i]);
                    folded = true;
                }
#endif
                if (!folded)
                {
                    evaluator.multiply_reduced_error(basis(tree.tree[(std::size_t)k]), part.at(2 * k + 1), relin_keys, acc);
                    for (k *= 2; tree.tree[(std::size_t)k] != 0; k *= 2)
                    {
                        evaluator.multiply_reduced_error(basis(tree.tree[(std::size_t)k]), part.at(2 * k + 1), relin_keys, term);
                        evaluator.add_inplace_reduced_error(acc, term);
                    }
                }
#ifdef B200CKKS_FACADE
                if (fused_leaves() && part.at(k).coeff_modulus_size() >= acc.coeff_modulus_size())
                {
                    // the remainder joins the sum of products before its rescale
                    Ciphertext sum;
                    evaluator.scalar_linear_combination({ &acc, &part.at(k) }, { 1.0, 1.0 }, 0.0, acc.scale(), sum);
                    evaluator.rescale_to_next_inplace(sum);
                    acc = std::move(sum);
                    continue;
                }
#endif
                evaluator.rescale_to_next_inplace(acc);
                evaluator.add_inplace_reduced_error(acc, part.at(k));
            }
            // next power-of-two giant and the odd babies of this depth
            if (stage <= tree.m - 1)
                evalT(evaluator, public_key, secret_key, relin_keys, T[pow2(stage)], basis(pow2(stage - 1)),
                      basis(pow2(stage - 1)), basis(0));
            if (stage <= tree.l)
                for (long j = pow2(stage - 1) + 1; j <= pow2(stage) - 1; j += 2)
                    evalT(evaluator, public_key, secret_key, relin_keys, T[j], basis(pow2(stage - 1)),
                          basis(j - pow2(stage - 1)), basis(pow2(stage) - j));
        }
        res = part.at(1);
    }

    long ShowFailure_ReLU(Decryptor &decryptor, CKKSEncoder &encoder, Ciphertext &cipher, std::vector<double> &x,
                          long precision, long n)
    {
        Plaintext plain;
        std::vector<double> out;
        decryptor.decrypt(cipher, plain);
        encoder.decode(plain, out);
        const double bound = std::pow(2.0, static_cast<double>(-precision));
        long failure = 0;
        for (long i = 0; i < n; i++)
        {
            double relu = x[(std::size_t)i] > 0 ? x[(std::size_t)i] : 0.0;
            if (std::fabs(relu - out[(std::size_t)i]) > bound)
                failure++;
        }
        return failure;
    }
} // namespace seal
