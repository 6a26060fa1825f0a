// common/cached.h - "multiply by a slot vector that has a name".
//
// The reference encodes every plaintext operand at run time, every time (evaluator.h:1270-1278): 299 diagonals per
// bootstrap, 9q + co masks per convolution.  Those vectors are parameters of the network.  Call sites that can name
// their vector (owner object, index, variant) go through this helper; on the engine the encoded plaintext then stays
// resident in HBM and is reused when the same vector meets the same level and scale again (seal::Evaluator::
// multiply_vector_inplace_cached), on stock SEAL it is the reference's multiply_vector_inplace_reduced_error.
#pragma once
#include "seal/seal.h"
#include <cstdint>
#include <vector>

template <class Make>
inline void multiply_vector_named(seal::Evaluator &evaluator, seal::Ciphertext &encrypted, const void *owner,
                                  std::uint64_t index, std::uint64_t variant, Make &&make)
{
#ifdef B200CKKS_FACADE
    evaluator.multiply_vector_inplace_cached(encrypted, owner, index, variant, make);
#else
    (void)owner;
    (void)index;
    (void)variant;
    evaluator.multiply_vector_inplace_reduced_error(encrypted, make());
#endif
}

// sum <- sum + encrypted * (named vector); `started` tells whether sum holds a value yet.  The reference's loops read
//   multiply_vector_reduced_error(x, v, tmp); if (first) sum = tmp; else add_inplace_reduced_error(sum, tmp);
template <class Make>
inline void multiply_vector_named_accumulate(seal::Evaluator &evaluator, seal::Ciphertext &sum, bool &started,
                                             const seal::Ciphertext &encrypted, const void *owner, std::uint64_t index,
                                             std::uint64_t variant, Make &&make)
{
#ifdef B200CKKS_FACADE
    if (!started)
        sum.release();
    evaluator.multiply_vector_accumulate_cached(sum, encrypted, owner, index, variant, make);
#else
    (void)owner;
    (void)index;
    (void)variant;
    seal::Ciphertext product;
    evaluator.multiply_vector_reduced_error(const_cast<seal::Ciphertext &>(encrypted), make(), product);
    if (!started)
        sum = product;
    else
        evaluator.add_inplace_reduced_error(sum, product);
#endif
    started = true;
}

// sum <- sum over t of terms[t] * (vector named (owner, indices[t], variant)); make_at(index) builds a term's slot vector.
// On the engine one pass over all operands; on stock SEAL the reference's loop
//   multiply_vector_reduced_error(x_t, v_t, tmp); if (first) sum = tmp; else add_inplace_reduced_error(sum, tmp);
template <class MakeAt>
inline void multiply_vector_named_sum(seal::Evaluator &evaluator, seal::Ciphertext &sum,
                                      const std::vector<const seal::Ciphertext *> &terms, const void *owner,
                                      const std::vector<std::uint64_t> &indices, std::uint64_t variant, MakeAt &&make_at)
{
#ifdef B200CKKS_FACADE
    evaluator.multiply_vector_sum_cached(sum, terms, owner, indices, variant, make_at);
#else
    (void)owner;
    (void)variant;
    seal::Ciphertext product;
    for (std::size_t t = 0; t < terms.size(); t++)
    {
        evaluator.multiply_vector_reduced_error(const_cast<seal::Ciphertext &>(*terms[t]), make_at(indices[t]), product);
        if (t == 0)
            sum = product;
        else
            evaluator.add_inplace_reduced_error(sum, product);
    }
#endif
}

inline void forget_named(seal::Evaluator &evaluator, const void *owner)
{
#ifdef B200CKKS_FACADE
    evaluator.forget_cached(owner);
#else
    (void)evaluator;
    (void)owner;
#endif
}
