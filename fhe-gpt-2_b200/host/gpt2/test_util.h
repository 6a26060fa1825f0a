// gpt2/test_util.h - plain-value helpers of the reference's GPT-2 tests and microbenchmarks
// (gpt2_ckks/gpt2-ckks/single-key/gpt2/test_util.{h,cpp}).
#pragma once
#include "gpt2/util.h"
#include <map>

namespace gpt2
{
    enum TestType
    {
        CIPHER_PLAIN_128,
        ATTN_PROJ_ROW,
        QK_MATMUL,
        SV_MATMUL,
        SOFTMAX,
        SMAX,
        GELU,
        LAYERNORM,
        BOOTSTRAP
    };
    struct test_entry_t
    {
        std::string name, description;
        TestType type;
    };
    // A1 is m x n, A2 is n x k, A_out (m x k) is accumulated into
    void matrix_mul(vvec &A1, vvec &A2, vvec &A_out);
    // A is m x n, A_t is n x m
    void transpose(vvec &A, vvec &A_t);
    void compute_exp_plain(vec &A);
    // test_util.cpp:33-43 as written: subtract the maximum, then divide by the (integer-accumulated) sum of the
    // shifted values - the reference never exponentiates here; it only feeds printouts
    void compute_softmax_plain(vec &A, vec &out);
    // uniform in (-1, 1); seeded from B200CKKS_SEED when set (the reference seeds from std::random_device)
    void generate_random(vvec &v);
    void populate_tests(std::map<int, test_entry_t> &tests);
    void print_tests(std::map<int, test_entry_t> &tests);
} // namespace gpt2
