// gpt2/test_util.cpp - see test_util.h
#include "gpt2/test_util.h"
#include <algorithm>
#include <cstdlib>
#include <iostream>
#include <numeric>
#include <random>

namespace gpt2
{
    void matrix_mul(vvec &A1, vvec &A2, vvec &A_out)
    {
        const std::size_t m = A1.size(), n = A1[0].size(), k = A2[0].size();
        for (std::size_t i = 0; i < m; i++)
            for (std::size_t c = 0; c < n; c++)
            {
                const double a = A1[i][c];
                for (std::size_t j = 0; j < k; j++)
                    A_out[i][j] += a * A2[c][j];
            }
    }

    void transpose(vvec &A, vvec &A_t)
    {
        for (std::size_t i = 0; i < A.size(); i++)
            for (std::size_t j = 0; j < A[0].size(); j++)
                A_t[j][i] = A[i][j];
    }

    void compute_exp_plain(vec &A)
    {
        for (auto &x : A)
            x = std::exp(x);
    }

    void compute_softmax_plain(vec &A, vec &out)
    {
        const double max_ele = *std::max_element(A.begin(), A.end());
        std::transform(A.begin(), A.end(), out.begin(), [max_ele](double x) { return x - max_ele; });
        const double sum = std::accumulate(out.begin(), out.end(), 0);
        for (auto &x : out)
            x /= sum;
    }

    void generate_random(vvec &v)
    {
        const char *seed = std::getenv("B200CKKS_SEED");
        std::mt19937_64 rnd(seed ? std::strtoull(seed, nullptr, 0) : std::random_device{}());
        std::uniform_real_distribution<double> distribution(-1, 1);
        for (auto &row : v)
            for (auto &x : row)
                x = distribution(rnd);
    }

    void populate_tests(std::map<int, test_entry_t> &tests)
    {
        tests[CIPHER_PLAIN_128] = { "Performs ciphertext/plaintext multiplication", "Performs projection onto Q,K,V matrices",
                                    CIPHER_PLAIN_128 };
        tests[ATTN_PROJ_ROW] = { "Attention layer projection", "Performs projection onto Q,K,V matrices", ATTN_PROJ_ROW };
        tests[QK_MATMUL] = { "QK^T matrix multiplication", "Performs QK^T in the attention layer and splits output into heads",
                             QK_MATMUL };
        tests[SV_MATMUL] = { "SV matrix multiplication", "Performs multiplication by V and concatenates heads", SV_MATMUL };
        tests[SOFTMAX] = { "Softmax", "Performs softmax opperation on a single ciphertext", SOFTMAX };
        tests[SMAX] = { "Smax", "Performs optimized softmax on a single ciphertext", SOFTMAX };
        tests[GELU] = { "Gelu", "Performs GELU activation function", GELU };
        tests[LAYERNORM] = { "LayerNorm", "Performs LayerNorm on a packing of single ciphertext", BOOTSTRAP };
        tests[BOOTSTRAP] = { "Bootstrap", "Performs a bootstrap ona single Ciphertext", BOOTSTRAP };
    }

    void print_tests(std::map<int, test_entry_t> &tests)
    {
        for (const auto &it : tests)
            std::cout << it.first << ". " << it.second.name << ": " << it.second.description << std::endl;
    }
} // namespace gpt2
