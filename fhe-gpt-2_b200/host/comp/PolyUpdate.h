// comp/PolyUpdate.h - the evaluation tree of a Chebyshev-basis polynomial ("T-tree").
//
// Same meaning as minicomp::Tree in the reference (cnn_ckks/common/PolyUpdate.h:29-45, PolyUpdate.cpp:9-145):
// a binary heap, node k holds g > 0 when the polynomial at k is split as  p_k = T_g * p_{2k+1} + p_{2k},
// 0 for a leaf that is evaluated directly from the baby Chebyshev polynomials, -1 for "no node".
// Only the tree container and its merge are needed on the evaluation path; the reference's RR polynomial
// tooling in the same header (decomposition of new polynomials) is setup-time and out of scope.
#pragma once
#include "common/func.h"
#include <stdexcept>
#include <vector>

enum class evaltype : int
{
    none = 0,
    oddbaby = 1, // odd polynomial, odd baby steps, optimal level consumption
    baby = 2
};

namespace minicomp
{
    class Tree
    {
    public:
        int depth = 0;
        evaltype type = evaltype::none;
        std::vector<int> tree{ -1, 0 };
        int m = 0; // giants T_2, T_4, ..., T_{2^(m-1)}
        int l = 0; // odd babies T_1, T_3, ..., T_{2^l - 1}
        int b = 0;

        Tree() = default;
        explicit Tree(evaltype ty) : type(ty)
        {}
        Tree(const Tree &a, const Tree &b_, int g)
        {
            merge(a, b_, g);
        }
        void clear()
        {
            depth = 0;
            type = evaltype::none;
            tree.assign({ -1, 0 });
        }
        // root g, `rem` (degree < g) as the left subtree, `quo` as the right subtree
        void merge(const Tree &rem, const Tree &quo, int g)
        {
            if (rem.type != quo.type)
                throw std::invalid_argument("the types of two trees are not the same");
            Tree out(rem.type);
            out.depth = (rem.depth > quo.depth ? rem.depth : quo.depth) + 1;
            out.tree.assign((std::size_t)pow2(out.depth + 1), -1);
            out.tree[1] = g;
            auto graft = [&out](const Tree &sub, int side) {
                for (long i = 1; i < pow2(sub.depth + 1); i++)
                {
                    long row = 1;
                    while (row * 2 <= i)
                        row *= 2; // first index of i's heap row
                    out.tree[(std::size_t)(i + (side + 1) * row)] = sub.tree[(std::size_t)i];
                }
            };
            graft(rem, 0);
            graft(quo, 1);
            out.m = m;
            out.l = l;
            out.b = b;
            *this = out;
        }
        // number of non-scalar multiplications the tree costs
        int nonscalar() const
        {
            int n = type == evaltype::oddbaby ? m - 1 + (int)pow2(l - 1) - 1 : m + b - 2;
            for (int v : tree)
                if (v > 0)
                    n++;
            return n;
        }
    };
} // namespace minicomp
