// comp/program.cpp - see program.h.
#include "comp/program.h"
#include <cmath>

using minicomp::pow2;
using minicomp::Tree;

namespace
{
    constexpr int kInf = 10000;

    // cost[i][j] = fewest tree multiplications for an odd polynomial of degree i within depth j, given odd
    // babies up to 2^l - 1 and giants T_2..T_{2^(m-1)}; plan[i][j] = a tree achieving it
    struct Table
    {
        std::vector<std::vector<int>> cost;
        std::vector<std::vector<Tree>> plan;
    };

    Table solve(long n, long d, long l, long m)
    {
        Table t;
        t.cost.assign((std::size_t)n + 1, std::vector<int>((std::size_t)d + 1, 0));
        t.plan.assign((std::size_t)n + 1, std::vector<Tree>((std::size_t)d + 1, Tree(evaltype::oddbaby)));
        for (long i = 3; i <= n; i += 2)
            t.cost[(std::size_t)i][1] = kInf;
        for (long j = 2; j <= d; j++)
            for (long i = 1; i <= n; i += 2)
            {
                if (i <= pow2(l) - 1 && i <= pow2(j - 1))
                {
                    t.cost[(std::size_t)i][(std::size_t)j] = 0; // a leaf: linear combination of the babies
                    continue;
                }
                int best = kInf;
                Tree best_tree;
                for (long k = 1; k <= m - 1 && pow2(k) < i && k < j; k++)
                {
                    long g = pow2(k);
                    int c = t.cost[(std::size_t)(i - g)][(std::size_t)(j - 1)] + t.cost[(std::size_t)(g - 1)][(std::size_t)j] + 1;
                    if (c < best)
                    {
                        best = c;
                        best_tree.merge(t.plan[(std::size_t)(g - 1)][(std::size_t)j],
                                        t.plan[(std::size_t)(i - g)][(std::size_t)(j - 1)], (int)g);
                    }
                }
                t.cost[(std::size_t)i][(std::size_t)j] = best;
                t.plan[(std::size_t)i][(std::size_t)j] = best_tree;
            }
        return t;
    }
} // namespace

void upgrade_oddbaby(long n, Tree &tree)
{
    const long d = (long)minicomp::ceil_to_int(std::log(static_cast<double>(n)) / std::log(2.0));
    int total_best = kInf;
    Tree chosen;
    long best_m = 0, best_l = 0;
    for (long l = 1; pow2(l) - 1 <= n; l++)
        for (long m = 1; pow2(m - 1) < n; m++)
        {
            Table t = solve(n, d, l, m);
            int total = t.cost[(std::size_t)n][(std::size_t)d] + (int)pow2(l - 1) + (int)m - 2;
            if (total < total_best)
            {
                total_best = total;
                chosen = t.plan[(std::size_t)n][(std::size_t)d];
                best_m = m;
                best_l = l;
            }
        }
    tree = chosen;
    tree.m = (int)best_m;
    tree.l = (int)best_l;
}
