// comp/program.h - choice of the evaluation tree for an odd polynomial of degree n.
//
// upgrade_oddbaby follows the dynamic program of the reference (cnn_ckks/cpu-ckks/single-key/comp/
// program.cpp:3-60): over all (l, m) it minimises the number of non-scalar multiplications of an
// odd-baby-step / giant-step evaluation that uses the minimum depth ceil(log2 n), breaking ties exactly as
// the reference does (strict "<", increasing l then m, increasing k), because the tree-decomposed
// coefficient tables (comp/minimax_relu_alpha13.inc) are laid out for precisely these trees.
#pragma once
#include "comp/PolyUpdate.h"

void upgrade_oddbaby(long n, minicomp::Tree &tree);
