"""Algebra of level-aware hybrid key switching on a toy ring (oracle/hybrid_keyswitch_model.py, big integers): for every
shape the engine can pick - alpha special moduli, digits of dsize primes - the switched pair must satisfy
r0 + r1 s = c s' + (small error); and the engine's shape rule (hybrid_shape in engine.cu) restated here must stay
inside the constraints the model needs."""
import hybrid_keyswitch_model as hk


def test_key_switch_identity_for_all_shapes():
    toy = hk.Toy(L=6, seed=3)
    for l in range(1, 7):
        for alpha in range(1, 6 - l + 2):
            for dsize in sorted({1, max(1, alpha - 1), alpha}):
                err = toy.error(l, alpha, dsize)
                # error = sum_d ext_d e_d / P_S: with dsize < alpha it is ~N * |e| * dsize / (a 21-bit prime), i.e. single
                # digits; with dsize == alpha the digit is as large as P_S and the error is ~N * |e| * dsize
                bound = 64 if dsize < alpha else 64 * 16 * max(1, dsize)
                assert err < bound, (l, alpha, dsize, err)
                assert err < min(toy.q) // 1000


def hybrid_shape(l, top):
    """engine.cu: hybrid_shape"""
    alpha, dsize = 1, 1
    if l <= 5:
        return alpha, dsize
    best = l + l * (l + 1) - l + l * l / 8.0 + 2.0 + 2.0 * l / 8.0 + 2.0 * l
    for a in range(2, min(top - l + 1, 17) + 1):
        for ds in range(a - 1, (a if l >= top - 2 else a - 1) + 1):      # wide digits only at the two levels below the top
            d = (l + ds - 1) // ds
            cost = l + d * (l + a) - l + d * ds * l / 8.0 + 2.0 * a + 2.0 * a * l / 8.0 + 2.0 * l
            if cost < best - 1e-9:
                best, alpha, dsize = cost, a, ds
    return alpha, dsize


def test_shape_rule_respects_the_chain():
    for top in (31, 36):
        for l in range(1, top + 1):
            alpha, dsize = hybrid_shape(l, top)
            assert l + alpha - 1 <= top            # only idle primes are borrowed
            assert 1 <= alpha <= 17 and (dsize == max(1, alpha - 1) or (dsize == alpha and l >= top - 2))
        assert hybrid_shape(top, top) == (1, 1)    # nothing is idle at the top level: SEAL's own scheme
    assert hybrid_shape(20, 31)[0] > 2 and hybrid_shape(3, 31) == (1, 1)


def test_rounding_mod_down_and_merged_rescale_are_exact():
    """The hybrid ModDown as the engine computes it since the merged relinearize + rescale (rounding, exact conversion:
    engine.cu hyb_mod_down / hyb_mod_down_rescale, kernels.cuh k_hyb_conv<DS, true>) against exact big-integer
    arithmetic on the toy ring: round(x / P_S) limb by limb, and relinearization + rescale as ONE division
    round((acc + P_S base) / (q_{l-1} P_S)), which equals the two roundings in sequence except at ties."""
    import random
    from math import prod

    toy = hk.Toy(L=6, seed=5)
    rng = random.Random(9)
    for l in range(2, 7):
        for alpha in range(1, 6 - l + 2):
            S = [toy.q[l + i] for i in range(alpha - 1)] + [toy.P]
            E = toy.q[:l] + S
            PS, QP = prod(S), prod(E)
            x = [rng.randrange(QP) for _ in range(hk.N)]
            acc = [[v % m for v in x] for m in E]
            base_int = [rng.randrange(prod(toy.q[:l])) for _ in range(hk.N)]
            base = [[v % m for v in base_int] for m in toy.q[:l]]
            # (1) ModDown alone: round(x / P_S)
            got = hk.mod_down_rounded(acc, E, l, S)
            want = [(2 * v + PS) // (2 * PS) for v in x]
            assert got == [[w % m for w in want] for m in toy.q[:l]], (l, alpha)
            # (2) merged: round((x + P_S base) / (q_{l-1} P_S)) on l - 1 limbs
            D = toy.q[l - 1] * PS
            total = [(v + PS * b) % QP for v, b in zip(x, base_int)]
            got = hk.mod_down_rescale(acc, base, E, l, S)
            want = [(2 * v + D) // (2 * D) for v in total]
            assert got == [[w % m for w in want] for m in toy.q[: l - 1]], (l, alpha)
            # (3) the two calls it replaces: ModDown (rounded), add the base, rescale (rounded) - equal except at ties
            md = hk.crt(hk.mod_down_rounded(acc, E, l, S), toy.q[:l])
            ql = toy.q[l - 1]
            two = [((2 * ((m + b) % prod(toy.q[:l])) + ql) // (2 * ql)) for m, b in zip(md, base_int)]
            Ql1 = prod(toy.q[: l - 1])
            diff = [min((a - b) % Ql1, (b - a) % Ql1) for a, b in zip(two, want)]
            assert max(diff) <= 1 and sum(diff) <= 1, (l, alpha, diff)
