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
