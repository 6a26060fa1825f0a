"""Big-integer model of level-aware hybrid key switching - TEST INFRASTRUCTURE ONLY (tests/test_hybrid_cpu.py).

At level l of a chain q_0 .. q_{L-1} with special prime P, the primes q_l .. q_{l+alpha-2} are idle; together with P they
form the temporary special modulus P_S (alpha moduli).  The l limbs are grouped into digits of `dsize` primes.  The key
of digit d is an RLWE encryption under s of  P_S * s' * [1 on the digit's own limbs, 0 elsewhere]  over the l + alpha
extended moduli.  Key switch: fast basis conversion of each digit to the other extended moduli, inner product with the
keys, ModDown by P_S (conversion of the special limbs back to q_0 .. q_{l-1}, subtract, multiply by P_S^-1).
This is what fhe-gpt-2_b200/csrc/engine.cu: key_switch_hybrid computes limb-wise; the model checks the algebra on a toy
ring (N = 16, 20-bit primes): r0 + r1 s - c s' must be small compared with q."""
import random
from math import prod

N = 16


def _is_prime(p):
    if p < 2:
        return False
    i = 2
    while i * i <= p:
        if p % i == 0:
            return False
        i += 1
    return True


def gen_primes(bits, count, avoid=()):
    ps, p = [], (1 << bits) + 1
    while len(ps) < count:
        if p % (2 * N) == 1 and _is_prime(p) and p not in avoid:
            ps.append(p)
        p += 2 * N
    return ps


def polymul(a, b, m):
    r = [0] * N
    for i, x in enumerate(a):
        if x == 0:
            continue
        for j, y in enumerate(b):
            k = i + j
            if k < N:
                r[k] = (r[k] + x * y) % m
            else:
                r[k - N] = (r[k - N] - x * y) % m
    return r


def bconv(res, src, dst):
    """fast basis conversion: residues res[i] mod src[i] -> residues mod each dst (value + u * prod(src), 0 <= u < len(src))"""
    qs = prod(src)
    y = [[(c * pow(qs // m, -1, m)) % m for c in r] for r, m in zip(res, src)]
    return [[sum(y[i][j] * ((qs // src[i]) % t) for i in range(len(src))) % t for j in range(N)] for t in dst]


def bconv_rounding(res, src, dst):
    """The engine's exact, rounding form of the conversion (HybridPlan::d_shalf / d_spinv / d_negd / d_addc,
    k_hyb_conv<DS, true>): for x given by its residues modulo the moduli `src` (product D) returns
    [x + floor(D/2)]_D - floor(D/2) modulo each dst - so that (x - that) / D = round(x / D).  floor(D/2) joins the source
    residues first; the multiple u D by which the fast conversion overshoots is found from the fractional parts
    sum_a y_a / p_a in double precision, as the kernel does, and taken off again."""
    D = prod(src)
    half = D // 2
    y = [[((c + half) * pow(D // m, -1, m)) % m for c in r] for r, m in zip(res, src)]
    out = []
    u = [int(sum(float(y[a][j]) * (1.0 / src[a]) for a in range(len(src)))) for j in range(N)]
    for t in dst:
        out.append([(sum(y[a][j] * ((D // src[a]) % t) for a in range(len(src))) - u[j] * (D % t) - half) % t for j in range(N)])
    return out


def crt(res, mods):
    """the integer in [0, prod(mods)) with the given residues, coefficient by coefficient"""
    M = prod(mods)
    return [sum(r[j] * (M // m) * pow(M // m, -1, m) for r, m in zip(res, mods)) % M for j in range(N)]


def mod_down_rounded(acc_p, E, l, S):
    """hyb_mod_down: round(x / P_S) on the limbs 0 .. l-1, x given over the extended basis E = q_0..q_{l-1} + S"""
    PS = prod(S)
    conv = bconv_rounding(acc_p[l:], S, E[:l])
    return [[((acc_p[i][j] - conv[i][j]) * pow(PS % E[i], -1, E[i])) % E[i] for j in range(N)] for i in range(l)]


def mod_down_rescale(acc_p, base_p, E, l, S):
    """hyb_mod_down_rescale (relinearization followed by a rescale as ONE division): round((acc + P_S base) / D) with
    D = q_{l-1} P_S on the limbs 0 .. l-2.  The dropped basis is q_{l-1} + S; P_S base vanishes on S, so only the dropped
    limb l-1 sees the base (LdInvDropped); out_i = (acc_i - conv_i) D^-1 + base_i q_{l-1}^-1 (StModDownT<true>)."""
    PS = prod(S)
    ql = E[l - 1]
    dropped = [ql] + list(S)
    D = prod(dropped)
    first = [(a + (PS % ql) * b) % ql for a, b in zip(acc_p[l - 1], base_p[l - 1])]
    conv = bconv_rounding([first] + [list(r) for r in acc_p[l:]], dropped, E[: l - 1])
    return [[((acc_p[i][j] - conv[i][j]) * pow(D % E[i], -1, E[i]) + base_p[i][j] * pow(ql % E[i], -1, E[i])) % E[i]
             for j in range(N)] for i in range(l - 1)]


class Toy:
    def __init__(self, L=6, seed=1):
        self.rng = random.Random(seed)
        self.L = L
        self.q = gen_primes(20, L)
        self.P = gen_primes(21, 1, self.q)[0]
        self.s = [self.rng.choice([-1, 0, 1]) for _ in range(N)]
        self.s_from = [self.rng.choice([-1, 0, 1]) for _ in range(N)]

    def small(self):
        return [self.rng.choice([-2, -1, 0, 1, 2]) for _ in range(N)]

    def keygen(self, l, alpha, dsize):
        assert l + alpha - 1 <= self.L and 1 <= dsize <= max(1, alpha)
        S = [self.q[l + i] for i in range(alpha - 1)] + [self.P]
        E = self.q[:l] + S
        PS = prod(S)
        dnum = -(-l // dsize)
        keys = []
        for d in range(dnum):
            own = range(d * dsize, min((d + 1) * dsize, l))
            e = self.small()
            kb, ka = [], []
            for idx, m in enumerate(E):
                a = [self.rng.randrange(m) for _ in range(N)]
                b = [(-x + y) % m for x, y in zip(polymul(a, [v % m for v in self.s], m), e)]
                if idx < l and idx in own:
                    b = [(x + (PS % m) * (v % m)) % m for x, v in zip(b, self.s_from)]
                kb.append(b)
                ka.append(a)
            keys.append((kb, ka))
        return keys, E, S, PS, dnum

    def keyswitch(self, c, l, alpha, dsize):
        keys, E, S, PS, dnum = self.keygen(l, alpha, dsize)
        acc = [[[0] * N for _ in E] for _ in range(2)]
        for d in range(dnum):
            own = list(range(d * dsize, min((d + 1) * dsize, l)))
            conv = bconv([c[i] for i in own], [self.q[i] for i in own], E)
            for idx, m in enumerate(E):
                ext = c[idx] if idx in own else conv[idx]
                for p in range(2):
                    prod_p = polymul(ext, keys[d][p][idx], m)
                    acc[p][idx] = [(x + y) % m for x, y in zip(acc[p][idx], prod_p)]
        outs = []
        for p in range(2):
            conv = bconv(acc[p][l:], S, self.q[:l])
            outs.append([[((acc[p][i][j] - conv[i][j]) * pow(PS % self.q[i], -1, self.q[i])) % self.q[i] for j in range(N)]
                         for i in range(l)])
        return outs

    def error(self, l, alpha, dsize):
        """max |r0 + r1 s - c s'| over the coefficients (centred, CRT-composed) for a uniform c"""
        q = self.q[:l]
        Q = prod(q)
        cpoly = [self.rng.randrange(Q) for _ in range(N)]
        c = [[x % m for x in cpoly] for m in q]
        r0, r1 = self.keyswitch(c, l, alpha, dsize)
        err = []
        for i, m in enumerate(q):
            lhs = [(x + y) % m for x, y in zip(r0[i], polymul(r1[i], [v % m for v in self.s], m))]
            rhs = polymul(c[i], [v % m for v in self.s_from], m)
            err.append([(a - b) % m for a, b in zip(lhs, rhs)])
        worst = 0
        for j in range(N):
            x = sum(err[i][j] * (Q // m) * pow(Q // m, -1, m) for i, m in enumerate(q)) % Q
            worst = max(worst, min(x, Q - x))
        return worst


if __name__ == "__main__":
    toy = Toy()
    for l, alpha, dsize in [(6, 1, 1), (5, 2, 1), (5, 2, 2), (4, 3, 2), (4, 3, 3), (3, 4, 3), (2, 5, 4), (1, 1, 1)]:
        print(l, alpha, dsize, "max |err| =", toy.error(l, alpha, dsize))
