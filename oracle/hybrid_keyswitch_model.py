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
