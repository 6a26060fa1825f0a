import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fhe-gpt-2_b200", "python"))
import numpy as np
import b200ckks as bk
bits = [51] + [46] * 16 + [51] * 14 + [51]
for log_n in (14, 15):
    primes = bk.coeff_modulus_create(log_n, bits)
    eng = bk.Context(log_n, primes)
    sk = eng.generate_secret_key(64, 1); pk = eng.create_public_key(sk); rk = eng.create_relin_key(sk)
    conj_elt = 2 * eng.n - 1
    gk = eng.create_galois_keys(sk, [0, 1, 5])
    rng = np.random.default_rng(0)
    S = 2.0 ** 46
    x = rng.uniform(-1, 1, eng.slots) + 1j * rng.uniform(-1, 1, eng.slots)
    y = rng.uniform(-1, 1, eng.slots) + 1j * rng.uniform(-1, 1, eng.slots)
    dec = lambda c: eng.decode(eng.decrypt(sk, c))
    cx = eng.encrypt(pk, eng.encode(x, 31, S)); cy = eng.encrypt(pk, eng.encode(y, 31, S))
    c = cx.copy(); eng.complex_conjugate_inplace(c, gk); print(log_n, "conj31", np.abs(dec(c) - np.conj(x)).max())
    for L in (31, 28, 24, 20, 17):
        a = cx.copy(); b = cy.copy(); eng.mod_switch_to_inplace(a, L); eng.mod_switch_to_inplace(b, L)
        eng.multiply_inplace(a, b); eng.relinearize_inplace(a, rk); eng.rescale_to_next_inplace(a)
        print(log_n, "mul-relin-rescale", L, np.abs(dec(a) - x * y).max(), a.info())
        a = cx.copy(); eng.mod_switch_to_inplace(a, L); eng.square_inplace(a); eng.relinearize_inplace(a, rk); eng.rescale_to_next_inplace(a)
        print(log_n, "square", L, np.abs(dec(a) - x * x).max())
        a = cx.copy(); eng.mod_switch_to_inplace(a, L); eng.rotate_vector_inplace(a, 5, gk); print(log_n, "rot5", L, np.abs(dec(a) - np.roll(x, -5)).max())
        a = cx.copy(); eng.mod_switch_to_inplace(a, L); eng.complex_conjugate_inplace(a, gk); print(log_n, "conj", L, np.abs(dec(a) - np.conj(x)).max())
        a = cx.copy(); eng.mod_switch_to_inplace(a, L); eng.multiply_const_inplace(a, 0.37); eng.rescale_to_next_inplace(a); print(log_n, "mulconst", L, np.abs(dec(a) - 0.37 * x).max())
        a = cx.copy(); eng.mod_switch_to_inplace(a, L); eng.add_const_inplace(a, -1.0); print(log_n, "addconst", L, np.abs(dec(a) - (x - 1)).max())
        a = cx.copy(); eng.mod_switch_to_inplace(a, L); a.scale = float(primes[0]); p = eng.encode(y, L, float(primes[0]), top_dropped=True)
        eng.multiply_plain_inplace(a, p); eng.rescale_to_next_inplace(a)
        print(log_n, "mulplain@q0scale", L, np.abs(dec(a) * a.scale / (S * float(primes[0]) / float(primes[L-1])) - x * y).max())
    # modraise
    a = cx.copy(); eng.mod_switch_to_inplace(a, 1)
    b = a.copy(); eng.transform_from_ntt_inplace(b); coef = b.download()  # [2][1][N]
    eng.modraise_inplace(a); r = a.copy(); eng.transform_from_ntt_inplace(r); lifted = r.download()
    q0 = int(primes[0]); ok = True
    cc = coef[:, 0, :].astype(object)
    for j in range(31):
        q = int(primes[j])
        want = np.array([[(int(v) - q0 if int(v) > q0 // 2 else int(v)) % q for v in cc[p]] for p in range(2)], dtype=object)
        ok &= bool((lifted[:, j, :].astype(object) == want).all())
    print(log_n, "modraise exact", ok)
    eng.close()
