"""Child process of tests/test_switches_gpu.py: one seeded program under the environment the parent chose
(B200CKKS_NO_PDL / B200CKKS_TERMWISE_LEAVES are read once per process).  Prints one JSON line.
  python switches_child.py limbs    - SHA-256 of the limbs after rotate / multiply / relinearize / rescale / ModRaise
  python switches_child.py leaves   - sparse-slot bootstrap and minimax ReLU at N = 2^12: values, levels, rescale counts"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, "..", "fhe-gpt-2_b200", "python"), os.path.join(HERE, "..", "oracle")]
import numpy as np

mode = sys.argv[1]
if mode == "limbs":
    import b200ckks as bk

    hybrid = len(sys.argv) > 2 and sys.argv[2] == "hybrid"
    # the hybrid chain is long enough for levels with temporary special moduli (alpha > 1 below the top, above 5 limbs)
    primes = bk.coeff_modulus_create(13, [50] + [40] * 10 + [50] if hybrid else [50, 40, 40, 40, 40, 50])
    eng = bk.Context(13, primes, device=0)
    if hybrid:
        eng.set_hybrid(True)
    sk = eng.generate_secret_key(64, seed=1)
    pk = eng.create_public_key(sk)
    rk = eng.create_relin_key(sk)
    gk = eng.create_galois_keys(sk, [1, 5])
    rng = np.random.default_rng(0)
    x, y = rng.uniform(-1, 1, eng.slots), rng.uniform(-1, 1, eng.slots)
    top = len(primes) - 1
    a = eng.encrypt(pk, eng.encode(x, top, 2.0 ** 40))
    b = eng.encrypt(pk, eng.encode(y, top, 2.0 ** 40))
    h = hashlib.sha256()
    for step in (1, 5, 1):
        eng.rotate_vector_inplace(a, step, gk)
        h.update(a.download().tobytes())
        eng.multiply_inplace(a, b)
        eng.relinearize_inplace(a, rk)
        eng.rescale_to_next_inplace(a)
        eng.mod_switch_to_inplace(b, a.limbs)
        b.scale = a.scale
        h.update(a.download().tobytes())
    # hoisted rotations (one decomposition, two automorphisms) at the current level and at the top
    for ct in (a, eng.encrypt(pk, eng.encode(x, top, 2.0 ** 40))):
        for o in eng.apply_galois_hoisted(ct, [bk.galois_elt_from_step(13, 1), bk.galois_elt_from_step(13, 5)], gk):
            h.update(o.download().tobytes())
    on, key_bytes, n_keys = eng.hybrid_info()
    print(json.dumps({"sha256": h.hexdigest(), "limbs": a.limbs, "level_key_bytes": key_bytes, "level_keys": n_keys}))
    eng.close()
else:
    from b200ckks.app import App

    app = App()
    sess = app.session(12, [51] + [46] * 16 + [51] * 14 + [51], hamming_weight=64, rotation_steps=list(range(1, 2048)))
    rng = np.random.default_rng(1)
    xb = np.tile(rng.uniform(-1, 1, 512), sess.slots // 512)
    sess.stats(reset=True)
    boot = sess.bootstrapper(9).bootstrap(sess.encrypt(xb, 2.0 ** 46, limbs=1), real_message=True)
    boot_stats = sess.stats(reset=True)
    xr = rng.uniform(-0.9, 0.9, sess.slots)
    ct = sess.encrypt(xr, 2.0 ** 46, limbs=17)
    relu = sess.relu(ct)
    relu_stats = sess.stats(reset=True)
    out = {"boot": sess.decrypt(boot).real.tolist()[:512], "boot_in": xb[:512].tolist(), "boot_limbs": boot.info()[1],
           "boot_rescales": boot_stats["rescale"], "relu": sess.decrypt(relu).real.tolist()[:1024], "relu_in": xr[:1024].tolist(),
           "relu_limbs": relu.info()[1], "relu_rescales": relu_stats["rescale"],
           "boot_relins": boot_stats["key_switch_relin"], "relu_relins": relu_stats["key_switch_relin"],
           "boot_scale": boot.info()[2], "relu_scale": relu.info()[2]}
    print(json.dumps(out))
    sess.close()
