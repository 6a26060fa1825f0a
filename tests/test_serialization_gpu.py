"""SEAL 3.6's binary wire format: Ciphertext / RelinKeys / GaloisKeys / SecretKey / PublicKey
::save and ::load of the facade (fhe-gpt-2_b200/host/seal/serialization.h) against the reference's own SEAL
(oracle/_ref/libseal_ref.so: serialization.cpp, ciphertext.cpp:183-360, kswitchkeys.cpp:42-145).

* keys and ciphertexts the REFERENCE saved load into the engine, and rotations / multiplications with them are
  limb-identical to the reference's;
* what the ENGINE saved loads into the reference, byte-identical after a round trip;
* malformed input is refused with the reference's exception types."""
import os

import numpy as np
import pytest

import refseal
from util import SMALL_BITS, rand_slots, ref_fresh_ct

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pair():
    from b200ckks.app import App

    ref = refseal.RefSeal(13, SMALL_BITS, hamming_weight=64, seed=41)
    ref.make_galois_keys([1, -3, 0])
    sess = App().session(13, SMALL_BITS, hamming_weight=64, rotation_steps=[1, -3, 0])
    yield ref, sess
    sess.close()
    ref.close()


def test_reference_keys_and_ciphertexts_load_into_the_engine(pair, tmp_path):
    ref, sess = pair
    rng = np.random.default_rng(3)
    p = lambda name: str(tmp_path / name)
    for what in ("relin_keys", "galois_keys", "secret_key", "public_key"):
        ref.save(what, p(what))
        sess.load(what, p(what))
    x, y = rand_slots(rng, ref.n // 2), rand_slots(rng, ref.n // 2)
    a, b = ref_fresh_ct(ref, x, 5, 2.0 ** 40), ref_fresh_ct(ref, y, 5, 2.0 ** 40)
    ref.save("ciphertext", p("a"), a)
    ref.save("ciphertext", p("b"), b)
    ea, eb = sess.load("ciphertext", p("a")), sess.load("ciphertext", p("b"))
    assert ea.info() == ref.ct_info(a)[:3] and np.array_equal(ea.download(), ref.ct_get(a))
    # rotation with the loaded Galois key, multiplication + relinearization with the loaded relinearization key
    ref.op("rotate", a, iarg=1)
    sess.rotate(ea, 1)
    assert np.array_equal(ea.download(), ref.ct_get(a)), "rotation with the loaded Galois key differs"
    ref.op("multiply_reduced_error", a, b)
    sess.reduced_error_op("multiply", ea, eb)
    assert np.array_equal(ea.download(), ref.ct_get(a)), "relinearization with the loaded key differs"
    # the loaded secret key decrypts; the loaded public key encrypts for the reference's secret key
    assert np.abs(sess.decrypt(ea).real - np.roll(x, -1) * y).max() < 1e-4
    fresh = sess.encrypt(x, 2.0 ** 40, limbs=4)
    sess.save("ciphertext", p("fresh"), fresh)
    c = ref.ct_new()
    ref.load("ciphertext", p("fresh"), c)
    pt = ref.pt_new()
    ref.decrypt(c, pt)
    assert np.abs(ref.decode(pt).real - x).max() < 1e-5


def test_round_trips_are_byte_identical(pair, tmp_path):
    ref, sess = pair
    p = lambda name: str(tmp_path / name)
    for what in ("relin_keys", "galois_keys", "secret_key", "public_key"):
        ref.save(what, p(what))
        sess.load(what, p(what))
        sess.save(what, p(what + ".engine"))
        assert open(p(what), "rb").read() == open(p(what + ".engine"), "rb").read(), what
    a = ref_fresh_ct(ref, rand_slots(np.random.default_rng(5), ref.n // 2), 3, 2.0 ** 40)
    ref.save("ciphertext", p("ct"), a)
    sess.save("ciphertext", p("ct.engine"), sess.load("ciphertext", p("ct")))
    assert open(p("ct"), "rb").read() == open(p("ct.engine"), "rb").read()


def test_engine_generated_keys_load_into_the_reference(tmp_path):
    """keys generated on the device, saved in SEAL's format, used by the reference's Evaluator"""
    from b200ckks.app import App

    sess = App().session(13, SMALL_BITS, hamming_weight=64, rotation_steps=[1])
    ref = refseal.RefSeal(13, SMALL_BITS, hamming_weight=64, seed=43)
    p = lambda name: str(tmp_path / name)
    sess.save("galois_keys", p("gk"))
    sess.save("relin_keys", p("rk"))
    ref.load("galois_keys", p("gk"))
    ref.load("relin_keys", p("rk"))
    x = rand_slots(np.random.default_rng(7), ref.n // 2)
    ct = sess.encrypt(x, 2.0 ** 40, limbs=5)
    sess.save("ciphertext", p("ct"), ct)
    c = ref.ct_new()
    ref.load("ciphertext", p("ct"), c)
    ref.op("rotate", c, iarg=1)          # the reference's Evaluator on the engine's Galois key
    ref.op("square", c)
    ref.op("relinearize", c)             # ... and on the engine's relinearization key
    ref.save("ciphertext", p("out"), c)
    out = sess.load("ciphertext", p("out"))
    assert np.abs(sess.decrypt(out).real - np.roll(x, -1) ** 2).max() < 1e-4
    sess.close()
    ref.close()


def test_malformed_input_is_refused(pair, tmp_path):
    ref, sess = pair
    p = lambda name: str(tmp_path / name)
    a = ref_fresh_ct(ref, rand_slots(np.random.default_rng(9), ref.n // 2), 3, 2.0 ** 40)
    ref.save("ciphertext", p("ct"), a)
    raw = bytearray(open(p("ct"), "rb").read())
    bad_magic = bytes([0, 0]) + bytes(raw[2:])
    open(p("bad1"), "wb").write(bad_magic)
    with pytest.raises(RuntimeError, match="SEALHeader"):
        sess.load("ciphertext", p("bad1"))
    foreign = bytearray(raw)
    foreign[16] ^= 0xFF                     # parms_id of other parameters
    open(p("bad2"), "wb").write(foreign)
    with pytest.raises(RuntimeError, match="invalid"):
        sess.load("ciphertext", p("bad2"))
    open(p("bad3"), "wb").write(bytes(raw[: len(raw) // 2]))
    with pytest.raises(RuntimeError):
        sess.load("ciphertext", p("bad3"))
