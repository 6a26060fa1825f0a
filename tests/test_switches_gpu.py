"""Round-2 engine switches against their plain counterparts, each in its own process (the switches are read once):

* programmatic dependent launch ($B200CKKS_NO_PDL): identical limbs, bit for bit, on both key-switching paths - PDL
  only moves WHEN a kernel is set up, every kernel still waits for its predecessors before touching memory;
* seed-compressed level keys ($B200CKKS_COMPRESS_KEYS): identical limbs at half the resident key bytes;
* one-pass leaves of the polynomial evaluation trees ($B200CKKS_TERMWISE_LEAVES restores the reference's
  multiply_const + rescale + reduced-error add per term): same output level, same values within the bootstrapping /
  ReLU tolerance, and the rescale count the fusion exists to cut."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def child(args, **env):
    e = dict(os.environ, B200CKKS_SEED="11")
    for k in ("B200CKKS_NO_PDL", "B200CKKS_TERMWISE_LEAVES", "B200CKKS_HYBRID_KS", "B200CKKS_ENCRYPT_CONSTANTS",
              "B200CKKS_COMPRESS_KEYS"):
        e.pop(k, None)
    e.update(env)
    r = subprocess.run([sys.executable, os.path.join(HERE, "switches_child.py"), *args], env=e, capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("path", ["classic", "hybrid"])
def test_programmatic_dependent_launch_is_bit_identical(path):
    args = ["limbs"] + (["hybrid"] if path == "hybrid" else [])
    with_pdl = child(args)
    without = child(args, B200CKKS_NO_PDL="1")
    assert with_pdl == without


def test_seed_compressed_keys_are_bit_identical_at_half_the_bytes():
    """level keys keep only their non-uniform halves resident; the uniform halves are regenerated
    from their public ChaCha8 key each time a key is used.  Same seeds -> the regenerated words are the words keygen
    used, so every rotation / relinearization (hybrid levels, SEAL-shaped top level, hoisted) is identical bit for bit."""
    plain = child(["limbs", "hybrid"])
    packed = child(["limbs", "hybrid"], B200CKKS_COMPRESS_KEYS="1")
    assert packed["sha256"] == plain["sha256"] and packed["limbs"] == plain["limbs"]
    assert packed["level_keys"] == plain["level_keys"] > 0
    assert packed["level_key_bytes"] * 2 == plain["level_key_bytes"]


def test_one_pass_tree_leaves_match_the_termwise_sequence():
    fused = child(["leaves"], B200CKKS_HYBRID_KS="1")
    term = child(["leaves"], B200CKKS_HYBRID_KS="1", B200CKKS_TERMWISE_LEAVES="1")
    # same levels out
    assert fused["boot_limbs"] == term["boot_limbs"] and fused["relu_limbs"] == term["relu_limbs"]
    # the point of the fusion: one rescale per leaf and none for walking operands down
    assert fused["boot_rescales"] < 0.5 * term["boot_rescales"], (fused["boot_rescales"], term["boot_rescales"])
    assert fused["relu_rescales"] < term["relu_rescales"], (fused["relu_rescales"], term["relu_rescales"])
    x = np.array(fused["boot_in"])
    eb_f, eb_t = np.abs(np.array(fused["boot"]) - x).max(), np.abs(np.array(term["boot"]) - x).max()
    assert eb_f < 1e-4 and eb_t < 1e-4, (eb_f, eb_t)            # the bootstrapping tolerance of tests/app_cases.py
    assert np.abs(np.array(fused["boot"]) - np.array(term["boot"])).max() < 1e-4
    xr = np.array(fused["relu_in"])
    want = np.maximum(xr, 0)
    er_f, er_t = np.abs(np.array(fused["relu"]) - want).max(), np.abs(np.array(term["relu"]) - want).max()
    assert er_f < 2.0 ** -13 and er_t < 2.0 ** -13, (er_f, er_t)  # the minimax ReLU's bound (alpha = 13)
