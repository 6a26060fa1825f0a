"""Round-2 engine switches against their plain counterparts, each in its own process (the switches are read once):

* programmatic dependent launch ($B200CKKS_NO_PDL): identical limbs, bit for bit, on both key-switching paths - PDL
  only moves WHEN a kernel is set up, every kernel still waits for its predecessors before touching memory;
* seed-compressed level keys ($B200CKKS_COMPRESS_KEYS): identical limbs at half the resident key bytes;
* one-pass leaves of the polynomial evaluation trees ($B200CKKS_TERMWISE_LEAVES restores the reference's
  multiply_const + rescale + reduced-error add per term): same output level, same values within the bootstrapping /
  ReLU tolerance, and the rescale count the fusion exists to cut;
* relinearization and rescale as one division ($B200CKKS_MERGED_RESCALE=0 keeps the two calls and the reference's
  order of additions): same levels and scales out, same values, three relinearizations fewer per ReLU."""
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
              "B200CKKS_COMPRESS_KEYS", "B200CKKS_MERGED_RESCALE", "B200CKKS_EARLY_RESCALE"):
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


@pytest.mark.parametrize("hybrid", ["1", "0"])
def test_merged_relinearize_and_rescale_match_the_two_calls(hybrid):
    """Evaluator::relinearize_rescale_inplace (bk_relinearize_rescale_inplace: ModDown by P_S and division by q_last as
    one rounding division in hybrid mode, the two calls otherwise) with the additions of a product moved in front of
    its relinearization, against relinearize_inplace + rescale_to_next_inplace in the reference's order."""
    merged = child(["leaves"], B200CKKS_HYBRID_KS=hybrid)
    two = child(["leaves"], B200CKKS_HYBRID_KS=hybrid, B200CKKS_MERGED_RESCALE="0")
    for k in ("boot_limbs", "relu_limbs", "boot_rescales", "relu_rescales"):
        assert merged[k] == two[k], k
    assert abs(merged["boot_scale"] / two["boot_scale"] - 1) < 1e-9 and abs(merged["relu_scale"] / two["relu_scale"] - 1) < 1e-9
    # a chain of products T_g1 q_1 + T_g2 q_2 + r is relinearized once: 27 -> 24 per ReLU (one chain in each of the three
    # polynomials); EvalMod has no such chain
    assert (two["relu_relins"], merged["relu_relins"]) == (27, 24)
    assert two["boot_relins"] == merged["boot_relins"] == 18
    x = np.array(merged["boot_in"])
    eb_m, eb_t = np.abs(np.array(merged["boot"]) - x).max(), np.abs(np.array(two["boot"]) - x).max()
    assert eb_m < 1e-4 and eb_t < 1e-4, (eb_m, eb_t)
    assert eb_m < 2 * eb_t + 1e-6, (eb_m, eb_t)                # no precision given up (same keys, same input)
    xr = np.array(merged["relu_in"])
    want = np.maximum(xr, 0)
    er_m, er_t = np.abs(np.array(merged["relu"]) - want).max(), np.abs(np.array(two["relu"]) - want).max()
    assert er_m < 2.0 ** -13 and er_t < 2.0 ** -13, (er_m, er_t)
    assert np.abs(np.array(merged["relu"]) - np.array(two["relu"])).max() < 2.0 ** -13


def test_early_rescale_of_double_hoisted_transforms():
    """$B200CKKS_EARLY_RESCALE=1 (opt-in, common/func.h): the rescale after a linear transform is taken per inner sum,
    merged with its division by the special modulus (bk_bsgs_inner_sums, rescale = 1), and the giant-step rotations run
    one level lower.  Same level, scale and rescale count out; values within the bootstrapping tolerance."""
    early = child(["leaves"], B200CKKS_HYBRID_KS="1", B200CKKS_EARLY_RESCALE="1")
    late = child(["leaves"], B200CKKS_HYBRID_KS="1")
    assert early["boot_limbs"] == late["boot_limbs"] and early["boot_rescales"] == late["boot_rescales"]
    assert abs(early["boot_scale"] / late["boot_scale"] - 1) < 1e-9
    x = np.array(early["boot_in"])
    e_early, e_late = np.abs(np.array(early["boot"]) - x).max(), np.abs(np.array(late["boot"]) - x).max()
    assert e_early < 1e-4 and e_late < 1e-4, (e_early, e_late)
