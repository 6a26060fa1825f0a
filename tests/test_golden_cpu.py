"""Golden fixtures: the reference's committed run log (levels and scales per stage) and the per-level operation
histogram of one ResNet-20 inference, checked against what the reference's call graph implies."""
import json
import os

import numpy as np

import b200ckks as bk
import refseal

HERE = os.path.dirname(os.path.abspath(__file__))


def test_resnet20_operation_histogram_matches_the_reference_call_graph():
    h = json.load(open(os.path.join(HERE, "golden", "resnet20_op_histogram.json")))
    relins = 18 * 18 + 19 * 27          # 18 bootstraps x (16 + 2) + 19 ReLUs x (8 + 8 + 10 + 1)
    assert sum(h["multiply"]) == relins
    boot_rot = 6 * (76 + 69 + 62)       # logn 14 / 13 / 12, six bootstraps each
    conv_rot = 28 + 6 * 56 + 106 + 5 * 82 + 158 + 5 * 126
    assert sum(h["key_switch"]) - relins >= boot_rot + conv_rot
    assert sum(h["key_switch"]) - relins <= boot_rot + conv_rot + 200      # down-sampling, pooling, FC, replication
    boot_enc = 6 * (299 + 235 + 171)
    conv_enc = 34 + 6 * 88 + 176 + 5 * 104 + 208 + 5 * 136
    assert boot_enc + conv_enc <= sum(h["multiply_vector"]) <= boot_enc + conv_enc + 200
    # bootstrapping works at 18..31 limbs, the convolutions at <= 3 limbs (layer 0 at 20 -> 18)
    assert sum(h["multiply_vector"][4:18]) == 0


def test_trajectory_fixture_scales_follow_from_the_prime_chain():
    """result/resnet20_cifar10_image0.txt: scale after the first ReLU 7.0449e+13, after a convolution 7.06905e+13,
    after a bootstrap 2^46.  The convolution's scale is pure double arithmetic on the chain: two rescales from 3 limbs
    starting at the ReLU's output scale."""
    t = json.load(open(os.path.join(HERE, "golden", "resnet20_trajectory.json")))
    p = [float(int(v)) for v in bk.coeff_modulus_create(16, refseal.CNN_BITS)]
    relu = next(r for r in t["rows"] if r["op"] == "relu")
    conv = [r for r in t["rows"] if r["op"] == "conv"][1]
    assert relu["level"] == 2 and conv["level"] == 0
    s = relu["scale"]
    got = (s * s / p[2]) ** 2 / p[1]
    assert abs(got / conv["scale"] - 1) < 5e-5          # the log prints 6 significant digits
    assert all(r["level"] == 16 and abs(r["scale"] / 2.0 ** 46 - 1) < 1e-5 for r in t["rows"] if r["op"] == "bootstrap")
    assert [r["op"] for r in t["rows"]].count("bootstrap") == 18 and len(t["rows"]) == 88
    assert np.isclose(t["total_ms"], 2188790)
