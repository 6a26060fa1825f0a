"""Synthetic inputs of the reference's ResNet shapes (there is no network for datasets or checkpoints, and the
reference's CIFAR image file is missing): random-init weights in the parameter layout of
cnn_ckks/pretrained_parameters/resnet<L>_new (infer_seal.cpp:3-107) and SURVEY.md 8(d) images."""
import math

import numpy as np


def resnet_shapes(layer_num):
    """(end_num, [(ci, co) per convolution in the reference's file order])"""
    end_num = {20: 2, 32: 4, 44: 6, 56: 8, 110: 17}[layer_num]
    shapes = [(3, 16)]
    for j in range(3):
        co = 16 << j
        for k in range(end_num + 1):
            shapes.append((co // 2 if (j > 0 and k == 0) else co, co))
            shapes.append((co, co))
    return end_num, shapes


def random_weights(layer_num, seed=0, classes=10):
    """He-style convolutions and near-identity batch-norm statistics: activations stay well inside the [-B, B] = [-40, 40]
    range the network's approximate ReLU covers.  classes = 100 adds the 1x1 stride-2 shortcut convolutions of the
    CIFAR-100 network (infer_seal.cpp:108-250, :826-831)."""
    rng = np.random.default_rng(seed)
    _, shapes = resnet_shapes(layer_num)
    w = dict(conv_weight=[], bn_bias=[], bn_mean=[], bn_var=[], bn_weight=[])
    for ci, co in shapes:
        w["conv_weight"].append(rng.normal(0, math.sqrt(2.0 / (9 * ci)), 9 * ci * co) * 0.5)
        w["bn_bias"].append(rng.normal(0, 0.1, co))
        w["bn_mean"].append(rng.normal(0, 0.1, co))
        w["bn_var"].append(rng.uniform(0.5, 1.5, co))
        w["bn_weight"].append(rng.uniform(0.5, 1.0, co))
    w["linear_weight"] = rng.normal(0, 0.3, classes * 64)
    w["linear_bias"] = rng.normal(0, 0.1, classes)
    if classes == 100:
        for k in ("shortcut_weight", "shortcut_bn_bias", "shortcut_bn_mean", "shortcut_bn_var", "shortcut_bn_weight"):
            w[k] = []
        for ci, co in ((16, 32), (32, 64)):
            w["shortcut_weight"].append(rng.normal(0, math.sqrt(2.0 / ci), ci * co) * 0.5)
            w["shortcut_bn_bias"].append(rng.normal(0, 0.1, co))
            w["shortcut_bn_mean"].append(rng.normal(0, 0.1, co))
            w["shortcut_bn_var"].append(rng.uniform(0.5, 1.5, co))
            w["shortcut_bn_weight"].append(rng.uniform(0.5, 1.0, co))
    return w


def load_pretrained(directory, layer_num):
    """The reference's parameter files (pretrained_parameters/resnet<L>_new/*.txt) in the order
    import_parameters_cifar10 reads them (infer_seal.cpp:3-107): whitespace-separated values, one tensor per file."""
    import os

    end_num, shapes = resnet_shapes(layer_num)

    def rd(name, count):
        with open(os.path.join(directory, name + ".txt")) as f:
            v = np.array(f.read().split(), dtype=np.float64)
        if v.size < count:
            raise ValueError(f"{name}.txt holds {v.size} values, {count} expected")
        return v[:count]

    conv_names, bn_names = ["conv1"], ["bn1"]
    for j in range(1, 4):
        for k in range(end_num + 1):
            conv_names += [f"layer{j}_{k}_conv1", f"layer{j}_{k}_conv2"]
            bn_names += [f"layer{j}_{k}_bn1", f"layer{j}_{k}_bn2"]
    w = dict(conv_weight=[], bn_bias=[], bn_mean=[], bn_var=[], bn_weight=[])
    for (ci, co), cn, bn in zip(shapes, conv_names, bn_names):
        w["conv_weight"].append(rd(cn + "_weight", 9 * ci * co))
        w["bn_bias"].append(rd(bn + "_bias", co))
        w["bn_mean"].append(rd(bn + "_running_mean", co))
        w["bn_var"].append(rd(bn + "_running_var", co))
        w["bn_weight"].append(rd(bn + "_weight", co))
    w["linear_weight"] = rd("linear_weight", 10 * 64)
    w["linear_bias"] = rd("linear_bias", 10)
    return w


def pretrained_dir(layer_num=20):
    """tests/golden/pretrained_parameters/resnet<L>_new (a copy of the reference's trained parameters kept as a test
    fixture), or None when it is not there"""
    import os

    d = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "..", "tests", "golden",
                                      "pretrained_parameters", f"resnet{layer_num}_new"))
    return d if os.path.isdir(d) else None


def synthetic_image(image_id):
    """3072 i.i.d. N(0,1) values clipped to [-2.5, 2.5] (normalised-CIFAR-like), seed = image id, CHW order."""
    return np.clip(np.random.default_rng(image_id).normal(0, 1, 3072), -2.5, 2.5)


def shard(n_items, rank, world):
    """items of rank `rank` when n_items independent units are dealt round-robin to `world` ranks
    (image i -> GPU i mod G, the engine's counterpart of the reference's OpenMP loop over images, infer_seal.cpp:404)"""
    return list(range(rank, n_items, world))
