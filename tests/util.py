"""Shared helpers for the parity tests: the reference (oracle/_ref/libseal_ref.so, the
reference's own SEAL compiled in place) and the engine are driven on the same raw limbs."""
import numpy as np

import refseal

SMALL_BITS = [50, 40, 40, 40, 40, 50]      # fast chain for broad op coverage (log_n = 13)
CNN_BITS = refseal.CNN_BITS                # infer_seal.cpp:288-322
GPT2_BITS = refseal.GPT2_BITS


def rand_slots(rng, n, complex_=False, mag=1.0):
    if complex_:
        return mag * (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n))
    return mag * rng.uniform(-1, 1, n)


def ref_fresh_ct(ref, values, limbs, scale):
    """encode at the top level, encrypt, mod-switch to `limbs`; returns ref ct id."""
    pt = ref.pt_new()
    ct = ref.ct_new()
    top = ref.n_primes - 1
    ref.encode(pt, values, top, scale)
    ref.encrypt(pt, ct)
    if limbs < top:
        ref.op("mod_switch_to", ct, iarg=limbs)
    ref.pt_free(pt)
    return ct


def to_engine(eng, ref, ct_id):
    size, limbs, scale, ntt = ref.ct_info(ct_id)
    return eng.ciphertext(ref.ct_get(ct_id), scale, ntt)


def assert_ct_equal(eng_ct, ref, ct_id, what=""):
    size, limbs, scale, ntt = ref.ct_info(ct_id)
    esize, elimbs, escale, entt = eng_ct.info()
    assert (esize, elimbs, entt) == (size, limbs, ntt), f"{what}: shape {esize, elimbs, entt} vs {size, limbs, ntt}"
    assert escale == scale, f"{what}: scale {escale!r} vs {scale!r}"
    got = eng_ct.download()
    want = ref.ct_get(ct_id)
    if not np.array_equal(got, want):
        bad = np.argwhere(got != want)
        raise AssertionError(f"{what}: {len(bad)} of {got.size} words differ, first at {bad[0]}: "
                             f"{got[tuple(bad[0])]} vs {want[tuple(bad[0])]}")


REFERENCE_LOG_OPS = {"multiplexed parallel convolution...": "conv", "multiplexed parallel batch normalization...": "bn",
                     "approximate ReLU...": "relu", "bootstrapping...": "bootstrap", "cipher add...": "add",
                     "multiplexed parallel downsampling...": "downsample", "average pooling...": "avgpool",
                     "fully connected layer...": "fc"}


def parse_reference_log(text):
    """A result file in the reference's format (cnn_seal.cpp:101-283 *_print wrappers, infer_seal.cpp:543-575):
    rows of (op, ms, remaining level, scale), the ten logits, total time, image and inferred label.  Same rules as
    tools/make_golden_trajectory.py, which made tests/golden/resnet20_trajectory.json from the reference's own log."""
    import re

    rows, cur, out = [], None, {"logits": None, "total_ms": None, "inferred_label": None, "image_label": None}
    for line in text.splitlines():
        line = line.strip()
        if line in REFERENCE_LOG_OPS:
            cur = {"op": REFERENCE_LOG_OPS[line], "ms": None, "level": None, "scale": None}
            rows.append(cur)
        elif line.startswith("time :") and cur is not None:
            cur["ms"] = float(line.split()[2])
        elif line.startswith("remaining level :") and cur is not None:
            cur["level"] = int(line.split()[-1])
        elif line.startswith("scale:") and cur is not None:
            cur["scale"] = float(line.split()[-1])
        elif line.startswith("total time"):
            out["total_ms"] = float(line.split()[3])
        elif line.startswith("inferred label"):
            out["inferred_label"] = int(line.split()[-1])
        elif line.startswith("image label"):
            out["image_label"] = int(line.split()[-1])
        elif line.startswith("( (") and len(re.findall(r"\(", line)) == 11:
            out["logits"] = [float(m) for m in re.findall(r"\((-?[0-9.e+-]+),", line)]
    out["rows"] = rows
    return out
