"""CPU tests (no GPU): host-side setup math of the engine (hostmath.h through the C ABI),
the exported symbol set, and the loud failure without a device."""
import ctypes
import os
import re

import numpy as np
import pytest

import b200ckks as bk
import ckks_port as port
import refseal

GOLDEN_CNN_PRIMES = [  # CoeffModulus::Create(65536, {51, 46x16, 51x14, 51}) from the reference
    2251799780524033, 70368710492161, 70368712327169, 70368714424321, 70368714817537, 70368715210753,
    70368715603969, 70368718225409, 70368720322561, 70368721764353, 70368723337217, 70368732774401,
    70368736706561, 70368737230849, 70368738410497, 70368739065857, 70368740769793, 2251799780917249,
    2251799785504769, 2251799787601921, 2251799787995137, 2251799789568001, 2251799791403009, 2251799795466241,
    2251799797432321, 2251799799267329, 2251799805165569, 2251799806345217, 2251799807131649, 2251799809884161,
    2251799810670593, 2251799813554177]


def test_every_declared_symbol_is_exported():
    text = open(bk.HEADER_PATH).read()
    names = set(re.findall(r"\b(bk_[a-z0-9_]+)\s*\(", text))
    assert len(names) > 60
    lib = ctypes.CDLL(bk.LIB_PATH)
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing


def test_every_declared_app_symbol_is_exported():
    from b200ckks import app

    text = open(app.APP_HEADER_PATH).read()
    names = set(re.findall(r"\b(bka_[a-z0-9_]+)\s*\(", text))
    assert len(names) > 40
    lib = ctypes.CDLL(app.APP_LIB_PATH)
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing


def test_coeff_modulus_create_matches_golden_chain():
    got = bk.coeff_modulus_create(16, refseal.CNN_BITS)
    assert [int(v) for v in got] == GOLDEN_CNN_PRIMES


def test_gpt2_chain_shape():
    p = bk.coeff_modulus_create(16, refseal.GPT2_BITS)
    assert len(p) == 37 and len(set(int(v) for v in p)) == 37
    assert all((int(v) - 1) % (1 << 17) == 0 for v in p)
    assert int(p[-1]).bit_length() == 60 and int(p[0]).bit_length() == 49


@pytest.mark.skipif(not refseal.available(), reason="oracle/_ref not built")
def test_coeff_modulus_create_matches_reference_small():
    for log_n, bits in ((12, [40, 36, 36, 40]), (13, [50, 40, 40, 40, 40, 50]), (15, [55, 50, 50, 55])):
        r = refseal.RefSeal(log_n, bits, hamming_weight=0, seed=1)
        assert np.array_equal(bk.coeff_modulus_create(log_n, bits), r.primes)
        for i, q in enumerate(r.primes):
            assert np.array_equal(bk.ntt_root_powers(log_n, q), r.root_powers(i))
            assert np.array_equal(bk.ntt_root_powers(log_n, q, inverse=True), r.root_powers(i, True))
        r.close()


def test_minimal_primitive_root_kat():
    """tests/seal/util/ntt.cpp:55-75: psi for q = 0xffffffffffc0001, N = 2 -> 288794978602139552"""
    assert bk.minimal_primitive_root(1, 0xFFFFFFFFFFC0001) == 288794978602139552


@pytest.mark.skipif(not port.available(), reason="oracle port not built")
def test_galois_helpers_match_oracle():
    for step in (0, 1, -1, 5, -77, 16383):
        assert bk.galois_elt_from_step(16, step) == port.galois_elt_from_step(16, step)
    for elt in (5, 25, 131071, 3125):
        assert np.array_equal(bk.galois_table_ntt(16, elt), port.galois_table_ntt(16, elt))
    with pytest.raises(bk.InvalidArgument):
        bk.galois_elt_from_step(12, 2048)


def test_no_device_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(bk.NoDevice):
        bk.Context(16, GOLDEN_CNN_PRIMES)


def test_product_path_never_touches_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "fhe-gpt-2_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                for word in ("refseal", "ckks_port", "libseal_ref", "appref", "libapp_ref", "plain_model"):
                    assert word not in src, (f, word)


def test_chacha20_block_matches_rfc8439(tmp_path):
    """csrc/rng.cuh is the generator behind key generation and encryption; its block function against the test vector of
    RFC 8439 section 2.3.2 (key 00..1f, block counter 1, nonce 00:00:00:09:00:00:00:4a:00:00:00:00)."""
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "t.cpp"
    src.write_text(r'''
#include <cstdio>
#include "rng.cuh"
int main() {
    bk::RngKey k;
    for (int i = 0; i < 8; i++) k.k[i] = (4u*i) | ((4u*i+1) << 8) | ((4u*i+2) << 16) | ((4u*i+3) << 24);
    uint32_t out[16];
    bk::chacha20_block<16>(k, 1u, 0x09000000u, 0x4a000000u, 0u, out);
    for (int i = 0; i < 16; i++) printf("%08x ", out[i]);
    printf("\n");
    bk::RngKey d1 = bk::derive_call_key(k, 1), d2 = bk::derive_call_key(k, 2);
    printf("%d\n", d1.k[0] != d2.k[0] || d1.k[1] != d2.k[1]);
}''')
    exe = str(tmp_path / "t")
    subprocess.run(["g++", "-std=c++17", "-D__host__=", "-D__device__=", "-D__forceinline__=inline",
                    "-I" + os.path.join(root, "fhe-gpt-2_b200", "csrc"), str(src), "-o", exe], check=True)
    out = subprocess.run([exe], check=True, stdout=subprocess.PIPE, text=True).stdout.split("\n")
    assert out[0].split() == ["e4e7f110", "15593bd1", "1fdd0f50", "c47120a3", "c7f4d1c7", "0368c033", "9aaa2204", "4e6cd4c3",
                              "466482d2", "09aa9f07", "05d7c214", "a2028bd9", "d19c12b5", "b94e16de", "e883d0cb", "4e3c50a2"]
    assert out[1] == "1"


def test_every_kernel_waits_for_its_predecessors_before_touching_memory():
    """Programmatic dependent launch (engine.h launch_pdl) lets a kernel be scheduled while its predecessor still runs;
    that is only safe if EVERY kernel starts with pdl_prologue() (griddepcontrol.wait) and no launch bypasses
    launch_pdl.  Static check over fhe-gpt-2_b200/csrc."""
    import glob
    import re

    src_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "fhe-gpt-2_b200", "csrc")
    kernels = 0
    for path in glob.glob(os.path.join(src_dir, "*.cu")) + glob.glob(os.path.join(src_dir, "*.cuh")):
        text = open(path).read()
        assert "<<<" not in text, f"{path}: a triple-chevron launch bypasses launch_pdl"
        for m in re.finditer(r"__global__", text):
            pos = m.end()
            while True:  # skip __launch_bounds__(...) / __cluster_dims__(...) up to the parameter list
                par = text.index("(", pos)
                name = re.search(r"(\w+)\s*$", text[:par]).group(1)
                depth, q = 0, par
                while True:
                    depth += text[q] == "("
                    depth -= text[q] == ")"
                    if depth == 0:
                        break
                    q += 1
                if name in ("__launch_bounds__", "__cluster_dims__"):
                    pos = q + 1
                    continue
                break
            body = text[text.index("{", q):][:120]
            assert "pdl_prologue();" in body, f"{path}: kernel {name} does not begin with pdl_prologue()"
            kernels += 1
    assert kernels >= 30
