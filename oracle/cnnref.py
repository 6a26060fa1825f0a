"""TEST INFRASTRUCTURE ONLY.  The REFERENCE's OWN application code (cnn_ckks/common, ckks_bootstrapping, comp, cnn,
run/run_cnn.cpp), compiled unmodified from /root/reference by oracle/Makefile with oracle/ntl_shim standing in for NTL:

  oracle/_ref/libcnn_ref.so      on the reference's SEAL - the L2-L4 oracle (CPU)
  oracle/_ref/libcnn_dropin.so   on this repo's seal:: facade + libb200ckks.so - the reference's object code on the engine
  oracle/_ref/cnn_dropin         the reference's `cnn` program (run/run_cnn.cpp main) on the engine

Both libraries export the subset of include/b200ckks_app.h that oracle/cnn_ref_capi.cpp serves, so tests drive them
through b200ckks.app.App(path).  Never imported by the product path."""
import contextlib
import os
import shutil
import sys
import tempfile

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.normpath(os.path.join(_HERE, ".."))
sys.path.insert(0, os.path.join(ROOT, "fhe-gpt-2_b200", "python"))
LIB_REF = os.path.join(_HERE, "_ref", "libcnn_ref.so")
LIB_DROPIN = os.path.join(_HERE, "_ref", "libcnn_dropin.so")
CNN_DROPIN = os.path.join(_HERE, "_ref", "cnn_dropin")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def available():
    return os.path.exists(LIB_REF)


def dropin_available():
    return os.path.exists(LIB_DROPIN) and os.path.exists(CNN_DROPIN)


def app():
    from b200ckks.app import App

    a = App(LIB_REF)
    assert a.backend == "reference-app/reference-seal"
    return a


def dropin_app():
    from b200ckks.app import App

    a = App(LIB_DROPIN)
    assert a.backend == "reference-app/engine"
    return a


def make_tree(base, layer_num=20, images=None, labels=None):
    """The directory layout the reference's code expects around its working directory cnn_ckks/build
    (SEALcomp.cpp:12-15 ../result/d13.txt; infer_seal.cpp:3-107 ../../pretrained_parameters/resnet<L>_new/*.txt;
    :441,449 ../../../testFile/test_{values,label}.txt; :409 ../../result/): returns the working directory."""
    repo = os.path.join(base, "FHE-GPT-2")
    build = os.path.join(repo, "cnn_ckks", "build")
    os.makedirs(build, exist_ok=True)
    os.makedirs(os.path.join(repo, "cnn_ckks", "result"), exist_ok=True)
    os.makedirs(os.path.join(repo, "result"), exist_ok=True)
    shutil.copy(os.path.join(GOLDEN, "cnn_ckks_result", "d13.txt"), os.path.join(repo, "cnn_ckks", "result", "d13.txt"))
    src = os.path.join(GOLDEN, "pretrained_parameters", f"resnet{layer_num}_new")
    if os.path.isdir(src):
        dst = os.path.join(repo, "pretrained_parameters", f"resnet{layer_num}_new")
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.exists(dst):
            os.symlink(src, dst)
    if images is not None:
        tf = os.path.join(base, "testFile")
        os.makedirs(tf, exist_ok=True)
        with open(os.path.join(tf, "test_values.txt"), "w") as f:
            for img in images:
                f.write(" ".join(repr(float(v)) for v in img.reshape(-1)) + "\n")
        with open(os.path.join(tf, "test_label.txt"), "w") as f:
            f.write("\n".join(str(int(v)) for v in labels) + "\n")
    return build


@contextlib.contextmanager
def workdir(layer_num=20, images=None, labels=None):
    """chdir into a fresh copy of that layout for the duration of the block"""
    old = os.getcwd()
    base = tempfile.mkdtemp(prefix="cnnref_")
    try:
        os.chdir(make_tree(base, layer_num, images, labels))
        yield base
    finally:
        os.chdir(old)
        shutil.rmtree(base, ignore_errors=True)
