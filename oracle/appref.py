"""TEST INFRASTRUCTURE ONLY.  The application layers of fhe-gpt-2_b200/host (Bootstrapper, comp, cnn and their C ABI)
compiled against the REFERENCE's own SEAL (oracle/Makefile target `app_ref` -> oracle/_ref/libapp_ref.so): the
identical host code running on the reference library on the CPU.  Used by tests/ as the oracle for the application
layers and by bench.py's CPU-baseline legs; never by the product path."""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.normpath(os.path.join(_HERE, "..", "fhe-gpt-2_b200", "python")))
LIB = os.path.join(_HERE, "_ref", "libapp_ref.so")


def available():
    return os.path.exists(LIB)


def app():
    from b200ckks.app import App

    a = App(LIB)
    assert a.backend == "reference-seal"
    return a
