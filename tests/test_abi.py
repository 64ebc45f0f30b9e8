"""The C-ABI library loads and exports every symbol include/kmx.h declares (no compute calls:
this runs without a GPU)."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    h = open(os.path.join(ROOT, "include", "kmx.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(kmx_[a-z0-9_]+)\s*\(", h)))


def test_library_exports_every_declared_symbol():
    from kmtricks_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "run ./build.sh (or __graft_entry__.build())"
    L = C.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/kmx.h but not exported"


def test_python_binding_covers_the_header():
    from kmtricks_b200 import _lib
    assert sorted(_lib.SYMBOLS) == declared_symbols()


def test_no_device_is_a_loud_error_not_a_fallback():
    """Without a usable sm_100 device kmx_create must fail (there is no CPU path)."""
    import numpy as np
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("GPU present")
    from kmtricks_b200 import engine
    import pytest
    with pytest.raises(engine.KmxError):
        engine.Engine(engine.Config(kmer_size=31, nb_partitions=4), 1)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "kmtricks_b200")
    for r, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".inl", ".cpp", ".hpp")):
                src = open(os.path.join(r, f), errors="ignore").read()
                assert "oracle" not in src.replace("oracle/_ref", "").lower() or f == "synth.py" and False, f"{f} mentions the oracle"
