"""CPU checks of the C-ABI shared library: it loads, exports every symbol that
include/isfm_b200.h declares, its host-side integer entry point is bit-exact against the
numpy oracle, and the compute entry points refuse to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from instantsfm_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "isfm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(isfm_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/isfm_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes table out of sync with the header"


def test_version_and_error_string():
    lib = _lib.load()
    assert b"sm_100a" in lib.isfm_version()
    assert isinstance(lib.isfm_last_error(), bytes)
    assert _lib.timer_names()[0] == "linearize"


def test_partition_points_bit_exact():
    from instantsfm_b200.engine import partition_points
    from oracle.index_prep import partition_points as ref
    rng = np.random.default_rng(0)
    for n_pt in (1, 2, 7, 1000, 50_000):
        k = 2 + rng.geometric(0.3, n_pt)
        off = np.concatenate([[0], np.cumsum(k)]).astype(np.int64)
        for world in (1, 2, 3, 4, 8):
            got = partition_points(off, world)
            assert np.array_equal(got, ref(off, world))
            assert got[0] == 0 and got[-1] == n_pt and np.all(np.diff(got) >= 0)
            if n_pt >= 1000:
                loads = off[got[1:]] - off[got[:-1]]
                assert loads.max() - loads.min() <= 2 * k.max()      # balanced to one track


def test_no_cpu_fallback():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    from instantsfm_b200.engine import BAEngine, GPEngine
    with pytest.raises(_lib.IsfmError) as e:
        BAEngine(3)
    assert e.value.code == -3
    with pytest.raises(_lib.IsfmError):
        GPEngine()


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "instantsfm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
