"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/nrrt.h declares."""
import os
import re

from tests.conftest import ROOT


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from mgr_sampling_cnn_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "nrrt.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(nrrt_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for sym in declared:
        assert hasattr(lib, sym), f"libnrrt.so does not export {sym}"
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    assert lib.nrrt_version() == 100
    assert lib.nrrt_words_per_map(512 * 512) == 8192 and lib.nrrt_words_per_map(80 ** 3) == 16000


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "mgr_sampling_cnn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"
                assert "planner_oracle" not in src or f.endswith(".cuh") or "oracle/" in src, f


def test_missing_cuda_device_is_a_hard_error():
    import torch
    import pytest
    from mgr_sampling_cnn_b200 import _lib, engine
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(_lib.NrrtError):
        engine.pack_occupancy(torch.zeros(1, 8, 8, dtype=torch.uint8), 2)
