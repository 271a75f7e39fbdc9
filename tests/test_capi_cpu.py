"""CPU-side checks of the C-ABI library: it loads, exports every symbol the header declares, and the
Python prototypes cover the header 1:1. No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

from ingvio_b200 import capi


def _built():
    return os.path.exists(capi.LIB_PATH)


def test_header_and_prototypes_agree():
    assert set(capi.header_symbols()) == set(capi.SIGNATURES.keys())


@pytest.mark.skipif(not _built(), reason="libingvio_b200.so not built (run __graft_entry__.build())")
def test_library_exports_every_header_symbol():
    lib = capi.load()
    for name in capi.header_symbols():
        assert hasattr(lib, name), name


@pytest.mark.skipif(not _built(), reason="libingvio_b200.so not built")
def test_no_torch_types_in_abi():
    hdr = open(os.path.join(os.path.dirname(capi._HERE), "include", "ingvio_b200.h")).read()
    assert "torch" not in hdr and "at::" not in hdr and "#include <cuda" not in hdr


def test_product_never_imports_oracle():
    root = os.path.dirname(capi._HERE)
    bad = []
    for dp, _, files in os.walk(os.path.join(root, "ingvio_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+(ingvio_oracle|oracle)\b", txt, re.M) or \
                        re.search(r"(dlopen|CDLL|#include)[^\n]*oracle", txt):
                    bad.append(f)
    assert not bad, bad


@pytest.mark.skipif(not _built(), reason="libingvio_b200.so not built")
def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ingvio_b200.filter import BatchFilter
    with pytest.raises(capi.IgvError):
        BatchFilter(1, 4, 8, 4)
