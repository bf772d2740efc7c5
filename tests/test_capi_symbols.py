"""CPU test: the C-ABI library loads and exports every entry point include/*.h declares."""
import ctypes
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        text = open(h).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names.update(re.findall(r"\b(shb_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    from shasta_b200 import capi
    assert os.path.exists(capi.LIB_PATH), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(capi.LIB_PATH)
    syms = declared_symbols()
    assert "shb_lowhash0" in syms and "shb_set_markers" in syms
    for s in syms:
        assert hasattr(lib, s), f"{s} is declared in include/ but not exported"


def test_no_cpu_fallback_without_device():
    # Without a GPU the context creation must fail loudly (SHB_ERR_CUDA), never fall back to a CPU path.
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from shasta_b200 import capi
    with pytest.raises(capi.ShastaB200Error):
        capi.Context(0)


def test_product_does_not_import_oracle():
    # The oracle is test infrastructure; nothing under shasta_b200/ may import, include, link or call it.
    forbidden = re.compile(r"\bfrom\s+oracle\b|\bimport\s+oracle\b|oracle/|liboracle|libshasta_ref|\borc_[a-z]|\bref_[a-z]+\(")
    for path in glob.glob(os.path.join(ROOT, "shasta_b200", "**", "*"), recursive=True):
        if os.path.isfile(path) and path.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
            for line_no, line in enumerate(open(path), 1):
                assert not forbidden.search(line), f"{path}:{line_no}: {line.strip()}"
