"""CPU: the C-ABI library builds, loads, and exports every symbol include/emerge_b200.h declares."""
import os
import re

from emerge_b200 import lib as L

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_header_symbols():
    from emerge_b200.build import build
    build()
    lib = L.load_library()
    hdr = open(os.path.join(REPO, "include", "emerge_b200.h")).read()
    declared = set(re.findall(r"\b(emb_[a-z_A-Z0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(L.exported_symbols()), declared ^ set(L.exported_symbols())
    assert b"sm_100a" in lib.emb_version()


def test_no_gpu_is_a_loud_error():
    import torch
    if torch.cuda.is_available():
        return
    try:
        L.Context(0)
    except L.EmergeB200Error as e:
        assert "no CPU fallback" in str(e) or "CPU fallback" in str(e)
    else:
        raise AssertionError("Context() must fail without a GPU")


def test_product_never_imports_oracle():
    pkg = os.path.join(REPO, "emerge_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(root, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
