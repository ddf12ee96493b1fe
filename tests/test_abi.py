"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every declared symbol."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "seacells_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sc_[a-z_A-Z0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    from seacells_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        g.build()
    lib = _lib.load()
    syms = _header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/seacells_b200.h but not exported"
    assert set(_lib.SIGNATURES) == set(syms), set(_lib.SIGNATURES) ^ set(syms)
    assert b"sm_100a" in lib.sc_version()


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from seacells_b200 import SEACells, SeacellsB200Error
    from seacells_b200.synth import MiniAnnData, synth_embedding
    ad = MiniAnnData(synth_embedding(300, 10))
    with pytest.raises(SeacellsB200Error):
        SEACells(ad, "X_pca", 5, use_sparse=True)


def test_factory_contract():
    """core.py:44-46: use_sparse with use_gpu is an AssertionError."""
    from seacells_b200 import SEACells
    from seacells_b200.synth import MiniAnnData, synth_embedding
    ad = MiniAnnData(synth_embedding(100, 5))
    with pytest.raises(AssertionError):
        SEACells(ad, "X_pca", 5, use_gpu=True, use_sparse=True)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "seacells_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
