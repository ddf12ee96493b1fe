"""seacells_b200 - B200-native engine for the SEACells hot paths (kernel construction + Frank-Wolfe fit).

    from seacells_b200 import SEACells          # same signature as SEACells.core.SEACells

The CUDA library (seacells_b200/libseacells_b200.so, built by `__graft_entry__.build()`) is required; importing the
package works without it, using it raises `SeacellsB200Error`.
"""
from ._lib import SeacellsB200Error  # noqa: F401
from .core import FitEngine, KernelMatrix, SEACells, SEACellsB200, build_kernel, knn, spmm_csr  # noqa: F401
from . import evaluate  # noqa: F401
from .summarize import sparsify_assignments, summarize_by_SEACell, summarize_by_soft_SEACell  # noqa: F401

__version__ = "0.1.0"
