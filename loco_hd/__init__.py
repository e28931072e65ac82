"""Drop-in package for the reference's ``loco_hd`` (its ``loco_hd/__init__.py:1-2`` import surface): the five
native classes come from the extension module ``loco_hd.loco_hd`` (C++ over the CUDA C ABI, built in-tree by
``loco_hd_b200/build.py``), the helpers from the Python side of ``loco_hd_b200``."""
try:
    from .loco_hd import WeightFunction, PrimitiveAtom, TagPairingRule, LoCoHD, StatisticalDistance
except ImportError as exc:  # fail loudly: nothing else can score
    raise ImportError(
        "loco_hd.loco_hd (the CUDA-backed extension module) is not built: run `python loco_hd_b200/build.py` or "
        "`python -c 'import __graft_entry__ as g; g.build()'` from the repository root (needs nvcc for sm_100a and "
        "g++).  There is no CPU fallback.") from exc
from loco_hd_b200.atom_converter_utils import (PrimitiveAssigner, PrimitiveAtomTemplate, PrimitiveAtomSource,
                                               TypingSchemeElement)
