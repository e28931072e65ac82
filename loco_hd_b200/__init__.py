"""loco_hd_b200 — B200-native LoCoHD per-anchor scoring path (CUDA sm_100a behind a C ABI).

Layout:
    csrc/                 CUDA kernels + C ABI (include/locohd_b200.h) + the CPython host module source
    (loco_hd/loco_hd.*.so C++ CPython extension with the reference's five classes, built in-tree by build.py at the
                          reference's module path; re-exported here)
    _capi                 ctypes binding of the C ABI (parity tests, bench)
    atom_converter_utils  PrimitiveAssigner and its dataclasses (Python side of the reference API)
    batch                 structure-pair / frame / ensemble batches sharded over the GPUs of one box

There is no CPU fallback: the scoring classes need the built extension, and every scoring call needs a CUDA
device (a missing extension raises ImportError here, a missing device raises at the first scoring call).
"""
try:
    from loco_hd.loco_hd import (LoCoHD, PrimitiveAtom, StatisticalDistance, TagPairingRule, WeightFunction,
                                 device_count, get_device, set_device)
except ImportError as exc:  # fail loudly: nothing else can score
    raise ImportError(
        "loco_hd.loco_hd (the CUDA-backed extension module) is not built: run `python loco_hd_b200/build.py` or "
        "`python -c 'import __graft_entry__ as g; g.build()'` from the repository root (needs nvcc for sm_100a and "
        "g++). loco_hd_b200 has no CPU fallback.") from exc

from .atom_converter_utils import PrimitiveAssigner, PrimitiveAtomSource, PrimitiveAtomTemplate, TypingSchemeElement

__all__ = [
    "LoCoHD", "PrimitiveAtom", "StatisticalDistance", "TagPairingRule", "WeightFunction", "PrimitiveAssigner",
    "PrimitiveAtomSource", "PrimitiveAtomTemplate", "TypingSchemeElement", "device_count", "get_device", "set_device",
]
