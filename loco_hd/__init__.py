"""Drop-in package for the reference's ``loco_hd`` (its ``loco_hd/__init__.py:1-2`` import surface): the five
native classes come from the CUDA-backed host module, the helpers from the Python side of ``loco_hd_b200``."""
from .loco_hd import WeightFunction, PrimitiveAtom, TagPairingRule, LoCoHD, StatisticalDistance
from loco_hd_b200.atom_converter_utils import (PrimitiveAssigner, PrimitiveAtomTemplate, PrimitiveAtomSource,
                                               TypingSchemeElement)
