"""``loco_hd.loco_hd`` — the module name of the reference's native extension (src/lib.rs:9-17), served by
``loco_hd_b200._host`` (C++ over the CUDA C ABI)."""
from loco_hd_b200._host import WeightFunction, PrimitiveAtom, TagPairingRule, StatisticalDistance, LoCoHD

__all__ = ["WeightFunction", "PrimitiveAtom", "TagPairingRule", "StatisticalDistance", "LoCoHD"]
