"""Type stub of the extension module ``loco_hd.loco_hd`` (C++ over the CUDA C ABI ``include/locohd_b200.h``).

The first five classes are the API of the reference's PyO3 module (``/root/reference/src/lib.rs:9-17``, documented
upstream in ``loco_hd/loco_hd.pyi:7-290``): same names, positional order, keyword names and defaults.  Everything
marked *extension* is an addition of this implementation (array and resident-batch forms of the same scoring path).
All scoring runs on the GPU; without a CUDA device the scoring methods raise ``RuntimeError``.
"""
from typing import Any, Dict, List, Optional, Sequence, Tuple, Union, overload

from numpy import ndarray

VectorLike = Union[Sequence[float], ndarray]
MatrixLike = Union[Sequence[Sequence[float]], ndarray]
StringsLike = Union[Sequence[str], ndarray]

ABI_VERSION: int

def device_count() -> int:
    """*extension* Number of visible CUDA devices."""
def set_device(device: int) -> None:
    """*extension* GPU used by ``LoCoHD`` instances from their next scoring call on (default ``$LOCOHD_DEVICE`` or 0)."""
def get_device() -> int: ...

class WeightFunction:
    """Weight function of the radial integral, given by the name of its probability density and that density's
    parameters (``weight_function.rs:22-93``):

    * ``"hyper_exp"``: ``[a_1..a_n, b_1..b_n]`` (an even number of positive values), density ``sum a_i b_i exp(-b_i x) / sum a_i``
    * ``"dagum"``: ``[a, b, p]`` (non-negative)
    * ``"uniform"``: ``[x_min, x_max]`` with ``0 <= x_min < x_max``
    * ``"kumaraswamy"``: ``[x_min, x_max, a, b]`` with ``0 <= x_min < x_max`` and positive ``a``, ``b``

    Invalid names or parameters raise ``ValueError``.
    """
    parameters: List[float]
    function_name: str
    def __init__(self, function_name: str, parameters: VectorLike) -> None: ...
    def integral_point(self, point: float) -> float:
        """CDF at ``point`` (``ValueError`` for a negative point)."""
    def integral_vec(self, points: VectorLike) -> List[float]:
        """CDF at every element of ``points``."""
    def integral_range(self, point_from: float, point_to: float) -> float:
        """``CDF(point_to) - CDF(point_from)``."""

class PrimitiveAtom:
    """One labelled point: ``primitive_type`` is its category (must be one of the ``LoCoHD`` instance's
    ``categories`` when it turns up inside an environment), ``tag`` the label the ``TagPairingRule`` looks at
    (callers use ``"chain/resnum-RESNAME"``), ``coordinates`` three floats.  All three can be read and set."""
    primitive_type: str
    tag: str
    coordinates: List[float]
    def __init__(self, primitive_type: str, tag: str, coordinates: VectorLike) -> None: ...

class TagPairingRule:
    """Decides from the tags of (anchor, neighbour) whether the neighbour may enter the anchor's environment
    (the anchor itself always does).  ``variant`` is a dict with either

    * ``{"accept_same": bool}``: keep only neighbours with the same tag (``True``, the default rule of ``LoCoHD``)
      or only those with a different tag (``False``; "only hetero contacts"), or
    * ``{"tag_pairs": set[tuple[str, str]], "accepted_pairs": bool, "ordered": bool}``: the listed pairs are the
      accepted (or the rejected) ones; unordered lists also match the swapped pair.
    """
    def __init__(self, variant: Dict[str, Any]) -> None: ...
    def pair_accepted(self, pair: Tuple[str, str]) -> bool: ...
    def get_dbg_str(self) -> str: ...

class StatisticalDistance:
    """Distance between two normalised compositions: ``"Hellinger"`` ``[exponent]``, ``"Kolmogorov-Smirnov"`` ``[]``,
    ``"Kullback-Leibler"`` ``[epsilon]``, ``"Renyi"`` ``[alpha, epsilon]`` (``ValueError`` for another name or a
    wrong parameter count)."""
    def __init__(self, distance_name: str, parameters: List[float]) -> None: ...
    def run(self, p1: List[float], p2: List[float]) -> float: ...

class Structures:
    """*extension* Device-resident set of structures returned by ``LoCoHD.structures``."""
    n_structures: int
    n_primitives: int
    prim_offsets: List[int]
    def update_xyz(self, xyz: ndarray) -> None:
        """Replace all coordinates (same topology; float32 arrays travel as float32)."""
    def update_from_atoms(self, atom_xyz: ndarray, segment_start: ndarray, atom_index: ndarray,
                          first_structure: int = 0) -> None:
        """Frames of one compiled topology (``PrimitiveAssigner.compile_topology``): float32 atom coordinates
        ``[n_frames, n_atoms, 3]`` become the primitive centroids of structures ``first_structure ...`` on the device."""
    def close(self) -> None: ...

class Environments:
    """*extension* Device-resident sorted environments returned by ``LoCoHD.environments``."""
    def __len__(self) -> int: ...
    def close(self) -> None: ...

class LoCoHD:
    """Local Composition Hellinger Distance between environments of anchor points of two labelled point clouds.

    ``categories``: the primitive types; ``w_func``: one ``WeightFunction`` or a ``dict[str, WeightFunction]`` (then
    the scoring calls take one key per anchor pair); ``tag_pairing_rule``: default ``{"accept_same": True}``;
    ``n_of_threads``: accepted for compatibility (the GPU grid replaces the thread pool); ``category_weights``:
    positive weight per category (default all 1); ``statistical_distance``: default ``Hellinger [2.]``.
    """
    categories: Dict[str, int]
    category_weights: List[float]
    w_func: Union[WeightFunction, Dict[str, WeightFunction]]
    tag_pairing_rule: TagPairingRule
    def __init__(self, categories: StringsLike, w_func: Union[None, WeightFunction, Dict[str, WeightFunction]] = None,
                 tag_pairing_rule: Optional[TagPairingRule] = None, n_of_threads: Optional[int] = None,
                 category_weights: Optional[VectorLike] = None,
                 statistical_distance: Optional[StatisticalDistance] = None) -> None: ...
    def from_anchors(self, seq_a: StringsLike, seq_b: StringsLike, dists_a: VectorLike, dists_b: VectorLike,
                     w_func_key: Optional[str] = None) -> float:
        """One pair of environments given as category sequences and distances from the anchor (first distance 0;
        the lists are taken in the given order)."""
    def from_dmxs(self, seq_a: StringsLike, seq_b: StringsLike, dmx_a: MatrixLike, dmx_b: MatrixLike,
                  w_func_keys: Optional[List[str]] = None) -> List[float]:
        """Every row of the distance matrices is an anchor whose environment is the whole row (no cutoff, no tag rule)."""
    def from_coords(self, seq_a: StringsLike, seq_b: StringsLike, coords_a: MatrixLike, coords_b: MatrixLike,
                    w_func_keys: Optional[List[str]] = None) -> List[float]:
        """``from_dmxs`` on the Euclidean distance matrices of the two coordinate lists."""
    def from_primitives(self, prim_a: Sequence[PrimitiveAtom], prim_b: Sequence[PrimitiveAtom],
                        anchor_pairs: Union[Sequence[Tuple[int, int]], Sequence[Tuple[int, int, str]]],
                        threshold_distance: float) -> List[float]:
        """Scores of the anchor pairs ``(index in prim_a, index in prim_b[, weight function key])``: environments are
        the primitives closer than ``threshold_distance`` that the tag pairing rule accepts."""
    # ---- extensions: arrays in, arrays out; resident batches ---------------------------------------------------
    def from_arrays(self, xyz_a: ndarray, cat_a: ndarray, tag_a: ndarray, xyz_b: ndarray, cat_b: ndarray,
                    tag_b: ndarray, anchors: ndarray, threshold_distance: float,
                    wf_idx: Optional[Sequence[int]] = None) -> ndarray:
        """*extension* ``from_primitives`` on arrays: ``[n, 3]`` float64 coordinates, uint16 category ids
        (``category_ids``), uint32 tag ids (``intern_tags``; opaque integers for a rule without a tag list),
        ``[n_pairs, 2]`` anchors."""
    def to_arrays(self, primitives: Sequence[PrimitiveAtom]) -> Tuple[ndarray, ndarray, ndarray]:
        """*extension* ``list[PrimitiveAtom]`` -> ``(xyz [n, 3] float64, category ids uint16, tag ids uint32)``: convert a
        structure once, then use ``from_arrays`` / ``structures``."""
    def intern_tags(self, tags: StringsLike) -> ndarray:
        """*extension* Tag strings -> the uint32 ids of this instance (consistent with a ``WithList`` rule's pairs)."""
    def category_ids(self, primitive_types: StringsLike) -> ndarray:
        """*extension* Primitive type names -> uint16 category ids (0xFFFF for an unknown name)."""
    def structures(self, prim_offsets: ndarray, xyz: ndarray, categories: ndarray, tags: ndarray) -> Structures:
        """*extension* Upload concatenated structures once (structure ``s`` owns primitives
        ``prim_offsets[s]:prim_offsets[s + 1]``)."""
    def environments(self, structures: Structures, anchor_prim: ndarray, threshold_distance: float,
                     anchor_struct: Optional[ndarray] = None) -> Environments:
        """*extension* Sorted environments of the anchors ``(anchor_struct[e], anchor_prim[e])``."""
    @overload
    def score_batch(self, env_a: Environments, env_b: Environments, jobs: ndarray, reduce: None = None,
                    wf_idx: Optional[Sequence[int]] = None) -> ndarray: ...
    @overload
    def score_batch(self, env_a: Environments, env_b: Environments, jobs: ndarray,
                    reduce: Union[str, Sequence[str]], wf_idx: Optional[Sequence[int]] = None) -> Dict[str, ndarray]:
        """*extension* Runs of identity-paired environments, ``jobs`` = ``[n_jobs, 3]`` ``(a_first, b_first, n)``.
        ``reduce`` in ``"scores"``, ``"job_mean"``, ``"anchor_mean"``, ``"anchor_std"`` (device-side reductions)."""
