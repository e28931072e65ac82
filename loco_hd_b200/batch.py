"""Batches of structure pairs / trajectory frames / ensemble members over the GPUs of one box.

The scoring path has no exchange step (anchor pairs are independent): every rank owns one GPU,
scores its share of the jobs, and the per-rank score slabs are put together on the host.  This module holds the
host-side logic — dealing jobs to ranks, running a rank's share through the C ABI, assembling the results — and
nothing else.  It imports no communication library: a caller that wants the assembled results on every rank passes
its own `allgather` callable (the tests wrap torch.distributed.all_gather_object over gloo).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

JOB_DTYPE = np.dtype([("a_first", "<u8"), ("b_first", "<u8"), ("n", "<u8")])


def deal_jobs(job_sizes: Sequence[int], world: int) -> List[np.ndarray]:
    """Deal jobs to `world` ranks so that the anchor-pair totals are balanced (longest job first onto the least
    loaded rank; ties broken by rank, then job id, so every rank computes the same assignment)."""
    sizes = np.asarray(job_sizes, dtype=np.int64)
    order = np.lexsort((np.arange(len(sizes)), -sizes))
    load = np.zeros(world, dtype=np.int64)
    owner = np.empty(len(sizes), dtype=np.int64)
    for j in order:
        r = int(np.argmin(load))
        owner[j] = r
        load[r] += sizes[j]
    return [np.flatnonzero(owner == r) for r in range(world)]


def all_pairs(n: int) -> np.ndarray:
    """(i, j) with i < j, the all-vs-all job list of an ensemble of n structures (compare_ensembles.py:250-296)."""
    i, j = np.triu_indices(n, k=1)
    return np.stack([i, j], axis=1)


def blocked_pairs(n: int, block: int = 4) -> np.ndarray:
    """The same (i, j), i < j pairs in an order that keeps a working set of 2 * block structures: block x block tiles
    of the upper triangle, tile rows outermost.  Consecutive jobs then share their structures' environments, which
    stay in the 126 MB L2 of a B200 (one 5000-primitive structure has ~7.6 MB of environments at 10 A): 3.0 KB of
    DRAM reads per anchor pair in plain (i, j) order with strided CTAs, ~0.8 KB in this order with block = 4."""
    out = []
    for bi in range(0, n, block):
        for bj in range(bi, n, block):
            i = np.arange(bi, min(bi + block, n))
            j = np.arange(bj, min(bj + block, n))
            ii, jj = np.meshgrid(i, j, indexing="ij")
            keep = ii < jj
            if keep.any():
                out.append(np.stack([ii[keep], jj[keep]], axis=1))
    return np.concatenate(out) if out else np.zeros((0, 2), dtype=np.int64)


def contiguous_share(n_jobs: int, rank: int, world: int) -> slice:
    """Rank `rank`'s contiguous run of an ordered job list (equal counts within one job): keeps the locality of the
    order, unlike dealing every world-th job."""
    return slice(rank * n_jobs // world, (rank + 1) * n_jobs // world)


def assemble(per_rank_ids: Sequence[np.ndarray], per_rank_values: Sequence[np.ndarray], n_jobs: int) -> np.ndarray:
    """Per-rank (job ids, values) -> one array indexed by job id."""
    out = np.full(n_jobs, np.nan, dtype=np.float64)
    for ids, vals in zip(per_rank_ids, per_rank_values):
        out[np.asarray(ids, dtype=np.int64)] = np.asarray(vals, dtype=np.float64)
    return out


def gather_job_values(my_ids: np.ndarray, my_values: np.ndarray, n_jobs: int,
                      allgather: Optional[Callable[[object], Sequence[object]]] = None) -> np.ndarray:
    """All ranks get the values of all jobs.  `allgather(obj) -> [obj of rank 0, obj of rank 1, ...]` is whatever
    the launcher offers for exchanging small host objects (e.g. a wrapper of torch.distributed.all_gather_object or
    mpi4py's allgather); this module itself depends on no communication library.  None = single process."""
    if allgather is None:
        return assemble([my_ids], [my_values], n_jobs)
    gathered = allgather((np.asarray(my_ids), np.asarray(my_values)))
    return assemble([g[0] for g in gathered], [g[1] for g in gathered], n_jobs)


def run_sharded(job_sizes: Sequence[int], scorer: Callable[[np.ndarray], np.ndarray], rank: int, world: int,
                gather: bool = True, allgather: Optional[Callable[[object], Sequence[object]]] = None) -> np.ndarray:
    """Score this rank's share with `scorer(job_ids) -> one value per job` and (optionally) gather everything."""
    mine = deal_jobs(job_sizes, world)[rank]
    values = np.asarray(scorer(mine), dtype=np.float64) if len(mine) else np.zeros(0)
    if not gather or world == 1:
        return assemble([mine], [values], len(job_sizes))
    if allgather is None:
        raise ValueError("gather=True with world > 1 needs an `allgather` callable")
    return gather_job_values(mine, values, len(job_sizes), allgather)


def combine_anchor_stats(counts: Sequence[int], means: Sequence[np.ndarray], stds: Sequence[np.ndarray]):
    """Per-rank per-anchor (job count, mean, population std) -> the statistics over all jobs (exact pooling:
    total variance = mean of the within-rank variances + variance of the rank means, weighted by job counts)."""
    n = np.asarray(counts, dtype=np.float64)
    m = np.stack([np.asarray(x, dtype=np.float64) for x in means])
    s = np.stack([np.asarray(x, dtype=np.float64) for x in stds])
    total = n.sum()
    mean = (n[:, None] * m).sum(axis=0) / total
    var = (n[:, None] * (s * s + (m - mean) ** 2)).sum(axis=0) / total
    return mean, np.sqrt(var)


@dataclass
class ResidentEnsemble:
    """An ensemble resident on one GPU: structures uploaded once, one environment per (structure, anchor)."""

    ctx: object
    structs: object
    env: object
    anchors_per_structure: int

    @classmethod
    def build(cls, ctx, clouds, anchors: np.ndarray, threshold: float) -> "ResidentEnsemble":
        """clouds: objects with .xyz [n,3] f64, .cat u16, .tag u32 (same topology); anchors: primitive indices used in
        every structure."""
        offs = np.cumsum([0] + [len(c.cat) for c in clouds]).astype(np.uint64)
        st = ctx.structs_create(offs, np.concatenate([c.xyz for c in clouds]), np.concatenate([c.cat for c in clouds]),
                                np.concatenate([c.tag for c in clouds]))
        anchors = np.asarray(anchors, dtype=np.uint32)
        a_struct = np.repeat(np.arange(len(clouds), dtype=np.uint32), len(anchors))
        env = ctx.envset_build(st, np.tile(anchors, len(clouds)), threshold, anchor_struct=a_struct)
        return cls(ctx, st, env, len(anchors))

    def job_table(self, pairs: np.ndarray) -> np.ndarray:
        n = self.anchors_per_structure
        jobs = np.empty(len(pairs), dtype=JOB_DTYPE)
        jobs["a_first"] = np.asarray(pairs)[:, 0].astype(np.uint64) * n
        jobs["b_first"] = np.asarray(pairs)[:, 1].astype(np.uint64) * n
        jobs["n"] = n
        return jobs

    def pair_means(self, pairs: np.ndarray, chunk: int = 4096) -> np.ndarray:
        """Mean LoCoHD of every structure pair (the `lchd_dmx` entries of compare_ensembles.py:293-299), reduced on
        the device so that only one number per pair crosses the bus."""
        out = np.empty(len(pairs), dtype=np.float64)
        for lo in range(0, len(pairs), chunk):
            jobs = self.job_table(pairs[lo:lo + chunk])
            _, means = self.ctx.score_jobs(self.env, self.env, jobs, want_scores=False, want_means=True)
            out[lo:lo + chunk] = means
        return out

    def pair_means_and_anchor_stats(self, pairs: np.ndarray, chunk: int = 4096):
        """Per structure pair the mean score, and per anchor the mean and population standard deviation over all the
        given structure pairs (`lchd_dmx` entries and `lchd_by_atom` of compare_ensembles.py:293-299), all reduced on
        the device; chunks of structure pairs are pooled exactly on the host."""
        means = np.empty(len(pairs), dtype=np.float64)
        counts, amean, astd = [], [], []
        for lo in range(0, len(pairs), chunk):
            jobs = self.job_table(pairs[lo:lo + chunk])
            res = self.ctx.score_jobs_stats(self.env, self.env, jobs, job_means=True, anchor_means=True, anchor_stds=True)
            means[lo:lo + chunk] = res["job_means"]
            counts.append(len(jobs)); amean.append(res["anchor_means"]); astd.append(res["anchor_stds"])
        if not counts:
            empty = np.zeros(self.anchors_per_structure)
            return means, empty, empty
        mean, std = combine_anchor_stats(counts, amean, astd)
        return means, mean, std

    def close(self):
        self.env.close()
        self.structs.close()


def ensemble_all_vs_all(ctx, clouds, anchors, threshold: float, rank: int = 0, world: int = 1,
                        gather: bool = True, allgather=None) -> np.ndarray:
    """All-vs-all mean LoCoHD matrix entries (i < j, order of `all_pairs`) of an ensemble, sharded over `world`
    ranks: every rank keeps the whole ensemble resident (1000 x 5000 primitives = 145 MB) and scores one contiguous
    run of the structure pairs taken in `blocked_pairs` order (L2-friendly); entries of other ranks are NaN unless
    gathered."""
    n = len(clouds)
    pairs = all_pairs(n)
    order = blocked_pairs(n)
    i, j = order[:, 0].astype(np.int64), order[:, 1].astype(np.int64)
    ids = i * n - i * (i + 1) // 2 + (j - i - 1)          # position of (i, j) in the all_pairs order
    mine = ids[contiguous_share(len(ids), rank, world)]
    ens = ResidentEnsemble.build(ctx, clouds, anchors, threshold)
    try:
        values = ens.pair_means(pairs[mine]) if len(mine) else np.zeros(0)
    finally:
        ens.close()
    if not gather or world == 1:
        return assemble([mine], [values], len(pairs))
    if allgather is None:
        raise ValueError("gather=True with world > 1 needs an `allgather` callable")
    return gather_job_values(mine, values, len(pairs), allgather)


def trajectory_scores(lchd, topology, categories, tags, frame0_atoms, frame_blocks, anchors, threshold: float,
                      reduce: Optional[Sequence[str]] = None, max_block: Optional[int] = None):
    """Frame 0 of a trajectory against every later frame, from ATOM coordinates to scores on the device: what
    trajectory_analyzer.py:89-129 does per frame with `assign_from_universe` + `from_primitives`, without the per-frame
    regex pass and without re-uploading frame 0.

    lchd          a loco_hd.LoCoHD instance (public class)
    topology      PrimitiveAssigner.compile_topology(structure): which atoms feed which primitive
    categories    uint16 category id of every primitive (lchd.category_ids(topology.primitive_types))
    tags          uint32 tag id of every primitive (lchd.intern_tags(...), one tag per residue)
    frame0_atoms  [n_atoms, 3] float32
    frame_blocks  iterable of [B, n_atoms, 3] float32 blocks of frames (B may vary up to max_block / the first block's B)
    anchors       primitive indices used as anchors in frame 0 and in every frame (e.g. the "Cent" primitives)
    reduce        None -> yields one [B, n_anchors] score array per block; or names understood by LoCoHD.score_batch
                  ("anchor_mean", "anchor_std", "job_mean", "scores") -> yields the dict of that block

    Generator: one result per block, in order."""
    categories = np.ascontiguousarray(categories, dtype=np.uint16)
    tags = np.ascontiguousarray(tags, dtype=np.uint32)
    anchors = np.ascontiguousarray(anchors, dtype=np.uint32)
    seg = np.ascontiguousarray(topology.segment_start, dtype=np.uint32)
    idx = np.ascontiguousarray(topology.atom_index, dtype=np.uint32)
    n_prims = len(categories)
    if len(seg) != n_prims + 1 or len(tags) != n_prims:
        raise ValueError("categories / tags must have one entry per primitive of the topology")
    blocks = iter(frame_blocks)
    first = next(blocks, None)
    if first is None:
        return
    first = np.ascontiguousarray(first, dtype=np.float32)
    cap = int(max_block or first.shape[0])
    offsets = (np.arange(cap + 2, dtype=np.uint64) * n_prims)
    st = lchd.structures(offsets, np.zeros(((cap + 1) * n_prims, 3), dtype=np.float32), np.tile(categories, cap + 1),
                         np.tile(tags, cap + 1))
    env0 = None
    try:
        st.update_from_atoms(np.ascontiguousarray(frame0_atoms, dtype=np.float32), seg, idx, first_structure=0)
        env0 = lchd.environments(st, anchors, threshold)                      # frame 0: built once
        block = first
        while block is not None:
            b = block.shape[0]
            if b > cap:
                raise ValueError(f"a block of {b} frames exceeds max_block = {cap}")
            st.update_from_atoms(block, seg, idx, first_structure=1)
            env = lchd.environments(st, np.tile(anchors, b), threshold,
                                    anchor_struct=np.repeat(np.arange(1, b + 1, dtype=np.uint32), len(anchors)))
            jobs = np.array([(0, f * len(anchors), len(anchors)) for f in range(b)], dtype=np.uint64)
            try:
                out = lchd.score_batch(env0, env, jobs, reduce=reduce)
            finally:
                env.close()
            yield out.reshape(b, len(anchors)) if reduce is None else out
            nxt = next(blocks, None)
            block = None if nxt is None else np.ascontiguousarray(nxt, dtype=np.float32)
    finally:
        if env0 is not None:
            env0.close()
        st.close()


def pair_anchors_by_tag(ref_cat, ref_tag, model_cat, model_tag, anchor_category: int) -> np.ndarray:
    """Anchor pairs (reference index, model index) of one model: every reference primitive of the anchor category
    whose tag also carries an anchor primitive in the model, in reference order - the pairing loop of
    casp14_extend_with_locohd.py:48,66-79 (a model may lack residues; a tag that occurs twice in the model pairs
    with its last anchor primitive, as the caller's dict comprehension does)."""
    ref_cat, model_cat = np.asarray(ref_cat), np.asarray(model_cat)
    ref_tag, model_tag = np.asarray(ref_tag), np.asarray(model_tag)
    ref_anchor = np.flatnonzero(ref_cat == anchor_category)
    model_anchor = np.flatnonzero(model_cat == anchor_category)
    if len(ref_anchor) == 0 or len(model_anchor) == 0:
        return np.zeros((0, 2), dtype=np.uint32)
    mt = model_tag[model_anchor]
    order = np.argsort(mt, kind="stable")
    mt_sorted = mt[order]
    pos = np.searchsorted(mt_sorted, ref_tag[ref_anchor], side="right") - 1      # last occurrence
    hit = (pos >= 0) & (mt_sorted[np.maximum(pos, 0)] == ref_tag[ref_anchor])
    return np.stack([ref_anchor[hit], model_anchor[order[pos[hit]]]], axis=1).astype(np.uint32)


def models_against_reference(lchd, reference, models, anchor_category: int, threshold: float):
    """One reference structure against many models of it, per-residue scores and the per-model mean: the loop of
    casp14_extend_with_locohd.py:44-88 (BASELINE config 3) as one resident batch - the reference is uploaded once,
    all environments come from two gather launches and all models are scored by one call; the per-model means are
    reduced on the device.

    lchd             a loco_hd.LoCoHD instance (public class)
    reference        (xyz [n, 3], category ids uint16, tag ids uint32) - LoCoHD.to_arrays(list of PrimitiveAtom)
    models           sequence of such triples; sizes and residue sets may differ from the reference's
    anchor_category  category id of the anchor primitives (lchd.category_ids(["Cent"])[0])

    Returns a list with one (anchor_pairs [p, 2] uint32, scores [p] float64, mean float) per model; a model without
    common anchors gets empty arrays and nan."""
    models = list(models)
    if not models:
        return []
    rx, rc, rt = (np.asarray(v) for v in reference)
    pairs = [pair_anchors_by_tag(rc, rt, m[1], m[2], anchor_category) for m in models]
    sizes = [len(rc)] + [len(m[1]) for m in models]
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    f32 = all(np.asarray(v[0]).dtype == np.float32 for v in [reference] + models)
    xyz = np.concatenate([np.asarray(v[0], dtype=np.float32 if f32 else np.float64).reshape(-1, 3)
                          for v in [reference] + models])
    cats = np.concatenate([np.asarray(v[1], dtype=np.uint16) for v in [reference] + models])
    tags = np.concatenate([np.asarray(v[2], dtype=np.uint32) for v in [reference] + models])
    counts = np.array([len(p) for p in pairs], dtype=np.uint64)
    scored = np.flatnonzero(counts)
    out = [(p, np.zeros(0), float("nan")) for p in pairs]
    if len(scored) == 0:
        return out
    a_prim = np.concatenate([pairs[k][:, 0] for k in scored])
    b_prim = np.concatenate([pairs[k][:, 1] for k in scored])
    b_struct = np.repeat((scored + 1).astype(np.uint32), counts[scored].astype(np.int64))
    first = np.concatenate([[0], np.cumsum(counts[scored])[:-1]]).astype(np.uint64)
    jobs = np.stack([first, first, counts[scored]], axis=1).astype(np.uint64)
    st = lchd.structures(offsets, xyz, cats, tags)
    env_a = env_b = None
    try:
        env_a = lchd.environments(st, a_prim, threshold, anchor_struct=np.zeros(len(a_prim), dtype=np.uint32))
        env_b = lchd.environments(st, b_prim, threshold, anchor_struct=b_struct)
        red = lchd.score_batch(env_a, env_b, jobs, reduce=["scores", "job_mean"])
    finally:
        for h in (env_a, env_b, st):
            if h is not None:
                h.close()
    for n, k in enumerate(scored):
        lo = int(first[n])
        out[k] = (pairs[k], red["scores"][lo:lo + int(counts[k])], float(red["job_mean"][n]))
    return out
