"""Worker of tests/test_multi_gpu_sharding.py: one process per rank, gloo backend on CPU."""
import os
import sys
from pathlib import Path

import numpy as np
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from loco_hd_b200 import batch  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    rng = np.random.default_rng(0)
    sizes = rng.integers(1, 5000, size=101)
    truth = np.sqrt(np.arange(len(sizes)) + 0.25) * sizes          # what a scorer would return per job
    calls = []

    def scorer(ids):
        calls.append(np.array(ids))
        return truth[ids]

    def allgather(obj):   # the module itself carries no communication library: the launcher supplies the exchange
        box = [None] * world
        dist.all_gather_object(box, obj)
        return box

    got = batch.run_sharded(sizes, scorer, rank, world, gather=True, allgather=allgather)
    assert np.array_equal(got, truth), "gathered values differ"
    mine = batch.deal_jobs(sizes, world)[rank]
    assert len(calls) == 1 and np.array_equal(calls[0], mine)
    # every rank computed the same assignment, every job has exactly one owner
    owners = [None] * world
    dist.all_gather_object(owners, mine)
    allj = np.sort(np.concatenate(owners))
    assert np.array_equal(allj, np.arange(len(sizes)))
    loads = [int(sizes[o].sum()) for o in owners]
    assert max(loads) - min(loads) <= sizes.max()
    # without gathering a rank only sees its own jobs
    local = batch.run_sharded(sizes, scorer, rank, world, gather=False)
    assert np.array_equal(np.flatnonzero(~np.isnan(local)), mine)
    dist.barrier()
    if rank == 0:
        Path(os.environ["LOCOHD_TEST_OUT"]).write_text("ok")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
