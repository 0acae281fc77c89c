"""Multi-process strip path on CPU (gloo, world size 2): the row partition and the gather to the
presenting rank.  Each rank "renders" its strip with the CPU oracle (tests may use it), then the
same isend/irecv pattern StripRenderer uses with NCCL assembles the canvas on rank 0."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, result_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gudni_b200 import scenes
    from gudni_b200.strips import StripRenderer, partition_rows
    from oracle import oracle

    scene = scenes.fuzzy_circles(600, 700, 900, 5, 50, 0x61006)
    rows = partition_rows(scene, world, 256)
    y0, y1 = rows[rank]
    # the strip a rank owns only needs the shapes that touch its rows (what the rank would bin)
    import copy
    local = copy.copy(scene)
    local.entries = scene.subset_rows(y0, y1)
    image = oracle.render(local, taps=False, threads=2).image
    strip = torch.from_numpy(image[y0:y1].view(np.int32).copy())
    canvas = torch.zeros((scene.height, scene.width), dtype=torch.int32) if rank == 0 else None
    if rank == 0:
        canvas[y0:y1] = strip
    StripRenderer.gather_strips(dist, rank, 0, rows, strip, canvas)
    if rank == 0:
        full = oracle.render(scene, taps=False, threads=2).image
        np.save(result_path, np.array([int(np.array_equal(canvas.numpy().view(np.uint32), full))]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_strip_gather(tmp_path):
    port = 29600 + (os.getpid() % 300)
    result = str(tmp_path / "ok.npy")
    mp.spawn(_worker, args=(2, port, result), nprocs=2, join=True)
    assert np.load(result)[0] == 1


def test_partition_rows_is_contiguous_and_balanced():
    sys.path.insert(0, ROOT)
    from gudni_b200 import scenes
    from gudni_b200.strips import partition_rows
    scene = scenes.fuzzy_circles(2000, 4096, 4096, 20, 200, 5)
    for n in (1, 2, 4, 8, 16, 32):
        rows = partition_rows(scene, n, 256)
        assert rows[0][0] == 0 and rows[-1][1] == scene.height
        assert all(a[1] == b[0] and a[0] % 256 == 0 for a, b in zip(rows, rows[1:]))
        assert len(rows) == min(n, 16)
