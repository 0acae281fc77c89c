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


def _worker(rank, world, port, result_path, early):
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
    if early:
        # the presenting rank posts its receives before it has produced its own strip (StripRenderer.render)
        reqs = StripRenderer.post_receives(dist, rank, rows, canvas) if rank == 0 else None
        if rank == 0:
            canvas[y0:y1] = strip
            for q in reqs:
                q.wait()
        else:
            StripRenderer.gather_strips(dist, rank, 0, rows, strip, canvas)
    else:
        if rank == 0:
            canvas[y0:y1] = strip
        StripRenderer.gather_strips(dist, rank, 0, rows, strip, canvas)
    if rank == 0:
        full = oracle.render(scene, taps=False, threads=2).image
        np.save(result_path, np.array([int(np.array_equal(canvas.numpy().view(np.uint32), full))]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("early", [False, True], ids=["gather-after", "receives-posted-first"])
def test_two_rank_strip_gather(tmp_path, early):
    port = 29600 + (os.getpid() % 300) + (7 if early else 0)
    result = str(tmp_path / "ok.npy")
    mp.spawn(_worker, args=(2, port, result, early), nprocs=2, join=True)
    assert np.load(result)[0] == 1


def test_rebalance_rows_staggers_the_finishing_times():
    sys.path.insert(0, ROOT)
    from gudni_b200.strips import rebalance_rows
    height, tile = 64 * 256, 256
    even = [(k * 8 * tile, (k + 1) * 8 * tile) for k in range(8)]
    times = [4.0] * 8                      # uniform cost: 0.5 ms per tile row
    flat = rebalance_rows(even, times, height, tile)
    assert flat == even                    # nothing to gain without a transfer cost
    stag = rebalance_rows(even, times, height, tile, presenting=0, row_transfer_ms=0.02)
    assert stag[0][0] == 0 and stag[-1][1] == height
    assert all(a[1] == b[0] for a, b in zip(stag, stag[1:]))
    sizes = [(b - a) // tile for a, b in stag]
    # lower strips finish later (their data queues behind the strips above), so they may be larger
    assert sizes[1] <= sizes[-1] and sizes != [8] * 8
    done = [0.5 * n + (64 - a // tile) * 0.02 * (k != 0) for k, ((a, b), n) in enumerate(zip(stag, sizes))]
    done_even = [4.0 + (64 - 8 * k) * 0.02 * (k != 0) for k in range(8)]
    assert max(done) < max(done_even)


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


def test_rebalance_rows_late_receives_charges_the_presenting_rank():
    sys.path.insert(0, ROOT)
    from gudni_b200.strips import rebalance_rows
    height, tile = 64 * 256, 256
    halves = [(0, 32 * tile), (32 * tile, height)]
    # two ranks, uniform cost, receives posted after the presenting rank's strip: both must finish together
    # (the transfer follows whichever is later), so the halves stay
    assert rebalance_rows(halves, [10.0, 10.0], height, tile, presenting=0, row_transfer_ms=0.02, late_receives=True) == halves
    # early receives: the lower strip may finish later by less than its own transfer time
    early = rebalance_rows(halves, [10.0, 10.0], height, tile, presenting=0, row_transfer_ms=0.02)
    assert early[0][1] >= halves[0][1]
