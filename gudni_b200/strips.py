"""Multi-GPU strip partition: one process per GPU, each rank bins and rasterizes
a contiguous run of root-tile rows; the finished strips are gathered on the presenting rank.

The reference has no counterpart (one OpenCLState = one device, OpenCL/Setup.hs:118-120).  Tiles and
columns are independent, so the path shards with no data-path exchange except the final gather of
pixels.  Two gather modes:

  "nccl"  every rank renders into its own strip tensor; non-presenting ranks `isend` it, the
          presenting rank `irecv`s straight into the row slice of the canvas (NCCL over NVLink).
  "p2p"   the presenting rank's canvas is mapped into every process with CUDA IPC and the raster
          kernel stores its pixels there directly, so the transfer overlaps the rasterization and
          no separate gather pass exists; a barrier closes the frame.
"""
import numpy as np


def partition_rows(scene, n_ranks, tile_rows=256):
    """Contiguous runs of root-tile rows per rank, balanced by a cheap work estimate: shape-box
    area clipped to each tile row plus the row's pixel count (covers empty rows)."""
    n_rows = (scene.height + tile_rows - 1) // tile_rows
    e = scene.entries
    weight = np.zeros(n_rows, dtype=np.float64)
    if len(e):
        top = np.clip(e["top"].astype(np.float64), 0, scene.height)
        bottom = np.clip(e["bottom"].astype(np.float64), 0, scene.height)
        width = np.clip(e["right"], 0, scene.width).astype(np.float64) - np.clip(e["left"], 0, scene.width)
        for r in range(n_rows):
            y0, y1 = r * tile_rows, min((r + 1) * tile_rows, scene.height)
            weight[r] = float((np.clip(np.minimum(bottom, y1) - np.maximum(top, y0), 0, None) * width).sum())
    weight += float(scene.width) * tile_rows * 0.5
    n_ranks = min(n_ranks, n_rows)
    cum = np.concatenate([[0.0], np.cumsum(weight)])
    bounds = [0]
    for k in range(1, n_ranks):
        target = cum[-1] * k / n_ranks
        idx = int(np.searchsorted(cum, target))
        idx = max(bounds[-1] + 1, min(idx, n_rows - (n_ranks - k)))
        bounds.append(idx)
    bounds.append(n_rows)
    return [(bounds[k] * tile_rows, min(bounds[k + 1] * tile_rows, scene.height)) for k in range(n_ranks)]


def rebalance_rows(rows, times_ms, height, tile_rows=256, presenting=0, presenting_extra_ms=0.0, row_transfer_ms=0.0,
                   late_receives=False):
    """Feedback partition: given the strips of the last frame and the time each rank spent on its strip,
    assume cost is uniform inside a strip and cut the canvas again (contiguous tile rows, one strip per
    rank, in rank order) so that the frame completes as early as possible.  Exact by dynamic programming.

    Cost model.  The presenting rank posts its receives before it starts its own strip, so a strip starts
    to arrive as soon as its rank has finished it, and the strips share the presenting rank's inbound
    links: with `row_transfer_ms` per tile row, a strip that starts at tile row i cannot be complete
    before its own rank is done plus the time to move everything from row i down (the ranks below it
    finish later and queue behind it).  Minimising the latest of those makes the ranks finish staggered,
    top to bottom, each one just as the link frees up, instead of all together with the whole gather
    still to go.  With `late_receives` the presenting rank posts its receives after its own strip:
    nothing can arrive before it is done, so it is charged the transfer of every other strip as well.
    `presenting_extra_ms` is a flat charge on the presenting rank."""
    n_rows = (height + tile_rows - 1) // tile_rows
    cost = np.zeros(n_rows)
    for (y0, y1), t in zip(rows, times_ms):
        a, b = y0 // tile_rows, (y1 + tile_rows - 1) // tile_rows
        if b > a:
            cost[a:b] = max(t, 1e-3) / (b - a)
    n = len(rows)
    if n_rows < n:
        return list(rows)
    pre = np.concatenate([[0.0], np.cumsum(cost)])
    INF = float("inf")
    best = [[INF] * (n_rows + 1) for _ in range(n + 1)]
    cut = [[0] * (n_rows + 1) for _ in range(n + 1)]
    best[0][0] = 0.0
    for k in range(1, n + 1):
        is_presenting = (k - 1) == presenting
        for j in range(k, n_rows - (n - k) + 1):
            for i in range(k - 1, j):
                finish = pre[j] - pre[i]
                if is_presenting:
                    finish += presenting_extra_ms
                    if late_receives:
                        finish += (n_rows - (j - i)) * row_transfer_ms
                else:
                    finish += (n_rows - i) * row_transfer_ms if presenting == 0 else (j - i) * row_transfer_ms
                v = max(best[k - 1][i], finish)
                if v < best[k][j]:
                    best[k][j], cut[k][j] = v, i
    bounds = [n_rows]
    for k in range(n, 0, -1):
        bounds.append(cut[k][bounds[-1]])
    bounds = bounds[::-1]
    return [(bounds[k] * tile_rows, min(bounds[k + 1] * tile_rows, height)) for k in range(n)]


class StripRenderer:
    """Per-rank driver.  `dist` is torch.distributed (already initialised) or None for one GPU."""

    def __init__(self, rasterizer, scene, rank=0, world=1, dist=None, mode="nccl", presenting_rank=0, rows=None,
                 early_receives=True):
        import torch
        self.torch = torch
        self.r, self.scene, self.rank, self.world, self.dist = rasterizer, scene, rank, world, dist
        self.mode = mode if world > 1 else "local"
        # receives posted before the presenting rank's own strip (the strips then arrive while it rasterizes, at
        # the price of the receive kernel holding some SMs meanwhile) or after it
        self.early_receives = early_receives
        self.presenting = presenting_rank
        self.rows = rows if rows is not None else partition_rows(scene, world, rasterizer.spec.max_tile_size)
        while len(self.rows) < world:
            self.rows.append((scene.height, scene.height))   # more ranks than tile rows: idle ranks
        self.my_rows = self.rows[rank]
        self.entries = scene.subset_rows(*self.my_rows) if self.my_rows[1] > self.my_rows[0] else scene.entries[:0]
        dev = torch.device("cuda", torch.cuda.current_device())
        self.canvas = None
        self.strip = None
        self._peer_canvas = None
        h, w = scene.height, scene.width
        if rank == presenting_rank:
            self.canvas = torch.empty((h, w), dtype=torch.int32, device=dev)
        if self.mode == "p2p":
            self._setup_p2p(dev)
        elif rank != presenting_rank:
            rows = self.my_rows[1] - self.my_rows[0]
            self.strip = torch.empty((max(rows, 1), w), dtype=torch.int32, device=dev)

    # -- CUDA IPC: map the presenting rank's canvas into this process --------------------------------
    def _setup_p2p(self, dev):
        import ctypes
        torch, dist = self.torch, self.dist
        handle = torch.zeros(64, dtype=torch.uint8)
        if self.rank == self.presenting:
            h = (ctypes.c_ubyte * 64)()
            # export the torch-owned canvas through the runtime directly (cudaIpcGetMemHandle)
            rt = ctypes.CDLL("libcudart.so.12")
            rc = rt.cudaIpcGetMemHandle(ctypes.byref(h), ctypes.c_void_p(self.canvas.data_ptr()))
            if rc != 0:
                raise RuntimeError(f"cudaIpcGetMemHandle failed: {rc}")
            handle = torch.frombuffer(bytearray(h), dtype=torch.uint8).clone()
        hd = handle.to(dev)
        dist.broadcast(hd, src=self.presenting)
        if self.rank != self.presenting:
            raw = bytes(hd.cpu().numpy().tobytes())
            p = ctypes.c_void_p()
            buf = (ctypes.c_ubyte * 64).from_buffer_copy(raw)
            self.r._check(self.r._L.gudni_b200_ipc_open(self.r._ctx, buf, ctypes.byref(p)))
            self._peer_canvas = p.value

    # -- pipelined variant: chunks of the strip are sent while the next chunk is being rasterized -------
    def prepare_chunks(self, dscene_factory, chunk_rows=None):
        """Split this rank's strip into chunks of whole tile rows; `dscene_factory(entries)` uploads a
        chunk's shape entries and returns (device pointer, count)."""
        tile = self.r.spec.max_tile_size
        chunk_rows = chunk_rows or tile
        self.chunks = []
        for k, (y0, y1) in enumerate(self.rows):
            parts = [(a, min(a + chunk_rows, y1)) for a in range(y0, y1, chunk_rows)]
            self.chunks.append(parts)
        self.chunk_entries = []
        for (a, b) in self.chunks[self.rank]:
            self.chunk_entries.append(dscene_factory(self.scene.subset_rows(a, b)))

    def render_pipelined(self, frame, dscene):
        """Like render(), but the strip goes out chunk by chunk: NCCL carries chunk c to the presenting
        rank while chunk c+1 is in the raster kernels, so only the last chunk's transfer is exposed."""
        r, dist = self.r, self.dist
        reqs = []
        if self.rank == self.presenting:
            r.frame_target(self.canvas.data_ptr(), 0)
            for k, parts in enumerate(self.chunks):
                if k == self.rank:
                    continue
                for (a, b) in parts:
                    reqs.append(dist.irecv(self.canvas[a:b], k))
        else:
            r.frame_target(self.strip.data_ptr(), self.my_rows[0])
        for (a, b), (dev_entries, n) in zip(self.chunks[self.rank], self.chunk_entries):
            r.frame_begin_device(dscene, frame)
            r.frame_strip(a, b)
            r.raster_entries_device(dev_entries, n)
            _, self.last_stats = r.frame_end(want_image=False)
            if self.rank != self.presenting:
                reqs.append(dist.isend(self.strip[a - self.my_rows[0]: b - self.my_rows[0]], self.presenting))
        for q in reqs:
            q.wait()
        return self.canvas

    def close(self):
        if self._peer_canvas:
            import ctypes
            self.r._L.gudni_b200_ipc_close(self.r._ctx, ctypes.c_void_p(self._peer_canvas))
            self._peer_canvas = None
        self.r.frame_target(None)

    # -- one frame ---------------------------------------------------------------------------------
    def render_to_host(self, frame, host_canvas):
        """The host-presenter path: this rank rasterizes its strip from host buffers and copies it over its own
        PCIe link straight into its rows of `host_canvas` (a (H, W) uint32 array every rank maps — POSIX shared
        memory, page-locked by each rank).  No inter-GPU traffic; the caller's barrier closes the frame."""
        r = self.r
        rows = self.my_rows
        if rows[1] <= rows[0]:
            return
        r.frame_target(None)
        r.frame_begin(self.scene, frame)
        if self.world > 1:
            r.frame_strip(*rows)
        r.raster_entries(self.entries)
        _, self.last_stats = r.frame_end(out=host_canvas[rows[0]:rows[1]])

    def render(self, frame=0, dscene=None):
        """Rasterize this rank's strip and gather.  Returns the canvas tensor on the presenting rank."""
        r, torch = self.r, self.torch
        rows = self.my_rows
        active = rows[1] > rows[0]
        early = None
        if self.rank == self.presenting:
            r.frame_target(self.canvas.data_ptr(), 0)
            if self.world > 1 and self.mode == "nccl" and self.early_receives:
                # receives first: a strip then flows in as soon as its rank is done with it, while this rank
                # (and the slower ones) are still rasterizing
                early = self.post_receives(self.dist, self.rank, self.rows, self.canvas)
        elif self.mode == "p2p":
            r.frame_target(self._peer_canvas, 0)
        else:
            r.frame_target(self.strip.data_ptr(), rows[0])
        if active:
            if dscene is not None:
                r.frame_begin_device(dscene, frame)
            else:
                r.frame_begin(self.scene, frame)
            if self.world > 1:
                r.frame_strip(*rows)
            if dscene is not None:
                r.raster_entries_device(dscene.entries, dscene.n_entries)
            else:
                r.raster_entries(self.entries)
            _, self.last_stats = r.frame_end(want_image=False)
        if self.world > 1:
            if early is not None:
                for req in early:
                    req.wait()
            else:
                self._gather()
        return self.canvas

    @staticmethod
    def post_receives(dist, rank, rows, canvas):
        """One irecv per remote strip, straight into the canvas rows; returns the requests."""
        ops = [dist.P2POp(dist.irecv, canvas[y0:y1], k) for k, (y0, y1) in enumerate(rows) if k != rank and y1 > y0]
        return dist.batch_isend_irecv(ops) if ops else []

    def _gather(self):
        dist, torch = self.dist, self.torch
        if self.mode == "p2p":
            torch.cuda.synchronize()
            dist.barrier()
            return
        n = self.my_rows[1] - self.my_rows[0]
        self.gather_strips(dist, self.rank, self.presenting, self.rows,
                           self.strip[:n] if self.strip is not None else None, self.canvas)

    @staticmethod
    def gather_strips(dist, rank, presenting, rows, strip, canvas):
        """Finished strips to the presenting rank: it posts one irecv per remote strip straight into
        the canvas rows, every other rank one isend (grouped; NCCL over NVLink, or gloo in tests)."""
        ops = []
        if rank == presenting:
            for k, (y0, y1) in enumerate(rows):
                if k != rank and y1 > y0:
                    ops.append(dist.P2POp(dist.irecv, canvas[y0:y1], k))
        elif rows[rank][1] > rows[rank][0]:
            ops.append(dist.P2POp(dist.isend, strip, presenting))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
