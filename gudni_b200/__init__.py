"""gudni_b200 — sm_100a rasterizer behind Gudni's Raster/OpenCL boundary.

Only what the hot path needs lives here:
  csrc/         CUDA kernels + the C-ABI shim (libgudni_b200.so), and csrc/host/ the harness that
                produces wire-format inputs the way the Haskell front end would (libgudni_host.so)
  formats.py    numpy views of the wire formats
  scene.py      scene construction for tests / bench (harness, not product)
  raster.py     host-side mirror of the reference's Rasterizer interface over the C ABI
  strips.py     multi-GPU strip partition + gather
"""
from .formats import (SHAPE_DTYPE, TILE_DTYPE, ENTRY_DTYPE, PICTURE_USE_DTYPE, RasterSpec,
                      CANONICAL_SPEC)

__all__ = ["SHAPE_DTYPE", "TILE_DTYPE", "ENTRY_DTYPE", "PICTURE_USE_DTYPE", "RasterSpec",
           "CANONICAL_SPEC"]
