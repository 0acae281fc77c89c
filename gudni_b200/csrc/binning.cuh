// binning.cuh — GPU tile binning (level 2 of the C ABI).
#pragma once
#include "context.cuh"

namespace gudni_bin {
// Bins `n` entries (device pointer, scene order) into the context's tile / shape arrays, in
// tile-tree traversal order, and sets ctx->nTiles / nShapes / nColumns.  `whileBinning` (may be null) is called once
// the first kernels are queued and before the host waits for their counts.
int binScene(gudni_ctx* ctx, const gudni_shape_entry* devEntries, int n, int (*whileBinning)(gudni_ctx*) = nullptr);
}  // namespace gudni_bin
