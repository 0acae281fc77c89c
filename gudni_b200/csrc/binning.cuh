// binning.cuh — GPU tile binning (level 2 of the C ABI).
#pragma once
#include "context.cuh"

namespace gudni_bin {
// Bins `n` entries (device pointer, scene order) into the context's tile / shape arrays, in
// tile-tree traversal order, and sets ctx->nTiles / nShapes / nColumns.
int binScene(gudni_ctx* ctx, const gudni_shape_entry* devEntries, int n);
}  // namespace gudni_bin
