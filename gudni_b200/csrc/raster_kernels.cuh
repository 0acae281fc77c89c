// raster_kernels.cuh — launch wrappers of the raster kernels.
#pragma once
#include "context.cuh"
#include "raster_device.cuh"

namespace gudni_launch {
int rasterTiles(gudni_ctx* ctx, const gudni_dev::FrameParams& P, int tileBase, int nTiles);
int rasterSpill(gudni_ctx* ctx, const gudni_dev::FrameParams& P);
int strandBounds(gudni_ctx* ctx, const void* geometry, const void* records, int stride, int count, float2* bounds);
int selftestDiv3(gudni_ctx* ctx, unsigned long long n, unsigned long long seed, unsigned long long* devMismatches);
}  // namespace gudni_launch
