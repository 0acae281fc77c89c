// Host harness around the reference's own kernel source (see build_ref.py, cl_compat.hpp).
// TEST INFRASTRUCTURE, NOT PRODUCT.
//
// Plays the part of generateCall (OpenCL/CallKernels.hs:88-179): per raster job it allocates the four
// scratch buffers with the reference's sizes (:124-127), runs generateThresholds, sortThresholds and
// renderThresholds over the NDRange `Work2D numTiles threadsPerTile` (:141-142,153-154,173-174; global
// id 0 = tile, id 1 = column) one after the other, and releases the scratch.  Work-items are
// independent (no __local memory, the one barrier in the source is commented out, Kernels.cl:1991), so
// the NDRange is an OpenMP loop.  The entry point has the signature of the restated oracle's
// gudni_oracle_raster_job (oracle/kernels_oracle.cpp) so that one Python driver serves both.
#include <omp.h>

#include <cstdlib>
#include <vector>

#include "cl_compat.hpp"
#include "gudni_b200.h"

thread_local int cl_global_id[3] = {0, 0, 0};
static int cl_max_thresholds = 1024;  // MAXTHRESHOLDS (OpenCL/Setup.hs:48), set per call from the spec

namespace refcl {
#define new new_   // the source names parameters `new` (Kernels.cl:586-602,1098-1123): legal C, a keyword in C++
#include "kernels_cl.inc"
#undef new
}

static_assert(sizeof(refcl::Shape) == sizeof(gudni_shape), "Shape layout");
static_assert(sizeof(refcl::TileInfo) == sizeof(gudni_tile), "TileInfo layout");
static_assert(sizeof(refcl::PictureUse) == sizeof(gudni_picture_use), "PictureUse layout");
static_assert(sizeof(refcl::ShapeState) == 1088, "sIZEoFsHAPEsTATE, Raster/Constants.hs:60-63");
static_assert(sizeof(refcl::Slice) == 8, "Slice layout");

extern "C" {

int gudni_ref_threads(void) { return omp_get_max_threads(); }
void gudni_ref_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }

// Same contract as gudni_oracle_raster_job.  Returns the number of threads whose queue reached
// max_thresholds after generateThresholds (the reference has no overflow check — SURVEY.md App. B #9 —
// so their results, and possibly their neighbours', are garbage; the scratch has slack in front so
// that the stray writes stay inside the allocation).
int64_t gudni_ref_raster_job(const void* geometry, const float* substances, const uint8_t* picture_bytes,
                             const gudni_picture_use* picture_uses, const float* background_rgba, int width,
                             int height, const gudni_spec* spec, const gudni_shape* shapes, const gudni_tile* tiles,
                             int n_tiles, int columns_allocated, uint32_t* out, int32_t* n_thresholds,
                             int32_t* shape_bits, int64_t* total_thresholds) {
    using namespace refcl;
    const int threadsPerTile = spec->threads_per_tile;
    const int maxT = spec->max_thresholds;
    cl_max_thresholds = maxT;
    int computeDepth = 0;
    while ((1 << computeDepth) < threadsPerTile) computeDepth++;  // adjustedLog, Raster/TileTree.hs:74-75
    const size_t slack = 4 * (size_t)maxT;
    const size_t entries = (size_t)columns_allocated * maxT;      // CallKernels.hs:124-125
    // uninitialised, like the reference's clCreateBuffer (a zero fill of 1.3 GB per job is not its cost)
    float4* thresholdStore = static_cast<float4*>(std::malloc((entries + slack + 1) * sizeof(float4)));
    uint* headerStore = static_cast<uint*>(std::malloc((entries + slack + 1) * sizeof(uint)));
    float4* thresholdHeap = thresholdStore + slack;
    uint* headerHeap = headerStore + slack;
    ShapeState* shapeStateHeap = static_cast<ShapeState*>(std::calloc((size_t)columns_allocated + 1, sizeof(ShapeState)));
    Slice* qSliceHeap = static_cast<Slice*>(std::calloc((size_t)columns_allocated + 1, sizeof(Slice)));
    std::vector<float> randomField(4096, 0.0f);                   // geoRandomField; inert (STOCHASTIC_FACTOR 0)

    float4* geometryHeap = (float4*)geometry;
    Shape* shapeHeap = (Shape*)shapes;
    TileInfo* tileHeap = (TileInfo*)tiles;
    Substance* substanceHeap = (Substance*)substances;
    uchar* pictureData = (uchar*)picture_bytes;
    PictureUse* pictureRefs = (PictureUse*)picture_uses;
    const int2 bitmapSize = mk_int2(width, height);
    const float4 background = mk_float4(background_rgba[0], background_rgba[1], background_rgba[2], background_rgba[3]);
    const int frameNumber = 0, jobIndex = 0;
    const long total = (long)n_tiles * threadsPerTile;

#pragma omp parallel for schedule(dynamic, 64)
    for (long g = 0; g < total; g++) {
        cl_global_id[0] = (int)(g / threadsPerTile);
        cl_global_id[1] = (int)(g % threadsPerTile);
        generateThresholds(geometryHeap, shapeHeap, tileHeap, bitmapSize, computeDepth, frameNumber, jobIndex,
                           thresholdHeap, headerHeap, shapeStateHeap, qSliceHeap);
    }
    if (n_thresholds) for (int i = 0; i < columns_allocated; i++) n_thresholds[i] = -1;
    if (shape_bits) for (int i = 0; i < columns_allocated; i++) shape_bits[i] = -1;
    int64_t sum = 0, overflowed = 0;
    for (long g = 0; g < total; g++) {
        TileState tileS;
        initTileState(&tileS, getTileInfo(tileHeap, (int)(g / threadsPerTile)), bitmapSize, (int)(g % threadsPerTile),
                      jobIndex, computeDepth);
        if (!isActiveThread(&tileS)) continue;
        const int len = qSliceHeap[tileS.threadId].sLength;
        sum += len;
        overflowed += len >= maxT;
        if (n_thresholds) n_thresholds[tileS.threadId] = len;
        if (shape_bits) shape_bits[tileS.threadId] = (int32_t)shapeStateHeap[tileS.threadId].shapeBits;
    }
    if (total_thresholds) *total_thresholds = sum;

#pragma omp parallel for schedule(dynamic, 64)
    for (long g = 0; g < total; g++) {
        cl_global_id[0] = (int)(g / threadsPerTile);
        cl_global_id[1] = (int)(g % threadsPerTile);
        sortThresholds(thresholdHeap, headerHeap, qSliceHeap, tileHeap, bitmapSize, computeDepth, frameNumber, jobIndex);
    }
#pragma omp parallel for schedule(dynamic, 64)
    for (long g = 0; g < total; g++) {
        cl_global_id[0] = (int)(g / threadsPerTile);
        cl_global_id[1] = (int)(g % threadsPerTile);
        renderThresholds(thresholdHeap, headerHeap, shapeStateHeap, qSliceHeap, substanceHeap, pictureData, pictureRefs,
                         randomField.data(), shapeHeap, tileHeap, background, bitmapSize, computeDepth, frameNumber,
                         jobIndex, out);
    }
    std::free(thresholdStore);
    std::free(headerStore);
    std::free(shapeStateHeap);
    std::free(qSliceHeap);
    return overflowed;
}

}  // extern "C"
