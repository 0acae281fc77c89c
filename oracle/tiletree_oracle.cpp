// tiletree_oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT (see kernels_oracle.hpp).
//
// CPU restatement of the reference's tile binning and job packing, one shape at a time exactly as
// the Haskell does it (persistent tree replaced by an in-place tree; list order preserved):
//   buildTileTree / addShapeToTree / hSplit / vSplit / traverseTileTree
//       /root/reference/src/Graphics/Gudni/Raster/TileTree.hs:74-204
//   accumulateRasterJobs / addTileToRasterJob   Raster/Job.hs:121-178
//   buildRasterJobs (argument swap and job-list order)   OpenCL/CallKernels.hs:244-255
// How this file is pinned: the reference holds no fixtures for the tile tree and its Haskell cannot be compiled in
// this image, so tests/golden/tiletree_handworked.py derives six inputs by hand from TileTree.hs / Job.hs /
// CallKernels.hs, rule by rule with line citations (strict comparisons at a cut, the 127th shape splitting across then
// along, the strand cap, the 8-pixel floor keeping 130 shapes, a split child that splits again while being refilled,
// job packing with the swapped arguments); tests/test_tiletree_handworked.py holds this file to those literals and
// tests/test_gpu_tiletree_handworked.py the GPU binning.  Beyond them: invariants (every pixel in exactly one leaf, caps
// respected unless at the floor, shape order) and, end to end, images against oracle/_ref.
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#include "../include/gudni_b200.h"

namespace {

const int MIN_TILE_SIZE = 8;   // mINtILEsIZE, Raster/Constants.hs:65
const int MAXSHAPE = 127;      // mAXsHAPE, Raster/Constants.hs:55-58

struct Tile {  // Tile TileEntry, Raster/Types.hs:100-118 + TileTree.hs:42-49
    int left, top, right, bottom;
    int hDepth, vDepth;
    std::vector<int> shapes;      // entry indices, OLDEST first (the Haskell list is newest first)
    uint32_t strandCount = 0;
    int shapeCount = 0;
};

struct Node {  // HTree / VTree, TileTree.hs:55-70.  isV: cut is horizontal line (VTree/VLeaf)
    bool isV = true;
    bool isLeaf = true;
    float cut = 0.f;
    std::unique_ptr<Node> first, second;  // VTree: top, bottom.  HTree: left, right
    Tile tile;
};

int adjustedLog(int x) {  // TileTree.hs:74-75: ceiling (logBase 2 x), 0 for x < 1
    if (x < 1) return 0;
    int d = 0;
    while ((1 << d) < x) d++;
    return d;
}

Tile emptyTile(int hDepth, int vDepth, int l, int t, int r, int b) {
    Tile tl;
    tl.left = l; tl.top = t; tl.right = r; tl.bottom = b;
    tl.hDepth = hDepth; tl.vDepth = vDepth;
    return tl;
}

struct Builder {
    const gudni_shape_entry* entries;
    uint32_t maxStrandsPerTile;

    // buildTileTree, TileTree.hs:81-109
    std::unique_ptr<Node> goV(int depth, int tileDepth, int l, int t, int r, int b) {
        auto n = std::make_unique<Node>();
        n->isV = true;
        if (depth > tileDepth) {
            int cut = t + (1 << (depth - 1));
            n->isLeaf = false;
            n->cut = (float)cut;
            n->first = goH(depth, tileDepth, l, t, r, cut);
            n->second = goH(depth, tileDepth, l, cut, r, b);
        } else {
            n->tile = emptyTile(depth, depth, l, t, r, b);
        }
        return n;
    }
    std::unique_ptr<Node> goH(int depth, int tileDepth, int l, int t, int r, int b) {
        auto n = std::make_unique<Node>();
        n->isV = false;
        if (depth > tileDepth) {
            int cut = l + (1 << (depth - 1));
            n->isLeaf = false;
            n->cut = (float)cut;
            n->first = goV(depth - 1, tileDepth, l, t, cut, b);
            n->second = goV(depth - 1, tileDepth, cut, t, r, b);
        } else {
            n->tile = emptyTile(depth, depth, l, t, r, b);
        }
        return n;
    }

    // checkTileSpace, TileTree.hs:161-166
    bool checkTileSpace(const Tile& tile, int e) const {
        uint32_t withAdded = tile.strandCount + entries[e].num_strands;
        return tile.shapeCount < MAXSHAPE - 1 && withAdded < maxStrandsPerTile;
    }
    void insertShapeTile(Tile& tile, int e) {  // TileTree.hs:147-158
        tile.shapes.push_back(e);
        tile.strandCount += entries[e].num_strands;
        tile.shapeCount += 1;
    }
    // hSplit / vSplit, TileTree.hs:169-190: children inherit the box halves, one depth less in the
    // split axis; the old shapes are re-inserted oldest first.
    void hSplit(Node* n) {
        Tile old = std::move(n->tile);
        int cut = old.left + ((old.right - old.left) / 2);
        n->isLeaf = false;
        n->cut = (float)cut;
        n->first = std::make_unique<Node>();
        n->second = std::make_unique<Node>();
        n->first->isV = n->second->isV = true;
        n->first->tile = emptyTile(old.hDepth - 1, old.vDepth, old.left, old.top, cut, old.bottom);
        n->second->tile = emptyTile(old.hDepth - 1, old.vDepth, cut, old.top, old.right, old.bottom);
        for (int e : old.shapes) insertShapeH(n, e);
    }
    void vSplit(Node* n) {
        Tile old = std::move(n->tile);
        int cut = old.top + ((old.bottom - old.top) / 2);
        n->isLeaf = false;
        n->cut = (float)cut;
        n->first = std::make_unique<Node>();
        n->second = std::make_unique<Node>();
        n->first->isV = n->second->isV = false;
        n->first->tile = emptyTile(old.hDepth, old.vDepth - 1, old.left, old.top, old.right, cut);
        n->second->tile = emptyTile(old.hDepth, old.vDepth - 1, old.left, cut, old.right, old.bottom);
        for (int e : old.shapes) insertShapeV(n, e);
    }
    // insertShapeH / insertShapeV, TileTree.hs:117-145
    void insertShapeH(Node* n, int e) {
        if (!n->isLeaf) {
            if (entries[e].left < n->cut) insertShapeV(n->first.get(), e);
            if (entries[e].right > n->cut) insertShapeV(n->second.get(), e);
        } else if (checkTileSpace(n->tile, e) || (n->tile.right - n->tile.left) <= MIN_TILE_SIZE) {
            insertShapeTile(n->tile, e);
        } else {
            hSplit(n);
            insertShapeH(n, e);
        }
    }
    void insertShapeV(Node* n, int e) {
        if (!n->isLeaf) {
            if (entries[e].top < n->cut) insertShapeH(n->first.get(), e);
            if (entries[e].bottom > n->cut) insertShapeH(n->second.get(), e);
        } else if (checkTileSpace(n->tile, e) || (n->tile.bottom - n->tile.top) <= MIN_TILE_SIZE) {
            insertShapeTile(n->tile, e);
        } else {
            vSplit(n);
            insertShapeV(n, e);
        }
    }
    // traverseTileTree, TileTree.hs:193-204
    template <class F>
    void traverse(Node* n, F&& f) {
        if (n->isLeaf) { f(n->tile); return; }
        traverse(n->first.get(), f);
        traverse(n->second.get(), f);
    }
};

struct Job {
    std::vector<gudni_shape> shapes;
    std::vector<gudni_tile> tiles;
    int columnAllocation = 0;
};

struct Jobs {
    std::vector<Job> jobs;  // in the order the reference submits them (see below)
};

}  // namespace

extern "C" {

// Bins `entries` (scene order, already culled) and packs the leaves into RasterJobs.
// Returned handle lists jobs in the order `queueRasterJobs` receives them, which is
// `bsCurrentJob : bsJobs` = LAST-created job first (OpenCL/CallKernels.hs:255, Raster/Job.hs:166).
void* gudni_oracle_build_jobs(const gudni_shape_entry* entries, int n_entries, int canvas_w, int canvas_h,
                              const gudni_spec* spec) {
    Builder b{entries, (uint32_t)spec->max_strands_per_tile};
    int canvasDepth = adjustedLog(canvas_w > canvas_h ? canvas_w : canvas_h);
    int tileDepth = adjustedLog(spec->max_tile_size);
    int side = 1 << canvasDepth;
    std::unique_ptr<Node> root = b.goV(canvasDepth, tileDepth, 0, 0, side, side);
    for (int e = 0; e < n_entries; e++) b.insertShapeV(root.get(), e);  // Serialize.hs:177 via onShape

    // buildRasterJobs passes (threadsPerTile, tilesPerCall) to accumulateRasterJobs, whose
    // parameters are (maxTilesPerJob, threadsPerTile): the two are swapped (CallKernels.hs:254 vs
    // Job.hs:151-156).  Restated as is.
    const int maxTilesPerJob = spec->threads_per_tile;
    const int columnsPerTile = spec->max_tiles_per_call;
    std::vector<Job> created(1);
    int tileCount = 0;
    b.traverse(root.get(), [&](const Tile& tile) {
        if (tileCount >= maxTilesPerJob) {
            created.emplace_back();
            tileCount = 0;
        }
        Job& job = created.back();
        gudni_tile ti{};
        ti.left = tile.left; ti.top = tile.top; ti.right = tile.right; ti.bottom = tile.bottom;
        ti.h_depth = (int16_t)tile.hDepth;
        ti.v_depth = (int16_t)tile.vDepth;
        ti.column_allocation = job.columnAllocation;
        ti.shape_start = (uint32_t)job.shapes.size();
        ti.shape_count = (uint32_t)tile.shapes.size();
        for (size_t i = tile.shapes.size(); i-- > 0;) {  // tileShapes is newest first (TileTree.hs:151)
            const gudni_shape_entry& se = entries[tile.shapes[i]];
            job.shapes.push_back(gudni_shape{se.tag, se.geo_start, se.num_strands});
        }
        job.tiles.push_back(ti);
        tileCount += 1;
        job.columnAllocation += columnsPerTile;
    });
    Jobs* out = new Jobs();
    out->jobs.assign(std::make_move_iterator(created.rbegin()), std::make_move_iterator(created.rend()));
    return out;
}

int gudni_oracle_jobs_count(void* h) { return (int)static_cast<Jobs*>(h)->jobs.size(); }
void gudni_oracle_job_info(void* h, int job, int* n_shapes, int* n_tiles, int* columns) {
    const Job& j = static_cast<Jobs*>(h)->jobs[job];
    *n_shapes = (int)j.shapes.size();
    *n_tiles = (int)j.tiles.size();
    *columns = j.columnAllocation;
}
const gudni_shape* gudni_oracle_job_shapes(void* h, int job) { return static_cast<Jobs*>(h)->jobs[job].shapes.data(); }
const gudni_tile* gudni_oracle_job_tiles(void* h, int job) { return static_cast<Jobs*>(h)->jobs[job].tiles.data(); }
void gudni_oracle_jobs_free(void* h) { delete static_cast<Jobs*>(h); }

}  // extern "C"
