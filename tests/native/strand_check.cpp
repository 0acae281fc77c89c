// strand_check.cpp — TEST HELPER: the per-shape logic of the level-3 kernels (gudni_b200/csrc/strand_build.cuh,
// the very text nvcc compiles for the device) compiled for the host, with the three launches of strands.cu
// replaced by loops.  tests/test_strands_host.py holds its output byte for byte to the harness's own
// restatement of Raster/Strand.hs (gudni_b200/csrc/host/strand.hpp), so the kernels' logic is checked on the
// CPU suite before a GPU sees it.  Not product, not oracle.
#include <cstring>
#include <vector>

#include "../../gudni_b200/csrc/strand_build.cuh"

using namespace gudni_strands;

extern "C" {

// Returns the number of kept shapes; *geometry_bytes gets the heap size.  Call with geometry == NULL to size.
int64_t strand_check_build(const gudni_outline_shape* shapes, int n_shapes, const gudni_outline* outlines,
                           const gudni_curve_pair* pairs, const gudni_transform* transforms, int width, int height,
                           uint8_t* geometry, size_t geometry_capacity, size_t* geometry_bytes, gudni_shape_entry* entries,
                           int64_t* n_strands) {
    static ReorderTable table;
    static bool built = false;
    if (!built) { buildReorderTable(table); built = true; }
    std::vector<ShapeMeasure> m(n_shapes);
    for (int i = 0; i < n_shapes; i++) m[i] = measureShape(shapes[i], outlines, pairs, transforms);
    uint64_t units = 0;
    int64_t kept = 0, strands = 0;
    std::vector<uint64_t> start(n_shapes);
    std::vector<int64_t> index(n_shapes, -1);
    for (int i = 0; i < n_shapes; i++) {
        if (culled(m[i], width, height)) continue;
        start[i] = units;
        index[i] = kept++;
        units += m[i].units;
        strands += m[i].strands;
    }
    if (geometry_bytes) *geometry_bytes = (size_t)units * 16;
    if (n_strands) *n_strands = strands;
    if (!geometry || !entries) return kept;
    if (geometry_capacity < units * 16) return -1;
    std::memset(geometry, 0xCD, (size_t)units * 16);   // every byte must be written by emitShape
    for (int i = 0; i < n_shapes; i++) {
        if (index[i] < 0) continue;
        gudni_shape_entry e{};
        e.tag = shapes[i].tag;
        e.geo_start = (uint32_t)start[i];
        e.num_strands = m[i].strands;
        e.left = m[i].left; e.top = m[i].top; e.right = m[i].right; e.bottom = m[i].bottom;
        entries[index[i]] = e;
        emitShape(shapes[i], outlines, pairs, transforms, &table, geometry, 16ull * start[i]);
    }
    return kept;
}

// row of the reorder table for a strand of n Béziers (2n+1 entries), for the known-answer test
void strand_check_table_row(int n, uint8_t* out) {
    ReorderTable t;
    buildReorderTable(t);
    std::memcpy(out, t.row[n], 2 * n + 1);
}

}  // extern "C"
