/* gudni_b200.h — C ABI of libgudni_b200.so, the sm_100a rasterizer that stands in for
 * Gudni's OpenCL host layer + Kernels.cl.
 *
 * Every entry point below replaces a Haskell function of the reference whose body calls CLUtil /
 * OpenCL today (the reference has no FFI for this path).  Paths are relative
 * to /root/reference/src/Graphics/Gudni/.
 *
 * Conventions
 *   - plain C, no torch / CUDA types in any signature; pointers named `dev_*` are device pointers,
 *     all others are host pointers owned by the caller and only read during the call.
 *   - every function returns GUDNI_OK (0) or a negative gudni_status; the message for the last
 *     failure on a context is available from gudni_b200_last_error().  Nothing aborts or throws.
 *   - one caller thread per context; calls may block (safe for `foreign import ccall safe`).
 *   - there is no CPU fallback: if no sm_100-class device is usable, init fails.
 */
#ifndef GUDNI_B200_H
#define GUDNI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------
 * Wire formats (little endian).  These are the byte layouts the Haskell `StorableM` instances
 * poke and `Kernels.cl` reads; they are the ABI contract.
 * ---------------------------------------------------------------------------------------------- */

/* RasterSpec — OpenCL/Rasterizer.hs:37-50, derived in OpenCL/Setup.hs:71-87. */
typedef struct gudni_spec {
    int32_t max_tile_size;        /* specMaxTileSize      (pixels, power of two)            */
    int32_t threads_per_tile;     /* specThreadsPerTile   (power of two, 32..1024)          */
    int32_t max_tiles_per_call;   /* specMaxTilesPerCall  (job packing only)                */
    int32_t max_thresholds;       /* specMaxThresholds    (MAXTHRESHOLDS, per column-thread)*/
    int32_t max_strands_per_tile; /* specMaxStrandsPerTile                                  */
    int32_t max_shapes;           /* specMaxShapes        (MAXSHAPE, <= 127)                */
} gudni_spec;

/* Shape GeoReference — Raster/Types.hs:159-168 / Kernels.cl:318-321.  16 bytes. */
typedef struct gudni_shape {
    uint64_t tag;          /* ShapeTag: Raster/Constants.hs:74-89                            */
    uint32_t geo_start;    /* offset into the geometry heap in 16-byte units                 */
    uint32_t num_strands;
} gudni_shape;

#define GUDNI_TAG_SUBSTANCETYPE_MASK  0xC000000000000000ull
#define GUDNI_TAG_SUBSTANCE_SOLID     0x8000000000000000ull
#define GUDNI_TAG_SUBSTANCE_PICTURE   0x4000000000000000ull
#define GUDNI_TAG_COMPOUND_MASK       0x3000000000000000ull
#define GUDNI_TAG_COMPOUND_CONTINUE   0x1000000000000000ull
#define GUDNI_TAG_COMPOUND_ADD        0x2000000000000000ull
#define GUDNI_TAG_COMPOUND_SUBTRACT   0x3000000000000000ull
#define GUDNI_TAG_SUBSTANCEID_MASK    0x0FFFFFFFFFFFFFFFull

/* Tile (Slice (Shape GeoReference), Int) — Raster/Types.hs:176-198 / Kernels.cl:335-341. 32 B. */
typedef struct gudni_tile {
    int32_t  left, top, right, bottom; /* tileBox in pixels                                   */
    int16_t  h_depth;                  /* log2 width                                          */
    int16_t  v_depth;                  /* log2 height                                         */
    int32_t  column_allocation;        /* first column-thread id of the tile inside its job   */
    uint32_t shape_start;              /* slice into the job's shape array                    */
    uint32_t shape_count;
} gudni_tile;

/* PictureUsage PictureMemoryReference — Figure/Picture.hs:183-199 / Kernels.cl:349-354. 24 B. */
typedef struct gudni_picture_use {
    float    translate_x, translate_y;
    int32_t  width, height;
    uint32_t mem_offset;               /* byte offset into the picture heap                   */
    float    scale;
} gudni_picture_use;

/* Un-binned shape entry = Shape ShapeEntry (Raster/Types.hs:139-157): what `addShapeToTree`
 * (Raster/TileTree.hs:113) receives.  Not a reference wire format (Haskell never serialises it);
 * this is the level-2 input that moves tile binning behind the shim.  32 bytes. */
typedef struct gudni_shape_entry {
    uint64_t tag;
    uint32_t geo_start;
    uint32_t num_strands;
    float    left, top, right, bottom; /* shapeBox (includes control points)                  */
} gudni_shape_entry;

/* ---- level 3 inputs: raw outlines + transformer stacks (what the scene holds BEFORE serialisation) ----
 * CurvePair — Figure/Outline.hs:55-57: an on-curve point and the control point towards the next one. */
typedef struct gudni_curve_pair {
    float on_x, on_y, off_x, off_y;
} gudni_curve_pair;
/* Outline — Figure/Outline.hs:59-61: a closed loop of curve pairs, `n_pairs` of them from `first_pair`.
 * Outlines may be shared by any number of shapes (one unit circle, 100,000 placements). */
typedef struct gudni_outline {
    uint32_t first_pair, n_pairs;
} gudni_outline;
/* One simple transformation — Figure/Transformer.hs:94-105.  A rotation carries cos and sin of its angle
 * (the caller's libm computes them, as the Haskell side would), so the library only multiplies. */
enum { GUDNI_TRANSFORM_TRANSLATE = 0, GUDNI_TRANSFORM_SCALE = 1, GUDNI_TRANSFORM_ROTATE = 2 };
typedef struct gudni_transform {
    uint32_t kind;     /* GUDNI_TRANSFORM_*                                             */
    float a, b;        /* translate: (dx, dy); scale: (factor, -); rotate: (cos, sin)   */
    uint32_t reserved;
} gudni_transform;
/* A shape as the scene tree holds it (Raster/TraverseShapeTree.hs:35-80 hands onShape exactly this):
 * tag, its outlines, and the transformer stack above it, listed outermost first — `tTranslate p .
 * tScale s $ shape` is {translate p, scale s} — and applied last to first (CombineTransform, :105). */
typedef struct gudni_outline_shape {
    uint64_t tag;                          /* as gudni_shape_entry.tag */
    uint32_t first_outline, n_outlines;    /* into outlines[]          */
    uint32_t first_transform, n_transforms;/* into transforms[]        */
    uint32_t reserved[2];
} gudni_outline_shape;                     /* 32 bytes */

/* Per-frame statistics — replaces the reference's putStrLn/`tr` logging. */
typedef struct gudni_stats {
    int64_t n_tiles;            /* leaf tiles rendered                                        */
    int64_t n_shape_refs;       /* sum over tiles of shape_count                              */
    int64_t n_thresholds;       /* thresholds kept by the generate phase, summed over threads */
    int64_t n_spilled_threads;  /* column-threads that left the on-chip queue for the HBM one */
    int64_t n_overflow_threads; /* column-threads that exceeded max_thresholds (UB in ref.)   */
    int64_t algorithmic_bytes;  /* A(frame), SURVEY.md §8(d)                                  */
    float   ms_upload, ms_bin, ms_raster, ms_download; /* CUDA-event stage times              */
    float   ms_strands;          /* level 3 only: outlines -> geometry heap + shape entries */
    float   reserved;
} gudni_stats;

typedef enum gudni_status {
    GUDNI_OK            =  0,
    GUDNI_ERR_ARGUMENT  = -1,
    GUDNI_ERR_NO_DEVICE = -2,
    GUDNI_ERR_CUDA      = -3,
    GUDNI_ERR_STATE     = -4,
    GUDNI_ERR_OOM       = -5
} gudni_status;

typedef struct gudni_ctx gudni_ctx;

/* ------------------------------------------------------------------------------------------------
 * Entry points
 * ---------------------------------------------------------------------------------------------- */

/* Replaces setupOpenCL (OpenCL/Setup.hs:102-147) + determineRasterSpec (:71-87).
 * `device` = CUDA ordinal or -1 for the current device.  `want` may be NULL (canonical spec:
 * 256,256,256,1024,1022,127).  `got` receives the spec in force; the Haskell side reads
 * max_tile_size / threads_per_tile / max_tiles_per_call / max_strands_per_tile from it
 * (Application.hs:224, Raster/Serialize.hs:131, OpenCL/CallKernels.hs:251-252). */
int gudni_b200_init(int device, const gudni_spec* want, gudni_spec* got, gudni_ctx** out);

/* Replaces the frame-constant uploads of queueRasterJobs (OpenCL/CallKernels.hs:223-242:
 * pileToBuffer x4 + vectorToBuffer) and picks up bitmapSize / frameCount / background of
 * `raster` (:189-197) and generateCall (:159-171).  The random field is not taken: it is inert
 * (STOCHASTIC_FACTOR = 0, Raster/Constants.hs:54). */
int gudni_b200_frame_begin(gudni_ctx* ctx,
                           const void* geometry, size_t geometry_bytes,
                           const float* substances /* 4 floats each */, int n_substances,
                           const uint8_t* picture_bytes, size_t n_picture_bytes,
                           const gudni_picture_use* picture_uses, int n_picture_uses,
                           const float background_rgba[4],
                           int width, int height, int frame_number);

/* Persistent input cache.  The reference packs and uploads pictures, geometry and
 * substances again every frame (OpenCL/CallKernels.hs:229-235), changed or not.  The caller knows when a Pile
 * changed; it says so with a generation counter per input: a buffer is uploaded unless its generation is nonzero
 * and equal to the generation, pointer-independent, under which the same number of bytes was last uploaded into
 * this context.  Counters, not content hashes: hashing 30 MB on the host costs more than sending it, and a
 * pointer comparison would be wrong for Piles that are refilled in place — the caller bumps the counter when it
 * refills.  generation 0 = always upload (what gudni_b200_frame_begin does).  `entries` covers the shape-entry
 * array of gudni_b200_raster_scene_cached. */
typedef struct gudni_generations {
    uint64_t geometry, substances, pictures, picture_uses, entries;
} gudni_generations;
int gudni_b200_frame_begin_cached(gudni_ctx* ctx,
                                  const void* geometry, size_t geometry_bytes,
                                  const float* substances, int n_substances,
                                  const uint8_t* picture_bytes, size_t n_picture_bytes,
                                  const gudni_picture_use* picture_uses, int n_picture_uses,
                                  const float background_rgba[4],
                                  int width, int height, int frame_number,
                                  const gudni_generations* generations);
int gudni_b200_raster_scene_cached(gudni_ctx* ctx, const gudni_shape_entry* entries, int n_entries, uint64_t generation);

/* Restrict this context to canvas rows [row_begin, row_end) (whole rows of root tiles).  Used by
 * the multi-GPU strip partition (no reference counterpart: one OpenCLState = one device,
 * OpenCL/Setup.hs:118-120).  row_begin = 0, row_end = height restores the full frame. Must be
 * called after frame_begin and before any raster call of the frame. */
int gudni_b200_frame_strip(gudni_ctx* ctx, int row_begin, int row_end);

/* Level 1 — exact stand-in for `raster`/`generateCall` (OpenCL/CallKernels.hs:182-206, 88-179):
 * one RasterJob whose tiles were binned by the caller (Raster/Job.hs:132-178).  Asynchronous:
 * returns once the job is enqueued. */
int gudni_b200_raster_job(gudni_ctx* ctx,
                          const gudni_shape* shapes, int n_shapes,
                          const gudni_tile* tiles, int n_tiles,
                          int columns_allocated, int job_index);

/* Level 2 — tile binning behind the shim: replaces buildTileTree/addShapeToTree
 * (Raster/TileTree.hs:81-190), traverseTileTree (:193-204), accumulateRasterJobs
 * (Raster/Job.hs:151-178) and the per-job loop of queueRasterJobs.  Entries are in scene order
 * (first = top-most), already culled against the canvas (Raster/Serialize.hs:97-104). */
int gudni_b200_raster_scene(gudni_ctx* ctx, const gudni_shape_entry* entries, int n_entries);

/* Replaces the OutputPtr read-back (OpenCL/Instances.hs:60-75): waits for the frame, copies the
 * BGRA8 words (B | G<<8 | R<<16 | 0xFF<<24) of the context's rows into `out_bgra`
 * (width * (row_end-row_begin) words, may be NULL to leave the frame on the device) and fills
 * `stats` (may be NULL).
 * Returns GUDNI_ERR_ARGUMENT — and a bitmap that means nothing — if the frame's geometry held a point at
 * +-infinity: the curve bisection of Kernels.cl:1226-1258 does not terminate on one (the reference's kernels
 * hang), so the raster kernels were not let near it.  NaN and large finite coordinates render as in the
 * reference.  The context stays usable. */
int gudni_b200_frame_end(gudni_ctx* ctx, uint32_t* out_bgra, gudni_stats* stats);

/* ---- several GPUs of one box, one process ------------------------------------------------------
 * No reference counterpart: one OpenCLState is one device (OpenCL/Setup.hs:118-120) and queueRasterJobs
 * (OpenCL/CallKernels.hs:218-242) hands it every job.  gudni_b200_multi_frame is queueRasterJobs for a canvas that
 * is worth sharding (a 16K x 16K canvas; a frame that fits one GPU is better left on one): the canvas is cut into
 * strips of whole root-tile rows, one per device, re-cut every frame from the devices' measured times; each device
 * bins and rasterizes the shapes that touch its strip (its own context, its own host thread) and copies its rows
 * straight into `out_bgra` over its own PCIe link.  With a presenting device set (gudni_b200_multi_set_presenting)
 * the strips are also pushed over NVLink into one canvas on that device (gudni_b200_multi_canvas), for a presenter
 * that keeps the frame on the GPU.  `devices` may be NULL (0 .. n_devices-1) and may name a device twice (two
 * contexts on one GPU: how the path is tested on a one-GPU box). */
#define GUDNI_MULTI_MAX_DEVICES 16
typedef struct gudni_multi gudni_multi;
typedef struct gudni_multi_stats {
    int32_t n_devices;
    int32_t device[GUDNI_MULTI_MAX_DEVICES];
    int32_t row_begin[GUDNI_MULTI_MAX_DEVICES], row_end[GUDNI_MULTI_MAX_DEVICES];   /* this frame's strips */
    float   ms_device[GUDNI_MULTI_MAX_DEVICES];   /* strands + binning + raster kernels of the strip (CUDA events) */
    float   ms_frame;                             /* wall clock of the call                                         */
    float   ms_gather_exposed;                    /* last strip landed - last device done rasterizing               */
    gudni_stats total;                            /* counts summed over the devices, stage times the maximum        */
} gudni_multi_stats;
int gudni_b200_multi_init(int n_devices, const int* devices, const gudni_spec* want, gudni_spec* got, gudni_multi** out);
void gudni_b200_multi_destroy(gudni_multi* m);
const char* gudni_b200_multi_last_error(gudni_multi* m);
/* device_index: index into the devices given to multi_init, or -1 for no device-side canvas (the default) */
int gudni_b200_multi_set_presenting(gudni_multi* m, int device_index);
int gudni_b200_multi_canvas(gudni_multi* m, void** dev_bgra, int* device);
/* frame_begin_cached + raster_scene_cached + frame_end for the whole canvas.  out_bgra: width*height words, may be NULL. */
int gudni_b200_multi_frame(gudni_multi* m,
                           const void* geometry, size_t geometry_bytes,
                           const float* substances, int n_substances,
                           const uint8_t* picture_bytes, size_t n_picture_bytes,
                           const gudni_picture_use* picture_uses, int n_picture_uses,
                           const float background_rgba[4], int width, int height, int frame_number,
                           const gudni_shape_entry* entries, int n_entries,
                           const gudni_generations* generations,
                           uint32_t* out_bgra, gudni_multi_stats* stats);
/* The strip partition by itself (host arithmetic, no GPU needed): first cut from the shapes' boxes, and the feedback
 * cut from last frame's strips and per-device times.  rows: n_devices pairs (row_begin, row_end). */
int gudni_b200_partition_rows(const gudni_shape_entry* entries, int n_entries, int width, int height, int tile_rows, int n_devices,
                              int* rows_out);
int gudni_b200_rebalance_rows(const int* rows_in, const double* ms, int n_devices, int height, int tile_rows, int* rows_out);

/* Device-side access for callers that keep data on the GPU (bench, multi-GPU gather).  The frame
 * pointer stays valid until the next frame_begin with a different size, or destroy. */
int gudni_b200_frame_device_ptr(gudni_ctx* ctx, void** dev_bgra, size_t* n_bytes);
/* Redirect pixel stores of this context to a caller-owned device buffer whose first row is canvas
 * row `row_origin` (row y of the canvas at word (y - row_origin)*width).  Two uses: a torch-owned
 * strip tensor (row_origin = the strip's first row), or a peer GPU's full canvas mapped through
 * CUDA IPC (row_origin = 0), so strips land on the presenting GPU over NVLink without a separate
 * gather.  NULL restores the context's own frame buffer. */
int gudni_b200_frame_target(gudni_ctx* ctx, void* dev_bgra, int row_origin);
/* The same for a bitmap in HOST memory: `host_bgra` — width * height words, page-locked through
 * gudni_b200_host_register — becomes the frame's target, and the kernels store their finished rows into it
 * across PCIe while the rest of the frame is still being rasterized, the way strips cross NVLink into a
 * presenting GPU's canvas.  gudni_b200_frame_end called with the same pointer then has nothing left to copy
 * (replaces the read-back of the OutputPtr target, OpenCL/Instances.hs:60-75, CallKernels.hs:196-201).  Stays
 * in force until called with NULL; whole frames only (no gudni_b200_frame_strip). */
int gudni_b200_frame_target_host(gudni_ctx* ctx, uint32_t* host_bgra);
/* Run this context's work on a caller-owned CUDA stream (a cudaStream_t passed as void*), e.g.
 * torch's current stream so the caller's events bracket the kernels.  NULL restores the context's
 * own stream. */
int gudni_b200_set_stream(gudni_ctx* ctx, void* cuda_stream);
/* CUDA-IPC plumbing for the above (one process per GPU). `handle` is 64 bytes. */
int gudni_b200_ipc_export_frame(gudni_ctx* ctx, void* handle_64b);
int gudni_b200_ipc_open(gudni_ctx* ctx, const void* handle_64b, void** dev_ptr);
int gudni_b200_ipc_close(gudni_ctx* ctx, void* dev_ptr);

/* Same as frame_begin/raster_scene but with inputs already resident on the device
 * (pointers previously obtained from gudni_b200_device_alloc).  Used by bench.py's
 * inputs-in-HBM leg. */
int gudni_b200_device_alloc(gudni_ctx* ctx, size_t bytes, void** dev_ptr);
int gudni_b200_device_free(gudni_ctx* ctx, void* dev_ptr);
int gudni_b200_upload(gudni_ctx* ctx, void* dev_dst, const void* host_src, size_t bytes);
int gudni_b200_download(gudni_ctx* ctx, void* host_dst, const void* dev_src, size_t bytes);
int gudni_b200_frame_begin_device(gudni_ctx* ctx,
                                  const void* dev_geometry, size_t geometry_bytes,
                                  const void* dev_substances, int n_substances,
                                  const void* dev_picture_bytes, size_t n_picture_bytes,
                                  const void* dev_picture_uses, int n_picture_uses,
                                  const float background_rgba[4],
                                  int width, int height, int frame_number);
int gudni_b200_raster_scene_device(gudni_ctx* ctx, const void* dev_entries, int n_entries);
/* Level 3 — strand building behind the shim as well: replaces onShape's
 * geometry work (Raster/Serialize.hs:148-177: applyTransformer, boxOf, excludeBox, appendGeoRef),
 * enclose (Raster/Enclosure.hs:62-73), outlineToStrands (Raster/Strand.hs:153-178), replaceKnobs
 * (Raster/Deknob.hs:104-108) and the reorder table (Raster/ReorderTable.hs:96-110), then continues as
 * level 2.  Call frame_begin with no geometry (NULL, 0); shapes in scene order (first = top-most).
 * Shapes whose transformed bounding box misses the canvas are dropped, as excludeBox drops them. */
int gudni_b200_raster_outlines(gudni_ctx* ctx,
                               const gudni_outline_shape* shapes, int n_shapes,
                               const gudni_outline* outlines, int n_outlines,
                               const gudni_curve_pair* pairs, int64_t n_pairs,
                               const gudni_transform* transforms, int n_transforms);
/* Same with the four arrays already on the device (bench.py's inputs-in-HBM leg; indices are trusted). */
int gudni_b200_raster_outlines_device(gudni_ctx* ctx,
                                      const void* dev_shapes, int n_shapes,
                                      const void* dev_outlines, int n_outlines,
                                      const void* dev_pairs, int64_t n_pairs,
                                      const void* dev_transforms, int n_transforms);
/* Debug: the geometry heap and shape entries level 3 built for the last frame (either may be NULL to
 * query sizes). */
int gudni_b200_debug_strands(gudni_ctx* ctx, void* geometry, size_t geometry_capacity, size_t* geometry_bytes,
                             gudni_shape_entry* entries, int64_t entry_capacity, int64_t* n_entries);
/* Optional: page-lock a caller-owned host buffer that stays at the same address across frames (a
 * Haskell `Pile`'s storage, the SDL texture) so the copies in frame_begin / raster_scene / frame_end run
 * as direct DMA instead of going through the driver's staging buffer.  Must be unregistered before the
 * buffer is freed or reallocated.  (The reference got the same effect from CL_MEM_USE_HOST_PTR,
 * OpenCL/Instances.hs:39-43.) */
int gudni_b200_host_register(gudni_ctx* ctx, void* host_ptr, size_t bytes);
int gudni_b200_host_unregister(gudni_ctx* ctx, void* host_ptr);
/* Waits for all queued work of the context. */
int gudni_b200_sync(gudni_ctx* ctx);
/* Milliseconds of device time between the first and last kernel of the last frame. */
int gudni_b200_last_frame_ms(gudni_ctx* ctx, float* ms);
/* Number of kernels this library launched on the context since init. */
int gudni_b200_launch_count(gudni_ctx* ctx, int64_t* n);

/* Parity taps (tests only; no reference counterpart — the reference's DEBUG_OUTPUT printf,
 * Kernels.cl:54-68, is the closest).  When enabled before a raster call, the generate phase
 * records, per column-thread id (column_allocation + column, jobs laid end to end), the queue
 * length qSlice.sLength and ShapeState.shapeBits it would have stored (Kernels.cl:2078-2080). */
int gudni_b200_debug_enable(gudni_ctx* ctx, int on);
int gudni_b200_debug_thread_counts(gudni_ctx* ctx, int32_t* n_thresholds, int32_t* shape_bits,
                                   int64_t capacity, int64_t* n_threads);
/* Device self-test of the arithmetic helpers that are not a literal transcription of the reference:
 * the shared-reciprocal IEEE division used by `composite` is compared bit for bit with the
 * compiler's division on `n` pseudo-random operand triples; *mismatches must come back 0. */
int gudni_b200_debug_selftest(gudni_ctx* ctx, uint64_t n, uint64_t seed, uint64_t* mismatches);
/* Tiles and per-tile shape lists produced by the last level-2 binning, in job order. */
int gudni_b200_debug_binned(gudni_ctx* ctx, gudni_tile* tiles, int64_t tile_capacity,
                            int64_t* n_tiles, gudni_shape* shapes, int64_t shape_capacity,
                            int64_t* n_shapes);

const char* gudni_b200_last_error(gudni_ctx* ctx);
void gudni_b200_destroy(gudni_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* GUDNI_B200_H */
