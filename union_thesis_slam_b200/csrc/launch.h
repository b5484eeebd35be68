// launch.h -- host-side launch wrappers shared by the .cu translation units and capi.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct FrameParams;
struct VolumeParams;
struct CullParams;
struct BrickRef;
struct FramePixel;

// Per-frame scratch.  No memset per frame: k_prep_frame zeroes the counters (it is the first kernel of a frame's
// chain and does not use them) and the maxima are TAGGED -- (frame tag << 32) | float bits, raised with a 64-bit
// atomicMax -- so a value left by another frame simply reads as "no valid depth" (frame_max_depth()).  One launch less
// on the preparation stream per frame, which at 8 ranks is close to the critical path (profiles/r2_dense_ab.md).
struct FrameScratch {
    unsigned long long max_depth_tagged; // tag << 32 | float bits of the largest valid depth (rounded up)
    int vis_count;               // entries in the visible-brick list
    unsigned long long updated;  // voxels updated by this frame
    int cand_count;              // hash: candidate keys emitted by the ray walk (after dedup)
    int new_keys;                // hash: keys inserted this frame
    int error;                   // sticky per-frame error flag (tsdf_status), 0 = ok
    int vis_extra[3];            // dense: entries in the visible-brick lists of brick classes 1..3
    int pad;
};
// Dense grid sharding: brick-thick x-sheets dealt to the ranks in SNAKE order (0 1 .. W-1 W-1 .. 1 0 0 1 ..).
// A frustum's load per sheet changes smoothly along x; plain round-robin hands rank W-1 every sheet's
// "later" neighbour, which adds up to a 7-8 % per-frame imbalance at 8 ranks, while the snake pairs each
// rank's sheets so that a linear trend cancels.  sheet_of(): global sheet of rank's l-th local sheet;
// sheets_of(): how many sheets of nbx the rank owns.  dist.owned_x() is the Python mirror.
#ifdef __CUDACC__
#define TSDF_HD __host__ __device__
#else
#define TSDF_HD
#endif
static inline TSDF_HD int sheet_of(int l, int rank, int world)
{
    return 2 * world * (l >> 1) + ((l & 1) ? 2 * world - 1 - rank : rank);
}
static inline int sheets_of(int nbx, int rank, int world)
{
    const int period = 2 * world, full = nbx / period, rem = nbx % period;
    return 2 * full + (rank < rem ? 1 : 0) + (period - 1 - rank < rem ? 1 : 0);
}
#define DENSE_MAX_CLASSES 4
#define DENSE_MAX_BATCH 4 // frames applied to a brick in one visit (temporal blocking)

// Max valid depth per pixel tile, 4 levels: 8, 16, 32, 64 px tiles (0 = no valid pixel in the tile).
// Level 0 is plain floats, fully rewritten by prep_frame every frame; levels 1..3 are tagged 64-bit words raised by
// atomicMax from the same kernel (see FrameScratch) and read through pyramid_depth().
struct TilePyramid {
    float *level0;
    unsigned long long *up[3];
    int tx[4], ty[4];
};

// Per-warp update counters of the integrate kernels: every warp owns one slot of each array, so
// counting costs no atomics (thousands of same-address atomics per launch cost ~15-25 us).
// frame[] is overwritten every frame, total[] accumulates; the host sums them when asked.
struct UpdateCounters {
    unsigned long long *frame;
    unsigned long long *total;
};

// Everything the stats calls report, gathered by one small kernel and one D2H copy.
struct StatsBlock {
    unsigned long long updated_last, updated_total;
    unsigned long long last_visited, max_probe, error;
    int cand_count, new_keys, vis_count, n_entries, heap_top, pad;
};

// Lives for the lifetime of the handle.
struct Totals {
    unsigned long long updated;      // voxels updated since create
    unsigned long long last_visited; // bricks/blocks visited by the last frame
    unsigned long long max_probe;    // hash: longest probe sequence
    unsigned long long error;        // hash: sticky error (tsdf_status) raised inside a kernel
};

void launch_fill_bricks(float *bricks, size_t nbricks, int sm_count, cudaStream_t s);
void launch_prep_frame(const uint8_t *rgb, const double *depth, FramePixel *px, FrameScratch *fs,
                       const TilePyramid &pyr, int H, int W, unsigned tag, cudaStream_t s);
void launch_cull_bricks(const FrameParams &f, const CullParams &c, int nlbx, int nby, int nbz, int rank, int world,
                        int classes, BrickRef *lists, size_t class_cap, FrameScratch *fs, cudaStream_t s);
void launch_depth_u16(const uint16_t *raw, double *depth, int npix, double scale, int invalid_raw, cudaStream_t s);
void launch_collect_stats(UpdateCounters uc, int n_warps, const FrameScratch *fs, const Totals *tot, const int *n_entries,
                          const int *heap_top, StatsBlock *out, cudaStream_t s, bool with_total = true);
int integrate_dense_ctas_per_sm();
int integrate_reserve_ctas(); // CTA slots per SM left free for the concurrent preparation stream
void launch_integrate_dense(const FrameParams &f, const VolumeParams &v, const BrickRef *list, const int *count,
                            float *bricks, UpdateCounters uc, int grid, cudaStream_t s);
void launch_cull_bricks_batch(const FrameParams *f, const CullParams *c, FrameScratch *const *fs, int k, int nlbx, int nby,
                              int nbz, int rank, int world, int classes, BrickRef *lists, size_t class_cap, cudaStream_t s);
void launch_integrate_dense_batch(const FrameParams *f, int k, const VolumeParams &v, const BrickRef *list,
                                  const int *count, float *bricks, UpdateCounters uc, int grid, cudaStream_t s);
void launch_unbrick(const float *bricks, int n_x, int dim_y, int dim_z, int nby, int nbz, float *tsdf, float *wgt,
                    float *col, int sm_count, cudaStream_t s);

// ---- voxel-block hash map -----------------------------------------------------------------------
#define TSDF_ERR_HEAP_EXHAUSTED_CODE (-5)
#define TSDF_ERR_TABLE_FULL_CODE (-6)

struct HashDev {
    unsigned long long *keys; // [mask + 1]
    uint32_t *vals;           // [mask + 1] heap index or NONE
    uint32_t mask;
    uint32_t *entries;        // [mask + 1] slots of inserted keys
    int *n_entries;
    uint32_t *heap;           // [heap_cap] free block records
    int *heap_top;            // number of free records
    int heap_cap;
    float *blocks;            // [heap_cap][1536]
    uint32_t *known;          // bitmap over the bounded block grid: key already in the table
    int nbx, nby, nbz;        // block grid of the volume
    int nsx, nsy, nsz;        // super-block grid (4x4x4 blocks)
    int rank, world;
};

// Ray-walk constants of one frame (forward pose = inverse of Tinv, computed on the host).
struct WalkParams {
    double P[12];             // camera -> world, rows 0..2
    double ox, oy, oz;        // volume origin
    double inv_vs;            // 1 / voxel size
    double step;              // sample spacing in metres of z-depth
    double trunc;
    double fx, fy, cx, cy;
    double pose_scale;        // upper bound of the 2-norm of the pose's 3x3 block
    int W, H, tiles_x, tiles_y;
    int chunks;               // warps per tile: 32 samples each
    int dim_x, dim_y, dim_z;
};

void launch_hash_init(const HashDev &h, int sm_count, cudaStream_t s);
void launch_walk_coarse(const WalkParams &w, const HashDev &h, const float *tile_dmax, uint32_t *sb_flags,
                        FrameScratch *fs, cudaStream_t s);
void launch_refine_insert(const FrameParams &f, const CullParams &c, const HashDev &h, uint32_t *sb_flags, bool walked,
                          FrameScratch *fs, Totals *tot, cudaStream_t s);
void launch_compact_visible(const FrameParams &f, const CullParams &c, const HashDev &h, const uint32_t *sb_flags,
                            int max_entries, BrickRef *list, FrameScratch *fs, cudaStream_t s);
int integrate_hash_ctas_per_sm();
void launch_integrate_hash(const FrameParams &f, const VolumeParams &v, const HashDev &h, const BrickRef *list,
                           FrameScratch *fs, Totals *tot, UpdateCounters uc, int grid, cudaStream_t s);
void launch_scatter_dense(const HashDev &h, int dim_x, int dim_y, int dim_z, float *tsdf, float *wgt, float *col,
                          int sm_count, cudaStream_t s);
void launch_export_blocks(const HashDev &h, int capacity, int *count, int32_t *keys_out, float *payload, int sm_count,
                          cudaStream_t s);
void launch_scatter_block_list(const int32_t *keys, const float *payload, int n, int dim_x, int dim_y, int dim_z, float *tsdf,
                               float *wgt, float *col, bool fill, int sm_count, cudaStream_t s);
void launch_voxel_query(const HashDev &h, const int32_t *coords, int n, int mode, float *out, uint8_t *found, cudaStream_t s);
void launch_count_voxels(const HashDev &h, unsigned long long *total, int sm_count, cudaStream_t s);
void launch_remove_blocks(const HashDev &h, const int32_t *keys, int n, int *removed, cudaStream_t s);
