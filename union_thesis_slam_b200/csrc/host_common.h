// host_common.h -- host-side plumbing shared by the dense-grid and hash-map C ABI:
// error reporting, per-frame constants, and the host->device frame path (FRAME_SLOTS buffers in flight).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/tsdf_b200.h"
#include "launch.h"
#include "tsdf_common.cuh"

void tsdf_set_error(const char *fmt, ...);

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            tsdf_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return _e == cudaErrorMemoryAllocation ? TSDF_ERR_OOM : TSDF_ERR_CUDA;                  \
        }                                                                                           \
    } while (0)

#define BAD_ARG(cond, ...)                \
    do {                                  \
        if (cond) {                       \
            tsdf_set_error(__VA_ARGS__);  \
            return TSDF_ERR_BAD_ARG;      \
        }                                 \
    } while (0)

// Geometry common to both maps (grid_fusion.py:33-44 / hash_fusion.py:42-53).
struct VolumeGeom {
    int64_t dim[3];
    float origin[3];
    double vs, trunc;
    int nb[3];           // bricks per axis (global)
    double *d_lat = nullptr; // device: lat_x | lat_y | lat_z, each padded to a multiple of 8
    int lat_off[3];
    double max_abs_coord;

    int init(const int64_t dim_[3], const float origin_[3], double vs_, double trunc_);
    void release();
    VolumeParams params() const;
    // world-space radius that covers every voxel centre of a brick around the brick centre
    double brick_radius() const;
};

int tsdf_select_device(int device, int *sm_count);

// frame fan-out (capi_fanout.cu): ranks of one box and ring slots per rank
#define FANOUT_MAX_WORLD 16
#define FANOUT_MAX_SLOTS 8

// Per-warp update counters owned by a handle (see UpdateCounters).
#ifndef FRAME_SLOTS
#define FRAME_SLOTS 3
#endif
// FRAME_SLOTS: frames in flight per handle: preparation of frame n+2 may start while frame n integrates
struct CounterStore {
    unsigned long long *frame[FRAME_SLOTS] = {}; // one per frame slot
    unsigned long long *total = nullptr;
    int n = 0;
    int init(int n_warps);
    void release();
    UpdateCounters view(int slot, int first_warp = 0) const
    {
        return UpdateCounters{frame[slot] + first_warp, total + first_warp};
    }
    StatsBlock *d_stats = nullptr, *h_stats = nullptr; // device block + pinned host mirror
    // one kernel + one D2H copy + one stream sync; returns the pinned mirror
    int collect(int slot, const FrameScratch *fs, const Totals *tot, const int *n_entries, const int *heap_top,
                cudaStream_t s, const StatsBlock **out);
    // Non-blocking variant: post() enqueues the same kernel + D2H copy into ring slot (ticket % STATS_RING)
    // and records an event; wait() blocks on that event only, so a caller that reads frame k-1's
    // counters while frame k runs never stalls the pipeline (the upload of frame k+1 overlaps frame k).
    static const int STATS_RING = 16;
    StatsBlock *d_ring = nullptr, *h_ring = nullptr;
    cudaEvent_t ring_done[STATS_RING] = {};
    int64_t ring_frames[STATS_RING] = {};
    int64_t next_ticket = 0;
    int post(int slot, const FrameScratch *fs, const Totals *tot, const int *n_entries, const int *heap_top,
             cudaStream_t s, int64_t frames, int64_t *ticket);
    int wait(int64_t ticket, const StatsBlock **out, int64_t *frames);
};
int make_frame_params(FrameParams *f, CullParams *c, const VolumeGeom &g, int H, int W, const double K[9],
                      const double Tinv[16], double obs);

// Frame plumbing of one handle.
//  * Host frames go through pinned staging (or straight from the caller's buffer when that is
//    already pinned) on the `copy` stream.
//  * Per-frame preparation kernels (colour fold, tile pyramid, culling / ray walk, compaction) run
//    on the `pre` stream, the integrate kernel on `compute` (the dense grid's second, independent
//    classes of bricks on the `side` streams).  All per-frame scratch (folded colour,
//    FrameScratch, tile pyramid, and the handles' visible lists) exists FRAME_SLOTS times, so
//    the preparation of frames n+1 and n+2 overlaps the integration of frame n (the preparation
//    chain is three dependent launches; with two slots it sat on the critical path of short frames).
struct FrameIO {
    cudaStream_t compute = nullptr, copy = nullptr, pre = nullptr;
    // Snapshots of the counters (stats_post) are taken on their own stream behind the frame's integrate
    // events, so that they never order one frame's kernels before the next one's.
    cudaStream_t stat = nullptr;
    cudaEvent_t stat_done[FRAME_SLOTS] = {};
    bool stat_used = false;
    int stat_begin();  // `stat` waits for the integrate kernels of the last frame issued
    int stat_end();    // marks the snapshot's reads of the frame slot as done (the slot's next user waits for it)
    // sibling integrate streams of the dense grid (independent classes of bricks, see k_cull_bricks);
    // n_side is set by the handle: how many of them the frame being issued uses (0: all on `compute`)
    static const int MAX_SIDE = DENSE_MAX_CLASSES - 1;
    cudaStream_t side[MAX_SIDE] = {};
    int n_side = 0;
    cudaEvent_t side_int_done[MAX_SIDE][FRAME_SLOTS] = {}, side_slot_free[MAX_SIDE][FRAME_SLOTS] = {}, side_end[MAX_SIDE] = {};
    bool side_pending[MAX_SIDE] = {}; // work issued on side[k] that `compute` has not been made to wait for
    int ended_side = 0;               // side streams used by the last frame issued (for last_ms)
    int side_slot_last = 0;           // frame slot whose side_int_done events were recorded last
    int side_used = 0;                // side streams ever used: their events are the only ones worth waiting on
    int cap_pix = 0;
    uint8_t *d_rgb[FRAME_SLOTS] = {};
    double *d_depth[FRAME_SLOTS] = {};
    uint8_t *h_rgb[FRAME_SLOTS] = {};
    double *h_depth[FRAME_SLOTS] = {};
    uint16_t *d_depth16[FRAME_SLOTS] = {}; // raw sensor depth (frame ingest path)
    // per-frame scratch, one set per frame slot: [FrameScratch | tile pyramid levels 1..3]
    // (cleared every frame), then pyramid level 0
    // The dense grid applies up to DENSE_MAX_BATCH frames per brick visit: every frame slot then holds
    // `subs` prepared frames.  Flat index = sub * FRAME_SLOTS + slot, so [slot] alone is sub 0 (hash map).
    static const int MAX_PREP = FRAME_SLOTS * DENSE_MAX_BATCH;
    int subs = 1;
    FramePixel *px_slot[MAX_PREP] = {}; // prepared frames (depth + folded colour, 16 B per pixel)
    uint8_t *scratch_slot[MAX_PREP] = {};
    TilePyramid pyr_slot[MAX_PREP] = {};
    unsigned frame_tag = 0;                     // tag of the last prepared frame (see FrameScratch); 0 is never used
    unsigned next_tag() { return ++frame_tag ? frame_tag : ++frame_tag; }
    int slot = FRAME_SLOTS - 1; // scratch slot of the frame being issued (advanced by begin_kernels)
    // views of the current slot (after a frame: of the last frame issued)
    FramePixel *d_px = nullptr;
    FrameScratch *fs = nullptr;
    TilePyramid pyr{};
    int tiles_H = 0, tiles_W = 0;
    cudaEvent_t h2d_done[FRAME_SLOTS] = {}, slot_free[FRAME_SLOTS] = {};
    cudaEvent_t pre_done[FRAME_SLOTS] = {}, int_done[FRAME_SLOTS] = {};
    cudaEvent_t k_begin = nullptr, k_end = nullptr;
    cudaEvent_t k_mark[3] = {nullptr, nullptr, nullptr}; // optional stage boundaries inside a frame
    int n_marks = 0;
    int next = 0, cur = -1;
    int borrowed = -1; // upload slot whose H2D copy still reads the caller's pinned buffers
    bool timed = false;
    // Frame timing events (k_begin / k_end / side_end) cost three API calls per frame; the dense
    // handle records them only while timing or stage profiling is switched on.
    bool timing = true;

    int init();
    int ensure(int npix);
    // stages the frame and makes `pre` wait for it; returns device pointers of the slot
    int upload(const uint8_t *rgb, const double *depth, int npix, const uint8_t **d_rgb_out,
               const double **d_depth_out);
    // frame ingest: raw uint16 depth (grid_demo1.py:81-83) is uploaded as is (4x fewer bytes) and
    // converted on the device to the float64 metres the rule consumes
    int upload_u16(const uint8_t *rgb, const uint16_t *depth_raw, int npix, double depth_scale, int invalid_raw,
                   const uint8_t **d_rgb_out, const double **d_depth_out);
    int finish_upload(); // waits until the copy engine has left the caller's buffers (end of an integrate call)
    // A frame produced on the device by a stream of the caller (a collective's stream, a decoder): adopt()
    // orders the next frame's preparation kernel after the work enqueued on `producer` so far; once that
    // kernel has consumed the frame (release_upload_after_prep) `producer` is made to wait for it, so
    // whatever the caller enqueues there next may overwrite the same buffers.
    cudaStream_t ext = nullptr;
    cudaEvent_t ext_ready = nullptr, ext_consumed = nullptr;
    int adopt(cudaStream_t producer);
    // raw uint16 depth already on the device: converted into the next upload slot's float64 image (stream `pre`)
    int convert_u16(const uint16_t *d_raw, int npix, double depth_scale, int invalid_raw, const double **d_depth_out);
    int ensure_px(int npix);
    int set_subs(int n);              // prepared frames per slot (drains the streams when it grows)
    static int prep_index(int slot, int sub) { return sub * FRAME_SLOTS + slot; }
    int release_upload_after_prep();  // the uploaded frame is consumed by the preparation kernel alone
    int ensure_tiles(int H, int W);   // (re)allocates fs + pyramid for this image size
    // picks the frame's scratch slot and waits (on `pre`) until the frame FRAME_SLOTS calls ago has left it
    int begin_kernels();
    int prepared();    // preparation kernels issued: `compute` waits for them
    int end_kernels(); // integrate issued: records timing, frees the scratch and upload slots
    int mark(); // records a stage boundary (up to 3 per frame)
    int join(); // `compute` waits (on the device) for everything issued on the side streams so far
    int sync_side(); // host waits for the side streams
    int last_ms(float *ms);
    int last_stage_ms(float ms[4], int *n); // durations of the stages delimited by mark()
    void release();
};
