/*
 * tsdf_b200.h -- C ABI of libtsdf_b200.so: B200 (sm_100a) TSDF integration into a dense brick
 * grid and into a voxel-block hash map.
 *
 * This is the drop-in boundary for the mapping hot path of DiWu9/Union-Thesis-SLAM.  Every entry
 * point names the reference interface it replaces (file:line under the reference root).  Plain C
 * types only (pointers + sizes): bind with ctypes / cffi / a 50-line extension; see INTEGRATION.md.
 *
 * Conventions
 *   - every function returns 0 (TSDF_OK) or a negative tsdf_status; tsdf_last_error() gives the
 *     text of the calling thread's last failure.  There is NO CPU fallback: without a usable CUDA
 *     device create() fails with TSDF_ERR_NO_DEVICE.
 *   - a handle is one rank's shard, bound to one CUDA device and its own streams; not thread-safe.
 *   - integrate() is stream-ordered and may return before the kernels finish; download / stats /
 *     sync wait.  Host frame buffers are borrowed only for the duration of the call.
 *   - frames: rgb = H*W*3 uint8 (R,G,B), depth = H*W float64 metres with 0 = invalid,
 *     K = 3x3 row-major intrinsics, Tinv = 4x4 row-major float64 inverse of the camera->world pose
 *     (the host computes np.linalg.inv(cam_pose) exactly like grid_fusion.py:265 and passes it
 *     down; the device never inverts a pose), obs_weight > 0.
 *   - voxel arithmetic follows the reference CPU path bit for bit.
 */
#ifndef TSDF_B200_H
#define TSDF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum tsdf_status {
    TSDF_OK = 0,
    TSDF_ERR_BAD_ARG = -1,
    TSDF_ERR_NO_DEVICE = -2,
    TSDF_ERR_CUDA = -3,
    TSDF_ERR_OOM = -4,
    TSDF_ERR_HEAP_EXHAUSTED = -5, /* voxel-block heap has no free block left */
    TSDF_ERR_TABLE_FULL = -6,     /* hash table probe sequence exhausted */
    TSDF_ERR_CANDIDATES = -7      /* per-frame candidate / visible list overflow */
} tsdf_status;

typedef struct tsdf_grid tsdf_grid; /* opaque: dense brick grid shard */
typedef struct tsdf_hash tsdf_hash; /* opaque: voxel-block hash map shard */

#define TSDF_BRICK 8          /* voxels per brick edge (8^3 = 512 voxels) */
#define TSDF_BRICK_FLOATS 1536 /* one brick record: 512 tsdf | 512 weight | 512 colour, float32 */

const char *tsdf_last_error(void);
/* Library / device facts: sm count, HBM bytes, compiled arch.  No handle needed. */
int tsdf_device_info(int device, int *sm_count, int64_t *total_mem, int *cc_major, int *cc_minor);

/* ---------------------------------------------------------------------------------------------
 * Dense grid.  Replaces grid_fusion.TSDFVolume (grid_fusion.py:19-320).
 * --------------------------------------------------------------------------------------------- */

/* TSDFVolume.__init__ (grid_fusion.py:22-57): dim = _vol_dim, origin = _vol_origin (float32),
 * voxel_size, trunc = _trunc_margin.  State starts at tsdf=1, weight=0, colour=0.
 * Sharding: the grid is cut into x-sheets one brick (8 voxels) thick; sheet s belongs to rank
 * s % world.  rank=0, world=1 gives the whole grid.  No per-frame communication is needed. */
int tsdf_grid_create(const int64_t dim[3], const float origin[3], double voxel_size, double trunc,
                     int device, int rank, int world, tsdf_grid **out);
int tsdf_grid_destroy(tsdf_grid *g);

/* TSDFVolume.integrate (grid_fusion.py:214-314, CPU branch semantics) -- HOST frame buffers. */
int tsdf_grid_integrate(tsdf_grid *g, const uint8_t *rgb, const double *depth, int H, int W,
                        const double K[9], const double Tinv[16], double obs_weight);
/* Frame ingest (grid_demo1.py:80-84): raw uint16 sensor depth.  The device computes
 * depth = (double)raw / depth_scale (IEEE division, bit-identical to `astype(float); /= 1000.`) and
 * maps raw == invalid_raw to 0 (`depth_im[depth_im == 65.535] = 0` for invalid_raw = 65535; pass -1
 * to disable).  1.54 MB instead of 3.38 MB per 640x480 frame over PCIe. */
int tsdf_grid_integrate_u16(tsdf_grid *g, const uint8_t *rgb, const uint16_t *depth_raw, int H, int W,
                            const double K[9], const double Tinv[16], double obs_weight, double depth_scale,
                            int invalid_raw);
/* Same, frame already resident in device memory (throughput runs; no H2D inside). */
int tsdf_grid_integrate_device(tsdf_grid *g, const uint8_t *d_rgb, const double *d_depth, int H, int W,
                               const double K[9], const double Tinv[16], double obs_weight);
/* Same for the sensor's raw uint16 depth resident in device memory (converted as in tsdf_grid_integrate_u16). */
int tsdf_grid_integrate_device_u16(tsdf_grid *g, const uint8_t *d_rgb, const uint16_t *d_depth_raw, int H, int W,
                                   const double K[9], const double Tinv[16], double obs_weight, double depth_scale,
                                   int invalid_raw);
/* Frame fan-out (the "each frame's depth, colour and pose broadcast to all ranks" step of the multi-GPU scheme;
 * the reference has no counterpart: grid_demo1.py:76-87 reads every frame in the one process it has).
 * Declares that the device buffers passed to the NEXT tsdf_grid_integrate_device[_u16] call are being written
 * by work already enqueued on the caller's CUDA stream `producer_stream` (an upload, an NCCL broadcast): the
 * handle orders its preparation kernel after that work, and orders `producer_stream` after the kernel has
 * consumed the frame, so the caller may enqueue the next write into the same buffers at once.  No host wait. */
int tsdf_grid_frame_source(tsdf_grid *g, void *producer_stream);

/* TSDFVolume.get_volume (grid_fusion.py:316-320) plus the weight plane (_weight_vol_cpu).
 * Writes this rank's sheets, un-bricked, as C-order [nx_local][Y][Z] float32 host arrays; any
 * pointer may be NULL.  nx_local and the global x index of every local row come from
 * tsdf_grid_local_x(). */
int tsdf_grid_download(tsdf_grid *g, float *tsdf, float *weight, float *color);
/* Same into DEVICE buffers (for the NCCL gather done by the caller's torch.distributed). */
int tsdf_grid_download_device(tsdf_grid *g, float *d_tsdf, float *d_weight, float *d_color);
/* n_x: number of local x rows (voxels).  xs (may be NULL): their global x indices, ascending. */
int tsdf_grid_local_x(tsdf_grid *g, int64_t *n_x, int64_t *xs);

/* counters[0] voxels updated by the last frame, [1] voxels updated since create,
 * [2] bricks visited by the last frame, [3] frames integrated. */
int tsdf_grid_stats(tsdf_grid *g, int64_t counters[4]);
/* The same counters without stalling the pipeline: post() enqueues a snapshot behind everything issued
 * so far (one tiny kernel + a 64-byte device->host copy into pinned memory) and returns a ticket at once;
 * wait() blocks until THAT snapshot has landed -- not until the device is idle -- and returns it.  A
 * caller that integrates frame k, posts, and then reads the ticket of frame k-1 gets every frame's
 * result while the upload of frame k+1 overlaps the integration of frame k (the reference's loop,
 * grid_demo1.py:76-87, has no per-frame read-back at all).  The last 16 tickets stay readable.
 * counters[1] (running total) is -1 in a snapshot: later frames may already be adding to it;
 * tsdf_grid_stats reports it. */
int tsdf_grid_stats_post(tsdf_grid *g, int64_t *ticket);
int tsdf_grid_stats_wait(tsdf_grid *g, int64_t ticket, int64_t counters[4]);
int tsdf_grid_sync(tsdf_grid *g);
/* Measurement aids.  Timing events are recorded only while profiling is on:
 *   set_profiling(2): frame timing -- last_kernel_ms() = device milliseconds from the first
 *     preparation kernel to the end of the last integrate kernel of the last frame (waits for it);
 *   set_profiling(1): stage timing -- additionally the frame's integrate kernels run on one stream
 *     behind an extra event, and last_stage_ms() returns n = 2 durations {prepare+cull, integrate};
 *   set_profiling(0) (default): no timing events (three API calls less per frame).
 * stream() exposes the handle's cudaStream_t so callers can record their own events on it. */
int tsdf_grid_last_kernel_ms(tsdf_grid *g, float *ms);
int tsdf_grid_set_profiling(tsdf_grid *g, int on);
int tsdf_grid_last_stage_ms(tsdf_grid *g, float ms[4], int *n);
int tsdf_grid_stream(tsdf_grid *g, void **stream);
/* The integrate kernels of a frame run on two device streams (two independent halves of the
 * bricks).  flush() makes the stream returned by tsdf_grid_stream() wait -- on the device, the
 * host does not block -- for every kernel issued so far, so that an event recorded on it
 * afterwards marks the completion of all of them.  Read-out calls and sync() do this themselves. */
int tsdf_grid_flush(tsdf_grid *g);
/* Temporal blocking.  The reference applies one frame per integrate call (grid_fusion.py:214); its
 * per-voxel rule has no cross-voxel term, so applying the frames k, k+1, ... to a voxel back to back
 * gives the same bits as applying them volume by volume.  With frames > 1 (max 4) every integrate
 * call prepares its frame at once but the volume update is deferred until `frames` calls have
 * accumulated; then each visible brick is read once, updated by all of them in call order, and
 * written once.  stats / download / sync / flush apply whatever is pending first, so results are
 * always those of the calls made so far.  While frames > 1, stats' counters[0] and [2] describe the
 * last batch issued, and device buffers passed to integrate_device are consumed before the call's
 * successor work on the handle's streams, exactly as with frames == 1.  Default 1. */
int tsdf_grid_set_batch(tsdf_grid *g, int frames);

/* ---------------------------------------------------------------------------------------------
 * Voxel-block hash map.  Replaces hash_fusion.HashTable (hash_fusion.py:29-507) on the
 * integrate / get_volume path.  Storage is 8^3-voxel blocks found through a lock-free
 * open-addressing table keyed on packed 64-bit block coordinates (hash: hash_fusion.py:182-190
 * primes), block storage from an atomic free-list heap.  A block exists iff at least one of its
 * voxels has been updated, so {exported blocks} == {v // 8 : the reference map has an entry v}.
 * --------------------------------------------------------------------------------------------- */

/* HashTable.__init__ (hash_fusion.py:34-69).  table_slots: open-addressing capacity (rounded up
 * to a power of two); heap_blocks: voxel-block capacity of this rank.  Sharding: a block belongs
 * to rank owner(key) % world (owner hash decorrelated from the slot hash). */
int vh_create(const int64_t dim[3], const float origin[3], double voxel_size, double trunc,
              int64_t table_slots, int64_t heap_blocks, int device, int rank, int world,
              tsdf_hash **out);
int vh_destroy(tsdf_hash *h);

/* HashTable.integrate (hash_fusion.py:103-145).  obs_weight follows the dense rule. */
int vh_integrate(tsdf_hash *h, const uint8_t *rgb, const double *depth, int H, int W,
                 const double K[9], const double Tinv[16], double obs_weight);
int vh_integrate_u16(tsdf_hash *h, const uint8_t *rgb, const uint16_t *depth_raw, int H, int W, const double K[9],
                     const double Tinv[16], double obs_weight, double depth_scale, int invalid_raw);
int vh_integrate_device(tsdf_hash *h, const uint8_t *d_rgb, const double *d_depth, int H, int W,
                        const double K[9], const double Tinv[16], double obs_weight);
int vh_integrate_device_u16(tsdf_hash *h, const uint8_t *d_rgb, const uint16_t *d_depth_raw, int H, int W,
                            const double K[9], const double Tinv[16], double obs_weight, double depth_scale,
                            int invalid_raw);
/* As tsdf_grid_frame_source, for the hash map. */
int vh_frame_source(tsdf_hash *h, void *producer_stream);

/* HashTable.count_num_hash_entries / get_load_factor / get_num_collisions analogues
 * (hash_fusion.py:147-180).  stats[0] allocated blocks, [1] table keys (incl. keys whose block
 * never received an update), [2] table slots, [3] heap capacity, [4] free heap blocks,
 * [5] voxels updated by the last frame, [6] voxels updated since create, [7] blocks visited by
 * the last frame, [8] longest probe sequence seen, [9] keys inserted by the last frame,
 * [10] frames integrated, [11] candidate keys emitted by the last frame's ray walk. */
int vh_stats(tsdf_hash *h, int64_t stats[12]);
int vh_num_blocks(tsdf_hash *h, int64_t *n);
/* Block-list read-out for the multi-GPU gather ("read-out = gather of per-rank (keys, voxel payload)
 * lists"): this rank's blocks into DEVICE buffers -- keys [capacity][3] int32 block coordinates, payload
 * [capacity][3][512] float32 (tsdf | weight | colour in the float encoding); capacity 0 counts only -- and the
 * inverse: any block list scattered into dense DEVICE volumes [X][Y][Z] (fill != 0 first writes the defaults
 * tsdf 1, weight 0, colour 0, hash_fusion.py:447-448).  The caller's torch.distributed exchanges the lists. */
int vh_export_blocks_device(tsdf_hash *h, int64_t capacity, int64_t *n, int32_t *d_keys, float *d_payload);
int vh_scatter_blocks_device(tsdf_hash *h, const int32_t *d_keys, const float *d_payload, int64_t n, int fill,
                             float *d_tsdf, float *d_weight, float *d_color);
/* Per-voxel view of the block map -- HashTable.get_hash_entry / get_voxel (hash_fusion.py:285-320) and
 * remove (hash_fusion.py:322-393) on the table that integrate() fills.  coords: n voxel indices
 * [n][3]; a voxel has an entry iff its block owns storage and its weight is > 0 (the reference
 * creates the entry when it first updates the voxel, hash_fusion.py:138-143).
 *   mode 0: out[i] = {sdf, weight, colour (float encoding B*65536+G*256+R)}, found[i] = 0/1
 *   mode 1: like 0, and every found voxel returns to the no-entry state (sdf 1, weight 0, colour 0).
 * With world > 1 each rank answers for the blocks it owns.  out / found are HOST arrays, may be NULL. */
int vh_voxel_query(tsdf_hash *h, const int32_t *coords, int64_t n, int mode, float *out, uint8_t *found);
/* count_num_hash_entries (hash_fusion.py:147-154) in the reference's unit: voxels that have an entry. */
int vh_count_voxels(tsdf_hash *h, int64_t *count);
/* Non-blocking variant of vh_stats, same contract as tsdf_grid_stats_post / _wait. */
int vh_stats_post(tsdf_hash *h, int64_t *ticket);
int vh_stats_wait(tsdf_hash *h, int64_t ticket, int64_t stats[12]);

/* Block export: keys[n][3] int32 block coordinates (voxel // 8), payload[n][1536] float32
 * (tsdf | weight | colour, voxel order x*64 + y*8 + z inside the block).  Call with NULL
 * pointers to get n only.  Order is unspecified (compare as a set). */
int vh_export_blocks(tsdf_hash *h, int64_t capacity, int64_t *n, int32_t *keys, float *payload);

/* HashTable.get_volume (hash_fusion.py:442-463): scatter into dense C-order [X][Y][Z] float32
 * volumes (tsdf default 1, weight 0, colour 0).  Host pointers, any may be NULL.  With world > 1
 * each rank writes only its own blocks; the caller merges (blocks are disjoint). */
int vh_to_dense(tsdf_hash *h, float *tsdf, float *weight, float *color);
int vh_to_dense_device(tsdf_hash *h, float *d_tsdf, float *d_weight, float *d_color);

/* Free every block whose key is listed (keys[n][3]); their storage returns to the heap.
 * Mirrors HashTable.remove (hash_fusion.py:322-393) at block granularity. */
int vh_remove_blocks(tsdf_hash *h, const int32_t *keys, int64_t n, int64_t *removed);
int vh_sync(tsdf_hash *h);
int vh_last_kernel_ms(tsdf_hash *h, float *ms);
/* The cudaStream_t the integrate kernel of every frame runs on (it starts after the frame's
 * allocation / compaction kernels): an event recorded on it after a frame marks that frame's end. */
int vh_stream(tsdf_hash *h, void **stream);

/* ---------------------------------------------------------------------------------------------
 * Surface extraction for get_mesh / get_point_cloud (grid_fusion.py:322-360,
 * hash_fusion.py:465-507: marching cubes at level 0 on the volume get_volume() returns, vertex
 * colours looked up at round(vertex)).  d_tsdf / d_color: dense C-order [X][Y][Z] float32 planes in
 * DEVICE memory (tsdf_grid_download_device / vh_to_dense_device); d_color may be NULL.
 * Vertices are world coordinates (voxel index * voxel_size + origin), one per crossed grid edge,
 * faces index them; normals are unit TSDF gradients; colours are r, g, b bytes.
 * --------------------------------------------------------------------------------------------- */
typedef struct tsdf_mesh tsdf_mesh;
int tsdf_mesh_extract(const float *d_tsdf, const float *d_color, const int64_t dim[3], double voxel_size,
                      const float origin[3], int device, tsdf_mesh **out, int64_t *n_verts, int64_t *n_faces);
/* verts[n_verts][3] float32, faces[n_faces][3] int32, norms[n_verts][3] float32, colors[n_verts][3]
 * uint8; host pointers, any may be NULL. */
int tsdf_mesh_download(tsdf_mesh *m, float *verts, int32_t *faces, float *norms, uint8_t *colors);
int tsdf_mesh_destroy(tsdf_mesh *m);

/* ---------------------------------------------------------------------------------------------
 * Per-voxel hash CONTAINER with the reference's slot layout.
 * Replaces the container methods of hash_fusion.HashTable: hash_function (hash_fusion.py:182-190),
 * add_hash_entry / _add_to_linked_list / _find_next_free_space_and_add_entry (:199-276),
 * get_hash_entry (:285-320), remove_hash_entry (:331-393), needs_resize / double_table_size
 * (:156-161, :414-437), the occupancy metrics (:147-180) and the per-voxel integrate loop
 * (:135-145 with data_structures/voxel.py:19-49).  Buckets of `bucket_size` slots (reference: 5,
 * last slot = head of the bucket's collision chain), chains as (bucket, slot) offsets, tables in
 * HBM.  Mutating operations are order-dependent in the reference (the slot an entry lands in
 * depends on every earlier insert), so vbt_apply executes a batch strictly in order on the device
 * and returns the reference's own return values; vbt_lookup is a parallel read-only batch.
 * --------------------------------------------------------------------------------------------- */
typedef struct tsdf_buckets tsdf_buckets;
enum { VBT_OP_ADD = 0, VBT_OP_GET = 1, VBT_OP_REMOVE = 2, VBT_OP_INTEGRATE = 3 };

int vbt_create(int64_t n_buckets, int bucket_size, double load_factor, int device, tsdf_buckets **out);
int vbt_destroy(tsdf_buckets *h);
/* ops[n] (VBT_OP_*), positions[n][3] voxel coordinates, args[n][5] (may be NULL) =
 * {dist, folded colour} for INTEGRATE, {.., .., sdf, colour, weight} voxel payload for ADD.
 * out[n][2]: ADD / GET / INTEGRATE -> (bucket, slot) where the entry is stored / was found, or
 * (-1, -1); REMOVE -> (1 removed / 0 not found, -1).  An ADD that finds the table at its load
 * factor doubles it first, like add_hash_entry. */
int vbt_apply(tsdf_buckets *h, const int32_t *ops, const int64_t *positions, const double *args, int64_t n,
              int64_t *out);
/* get_hash_entry for a batch, one device thread per query: out[n][2] = (bucket, slot) or (-1, -1);
 * voxels (may be NULL) [n][3] = (sdf, colour, weight) of the entries found. */
int vbt_lookup(tsdf_buckets *h, const int64_t *positions, int64_t n, int64_t *out, double *voxels);
int vbt_double_table_size(tsdf_buckets *h);
/* stats[0] buckets (table size), [1] non-empty buckets, [2] entries stored, [3] buckets holding
 * more than one entry (get_num_collisions), [4] resizes so far, [5] bucket size. */
int vbt_stats(tsdf_buckets *h, int64_t stats[6]);
/* The reference's HashTable.integrate as a whole (hash_fusion.py:103-145) on the per-voxel container:
 * every voxel of the bounded volume is projected in vox_coords order with the reference's own
 * arithmetic, the voxels it would update are streamed in that order through the per-voxel loop
 * (look-up, insert on a miss, float64 running mean).  obs_weight does not exist here: the reference
 * never passes it on (:141, :145).  Frame arguments as tsdf_grid_integrate (host pointers).
 * *n_valid (may be NULL): voxels updated by the frame. */
int vbt_integrate_frame(tsdf_buckets *h, const int64_t dim[3], const float origin[3], double voxel_size, double trunc,
                        const uint8_t *rgb, const double *depth, int H, int W, const double K[9], const double Tinv[16],
                        int64_t *n_valid);
/* get_volume of the reference hash map (hash_fusion.py:442-463): [X][Y][Z] float32 planes, tsdf 1 /
 * colour 0 where no entry is stored, else float32(sdf) / float32(colour) of the entry's voxel. */
int vbt_to_dense(tsdf_buckets *h, const int64_t dim[3], float *tsdf, float *color);
/* Slot layout in (bucket, slot) order.  rows == NULL: only *count.  rows[count][7] =
 * {bucket, slot, x, y, z, offset bucket, offset slot} (-1, -1 = empty offset); voxels (may be
 * NULL) [count][3] = (sdf, colour, weight), NaN sdf for an entry stored without a voxel. */
int vbt_dump(tsdf_buckets *h, int64_t capacity, int64_t *count, int64_t *rows, double *voxels);

/* ---- frame fan-out over NVLink peer memory (multi-GPU, one process per GPU) -------------------------------
 * The "each frame's depth, colour and pose broadcast to all ranks" step of the multi-GPU scheme; the reference is
 * one process and reads every frame itself (grid_fusion_demos/grid_demo1.py:76-87), so there is nothing it binds
 * here -- a multi-process host of the reference would.  A frame is uploaded once, on its source rank, into a ring
 * slot of that rank's arena; every peer's copy engine pulls it over NVLink from there (CUDA IPC mappings) into the
 * same slot of its own arena.  "Staged" and "pulled" are sequence-number flags in the arenas, written by peer
 * stores and waited for on the device, never on the host.
 *
 * create -> export (64-byte handle) -> [the caller all-gathers the handles] -> connect (handles of all ranks, in rank
 * order; the own entry is ignored).  Per frame `seq` (0, 1, 2, ... on every rank alike), source rank `src`:
 *   src only : tsdf_fanout_push(seq, host depth, host colour)      -- the slot is seq % slots
 *   all ranks: tsdf_fanout_acquire(seq, src, ...) -> device pointers of the frame and the stream they arrive on;
 *              tsdf_grid_frame_source / vh_frame_source(that stream); tsdf_grid_integrate_device[_u16] / vh_...;
 *              tsdf_fanout_release(seq)                            -- marks the slot's consumption on that stream
 * Pinned host buffers given to push are read asynchronously: they must stay unchanged until the rank's next push or
 * tsdf_fanout_drain returns.  drain also reports a peer that failed to deliver / consume within the time-out
 * (default 20 s).  Before destroy the caller synchronises the ranks (peers write into this rank's arena). */
typedef struct tsdf_fanout tsdf_fanout;
int tsdf_fanout_create(int device, int rank, int world, int64_t slot_bytes, int slots, tsdf_fanout **out);
int tsdf_fanout_export(tsdf_fanout *f, uint8_t handle[64]);
int tsdf_fanout_connect(tsdf_fanout *f, const uint8_t *handles /* [world][64] */);
/* depth goes to offset 0 of the slot, colour to rgb_offset (>= depth_bytes). */
int tsdf_fanout_push(tsdf_fanout *f, int64_t seq, const void *depth, int64_t depth_bytes, const void *rgb,
                     int64_t rgb_bytes, int64_t rgb_offset);
/* frame_bytes = rgb_offset + rgb_bytes: what a non-source rank pulls from the source's slot. */
int tsdf_fanout_acquire(tsdf_fanout *f, int64_t seq, int src, int64_t frame_bytes, int64_t rgb_offset, void **d_depth,
                        void **d_rgb, void **stream);
int tsdf_fanout_release(tsdf_fanout *f, int64_t seq);
/* The per-frame sequence above in one call, for a dense grid (hash == NULL) or a hash map (grid == NULL): push if
 * src is this rank (depth: float64 metres, or raw uint16 with raw_u16 != 0; colour at the 256-byte-rounded offset
 * behind it), acquire, frame_source, integrate_device[_u16], release.  depth / rgb are read on the source only. */
int tsdf_fanout_integrate(tsdf_fanout *f, tsdf_grid *grid, tsdf_hash *hash, int64_t seq, int src, const void *depth,
                          const void *rgb, int H, int W, int raw_u16, const double K[9], const double Tinv[16],
                          double obs_weight, double depth_scale, int invalid_raw);
int tsdf_fanout_drain(tsdf_fanout *f);
int tsdf_fanout_set_timeout(tsdf_fanout *f, double seconds);
int tsdf_fanout_destroy(tsdf_fanout *f);

#ifdef __cplusplus
}
#endif
#endif /* TSDF_B200_H */
