// capi_hash.cu -- C ABI of the voxel-block hash map (include/tsdf_b200.h, vh_*).
#include <algorithm>
#include <new>
#include <vector>

#include "host_common.h"

static cudaError_t alloc_lists(BrickRef *lists[FRAME_SLOTS], size_t cap)
{
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < FRAME_SLOTS && e == cudaSuccess; ++i) e = cudaMalloc(&lists[i], cap * sizeof(BrickRef));
    return e;
}

#define WALK_SB_PER_TILE 400.0 // the scan still wins at 83 super-blocks per tile (16x the scene's bounds, profiles/r2_hash_alloc.md)
struct tsdf_hash {
    int device = 0, sm_count = 0, rank = 0, world = 1;
    VolumeGeom geom;
    FrameIO io;
    HashDev dev{};
    int64_t table_slots = 0, heap_blocks = 0;
    BrickRef *vis_list[FRAME_SLOTS] = {}; // one per frame slot
    Totals *tot = nullptr;
    int *scratch_count = nullptr; // device int used by export / remove
    CounterStore counters;
    uint32_t *sb_flags = nullptr; // per-frame bitmap of super-blocks touched by the ray walk
    size_t sb_words = 0;
    int integrate_grid = 0;
    int64_t frames = 0;
    bool unit_w = true;
    double walk_step_voxels = 16.0; // ray-walk sample spacing: half a super-block
    // Which super-blocks can the frame touch?  false: K3 tests every super-block of the bounded grid
    // itself (one warp each, sphere vs frustum + tile depth pyramid; cost scales with the volume).
    // true: K2's cone walk per pixel tile flags them first (cost scales with pixels x depth).
    // Both are conservative and K3's per-block test decides, so the allocated set is identical;
    // TSDF_HASH_ALLOC=walk|grid overrides the default for measurement (profiles/r1_hash_alloc.md).
    // Allocation front end: which 32^3-voxel super-blocks can the frame touch?
    //   grid (0): one warp per super-block of the bounded volume tests it against the frame (cost ~ volume)
    //   walk (1): the per-pixel-tile cone walk flags them (cost ~ pixels x depth)
    //   auto (2, default): per frame, whichever the cost model below says is cheaper -- the walk once the
    //   bounded volume holds more than WALK_SB_PER_TILE super-blocks per 8x8-pixel tile
    //   (profiles/r2_hash_alloc.md).  TSDF_HASH_ALLOC=grid|walk|auto overrides.
    int alloc_mode = 2;
};

static int hash_free(tsdf_hash *h)
{
    if (!h) return TSDF_OK;
    cudaSetDevice(h->device);
    h->io.release();
    h->geom.release();
    void *ptrs[] = {h->dev.keys, h->dev.vals, h->dev.entries, h->dev.n_entries, h->dev.heap, h->dev.heap_top,
                    h->dev.blocks, h->dev.known, h->vis_list[0], h->vis_list[1], h->vis_list[2], h->tot, h->scratch_count,
                    h->sb_flags};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    h->counters.release();
    delete h;
    return TSDF_OK;
}

extern "C" int vh_create(const int64_t dim[3], const float origin[3], double voxel_size, double trunc,
                         int64_t table_slots, int64_t heap_blocks, int device, int rank, int world, tsdf_hash **out)
{
    BAD_ARG(!out, "out is NULL");
    *out = nullptr;
    BAD_ARG(world < 1 || rank < 0 || rank >= world, "bad rank/world %d/%d", rank, world);
    BAD_ARG(table_slots <= 0 || table_slots > ((int64_t)1 << 30), "table_slots %lld out of range", (long long)table_slots);
    BAD_ARG(heap_blocks <= 0 || heap_blocks > ((int64_t)1 << 30), "heap_blocks %lld out of range", (long long)heap_blocks);
    int sm = 0;
    int rc = tsdf_select_device(device, &sm);
    if (rc) return rc;
    tsdf_hash *h = new (std::nothrow) tsdf_hash;
    BAD_ARG(!h, "out of host memory");
    h->device = device;
    h->sm_count = sm;
    h->rank = rank;
    h->world = world;
    if ((rc = h->geom.init(dim, origin, voxel_size, trunc)) || (rc = h->io.init())) {
        hash_free(h);
        return rc;
    }
    for (int a = 0; a < 3; ++a)
        if (h->geom.nb[a] >= (1 << 15)) { // visible-list entries pack bx | by << 16
            hash_free(h);
            tsdf_set_error("volume has too many blocks along axis %d (limit 32767)", a);
            return TSDF_ERR_BAD_ARG;
        }
    // the presence bitmap, the super-block flags and the launch grids index the bounded block grid with
    // 32-bit integers: a wrapped index would alias `known` bits and silently drop allocations
    if ((int64_t)h->geom.nb[0] * h->geom.nb[1] * h->geom.nb[2] >= ((int64_t)1 << 31)) {
        const long long nblk = (long long)h->geom.nb[0] * h->geom.nb[1] * h->geom.nb[2];
        hash_free(h);
        tsdf_set_error("bounded volume holds %lld blocks (limit 2^31 - 1): use a coarser voxel size or smaller bounds", nblk);
        return TSDF_ERR_BAD_ARG;
    }
    if (const char *e = getenv("TSDF_HASH_ALLOC")) h->alloc_mode = !strcmp(e, "walk") ? 1 : (!strcmp(e, "grid") ? 0 : 2);
    int64_t S = 1024;
    while (S < table_slots) S <<= 1;
    h->table_slots = S;
    h->heap_blocks = heap_blocks;
    HashDev &d = h->dev;
    d.mask = (uint32_t)(S - 1);
    d.heap_cap = (int)heap_blocks;
    d.rank = rank;
    d.world = world;
    d.nbx = h->geom.nb[0]; d.nby = h->geom.nb[1]; d.nbz = h->geom.nb[2];
    d.nsx = (d.nbx + 3) / 4; d.nsy = (d.nby + 3) / 4; d.nsz = (d.nbz + 3) / 4;
    const size_t known_words = ((size_t)d.nbx * d.nby * d.nbz + 31) / 32;
    h->sb_words = ((size_t)d.nsx * d.nsy * d.nsz + 31) / 32;
    // a key exists at most once per block of the bounded grid: that bounds the visible list too
    const size_t list_cap = (size_t)std::min<int64_t>(S, (int64_t)d.nbx * d.nby * d.nbz);
    cudaError_t e;
    if ((e = cudaMalloc(&d.keys, S * sizeof(unsigned long long))) != cudaSuccess ||
        (e = cudaMalloc(&d.vals, S * sizeof(uint32_t))) != cudaSuccess ||
        (e = cudaMalloc(&d.entries, S * sizeof(uint32_t))) != cudaSuccess ||
        (e = cudaMalloc(&d.n_entries, sizeof(int))) != cudaSuccess ||
        (e = cudaMalloc(&d.heap, heap_blocks * sizeof(uint32_t))) != cudaSuccess ||
        (e = cudaMalloc(&d.heap_top, sizeof(int))) != cudaSuccess ||
        (e = cudaMalloc(&d.blocks, (size_t)heap_blocks * BRICK_FLOATS * sizeof(float))) != cudaSuccess ||
        (e = cudaMalloc(&d.known, known_words * sizeof(uint32_t))) != cudaSuccess ||
        (e = cudaMalloc(&h->sb_flags, h->sb_words * sizeof(uint32_t))) != cudaSuccess ||
        (e = alloc_lists(h->vis_list, list_cap)) != cudaSuccess ||
        (e = cudaMalloc(&h->tot, sizeof(Totals))) != cudaSuccess ||
        (e = cudaMalloc(&h->scratch_count, sizeof(int))) != cudaSuccess) {
        tsdf_set_error("device allocation failed (%lld slots, %lld blocks = %.2f GB): %s", (long long)S,
                       (long long)heap_blocks, heap_blocks * 6144.0 / 1e9, cudaGetErrorString(e));
        cudaGetLastError();
        hash_free(h);
        return TSDF_ERR_OOM;
    }
    cudaMemsetAsync(h->tot, 0, sizeof(Totals), h->io.compute);
    launch_hash_init(d, sm, h->io.compute);
    h->integrate_grid = sm * std::max(1, integrate_hash_ctas_per_sm() - integrate_reserve_ctas());
    if ((rc = h->counters.init(h->integrate_grid * (INTEGRATE_THREADS / 32)))) {
        hash_free(h);
        return rc;
    }
    e = cudaStreamSynchronize(h->io.compute);
    if (e != cudaSuccess) {
        tsdf_set_error("hash map initialisation failed: %s", cudaGetErrorString(e));
        hash_free(h);
        return TSDF_ERR_CUDA;
    }
    *out = h;
    return TSDF_OK;
}

extern "C" int vh_destroy(tsdf_hash *h) { return hash_free(h); }

// 3x4 camera->world pose from the 4x4 inverse pose (general affine inverse, double precision).
static int invert_affine(const double Ti[16], double P[12])
{
    const double a = Ti[0], b = Ti[1], c = Ti[2], d = Ti[4], e = Ti[5], f = Ti[6], g = Ti[8], hh = Ti[9], i = Ti[10];
    const double det = a * (e * i - f * hh) - b * (d * i - f * g) + c * (d * hh - e * g);
    if (!(fabs(det) > 1e-300) || !isfinite(det)) return -1;
    const double r[9] = {(e * i - f * hh) / det, (c * hh - b * i) / det, (b * f - c * e) / det,
                         (f * g - d * i) / det,  (a * i - c * g) / det,  (c * d - a * f) / det,
                         (d * hh - e * g) / det, (b * g - a * hh) / det, (a * e - b * d) / det};
    for (int row = 0; row < 3; ++row) {
        for (int col = 0; col < 3; ++col) P[4 * row + col] = r[3 * row + col];
        P[4 * row + 3] = -(r[3 * row] * Ti[3] + r[3 * row + 1] * Ti[7] + r[3 * row + 2] * Ti[11]);
    }
    return 0;
}

// Sticky error raised inside a kernel by an earlier frame (heap exhausted / table full).
static int check_device_error(tsdf_hash *h)
{
    Totals tot;
    CUDA_TRY(cudaMemcpy(&tot, h->tot, sizeof(tot), cudaMemcpyDeviceToHost));
    if (tot.error) {
        const int code = -(int)tot.error;
        tsdf_set_error(code == TSDF_ERR_HEAP_EXHAUSTED
                           ? "voxel-block heap exhausted (%lld blocks): blocks were dropped; create the map with more heap_blocks"
                           : "hash table full (%lld slots): keys were dropped; create the map with more table_slots",
                       (long long)(code == TSDF_ERR_HEAP_EXHAUSTED ? h->heap_blocks : h->table_slots));
        return code;
    }
    return TSDF_OK;
}

static int hash_run(tsdf_hash *h, const uint8_t *d_rgb, const double *d_depth, int H, int W, FrameParams &f,
                    CullParams &c, const double Tinv[16])
{
    int rc = h->io.ensure_px(H * W);
    if (rc || (rc = h->io.ensure_tiles(H, W))) return rc;
    const int tiles_x = h->io.pyr_slot[0].tx[0], tiles_y = h->io.pyr_slot[0].ty[0];
    WalkParams w;
    BAD_ARG(invert_affine(Tinv, w.P) != 0, "Tinv is singular");
    w.ox = (double)h->geom.origin[0]; w.oy = (double)h->geom.origin[1]; w.oz = (double)h->geom.origin[2];
    w.inv_vs = 1.0 / h->geom.vs;
    w.step = h->walk_step_voxels * h->geom.vs;
    w.trunc = h->geom.trunc;
    w.fx = f.fx; w.fy = f.fy; w.cx = f.cx; w.cy = f.cy;
    BAD_ARG(f.fx == 0.0 || f.fy == 0.0, "zero focal length");
    // ||R||_2^2 = lambda_max(R^T R) <= 1 + ||R^T R - I||_F
    double ef = 0.0;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double dot = 0.0;
            for (int k = 0; k < 3; ++k) dot += w.P[4 * k + i] * w.P[4 * k + j];
            dot -= (i == j) ? 1.0 : 0.0;
            ef += dot * dot;
        }
    w.pose_scale = sqrt(1.0 + sqrt(ef)) * 1.000001;
    w.W = W; w.H = H; w.tiles_x = tiles_x; w.tiles_y = tiles_y;
    {   // a ray is inside the (dilated) volume for at most its diagonal: bound the samples per tile
        double diag = 0.0;
        for (int a = 0; a < 3; ++a) diag += (double)h->geom.dim[a] * h->geom.dim[a];
        const double max_samples = (sqrt(diag) + 64.0) * h->geom.vs / w.step * 1.5 + 4.0;
        w.chunks = std::max(1, std::min(64, (int)ceil(max_samples / 32.0)));
    }
    w.dim_x = (int)h->geom.dim[0]; w.dim_y = (int)h->geom.dim[1]; w.dim_z = (int)h->geom.dim[2];

    if ((rc = h->io.begin_kernels())) return rc;
    BrickRef *list = h->vis_list[h->io.slot];
    c.pyr = h->io.pyr;
    f.depth = d_depth;
    f.px = h->io.d_px;
    h->unit_w = h->unit_w && f.obs == 1.0;
    f.unit_w = h->unit_w ? 1 : 0;
    // allocation + compaction (stream `pre`; they only insert keys and read table state, so they
    // overlap the previous frame's integrate kernel, which only touches records / vals / the heap)
    cudaStream_t s = h->io.pre;
    c.tag = h->io.next_tag();
    launch_prep_frame(d_rgb, d_depth, h->io.d_px, h->io.fs, h->io.pyr, H, W, c.tag, s);
    CUDA_TRY(cudaGetLastError());
    if ((rc = h->io.release_upload_after_prep())) return rc; // every later kernel reads the prepared pixels
    const double n_sb = (double)h->dev.nsx * h->dev.nsy * h->dev.nsz;
    const bool ray_walk = h->alloc_mode == 1 || (h->alloc_mode == 2 && n_sb > WALK_SB_PER_TILE * tiles_x * tiles_y);
    // the frame's visible-super-block bitmap: written by the walk or by the scan, read by refine and compaction
    CUDA_TRY(cudaMemsetAsync(h->sb_flags, 0, h->sb_words * sizeof(uint32_t), s));
    if (ray_walk) {
        launch_walk_coarse(w, h->dev, h->io.pyr.level0, h->sb_flags, h->io.fs, s);
        CUDA_TRY(cudaGetLastError());
    }
    launch_refine_insert(f, c, h->dev, h->sb_flags, ray_walk, h->io.fs, h->tot, s);
    CUDA_TRY(cudaGetLastError());
    // a key exists at most once per block of the bounded grid: bound the scan by that, not by the table size
    launch_compact_visible(f, c, h->dev, h->sb_flags, (int)std::min<int64_t>(h->table_slots, (int64_t)h->dev.nbx * h->dev.nby * h->dev.nbz),
                           list, h->io.fs, s);
    CUDA_TRY(cudaGetLastError());
    if ((rc = h->io.prepared())) return rc;
    launch_integrate_hash(f, h->geom.params(), h->dev, list, h->io.fs, h->tot, h->counters.view(h->io.slot),
                          h->integrate_grid, h->io.compute);
    CUDA_TRY(cudaGetLastError());
    if ((rc = h->io.end_kernels())) return rc;
    h->frames++;
    return TSDF_OK;
}

extern "C" int vh_integrate(tsdf_hash *h, const uint8_t *rgb, const double *depth, int H, int W, const double K[9],
                            const double Tinv[16], double obs_weight)
{
    BAD_ARG(!h || !rgb || !depth, "NULL argument");
    CUDA_TRY(cudaSetDevice(h->device));
    FrameParams f;
    CullParams c;
    int rc = make_frame_params(&f, &c, h->geom, H, W, K, Tinv, obs_weight);
    if (rc) return rc;
    const uint8_t *d_rgb;
    const double *d_depth;
    if ((rc = h->io.upload(rgb, depth, H * W, &d_rgb, &d_depth))) return rc;
    rc = hash_run(h, d_rgb, d_depth, H, W, f, c, Tinv);
    const int rc2 = h->io.finish_upload(); // the caller's buffers are only borrowed for the duration of the call
    return rc ? rc : rc2;
}

extern "C" int vh_integrate_u16(tsdf_hash *h, const uint8_t *rgb, const uint16_t *depth_raw, int H, int W,
                                const double K[9], const double Tinv[16], double obs_weight, double depth_scale,
                                int invalid_raw)
{
    BAD_ARG(!h || !rgb || !depth_raw, "NULL argument");
    BAD_ARG(!(depth_scale > 0.0), "depth_scale must be positive");
    CUDA_TRY(cudaSetDevice(h->device));
    FrameParams f;
    CullParams c;
    int rc = make_frame_params(&f, &c, h->geom, H, W, K, Tinv, obs_weight);
    if (rc) return rc;
    const uint8_t *d_rgb;
    const double *d_depth;
    if ((rc = h->io.upload_u16(rgb, depth_raw, H * W, depth_scale, invalid_raw, &d_rgb, &d_depth))) return rc;
    rc = hash_run(h, d_rgb, d_depth, H, W, f, c, Tinv);
    const int rc2 = h->io.finish_upload();
    return rc ? rc : rc2;
}

extern "C" int vh_integrate_device(tsdf_hash *h, const uint8_t *d_rgb, const double *d_depth, int H, int W,
                                   const double K[9], const double Tinv[16], double obs_weight)
{
    BAD_ARG(!h || !d_rgb || !d_depth, "NULL argument");
    CUDA_TRY(cudaSetDevice(h->device));
    FrameParams f;
    CullParams c;
    int rc = make_frame_params(&f, &c, h->geom, H, W, K, Tinv, obs_weight);
    if (rc) return rc;
    return hash_run(h, d_rgb, d_depth, H, W, f, c, Tinv);
}

extern "C" int vh_integrate_device_u16(tsdf_hash *h, const uint8_t *d_rgb, const uint16_t *d_depth_raw, int H, int W,
                                       const double K[9], const double Tinv[16], double obs_weight, double depth_scale,
                                       int invalid_raw)
{
    BAD_ARG(!h || !d_rgb || !d_depth_raw, "NULL argument");
    BAD_ARG(!(depth_scale > 0.0), "depth_scale must be positive");
    CUDA_TRY(cudaSetDevice(h->device));
    FrameParams f;
    CullParams c;
    int rc = make_frame_params(&f, &c, h->geom, H, W, K, Tinv, obs_weight);
    if (rc) return rc;
    const double *d_depth;
    if ((rc = h->io.convert_u16(d_depth_raw, H * W, depth_scale, invalid_raw, &d_depth))) return rc;
    return hash_run(h, d_rgb, d_depth, H, W, f, c, Tinv);
}

extern "C" int vh_frame_source(tsdf_hash *h, void *producer_stream)
{
    BAD_ARG(!h, "NULL handle");
    CUDA_TRY(cudaSetDevice(h->device));
    return h->io.adopt(static_cast<cudaStream_t>(producer_stream));
}

static int hash_fill_stats(tsdf_hash *h, const StatsBlock *st, int64_t frames, int64_t stats[12])
{
    stats[0] = h->heap_blocks - st->heap_top;
    stats[1] = st->n_entries;
    stats[2] = h->table_slots;
    stats[3] = h->heap_blocks;
    stats[4] = st->heap_top;
    stats[5] = frames > 0 ? (int64_t)st->updated_last : 0;
    stats[6] = st->updated_total == ~0ull ? -1 : (int64_t)st->updated_total; // -1 in asynchronous snapshots
    stats[7] = (int64_t)st->last_visited;
    stats[8] = (int64_t)st->max_probe;
    stats[9] = st->new_keys;
    stats[10] = frames;
    stats[11] = st->cand_count;
    if (st->error) {
        const int code = -(int)st->error;
        tsdf_set_error(code == TSDF_ERR_HEAP_EXHAUSTED
                           ? "voxel-block heap exhausted (%lld blocks): blocks were dropped; create the map with more heap_blocks"
                           : "hash table full (%lld slots): keys were dropped; create the map with more table_slots",
                       (long long)(code == TSDF_ERR_HEAP_EXHAUSTED ? h->heap_blocks : h->table_slots));
        return code;
    }
    return TSDF_OK;
}

extern "C" int vh_stats(tsdf_hash *h, int64_t stats[12])
{
    BAD_ARG(!h || !stats, "NULL argument");
    CUDA_TRY(cudaSetDevice(h->device));
    const StatsBlock *st = nullptr;
    const int slot = h->frames > 0 ? h->io.slot : 0;
    int rc = h->counters.collect(slot, h->io.fs, h->tot, h->dev.n_entries, h->dev.heap_top, h->io.compute, &st);
    if (rc) return rc;
    return hash_fill_stats(h, st, h->frames, stats);
}

extern "C" int vh_stats_post(tsdf_hash *h, int64_t *ticket)
{
    BAD_ARG(!h || !ticket, "NULL argument");
    CUDA_TRY(cudaSetDevice(h->device));
    const int slot = h->frames > 0 ? h->io.slot : 0;
    int rc = h->io.stat_begin();
    if (rc || (rc = h->counters.post(slot, h->io.fs, h->tot, h->dev.n_entries, h->dev.heap_top, h->io.stat, h->frames, ticket)))
        return rc;
    return h->io.stat_end();
}

extern "C" int vh_stats_wait(tsdf_hash *h, int64_t ticket, int64_t stats[12])
{
    BAD_ARG(!h || !stats, "NULL argument");
    CUDA_TRY(cudaSetDevice(h->device));
    const StatsBlock *st = nullptr;
    int64_t frames = 0;
    int rc = h->counters.wait(ticket, &st, &frames);
    if (rc) return rc;
    return hash_fill_stats(h, st, frames, stats);
}

extern "C" int vh_num_blocks(tsdf_hash *h, int64_t *n)
{
    BAD_ARG(!h || !n, "NULL argument");
    int64_t s[12];
    const int rc = vh_stats(h, s);
    *n = s[0];
    return rc;
}

extern "C" int vh_sync(tsdf_hash *h)
{
    BAD_ARG(!h, "NULL handle");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaStreamSynchronize(h->io.stat));
    CUDA_TRY(cudaStreamSynchronize(h->io.copy));
    CUDA_TRY(cudaStreamSynchronize(h->io.pre));
    CUDA_TRY(cudaStreamSynchronize(h->io.compute));
    return check_device_error(h);
}

extern "C" int vh_last_kernel_ms(tsdf_hash *h, float *ms)
{
    BAD_ARG(!h, "NULL handle");
    CUDA_TRY(cudaSetDevice(h->device));
    return h->io.last_ms(ms);
}

extern "C" int vh_stream(tsdf_hash *h, void **stream)
{
    BAD_ARG(!h || !stream, "NULL argument");
    *stream = (void *)h->io.compute;
    return TSDF_OK;
}

extern "C" int vh_to_dense_device(tsdf_hash *h, float *d_tsdf, float *d_weight, float *d_color)
{
    BAD_ARG(!h, "NULL handle");
    CUDA_TRY(cudaSetDevice(h->device));
    launch_scatter_dense(h->dev, (int)h->geom.dim[0], (int)h->geom.dim[1], (int)h->geom.dim[2], d_tsdf, d_weight,
                         d_color, h->sm_count, h->io.compute);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(h->io.compute));
    return check_device_error(h);
}

extern "C" int vh_to_dense(tsdf_hash *h, float *tsdf, float *weight, float *color)
{
    BAD_ARG(!h, "NULL handle");
    CUDA_TRY(cudaSetDevice(h->device));
    const size_t n = (size_t)h->geom.dim[0] * h->geom.dim[1] * h->geom.dim[2];
    float *scratch = nullptr;
    CUDA_TRY(cudaMalloc(&scratch, n * sizeof(float)));
    float *host[3] = {tsdf, weight, color};
    for (int p = 0; p < 3; ++p) {
        if (!host[p]) continue;
        launch_scatter_dense(h->dev, (int)h->geom.dim[0], (int)h->geom.dim[1], (int)h->geom.dim[2],
                             p == 0 ? scratch : nullptr, p == 1 ? scratch : nullptr, p == 2 ? scratch : nullptr,
                             h->sm_count, h->io.compute);
        cudaError_t e = cudaMemcpyAsync(host[p], scratch, n * sizeof(float), cudaMemcpyDeviceToHost, h->io.compute);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->io.compute);
        if (e != cudaSuccess) {
            cudaFree(scratch);
            tsdf_set_error("to_dense failed: %s", cudaGetErrorString(e));
            return TSDF_ERR_CUDA;
        }
    }
    cudaFree(scratch);
    return check_device_error(h);
}

extern "C" int vh_export_blocks(tsdf_hash *h, int64_t capacity, int64_t *n, int32_t *keys, float *payload)
{
    BAD_ARG(!h || !n, "NULL argument");
    BAD_ARG(capacity < 0, "negative capacity");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t s = h->io.compute;
    const bool want = keys && payload && capacity > 0;
    int32_t *d_keys = nullptr;
    float *d_payload = nullptr;
    if (want) {
        CUDA_TRY(cudaMalloc(&d_keys, (size_t)capacity * 3 * sizeof(int32_t)));
        cudaError_t e = cudaMalloc(&d_payload, (size_t)capacity * BRICK_FLOATS * sizeof(float));
        if (e != cudaSuccess) {
            cudaFree(d_keys);
            cudaGetLastError();
            tsdf_set_error("export scratch allocation failed: %s", cudaGetErrorString(e));
            return TSDF_ERR_OOM;
        }
    }
    cudaMemsetAsync(h->scratch_count, 0, sizeof(int), s);
    launch_export_blocks(h->dev, want ? (int)capacity : 0, h->scratch_count, d_keys, d_payload, h->sm_count, s);
    int count = 0;
    cudaError_t e = cudaMemcpyAsync(&count, h->scratch_count, sizeof(int), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e == cudaSuccess && want) {
        const size_t m = (size_t)std::min<int64_t>(count, capacity);
        e = cudaMemcpy(keys, d_keys, m * 3 * sizeof(int32_t), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(payload, d_payload, m * BRICK_FLOATS * sizeof(float), cudaMemcpyDeviceToHost);
    }
    if (d_keys) cudaFree(d_keys);
    if (d_payload) cudaFree(d_payload);
    if (e != cudaSuccess) {
        tsdf_set_error("export failed: %s", cudaGetErrorString(e));
        return TSDF_ERR_CUDA;
    }
    *n = count;
    return check_device_error(h);
}

// Block-list read-out for the multi-GPU gather: this rank's blocks into DEVICE buffers (capacity blocks; pass
// capacity 0 to count), and the inverse -- any block list scattered into dense DEVICE volumes.
extern "C" int vh_export_blocks_device(tsdf_hash *h, int64_t capacity, int64_t *n, int32_t *d_keys, float *d_payload)
{
    BAD_ARG(!h || !n, "NULL argument");
    BAD_ARG(capacity < 0 || capacity > ((int64_t)1 << 30), "bad capacity");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t s = h->io.compute;
    const bool want = d_keys && d_payload && capacity > 0;
    CUDA_TRY(cudaMemsetAsync(h->scratch_count, 0, sizeof(int), s));
    launch_export_blocks(h->dev, want ? (int)capacity : 0, h->scratch_count, d_keys, d_payload, h->sm_count, s);
    CUDA_TRY(cudaGetLastError());
    int count = 0;
    CUDA_TRY(cudaMemcpyAsync(&count, h->scratch_count, sizeof(int), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    *n = count;
    return check_device_error(h);
}

extern "C" int vh_scatter_blocks_device(tsdf_hash *h, const int32_t *d_keys, const float *d_payload, int64_t n, int fill,
                                        float *d_tsdf, float *d_weight, float *d_color)
{
    BAD_ARG(!h || ((!d_keys || !d_payload) && n > 0), "NULL argument");
    BAD_ARG(n < 0 || n > ((int64_t)1 << 30), "bad block count");
    CUDA_TRY(cudaSetDevice(h->device));
    launch_scatter_block_list(d_keys, d_payload, (int)n, (int)h->geom.dim[0], (int)h->geom.dim[1], (int)h->geom.dim[2],
                              d_tsdf, d_weight, d_color, fill != 0, h->sm_count, h->io.compute);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(h->io.compute));
    return TSDF_OK;
}

extern "C" int vh_remove_blocks(tsdf_hash *h, const int32_t *keys, int64_t n, int64_t *removed)
{
    BAD_ARG(!h || (!keys && n > 0), "NULL argument");
    BAD_ARG(n < 0 || n > ((int64_t)1 << 30), "bad key count");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t s = h->io.compute;
    int count = 0;
    if (n > 0) {
        int32_t *d_keys = nullptr;
        CUDA_TRY(cudaMalloc(&d_keys, (size_t)n * 3 * sizeof(int32_t)));
        cudaError_t e = cudaMemcpyAsync(d_keys, keys, (size_t)n * 3 * sizeof(int32_t), cudaMemcpyHostToDevice, s);
        cudaMemsetAsync(h->scratch_count, 0, sizeof(int), s);
        launch_remove_blocks(h->dev, d_keys, (int)n, h->scratch_count, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(&count, h->scratch_count, sizeof(int), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        cudaFree(d_keys);
        if (e != cudaSuccess) {
            tsdf_set_error("remove failed: %s", cudaGetErrorString(e));
            return TSDF_ERR_CUDA;
        }
    }
    if (removed) *removed = count;
    return TSDF_OK;
}

// Per-voxel view of the block map: HashTable.get_hash_entry / get_voxel / remove on the table that
// integrate() fills (reference hash_fusion.py:285-320, 322-393).  mode 0 = look up, 1 = remove.
extern "C" int vh_voxel_query(tsdf_hash *h, const int32_t *coords, int64_t n, int mode, float *out, uint8_t *found)
{
    BAD_ARG(!h || (!coords && n > 0), "NULL argument");
    BAD_ARG(n < 0 || n > ((int64_t)1 << 28), "bad query count");
    BAD_ARG(mode != 0 && mode != 1, "mode %d (0 look up, 1 remove)", mode);
    CUDA_TRY(cudaSetDevice(h->device));
    if (!n) return TSDF_OK;
    cudaStream_t s = h->io.compute;
    int32_t *d_coords = nullptr;
    float *d_out = nullptr;
    uint8_t *d_found = nullptr;
    cudaError_t e = cudaMalloc(&d_coords, (size_t)n * 3 * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc(&d_out, (size_t)n * 3 * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&d_found, (size_t)n);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_coords, coords, (size_t)n * 3 * sizeof(int32_t), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) {
        launch_voxel_query(h->dev, d_coords, (int)n, mode, d_out, d_found, s);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess && out) e = cudaMemcpyAsync(out, d_out, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess && found) e = cudaMemcpyAsync(found, d_found, (size_t)n, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (d_coords) cudaFree(d_coords);
    if (d_out) cudaFree(d_out);
    if (d_found) cudaFree(d_found);
    if (e != cudaSuccess) {
        tsdf_set_error("voxel query failed: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return e == cudaErrorMemoryAllocation ? TSDF_ERR_OOM : TSDF_ERR_CUDA;
    }
    return TSDF_OK;
}

// Number of voxels that have an entry in the reference's sense (weight > 0): count_num_hash_entries.
extern "C" int vh_count_voxels(tsdf_hash *h, int64_t *count)
{
    BAD_ARG(!h || !count, "NULL argument");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t s = h->io.compute;
    unsigned long long *d_total = nullptr, total = 0;
    CUDA_TRY(cudaMalloc(&d_total, sizeof(unsigned long long)));
    cudaError_t e = cudaMemsetAsync(d_total, 0, sizeof(unsigned long long), s);
    if (e == cudaSuccess) {
        launch_count_voxels(h->dev, d_total, h->sm_count, s);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(&total, d_total, sizeof(total), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(d_total);
    if (e != cudaSuccess) {
        tsdf_set_error("voxel count failed: %s", cudaGetErrorString(e));
        return TSDF_ERR_CUDA;
    }
    *count = (int64_t)total;
    return TSDF_OK;
}
