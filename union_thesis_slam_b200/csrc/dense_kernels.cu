// dense_kernels.cu -- sm_100a kernels of the dense brick grid (K1 of SURVEY.md section 2a).
//
// Data layout in HBM: the grid is stored as 8x8x8-voxel bricks, one 6 KB record per brick:
//   record = [512 tsdf f32 | 512 weight f32 | 512 colour u32 (B<<16|G<<8|R, the integer value of
//   the reference's float encoding)], voxel order (x*8 + y)*8 + z.
// A z-run of 8 voxels is one 32 B sector per plane; a warp's float4 accesses cover 512 B
// contiguous, so every DRAM transaction is a full sector and bricks the frame cannot touch are
// never read.  Brick (bx,by,bz) of this rank lives at index ((bx/world)*nby + by)*nbz + bz.
//
// Per frame:  prep_frame (fold colour, max depth)  ->  cull_bricks (frustum + depth range,
// warp-aggregated compaction)  ->  integrate_bricks (persistent CTAs over the visible list).
#include <stdlib.h>
#include <string.h>

#include "launch.h"
#include "integrate.cuh"

// ---------------------------------------------------------------------------------------------
__global__ void k_fill_bricks(float4 *bricks, size_t n_float4)
{
    // record pattern: first 128 float4 of every 384 are tsdf = 1, the rest 0
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_float4; i += stride) {
        const float v = (i % 384) < 128 ? 1.0f : 0.0f;
        bricks[i] = make_float4(v, v, v, v);
    }
}

// grid_fusion.py:228-232 colour fold (as an exact integer), max valid depth of the frame and the
// per-tile max-depth pyramid.  One warp per 8x8-pixel tile, two horizontally adjacent pixels per lane.
__global__ void k_prep_frame(const uint8_t *__restrict__ rgb, const double *__restrict__ depth,
                             FramePixel *__restrict__ px, FrameScratch *fs, TilePyramid pyr, int H, int W, unsigned tag)
{
    const int tile = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (tile >= pyr.tx[0] * pyr.ty[0]) return;
    if (tile == 0 && lane == 0) {
        // sentinel pixel behind the image: what a voxel that projects nowhere reads
        *reinterpret_cast<uint4 *>(px + (size_t)H * W) = make_uint4(0u, 0xFFF00000u, 0u, 0u);
        // the frame's counters start at zero (nothing in this kernel uses them; the maxima are tagged instead)
        fs->vis_count = 0; fs->updated = 0ull; fs->cand_count = 0; fs->new_keys = 0; fs->error = 0;
        fs->vis_extra[0] = fs->vis_extra[1] = fs->vis_extra[2] = 0;
    }
    const int tx = tile % pyr.tx[0], ty = tile / pyr.tx[0];
    const int v = ty * 8 + (lane >> 2), u = tx * 8 + (lane & 3) * 2;
    float m = 0.0f;
    if (v < H) {
#pragma unroll
        for (int k = 0; k < 2; ++k)
            if (u + k < W) {
                const int p = v * W + u + k;
                const uint32_t r = __ldg(rgb + 3 * p), g = __ldg(rgb + 3 * p + 1), b = __ldg(rgb + 3 * p + 2);
                const double d = __ldg(depth + p);
                const bool valid = d > 0.0; // grid_fusion.py:284; false for 0, negatives and NaN
                // one 16-byte store: {depth or -inf, folded colour, 0}
                const double ds = valid ? d : __longlong_as_double(0xFFF0000000000000ll);
                *reinterpret_cast<uint4 *>(px + p) = make_uint4((unsigned)__double2loint(ds), (unsigned)__double2hiint(ds),
                                                                (b << 16) | (g << 8) | r, 0u);
                if (valid) m = fmaxf(m, __double2float_ru(d));
            }
    }
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) {
        pyr.level0[tile] = m;
        if (m > 0.0f) {
            const unsigned long long v = ((unsigned long long)tag << 32) | __float_as_uint(m);
            atomicMax(pyr.up[0] + (ty >> 1) * pyr.tx[1] + (tx >> 1), v);
            atomicMax(pyr.up[1] + (ty >> 2) * pyr.tx[2] + (tx >> 2), v);
            atomicMax(pyr.up[2] + (ty >> 3) * pyr.tx[3] + (tx >> 3), v);
            atomicMax(&fs->max_depth_tagged, v);
        }
    }
}

// Frame ingest (grid_demo1.py:81-83): depth = raw.astype(float) / depth_scale, raw == invalid_raw -> 0.
// IEEE division, so the result is bit-identical to NumPy's.
__global__ void k_depth_u16(const uint16_t *__restrict__ raw, double *__restrict__ depth, int npix, double scale,
                            int invalid_raw)
{
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += gridDim.x * blockDim.x) {
        const int r = raw[p];
        depth[p] = (r == invalid_raw) ? 0.0 : __ddiv_rn((double)r, scale);
    }
}

// One thread per local brick; visible bricks are appended with one atomic per warp and list.
// The bricks are dealt into `classes` lists by (local bx + by + bz) % classes (a 3-D checkerboard
// for two classes).  The lists never share a brick, so their integrate kernels are independent and
// run on separate streams: the next list's CTAs fill the SM slots that a kernel's tail leaves idle,
// and the next frame's first kernel does the same for the last one (capi_grid.cu).
// Class k's list starts at lists + k * class_cap; its length is counted in fs (dense_count()).
__device__ __forceinline__ int *dense_count(FrameScratch *fs, int k) { return k ? &fs->vis_extra[k - 1] : &fs->vis_count; }

__global__ void k_cull_bricks(FrameParams f, CullParams c, int nlbx, int nby, int nbz, int rank, int world,
                              int classes, BrickRef *__restrict__ lists, size_t class_cap, FrameScratch *fs)
{
    const int n = nlbx * nby * nbz;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool vis = false;
    int bx = 0, by = 0, bz = 0, cls = 0;
    if (i < n) {
        bz = i % nbz;
        by = (i / nbz) % nby;
        const int lbx = i / (nbz * nby);
        bx = sheet_of(lbx, rank, world);
        cls = (lbx + by + bz) % classes;
        c.max_depth = (double)frame_max_depth(fs, c.tag);
        vis = c.max_depth > 0.0 && brick_visible(f, c, bx, by, bz);
    }
    if (!__ballot_sync(0xffffffffu, vis)) return;
    const int lane = threadIdx.x & 31;
    for (int k = 0; k < classes; ++k) {
        const unsigned m = __ballot_sync(0xffffffffu, vis && cls == k);
        if (!m) continue;
        int base = 0;
        if (lane == 0) base = atomicAdd(dense_count(fs, k), __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (vis && cls == k) {
            BrickRef r;
            r.bx = bx; r.by = by; r.bz = bz; r.storage = (uint32_t)i;
            lists[(size_t)k * class_cap + base + __popc(m & ((1u << lane) - 1u))] = r;
        }
    }
}

// Batch culling: the same test against every frame of a batch; a brick is listed once, with the mask
// of the frames that can touch it in the top byte of bz (the integrate kernel skips the others).
// Per-frame max depth comes from each frame's own FrameScratch; the list counters live in frame 0's.
struct CullBatch {
    FrameParams f[DENSE_MAX_BATCH];
    CullParams c[DENSE_MAX_BATCH];
    const FrameScratch *fs[DENSE_MAX_BATCH];
    int n;
};

__global__ void k_cull_bricks_batch(CullBatch cb, int nlbx, int nby, int nbz, int rank, int world, int classes,
                                    BrickRef *__restrict__ lists, size_t class_cap, FrameScratch *fs0)
{
    const int n = nlbx * nby * nbz;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned mask = 0;
    int bx = 0, by = 0, bz = 0, cls = 0;
    if (i < n) {
        bz = i % nbz;
        by = (i / nbz) % nby;
        const int lbx = i / (nbz * nby);
        bx = sheet_of(lbx, rank, world);
        cls = (lbx + by + bz) % classes;
#pragma unroll 1
        for (int k = 0; k < cb.n; ++k) {
            CullParams c = cb.c[k];
            c.max_depth = (double)frame_max_depth(cb.fs[k], c.tag);
            if (c.max_depth > 0.0 && brick_visible(cb.f[k], c, bx, by, bz)) mask |= 1u << k;
        }
    }
    const bool vis = mask != 0;
    if (!__ballot_sync(0xffffffffu, vis)) return;
    const int lane = threadIdx.x & 31;
    for (int k = 0; k < classes; ++k) {
        const unsigned m = __ballot_sync(0xffffffffu, vis && cls == k);
        if (!m) continue;
        int base = 0;
        if (lane == 0) base = atomicAdd(dense_count(fs0, k), __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (vis && cls == k) {
            BrickRef r;
            r.bx = bx; r.by = by; r.bz = bz | (int)(mask << 24); r.storage = (uint32_t)i;
            lists[(size_t)k * class_cap + base + __popc(m & ((1u << lane) - 1u))] = r;
        }
    }
}

// Persistent CTAs of two warps; one brick per iteration, thread t owns the z-column
// (x = t>>3, y = t&7): two quads of four voxels, i.e. two float4 (= one 32 B sector) per plane.
// One launch per class of bricks (see k_cull_bricks); `count` is that class's list length.
//
// k_integrate_dense<UNIT_W, K>: K frames are applied to a brick during ONE visit (temporal blocking;
// K = 1 is the plain per-frame kernel).  A list entry carries, in the top byte of bz, the mask of
// the batch's frames that can touch the brick; every voxel applies them in frame order, so the
// result is bit-identical to K single-frame launches (the rule has no cross-voxel term).  The
// record quads are staged in thread-private shared-memory slots by cp.async one quad ahead of use
// (integrate.cuh): the record's DRAM latency hides behind the previous quad's geometry, and between
// the frames of a visit the record lives in the slot instead of going back to HBM.
//
// k_integrate_dense_v1: the round-1 kernel (record requested with the pixels, straight from global
// memory), kept behind TSDF_DENSE_V1=1 for A/B measurements.
#ifndef DENSE_UNROLL_FRAMES
#define DENSE_UNROLL_FRAMES 0
#endif
template <int K>
struct FrameBatch {
    FrameParams f[K];
};

#ifndef DENSE_BATCH_MIN_CTAS
#define DENSE_BATCH_MIN_CTAS 12 // K > 1: 80 registers (the frame loop carries more state; 64 spills ~200 B)
#endif
#ifndef DENSE_FAST_START
#define DENSE_FAST_START 1
#endif
#ifndef DENSE_SM_CONTIGUOUS
#define DENSE_SM_CONTIGUOUS 1
#endif

// Which list entry a persistent CTA takes in round r.  CTAs are dealt to SMs round-robin
// (CTA i -> SM i % n_sm for the resident wave), so with DENSE_SM_CONTIGUOUS the CTAs that share an SM
// -- and its L1 -- take CONSECUTIVE list entries: neighbouring bricks, whose pixel footprints and
// lattice rows overlap.  Any placement is correct; the mapping is a permutation of each round.
__device__ __forceinline__ int entry_of(int round, int n_sm)
{
#if DENSE_SM_CONTIGUOUS
    const int per_sm = gridDim.x / n_sm; // the launcher sizes the grid as a multiple of the SM count
    const int sm = blockIdx.x % n_sm, j = blockIdx.x / n_sm;
    return round * (int)gridDim.x + sm * per_sm + j;
#else
    return round * (int)gridDim.x + blockIdx.x;
#endif
}

template <bool UNIT_W, int K>
__global__ void __launch_bounds__(INTEGRATE_THREADS, K == 1 ? INTEGRATE_MIN_CTAS : DENSE_BATCH_MIN_CTAS)
k_integrate_dense(FrameBatch<K> fb, VolumeParams v, const BrickRef *__restrict__ list, const int *__restrict__ count,
                  float *__restrict__ bricks, UpdateCounters uc, int n_sm)
{
    // one quad record per thread: [plane][thread] float4, filled by the thread's own cp.async
    __shared__ __align__(16) float4 s_rec[3][INTEGRATE_THREADS];
    __shared__ unsigned s_updates[INTEGRATE_THREADS];
    const int n = *count;
    const int t = threadIdx.x;
    const int lx = t >> 3, ly = t & 7;
    s_updates[t] = 0; // only this thread touches its counter: no barrier needed
    float4 *slot = &s_rec[0][t];
    int round = 0;
    int b = entry_of(0, n_sm);
    int4 br = make_int4(0, 0, 0, 0);
    bool inb_next = false;
    float *rec_next = nullptr;
    if (b < n) {
        br = __ldg(reinterpret_cast<const int4 *>(list) + b);
        inb_next = br.x * BRICK + lx < v.dim_x && br.y * BRICK + ly < v.dim_y;
        rec_next = bricks + (size_t)(uint32_t)br.w * BRICK_FLOATS + 8 * t;
        if (inb_next) prefetch_quad(slot, rec_next);
    }
    prefetch_commit();
    while (b < n) {
        const int4 cur = br;
        const bool inb = inb_next;
        float *rec = rec_next;
        const int b_next = entry_of(++round, n_sm);
        inb_next = false;
        if (b_next < n) {
            br = __ldg(reinterpret_cast<const int4 *>(list) + b_next);
            inb_next = br.x * BRICK + lx < v.dim_x && br.y * BRICK + ly < v.dim_y;
            rec_next = bricks + (size_t)(uint32_t)br.w * BRICK_FLOATS + 8 * t;
        }
        b = b_next;
        const int gx = cur.x * BRICK + lx, gy = cur.y * BRICK + ly;
        const int gz0 = (K == 1 ? cur.z : (cur.z & 0xFFFFFF)) * BRICK;
        const unsigned mask = K == 1 ? 1u : (unsigned)cur.z >> 24; // CTA-uniform
        double wx = 0.0, wy = 0.0;
        if (inb) {
            wx = __ldg(v.lat_x + gx);
            wy = __ldg(v.lat_y + gy);
        }
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            const int gz = gz0 + 4 * h;
            const int nvalid = v.dim_z - gz;
            unsigned nupd = 0;
            if (inb && nvalid > 0) {
                const double2 wz01 = __ldg(reinterpret_cast<const double2 *>(v.lat_z + gz));
                const double2 wz23 = __ldg(reinterpret_cast<const double2 *>(v.lat_z + gz) + 1);
                const double wz[4] = {wz01.x, wz01.y, wz23.x, wz23.y};
                unsigned flags = 0;
#if DENSE_UNROLL_FRAMES
#pragma unroll
#else
#pragma unroll 1
#endif
                for (int k = 0; k < K; ++k) {
                    if (K > 1 && !((mask >> k) & 1u)) continue;
                    double cXs, cYs, cZ;
                    column_sums(fb.f[k], wx, wy, cXs, cYs, cZ);
                    if (nvalid >= 4)
                        nupd += integrate_quad_slot<UNIT_W, true, K == 1>(fb.f[k], wx, wy, cXs, cYs, cZ, wz, 4, slot,
                                                                           rec + 4 * h, flags);
                    else
                        nupd += integrate_quad_slot<UNIT_W, false, K == 1>(fb.f[k], wx, wy, cXs, cYs, cZ, wz, nvalid, slot,
                                                                            rec + 4 * h, flags);
                }
                if (K > 1) flush_quad(slot, rec + 4 * h, flags);
            }
            // Refill the slot with the quad that comes next (the brick's second quad, or the first one of
            // this CTA's next brick): it lands while that quad's geometry and pixel loads run.  The slot's
            // last reads (LDS) fed the arithmetic / stores above, so they have completed; a copy that was
            // never waited for (the quad left early) must land before the next one is issued.
            prefetch_wait<0>();
            if (h == 0) {
                if (inb) prefetch_quad(slot, rec + 4);
            } else if (inb_next) {
                prefetch_quad(slot, rec_next);
            }
            prefetch_commit();
            if (nupd) atomicAdd(&s_updates[t], nupd); // no return value used: nothing waits for it
        }
    }
    prefetch_wait<0>();
    unsigned my_updates = s_updates[t];
    for (int o = 16; o; o >>= 1) my_updates += __shfl_xor_sync(0xffffffffu, my_updates, o);
    if ((t & 31) == 0) { // one private slot per warp: no atomics
        const int w = blockIdx.x * (INTEGRATE_THREADS / 32) + (t >> 5);
        uc.frame[w] = my_updates;
        uc.total[w] += my_updates;
    }
}

// (Quad layers instead of whole bricks as the work unit -- for every brick, or only for the bricks of the last,
// partial round of a launch -- were measured at 1 / 2 / 4 / 8 ranks and changed nothing or lost: the second
// class's launch already fills a launch's tail.  profiles/r2_dense_ab.md.)
template <bool UNIT_W>
__global__ void __launch_bounds__(INTEGRATE_THREADS, INTEGRATE_MIN_CTAS)
k_integrate_dense_v1(FrameParams f, VolumeParams v, const BrickRef *__restrict__ list, const int *__restrict__ count,
                     float *__restrict__ bricks, UpdateCounters uc, int n_sm)
{
    // per-thread update counter in shared memory: a fire-and-forget add per column instead of a register
    // that the 64-register budget would spill (its reload sat on the long scoreboard)
    __shared__ unsigned s_updates[INTEGRATE_THREADS];
    const int n = *count;
    const int t = threadIdx.x;
    const int lx = t >> 3, ly = t & 7;
    s_updates[t] = 0; // only this thread touches its counter: no barrier needed
    int round = 0;
    int b = entry_of(0, n_sm);
#if DENSE_FAST_START
    // The first entry is requested together with the count (the list is padded by a grid's worth of entries, so
    // the read is in bounds whatever the count turns out to be), and the first brick's record is on its way to
    // L2 while the lattice -> pixel chain of its first quad runs: a persistent CTA of a small shard lives for
    // only ~6 bricks, so its start-up chain of dependent loads is a visible part of the launch.
    int4 br_next = __ldg(reinterpret_cast<const int4 *>(list) + b);
    if (b < n) {
        const float *r0 = bricks + (size_t)(uint32_t)br_next.w * BRICK_FLOATS + 8 * t;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(r0));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(r0 + BRICK_VOX));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(r0 + 2 * BRICK_VOX));
    }
#else
    int4 br_next = make_int4(0, 0, 0, 0);
    if (b < n) br_next = __ldg(reinterpret_cast<const int4 *>(list) + b);
#endif
    while (b < n) {
        const int4 br = br_next;
        b = entry_of(++round, n_sm);
        if (b < n) br_next = __ldg(reinterpret_cast<const int4 *>(list) + b);
        const int gx = br.x * BRICK + lx, gy = br.y * BRICK + ly;
        if (gx >= v.dim_x || gy >= v.dim_y) continue;
        const unsigned nupd = integrate_column8<UNIT_W, false>(f, v, gx, gy, br.z * BRICK,
                                                                bricks + (size_t)(uint32_t)br.w * BRICK_FLOATS + 8 * t);
        if (nupd) atomicAdd(&s_updates[t], nupd);
    }
    unsigned my_updates = s_updates[t];
    for (int o = 16; o; o >>= 1) my_updates += __shfl_xor_sync(0xffffffffu, my_updates, o);
    if ((t & 31) == 0) { // one private slot per warp: no atomics
        const int w = blockIdx.x * (INTEGRATE_THREADS / 32) + (t >> 5);
        uc.frame[w] = my_updates;
        uc.total[w] += my_updates;
    }
}

// Stats read-out: sums the per-warp update counters and gathers the scalar counters into one block.
__global__ void k_collect_stats(UpdateCounters uc, int n_warps, const FrameScratch *fs, const Totals *tot,
                                const int *n_entries, const int *heap_top, StatsBlock *out, int with_total)
{
    __shared__ unsigned long long s_a[8], s_b[8];
    unsigned long long a = 0, b = 0;
    for (int i = threadIdx.x; i < n_warps; i += blockDim.x) {
        a += uc.frame[i];
        if (with_total) b += uc.total[i]; // not for snapshots taken while later frames run (they add to total[])
    }
    for (int o = 16; o; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if ((threadIdx.x & 31) == 0) { s_a[threadIdx.x >> 5] = a; s_b[threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        a = b = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a += s_a[i]; b += s_b[i]; }
        StatsBlock r;
        r.updated_last = a; r.updated_total = with_total ? b : ~0ull;
        r.last_visited = tot->last_visited; r.max_probe = tot->max_probe; r.error = tot->error;
        r.cand_count = fs ? fs->cand_count : 0; r.new_keys = fs ? fs->new_keys : 0;
        r.vis_count = fs ? fs->vis_count + fs->vis_extra[0] + fs->vis_extra[1] + fs->vis_extra[2] : 0;
        r.n_entries = n_entries ? *n_entries : 0; r.heap_top = heap_top ? *heap_top : 0; r.pad = 0;
        *out = r;
    }
}

// Read-out: bricks -> C-order [n_x][Y][Z] planes (grid_fusion.py:316-320 layout).
__global__ void k_unbrick(const float *__restrict__ bricks, int n_x, int dim_y, int dim_z, int nby, int nbz,
                          float *__restrict__ tsdf, float *__restrict__ wgt, float *__restrict__ col)
{
    const size_t total = (size_t)n_x * dim_y * dim_z;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += stride) {
        const int k = (int)(o % dim_z);
        const int j = (int)((o / dim_z) % dim_y);
        const int i = (int)(o / ((size_t)dim_z * dim_y));
        const size_t brick = ((size_t)(i >> 3) * nby + (j >> 3)) * nbz + (k >> 3);
        const float *rec = bricks + brick * BRICK_FLOATS + ((i & 7) * 64 + (j & 7) * 8 + (k & 7));
        if (tsdf) tsdf[o] = __ldg(rec);
        if (wgt) wgt[o] = __ldg(rec + BRICK_VOX);
        if (col) col[o] = (float)__float_as_uint(__ldg(rec + 2 * BRICK_VOX)); // packed B<<16|G<<8|R -> f32 encoding
    }
}

// ---------------------------------------------------------------------------------------------
// host-side launchers (declared in launch.h)
void launch_fill_bricks(float *bricks, size_t nbricks, int sm_count, cudaStream_t s)
{
    const size_t n4 = nbricks * (BRICK_FLOATS / 4);
    if (!n4) return;
    k_fill_bricks<<<sm_count * 8, 256, 0, s>>>(reinterpret_cast<float4 *>(bricks), n4);
}

void launch_prep_frame(const uint8_t *rgb, const double *depth, FramePixel *px, FrameScratch *fs,
                       const TilePyramid &pyr, int H, int W, unsigned tag, cudaStream_t s)
{
    const int warps = pyr.tx[0] * pyr.ty[0];
    k_prep_frame<<<(warps * 32 + 127) / 128, 128, 0, s>>>(rgb, depth, px, fs, pyr, H, W, tag);
}

void launch_depth_u16(const uint16_t *raw, double *depth, int npix, double scale, int invalid_raw, cudaStream_t s)
{
    k_depth_u16<<<(npix + 255) / 256, 256, 0, s>>>(raw, depth, npix, scale, invalid_raw);
}

void launch_cull_bricks(const FrameParams &f, const CullParams &c, int nlbx, int nby, int nbz, int rank, int world,
                        int classes, BrickRef *lists, size_t class_cap, FrameScratch *fs, cudaStream_t s)
{
    const long long n = (long long)nlbx * nby * nbz;
    if (!n) return;
    k_cull_bricks<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(f, c, nlbx, nby, nbz, rank, world, classes, lists,
                                                              class_cap, fs);
}

void launch_cull_bricks_batch(const FrameParams *f, const CullParams *c, FrameScratch *const *fs, int k, int nlbx, int nby,
                              int nbz, int rank, int world, int classes, BrickRef *lists, size_t class_cap, cudaStream_t s)
{
    const long long n = (long long)nlbx * nby * nbz;
    if (!n) return;
    CullBatch cb;
    for (int i = 0; i < DENSE_MAX_BATCH; ++i) {
        cb.f[i] = f[i < k ? i : k - 1];
        cb.c[i] = c[i < k ? i : k - 1];
        cb.fs[i] = fs[i < k ? i : k - 1];
    }
    cb.n = k;
    k_cull_bricks_batch<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(cb, nlbx, nby, nbz, rank, world, classes, lists, class_cap,
                                                                    fs[0]);
}

#ifndef INTEGRATE_RESERVE_CTAS
#define INTEGRATE_RESERVE_CTAS 2
#endif
int integrate_reserve_ctas() { return INTEGRATE_RESERVE_CTAS; }

void launch_collect_stats(UpdateCounters uc, int n_warps, const FrameScratch *fs, const Totals *tot, const int *n_entries,
                          const int *heap_top, StatsBlock *out, cudaStream_t s, bool with_total)
{
    k_collect_stats<<<1, 256, 0, s>>>(uc, n_warps, fs, tot, n_entries, heap_top, out, with_total ? 1 : 0);
}

int integrate_dense_ctas_per_sm()
{
    int n = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_integrate_dense_v1<true>, INTEGRATE_THREADS, 0);
    return n > 0 ? n : 1;
}

// Single-frame launches use k_integrate_dense_v1 (measured faster at K = 1, profiles/r2_dense_ab.md);
// TSDF_DENSE_V2=1 routes them through the staged kernel instead (A/B measurements).
static bool dense_v1()
{
    static const bool v2 = getenv("TSDF_DENSE_V2") && atoi(getenv("TSDF_DENSE_V2")) != 0;
    return !v2;
}

// SM count for entry_of(): the mapping needs the grid to be a multiple of it (1 = plain order);
// TSDF_DENSE_ORDER=plain switches the SM-contiguous order off (A/B measurements).
static int sm_count_for(int grid)
{
    static int n_sm = 0;
    static bool plain = false;
    if (!n_sm) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        const char *e = getenv("TSDF_DENSE_ORDER");
        plain = e && !strcmp(e, "plain");
    }
    return (!plain && n_sm > 0 && grid % n_sm == 0) ? n_sm : 1;
}

template <int K>
static void launch_batch(const FrameParams *f, const VolumeParams &v, const BrickRef *list, const int *count, float *bricks,
                         UpdateCounters uc, int grid, cudaStream_t s, bool unit_w)
{
    FrameBatch<K> fb;
    for (int k = 0; k < K; ++k) fb.f[k] = f[k];
    const int sms = sm_count_for(grid);
    if (unit_w)
        k_integrate_dense<true, K><<<grid, INTEGRATE_THREADS, 0, s>>>(fb, v, list, count, bricks, uc, sms);
    else
        k_integrate_dense<false, K><<<grid, INTEGRATE_THREADS, 0, s>>>(fb, v, list, count, bricks, uc, sms);
}

void launch_integrate_dense(const FrameParams &f, const VolumeParams &v, const BrickRef *list, const int *count,
                            float *bricks, UpdateCounters uc, int grid, cudaStream_t s)
{
    if (dense_v1()) {
        const int sms = sm_count_for(grid);
        if (f.unit_w)
            k_integrate_dense_v1<true><<<grid, INTEGRATE_THREADS, 0, s>>>(f, v, list, count, bricks, uc, sms);
        else
            k_integrate_dense_v1<false><<<grid, INTEGRATE_THREADS, 0, s>>>(f, v, list, count, bricks, uc, sms);
        return;
    }
    launch_batch<1>(&f, v, list, count, bricks, uc, grid, s, f.unit_w != 0);
}

// k frames (1..DENSE_MAX_BATCH) applied in one visit per brick; list entries carry the frame mask.
// Batches of 3 run the K = 4 kernel with the fourth frame masked out of every entry.
void launch_integrate_dense_batch(const FrameParams *f, int k, const VolumeParams &v, const BrickRef *list,
                                  const int *count, float *bricks, UpdateCounters uc, int grid, cudaStream_t s)
{
    bool unit_w = true;
    for (int i = 0; i < k; ++i) unit_w = unit_w && f[i].unit_w != 0;
    FrameParams pad[DENSE_MAX_BATCH];
    for (int i = 0; i < DENSE_MAX_BATCH; ++i) pad[i] = f[i < k ? i : k - 1];
    if (k == 1) launch_batch<1>(pad, v, list, count, bricks, uc, grid, s, unit_w);
    else if (k == 2) launch_batch<2>(pad, v, list, count, bricks, uc, grid, s, unit_w);
    else launch_batch<4>(pad, v, list, count, bricks, uc, grid, s, unit_w);
}

void launch_unbrick(const float *bricks, int n_x, int dim_y, int dim_z, int nby, int nbz, float *tsdf, float *wgt,
                    float *col, int sm_count, cudaStream_t s)
{
    if (!n_x) return;
    k_unbrick<<<sm_count * 16, 256, 0, s>>>(bricks, n_x, dim_y, dim_z, nby, nbz, tsdf, wgt, col);
}
