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
                             FramePixel *__restrict__ px, FrameScratch *fs, TilePyramid pyr, int H, int W)
{
    const int tile = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (tile >= pyr.tx[0] * pyr.ty[0]) return;
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
        pyr.level[0][tile] = m;
        if (m > 0.0f) {
            const unsigned bits = __float_as_uint(m);
            atomicMax(reinterpret_cast<unsigned *>(pyr.level[1]) + (ty >> 1) * pyr.tx[1] + (tx >> 1), bits);
            atomicMax(reinterpret_cast<unsigned *>(pyr.level[2]) + (ty >> 2) * pyr.tx[2] + (tx >> 2), bits);
            atomicMax(reinterpret_cast<unsigned *>(pyr.level[3]) + (ty >> 3) * pyr.tx[3] + (tx >> 3), bits);
            atomicMax(&fs->max_depth_bits, bits);
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
        bx = lbx * world + rank;
        cls = (lbx + by + bz) % classes;
        c.max_depth = (double)__uint_as_float(fs->max_depth_bits);
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

// Persistent CTAs of two warps; one brick per iteration, thread t owns the z-column
// (x = t>>3, y = t&7): two float4 (= one 32 B sector) per plane, read with streaming loads right
// after the depth pixels are requested and written back straight from registers with streaming
// stores.  One launch per class of bricks (see k_cull_bricks); `count` is that class's list length.
// -DINTEGRATE_STAGED=1 builds the measured-and-rejected variant instead (profiles/r1_dense_ab.txt):
// records double-buffered through shared memory by one thread's 6 KB bulk copy (cp.async.bulk +
// mbarrier, the TMA path) of brick i+1 while brick i is integrated.
#ifndef INTEGRATE_L2_PREFETCH
#define INTEGRATE_L2_PREFETCH 0
#endif
#ifndef INTEGRATE_STAGED
#define INTEGRATE_STAGED 0 // measured: 0.63 vs 0.59 of the HBM roofline on the 1 cm room (profiles/)
#endif
#if INTEGRATE_STAGED
template <bool UNIT_W>
__global__ void __launch_bounds__(INTEGRATE_THREADS, INTEGRATE_MIN_CTAS)
k_integrate_dense(FrameParams f, VolumeParams v, const BrickRef *__restrict__ list, const int *__restrict__ count,
                  float *__restrict__ bricks, UpdateCounters uc)
{
    __shared__ __align__(128) float s_rec[2][BRICK_FLOATS];
    __shared__ __align__(8) uint64_t s_bar[2];
    const int n = *count;
    const int t = threadIdx.x;
    const int lx = t >> 3, ly = t & 7;
    if (t == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    unsigned my_updates = 0;
    int b = blockIdx.x;
    int4 br_next = make_int4(0, 0, 0, 0);
    if (b < n) {
        br_next = __ldg(reinterpret_cast<const int4 *>(list) + b);
        if (t == 0) bulk_load_record(s_rec[0], bricks + (size_t)(uint32_t)br_next.w * BRICK_FLOATS, &s_bar[0]);
    }
    uint32_t it = 0;
    for (; b < n; b += gridDim.x, ++it) {
        const int4 br = br_next;
        const uint32_t cur = it & 1u;
        // everyone is done reading the other buffer (it held brick it-1): refill it with brick it+1
        __syncthreads();
        if (b + (int)gridDim.x < n) {
            br_next = __ldg(reinterpret_cast<const int4 *>(list) + b + gridDim.x);
            if (t == 0)
                bulk_load_record(s_rec[cur ^ 1u], bricks + (size_t)(uint32_t)br_next.w * BRICK_FLOATS, &s_bar[cur ^ 1u]);
        }
        Staged st;
        st.smem = s_rec[cur] + 8 * t;
        st.bar = &s_bar[cur];
        st.parity = (it >> 1) & 1u;
        st.waited = false;
        const int gx = br.x * BRICK + lx, gy = br.y * BRICK + ly;
        if (gx < v.dim_x && gy < v.dim_y)
            my_updates += integrate_column8<UNIT_W, false, true>(f, v, gx, gy, br.z * BRICK,
                                                                  bricks + (size_t)(uint32_t)br.w * BRICK_FLOATS + 8 * t, &st);
        // every thread observes the completion of this phase, so the barrier can be re-armed safely
        if (!st.waited) mbar_wait(st.bar, st.parity);
    }
    for (int o = 16; o; o >>= 1) my_updates += __shfl_xor_sync(0xffffffffu, my_updates, o);
    if ((t & 31) == 0) { // one private slot per warp: no atomics
        const int w = blockIdx.x * (INTEGRATE_THREADS / 32) + (t >> 5);
        uc.frame[w] = my_updates;
        uc.total[w] += my_updates;
    }
}

#else
template <bool UNIT_W>
__global__ void __launch_bounds__(INTEGRATE_THREADS, INTEGRATE_MIN_CTAS)
k_integrate_dense(FrameParams f, VolumeParams v, const BrickRef *__restrict__ list, const int *__restrict__ count,
                  float *__restrict__ bricks, UpdateCounters uc)
{
    const int n = *count;
    const int t = threadIdx.x;
    const int lx = t >> 3, ly = t & 7;
    unsigned my_updates = 0;
    int b = blockIdx.x;
    int4 br_next = make_int4(0, 0, 0, 0);
    if (b < n) br_next = __ldg(reinterpret_cast<const int4 *>(list) + b);
    for (; b < n; b += gridDim.x) {
        const int4 br = br_next;
        if (b + (int)gridDim.x < n) {
            br_next = __ldg(reinterpret_cast<const int4 *>(list) + b + gridDim.x);
#if INTEGRATE_L2_PREFETCH
            // one bulk-prefetch instruction pulls the next brick's 6 KB record into L2
            if (t == 0) prefetch_record_l2(bricks + (size_t)(uint32_t)br_next.w * BRICK_FLOATS);
#endif
        }
        const int gx = br.x * BRICK + lx, gy = br.y * BRICK + ly;
        if (gx >= v.dim_x || gy >= v.dim_y) continue;
        my_updates += integrate_column8<UNIT_W, false>(f, v, gx, gy, br.z * BRICK,
                                                       bricks + (size_t)(uint32_t)br.w * BRICK_FLOATS + 8 * t);
    }
    for (int o = 16; o; o >>= 1) my_updates += __shfl_xor_sync(0xffffffffu, my_updates, o);
    if ((t & 31) == 0) { // one private slot per warp: no atomics
        const int w = blockIdx.x * (INTEGRATE_THREADS / 32) + (t >> 5);
        uc.frame[w] = my_updates;
        uc.total[w] += my_updates;
    }
}
#endif

// Stats read-out: sums the per-warp update counters and gathers the scalar counters into one block.
__global__ void k_collect_stats(UpdateCounters uc, int n_warps, const FrameScratch *fs, const Totals *tot,
                                const int *n_entries, const int *heap_top, StatsBlock *out)
{
    __shared__ unsigned long long s_a[8], s_b[8];
    unsigned long long a = 0, b = 0;
    for (int i = threadIdx.x; i < n_warps; i += blockDim.x) {
        a += uc.frame[i];
        b += uc.total[i];
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
        r.updated_last = a; r.updated_total = b;
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
                       const TilePyramid &pyr, int H, int W, cudaStream_t s)
{
    const int warps = pyr.tx[0] * pyr.ty[0];
    k_prep_frame<<<(warps * 32 + 127) / 128, 128, 0, s>>>(rgb, depth, px, fs, pyr, H, W);
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

#ifndef INTEGRATE_RESERVE_CTAS
#define INTEGRATE_RESERVE_CTAS 2
#endif
int integrate_reserve_ctas() { return INTEGRATE_RESERVE_CTAS; }

void launch_collect_stats(UpdateCounters uc, int n_warps, const FrameScratch *fs, const Totals *tot, const int *n_entries,
                          const int *heap_top, StatsBlock *out, cudaStream_t s)
{
    k_collect_stats<<<1, 256, 0, s>>>(uc, n_warps, fs, tot, n_entries, heap_top, out);
}

int integrate_dense_ctas_per_sm()
{
    int n = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_integrate_dense<true>, INTEGRATE_THREADS, 0);
    return n > 0 ? n : 1;
}

void launch_integrate_dense(const FrameParams &f, const VolumeParams &v, const BrickRef *list, const int *count,
                            float *bricks, UpdateCounters uc, int grid, cudaStream_t s)
{
    if (f.unit_w)
        k_integrate_dense<true><<<grid, INTEGRATE_THREADS, 0, s>>>(f, v, list, count, bricks, uc);
    else
        k_integrate_dense<false><<<grid, INTEGRATE_THREADS, 0, s>>>(f, v, list, count, bricks, uc);
}

void launch_unbrick(const float *bricks, int n_x, int dim_y, int dim_z, int nby, int nbz, float *tsdf, float *wgt,
                    float *col, int sm_count, cudaStream_t s)
{
    if (!n_x) return;
    k_unbrick<<<sm_count * 16, 256, 0, s>>>(bricks, n_x, dim_y, dim_z, nby, nbz, tsdf, wgt, col);
}
