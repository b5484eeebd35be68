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
#include "tsdf_common.cuh"

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

// grid_fusion.py:228-232 colour fold (as an exact integer) + max valid depth of the frame.
__global__ void k_prep_frame(const uint8_t *__restrict__ rgb, const double *__restrict__ depth,
                             uint32_t *__restrict__ cfold, FrameScratch *fs, int npix)
{
    float m = 0.0f;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += gridDim.x * blockDim.x) {
        const uint32_t r = rgb[3 * p], g = rgb[3 * p + 1], b = rgb[3 * p + 2];
        cfold[p] = (b << 16) | (g << 8) | r;
        const double d = depth[p];
        if (d > 0.0) m = fmaxf(m, __double2float_ru(d));
    }
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(&fs->max_depth_bits, __float_as_uint(m));
}

// One thread per local brick; visible bricks are appended with one atomic per warp.
__global__ void k_cull_bricks(FrameParams f, CullParams c, int nlbx, int nby, int nbz, int rank, int world,
                              BrickRef *__restrict__ list, FrameScratch *fs)
{
    const int n = nlbx * nby * nbz;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool vis = false;
    int bx = 0, by = 0, bz = 0;
    if (i < n) {
        bz = i % nbz;
        by = (i / nbz) % nby;
        bx = (i / (nbz * nby)) * world + rank;
        c.max_depth = (double)__uint_as_float(fs->max_depth_bits);
        vis = c.max_depth > 0.0 && brick_visible(f, c, bx, by, bz);
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, vis);
    if (!ballot) return;
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0) base = atomicAdd(&fs->vis_count, __popc(ballot));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (vis) {
        BrickRef r;
        r.bx = bx; r.by = by; r.bz = bz; r.storage = (uint32_t)i;
        list[base + __popc(ballot & ((1u << lane) - 1u))] = r;
    }
}

// Per-thread body shared by the dense and the hash integrate kernels: geometry of four
// z-consecutive voxels of one (x, y) column, then read-modify-write of their float4 per plane.
//   cX/cY/cZ: the column's partial sums fma(T[r][1], wy, T[r][0]*wx); wz[4]: lattice z of the
//   voxels; nvalid: how many of the four exist (z < dim_z).  rec: the float4 slot inside the brick
//   record.  Returns the number of voxels updated.
// Straight-line fast paths for 4 voxels (ILP), rare cases (IEEE-division pixel, band / general
// weights) peeled off.  The voxel record is requested as soon as one voxel projects into the
// image, together with the depth pixels, so the three dependent latencies (depth, record,
// colour) overlap; the record bypasses L1 (streamed once per frame) so L1 keeps the depth /
// colour pixels.
template <bool UNIT_W>
__device__ __forceinline__ unsigned integrate_quad(const FrameParams &f, double cX, double cY, double cZ,
                                                   const double wz[4], int nvalid, float *rec)
{
    double Z[4], ax[4], ay[4];
    int pix[4];
    bool cand[4], exact[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        Z[e] = project_fast(f, cX, cY, cZ, wz[e], ax[e], ay[e], pix[e], cand[e], exact[e]);
        if (e >= nvalid) { cand[e] = false; exact[e] = false; pix[e] = 0; }
    }
    if (exact[0] || exact[1] || exact[2] || exact[3]) { // ~2^-15 of the voxels
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (exact[e]) {
                const int p = exact_pixel(ax[e], ay[e], Z[e], f.cx, f.cy, f.W, f.H);
                if (p >= 0) { pix[e] = p; cand[e] = true; }
            }
    }
    if (!(cand[0] || cand[1] || cand[2] || cand[3])) return 0;
    // depth test (grid_fusion.py:278-286); record and colour requests ride along
    double d[4];
    uint32_t nc[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) d[e] = __ldg(f.depth + pix[e]);
    const float4 T4 = __ldcs(reinterpret_cast<const float4 *>(rec));
    const float4 W4 = __ldcs(reinterpret_cast<const float4 *>(rec + BRICK_VOX));
    const uint4 C4 = __ldcs(reinterpret_cast<const uint4 *>(rec + 2 * BRICK_VOX));
#pragma unroll
    for (int e = 0; e < 4; ++e) nc[e] = __ldg(f.cfold + pix[e]);
    double diff[4];
    bool pass[4], band[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        diff[e] = __dadd_rn(d[e], -Z[e]);
        pass[e] = cand[e] && d[e] > 0.0 && diff[e] >= f.neg_trunc;
        band[e] = diff[e] < f.trunc;
    }
    if (!(pass[0] || pass[1] || pass[2] || pass[3])) return 0;
    float t[4] = {T4.x, T4.y, T4.z, T4.w}, w[4] = {W4.x, W4.y, W4.z, W4.w};
    uint32_t c[4] = {C4.x, C4.y, C4.z, C4.w};
    bool slow[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        float we = w[e];
        uint32_t ce = c[e];
        const bool ok = update_voxel_fast<UNIT_W>(f, t[e], we, ce, band[e], nc[e]);
        if (pass[e] && ok) { w[e] = we; c[e] = ce; }
        slow[e] = pass[e] && !ok;
    }
    if (slow[0] || slow[1] || slow[2] || slow[3]) { // truncation band, re-observed surface voxels, unusual obs_weight
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (slow[e]) {
                // grid_fusion.py:286: min(1, diff/trunc); the quotient is >= 1 exactly when diff >= trunc
                const double dist = band[e] ? fmin(1.0, __ddiv_rn(diff[e], f.trunc)) : 1.0;
                Voxel vin;
                vin.tsdf = t[e]; vin.wgt = w[e]; vin.col = c[e];
                const Voxel vo = update_voxel_slow(vin, dist, nc[e], f.obs, f.obs_f);
                t[e] = vo.tsdf; w[e] = vo.wgt; c[e] = vo.col;
            }
        __stcs(reinterpret_cast<float4 *>(rec), make_float4(t[0], t[1], t[2], t[3]));
    }
    __stcs(reinterpret_cast<float4 *>(rec + BRICK_VOX), make_float4(w[0], w[1], w[2], w[3]));
    __stcs(reinterpret_cast<uint4 *>(rec + 2 * BRICK_VOX), make_uint4(c[0], c[1], c[2], c[3]));
    return (unsigned)pass[0] + (unsigned)pass[1] + (unsigned)pass[2] + (unsigned)pass[3];
}

// One (x, y) column of a brick: 8 voxels as two quads.  Returns the number of voxels updated.
template <bool UNIT_W>
__device__ __forceinline__ unsigned integrate_column8(const FrameParams &f, const VolumeParams &v, int gx, int gy,
                                                      int gz0, float *rec)
{
    const double wx = __ldg(v.lat_x + gx), wy = __ldg(v.lat_y + gy);
    const double cX = __fma_rn(f.T[1], wy, __dmul_rn(f.T[0], wx));
    const double cY = __fma_rn(f.T[5], wy, __dmul_rn(f.T[4], wx));
    const double cZ = __fma_rn(f.T[9], wy, __dmul_rn(f.T[8], wx));
    unsigned n = 0;
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
        const int gz = gz0 + 4 * h;
        const int nvalid = v.dim_z - gz;
        if (nvalid <= 0) break;
        const double2 wz01 = __ldg(reinterpret_cast<const double2 *>(v.lat_z + gz));
        const double2 wz23 = __ldg(reinterpret_cast<const double2 *>(v.lat_z + gz) + 1);
        const double wz[4] = {wz01.x, wz01.y, wz23.x, wz23.y};
        n += integrate_quad<UNIT_W>(f, cX, cY, cZ, wz, nvalid, rec + 4 * h);
    }
    return n;
}

// Persistent CTAs of two warps; one brick per iteration, thread t owns the z-column
// (x = t>>3, y = t&7): two float4 (= one 32 B sector) per plane.  The next list entry is
// fetched while the current brick is processed.
template <bool UNIT_W>
__global__ void __launch_bounds__(INTEGRATE_THREADS, INTEGRATE_MIN_CTAS)
k_integrate_dense(FrameParams f, VolumeParams v, const BrickRef *__restrict__ list, float *__restrict__ bricks,
                  FrameScratch *fs, Totals *tot)
{
    const int n = fs->vis_count;
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
            // pull the next brick's 6 KB record into L2 while this one is integrated
            if (t == 0) prefetch_record_l2(bricks + (size_t)(uint32_t)br_next.w * BRICK_FLOATS);
        }
        const int gx = br.x * BRICK + lx, gy = br.y * BRICK + ly;
        if (gx >= v.dim_x || gy >= v.dim_y) continue;
        my_updates += integrate_column8<UNIT_W>(f, v, gx, gy, br.z * BRICK,
                                                bricks + (size_t)(uint32_t)br.w * BRICK_FLOATS + 8 * t);
    }
    for (int o = 16; o; o >>= 1) my_updates += __shfl_xor_sync(0xffffffffu, my_updates, o);
    if ((t & 31) == 0 && my_updates) {
        atomicAdd(&fs->updated, (unsigned long long)my_updates);
        atomicAdd(&tot->updated, (unsigned long long)my_updates);
    }
    if (blockIdx.x == 0 && t == 0) tot->last_visited = (unsigned long long)n;
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

void launch_prep_frame(const uint8_t *rgb, const double *depth, uint32_t *cfold, FrameScratch *fs, int npix,
                       int sm_count, cudaStream_t s)
{
    int blocks = (npix + 255) / 256;
    if (blocks > sm_count * 8) blocks = sm_count * 8;
    k_prep_frame<<<blocks, 256, 0, s>>>(rgb, depth, cfold, fs, npix);
}

void launch_cull_bricks(const FrameParams &f, const CullParams &c, int nlbx, int nby, int nbz, int rank, int world,
                        BrickRef *list, FrameScratch *fs, cudaStream_t s)
{
    const long long n = (long long)nlbx * nby * nbz;
    if (!n) return;
    k_cull_bricks<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(f, c, nlbx, nby, nbz, rank, world, list, fs);
}

int integrate_dense_ctas_per_sm()
{
    int n = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_integrate_dense<true>, INTEGRATE_THREADS, 0);
    return n > 0 ? n : 1;
}

void launch_integrate_dense(const FrameParams &f, const VolumeParams &v, const BrickRef *list, float *bricks,
                            FrameScratch *fs, Totals *tot, int grid, cudaStream_t s)
{
    if (f.unit_w)
        k_integrate_dense<true><<<grid, INTEGRATE_THREADS, 0, s>>>(f, v, list, bricks, fs, tot);
    else
        k_integrate_dense<false><<<grid, INTEGRATE_THREADS, 0, s>>>(f, v, list, bricks, fs, tot);
}

void launch_unbrick(const float *bricks, int n_x, int dim_y, int dim_z, int nby, int nbz, float *tsdf, float *wgt,
                    float *col, int sm_count, cudaStream_t s)
{
    if (!n_x) return;
    k_unbrick<<<sm_count * 16, 256, 0, s>>>(bricks, n_x, dim_y, dim_z, nby, nbz, tsdf, wgt, col);
}
