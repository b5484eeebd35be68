// integrate.cuh -- the per-thread integrate body shared by the dense-grid kernel
// (dense_kernels.cu) and the hash-map kernel (hash_kernels.cu).
#pragma once
#include "tsdf_common.cuh"

// ---- record staging through shared memory (TMA-class bulk copy + mbarrier) ----------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// One elected thread: arm the barrier with the byte count and start the 6 KB bulk copy.
__device__ __forceinline__ void bulk_load_record(float *dst_smem, const float *src, uint64_t *bar)
{
    const uint32_t b = smem_u32(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(BRICK_FLOATS * 4) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src), "r"(BRICK_FLOATS * 4), "r"(b)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    const uint32_t b = smem_u32(bar);
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "WAIT_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra DONE_%=;\n"
                 "bra WAIT_%=;\n"
                 "DONE_%=:\n"
                 "}" ::"r"(b),
                 "r"(parity)
                 : "memory");
}

// Measurement switches (tools/ab_variants.sh; defaults are the shipped configuration).
//   PIXEL_EVICT_LAST=1: depth / colour pixels are loaded with an L1 evict-last hint, so the record
//                       stream (evict-first) does not push them out of L1.
//   QUAD_UNROLL=1:      the two quads of a column are unrolled into one instruction stream.
#ifndef PIXEL_EVICT_LAST
#define PIXEL_EVICT_LAST 0
#endif
#ifndef QUAD_UNROLL
#define QUAD_UNROLL 0
#endif
// One prepared pixel (depth + folded colour) in a single 16-byte read-only load.
__device__ __forceinline__ void load_pixel(const FramePixel *p, double &depth, uint32_t &colour)
{
    uint4 v;
#if PIXEL_EVICT_LAST
    asm volatile("ld.global.nc.L1::evict_last.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
#else
    v = __ldg(reinterpret_cast<const uint4 *>(p));
#endif
    depth = __hiloint2double((int)v.y, (int)v.x);
    colour = v.z;
}
__device__ __forceinline__ double load_pixel_depth(const FramePixel *p) { return __ldg(&p->depth); }

// Where a quad reads its record from: global memory (rec itself) or a staged copy in shared
// memory that becomes valid when the mbarrier phase completes.
struct Staged {
    const float *smem; // this thread's slot inside the staged record (nullptr: read global)
    uint64_t *bar;
    uint32_t parity;
    bool waited;
};

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
//   PROBE = true: geometry and depth test only -- returns how many voxels WOULD be updated and
//   never touches rec (used by the hash map to decide whether a block needs storage).
//   FULL = true: all four voxels exist (nvalid is ignored and the masking disappears).
template <bool UNIT_W, bool PROBE, bool STAGED, bool FULL>
__device__ __forceinline__ unsigned integrate_quad(const FrameParams &f, double wx, double wy, double cXs, double cYs,
                                                   double cZ, const double wz[4], int nvalid, float *rec, Staged *st,
                                                   int h)
{
    double Z[4];
    int pix[4];
    bool cand[4], dirty[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        Z[e] = project_fast(f, cXs, cYs, cZ, wz[e], pix[e], cand[e], dirty[e]);
        if (!FULL && e >= nvalid) { cand[e] = false; dirty[e] = false; pix[e] = 0; }
    }
    if (dirty[0] || dirty[1] || dirty[2] || dirty[3]) { // ~2^-15 of the voxels, plus any voxel with Z <= 0
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (dirty[e] && Z[e] > 0.0) { // grid_fusion.py:277
                const int p = exact_pixel(f.T[0], f.T[1], f.T[2], f.T[3], f.T[4], f.T[5], f.T[6], f.T[7], wx, wy, wz[e], Z[e],
                                          f.fx, f.fy, f.cx, f.cy, f.W, f.H);
                if (p >= 0) { pix[e] = p; cand[e] = true; }
            }
    }
    if (!(cand[0] || cand[1] || cand[2] || cand[3])) return 0;
    // depth test (grid_fusion.py:278-286); record and colour requests ride along
    double d[4];
    uint32_t nc[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        if (PROBE) d[e] = load_pixel_depth(f.px + pix[e]);
        else load_pixel(f.px + pix[e], d[e], nc[e]);
    }
    float4 T4 = make_float4(1.f, 1.f, 1.f, 1.f), W4 = make_float4(0.f, 0.f, 0.f, 0.f);
    uint4 C4 = make_uint4(0u, 0u, 0u, 0u);
    if (!PROBE) {
        if (!STAGED) {
            T4 = __ldcs(reinterpret_cast<const float4 *>(rec));
            W4 = __ldcs(reinterpret_cast<const float4 *>(rec + BRICK_VOX));
            C4 = __ldcs(reinterpret_cast<const uint4 *>(rec + 2 * BRICK_VOX));
        }
        if (STAGED) { // the bulk copy was started a whole brick ago; normally this does not block
            if (!st->waited) {
                mbar_wait(st->bar, st->parity);
                st->waited = true;
            }
            const float *sr = st->smem + 4 * h;
            T4 = *reinterpret_cast<const float4 *>(sr);
            W4 = *reinterpret_cast<const float4 *>(sr + BRICK_VOX);
            C4 = *reinterpret_cast<const uint4 *>(sr + 2 * BRICK_VOX);
        }
    }
    double diff[4];
    bool pass[4], band[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        diff[e] = __dadd_rn(d[e], -Z[e]);
        pass[e] = cand[e] && diff[e] >= f.neg_trunc; // rejected depths are stored as -inf (FramePixel)
        band[e] = diff[e] < f.trunc;
    }
    if (!(pass[0] || pass[1] || pass[2] || pass[3])) return 0;
    if (PROBE) return (unsigned)pass[0] + (unsigned)pass[1] + (unsigned)pass[2] + (unsigned)pass[3];
    float t[4] = {T4.x, T4.y, T4.z, T4.w}, w[4] = {W4.x, W4.y, W4.z, W4.w};
    uint32_t c[4] = {C4.x, C4.y, C4.z, C4.w};
    bool slow[4];
    // non-passing voxels of the quad must not veto the colour shortcut: give them a zero difference
    uint32_t ncq[4];
    float wq[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        ncq[e] = pass[e] ? nc[e] : c[e];
        wq[e] = pass[e] ? w[e] : 1024.0f;
    }
    const bool skip_color = UNIT_W && colour_converged4(c, ncq, wq);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        float we = w[e];
        uint32_t ce = c[e];
        const bool ok = update_voxel_fast<UNIT_W>(f, t[e], we, ce, band[e], nc[e], skip_color);
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
template <bool UNIT_W, bool PROBE, bool STAGED = false>
__device__ __forceinline__ unsigned integrate_column8(const FrameParams &f, const VolumeParams &v, int gx, int gy,
                                                      int gz0, float *rec, Staged *st = nullptr)
{
    const double wx = __ldg(v.lat_x + gx), wy = __ldg(v.lat_y + gy);
    const double cZ = __fma_rn(f.T[9], wy, __dmul_rn(f.T[8], wx)); // exact dgemm partial sum of the Z row
    const double cXs = __fma_rn(f.sT[1], wy, __fma_rn(f.sT[0], wx, f.sT[3]));
    const double cYs = __fma_rn(f.sT[5], wy, __fma_rn(f.sT[4], wx, f.sT[7]));
    unsigned n = 0;
#if QUAD_UNROLL
#pragma unroll
#else
#pragma unroll 1
#endif
    for (int h = 0; h < 2; ++h) {
        const int gz = gz0 + 4 * h;
        const int nvalid = v.dim_z - gz;
        if (nvalid <= 0) break;
        const double2 wz01 = __ldg(reinterpret_cast<const double2 *>(v.lat_z + gz));
        const double2 wz23 = __ldg(reinterpret_cast<const double2 *>(v.lat_z + gz) + 1);
        const double wz[4] = {wz01.x, wz01.y, wz23.x, wz23.y};
        if (nvalid >= 4)
            n += integrate_quad<UNIT_W, PROBE, STAGED, true>(f, wx, wy, cXs, cYs, cZ, wz, 4, rec + 4 * h, st, h);
        else
            n += integrate_quad<UNIT_W, PROBE, STAGED, false>(f, wx, wy, cXs, cYs, cZ, wz, nvalid, rec + 4 * h, st, h);
    }
    return n;
}

