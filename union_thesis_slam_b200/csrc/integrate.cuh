// integrate.cuh -- the per-thread integrate body shared by the dense-grid kernels
// (dense_kernels.cu) and the hash-map kernel (hash_kernels.cu).
#pragma once
#include "tsdf_common.cuh"

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Measurement switch (tools/ab_variants.sh; the default is the shipped configuration).
//   PIXEL_EVICT_LAST=1: prepared pixels are loaded with an L1 evict-last hint, so the record
//                       stream does not push them out of L1.
#ifndef PIXEL_EVICT_LAST
#define PIXEL_EVICT_LAST 0
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
#ifndef RECORD_LOAD_EARLY
#define RECORD_LOAD_EARLY 0
#endif
__device__ __forceinline__ float4 ldcs_early(const float *p)
{
    float4 v;
    asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ double load_pixel_depth(const FramePixel *p) { return __ldg(&p->depth); }

// Geometry of four z-consecutive voxels of one (x, y) column (Appendix A steps 3-5): straight-line
// fast path for the four (ILP), the rare IEEE-division pixel peeled off.
//   cXs/cYs/cZ: the column's partial sums (see integrate_column8); wz[4]: lattice z of the voxels;
//   FULL = true: all four voxels exist (nvalid is ignored and the masking disappears).
// Returns whether any voxel projects into the image (cand[e], pix[e] valid where cand[e]).
template <bool FULL>
__device__ __forceinline__ bool quad_geometry(const FrameParams &f, double wx, double wy, double cXs, double cYs, double cZ,
                                              const double wz[4], int nvalid, double Z[4], int pix[4])
{
    bool cand[4], dirty[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        Z[e] = project_fast(f, cXs, cYs, cZ, wz[e], pix[e], cand[e], dirty[e]);
        if (!FULL && e >= nvalid) { cand[e] = false; dirty[e] = false; pix[e] = f.npix; }
    }
    bool any = cand[0] || cand[1] || cand[2] || cand[3];
    if (dirty[0] || dirty[1] || dirty[2] || dirty[3]) { // ~2^-14 of the voxels, plus any voxel with Z <= 0
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (dirty[e] && Z[e] > 0.0) { // grid_fusion.py:277
                const int p = exact_pixel(f.T[0], f.T[1], f.T[2], f.T[3], f.T[4], f.T[5], f.T[6], f.T[7], wx, wy, wz[e], Z[e],
                                          f.fx, f.fy, f.cx, f.cy, f.W, f.H);
                if (p >= 0) { pix[e] = p; any = true; }
            }
    }
    return any;
}

// Depth test of the four voxels (grid_fusion.py:278-286) given their pixels' depths.  Rejected
// depths are stored as -inf in the prepared frame and a voxel that projects nowhere reads the
// sentinel pixel (also -inf), so `depth - Z >= -trunc` alone decides.  The float64 comparisons are
// done once; everything downstream works on the integer class:
//   cls = 0: not updated;  1: updated, inside the truncation band (diff < trunc);  2: updated, free
//   space in front of the band (min(1, diff/trunc) == 1 exactly, grid_fusion.py:286).
__device__ __forceinline__ bool quad_depth_test(const FrameParams &f, const double Z[4], const double d[4], double diff[4],
                                                int cls[4])
{
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        diff[e] = __dadd_rn(d[e], -Z[e]);
        cls[e] = (diff[e] >= f.neg_trunc ? 1 : 0) + (diff[e] >= f.trunc ? 1 : 0);
    }
    return (cls[0] | cls[1] | cls[2] | cls[3]) != 0;
}

// Out of line on purpose: inlined, the compiler runs the test speculatively for every quad; as a call
// it costs nothing until weights are large enough for it to have a chance (see quad_update).
static __device__ __noinline__ bool colour_converged4_call(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t n0,
                                                           uint32_t n1, uint32_t n2, uint32_t n3, float w0, float w1, float w2,
                                                           float w3)
{
    const uint32_t c[4] = {c0, c1, c2, c3}, n[4] = {n0, n1, n2, n3};
    const float w[4] = {w0, w1, w2, w3};
    return colour_converged4(c, n, w);
}

// Read-modify-write of the four voxels in registers (Appendix A steps 7-9).  Returns whether any
// voxel's tsdf changed (only then does the tsdf plane need to be stored: free space seen with
// tsdf still 1 keeps it at 1 exactly).
//   Fast path per voxel: free space (diff >= trunc, so dist == 1) seen with tsdf still 1 and a weight whose
//   float32 add is exact -- then (w*1 + obs*1) / w_new is exactly 1 and f32(f64(w) + obs) == w + obs_f, so
//   only weight and colour change.  UNIT_W (every frame so far had obs_weight == 1, weights are integers):
//   exact for w < 32767; otherwise the two-sum identity checks the add.  Everything else (truncation band,
//   re-observed surface voxels, unusual weights) goes through update_voxel_slow.
//   The four colour blends sit behind ONE branch (skip_colour is per thread, not per voxel).
template <bool UNIT_W>
__device__ __forceinline__ bool quad_update(const FrameParams &f, const double diff[4], const uint32_t nc[4], float t[4],
                                            float w[4], uint32_t c[4])
{
    // predicates straight from the float64 comparisons (measured cheaper than carrying the integer class of
    // quad_depth_test through the update, profiles/r2_dense_ab.md)
    bool pass[4], free_space[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        pass[e] = diff[e] >= f.neg_trunc;
        free_space[e] = diff[e] >= f.trunc;
    }
    float wn[4];
    bool ok[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        wn[e] = __fadd_rn(w[e], UNIT_W ? 1.0f : f.obs_f);
        ok[e] = free_space[e] && t[e] == 1.0f;
        if (UNIT_W)
            ok[e] = ok[e] && w[e] < 32767.0f; // integer weights: w + 1 exact, and blend_color_unit's range
        else
            ok[e] = ok[e] && f.fast_w && (__fadd_rn(wn[e], -w[e]) == f.obs_f) && (__fadd_rn(wn[e], -f.obs_f) == w[e]);
    }
    // Converged-colour shortcut (integer weights): worth testing only once weights are large enough
    // for 2*|n - o| < w + 1 to hold against sensor noise; the cheap pre-check only gates the attempt.
    bool skip_colour = false;
    if (UNIT_W && fmaxf(w[0], w[3]) >= 12.0f) {
        // non-passing voxels of the quad must not veto the shortcut: give them a zero difference
        uint32_t ncq[4];
        float wq[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            ncq[e] = pass[e] ? nc[e] : c[e];
            wq[e] = pass[e] ? w[e] : 1024.0f;
        }
        skip_colour = colour_converged4_call(c[0], c[1], c[2], c[3], ncq[0], ncq[1], ncq[2], ncq[3], wq[0], wq[1], wq[2], wq[3]);
    }
    uint32_t cn[4] = {c[0], c[1], c[2], c[3]};
    if (!skip_colour) {
#pragma unroll
        for (int e = 0; e < 4; ++e)
            cn[e] = UNIT_W ? blend_color_unit(c[e], nc[e], wn[e], f.k23)
                           : blend_color<false>(c[e], nc[e], w[e], wn[e], f.obs_f, f.k23);
    }
    bool slow[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        slow[e] = pass[e] && !ok[e];
        if (ok[e]) { w[e] = wn[e]; c[e] = cn[e]; }
    }
    const bool any_slow = slow[0] || slow[1] || slow[2] || slow[3];
    if (any_slow) {
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (slow[e]) {
                // grid_fusion.py:286: min(1, diff/trunc); the quotient is >= 1 exactly when diff >= trunc
                const double dist = free_space[e] ? 1.0 : fmin(1.0, __ddiv_rn(diff[e], f.trunc));
                Voxel vin;
                vin.tsdf = t[e]; vin.wgt = w[e]; vin.col = c[e];
                const Voxel vo = update_voxel_slow(vin, dist, nc[e], f.obs, f.obs_f);
                t[e] = vo.tsdf; w[e] = vo.wgt; c[e] = vo.col;
            }
    }
    return any_slow;
}

// voxels updated = voxels of class 1 or 2
__device__ __forceinline__ unsigned quad_count(const int cls[4])
{
    const unsigned s = (unsigned)cls[0] | ((unsigned)cls[1] << 2) | ((unsigned)cls[2] << 4) | ((unsigned)cls[3] << 6);
    return (unsigned)__popc((s | (s >> 1)) & 0x55u);
}

// One quad against a record in global memory (hash-map kernel; PROBE mode of both maps).  The record
// is requested together with the pixels, as soon as one voxel projects into the image, so the
// dependent latencies overlap; it bypasses L1 (streamed once per frame) so L1 keeps the pixels.
//   PROBE = true: geometry and depth test only -- returns how many voxels WOULD be updated and
//   never touches rec (used by the hash map to decide whether a block needs storage).
template <bool UNIT_W, bool PROBE, bool FULL>
__device__ __forceinline__ unsigned integrate_quad(const FrameParams &f, double wx, double wy, double cXs, double cYs,
                                                   double cZ, const double wz[4], int nvalid, float *rec)
{
    double Z[4];
    int pix[4];
    if (!quad_geometry<FULL>(f, wx, wy, cXs, cYs, cZ, wz, nvalid, Z, pix)) return 0;
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
#if RECORD_LOAD_EARLY
        // volatile: the compiler otherwise sinks these loads below the depth-test branch, which puts the
        // record's DRAM latency AFTER the pixels' instead of beside it.  Mode 2 issues only the weight and
        // colour planes early (8 registers instead of 12) and leaves the tsdf plane where ptxas puts it.
        if (RECORD_LOAD_EARLY == 1) T4 = ldcs_early(rec);
        else T4 = __ldcs(reinterpret_cast<const float4 *>(rec));
        W4 = ldcs_early(rec + BRICK_VOX);
        const float4 c4 = ldcs_early(rec + 2 * BRICK_VOX);
        C4 = make_uint4(__float_as_uint(c4.x), __float_as_uint(c4.y), __float_as_uint(c4.z), __float_as_uint(c4.w));
#else
        T4 = __ldcs(reinterpret_cast<const float4 *>(rec));
        W4 = __ldcs(reinterpret_cast<const float4 *>(rec + BRICK_VOX));
        C4 = __ldcs(reinterpret_cast<const uint4 *>(rec + 2 * BRICK_VOX));
#endif
    }
    double diff[4];
    int cls[4];
    if (!quad_depth_test(f, Z, d, diff, cls)) return 0;
    const unsigned npass = quad_count(cls);
    if (PROBE) return npass;
    float t[4] = {T4.x, T4.y, T4.z, T4.w}, w[4] = {W4.x, W4.y, W4.z, W4.w};
    uint32_t c[4] = {C4.x, C4.y, C4.z, C4.w};
    if (quad_update<UNIT_W>(f, diff, nc, t, w, c))
        __stcs(reinterpret_cast<float4 *>(rec), make_float4(t[0], t[1], t[2], t[3]));
    __stcs(reinterpret_cast<float4 *>(rec + BRICK_VOX), make_float4(w[0], w[1], w[2], w[3]));
    __stcs(reinterpret_cast<uint4 *>(rec + 2 * BRICK_VOX), make_uint4(c[0], c[1], c[2], c[3]));
    return npass;
}

// Per-column partial sums of one frame: the exact dgemm partial sum of the Z row and the
// fx- / fy-scaled numerator rows of the fast path.
__device__ __forceinline__ void column_sums(const FrameParams &f, double wx, double wy, double &cXs, double &cYs, double &cZ)
{
    cZ = __fma_rn(f.T[9], wy, __dmul_rn(f.T[8], wx));
    cXs = __fma_rn(f.sT[1], wy, __fma_rn(f.sT[0], wx, f.sT[3]));
    cYs = __fma_rn(f.sT[5], wy, __fma_rn(f.sT[4], wx, f.sT[7]));
}

// One (x, y) column of a brick: 8 voxels as two quads, record in global memory.  Returns the number of
// voxels updated.
template <bool UNIT_W, bool PROBE>
__device__ __forceinline__ unsigned integrate_column8(const FrameParams &f, const VolumeParams &v, int gx, int gy,
                                                      int gz0, float *rec)
{
    const double wx = __ldg(v.lat_x + gx), wy = __ldg(v.lat_y + gy);
    double cXs, cYs, cZ;
    column_sums(f, wx, wy, cXs, cYs, cZ);
    unsigned n = 0;
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
        const int gz = gz0 + 4 * h;
        const int nvalid = v.dim_z - gz;
        if (nvalid <= 0) break;
        const double2 wz01 = __ldg(reinterpret_cast<const double2 *>(v.lat_z + gz));
        const double2 wz23 = __ldg(reinterpret_cast<const double2 *>(v.lat_z + gz) + 1);
        const double wz[4] = {wz01.x, wz01.y, wz23.x, wz23.y};
        if (nvalid >= 4)
            n += integrate_quad<UNIT_W, PROBE, true>(f, wx, wy, cXs, cYs, cZ, wz, 4, rec + 4 * h);
        else
            n += integrate_quad<UNIT_W, PROBE, false>(f, wx, wy, cXs, cYs, cZ, wz, nvalid, rec + 4 * h);
    }
    return n;
}

// ---- dense grid: record quads staged in thread-private shared-memory slots ------------------------
// Every thread owns three 16-byte slots (tsdf, weight, colour of one quad of its column) in shared
// memory.  They are filled by its own cp.async (LDGSTS, L1 bypassed) as soon as the previous quad is
// done with them, so the DRAM latency of a record hides behind the quad's geometry and pixel loads
// instead of stalling its first use, and no register is held for it.  Only the issuing thread reads a slot, so the
// completion wait (cp.async.wait_group) is all the synchronisation there is: no CTA barrier.
// With several frames applied to a brick in one visit (temporal blocking) the slot is also where
// the record lives between frames.
__device__ __forceinline__ void prefetch_quad(float4 *slot, const float *rec)
{
    // slot: this thread's tsdf slot; weight and colour slots follow at +INTEGRATE_THREADS float4 each
    const uint32_t s = smem_u32(slot);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(rec) : "memory");
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s + 16u * INTEGRATE_THREADS), "l"(rec + BRICK_VOX) : "memory");
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s + 32u * INTEGRATE_THREADS), "l"(rec + 2 * BRICK_VOX)
                 : "memory");
}
__device__ __forceinline__ void prefetch_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void prefetch_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// One frame applied to one quad whose record sits in the thread's shared-memory slots.
//   FINAL = true : the frame is the only one of this visit: results go straight to global memory.
//   FINAL = false: results go back to the slots; flags accumulates QUAD_DIRTY / QUAD_TSDF_DIRTY and
//                  the caller flushes the slots once after the last frame (flush_quad).
#define QUAD_DIRTY 1u
#define QUAD_TSDF_DIRTY 2u
template <bool UNIT_W, bool FULL, bool FINAL>
__device__ __forceinline__ unsigned integrate_quad_slot(const FrameParams &f, double wx, double wy, double cXs, double cYs,
                                                        double cZ, const double wz[4], int nvalid, float4 *slot, float *rec,
                                                        unsigned &flags)
{
    double Z[4];
    int pix[4];
    if (!quad_geometry<FULL>(f, wx, wy, cXs, cYs, cZ, wz, nvalid, Z, pix)) return 0;
    double d[4];
    uint32_t nc[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) load_pixel(f.px + pix[e], d[e], nc[e]);
    double diff[4];
    int cls[4];
    if (!quad_depth_test(f, Z, d, diff, cls)) return 0;
    prefetch_wait<0>(); // this thread's copy of the quad has landed (issued a whole geometry + pixel wait ago)
    const float4 T4 = slot[0], W4 = slot[INTEGRATE_THREADS];
    const uint4 C4 = *reinterpret_cast<const uint4 *>(slot + 2 * INTEGRATE_THREADS);
    float t[4] = {T4.x, T4.y, T4.z, T4.w}, w[4] = {W4.x, W4.y, W4.z, W4.w};
    uint32_t c[4] = {C4.x, C4.y, C4.z, C4.w};
    const bool tsdf_changed = quad_update<UNIT_W>(f, diff, nc, t, w, c);
    if (FINAL) {
        if (tsdf_changed) __stcs(reinterpret_cast<float4 *>(rec), make_float4(t[0], t[1], t[2], t[3]));
        __stcs(reinterpret_cast<float4 *>(rec + BRICK_VOX), make_float4(w[0], w[1], w[2], w[3]));
        __stcs(reinterpret_cast<uint4 *>(rec + 2 * BRICK_VOX), make_uint4(c[0], c[1], c[2], c[3]));
    } else {
        if (tsdf_changed) {
            slot[0] = make_float4(t[0], t[1], t[2], t[3]);
            flags |= QUAD_TSDF_DIRTY;
        }
        slot[INTEGRATE_THREADS] = make_float4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<uint4 *>(slot + 2 * INTEGRATE_THREADS) = make_uint4(c[0], c[1], c[2], c[3]);
        flags |= QUAD_DIRTY;
    }
    return quad_count(cls);
}

__device__ __forceinline__ void flush_quad(const float4 *slot, float *rec, unsigned flags)
{
    if (!(flags & QUAD_DIRTY)) return;
    if (flags & QUAD_TSDF_DIRTY) __stcs(reinterpret_cast<float4 *>(rec), slot[0]);
    __stcs(reinterpret_cast<float4 *>(rec + BRICK_VOX), slot[INTEGRATE_THREADS]);
    __stcs(reinterpret_cast<float4 *>(rec + 2 * BRICK_VOX), slot[2 * INTEGRATE_THREADS]);
}
