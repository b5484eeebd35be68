// bucket_kernels.cu -- device-resident per-voxel hash CONTAINER with the reference's exact slot
// layout: 5-slot buckets, collision chains threaded through
// (bucket, slot) offsets, resize by re-insertion.  C ABI: vbt_* (include/tsdf_b200.h).
//
// Reference: hash_fusion.py:147-180 (metrics), :182-190 (hash), :199-276 (insert + chain),
// :285-320 (lookup), :322-393 (removal), :414-437 (resize), data_structures/bucket.py,
// data_structures/hash_entry.py, data_structures/voxel.py:19-49 (per-voxel running mean).
//
// The reference container is order-dependent by construction: an insert takes the first free slot
// of its home bucket, else walks its chain to the tail and scans forward for the next bucket with
// a free slot, so where an entry lands depends on everything inserted before it.  Mutating ops
// (add / remove / integrate) are therefore applied IN ORDER by one device thread over tables that
// live in HBM (the point is layout parity with the reference container -- returned (bucket, slot)
// pairs, chains, resize points -- not throughput; the throughput path is the lock-free block table
// of hash_kernels.cu).  Pure lookups of a batch run one thread per query.
// The host only drives the control flow the device cannot: a resize allocates the doubled table
// and re-adds the old entries in slot order, recursively when re-adding crosses the load factor
// again, exactly as the reference's add_hash_entry -> double_table_size -> add_hash_entry does.
#include <new>
#include <vector>

#include "host_common.h"

#define VBT_P1 73856093LL
#define VBT_P2 19349669LL
#define VBT_P3 83492791LL

enum { VBT_ADD = 0, VBT_GET = 1, VBT_REMOVE = 2, VBT_INTEGRATE = 3 };
enum { VBT_DONE = 0, VBT_NEED_RESIZE = 1, VBT_DANGLING = -1 };

struct BucketDev {
    int32_t *pos;     // [S*B][3] voxel coordinate of the entry in slot (b, e) = index b*B + e
    int32_t *off;     // [S*B][2] chain pointer (bucket, slot); (-1, -1) = empty offset
    double *vox;      // [S*B][3] voxel payload (sdf, colour, weight); NaN sdf = entry without voxel
    uint8_t *occ;     // [S*B]   slot holds an entry
    uint8_t *exists;  // [S]     the reference's table[b] is a Bucket object (not None)
    long long *nonempty; // _num_non_empty_bucket
    long long S;
    int B;
};

struct VbtStatus {
    long long next; // op / slot index at which the kernel stopped
    int code;       // VBT_DONE, VBT_NEED_RESIZE, VBT_DANGLING
    int pad;
};

struct Entry {
    int x, y, z;
    double sdf, col, wgt;
};

// hash_fusion.py:182-190: ((x*p1) xor (y*p2) xor (z*p3)) mod n, floor-mod on 64-bit integers
__host__ __device__ __forceinline__ long long vbt_hash(long long x, long long y, long long z, long long n)
{
    const long long v = (x * VBT_P1) ^ (y * VBT_P2) ^ (z * VBT_P3);
    const long long r = v % n;
    return r < 0 ? r + n : r;
}

__device__ __forceinline__ int vbt_count(const BucketDev &t, long long b) // bucket.py:62-70
{
    int c = 0;
    for (int e = 0; e < t.B; ++e) c += t.occ[b * t.B + e];
    return c;
}
__device__ __forceinline__ int vbt_first_free(const BucketDev &t, long long b, int upto) // bucket.py:72-82
{
    for (int e = 0; e < upto; ++e)
        if (!t.occ[b * t.B + e]) return e;
    return -1;
}
__device__ __forceinline__ bool vbt_match(const BucketDev &t, long long i, int x, int y, int z)
{
    return t.occ[i] && t.pos[3 * i] == x && t.pos[3 * i + 1] == y && t.pos[3 * i + 2] == z;
}
__device__ __forceinline__ void vbt_put(const BucketDev &t, long long i, const Entry &en, int ob, int oe)
{
    t.occ[i] = 1;
    t.pos[3 * i] = en.x; t.pos[3 * i + 1] = en.y; t.pos[3 * i + 2] = en.z;
    t.off[2 * i] = ob; t.off[2 * i + 1] = oe;
    t.vox[3 * i] = en.sdf; t.vox[3 * i + 1] = en.col; t.vox[3 * i + 2] = en.wgt;
}
__device__ __forceinline__ void vbt_clear(const BucketDev &t, long long i)
{
    t.occ[i] = 0;
    t.off[2 * i] = -1; t.off[2 * i + 1] = -1;
}
__device__ __forceinline__ bool vbt_needs_resize(const BucketDev &t, double load_factor) // hash_fusion.py:156-161
{
    return __ddiv_rn((double)*t.nonempty, (double)t.S) >= load_factor;
}

// hash_fusion.py:241-276.  prev: slot index of the chain tail whose offset is set.
__device__ void vbt_find_free_and_add(const BucketDev &t, long long begin_b, int begin_e, long long prev, const Entry &en,
                                      long long &rb, int &re)
{
    long long b = begin_b;
    if (begin_e < t.B - 1) {
        for (int i = begin_e; i < t.B - 1; ++i)
            if (!t.occ[b * t.B + i]) {
                vbt_put(t, b * t.B + i, en, -1, -1);
                t.off[2 * prev] = (int)begin_b; t.off[2 * prev + 1] = i;
                rb = begin_b; re = i;
                return;
            }
    }
    bool scanned = false;
    while (!scanned) {
        b = (b >= t.S - 1) ? 0 : b + 1;
        if (!t.exists[b]) {
            t.exists[b] = 1;
            vbt_put(t, b * t.B, en, -1, -1);
            *t.nonempty += 1;
            t.off[2 * prev] = (int)b; t.off[2 * prev + 1] = 0;
            rb = b; re = 0;
            return;
        }
        if (vbt_first_free(t, b, t.B - 1) < 0) { // is_full_for_alien_entries
            scanned = b == begin_b;
        } else {
            const int e = vbt_first_free(t, b, t.B);
            vbt_put(t, b * t.B + e, en, -1, -1);
            t.off[2 * prev] = (int)b; t.off[2 * prev + 1] = e;
            rb = b; re = e;
            return;
        }
    }
    rb = -1; re = -1;
}

// hash_fusion.py:199-239 without the resize check (the caller did it).  Returns false on a dangling chain.
__device__ bool vbt_add(const BucketDev &t, const Entry &en, long long &rb, int &re)
{
    const long long h = vbt_hash(en.x, en.y, en.z, t.S);
    if (!t.exists[h]) {
        t.exists[h] = 1;
        vbt_put(t, h * t.B, en, -1, -1);
        *t.nonempty += 1;
        rb = h; re = 0;
        return true;
    }
    const int e = vbt_first_free(t, h, t.B);
    if (e >= 0) {
        vbt_put(t, h * t.B + e, en, -1, -1);
        rb = h; re = e;
        return true;
    }
    // _add_to_linked_list: walk from the list head (last slot of the home bucket) to the tail
    long long tail = h * t.B + (t.B - 1), tb = h;
    int te = t.B - 1;
    while (t.off[2 * tail] >= 0) {
        tb = t.off[2 * tail]; te = t.off[2 * tail + 1];
        tail = tb * t.B + te;
        if (!t.occ[tail]) return false;
    }
    vbt_find_free_and_add(t, tb, te, tail, en, rb, re);
    return true;
}

// hash_fusion.py:285-310.  Returns false on a dangling chain.
__device__ bool vbt_find(const BucketDev &t, int x, int y, int z, long long &rb, int &re)
{
    rb = -1; re = -1;
    const long long h = vbt_hash(x, y, z, t.S);
    if (!t.exists[h] || vbt_count(t, h) == 0) return true;
    for (int e = 0; e < t.B; ++e)
        if (vbt_match(t, h * t.B + e, x, y, z)) { rb = h; re = e; return true; }
    const long long last = h * t.B + (t.B - 1);
    if (t.occ[last] && t.off[2 * last] >= 0) { // bucket.is_overflow()
        long long b = t.off[2 * last];
        int e = t.off[2 * last + 1];
        while (true) {
            const long long i = b * t.B + e;
            if (!t.occ[i]) return false;
            if (vbt_match(t, i, x, y, z)) { rb = b; re = e; return true; }
            if (t.off[2 * i] < 0) break;
            b = t.off[2 * i]; e = t.off[2 * i + 1];
        }
    }
    return true;
}

// hash_fusion.py:386-393 _remove_hash_entry
__device__ void vbt_remove_at(const BucketDev &t, long long b, int e)
{
    if (!t.exists[b]) return;
    vbt_clear(t, b * t.B + e);
    if (vbt_count(t, b) == 0) {
        t.exists[b] = 0;
        *t.nonempty -= 1;
    }
}

// hash_fusion.py:331-384.  removed: 1 / 0.  Returns false on a dangling chain.
__device__ bool vbt_remove(const BucketDev &t, int x, int y, int z, int &removed)
{
    removed = 0;
    const long long h = vbt_hash(x, y, z, t.S);
    if (!t.exists[h] || vbt_count(t, h) == 0) return true;
    long long last_entry = -1;
    for (int i = 0; i < t.B; ++i) { // case 1: in the home bucket
        const long long s = h * t.B + i;
        if (!t.occ[s]) continue;
        last_entry = s;
        if (!vbt_match(t, s, x, y, z)) continue;
        if (t.off[2 * s] < 0) { // 1a
            vbt_clear(t, s);
            if (vbt_count(t, h) == 0) {
                t.exists[h] = 0;
                *t.nonempty -= 1;
            }
        } else { // 1b: the next chain element moves into this slot (with its own offset)
            const long long pb = t.off[2 * s];
            const int pe = t.off[2 * s + 1];
            const long long n = pb * t.B + pe;
            const bool has = t.occ[n];
            Entry en;
            int ob = -1, oe = -1;
            if (has) {
                en.x = t.pos[3 * n]; en.y = t.pos[3 * n + 1]; en.z = t.pos[3 * n + 2];
                en.sdf = t.vox[3 * n]; en.col = t.vox[3 * n + 1]; en.wgt = t.vox[3 * n + 2];
                ob = t.off[2 * n]; oe = t.off[2 * n + 1];
            }
            vbt_remove_at(t, pb, pe);
            if (has) vbt_put(t, s, en, ob, oe);
            else vbt_clear(t, s); // the reference stores None here and keeps the (possibly empty) Bucket object
        }
        removed = 1;
        return true;
    }
    if (last_entry >= 0) { // case 2: further down the chain that starts at the last non-empty home slot
        long long cur = last_entry;
        while (t.off[2 * cur] >= 0) {
            const long long pb = t.off[2 * cur];
            const int pe = t.off[2 * cur + 1];
            const long long n = pb * t.B + pe;
            if (!t.occ[n]) return false;
            if (vbt_match(t, n, x, y, z)) {
                t.off[2 * cur] = t.off[2 * n]; t.off[2 * cur + 1] = t.off[2 * n + 1]; // 2a: empty, 2b: successor
                vbt_remove_at(t, pb, pe);
                removed = 1;
                return true;
            }
            cur = n;
        }
    }
    return true;
}

// data_structures/voxel.py:19-49 with the reference's default obs_weight = 1 (hash_fusion.py:141,145
// never pass it).  All float64: the colour operands are integers < 2^24, so the reference's mixed
// float32/float64 evaluation is exact up to the final float64 division; rint = numpy's half-to-even.
__device__ void vbt_voxel_integrate(double &sdf, double &col, double &wgt, double dist, double ncol)
{
    const double w_old = wgt, w_new = __dadd_rn(w_old, 1.0);
    sdf = __ddiv_rn(__dadd_rn(__dmul_rn(sdf, w_old), __dmul_rn(dist, 1.0)), w_new);
    wgt = w_new;
    const double ob = floor(__ddiv_rn(col, 65536.0));
    const double og = floor(__ddiv_rn(__dadd_rn(col, -__dmul_rn(ob, 65536.0)), 256.0));
    const double orr = __dadd_rn(__dadd_rn(col, -__dmul_rn(ob, 65536.0)), -__dmul_rn(og, 256.0));
    const double nb = floor(__ddiv_rn(ncol, 65536.0));
    const double ng = floor(__ddiv_rn(__dadd_rn(ncol, -__dmul_rn(nb, 65536.0)), 256.0));
    const double nr = __dadd_rn(__dadd_rn(ncol, -__dmul_rn(nb, 65536.0)), -__dmul_rn(ng, 256.0));
    const double b = fmin(255.0, rint(__ddiv_rn(__dadd_rn(__dmul_rn(w_old, ob), nb), w_new)));
    const double g = fmin(255.0, rint(__ddiv_rn(__dadd_rn(__dmul_rn(w_old, og), ng), w_new)));
    const double r = fmin(255.0, rint(__ddiv_rn(__dadd_rn(__dmul_rn(w_old, orr), nr), w_new)));
    col = __dadd_rn(__dadd_rn(__dmul_rn(b, 65536.0), __dmul_rn(g, 256.0)), r);
}

// Applies ops[begin..n) in order.  Stops with VBT_NEED_RESIZE *before* an insert that finds the
// table at its load factor (hash_fusion.py:207-208); the host doubles the table and relaunches from
// that op with skip_check = 1 (the reference does not re-check after resizing).
//   pos[i][3] int32; args[i][5] = {dist, colour, sdf, col, wgt} (INTEGRATE uses 0..1, ADD uses 2..4)
//   out[i][2]: ADD/GET/INTEGRATE -> (bucket, slot) or (-1, -1); REMOVE -> (1 / 0, -1)
//   ops == nullptr: every op is INTEGRATE (the per-frame voxel stream); out == nullptr: no results wanted
__global__ void k_vbt_apply(BucketDev t, double load_factor, const int32_t *__restrict__ ops,
                            const int32_t *__restrict__ pos, const double *__restrict__ args, long long begin,
                            long long n, int skip_check, long long *__restrict__ out, VbtStatus *st)
{
    if (blockIdx.x || threadIdx.x) return;
    for (long long i = begin; i < n; ++i) {
        const int op = ops ? ops[i] : VBT_INTEGRATE;
        const int x = pos[3 * i], y = pos[3 * i + 1], z = pos[3 * i + 2];
        long long rb = -1;
        int re = -1;
        bool ok = true;
        if (op == VBT_GET) {
            ok = vbt_find(t, x, y, z, rb, re);
        } else if (op == VBT_REMOVE) {
            int removed = 0;
            ok = vbt_remove(t, x, y, z, removed);
            rb = removed;
        } else {
            bool insert = true;
            Entry en;
            en.x = x; en.y = y; en.z = z;
            if (op == VBT_INTEGRATE) { // hash_fusion.py:135-145
                ok = vbt_find(t, x, y, z, rb, re);
                if (ok && rb >= 0) {
                    const long long s = rb * t.B + re;
                    vbt_voxel_integrate(t.vox[3 * s], t.vox[3 * s + 1], t.vox[3 * s + 2], args[5 * i], args[5 * i + 1]);
                    insert = false;
                } else {
                    en.sdf = 1.0; en.col = 0.0; en.wgt = 0.0; // Voxel() defaults, voxel.py:8
                    vbt_voxel_integrate(en.sdf, en.col, en.wgt, args[5 * i], args[5 * i + 1]);
                }
            } else {
                en.sdf = args[5 * i + 2]; en.col = args[5 * i + 3]; en.wgt = args[5 * i + 4];
            }
            if (ok && insert) {
                if (!(skip_check && i == begin) && vbt_needs_resize(t, load_factor)) {
                    st->next = i; st->code = VBT_NEED_RESIZE;
                    return;
                }
                ok = vbt_add(t, en, rb, re);
            }
        }
        if (!ok) {
            st->next = i; st->code = VBT_DANGLING;
            return;
        }
        if (out) { out[2 * i] = rb; out[2 * i + 1] = re; }
    }
    st->next = n; st->code = VBT_DONE;
}

// Resize, hash_fusion.py:425-436: re-adds the entries of `old` in bucket / slot order with their
// offsets cleared, through the same insert (which may ask for another resize).
__global__ void k_vbt_readd(BucketDev old, BucketDev t, double load_factor, long long begin, int skip_check, VbtStatus *st)
{
    if (blockIdx.x || threadIdx.x) return;
    const long long n = old.S * old.B;
    for (long long i = begin; i < n; ++i) {
        if (!old.occ[i]) continue;
        if (!(skip_check && i == begin) && vbt_needs_resize(t, load_factor)) {
            st->next = i; st->code = VBT_NEED_RESIZE;
            return;
        }
        Entry en;
        en.x = old.pos[3 * i]; en.y = old.pos[3 * i + 1]; en.z = old.pos[3 * i + 2];
        en.sdf = old.vox[3 * i]; en.col = old.vox[3 * i + 1]; en.wgt = old.vox[3 * i + 2];
        long long rb;
        int re;
        if (!vbt_add(t, en, rb, re)) {
            st->next = i; st->code = VBT_DANGLING;
            return;
        }
    }
    st->next = n; st->code = VBT_DONE;
}

// Read-only batch lookup: one thread per query (bucket, slot) or (-1, -1); (-2, -2) = dangling chain.
__global__ void k_vbt_lookup(BucketDev t, const int32_t *__restrict__ pos, long long n, long long *__restrict__ out,
                             double *__restrict__ vox_out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    long long rb;
    int re;
    if (!vbt_find(t, pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], rb, re)) { rb = -2; re = -2; }
    out[2 * i] = rb; out[2 * i + 1] = re;
    if (vox_out) {
        const long long s = rb >= 0 ? rb * t.B + re : -1;
        for (int k = 0; k < 3; ++k) vox_out[3 * i + k] = s >= 0 ? t.vox[3 * s + k] : 0.0;
    }
}

// counts[0] entries, [1] buckets holding more than one entry (hash_fusion.py:147-180)
__global__ void k_vbt_count(BucketDev t, unsigned long long *counts)
{
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int c = 0;
    if (b < t.S && t.exists[b]) c = vbt_count(t, b);
    unsigned long long e = (unsigned long long)c, m = c > 1 ? 1ull : 0ull;
    for (int o = 16; o; o >>= 1) {
        e += __shfl_xor_sync(0xffffffffu, e, o);
        m += __shfl_xor_sync(0xffffffffu, m, o);
    }
    if ((threadIdx.x & 31) == 0 && e) {
        atomicAdd(counts, e);
        if (m) atomicAdd(counts + 1, m);
    }
}

__global__ void k_vbt_init(BucketDev t)
{
    const long long n = t.S * t.B;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        t.occ[i] = 0;
        t.off[2 * i] = -1; t.off[2 * i + 1] = -1;
        if (i < t.S) t.exists[i] = 0;
        if (i == 0) *t.nonempty = 0;
    }
}

// ---------------------------------------------------------------------------------------------
// The pre-pass of the reference's HashTable.integrate (hash_fusion.py:108-132) as an ORDERED voxel
// stream: every voxel of the bounded volume in vox_coords order (x slowest, z fastest,
// hash_fusion.py:58-69) is projected with the reference's own chain (exact_pixel: dgemm-ordered
// transform, IEEE division, rint), and the voxels with valid_pts set are emitted in that order with
// their dist = min(1, depth_diff / trunc) (float64) and the folded colour of their pixel.  Order
// matters: it is the insertion order of the container, hence its slot layout.
// Two passes over the same test (count per 256-voxel chunk, exclusive scan, ordered emit).
#define VS_CHUNK 256
__device__ __forceinline__ bool voxel_valid(const FrameParams &f, const VolumeParams &v, long long i, int &x, int &y, int &z,
                                            int &pix, double &dist)
{
    z = (int)(i % v.dim_z);
    y = (int)((i / v.dim_z) % v.dim_y);
    x = (int)(i / ((long long)v.dim_z * v.dim_y));
    const double wx = v.lat_x[x], wy = v.lat_y[y], wz = v.lat_z[z];
    const double Z = __dadd_rn(__fma_rn(f.T[10], wz, __fma_rn(f.T[9], wy, __dmul_rn(f.T[8], wx))), f.T[11]);
    if (!(Z > 0.0)) return false; // hash_fusion.py:119
    pix = exact_pixel(f.T[0], f.T[1], f.T[2], f.T[3], f.T[4], f.T[5], f.T[6], f.T[7], wx, wy, wz, Z, f.fx, f.fy, f.cx, f.cy,
                      f.W, f.H);
    if (pix < 0) return false;
    const double d = f.depth[pix];
    const double diff = __dadd_rn(d, -Z);
    if (!(d > 0.0 && diff >= f.neg_trunc)) return false; // hash_fusion.py:124
    dist = fmin(1.0, __ddiv_rn(diff, f.trunc));           // hash_fusion.py:125
    return true;
}

template <bool EMIT>
__global__ void __launch_bounds__(VS_CHUNK)
k_voxel_stream(FrameParams f, VolumeParams v, const uint8_t *__restrict__ rgb, long long n_vox, unsigned *__restrict__ counts,
               const unsigned long long *__restrict__ offsets, int32_t *__restrict__ pos_out, double *__restrict__ args_out)
{
    __shared__ unsigned s_warp[VS_CHUNK / 32];
    const long long i = (long long)blockIdx.x * VS_CHUNK + threadIdx.x;
    int x = 0, y = 0, z = 0, pix = 0;
    double dist = 0.0;
    const bool valid = i < n_vox && voxel_valid(f, v, i, x, y, z, pix, dist);
    const unsigned ballot = __ballot_sync(0xffffffffu, valid);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_warp[warp] = __popc(ballot);
    __syncthreads();
    if (!EMIT) {
        if (threadIdx.x == 0) {
            unsigned c = 0;
            for (int w = 0; w < VS_CHUNK / 32; ++w) c += s_warp[w];
            counts[blockIdx.x] = c;
        }
        return;
    }
    if (!valid) return;
    unsigned before = __popc(ballot & ((1u << lane) - 1u));
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    const unsigned long long k = offsets[blockIdx.x] + before;
    pos_out[3 * k] = x; pos_out[3 * k + 1] = y; pos_out[3 * k + 2] = z;
    // hash_fusion.py:109-110: floor(B * 65536 + G * 256 + R) of the float32 image, an exact integer
    const unsigned r = rgb[3 * pix], g = rgb[3 * pix + 1], b = rgb[3 * pix + 2];
    args_out[5 * k] = dist;
    args_out[5 * k + 1] = (double)((b << 16) | (g << 8) | r);
    args_out[5 * k + 2] = 0.0; args_out[5 * k + 3] = 0.0; args_out[5 * k + 4] = 0.0;
}

// Exclusive scan of the chunk counts by one block (tiles of 1024, running carry).
__global__ void k_scan_counts(const unsigned *__restrict__ counts, unsigned long long *__restrict__ offsets, long long n,
                              unsigned long long *total)
{
    __shared__ unsigned long long s[1024];
    __shared__ unsigned long long carry;
    const int t = threadIdx.x;
    if (t == 0) carry = 0;
    __syncthreads();
    for (long long base = 0; base < n; base += 1024) {
        const long long idx = base + t;
        const unsigned long long val = idx < n ? counts[idx] : 0ull;
        s[t] = val;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            const unsigned long long add = t >= o ? s[t - o] : 0ull;
            __syncthreads();
            s[t] += add;
            __syncthreads();
        }
        if (idx < n) offsets[idx] = carry + s[t] - val;
        __syncthreads();
        if (t == 1023) carry += s[1023];
        __syncthreads();
    }
    if (t == 0) *total = carry;
}

// get_volume of the reference hash map (hash_fusion.py:442-463): tsdf = 1 / colour = 0 everywhere,
// then every entry that carries a voxel writes float32(sdf), float32(colour) at its position.
__global__ void k_vbt_fill_dense(float *tsdf, float *color, long long n)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        tsdf[i] = 1.0f;
        color[i] = 0.0f;
    }
}
__global__ void k_vbt_scatter_dense(BucketDev t, int dx, int dy, int dz, float *tsdf, float *color)
{
    const long long n = t.S * t.B;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (!t.occ[i] || t.vox[3 * i] != t.vox[3 * i]) continue; // empty slot, or an entry without a voxel (NaN)
        const int x = t.pos[3 * i], y = t.pos[3 * i + 1], z = t.pos[3 * i + 2];
        if (x < 0 || y < 0 || z < 0 || x >= dx || y >= dy || z >= dz) continue;
        const long long o = ((long long)x * dy + y) * dz + z;
        tsdf[o] = (float)t.vox[3 * i];
        color[o] = (float)t.vox[3 * i + 1];
    }
}

// ---------------------------------------------------------------------------------------------
struct tsdf_buckets {
    int device = 0;
    double load_factor = 0.75;
    BucketDev t{};
    cudaStream_t stream = nullptr;
    VbtStatus *d_status = nullptr, *h_status = nullptr;
    unsigned long long *d_counts = nullptr;
    long long resizes = 0;
    // per-frame voxel stream (vbt_integrate_frame): cached lattice tables and growable device buffers
    VolumeGeom geom{};
    bool geom_valid = false;
    uint8_t *d_rgb = nullptr;
    double *d_depth = nullptr;
    size_t cap_pix = 0;
    unsigned *d_chunk_counts = nullptr;
    unsigned long long *d_chunk_offsets = nullptr, *d_total = nullptr;
    size_t cap_chunks = 0;
    int32_t *d_pos = nullptr;
    double *d_args = nullptr;
    size_t cap_stream = 0;
};

static void table_free(BucketDev &t)
{
    if (t.pos) cudaFree(t.pos);
    if (t.off) cudaFree(t.off);
    if (t.vox) cudaFree(t.vox);
    if (t.occ) cudaFree(t.occ);
    if (t.exists) cudaFree(t.exists);
    if (t.nonempty) cudaFree(t.nonempty);
    t = BucketDev{};
}

static int table_alloc(BucketDev &t, long long S, int B, cudaStream_t s)
{
    t = BucketDev{};
    t.S = S;
    t.B = B;
    const size_t n = (size_t)S * B;
    cudaError_t e;
    if ((e = cudaMalloc(&t.pos, n * 3 * sizeof(int32_t))) != cudaSuccess ||
        (e = cudaMalloc(&t.off, n * 2 * sizeof(int32_t))) != cudaSuccess ||
        (e = cudaMalloc(&t.vox, n * 3 * sizeof(double))) != cudaSuccess ||
        (e = cudaMalloc(&t.occ, n)) != cudaSuccess || (e = cudaMalloc(&t.exists, (size_t)S)) != cudaSuccess ||
        (e = cudaMalloc(&t.nonempty, sizeof(long long))) != cudaSuccess) {
        cudaGetLastError();
        table_free(t);
        tsdf_set_error("bucket table of %lld buckets: device allocation failed: %s", S, cudaGetErrorString(e));
        return TSDF_ERR_OOM;
    }
    k_vbt_init<<<(unsigned)std::min<size_t>((n + 255) / 256, 4096), 256, 0, s>>>(t);
    CUDA_TRY(cudaGetLastError());
    return TSDF_OK;
}

static int read_status(tsdf_buckets *h, VbtStatus *out)
{
    CUDA_TRY(cudaMemcpyAsync(h->h_status, h->d_status, sizeof(VbtStatus), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    *out = *h->h_status;
    if (out->code == VBT_DANGLING) {
        tsdf_set_error("bucket table: a collision chain points at an empty slot (the reference raises "
                       "AttributeError on None here); stopped at index %lld", out->next);
        return TSDF_ERR_BAD_ARG;
    }
    return TSDF_OK;
}

// hash_fusion.py:414-437, including the recursion through add_hash_entry -> needs_resize.
static int do_resize(tsdf_buckets *h, int depth)
{
    if (depth > 40 || h->t.S > (1LL << 40)) {
        tsdf_set_error("bucket table: resize recursion does not terminate");
        return TSDF_ERR_BAD_ARG;
    }
    BucketDev old = h->t;
    BAD_ARG(old.S * 2 * old.B >= (1LL << 31), "bucket table too large to double (%lld buckets)", old.S);
    int rc = table_alloc(h->t, old.S * 2, old.B, h->stream);
    if (rc) {
        h->t = old;
        return rc;
    }
    h->resizes++;
    long long begin = 0;
    int skip = 0;
    while (true) {
        k_vbt_readd<<<1, 32, 0, h->stream>>>(old, h->t, h->load_factor, begin, skip, h->d_status);
        VbtStatus st;
        if ((rc = cudaGetLastError() != cudaSuccess ? TSDF_ERR_CUDA : read_status(h, &st))) break;
        if (st.code == VBT_DONE) break;
        begin = st.next;
        skip = 1;
        if ((rc = do_resize(h, depth + 1))) break;
    }
    table_free(old);
    return rc;
}

// Runs ops[0..n) (device arrays) in order, doubling the table whenever an insert finds it at its
// load factor (the kernel stops there; the reference does not re-check after resizing).
static int apply_device(tsdf_buckets *h, const int32_t *d_ops, const int32_t *d_pos, const double *d_args, long long n,
                        long long *d_out)
{
    long long begin = 0;
    int skip = 0;
    while (true) {
        k_vbt_apply<<<1, 32, 0, h->stream>>>(h->t, h->load_factor, d_ops, d_pos, d_args, begin, n, skip, d_out, h->d_status);
        if (cudaGetLastError() != cudaSuccess) {
            tsdf_set_error("bucket table: kernel launch failed");
            return TSDF_ERR_CUDA;
        }
        VbtStatus st;
        int rc = read_status(h, &st);
        if (rc || st.code == VBT_DONE) return rc;
        begin = st.next;
        skip = 1;
        if ((rc = do_resize(h, 0))) return rc;
    }
}

extern "C" int vbt_create(int64_t n_buckets, int bucket_size, double load_factor, int device, tsdf_buckets **out)
{
    BAD_ARG(!out, "out is NULL");
    *out = nullptr;
    BAD_ARG(n_buckets < 1 || n_buckets * (int64_t)bucket_size >= (1LL << 31), "n_buckets out of range");
    BAD_ARG(bucket_size < 2 || bucket_size > 32, "bucket_size must be in 2..32");
    BAD_ARG(!(load_factor > 0.0), "load_factor must be positive");
    int sm = 0;
    int rc = tsdf_select_device(device, &sm);
    if (rc) return rc;
    tsdf_buckets *h = new (std::nothrow) tsdf_buckets;
    BAD_ARG(!h, "out of host memory");
    h->device = device;
    h->load_factor = load_factor;
    cudaError_t e;
    if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaMalloc(&h->d_status, sizeof(VbtStatus))) != cudaSuccess ||
        (e = cudaMallocHost(&h->h_status, sizeof(VbtStatus))) != cudaSuccess ||
        (e = cudaMalloc(&h->d_counts, 2 * sizeof(unsigned long long))) != cudaSuccess) {
        tsdf_set_error("bucket table: %s", cudaGetErrorString(e));
        vbt_destroy(h);
        return TSDF_ERR_CUDA;
    }
    if ((rc = table_alloc(h->t, n_buckets, bucket_size, h->stream))) {
        vbt_destroy(h);
        return rc;
    }
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) {
        tsdf_set_error("bucket table initialisation failed");
        vbt_destroy(h);
        return TSDF_ERR_CUDA;
    }
    *out = h;
    return TSDF_OK;
}

extern "C" int vbt_destroy(tsdf_buckets *h)
{
    if (!h) return TSDF_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    table_free(h->t);
    if (h->d_status) cudaFree(h->d_status);
    if (h->h_status) cudaFreeHost(h->h_status);
    if (h->d_counts) cudaFree(h->d_counts);
    if (h->geom_valid) h->geom.release();
    void *bufs[] = {h->d_rgb, h->d_depth, h->d_chunk_counts, h->d_chunk_offsets, h->d_total, h->d_pos, h->d_args};
    for (void *b : bufs)
        if (b) cudaFree(b);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return TSDF_OK;
}

extern "C" int vbt_apply(tsdf_buckets *h, const int32_t *ops, const int64_t *positions, const double *args, int64_t n,
                         int64_t *out)
{
    BAD_ARG(!h || !out || (n > 0 && (!ops || !positions)), "NULL argument");
    BAD_ARG(n < 0, "negative count");
    if (!n) return TSDF_OK;
    CUDA_TRY(cudaSetDevice(h->device));
    std::vector<int32_t> p32((size_t)n * 3);
    std::vector<double> a((size_t)n * 5, 0.0);
    for (int64_t i = 0; i < n; ++i) {
        BAD_ARG(ops[i] < VBT_ADD || ops[i] > VBT_INTEGRATE, "op %lld: unknown op code %d", (long long)i, ops[i]);
        for (int k = 0; k < 3; ++k) {
            const int64_t v = positions[3 * i + k];
            BAD_ARG(v < -(1LL << 31) || v >= (1LL << 31), "op %lld: coordinate %lld does not fit 32 bits", (long long)i,
                    (long long)v);
            p32[3 * i + k] = (int32_t)v;
        }
        if (args)
            for (int k = 0; k < 5; ++k) a[5 * i + k] = args[5 * i + k];
        else
            a[5 * i + 2] = nan(""); // entry without a voxel (HashEntry(position, None, None))
    }
    int32_t *d_ops = nullptr, *d_pos = nullptr;
    double *d_args = nullptr;
    long long *d_out = nullptr;
    int rc = TSDF_OK;
    cudaError_t e;
    if ((e = cudaMalloc(&d_ops, n * sizeof(int32_t))) != cudaSuccess || (e = cudaMalloc(&d_pos, n * 3 * sizeof(int32_t))) != cudaSuccess ||
        (e = cudaMalloc(&d_args, n * 5 * sizeof(double))) != cudaSuccess || (e = cudaMalloc(&d_out, n * 2 * sizeof(long long))) != cudaSuccess) {
        tsdf_set_error("bucket table: %s", cudaGetErrorString(e));
        rc = TSDF_ERR_OOM;
    }
    if (!rc) {
        cudaMemcpyAsync(d_ops, ops, n * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream);
        cudaMemcpyAsync(d_pos, p32.data(), n * 3 * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream);
        cudaMemcpyAsync(d_args, a.data(), n * 5 * sizeof(double), cudaMemcpyHostToDevice, h->stream);
        cudaMemsetAsync(d_out, 0xFF, n * 2 * sizeof(long long), h->stream);
        rc = apply_device(h, d_ops, d_pos, d_args, n, d_out);
    }
    if (!rc) {
        e = cudaMemcpyAsync(out, d_out, n * 2 * sizeof(long long), cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) {
            tsdf_set_error("bucket table: %s", cudaGetErrorString(e));
            rc = TSDF_ERR_CUDA;
        }
    }
    cudaFree(d_ops); cudaFree(d_pos); cudaFree(d_args); cudaFree(d_out);
    return rc;
}

extern "C" int vbt_lookup(tsdf_buckets *h, const int64_t *positions, int64_t n, int64_t *out, double *voxels)
{
    BAD_ARG(!h || !out || (n > 0 && !positions), "NULL argument");
    BAD_ARG(n < 0, "negative count");
    if (!n) return TSDF_OK;
    CUDA_TRY(cudaSetDevice(h->device));
    std::vector<int32_t> p32((size_t)n * 3);
    for (int64_t i = 0; i < 3 * n; ++i) {
        BAD_ARG(positions[i] < -(1LL << 31) || positions[i] >= (1LL << 31), "coordinate does not fit 32 bits");
        p32[i] = (int32_t)positions[i];
    }
    int32_t *d_pos = nullptr;
    long long *d_out = nullptr;
    double *d_vox = nullptr;
    int rc = TSDF_OK;
    cudaError_t e;
    if ((e = cudaMalloc(&d_pos, n * 3 * sizeof(int32_t))) != cudaSuccess || (e = cudaMalloc(&d_out, n * 2 * sizeof(long long))) != cudaSuccess ||
        (voxels && (e = cudaMalloc(&d_vox, n * 3 * sizeof(double))) != cudaSuccess)) {
        tsdf_set_error("bucket table: %s", cudaGetErrorString(e));
        rc = TSDF_ERR_OOM;
    }
    if (!rc) {
        cudaMemcpyAsync(d_pos, p32.data(), n * 3 * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream);
        k_vbt_lookup<<<(unsigned)((n + 127) / 128), 128, 0, h->stream>>>(h->t, d_pos, n, d_out, d_vox);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, n * 2 * sizeof(long long), cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess && voxels)
            e = cudaMemcpyAsync(voxels, d_vox, n * 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) {
            tsdf_set_error("bucket table: %s", cudaGetErrorString(e));
            rc = TSDF_ERR_CUDA;
        }
    }
    cudaFree(d_pos); cudaFree(d_out); cudaFree(d_vox);
    if (!rc)
        for (int64_t i = 0; i < n; ++i)
            if (out[2 * i] == -2) {
                tsdf_set_error("bucket table: a collision chain points at an empty slot (query %lld)", (long long)i);
                return TSDF_ERR_BAD_ARG;
            }
    return rc;
}

extern "C" int vbt_double_table_size(tsdf_buckets *h)
{
    BAD_ARG(!h, "NULL handle");
    CUDA_TRY(cudaSetDevice(h->device));
    return do_resize(h, 0);
}

extern "C" int vbt_stats(tsdf_buckets *h, int64_t stats[6])
{
    BAD_ARG(!h || !stats, "NULL argument");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaMemsetAsync(h->d_counts, 0, 2 * sizeof(unsigned long long), h->stream));
    k_vbt_count<<<(unsigned)((h->t.S + 127) / 128), 128, 0, h->stream>>>(h->t, h->d_counts);
    CUDA_TRY(cudaGetLastError());
    unsigned long long c[2];
    long long nonempty = 0;
    CUDA_TRY(cudaMemcpyAsync(c, h->d_counts, sizeof(c), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaMemcpyAsync(&nonempty, h->t.nonempty, sizeof(nonempty), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    stats[0] = h->t.S;
    stats[1] = nonempty;
    stats[2] = (int64_t)c[0];
    stats[3] = (int64_t)c[1];
    stats[4] = h->resizes;
    stats[5] = h->t.B;
    return TSDF_OK;
}

extern "C" int vbt_dump(tsdf_buckets *h, int64_t capacity, int64_t *count, int64_t *rows, double *voxels)
{
    BAD_ARG(!h || !count, "NULL argument");
    CUDA_TRY(cudaSetDevice(h->device));
    const size_t n = (size_t)h->t.S * h->t.B;
    std::vector<uint8_t> occ(n);
    CUDA_TRY(cudaMemcpyAsync(occ.data(), h->t.occ, n, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    int64_t k = 0;
    for (size_t i = 0; i < n; ++i) k += occ[i];
    *count = k;
    if (!rows) return TSDF_OK;
    BAD_ARG(capacity < k, "capacity %lld < %lld stored entries", (long long)capacity, (long long)k);
    std::vector<int32_t> pos(n * 3), off(n * 2);
    std::vector<double> vox(voxels ? n * 3 : 0);
    CUDA_TRY(cudaMemcpyAsync(pos.data(), h->t.pos, n * 3 * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaMemcpyAsync(off.data(), h->t.off, n * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    if (voxels) CUDA_TRY(cudaMemcpyAsync(vox.data(), h->t.vox, n * 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    k = 0;
    for (size_t i = 0; i < n; ++i) {
        if (!occ[i]) continue;
        int64_t *r = rows + 7 * k;
        r[0] = (int64_t)(i / h->t.B); r[1] = (int64_t)(i % h->t.B);
        r[2] = pos[3 * i]; r[3] = pos[3 * i + 1]; r[4] = pos[3 * i + 2];
        r[5] = off[2 * i]; r[6] = off[2 * i + 1];
        if (voxels)
            for (int c = 0; c < 3; ++c) voxels[3 * k + c] = vox[3 * i + c];
        ++k;
    }
    return TSDF_OK;
}

// ---------------------------------------------------------------------------------------------
// The reference's HashTable.integrate, whole (hash_fusion.py:103-145): ordered voxel stream of the
// frame, then the per-voxel loop over it.  obs_weight does not exist here: the reference accepts it
// and never passes it on (:141, :145).
template <typename T>
static int grow(T **buf, size_t *cap, size_t want, const char *what)
{
    if (want <= *cap && *buf) return TSDF_OK;
    if (*buf) cudaFree(*buf);
    *buf = nullptr;
    *cap = 0;
    const size_t n = want + want / 4 + 64;
    if (cudaMalloc(buf, n * sizeof(T)) != cudaSuccess) {
        cudaGetLastError();
        tsdf_set_error("bucket table: device allocation of %s (%zu elements) failed", what, n);
        return TSDF_ERR_OOM;
    }
    *cap = n;
    return TSDF_OK;
}

extern "C" int vbt_integrate_frame(tsdf_buckets *h, const int64_t dim[3], const float origin[3], double voxel_size,
                                   double trunc, const uint8_t *rgb, const double *depth, int H, int W, const double K[9],
                                   const double Tinv[16], int64_t *n_valid)
{
    BAD_ARG(!h || !dim || !origin || !rgb || !depth, "NULL argument");
    CUDA_TRY(cudaSetDevice(h->device));
    int rc;
    if (!h->geom_valid || h->geom.dim[0] != dim[0] || h->geom.dim[1] != dim[1] || h->geom.dim[2] != dim[2] ||
        h->geom.origin[0] != origin[0] || h->geom.origin[1] != origin[1] || h->geom.origin[2] != origin[2] ||
        h->geom.vs != voxel_size || h->geom.trunc != trunc) {
        if (h->geom_valid) h->geom.release();
        h->geom_valid = false;
        if ((rc = h->geom.init(dim, origin, voxel_size, trunc))) return rc;
        h->geom_valid = true;
    }
    FrameParams f;
    CullParams c;
    if ((rc = make_frame_params(&f, &c, h->geom, H, W, K, Tinv, 1.0))) return rc;
    const size_t npix = (size_t)H * W;
    {   // frame buffers share one capacity
        size_t cap_a = h->cap_pix * 3, cap_b = h->cap_pix;
        if ((rc = grow(&h->d_rgb, &cap_a, npix * 3, "colour image")) || (rc = grow(&h->d_depth, &cap_b, npix, "depth image")))
            return rc;
        h->cap_pix = std::min(cap_a / 3, cap_b);
    }
    CUDA_TRY(cudaMemcpyAsync(h->d_rgb, rgb, npix * 3, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->d_depth, depth, npix * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    f.depth = h->d_depth;
    const VolumeParams v = h->geom.params();
    const long long n_vox = (long long)dim[0] * dim[1] * dim[2];
    const long long n_chunks = (n_vox + VS_CHUNK - 1) / VS_CHUNK;
    BAD_ARG(n_chunks >= (1LL << 31), "volume too large for the per-voxel stream (%lld voxels)", n_vox);
    {
        size_t cap_a = h->cap_chunks, cap_b = h->cap_chunks, one = h->d_total ? 1 : 0;
        if ((rc = grow(&h->d_chunk_counts, &cap_a, (size_t)n_chunks, "chunk counts")) ||
            (rc = grow(&h->d_chunk_offsets, &cap_b, (size_t)n_chunks, "chunk offsets")) ||
            (rc = grow(&h->d_total, &one, 1, "stream length")))
            return rc;
        h->cap_chunks = std::min(cap_a, cap_b);
    }
    k_voxel_stream<false><<<(unsigned)n_chunks, VS_CHUNK, 0, h->stream>>>(f, v, h->d_rgb, n_vox, h->d_chunk_counts, nullptr,
                                                                          nullptr, nullptr);
    CUDA_TRY(cudaGetLastError());
    k_scan_counts<<<1, 1024, 0, h->stream>>>(h->d_chunk_counts, h->d_chunk_offsets, n_chunks, h->d_total);
    CUDA_TRY(cudaGetLastError());
    unsigned long long total = 0;
    CUDA_TRY(cudaMemcpyAsync(&total, h->d_total, sizeof(total), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (n_valid) *n_valid = (int64_t)total;
    if (!total) return TSDF_OK;
    {
        size_t cap_a = h->cap_stream * 3, cap_b = h->cap_stream * 5;
        if ((rc = grow(&h->d_pos, &cap_a, (size_t)total * 3, "voxel stream positions")) ||
            (rc = grow(&h->d_args, &cap_b, (size_t)total * 5, "voxel stream values")))
            return rc;
        h->cap_stream = std::min(cap_a / 3, cap_b / 5);
    }
    k_voxel_stream<true><<<(unsigned)n_chunks, VS_CHUNK, 0, h->stream>>>(f, v, h->d_rgb, n_vox, nullptr, h->d_chunk_offsets,
                                                                         h->d_pos, h->d_args);
    CUDA_TRY(cudaGetLastError());
    return apply_device(h, nullptr, h->d_pos, h->d_args, (long long)total, nullptr);
}

extern "C" int vbt_to_dense(tsdf_buckets *h, const int64_t dim[3], float *tsdf, float *color)
{
    BAD_ARG(!h || !dim || !tsdf || !color, "NULL argument");
    BAD_ARG(dim[0] <= 0 || dim[1] <= 0 || dim[2] <= 0 || dim[0] > (1 << 20) || dim[1] > (1 << 20) || dim[2] > (1 << 20),
            "bad volume dimensions");
    CUDA_TRY(cudaSetDevice(h->device));
    const long long n = (long long)dim[0] * dim[1] * dim[2];
    float *d_t = nullptr, *d_c = nullptr;
    cudaError_t e;
    if ((e = cudaMalloc(&d_t, n * sizeof(float))) != cudaSuccess || (e = cudaMalloc(&d_c, n * sizeof(float))) != cudaSuccess) {
        cudaGetLastError();
        if (d_t) cudaFree(d_t);
        tsdf_set_error("bucket table: dense read-out of %lld voxels: %s", n, cudaGetErrorString(e));
        return TSDF_ERR_OOM;
    }
    k_vbt_fill_dense<<<2048, 256, 0, h->stream>>>(d_t, d_c, n);
    k_vbt_scatter_dense<<<2048, 256, 0, h->stream>>>(h->t, (int)dim[0], (int)dim[1], (int)dim[2], d_t, d_c);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(tsdf, d_t, n * sizeof(float), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(color, d_c, n * sizeof(float), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d_t);
    cudaFree(d_c);
    if (e != cudaSuccess) {
        tsdf_set_error("bucket table: dense read-out failed: %s", cudaGetErrorString(e));
        return TSDF_ERR_CUDA;
    }
    return TSDF_OK;
}
