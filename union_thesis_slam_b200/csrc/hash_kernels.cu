// hash_kernels.cu -- sm_100a kernels of the voxel-block hash map (K2-K6 of SURVEY.md section 2a).
//
// Layout in HBM
//   keys[S]    uint64  open-addressing table, S a power of two; key = bx<<42 | by<<21 | bz,
//                      empty = all ones.  Slot = ((bx*P1) ^ (by*P2) ^ (bz*P3)) mod S with the
//                      reference's primes (hash_fusion.py:24-26,182-190), linear probing.
//   vals[S]    uint32  heap index of the block record, or NONE while the key has no storage
//   entries[]  uint32  slots of all keys ever inserted (append-only), n_entries
//   heap[H]    uint32  free list of block records, heap_top = number of free records
//   blocks[H]  6 KB records, same format as the dense grid's bricks
//
// Per frame
//   (prep_frame builds the per-tile max-depth pyramid, dense_kernels.cu)
//   K2   walk_coarse    one warp per (8x8-pixel tile, 32 samples) walks the tile's cone from the
//                       camera to (max depth of the tile + trunc) and flags every 32^3-voxel
//                       super-block the conservatively dilated samples touch; duplicates inside the
//                       warp are removed with __match_any_sync before the atomicOr
//   K3   refine_insert  one warp per flagged super-block: blocks whose key is not yet known are
//                       tested with the dense grid's conservative visibility test and inserted
//                       lock-free (atomicCAS on the 64-bit key word).  Keys only -- no storage yet.
//   K4   compact        frustum/depth-range test of every table entry, warp ballot + prefix sum,
//                       one atomic per warp -> visible block list
//   K5   integrate      persistent CTAs, one block per iteration, the dense per-voxel rule.  A key
//                       without storage is probed first (geometry only); storage is taken from the
//                       free-list heap only if a voxel passes, so a block owns storage iff one of
//                       its voxels has been updated == the reference's entry set, block-wise.
#include "integrate.cuh"
#include "launch.h"

#define KEY_EMPTY 0xFFFFFFFFFFFFFFFFull
#define VAL_NONE 0xFFFFFFFFu
#define TILE_PX 8

__device__ __forceinline__ unsigned long long pack_key(int bx, int by, int bz)
{
    return ((unsigned long long)bx << 42) | ((unsigned long long)by << 21) | (unsigned long long)bz;
}
__device__ __forceinline__ void unpack_key(unsigned long long k, int &bx, int &by, int &bz)
{
    bx = (int)(k >> 42);
    by = (int)((k >> 21) & 0x1FFFFFull);
    bz = (int)(k & 0x1FFFFFull);
}
// hash_fusion.py:182-190; for a power-of-two table the floor-mod is a mask
__device__ __forceinline__ uint32_t slot_hash(int bx, int by, int bz, uint32_t mask)
{
    const long long h = ((long long)bx * 73856093LL) ^ ((long long)by * 19349669LL) ^ ((long long)bz * 83492791LL);
    return (uint32_t)h & mask;
}
// shard owner: decorrelated from slot_hash (must match dist.block_owner in Python)
__device__ __forceinline__ int owner_of(unsigned long long key, int world)
{
    return (int)(((key * 0x9E3779B97F4A7C15ull) >> 33) % (unsigned long long)world);
}

__device__ __forceinline__ void raise_error(Totals *tot, int code)
{
    atomicCAS(&tot->error, 0ull, (unsigned long long)(unsigned)(-code));
}

// Lookup; returns the slot or -1.
__device__ __forceinline__ int table_find(const HashDev &h, unsigned long long key, int bx, int by, int bz)
{
    uint32_t slot = slot_hash(bx, by, bz, h.mask);
    for (uint32_t probe = 0; probe <= h.mask; ++probe) {
        const unsigned long long k = h.keys[slot];
        if (k == key) return (int)slot;
        if (k == KEY_EMPTY) return -1;
        slot = (slot + 1) & h.mask;
    }
    return -1;
}

// Lock-free find-or-insert of a key (no storage).  atomicCAS on the 64-bit key word; the winner
// appends the slot to the entry list.  Keys are never removed, so a key found once stays put.
__device__ __forceinline__ void table_insert(const HashDev &h, unsigned long long key, int bx, int by, int bz,
                                             FrameScratch *fs, Totals *tot)
{
    uint32_t slot = slot_hash(bx, by, bz, h.mask);
    for (uint32_t probe = 0; probe <= h.mask; ++probe) {
        unsigned long long k = h.keys[slot];
        if (k == key) return;
        if (k == KEY_EMPTY) {
            k = atomicCAS(&h.keys[slot], KEY_EMPTY, key);
            if (k == KEY_EMPTY) {
                const int e = atomicAdd(h.n_entries, 1);
                h.entries[e] = slot; // e < S because every slot is appended at most once
                atomicAdd(&fs->new_keys, 1);
                if (probe > 0) atomicMax(&tot->max_probe, (unsigned long long)probe);
                return;
            }
            if (k == key) return; // another thread inserted the same key first
        }
        slot = (slot + 1) & h.mask;
    }
    raise_error(tot, TSDF_ERR_TABLE_FULL_CODE);
}

// ---------------------------------------------------------------------------------------------
// K2: cone walk per pixel tile over SUPER-BLOCKS (4x4x4 blocks = 32^3 voxels).  One warp per
// (tile, chunk of 32 samples); a sample stands for the slice of the tile's cone within half a step
// of its depth, and every super-block that slice can touch (axis-aligned box of the sample dilated
// by the along-ray half step and the cone half width, all conservative) is flagged in a per-frame
// bitmap.  Lanes that hit the same super-block elect one leader (__match_any_sync) for the atomicOr.
__global__ void k_walk_coarse(WalkParams w, HashDev h, const float *__restrict__ tile_dmax, uint32_t *sb_flags,
                              FrameScratch *fs)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int tile = warp / w.chunks, chunk = warp - tile * w.chunks;
    if (tile >= w.tiles_x * w.tiles_y) return;
    const float dmax = tile_dmax[tile];
    if (!(dmax > 0.0f)) return;
    const int tx = tile % w.tiles_x, ty = tile / w.tiles_x;
    const int u0 = tx * TILE_PX, v0 = ty * TILE_PX;
    const int u1 = min(u0 + TILE_PX, w.W) - 1, v1 = min(v0 + TILE_PX, w.H) - 1;
    // centre ray of the tile and its lateral half-extent per unit depth (+0.51 px: pixel rounding)
    const double dcx = (0.5 * (u0 + u1) - w.cx) / w.fx, dcy = (0.5 * (v0 + v1) - w.cy) / w.fy;
    const double hx = (0.5 * (u1 - u0) + 0.51) / fabs(w.fx), hy = (0.5 * (v1 - v0) + 0.51) / fabs(w.fy);
    const double lat = sqrt(hx * hx + hy * hy) * w.pose_scale * w.inv_vs; // cone half width, voxels per metre of depth
    // ray in voxel-index space: g(s) = e + s * d
    const double dd[3] = {(w.P[0] * dcx + w.P[1] * dcy + w.P[2]) * w.inv_vs,
                          (w.P[4] * dcx + w.P[5] * dcy + w.P[6]) * w.inv_vs,
                          (w.P[8] * dcx + w.P[9] * dcy + w.P[10]) * w.inv_vs};
    const double ee[3] = {(w.P[3] - w.ox) * w.inv_vs, (w.P[7] - w.oy) * w.inv_vs, (w.P[11] - w.oz) * w.inv_vs};
    const double hi[3] = {(double)w.dim_x, (double)w.dim_y, (double)w.dim_z};
    const double s_end = (double)dmax + w.trunc;
    // clip [0, s_end] to where the cone can touch the volume (index box dilated by the largest radius)
    double s0 = 0.0, s1 = s_end;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const double rmax = 0.5 * w.step * fabs(dd[a]) + lat * (s_end + w.step) + 2.0;
        const double lo_b = -rmax, hi_b = hi[a] + rmax;
        if (fabs(dd[a]) < 1e-12) {
            if (ee[a] < lo_b || ee[a] > hi_b) return;
        } else {
            double ta = (lo_b - ee[a]) / dd[a], tb = (hi_b - ee[a]) / dd[a];
            if (ta > tb) { const double t = ta; ta = tb; tb = t; }
            s0 = fmax(s0, ta);
            s1 = fmin(s1, tb);
        }
    }
    if (!(s0 <= s1)) return;
    const int nsamp = (int)ceil((s1 - s0) / w.step) + 1;
    int my_cand = 0;
    // chunks beyond w.chunks - 1 (never needed when the host bound holds) fold into the last one
    const int j_end = (chunk == w.chunks - 1) ? nsamp : min(nsamp, (chunk + 1) * 32);
    for (int j = chunk * 32 + lane; j < j_end; j += 32) {
        const double s = fmin(s0 + j * w.step, s_end);
        // every voxel centre a pixel of this tile can update at depth in [s - step/2, s + step/2] lies,
        // per axis, within (half step along the ray + cone half width) of the sample
        const double rl = lat * (s + 0.5 * w.step) * 1.001 + 0.02;
        int lo[3], hi_i[3];
        bool empty = false;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double g = ee[a] + s * dd[a], r = 0.5 * w.step * fabs(dd[a]) * 1.001 + rl;
            lo[a] = max((int)ceil(g - r), 0);
            hi_i[a] = min((int)floor(g + r), (int)hi[a] - 1);
            empty = empty || lo[a] > hi_i[a];
        }
        if (empty) continue;
        for (int sx = lo[0] >> 5; sx <= (hi_i[0] >> 5); ++sx)
            for (int sy = lo[1] >> 5; sy <= (hi_i[1] >> 5); ++sy)
                for (int sz = lo[2] >> 5; sz <= (hi_i[2] >> 5); ++sz) {
                    ++my_cand;
                    const uint32_t bit = (uint32_t)((sx * h.nsy + sy) * h.nsz + sz);
                    if ((sb_flags[bit >> 5] >> (bit & 31u)) & 1u) continue; // already flagged this frame
                    const unsigned peers = __match_any_sync(__activemask(), bit);
                    if (lane == __ffs(peers) - 1) atomicOr(sb_flags + (bit >> 5), 1u << (bit & 31u));
                }
    }
    __syncwarp();
    for (int o = 16; o; o >>= 1) my_cand += __shfl_xor_sync(0xffffffffu, my_cand, o);
    if (lane == 0 && my_cand) atomicAdd(&fs->cand_count, my_cand);
}

// K3: one warp per flagged super-block; each lane tests two of its 64 blocks.  Called with
// sb_flags == nullptr it also decides itself which super-blocks the frame can touch (sphere-vs-
// frustum + tile-depth-pyramid test of the bounded super-block grid) and K2 is not run at all.  A block whose key
// is already in the table (presence bitmap) costs one bit test; otherwise the conservative
// visibility test of the dense grid decides and the key is inserted lock-free (keys only).
__global__ void k_refine_insert(FrameParams f, CullParams c, HashDev h, uint32_t *__restrict__ sb_flags, int walked,
                                FrameScratch *fs, Totals *tot)
{
    const int sb = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (sb >= h.nsx * h.nsy * h.nsz) return;
    if (walked && !((sb_flags[sb >> 5] >> (sb & 31)) & 1u)) return;
    const int sz = sb % h.nsz, sy = (sb / h.nsz) % h.nsy, sx = sb / (h.nsz * h.nsy);
    c.max_depth = (double)frame_max_depth(fs, c.tag);
    if (!(c.max_depth > 0.0)) return;
    // grid-scan allocation (no ray walk ran): the conservative visibility test of the whole
    // super-block stands in for the walk's flag (warp-uniform)
    if (!walked) {
        if (!superblock_visible(f, c, sx, sy, sz)) return;
        // the frame's visible-super-block bitmap also gates the compaction pass (k_compact_visible)
        if (lane == 0) atomicOr(sb_flags + (sb >> 5), 1u << (sb & 31));
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int child = lane * 2 + k;
        const int bx = sx * 4 + (child >> 4), by = sy * 4 + ((child >> 2) & 3), bz = sz * 4 + (child & 3);
        if (bx >= h.nbx || by >= h.nby || bz >= h.nbz) continue;
        const uint32_t bit = (uint32_t)((bx * h.nby + by) * h.nbz + bz);
        if ((h.known[bit >> 5] >> (bit & 31u)) & 1u) continue;
        const unsigned long long key = pack_key(bx, by, bz);
        const bool ours = h.world == 1 || owner_of(key, h.world) == h.rank;
        if (ours) { // the owner decides first: another rank's block costs no visibility test at all
            if (!brick_visible(f, c, bx, by, bz)) continue;
            table_insert(h, key, bx, by, bz, fs, tot);
        }
        atomicOr(h.known + (bit >> 5), 1u << (bit & 31u)); // ours and inserted, or not ours: never test again
    }
}

// K4: visible-block compaction over the entry list.
__global__ void k_compact_visible(FrameParams f, CullParams c, HashDev h, const uint32_t *__restrict__ sb_flags,
                                  BrickRef *__restrict__ list, FrameScratch *fs)
{
    const int n = *h.n_entries;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if ((i & ~31) >= n) return;
    bool vis = false;
    int bx = 0, by = 0, bz = 0;
    uint32_t slot = 0;
    if (i < n) {
        slot = h.entries[i];
        unpack_key(h.keys[slot], bx, by, bz);
        c.max_depth = (double)frame_max_depth(fs, c.tag);
        // entries of super-blocks the frame cannot touch (bitmap written by the allocation pass) cost one bit test
        const int sb = ((bx >> 2) * h.nsy + (by >> 2)) * h.nsz + (bz >> 2);
        vis = c.max_depth > 0.0 && ((sb_flags[sb >> 5] >> (sb & 31)) & 1u) && brick_visible(f, c, bx, by, bz);
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, vis);
    if (!ballot) return;
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0) base = atomicAdd(&fs->vis_count, __popc(ballot));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (vis) {
        // list entry of the hash map: (bx | by << 16, bz, table slot, record index or NONE)
        BrickRef r;
        r.bx = bx | (by << 16); r.by = bz; r.bz = (int)slot; r.storage = h.vals[slot];
        list[base + __popc(ballot & ((1u << lane) - 1u))] = r;
    }
}

// K5: per-block integration.
#ifndef INTEGRATE_HASH_MIN_CTAS
#define INTEGRATE_HASH_MIN_CTAS INTEGRATE_MIN_CTAS
#endif
template <bool UNIT_W>
__global__ void __launch_bounds__(INTEGRATE_THREADS, INTEGRATE_HASH_MIN_CTAS)
k_integrate_hash(FrameParams f, VolumeParams v, HashDev h, const BrickRef *__restrict__ list, FrameScratch *fs, Totals *tot,
                 UpdateCounters uc)
{
    __shared__ uint32_t s_storage;
    // per-thread update counter in shared memory (a fire-and-forget add per column instead of a spilled register)
    __shared__ unsigned s_updates[INTEGRATE_THREADS];
    const int n = fs->vis_count;
    const int t = threadIdx.x;
    const int lx = t >> 3, ly = t & 7;
    s_updates[t] = 0;
    int b = blockIdx.x;
    int4 br_next = make_int4(0, 0, 0, 0);
    if (b < n) br_next = __ldg(reinterpret_cast<const int4 *>(list) + b);
    for (; b < n; b += gridDim.x) {
        const int4 br = br_next; // (bx | by << 16, bz, slot, storage)
        if (b + (int)gridDim.x < n) br_next = __ldg(reinterpret_cast<const int4 *>(list) + b + gridDim.x);
        const int gx = (br.x & 0xFFFF) * BRICK + lx, gy = (br.x >> 16) * BRICK + ly, gz0 = br.y * BRICK;
        const bool inb = gx < v.dim_x && gy < v.dim_y;
        uint32_t storage = (uint32_t)br.w; // CTA-uniform
        // compaction of this frame may have run while the previous frame's integrate kernel was
        // still allocating records: vals[] is authoritative for entries listed without storage
        if (storage == VAL_NONE) storage = h.vals[(uint32_t)br.z];
        if (storage == VAL_NONE) {
            // key without storage: geometry-only probe, allocate iff some voxel will be updated
            const unsigned would = inb ? integrate_column8<UNIT_W, true>(f, v, gx, gy, gz0, nullptr) : 0u;
            if (!__syncthreads_or(would != 0u)) continue;
            if (t == 0) {
                const int top = atomicSub(h.heap_top, 1) - 1; // atomic free-list pop
                uint32_t blk = VAL_NONE;
                if (top >= 0) {
                    blk = h.heap[top];
                    h.vals[(uint32_t)br.z] = blk;
                } else {
                    atomicAdd(h.heap_top, 1);
                    raise_error(tot, TSDF_ERR_HEAP_EXHAUSTED_CODE);
                }
                s_storage = blk;
            }
            __syncthreads();
            storage = s_storage;
            if (storage == VAL_NONE) continue;
            // initialise the record: tsdf = 1, weight = 0, colour = 0 (voxel.py:8)
            float *rec = h.blocks + (size_t)storage * BRICK_FLOATS + 8 * t;
            const float4 one = make_float4(1.f, 1.f, 1.f, 1.f), zero = make_float4(0.f, 0.f, 0.f, 0.f);
            __stcs(reinterpret_cast<float4 *>(rec), one);
            __stcs(reinterpret_cast<float4 *>(rec) + 1, one);
            __stcs(reinterpret_cast<float4 *>(rec + BRICK_VOX), zero);
            __stcs(reinterpret_cast<float4 *>(rec + BRICK_VOX) + 1, zero);
            __stcs(reinterpret_cast<float4 *>(rec + 2 * BRICK_VOX), zero);
            __stcs(reinterpret_cast<float4 *>(rec + 2 * BRICK_VOX) + 1, zero);
        }
        if (inb) {
            const unsigned nupd = integrate_column8<UNIT_W, false>(f, v, gx, gy, gz0,
                                                                    h.blocks + (size_t)storage * BRICK_FLOATS + 8 * t);
            if (nupd) atomicAdd(&s_updates[t], nupd);
        }
    }
    unsigned my_updates = s_updates[t];
    for (int o = 16; o; o >>= 1) my_updates += __shfl_xor_sync(0xffffffffu, my_updates, o);
    if ((t & 31) == 0) { // one private slot per warp: no atomics
        const int w = blockIdx.x * (INTEGRATE_THREADS / 32) + (t >> 5);
        uc.frame[w] = my_updates;
        uc.total[w] += my_updates;
    }
    if (blockIdx.x == 0 && t == 0) tot->last_visited = (unsigned long long)n;
}

// ---------------------------------------------------------------------------------------------
// Table initialisation: keys = EMPTY, vals = NONE, heap[i] = H-1-i (record 0 is handed out first).
__global__ void k_hash_init(HashDev h)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (size_t i = i0; i <= h.mask; i += stride) {
        h.keys[i] = KEY_EMPTY;
        h.vals[i] = VAL_NONE;
    }
    for (size_t i = i0; i < (size_t)h.heap_cap; i += stride) h.heap[i] = (uint32_t)(h.heap_cap - 1 - i);
    const size_t nwords = ((size_t)h.nbx * h.nby * h.nbz + 31) / 32;
    for (size_t i = i0; i < nwords; i += stride) h.known[i] = 0u;
    if (i0 == 0) {
        *h.n_entries = 0;
        *h.heap_top = h.heap_cap;
    }
}

__global__ void k_fill_dense(float *tsdf, float *wgt, float *col, size_t n)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (tsdf) tsdf[i] = 1.0f;
        if (wgt) wgt[i] = 0.0f;
        if (col) col[i] = 0.0f;
    }
}

// K6: scatter every block that owns storage into dense [X][Y][Z] volumes (hash_fusion.py:442-463).
__global__ void k_scatter_dense(HashDev h, int dim_x, int dim_y, int dim_z, float *tsdf, float *wgt, float *col)
{
    const int n = *h.n_entries;
    const int t = threadIdx.x, lx = t >> 3, ly = t & 7;
    for (int e = blockIdx.x; e < n; e += gridDim.x) {
        const uint32_t slot = h.entries[e];
        const uint32_t storage = h.vals[slot];
        if (storage == VAL_NONE) continue;
        int bx, by, bz;
        unpack_key(h.keys[slot], bx, by, bz);
        const int gx = bx * BRICK + lx, gy = by * BRICK + ly;
        if (gx >= dim_x || gy >= dim_y) continue;
        const float *rec = h.blocks + (size_t)storage * BRICK_FLOATS + 8 * t;
        const size_t base = ((size_t)gx * dim_y + gy) * dim_z + (size_t)bz * BRICK;
        for (int k = 0; k < BRICK && bz * BRICK + k < dim_z; ++k) {
            if (tsdf) tsdf[base + k] = rec[k];
            if (wgt) wgt[base + k] = rec[BRICK_VOX + k];
            if (col) col[base + k] = (float)__float_as_uint(rec[2 * BRICK_VOX + k]);
        }
    }
}

// Block export: compacts the blocks that own storage; colour converted to the float encoding.
__global__ void k_export_blocks(HashDev h, int capacity, int *count, int32_t *keys_out, float *payload)
{
    __shared__ int s_idx;
    const int n = *h.n_entries;
    for (int e = blockIdx.x; e < n; e += gridDim.x) {
        const uint32_t slot = h.entries[e];
        const uint32_t storage = h.vals[slot];
        if (storage == VAL_NONE) continue; // CTA-uniform
        if (threadIdx.x == 0) s_idx = atomicAdd(count, 1);
        __syncthreads();
        const int idx = s_idx;
        __syncthreads();
        if (idx >= capacity || !keys_out || !payload) continue;
        if (threadIdx.x == 0) {
            int bx, by, bz;
            unpack_key(h.keys[slot], bx, by, bz);
            keys_out[3 * idx] = bx; keys_out[3 * idx + 1] = by; keys_out[3 * idx + 2] = bz;
        }
        const float *rec = h.blocks + (size_t)storage * BRICK_FLOATS;
        float *dst = payload + (size_t)idx * BRICK_FLOATS;
        for (int i = threadIdx.x; i < BRICK_FLOATS; i += blockDim.x)
            dst[i] = i < 2 * BRICK_VOX ? rec[i] : (float)__float_as_uint(rec[i]);
    }
}

// Block removal (hash_fusion.py:322-393 at block granularity): the record returns to the free
// list (atomic push); the key stays in the table without storage, so probe chains never break.
__global__ void k_remove_blocks(HashDev h, const int32_t *keys, int n, int *removed)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int bx = keys[3 * i], by = keys[3 * i + 1], bz = keys[3 * i + 2];
    if (bx < 0 || by < 0 || bz < 0 || bx >= (1 << 21) || by >= (1 << 21) || bz >= (1 << 21)) return;
    const int slot = table_find(h, pack_key(bx, by, bz), bx, by, bz);
    if (slot < 0) return;
    const uint32_t storage = atomicExch(&h.vals[slot], VAL_NONE); // duplicates in `keys` free it once
    if (storage == VAL_NONE) return;
    h.heap[atomicAdd(h.heap_top, 1)] = storage;
    atomicAdd(removed, 1);
}

// Per-voxel view of the block map (reference hash_fusion.py:285-320: get_hash_entry / get_voxel on the
// table that integrate() fills).  coords are VOXEL indices; a voxel "has an entry" iff its block owns
// storage and its weight is > 0 (the reference creates an entry exactly when it first updates a voxel).
//   mode 0: look up -- out[i] = (tsdf, weight, colour as the reference's float encoding), found[i] = 0/1
//   mode 1: remove  -- a found voxel goes back to the no-entry state (tsdf 1, weight 0, colour 0;
//                      hash_fusion.py:322-393 at voxel granularity), found[i] reports whether it existed
__global__ void k_voxel_query(HashDev h, const int32_t *__restrict__ coords, int n, int mode, float *__restrict__ out,
                              uint8_t *__restrict__ found)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int x = coords[3 * i], y = coords[3 * i + 1], z = coords[3 * i + 2];
    bool ok = false;
    float t = 1.0f, w = 0.0f, c = 0.0f;
    if (x >= 0 && y >= 0 && z >= 0 && (x >> 3) < h.nbx && (y >> 3) < h.nby && (z >> 3) < h.nbz) {
        const int bx = x >> 3, by = y >> 3, bz = z >> 3;
        const int slot = table_find(h, pack_key(bx, by, bz), bx, by, bz);
        const uint32_t storage = slot >= 0 ? h.vals[slot] : VAL_NONE;
        if (storage != VAL_NONE) {
            float *rec = h.blocks + (size_t)storage * BRICK_FLOATS + ((x & 7) * 64 + (y & 7) * 8 + (z & 7));
            w = rec[BRICK_VOX];
            if (w > 0.0f) {
                ok = true;
                t = rec[0];
                c = (float)__float_as_uint(rec[2 * BRICK_VOX]);
                if (mode == 1) {
                    rec[0] = 1.0f;
                    rec[BRICK_VOX] = 0.0f;
                    rec[2 * BRICK_VOX] = 0.0f;
                }
            }
        }
    }
    if (out) { out[3 * i] = t; out[3 * i + 1] = ok ? w : 0.0f; out[3 * i + 2] = ok ? c : 0.0f; }
    if (found) found[i] = ok ? 1 : 0;
}

// Number of voxels with weight > 0 over all blocks that own storage == the reference map's entry
// count (hash_fusion.py:147-154), which is what the thesis' space comparison reports.
__global__ void k_count_voxels(HashDev h, unsigned long long *total)
{
    const int n = *h.n_entries;
    unsigned cnt = 0;
    for (int e = blockIdx.x; e < n; e += gridDim.x) {
        const uint32_t storage = h.vals[h.entries[e]];
        if (storage == VAL_NONE) continue;
        const float *wp = h.blocks + (size_t)storage * BRICK_FLOATS + BRICK_VOX;
        for (int i = threadIdx.x; i < BRICK_VOX; i += blockDim.x) cnt += wp[i] > 0.0f;
    }
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(total, (unsigned long long)cnt);
}

// Read-out gather (world > 1): block lists exported by every rank (keys [n][3], payload [n][3][512] with colour in
// the float encoding) are scattered into dense volumes that the caller pre-filled with the defaults.  One CTA per
// block, same voxel order as k_scatter_dense.
__global__ void k_scatter_block_list(const int32_t *__restrict__ keys, const float *__restrict__ payload, int n, int dim_x,
                                     int dim_y, int dim_z, float *tsdf, float *wgt, float *col)
{
    const int t = threadIdx.x, lx = t >> 3, ly = t & 7;
    for (int b = blockIdx.x; b < n; b += gridDim.x) {
        const int bx = keys[3 * b], by = keys[3 * b + 1], bz = keys[3 * b + 2];
        const int gx = bx * BRICK + lx, gy = by * BRICK + ly;
        if (gx >= dim_x || gy >= dim_y) continue;
        const float *rec = payload + (size_t)b * BRICK_FLOATS + 8 * t;
        const size_t base = ((size_t)gx * dim_y + gy) * dim_z + (size_t)bz * BRICK;
        for (int k = 0; k < BRICK && bz * BRICK + k < dim_z; ++k) {
            if (tsdf) tsdf[base + k] = rec[k];
            if (wgt) wgt[base + k] = rec[BRICK_VOX + k];
            if (col) col[base + k] = rec[2 * BRICK_VOX + k];
        }
    }
}

void launch_scatter_block_list(const int32_t *keys, const float *payload, int n, int dim_x, int dim_y, int dim_z, float *tsdf,
                               float *wgt, float *col, bool fill, int sm_count, cudaStream_t s)
{
    if (fill) k_fill_dense<<<sm_count * 8, 256, 0, s>>>(tsdf, wgt, col, (size_t)dim_x * dim_y * dim_z);
    if (n > 0) k_scatter_block_list<<<sm_count * 16, INTEGRATE_THREADS, 0, s>>>(keys, payload, n, dim_x, dim_y, dim_z, tsdf, wgt, col);
}

void launch_voxel_query(const HashDev &h, const int32_t *coords, int n, int mode, float *out, uint8_t *found, cudaStream_t s)
{
    if (n <= 0) return;
    k_voxel_query<<<(n + 127) / 128, 128, 0, s>>>(h, coords, n, mode, out, found);
}

void launch_count_voxels(const HashDev &h, unsigned long long *total, int sm_count, cudaStream_t s)
{
    k_count_voxels<<<sm_count * 8, 128, 0, s>>>(h, total);
}

// ---------------------------------------------------------------------------------------------
void launch_hash_init(const HashDev &h, int sm_count, cudaStream_t s) { k_hash_init<<<sm_count * 4, 256, 0, s>>>(h); }

void launch_walk_coarse(const WalkParams &w, const HashDev &h, const float *tile_dmax, uint32_t *sb_flags,
                        FrameScratch *fs, cudaStream_t s)
{
    const long long warps = (long long)w.tiles_x * w.tiles_y * w.chunks;
    k_walk_coarse<<<(unsigned)((warps * 32 + 127) / 128), 128, 0, s>>>(w, h, tile_dmax, sb_flags, fs);
}

void launch_refine_insert(const FrameParams &f, const CullParams &c, const HashDev &h, uint32_t *sb_flags, bool walked,
                          FrameScratch *fs, Totals *tot, cudaStream_t s)
{
    const long long warps = (long long)h.nsx * h.nsy * h.nsz;
    k_refine_insert<<<(unsigned)((warps * 32 + 127) / 128), 128, 0, s>>>(f, c, h, sb_flags, walked ? 1 : 0, fs, tot);
}

void launch_compact_visible(const FrameParams &f, const CullParams &c, const HashDev &h, const uint32_t *sb_flags,
                            int max_entries, BrickRef *list, FrameScratch *fs, cudaStream_t s)
{
    if (max_entries <= 0) return;
    k_compact_visible<<<(max_entries + 127) / 128, 128, 0, s>>>(f, c, h, sb_flags, list, fs);
}

int integrate_hash_ctas_per_sm()
{
    int n = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_integrate_hash<true>, INTEGRATE_THREADS, 0);
    return n > 0 ? n : 1;
}

void launch_integrate_hash(const FrameParams &f, const VolumeParams &v, const HashDev &h, const BrickRef *list,
                           FrameScratch *fs, Totals *tot, UpdateCounters uc, int grid, cudaStream_t s)
{
    if (f.unit_w)
        k_integrate_hash<true><<<grid, INTEGRATE_THREADS, 0, s>>>(f, v, h, list, fs, tot, uc);
    else
        k_integrate_hash<false><<<grid, INTEGRATE_THREADS, 0, s>>>(f, v, h, list, fs, tot, uc);
}

void launch_scatter_dense(const HashDev &h, int dim_x, int dim_y, int dim_z, float *tsdf, float *wgt, float *col,
                          int sm_count, cudaStream_t s)
{
    const size_t n = (size_t)dim_x * dim_y * dim_z;
    k_fill_dense<<<sm_count * 8, 256, 0, s>>>(tsdf, wgt, col, n);
    k_scatter_dense<<<sm_count * 16, INTEGRATE_THREADS, 0, s>>>(h, dim_x, dim_y, dim_z, tsdf, wgt, col);
}

void launch_export_blocks(const HashDev &h, int capacity, int *count, int32_t *keys_out, float *payload, int sm_count,
                          cudaStream_t s)
{
    k_export_blocks<<<sm_count * 8, 128, 0, s>>>(h, capacity, count, keys_out, payload);
}

void launch_remove_blocks(const HashDev &h, const int32_t *keys, int n, int *removed, cudaStream_t s)
{
    if (n <= 0) return;
    k_remove_blocks<<<(n + 255) / 256, 256, 0, s>>>(h, keys, n, removed);
}
