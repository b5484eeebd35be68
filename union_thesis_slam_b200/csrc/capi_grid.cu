// capi_grid.cu -- C ABI of the dense brick grid (include/tsdf_b200.h, tsdf_grid_*) and the host
// plumbing shared with the hash map (errors, frame constants, frame upload).
#include <algorithm>
#include <new>
#include <vector>

#include "host_common.h"

// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

void tsdf_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *tsdf_last_error(void) { return g_err; }

extern "C" int tsdf_device_info(int device, int *sm_count, int64_t *total_mem, int *cc_major, int *cc_minor)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        tsdf_set_error("no CUDA device available (this library has no CPU fallback)");
        return TSDF_ERR_NO_DEVICE;
    }
    BAD_ARG(device < 0 || device >= n, "device %d out of range (%d devices)", device, n);
    cudaDeviceProp p;
    CUDA_TRY(cudaGetDeviceProperties(&p, device));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (total_mem) *total_mem = (int64_t)p.totalGlobalMem;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    return TSDF_OK;
}

// ---------------------------------------------------------------------------------------------
int VolumeGeom::init(const int64_t dim_[3], const float origin_[3], double vs_, double trunc_)
{
    BAD_ARG(!dim_ || !origin_, "dim/origin is NULL");
    BAD_ARG(!(vs_ > 0.0) || !(trunc_ > 0.0), "voxel_size and trunc must be positive");
    int total_pad = 0;
    max_abs_coord = 0.0;
    for (int a = 0; a < 3; ++a) {
        BAD_ARG(dim_[a] <= 0 || dim_[a] > (1 << 20), "dim[%d]=%lld out of range", a, (long long)dim_[a]);
        dim[a] = dim_[a];
        origin[a] = origin_[a];
        nb[a] = (int)((dim_[a] + BRICK - 1) / BRICK);
        lat_off[a] = total_pad;
        total_pad += nb[a] * BRICK;
    }
    vs = vs_;
    trunc = trunc_;
    // grid_fusion.py:175-180: cam_pts = f32( f64(f32 origin) + vs * f64(f32(idx)) ), later upcast to f64
    std::vector<double> lat((size_t)total_pad);
    for (int a = 0; a < 3; ++a)
        for (int i = 0; i < nb[a] * BRICK; ++i) {
            volatile double prod = vs * (double)(float)i; // volatile: keep mul and add separately rounded
            const float w = (float)((double)origin[a] + prod);
            lat[(size_t)lat_off[a] + i] = (double)w;
            if (i < dim[a]) max_abs_coord = std::max(max_abs_coord, fabs((double)w));
        }
    CUDA_TRY(cudaMalloc(&d_lat, lat.size() * sizeof(double)));
    CUDA_TRY(cudaMemcpy(d_lat, lat.data(), lat.size() * sizeof(double), cudaMemcpyHostToDevice));
    return TSDF_OK;
}

void VolumeGeom::release()
{
    if (d_lat) cudaFree(d_lat);
    d_lat = nullptr;
}

VolumeParams VolumeGeom::params() const
{
    VolumeParams v;
    v.lat_x = d_lat + lat_off[0];
    v.lat_y = d_lat + lat_off[1];
    v.lat_z = d_lat + lat_off[2];
    v.dim_x = (int)dim[0]; v.dim_y = (int)dim[1]; v.dim_z = (int)dim[2];
    v.nbx = nb[0]; v.nby = nb[1]; v.nbz = nb[2];
    return v;
}

double VolumeGeom::brick_radius() const
{
    // centres span 7 voxels per axis -> half diagonal 3.5*sqrt(3)*vs; the lattice itself is
    // float32-rounded (<= 2^-24 relative per axis), covered by the second term
    return 3.5 * 1.7320508075688772 * vs * 1.0001 + ldexp(max_abs_coord + 1.0, -21);
}

int make_frame_params(FrameParams *f, CullParams *c, const VolumeGeom &g, int H, int W, const double K[9],
                      const double Tinv[16], double obs)
{
    BAD_ARG(!K || !Tinv, "K/Tinv is NULL");
    BAD_ARG(H <= 0 || W <= 0 || H > 4095 || W > 4095, "image size %dx%d unsupported (1..4095)", W, H);
    BAD_ARG(!(obs > 0.0) || !isfinite(obs), "obs_weight must be a positive finite number");
    for (int i = 0; i < 12; ++i) {
        BAD_ARG(!isfinite(Tinv[i]), "Tinv has a non-finite entry");
        f->T[i] = Tinv[i];
    }
    // grid_fusion.py:190-192: intrinsics are cast to float32 before use
    f->fx = (double)(float)K[0];
    f->fy = (double)(float)K[4];
    f->cx = (double)(float)K[2];
    f->cy = (double)(float)K[5];
    BAD_ARG(!isfinite(f->fx) || !isfinite(f->fy) || !(fabs(f->cx) <= 8192.0) || !(fabs(f->cy) <= 8192.0),
            "intrinsics out of range");
    for (int i = 0; i < 4; ++i) {
        f->sT[i] = f->fx * Tinv[i];
        f->sT[4 + i] = f->fy * Tinv[4 + i];
    }
    f->cxm = f->cx + 0.5 + FIX_MAGIC;
    f->cym = f->cy + 0.5 + FIX_MAGIC;
    f->trunc = g.trunc;
    f->neg_trunc = -g.trunc;
    f->obs = obs;
    f->obs_f = (float)obs;
    f->k23 = 0x4B000000u;
    f->fast_w = ((double)(float)obs == obs && obs >= 0.0009765625 && obs <= 1024.0) ? 1 : 0;
    f->unit_w = 0; // set by the handle, which knows the history of obs_weight
    f->H = H;
    f->W = W;
    f->npix = H * W;
    f->depth = nullptr;
    f->px = nullptr;
    for (int r = 0; r < 3; ++r)
        c->Tn[r] = sqrt(Tinv[4 * r] * Tinv[4 * r] + Tinv[4 * r + 1] * Tinv[4 * r + 1] + Tinv[4 * r + 2] * Tinv[4 * r + 2]) * 1.0000001;
    c->max_depth = 0.0;
    c->tag = 0; // set by the handle when the frame is prepared
    c->rho = g.brick_radius();
    c->vs = g.vs;
    c->ox = (double)g.origin[0];
    c->oy = (double)g.origin[1];
    c->oz = (double)g.origin[2];
    return TSDF_OK;
}

// ---------------------------------------------------------------------------------------------
int FrameIO::init()
{
    CUDA_TRY(cudaStreamCreateWithFlags(&compute, cudaStreamNonBlocking));
    for (int k = 0; k < MAX_SIDE; ++k) {
        CUDA_TRY(cudaStreamCreateWithFlags(&side[k], cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreate(&side_end[k]));
        for (int i = 0; i < FRAME_SLOTS; ++i) {
            CUDA_TRY(cudaEventCreateWithFlags(&side_int_done[k][i], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&side_slot_free[k][i], cudaEventDisableTiming));
        }
    }
    CUDA_TRY(cudaStreamCreateWithFlags(&copy, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&stat, cudaStreamNonBlocking));
    for (int i = 0; i < FRAME_SLOTS; ++i) CUDA_TRY(cudaEventCreateWithFlags(&stat_done[i], cudaEventDisableTiming));
    {   // preparation kernels are short and on the critical path of the next frame: highest priority
        int lo = 0, hi = 0;
        CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CUDA_TRY(cudaStreamCreateWithPriority(&pre, cudaStreamNonBlocking, hi));
    }
    for (int i = 0; i < FRAME_SLOTS; ++i) {
        CUDA_TRY(cudaEventCreateWithFlags(&h2d_done[i], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&slot_free[i], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&pre_done[i], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&int_done[i], cudaEventDisableTiming));
    }
    CUDA_TRY(cudaEventCreate(&k_begin));
    CUDA_TRY(cudaEventCreate(&k_end));
    for (int i = 0; i < 3; ++i) CUDA_TRY(cudaEventCreate(&k_mark[i]));
    CUDA_TRY(cudaEventCreateWithFlags(&ext_ready, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&ext_consumed, cudaEventDisableTiming));
    return TSDF_OK;
}

int FrameIO::ensure_px(int npix)
{
    if (npix <= cap_pix && px_slot[0]) return TSDF_OK;
    return ensure(npix);
}

int FrameIO::ensure(int npix)
{
    if (npix <= cap_pix) return TSDF_OK;
    CUDA_TRY(cudaStreamSynchronize(pre));
    CUDA_TRY(cudaStreamSynchronize(compute));
    { int rc_ = sync_side(); if (rc_) return rc_; }
    CUDA_TRY(cudaStreamSynchronize(copy));
    for (int i = 0; i < FRAME_SLOTS; ++i) {
        if (d_rgb[i]) cudaFree(d_rgb[i]);
        if (d_depth[i]) cudaFree(d_depth[i]);
        if (h_rgb[i]) cudaFreeHost(h_rgb[i]);
        if (h_depth[i]) cudaFreeHost(h_depth[i]);
        if (d_depth16[i]) cudaFree(d_depth16[i]);
        d_depth16[i] = nullptr;
        d_rgb[i] = nullptr; d_depth[i] = nullptr; h_rgb[i] = nullptr; h_depth[i] = nullptr;
    }
    for (int i = 0; i < MAX_PREP; ++i) {
        if (px_slot[i]) cudaFree(px_slot[i]);
        px_slot[i] = nullptr;
    }
    d_px = nullptr;
    cap_pix = 0;
    for (int i = 0; i < FRAME_SLOTS; ++i) {
        CUDA_TRY(cudaMalloc(&d_rgb[i], (size_t)npix * 3));
        CUDA_TRY(cudaMalloc(&d_depth[i], (size_t)npix * sizeof(double)));
        CUDA_TRY(cudaMallocHost(&h_rgb[i], (size_t)npix * 3));
        CUDA_TRY(cudaMallocHost(&h_depth[i], (size_t)npix * sizeof(double)));
        CUDA_TRY(cudaMalloc(&d_depth16[i], (size_t)npix * sizeof(uint16_t)));
    }
    for (int i = 0; i < FRAME_SLOTS * subs; ++i) CUDA_TRY(cudaMalloc(&px_slot[i], ((size_t)npix + 1) * sizeof(FramePixel)));
    cap_pix = npix;
    return TSDF_OK;
}

int FrameIO::set_subs(int n)
{
    if (n <= subs) return TSDF_OK;
    // more prepared frames per slot: drain everything and let ensure()/ensure_tiles() reallocate
    CUDA_TRY(cudaStreamSynchronize(pre));
    CUDA_TRY(cudaStreamSynchronize(compute));
    { int rc_ = sync_side(); if (rc_) return rc_; }
    CUDA_TRY(cudaStreamSynchronize(copy));
    subs = n;
    const int npix = cap_pix;
    if (npix > 0) {
        for (int i = 0; i < FRAME_SLOTS * subs; ++i)
            if (!px_slot[i]) CUDA_TRY(cudaMalloc(&px_slot[i], ((size_t)npix + 1) * sizeof(FramePixel)));
    }
    tiles_H = tiles_W = 0; // scratch is reallocated for all subs by the next ensure_tiles()
    return TSDF_OK;
}

int FrameIO::release_upload_after_prep()
{
    if (cur >= 0) {
        CUDA_TRY(cudaEventRecord(slot_free[cur], pre));
        cur = -1;
    }
    if (ext) { // the caller's producer stream may reuse the frame's buffers from here on
        CUDA_TRY(cudaEventRecord(ext_consumed, pre));
        CUDA_TRY(cudaStreamWaitEvent(ext, ext_consumed, 0));
        ext = nullptr;
    }
    return TSDF_OK;
}

int FrameIO::adopt(cudaStream_t producer)
{
    CUDA_TRY(cudaEventRecord(ext_ready, producer));
    CUDA_TRY(cudaStreamWaitEvent(pre, ext_ready, 0));
    ext = producer;
    return TSDF_OK;
}

int FrameIO::convert_u16(const uint16_t *d_raw, int npix, double depth_scale, int invalid_raw, const double **d_depth_out)
{
    int rc = ensure(npix);
    if (rc) return rc;
    const int s = next;
    next = (next + 1) % FRAME_SLOTS;
    // the slot's float64 image was last read by the preparation kernel FRAME_SLOTS frames ago
    CUDA_TRY(cudaStreamWaitEvent(pre, slot_free[s], 0));
    for (int k = 0; k < side_used; ++k) CUDA_TRY(cudaStreamWaitEvent(pre, side_slot_free[k][s], 0));
    launch_depth_u16(d_raw, d_depth[s], npix, depth_scale, invalid_raw, pre);
    CUDA_TRY(cudaGetLastError());
    cur = s;
    *d_depth_out = d_depth[s];
    return TSDF_OK;
}

int FrameIO::ensure_tiles(int H, int W)
{
    if (scratch_slot[0] && H == tiles_H && W == tiles_W) return TSDF_OK;
    CUDA_TRY(cudaStreamSynchronize(pre));
    CUDA_TRY(cudaStreamSynchronize(compute));
    { int rc_ = sync_side(); if (rc_) return rc_; }
    CUDA_TRY(cudaStreamSynchronize(stat)); // a counter snapshot may still be reading the scratch that is about to go
    size_t off[4], total = 256; // FrameScratch lives in the first 256 bytes
    TilePyramid p{};
    for (int l = 1; l <= 3; ++l) {
        p.tx[l] = (W + (8 << l) - 1) / (8 << l);
        p.ty[l] = (H + (8 << l) - 1) / (8 << l);
        off[l] = total;
        total += (((size_t)p.tx[l] * p.ty[l] * sizeof(unsigned long long)) + 255) / 256 * 256; // tagged 64-bit maxima
    }
    p.tx[0] = (W + 7) / 8;
    p.ty[0] = (H + 7) / 8;
    off[0] = total;
    total += (size_t)p.tx[0] * p.ty[0] * sizeof(float);
    for (int i = 0; i < MAX_PREP; ++i) {
        if (scratch_slot[i]) cudaFree(scratch_slot[i]);
        scratch_slot[i] = nullptr;
    }
    for (int i = 0; i < FRAME_SLOTS * subs; ++i) {
        CUDA_TRY(cudaMalloc(&scratch_slot[i], total));
        CUDA_TRY(cudaMemset(scratch_slot[i], 0, total)); // once: tag 0 is never used, so everything reads as "no depth"
        pyr_slot[i] = p;
        pyr_slot[i].level0 = reinterpret_cast<float *>(scratch_slot[i] + off[0]);
        for (int l = 1; l < 4; ++l) pyr_slot[i].up[l - 1] = reinterpret_cast<unsigned long long *>(scratch_slot[i] + off[l]);
    }
    fs = nullptr;
    tiles_H = H;
    tiles_W = W;
    return TSDF_OK;
}

static bool is_pinned_host(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

int FrameIO::upload(const uint8_t *rgb, const double *depth, int npix, const uint8_t **d_rgb_out,
                    const double **d_depth_out)
{
    int rc = ensure(npix);
    if (rc) return rc;
    const int s = next;
    next = (next + 1) % FRAME_SLOTS;
    const bool pinned = is_pinned_host(rgb) && is_pinned_host(depth);
    const uint8_t *src_rgb = rgb;
    const double *src_depth = depth;
    if (!pinned) {
        // the staging slot was last read by the H2D copy issued FRAME_SLOTS frames ago
        CUDA_TRY(cudaEventSynchronize(h2d_done[s]));
        memcpy(h_rgb[s], rgb, (size_t)npix * 3);
        memcpy(h_depth[s], depth, (size_t)npix * sizeof(double));
        src_rgb = h_rgb[s];
        src_depth = h_depth[s];
    }
    // the device slot was last read by the kernels of the frame FRAME_SLOTS calls ago
    CUDA_TRY(cudaStreamWaitEvent(copy, slot_free[s], 0));
    for (int k = 0; k < side_used; ++k) CUDA_TRY(cudaStreamWaitEvent(copy, side_slot_free[k][s], 0));
    CUDA_TRY(cudaMemcpyAsync(d_depth[s], src_depth, (size_t)npix * sizeof(double), cudaMemcpyHostToDevice, copy));
    CUDA_TRY(cudaMemcpyAsync(d_rgb[s], src_rgb, (size_t)npix * 3, cudaMemcpyHostToDevice, copy));
    CUDA_TRY(cudaEventRecord(h2d_done[s], copy));
    CUDA_TRY(cudaStreamWaitEvent(pre, h2d_done[s], 0));
    // the caller's pinned buffers are read by the copy engine until h2d_done: the call must not
    // return before that (finish_upload), but the frame's kernels can be issued meanwhile
    borrowed = pinned ? s : -1;
    cur = s;
    *d_rgb_out = d_rgb[s];
    *d_depth_out = d_depth[s];
    return TSDF_OK;
}

int FrameIO::finish_upload()
{
    if (borrowed < 0) return TSDF_OK;
    const int s = borrowed;
    borrowed = -1;
    CUDA_TRY(cudaEventSynchronize(h2d_done[s]));
    return TSDF_OK;
}

int FrameIO::upload_u16(const uint8_t *rgb, const uint16_t *depth_raw, int npix, double depth_scale, int invalid_raw,
                        const uint8_t **d_rgb_out, const double **d_depth_out)
{
    int rc = ensure(npix);
    if (rc) return rc;
    const int s = next;
    next = (next + 1) % FRAME_SLOTS;
    const bool pinned = is_pinned_host(rgb) && is_pinned_host(depth_raw);
    const uint8_t *src_rgb = rgb;
    const uint16_t *src_depth = depth_raw;
    if (!pinned) {
        CUDA_TRY(cudaEventSynchronize(h2d_done[s]));
        memcpy(h_rgb[s], rgb, (size_t)npix * 3);
        memcpy(h_depth[s], depth_raw, (size_t)npix * sizeof(uint16_t)); // the float64 staging buffer is large enough
        src_rgb = h_rgb[s];
        src_depth = reinterpret_cast<const uint16_t *>(h_depth[s]);
    }
    CUDA_TRY(cudaStreamWaitEvent(copy, slot_free[s], 0));
    for (int k = 0; k < side_used; ++k) CUDA_TRY(cudaStreamWaitEvent(copy, side_slot_free[k][s], 0));
    CUDA_TRY(cudaMemcpyAsync(d_depth16[s], src_depth, (size_t)npix * sizeof(uint16_t), cudaMemcpyHostToDevice, copy));
    CUDA_TRY(cudaMemcpyAsync(d_rgb[s], src_rgb, (size_t)npix * 3, cudaMemcpyHostToDevice, copy));
    CUDA_TRY(cudaEventRecord(h2d_done[s], copy));
    CUDA_TRY(cudaStreamWaitEvent(pre, h2d_done[s], 0));
    borrowed = pinned ? s : -1;
    launch_depth_u16(d_depth16[s], d_depth[s], npix, depth_scale, invalid_raw, pre);
    CUDA_TRY(cudaGetLastError());
    cur = s;
    *d_rgb_out = d_rgb[s];
    *d_depth_out = d_depth[s];
    return TSDF_OK;
}

int FrameIO::begin_kernels()
{
    slot = (slot + 1) % FRAME_SLOTS;
    d_px = px_slot[slot];
    fs = reinterpret_cast<FrameScratch *>(scratch_slot[slot]);
    pyr = pyr_slot[slot];
    // the integrate kernels of the frame FRAME_SLOTS calls ago were the last readers of this slot
    CUDA_TRY(cudaStreamWaitEvent(pre, int_done[slot], 0));
    for (int k = 0; k < side_used; ++k) CUDA_TRY(cudaStreamWaitEvent(pre, side_int_done[k][slot], 0));
    if (stat_used) CUDA_TRY(cudaStreamWaitEvent(pre, stat_done[slot], 0)); // a counter snapshot may still read the slot
    if (timing) CUDA_TRY(cudaEventRecord(k_begin, pre));
    n_marks = 0; // no memset: k_prep_frame zeroes the slot's counters, the maxima are tagged per frame
    return TSDF_OK;
}

int FrameIO::prepared()
{
    CUDA_TRY(cudaEventRecord(pre_done[slot], pre));
    CUDA_TRY(cudaStreamWaitEvent(compute, pre_done[slot], 0));
    for (int k = 0; k < n_side; ++k) CUDA_TRY(cudaStreamWaitEvent(side[k], pre_done[slot], 0));
    return TSDF_OK;
}

int FrameIO::stat_begin()
{
    CUDA_TRY(cudaStreamWaitEvent(stat, int_done[slot], 0));
    for (int k = 0; k < side_used; ++k) CUDA_TRY(cudaStreamWaitEvent(stat, side_int_done[k][slot], 0));
    return TSDF_OK;
}

int FrameIO::stat_end()
{
    CUDA_TRY(cudaEventRecord(stat_done[slot], stat));
    stat_used = true;
    return TSDF_OK;
}

int FrameIO::join()
{
    for (int k = 0; k < MAX_SIDE; ++k) {
        if (!side_pending[k]) continue;
        // the last event recorded on side[k]
        CUDA_TRY(cudaStreamWaitEvent(compute, side_int_done[k][side_slot_last], 0));
        side_pending[k] = false;
    }
    return TSDF_OK;
}

int FrameIO::sync_side()
{
    for (int k = 0; k < MAX_SIDE; ++k) {
        if (side[k]) CUDA_TRY(cudaStreamSynchronize(side[k]));
        side_pending[k] = false;
    }
    return TSDF_OK;
}

int FrameIO::mark()
{
    if (n_marks < 3) CUDA_TRY(cudaEventRecord(k_mark[n_marks++], compute));
    return TSDF_OK;
}

int FrameIO::last_stage_ms(float ms[4], int *n)
{
    BAD_ARG(!ms || !n, "NULL argument");
    BAD_ARG(!timed, "no frame has been timed (switch profiling on before the frame)");
    CUDA_TRY(cudaEventSynchronize(k_end));
    cudaEvent_t ev[5];
    int k = 0;
    ev[k++] = k_begin;
    for (int i = 0; i < n_marks; ++i) ev[k++] = k_mark[i];
    ev[k++] = k_end;
    for (int i = 0; i + 1 < k; ++i) CUDA_TRY(cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]));
    *n = k - 1;
    return TSDF_OK;
}

int FrameIO::end_kernels()
{
    if (timing) CUDA_TRY(cudaEventRecord(k_end, compute));
    CUDA_TRY(cudaEventRecord(int_done[slot], compute));
    ended_side = n_side;
    if (n_side > side_used) side_used = n_side;
    if (n_side) side_slot_last = slot;
    for (int k = 0; k < n_side; ++k) {
        if (timing) CUDA_TRY(cudaEventRecord(side_end[k], side[k]));
        CUDA_TRY(cudaEventRecord(side_int_done[k][slot], side[k]));
        side_pending[k] = true;
    }
    timed = timing;
    if (cur >= 0) {
        CUDA_TRY(cudaEventRecord(slot_free[cur], compute));
        for (int k = 0; k < n_side; ++k) CUDA_TRY(cudaEventRecord(side_slot_free[k][cur], side[k]));
        cur = -1;
    }
    return TSDF_OK;
}

int FrameIO::last_ms(float *ms)
{
    BAD_ARG(!ms, "ms is NULL");
    BAD_ARG(!timed, "no frame has been timed (switch profiling on before the frame)");
    CUDA_TRY(cudaEventSynchronize(k_end));
    CUDA_TRY(cudaEventElapsedTime(ms, k_begin, k_end));
    for (int k = 0; k < ended_side; ++k) { // sibling integrate kernels of the frame: whichever ended last
        float ms2 = 0.f;
        CUDA_TRY(cudaEventSynchronize(side_end[k]));
        CUDA_TRY(cudaEventElapsedTime(&ms2, k_begin, side_end[k]));
        if (ms2 > *ms) *ms = ms2;
    }
    return TSDF_OK;
}

void FrameIO::release()
{
    if (pre) cudaStreamSynchronize(pre);
    if (compute) cudaStreamSynchronize(compute);
    sync_side();
    if (copy) cudaStreamSynchronize(copy);
    if (stat) cudaStreamSynchronize(stat);
    for (int i = 0; i < FRAME_SLOTS; ++i) {
        if (stat_done[i]) cudaEventDestroy(stat_done[i]);
        for (int k = 0; k < MAX_SIDE; ++k) {
            if (side_int_done[k][i]) cudaEventDestroy(side_int_done[k][i]);
            if (side_slot_free[k][i]) cudaEventDestroy(side_slot_free[k][i]);
        }
        if (d_depth16[i]) cudaFree(d_depth16[i]);
        if (pre_done[i]) cudaEventDestroy(pre_done[i]);
        if (int_done[i]) cudaEventDestroy(int_done[i]);
        if (d_rgb[i]) cudaFree(d_rgb[i]);
        if (d_depth[i]) cudaFree(d_depth[i]);
        if (h_rgb[i]) cudaFreeHost(h_rgb[i]);
        if (h_depth[i]) cudaFreeHost(h_depth[i]);
        if (h2d_done[i]) cudaEventDestroy(h2d_done[i]);
        if (slot_free[i]) cudaEventDestroy(slot_free[i]);
    }
    for (int i = 0; i < MAX_PREP; ++i) {
        if (px_slot[i]) cudaFree(px_slot[i]);
        if (scratch_slot[i]) cudaFree(scratch_slot[i]);
    }
    for (int i = 0; i < 3; ++i)
        if (k_mark[i]) cudaEventDestroy(k_mark[i]);
    if (k_begin) cudaEventDestroy(k_begin);
    if (k_end) cudaEventDestroy(k_end);
    if (ext_ready) cudaEventDestroy(ext_ready);
    if (ext_consumed) cudaEventDestroy(ext_consumed);
    if (compute) cudaStreamDestroy(compute);
    for (int k = 0; k < MAX_SIDE; ++k) {
        if (side_end[k]) cudaEventDestroy(side_end[k]);
        if (side[k]) cudaStreamDestroy(side[k]);
    }
    if (copy) cudaStreamDestroy(copy);
    if (stat) cudaStreamDestroy(stat);
    if (pre) cudaStreamDestroy(pre);
}

int CounterStore::init(int n_warps)
{
    n = n_warps;
    for (int i = 0; i < FRAME_SLOTS; ++i) {
        CUDA_TRY(cudaMalloc(&frame[i], (size_t)n * sizeof(unsigned long long)));
        CUDA_TRY(cudaMemset(frame[i], 0, (size_t)n * sizeof(unsigned long long)));
    }
    CUDA_TRY(cudaMalloc(&total, (size_t)n * sizeof(unsigned long long)));
    CUDA_TRY(cudaMemset(total, 0, (size_t)n * sizeof(unsigned long long)));
    CUDA_TRY(cudaMalloc(&d_stats, sizeof(StatsBlock)));
    CUDA_TRY(cudaMallocHost(&h_stats, sizeof(StatsBlock)));
    CUDA_TRY(cudaMalloc(&d_ring, STATS_RING * sizeof(StatsBlock)));
    CUDA_TRY(cudaMallocHost(&h_ring, STATS_RING * sizeof(StatsBlock)));
    for (int i = 0; i < STATS_RING; ++i) CUDA_TRY(cudaEventCreateWithFlags(&ring_done[i], cudaEventDisableTiming));
    return TSDF_OK;
}

int CounterStore::post(int slot, const FrameScratch *fs, const Totals *tot, const int *n_entries, const int *heap_top,
                       cudaStream_t s, int64_t frames, int64_t *ticket)
{
    const int r = (int)(next_ticket % STATS_RING);
    launch_collect_stats(view(slot), n, fs, tot, n_entries, heap_top, d_ring + r, s, false);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(h_ring + r, d_ring + r, sizeof(StatsBlock), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaEventRecord(ring_done[r], s));
    ring_frames[r] = frames;
    *ticket = next_ticket++;
    return TSDF_OK;
}

int CounterStore::wait(int64_t ticket, const StatsBlock **out, int64_t *frames)
{
    BAD_ARG(ticket < 0 || ticket >= next_ticket, "unknown stats ticket %lld", (long long)ticket);
    BAD_ARG(next_ticket - ticket > STATS_RING, "stats ticket %lld is too old (only the last %d snapshots are kept)",
            (long long)ticket, STATS_RING);
    const int r = (int)(ticket % STATS_RING);
    CUDA_TRY(cudaEventSynchronize(ring_done[r]));
    *out = h_ring + r;
    *frames = ring_frames[r];
    return TSDF_OK;
}

void CounterStore::release()
{
    for (int i = 0; i < FRAME_SLOTS; ++i)
        if (frame[i]) cudaFree(frame[i]);
    if (total) cudaFree(total);
    if (d_stats) cudaFree(d_stats);
    if (h_stats) cudaFreeHost(h_stats);
    if (d_ring) cudaFree(d_ring);
    if (h_ring) cudaFreeHost(h_ring);
    for (int i = 0; i < STATS_RING; ++i)
        if (ring_done[i]) cudaEventDestroy(ring_done[i]);
}

int CounterStore::collect(int slot, const FrameScratch *fs, const Totals *tot, const int *n_entries,
                          const int *heap_top, cudaStream_t s, const StatsBlock **out)
{
    launch_collect_stats(view(slot), n, fs, tot, n_entries, heap_top, d_stats, s);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(h_stats, d_stats, sizeof(StatsBlock), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    *out = h_stats;
    return TSDF_OK;
}

int tsdf_select_device(int device, int *sm_count)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        tsdf_set_error("no CUDA device available (this library has no CPU fallback)");
        return TSDF_ERR_NO_DEVICE;
    }
    BAD_ARG(device < 0 || device >= n, "device %d out of range (%d devices)", device, n);
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp p;
    CUDA_TRY(cudaGetDeviceProperties(&p, device));
    if (p.major < 10) {
        tsdf_set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, p.major, p.minor);
        return TSDF_ERR_NO_DEVICE;
    }
    *sm_count = p.multiProcessorCount;
    return TSDF_OK;
}

// ---------------------------------------------------------------------------------------------
struct tsdf_grid {
    int device = 0, sm_count = 0, rank = 0, world = 1;
    VolumeGeom geom;
    FrameIO io;
    int nlbx = 0;              // local brick sheets
    int64_t n_x_local = 0;     // local voxel rows
    size_t nbricks = 0;
    float *bricks = nullptr;
    BrickRef *vis_list[FRAME_SLOTS] = {}; // one per frame slot: classes x [class_cap]
    int classes = 2;           // independent classes of bricks (1..DENSE_MAX_CLASSES), one integrate kernel each
    size_t class_cap = 0;      // capacity of one class's list inside vis_list[]
    Totals *tot = nullptr;
    int integrate_grid = 0;
    int64_t frames = 0;
    CounterStore counters;
    bool profile_stages = false;
    bool unit_w = true; // every frame so far had obs_weight == 1 (weights are integers)
    // Temporal blocking: up to `batch` consecutive frames are prepared as they arrive and applied to
    // every brick in ONE visit (tsdf_grid_set_batch).  q_* is the batch being collected.
    int batch = 1;
    int q_n = 0;
    FrameParams q_f[DENSE_MAX_BATCH];
    CullParams q_c[DENSE_MAX_BATCH];
    FrameScratch *q_fs[DENSE_MAX_BATCH] = {};
};

static int grid_flush_batch(tsdf_grid *g);
// entries behind a brick list that a persistent CTA may read before it knows the list's length (at most one per CTA)
static const size_t LIST_PAD = 32 * 256;

static int grid_free(tsdf_grid *g)
{
    if (!g) return TSDF_OK;
    cudaSetDevice(g->device);
    if (g->bricks && g->q_n) grid_flush_batch(g); // frames still waiting for their batch are applied, not dropped
    g->io.release();
    g->geom.release();
    if (g->bricks) cudaFree(g->bricks);
    for (int i = 0; i < FRAME_SLOTS; ++i)
        if (g->vis_list[i]) cudaFree(g->vis_list[i]);
    if (g->tot) cudaFree(g->tot);
    g->counters.release();
    delete g;
    return TSDF_OK;
}

extern "C" int tsdf_grid_create(const int64_t dim[3], const float origin[3], double voxel_size, double trunc,
                                int device, int rank, int world, tsdf_grid **out)
{
    BAD_ARG(!out, "out is NULL");
    *out = nullptr;
    BAD_ARG(world < 1 || rank < 0 || rank >= world, "bad rank/world %d/%d", rank, world);
    int sm = 0;
    int rc = tsdf_select_device(device, &sm);
    if (rc) return rc;
    tsdf_grid *g = new (std::nothrow) tsdf_grid;
    BAD_ARG(!g, "out of host memory");
    g->device = device;
    g->sm_count = sm;
    g->rank = rank;
    g->world = world;
    if ((rc = g->geom.init(dim, origin, voxel_size, trunc)) || (rc = g->io.init())) {
        grid_free(g);
        return rc;
    }
    const int nbx = g->geom.nb[0];
    g->nlbx = sheets_of(nbx, rank, world);
    g->n_x_local = 0;
    for (int l = 0; l < g->nlbx; ++l) {
        const int64_t x0 = (int64_t)sheet_of(l, rank, world) * BRICK;
        g->n_x_local += std::min<int64_t>(BRICK, dim[0] - x0);
    }
    g->nbricks = (size_t)g->nlbx * g->geom.nb[1] * g->geom.nb[2];
    // measurement knobs (profiles/r1_dense_classes.md); the defaults are the measured best
    if (const char *e = getenv("TSDF_DENSE_CLASSES")) g->classes = std::min(DENSE_MAX_CLASSES, std::max(1, atoi(e)));
    // (lbx + by + bz) % classes: every z-column holds at most ceil(nbz / classes) bricks of a class
    g->io.timing = false; // see tsdf_grid_set_profiling
    g->class_cap = (size_t)g->nlbx * g->geom.nb[1] * ((g->geom.nb[2] + g->classes - 1) / g->classes);
    if (g->nbricks >= (size_t)1 << 31) {
        grid_free(g);
        tsdf_set_error("too many bricks per rank (%zu)", g->nbricks);
        return TSDF_ERR_BAD_ARG;
    }
    cudaError_t e;
    const size_t nb_alloc = std::max<size_t>(g->nbricks, 1);
    e = cudaMalloc(&g->bricks, nb_alloc * BRICK_FLOATS * sizeof(float));
    for (int i = 0; i < FRAME_SLOTS && e == cudaSuccess; ++i)
        e = cudaMalloc(&g->vis_list[i], (g->classes * g->class_cap + LIST_PAD) * sizeof(BrickRef));
    if (e == cudaSuccess) e = cudaMalloc(&g->tot, sizeof(Totals));
    if (e != cudaSuccess) {
        tsdf_set_error("device allocation of %.2f GB failed: %s", nb_alloc * 6144.0 / 1e9, cudaGetErrorString(e));
        cudaGetLastError();
        grid_free(g);
        return TSDF_ERR_OOM;
    }
    cudaMemsetAsync(g->tot, 0, sizeof(Totals), g->io.compute);
    launch_fill_bricks(g->bricks, g->nbricks, sm, g->io.compute);
    // the persistent integrate grid leaves INTEGRATE_RESERVE_CTAS CTA slots per SM free so that the
    // next frame's preparation kernels (other stream) can run beside it
    g->integrate_grid = sm * std::max(1, integrate_dense_ctas_per_sm() - integrate_reserve_ctas());
    if (const char *e = getenv("TSDF_DENSE_CTAS_PER_SM")) g->integrate_grid = sm * std::min(32, std::max(1, atoi(e)));
    if ((rc = g->counters.init(g->classes * g->integrate_grid * (INTEGRATE_THREADS / 32)))) {
        grid_free(g);
        return rc;
    }
    e = cudaStreamSynchronize(g->io.compute);
    if (e != cudaSuccess) {
        tsdf_set_error("grid initialisation failed: %s", cudaGetErrorString(e));
        grid_free(g);
        return TSDF_ERR_CUDA;
    }
    *out = g;
    return TSDF_OK;
}

extern "C" int tsdf_grid_destroy(tsdf_grid *g) { return grid_free(g); }

// Issues cull + integrate for the frames collected so far (no-op when none are pending).
static int grid_flush_batch(tsdf_grid *g)
{
    const int k = g->q_n;
    if (!k) return TSDF_OK;
    g->q_n = 0;
    int rc;
    BrickRef *list = g->vis_list[g->io.slot];
    if (k == 1)
        launch_cull_bricks(g->q_f[0], g->q_c[0], g->nlbx, g->geom.nb[1], g->geom.nb[2], g->rank, g->world, g->classes, list,
                           g->class_cap, g->q_fs[0], g->io.pre);
    else
        launch_cull_bricks_batch(g->q_f, g->q_c, g->q_fs, k, g->nlbx, g->geom.nb[1], g->geom.nb[2], g->rank, g->world,
                                 g->classes, list, g->class_cap, g->io.pre);
    // Independent classes of bricks, one kernel each on its own stream: the next kernel's CTAs take
    // the SM slots a kernel's tail frees, and the next batch's first kernel does the same for the
    // last one, so neither the launch gap nor the drain of the persistent grids leaves the SMs idle.
    // With stage profiling on, all run on `compute` so that the stage marks delimit them.
    g->io.n_side = g->profile_stages ? 0 : g->classes - 1;
    if (!g->io.n_side && (rc = g->io.join())) return rc;
    if ((rc = g->io.prepared())) return rc;
    if (g->profile_stages && (rc = g->io.mark())) return rc;
    const VolumeParams vp = g->geom.params();
    const int warps = g->integrate_grid * (INTEGRATE_THREADS / 32);
    for (int c = 0; c < g->classes; ++c) {
        const int *count = c ? &g->q_fs[0]->vis_extra[c - 1] : &g->q_fs[0]->vis_count;
        cudaStream_t st = (c && g->io.n_side) ? g->io.side[c - 1] : g->io.compute;
        if (k == 1)
            launch_integrate_dense(g->q_f[0], vp, list + c * g->class_cap, count, g->bricks,
                                   g->counters.view(g->io.slot, c * warps), g->integrate_grid, st);
        else
            launch_integrate_dense_batch(g->q_f, k, vp, list + c * g->class_cap, count, g->bricks,
                                         g->counters.view(g->io.slot, c * warps), g->integrate_grid, st);
    }
    CUDA_TRY(cudaGetLastError());
    return g->io.end_kernels();
}

// Prepares one frame (stream `pre`, overlaps the previous batch's integrate kernels) and adds it to
// the batch being collected; a full batch is issued at once.
static int grid_enqueue(tsdf_grid *g, const uint8_t *d_rgb, const double *d_depth, int H, int W, FrameParams &f,
                        CullParams &c)
{
    int rc;
    // one image size per batch (a new size reallocates the per-frame scratch: nothing may be pending)
    if (g->q_n && (g->q_f[0].H != H || g->q_f[0].W != W) && (rc = grid_flush_batch(g))) return rc;
    if ((rc = g->io.ensure_px(H * W)) || (rc = g->io.ensure_tiles(H, W))) return rc;
    const int sub = g->q_n;
    if (sub == 0 && (rc = g->io.begin_kernels())) return rc; // picks the slot, waits for its previous user
    const int pi = FrameIO::prep_index(g->io.slot, sub);
    FrameScratch *fs = reinterpret_cast<FrameScratch *>(g->io.scratch_slot[pi]);
    c.pyr = g->io.pyr_slot[pi];
    f.depth = d_depth;
    f.px = g->io.px_slot[pi];
    g->unit_w = g->unit_w && f.obs == 1.0;
    f.unit_w = g->unit_w ? 1 : 0;
    c.tag = g->io.next_tag();
    launch_prep_frame(d_rgb, d_depth, g->io.px_slot[pi], fs, g->io.pyr_slot[pi], H, W, c.tag, g->io.pre);
    CUDA_TRY(cudaGetLastError());
    if ((rc = g->io.release_upload_after_prep())) return rc;
    g->q_f[sub] = f;
    g->q_c[sub] = c;
    g->q_fs[sub] = fs;
    g->q_n = sub + 1;
    g->frames++;
    if (g->q_n >= g->batch) return grid_flush_batch(g);
    return TSDF_OK;
}

extern "C" int tsdf_grid_integrate(tsdf_grid *g, const uint8_t *rgb, const double *depth, int H, int W,
                                   const double K[9], const double Tinv[16], double obs_weight)
{
    BAD_ARG(!g || !rgb || !depth, "NULL argument");
    CUDA_TRY(cudaSetDevice(g->device));
    FrameParams f;
    CullParams c;
    int rc = make_frame_params(&f, &c, g->geom, H, W, K, Tinv, obs_weight);
    if (rc) return rc;
    const uint8_t *d_rgb;
    const double *d_depth;
    if ((rc = g->io.upload(rgb, depth, H * W, &d_rgb, &d_depth))) return rc;
    rc = grid_enqueue(g, d_rgb, d_depth, H, W, f, c);
    const int rc2 = g->io.finish_upload(); // the caller's buffers are only borrowed for the duration of the call
    return rc ? rc : rc2;
}

extern "C" int tsdf_grid_integrate_u16(tsdf_grid *g, const uint8_t *rgb, const uint16_t *depth_raw, int H, int W,
                                       const double K[9], const double Tinv[16], double obs_weight,
                                       double depth_scale, int invalid_raw)
{
    BAD_ARG(!g || !rgb || !depth_raw, "NULL argument");
    BAD_ARG(!(depth_scale > 0.0), "depth_scale must be positive");
    CUDA_TRY(cudaSetDevice(g->device));
    FrameParams f;
    CullParams c;
    int rc = make_frame_params(&f, &c, g->geom, H, W, K, Tinv, obs_weight);
    if (rc) return rc;
    const uint8_t *d_rgb;
    const double *d_depth;
    if ((rc = g->io.upload_u16(rgb, depth_raw, H * W, depth_scale, invalid_raw, &d_rgb, &d_depth))) return rc;
    rc = grid_enqueue(g, d_rgb, d_depth, H, W, f, c);
    const int rc2 = g->io.finish_upload();
    return rc ? rc : rc2;
}

extern "C" int tsdf_grid_integrate_device(tsdf_grid *g, const uint8_t *d_rgb, const double *d_depth, int H, int W,
                                          const double K[9], const double Tinv[16], double obs_weight)
{
    BAD_ARG(!g || !d_rgb || !d_depth, "NULL argument");
    CUDA_TRY(cudaSetDevice(g->device));
    FrameParams f;
    CullParams c;
    int rc = make_frame_params(&f, &c, g->geom, H, W, K, Tinv, obs_weight);
    if (rc) return rc;
    return grid_enqueue(g, d_rgb, d_depth, H, W, f, c);
}

extern "C" int tsdf_grid_integrate_device_u16(tsdf_grid *g, const uint8_t *d_rgb, const uint16_t *d_depth_raw, int H, int W,
                                              const double K[9], const double Tinv[16], double obs_weight,
                                              double depth_scale, int invalid_raw)
{
    BAD_ARG(!g || !d_rgb || !d_depth_raw, "NULL argument");
    BAD_ARG(!(depth_scale > 0.0), "depth_scale must be positive");
    CUDA_TRY(cudaSetDevice(g->device));
    FrameParams f;
    CullParams c;
    int rc = make_frame_params(&f, &c, g->geom, H, W, K, Tinv, obs_weight);
    if (rc) return rc;
    const double *d_depth;
    if ((rc = g->io.convert_u16(d_depth_raw, H * W, depth_scale, invalid_raw, &d_depth))) return rc;
    return grid_enqueue(g, d_rgb, d_depth, H, W, f, c);
}

extern "C" int tsdf_grid_frame_source(tsdf_grid *g, void *producer_stream)
{
    BAD_ARG(!g, "NULL handle");
    CUDA_TRY(cudaSetDevice(g->device));
    return g->io.adopt(static_cast<cudaStream_t>(producer_stream));
}

extern "C" int tsdf_grid_local_x(tsdf_grid *g, int64_t *n_x, int64_t *xs)
{
    BAD_ARG(!g, "NULL handle");
    if (n_x) *n_x = g->n_x_local;
    if (xs) {
        int64_t k = 0;
        for (int l = 0; l < g->nlbx; ++l) {
            const int64_t x0 = (int64_t)sheet_of(l, g->rank, g->world) * BRICK;
            for (int64_t x = x0; x < std::min<int64_t>(x0 + BRICK, g->geom.dim[0]); ++x) xs[k++] = x;
        }
    }
    return TSDF_OK;
}

extern "C" int tsdf_grid_download_device(tsdf_grid *g, float *d_tsdf, float *d_weight, float *d_color)
{
    BAD_ARG(!g, "NULL handle");
    CUDA_TRY(cudaSetDevice(g->device));
    int rc = grid_flush_batch(g);
    if (rc || (rc = g->io.join())) return rc;
    launch_unbrick(g->bricks, (int)g->n_x_local, (int)g->geom.dim[1], (int)g->geom.dim[2], g->geom.nb[1],
                   g->geom.nb[2], d_tsdf, d_weight, d_color, g->sm_count, g->io.compute);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(g->io.compute));
    return TSDF_OK;
}

extern "C" int tsdf_grid_download(tsdf_grid *g, float *tsdf, float *weight, float *color)
{
    BAD_ARG(!g, "NULL handle");
    CUDA_TRY(cudaSetDevice(g->device));
    const size_t n = (size_t)g->n_x_local * g->geom.dim[1] * g->geom.dim[2];
    if (!n) return TSDF_OK;
    int rc = grid_flush_batch(g);
    if (rc || (rc = g->io.join())) return rc;
    // one plane at a time keeps the device-side scratch at 4 B/voxel
    float *scratch = nullptr;
    CUDA_TRY(cudaMalloc(&scratch, n * sizeof(float)));
    float *host[3] = {tsdf, weight, color};
    for (int p = 0; p < 3; ++p) {
        if (!host[p]) continue;
        launch_unbrick(g->bricks, (int)g->n_x_local, (int)g->geom.dim[1], (int)g->geom.dim[2], g->geom.nb[1],
                       g->geom.nb[2], p == 0 ? scratch : nullptr, p == 1 ? scratch : nullptr,
                       p == 2 ? scratch : nullptr, g->sm_count, g->io.compute);
        cudaError_t e = cudaMemcpyAsync(host[p], scratch, n * sizeof(float), cudaMemcpyDeviceToHost, g->io.compute);
        if (e == cudaSuccess) e = cudaStreamSynchronize(g->io.compute);
        if (e != cudaSuccess) {
            cudaFree(scratch);
            tsdf_set_error("download failed: %s", cudaGetErrorString(e));
            return TSDF_ERR_CUDA;
        }
    }
    cudaFree(scratch);
    return TSDF_OK;
}

extern "C" int tsdf_grid_stats(tsdf_grid *g, int64_t counters[4])
{
    BAD_ARG(!g || !counters, "NULL argument");
    CUDA_TRY(cudaSetDevice(g->device));
    const StatsBlock *st = nullptr;
    int rc = grid_flush_batch(g);
    if (rc || (rc = g->io.join())) return rc;
    const int slot = g->frames > 0 ? g->io.slot : 0;
    rc = g->counters.collect(slot, g->io.fs, g->tot, nullptr, nullptr, g->io.compute, &st);
    if (rc) return rc;
    counters[0] = g->frames > 0 ? (int64_t)st->updated_last : 0;
    counters[1] = (int64_t)st->updated_total;
    counters[2] = g->frames > 0 ? (int64_t)st->vis_count : 0;
    counters[3] = g->frames;
    return TSDF_OK;
}

extern "C" int tsdf_grid_stats_post(tsdf_grid *g, int64_t *ticket)
{
    BAD_ARG(!g || !ticket, "NULL argument");
    CUDA_TRY(cudaSetDevice(g->device));
    int rc = grid_flush_batch(g);
    if (rc || (rc = g->io.stat_begin())) return rc;
    const int slot = g->frames > 0 ? g->io.slot : 0;
    if ((rc = g->counters.post(slot, g->io.fs, g->tot, nullptr, nullptr, g->io.stat, g->frames, ticket))) return rc;
    return g->io.stat_end();
}

extern "C" int tsdf_grid_stats_wait(tsdf_grid *g, int64_t ticket, int64_t counters[4])
{
    BAD_ARG(!g || !counters, "NULL argument");
    CUDA_TRY(cudaSetDevice(g->device));
    const StatsBlock *st = nullptr;
    int64_t frames = 0;
    int rc = g->counters.wait(ticket, &st, &frames);
    if (rc) return rc;
    counters[0] = frames > 0 ? (int64_t)st->updated_last : 0;
    counters[1] = -1; // later frames may already be adding to the running total: tsdf_grid_stats has it
    counters[2] = frames > 0 ? (int64_t)st->vis_count : 0;
    counters[3] = frames;
    return TSDF_OK;
}

extern "C" int tsdf_grid_sync(tsdf_grid *g)
{
    BAD_ARG(!g, "NULL handle");
    CUDA_TRY(cudaSetDevice(g->device));
    { int rc_ = grid_flush_batch(g); if (rc_) return rc_; }
    CUDA_TRY(cudaStreamSynchronize(g->io.stat));
    CUDA_TRY(cudaStreamSynchronize(g->io.copy));
    CUDA_TRY(cudaStreamSynchronize(g->io.pre));
    CUDA_TRY(cudaStreamSynchronize(g->io.compute));
    return g->io.sync_side();
}

extern "C" int tsdf_grid_flush(tsdf_grid *g)
{
    BAD_ARG(!g, "NULL handle");
    CUDA_TRY(cudaSetDevice(g->device));
    int rc = grid_flush_batch(g);
    return rc ? rc : g->io.join();
}

extern "C" int tsdf_grid_set_batch(tsdf_grid *g, int frames)
{
    BAD_ARG(!g, "NULL handle");
    BAD_ARG(frames < 1 || frames > DENSE_MAX_BATCH, "batch of %d frames unsupported (1..%d)", frames, DENSE_MAX_BATCH);
    CUDA_TRY(cudaSetDevice(g->device));
    int rc = grid_flush_batch(g);
    if (rc || (rc = g->io.set_subs(frames))) return rc;
    g->batch = frames;
    return TSDF_OK;
}

extern "C" int tsdf_grid_last_kernel_ms(tsdf_grid *g, float *ms)
{
    BAD_ARG(!g, "NULL handle");
    CUDA_TRY(cudaSetDevice(g->device));
    return g->io.last_ms(ms);
}

extern "C" int tsdf_grid_set_profiling(tsdf_grid *g, int on)
{
    BAD_ARG(!g, "NULL handle");
    BAD_ARG(on < 0 || on > 2, "profiling mode %d (0 off, 1 stages, 2 frame time only)", on);
    const bool stages = on == 1;
    if (g->profile_stages != stages) { // the integrate kernels change streams: drain all of them first
        int rc = tsdf_grid_sync(g);
        if (rc) return rc;
    }
    g->profile_stages = stages;
    g->io.timing = on != 0;
    g->io.timed = false;
    return TSDF_OK;
}

extern "C" int tsdf_grid_last_stage_ms(tsdf_grid *g, float ms[4], int *n)
{
    BAD_ARG(!g, "NULL handle");
    CUDA_TRY(cudaSetDevice(g->device));
    return g->io.last_stage_ms(ms, n);
}

extern "C" int tsdf_grid_stream(tsdf_grid *g, void **stream)
{
    BAD_ARG(!g || !stream, "NULL argument");
    *stream = (void *)g->io.compute;
    return TSDF_OK;
}
