// capi_fanout.cu -- frame fan-out over NVLink peer memory (include/tsdf_b200.h, tsdf_fanout_*).
//
// BASELINE north_star (3): "each frame's depth, colour and pose broadcast to all ranks".  The reference has one
// process and reads every frame itself (grid_fusion_demos/grid_demo1.py:76-87); with one process per GPU every rank
// uploading every frame puts N copies of the same image on the host's PCIe complex per step.  Here a frame crosses
// PCIe once, on its source rank, and every peer's COPY ENGINE pulls it over NVLink into its own ring buffer:
//
//   * every rank owns one arena (cudaMalloc): `slots` frame buffers + two flag arrays; the arenas are mapped into
//     every process through CUDA IPC handles (exchanged by the caller, e.g. one all-gather at start-up);
//   * source of frame q (slot s = q % slots), on its upload stream: host-to-device copy into the own slot, then one
//     kernel stores q + 1 into ready[s] of every peer's arena.  Uploads of consecutive frames come from different
//     ranks and different PCIe links, and overlap;
//   * every other rank, on its arrival stream: a one-thread kernel waits for the local ready[s] >= q + 1, the copy
//     engine pulls the frame from the source's slot into the own slot, one kernel stores q + 1 into
//     pulled[s][rank] of the SOURCE's arena (which guards the source's slot against its next overwrite).  The map's
//     preparation kernel is ordered after the arrival stream (tsdf_grid_frame_source / vh_frame_source).
//
// No SM-resident collective kernel is involved (an NCCL broadcast kernel needs most of an SM's registers and starves
// behind the persistent integrate grids: measured 225 us per step at 2 ranks against 164 us of device time), no host
// thread ever waits for another process, and flags only grow, so nothing is reset.  A wait that is not satisfied
// within the time-out raises the arena's error word instead of hanging the GPU.
#include <new>
#include <vector>

#include "host_common.h"

struct FanFlags {
    uint32_t error;                     // set by a wait kernel that timed out
    uint32_t pad[119];                  // the ring slots behind the flags start on a 1 KiB boundary
    uint32_t ready[FANOUT_MAX_SLOTS]; // sequence number + 1 of the frame its source has staged in ITS copy of the slot
    uint32_t pulled[FANOUT_MAX_SLOTS][FANOUT_MAX_WORLD]; // [slot][rank]: last frame rank has pulled from THIS arena's slot, + 1
};

static_assert(sizeof(FanFlags) == 1024, "ring slots must stay 256-byte aligned");

struct PeerFlagPtrs {
    uint32_t *p[FANOUT_MAX_WORLD];
};

// Stores `value` into n flag words that live in peer arenas (or the own one).
__global__ void k_fan_signal(PeerFlagPtrs dst, int n, uint32_t value)
{
    if (threadIdx.x < n) {
        __threadfence_system(); // stream order already completed the copies; the fence orders this kernel's own view
        *reinterpret_cast<volatile uint32_t *>(dst.p[threadIdx.x]) = value;
        __threadfence_system();
    }
}

// Waits until every one of the n local flag words flags[i * stride] (i != skip) has reached `value`.
__global__ void k_fan_wait(const uint32_t *flags, int n, int stride, int skip, uint32_t value, unsigned long long timeout_ns,
                           uint32_t *error)
{
    const int i = threadIdx.x;
    if (i >= n || i == skip) return;
    const volatile uint32_t *f = flags + (size_t)i * stride;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    // sequence numbers compared as a signed distance: correct across the 32-bit wrap
    while ((int32_t)(*f - value) < 0) {
        __nanosleep(200);
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t - t0 > timeout_ns) {
            *error = 1u;
            break;
        }
    }
    __threadfence_system();
}

struct tsdf_fanout {
    int device = 0, rank = 0, world = 1, slots = 0;
    size_t slot_bytes = 0, arena_bytes = 0;
    uint8_t *arena = nullptr;  // [FanFlags | slots x slot_bytes]
    uint8_t *peer[FANOUT_MAX_WORLD] = {};
    bool connected = false;
    cudaStream_t up = nullptr;     // this rank's uploads (source side)
    cudaStream_t stream = nullptr; // arrivals: what the map's preparation kernel is ordered after
    cudaEvent_t pushed = nullptr;  // behind the last host-to-device copy that borrowed the caller's buffers
    cudaEvent_t staged[FANOUT_MAX_SLOTS] = {};   // own upload into the slot complete (stream `up`)
    cudaEvent_t consumed[FANOUT_MAX_SLOTS] = {}; // the map's preparation kernel has read the own slot (stream `stream`)
    int64_t last_pushed[FANOUT_MAX_SLOTS];       // last frame this rank staged in the slot as its source, or -1
    bool borrowed = false;
    double timeout_s = 20.0;

    FanFlags *flags(int r) const { return reinterpret_cast<FanFlags *>(peer[r]); }
    uint8_t *slot_of(int r, int s) const { return peer[r] + sizeof(FanFlags) + (size_t)s * slot_bytes; }
    // Before the own slot is overwritten: every peer has pulled the frame this rank last staged there.
    int guard(int s, cudaStream_t st)
    {
        if (world == 1 || last_pushed[s] < 0) return TSDF_OK;
        k_fan_wait<<<1, FANOUT_MAX_WORLD, 0, st>>>(flags(rank)->pulled[s], world, 1, rank, (uint32_t)(last_pushed[s] + 1),
                                                   (unsigned long long)(timeout_s * 1e9), &flags(rank)->error);
        CUDA_TRY(cudaGetLastError());
        return TSDF_OK;
    }
};

static int fan_check(const tsdf_fanout *f, int64_t seq, int src)
{
    BAD_ARG(!f, "NULL handle");
    BAD_ARG(seq < 0, "negative frame number");
    BAD_ARG(src < 0 || src >= f->world, "source rank %d outside the world of %d", src, f->world);
    BAD_ARG(f->world > 1 && !f->connected, "fan-out not connected to its peers (tsdf_fanout_connect)");
    return TSDF_OK;
}

extern "C" int tsdf_fanout_create(int device, int rank, int world, int64_t slot_bytes, int slots, tsdf_fanout **out)
{
    BAD_ARG(!out, "out is NULL");
    *out = nullptr;
    BAD_ARG(world < 1 || world > FANOUT_MAX_WORLD || rank < 0 || rank >= world, "bad rank/world %d/%d (at most %d ranks)",
            rank, world, FANOUT_MAX_WORLD);
    BAD_ARG(slots < 1 || slots > FANOUT_MAX_SLOTS, "slots must be 1..%d", FANOUT_MAX_SLOTS);
    BAD_ARG(slot_bytes <= 0, "slot_bytes must be positive");
    int sm = 0;
    int rc = tsdf_select_device(device, &sm);
    if (rc) return rc;
    tsdf_fanout *f = new (std::nothrow) tsdf_fanout;
    BAD_ARG(!f, "out of host memory");
    f->device = device; f->rank = rank; f->world = world; f->slots = slots;
    f->slot_bytes = ((size_t)slot_bytes + 255) / 256 * 256;
    // whole 2 MiB blocks: an IPC handle shares the driver's underlying block, which must hold nothing else
    f->arena_bytes = (sizeof(FanFlags) + (size_t)slots * f->slot_bytes + ((size_t)2 << 20) - 1) / ((size_t)2 << 20) * ((size_t)2 << 20);
    cudaError_t e = cudaMalloc(&f->arena, f->arena_bytes);
    if (e == cudaSuccess) e = cudaMemset(f->arena, 0, sizeof(FanFlags));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&f->up, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&f->pushed, cudaEventDisableTiming);
    for (int i = 0; i < slots; ++i) {
        f->last_pushed[i] = -1;
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&f->staged[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&f->consumed[i], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) {
        tsdf_set_error("fan-out arena of %.1f MB: %s", f->arena_bytes / 1e6, cudaGetErrorString(e));
        cudaGetLastError();
        if (f->arena) cudaFree(f->arena);
        delete f;
        return TSDF_ERR_CUDA;
    }
    f->peer[rank] = f->arena;
    f->connected = world == 1;
    *out = f;
    return TSDF_OK;
}

extern "C" int tsdf_fanout_export(tsdf_fanout *f, uint8_t handle[64])
{
    BAD_ARG(!f || !handle, "NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CUDA_TRY(cudaSetDevice(f->device));
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, f->arena));
    memcpy(handle, &h, 64);
    return TSDF_OK;
}

extern "C" int tsdf_fanout_connect(tsdf_fanout *f, const uint8_t *handles)
{
    BAD_ARG(!f || !handles, "NULL argument");
    BAD_ARG(f->connected, "fan-out already connected");
    CUDA_TRY(cudaSetDevice(f->device));
    for (int r = 0; r < f->world; ++r) {
        if (r == f->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + 64 * (size_t)r, 64);
        void *p = nullptr;
        CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        f->peer[r] = static_cast<uint8_t *>(p);
    }
    f->connected = true;
    return TSDF_OK;
}

extern "C" int tsdf_fanout_push(tsdf_fanout *f, int64_t seq, const void *depth, int64_t depth_bytes, const void *rgb,
                                int64_t rgb_bytes, int64_t rgb_offset)
{
    int rc = fan_check(f, seq, f ? f->rank : 0);
    if (rc) return rc;
    BAD_ARG(!depth || !rgb, "NULL frame buffer");
    BAD_ARG(depth_bytes <= 0 || rgb_bytes <= 0 || rgb_offset < depth_bytes || (size_t)(rgb_offset + rgb_bytes) > f->slot_bytes,
            "frame of %lld + %lld bytes does not fit the fan-out's slots", (long long)depth_bytes, (long long)rgb_bytes);
    CUDA_TRY(cudaSetDevice(f->device));
    const int s = (int)(seq % f->slots);
    if (f->borrowed) { // at most one borrowed host frame in flight
        CUDA_TRY(cudaEventSynchronize(f->pushed));
        f->borrowed = false;
    }
    // the slot's previous frame: pulled by every peer (if it was staged here) and consumed by this rank's own map
    if ((rc = f->guard(s, f->up))) return rc;
    if (seq >= f->slots) CUDA_TRY(cudaStreamWaitEvent(f->up, f->consumed[s], 0));
    uint8_t *own = f->slot_of(f->rank, s);
    CUDA_TRY(cudaMemcpyAsync(own, depth, (size_t)depth_bytes, cudaMemcpyHostToDevice, f->up));
    CUDA_TRY(cudaMemcpyAsync(own + rgb_offset, rgb, (size_t)rgb_bytes, cudaMemcpyHostToDevice, f->up));
    CUDA_TRY(cudaEventRecord(f->pushed, f->up));
    CUDA_TRY(cudaEventRecord(f->staged[s], f->up));
    f->borrowed = true;
    f->last_pushed[s] = seq;
    if (f->world > 1) {
        PeerFlagPtrs dst;
        int k = 0;
        for (int r = 0; r < f->world; ++r)
            if (r != f->rank) dst.p[k++] = &f->flags(r)->ready[s];
        k_fan_signal<<<1, FANOUT_MAX_WORLD, 0, f->up>>>(dst, k, (uint32_t)(seq + 1));
        CUDA_TRY(cudaGetLastError());
    }
    return TSDF_OK;
}

extern "C" int tsdf_fanout_acquire(tsdf_fanout *f, int64_t seq, int src, int64_t frame_bytes, int64_t rgb_offset,
                                   void **d_depth, void **d_rgb, void **stream)
{
    int rc = fan_check(f, seq, src);
    if (rc) return rc;
    BAD_ARG(!d_depth || !d_rgb || !stream, "NULL argument");
    BAD_ARG(rgb_offset < 0 || rgb_offset >= frame_bytes || (size_t)frame_bytes > f->slot_bytes, "frame does not fit the slot");
    CUDA_TRY(cudaSetDevice(f->device));
    const int s = (int)(seq % f->slots);
    uint8_t *own = f->slot_of(f->rank, s);
    if (src == f->rank) {
        BAD_ARG(f->last_pushed[s] != seq, "frame %lld was not pushed by this rank", (long long)seq);
        CUDA_TRY(cudaStreamWaitEvent(f->stream, f->staged[s], 0));
    } else {
        // the own slot is about to be overwritten: its previous frame has been consumed by the own map (ordered on
        // this stream by the map) and, if this rank staged it, pulled by every peer
        if ((rc = f->guard(s, f->stream))) return rc;
        f->last_pushed[s] = -1;
        k_fan_wait<<<1, 32, 0, f->stream>>>(&f->flags(f->rank)->ready[s], 1, 1, -1, (uint32_t)(seq + 1),
                                            (unsigned long long)(f->timeout_s * 1e9), &f->flags(f->rank)->error);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(own, f->slot_of(src, s), (size_t)frame_bytes, cudaMemcpyDeviceToDevice, f->stream));
        PeerFlagPtrs dst;
        dst.p[0] = &f->flags(src)->pulled[s][f->rank];
        k_fan_signal<<<1, 32, 0, f->stream>>>(dst, 1, (uint32_t)(seq + 1));
        CUDA_TRY(cudaGetLastError());
    }
    *d_depth = own;
    *d_rgb = own + rgb_offset;
    *stream = f->stream;
    return TSDF_OK;
}

// After the frame has been handed to the map (frame_source + integrate_device): marks where, on the arrival stream,
// the map's preparation kernel is through with the own slot.
extern "C" int tsdf_fanout_release(tsdf_fanout *f, int64_t seq)
{
    int rc = fan_check(f, seq, 0);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(f->device));
    CUDA_TRY(cudaEventRecord(f->consumed[(int)(seq % f->slots)], f->stream));
    return TSDF_OK;
}

// One frame of the fan-out loop in one call: push (source rank only), acquire, frame_source, integrate_device[_u16],
// release -- exactly that sequence (the per-step cost of five binding calls matters once a shard's frame is 50 us).
extern "C" int tsdf_fanout_integrate(tsdf_fanout *f, tsdf_grid *grid, tsdf_hash *hash, int64_t seq, int src,
                                     const void *depth, const void *rgb, int H, int W, int raw_u16, const double K[9],
                                     const double Tinv[16], double obs_weight, double depth_scale, int invalid_raw)
{
    BAD_ARG(!f || (!grid == !hash), "needs a fan-out and exactly one map");
    BAD_ARG(H <= 0 || W <= 0, "bad image size");
    const int64_t depth_bytes = (int64_t)H * W * (raw_u16 ? 2 : 8), rgb_bytes = (int64_t)H * W * 3;
    const int64_t rgb_off = (depth_bytes + 255) / 256 * 256;
    int rc;
    if (src == f->rank && (rc = tsdf_fanout_push(f, seq, depth, depth_bytes, rgb, rgb_bytes, rgb_off))) return rc;
    void *d_depth = nullptr, *d_rgb = nullptr, *st = nullptr;
    if ((rc = tsdf_fanout_acquire(f, seq, src, rgb_off + rgb_bytes, rgb_off, &d_depth, &d_rgb, &st))) return rc;
    if (grid) {
        if ((rc = tsdf_grid_frame_source(grid, st))) return rc;
        rc = raw_u16 ? tsdf_grid_integrate_device_u16(grid, (const uint8_t *)d_rgb, (const uint16_t *)d_depth, H, W, K, Tinv,
                                                      obs_weight, depth_scale, invalid_raw)
                     : tsdf_grid_integrate_device(grid, (const uint8_t *)d_rgb, (const double *)d_depth, H, W, K, Tinv, obs_weight);
    } else {
        if ((rc = vh_frame_source(hash, st))) return rc;
        rc = raw_u16 ? vh_integrate_device_u16(hash, (const uint8_t *)d_rgb, (const uint16_t *)d_depth, H, W, K, Tinv,
                                               obs_weight, depth_scale, invalid_raw)
                     : vh_integrate_device(hash, (const uint8_t *)d_rgb, (const double *)d_depth, H, W, K, Tinv, obs_weight);
    }
    if (rc) return rc;
    return tsdf_fanout_release(f, seq);
}

extern "C" int tsdf_fanout_drain(tsdf_fanout *f)
{
    BAD_ARG(!f, "NULL handle");
    CUDA_TRY(cudaSetDevice(f->device));
    CUDA_TRY(cudaStreamSynchronize(f->up));
    CUDA_TRY(cudaStreamSynchronize(f->stream));
    f->borrowed = false;
    uint32_t err = 0;
    CUDA_TRY(cudaMemcpy(&err, &f->flags(f->rank)->error, sizeof(err), cudaMemcpyDeviceToHost));
    if (err) {
        tsdf_set_error("frame fan-out: a peer did not deliver or consume a frame within %.0f s", f->timeout_s);
        return TSDF_ERR_CUDA;
    }
    return TSDF_OK;
}

extern "C" int tsdf_fanout_set_timeout(tsdf_fanout *f, double seconds)
{
    BAD_ARG(!f || !(seconds > 0.0), "bad argument");
    f->timeout_s = seconds;
    return TSDF_OK;
}

// The peers must have stopped using this rank's arena (a barrier of the caller's process group) before it is freed.
extern "C" int tsdf_fanout_destroy(tsdf_fanout *f)
{
    if (!f) return TSDF_OK;
    cudaSetDevice(f->device);
    if (f->up) cudaStreamSynchronize(f->up);
    if (f->stream) cudaStreamSynchronize(f->stream);
    for (int r = 0; r < f->world; ++r)
        if (r != f->rank && f->peer[r]) cudaIpcCloseMemHandle(f->peer[r]);
    if (f->pushed) cudaEventDestroy(f->pushed);
    for (int i = 0; i < f->slots; ++i) {
        if (f->staged[i]) cudaEventDestroy(f->staged[i]);
        if (f->consumed[i]) cudaEventDestroy(f->consumed[i]);
    }
    if (f->up) cudaStreamDestroy(f->up);
    if (f->stream) cudaStreamDestroy(f->stream);
    if (f->arena) cudaFree(f->arena);
    cudaGetLastError();
    delete f;
    return TSDF_OK;
}
