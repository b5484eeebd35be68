#!/usr/bin/env python
"""bench.py -- TSDF integration throughput on B200 (BASELINE.json metric: voxel-updates/s and
frames/s, dense grid and voxel-block hash map, against the HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A "step" is one RGB-D frame integrated into the map.  Workload = BASELINE.json configs[3]: a
synthetic 10 m x 10 m x 4 m room at 1 cm voxels (1000 x 1000 x 400 = 4.0e8 voxels, 4.8 GB of
voxel state, far larger than the 126 MB L2, so every frame streams its bricks from HBM), 640x480
frames on a seeded trajectory (union_thesis_slam_b200/synth.py), obs_weight 1.  With N > 1 the
same room is split into brick-thick cyclic x-sheets (rank = sheet % N): strong scaling, no
per-frame collective.  The hash-map line (BASELINE.json configs[4]: 1280x720 stream into a
40 x 40 x 6 m room at 2 cm voxels) rides in the "hash" object of the same JSON line.

Printed keys follow the driver contract; see DESIGN.md section "Measurement" for definitions.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

if "reference" in sys.argv:
    # --impl reference under torch.distributed.run inherits OMP_NUM_THREADS=1: the CPU arm must use
    # every host thread (numba prange, OpenBLAS dgemm behind np.dot, OpenMP of the C port) at any N
    for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "NUMBA_NUM_THREADS"):
        os.environ.pop(_k, None)
    os.environ["NUMBA_NUM_THREADS"] = str(os.cpu_count() or 1)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

_OUT = sys.stdout
ROOM = (10.0, 10.0, 4.0)
VOXEL = 0.01
IM_W, IM_H = 640, 480
N_FRAMES = 32           # distinct frames kept resident in HBM and replayed in order
HASH_VOXEL = 0.02
HASH_ROOM = (40.0, 40.0, 6.0)
HASH_W, HASH_H = 1280, 720
CPU_SHEET_STRIDE = 16   # cpu baseline integrates every 16th x-sheet of the same grid


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and clock-event (throttle) reasons sampled DURING the measurement: an NVML polling
    thread at ~1 ms (the timed region of a 20-step run lasts a few milliseconds, far below what an
    `nvidia-smi -lms` subprocess resolves), with nvidia-smi kept as the fallback.  mark()/unmark()
    bracket the timed region on the host clock; samples inside it are reported first, samples of the
    whole measured section (timed region + roofline pass + end-to-end pass) beside them."""

    def __init__(self, device):
        import threading
        self.samples = []   # (t, sm_mhz, reasons bitmask)
        self.t0 = self.t1 = None
        self.stop_flag = False
        self.max_mhz = None
        self.thread = None
        self.fallback = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(device))
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception:
            self.fallback = _SmiSampler(device)

    @staticmethod
    def _physical_index(device):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [x for x in vis.split(",") if x.strip()]
            if device < len(ids) and ids[device].strip().isdigit():
                return int(ids[device])
        return device

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.samples.append((time.perf_counter(), float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)),
                                     int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))))
            except Exception:
                pass
            time.sleep(0.001)

    def sample_now(self):
        """One sample taken by the calling thread (the timed region is shorter than a thread switch interval:
        the benchmark calls this while the device is still busy with the steps it has just been issued)."""
        if self.fallback is None:
            try:
                nv = self.nv
                self.samples.append((time.perf_counter(), float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)),
                                     int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))))
            except Exception:
                pass

    def mark(self):
        self.t0 = time.perf_counter()

    def unmark(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if self.fallback is not None:
            return self.fallback.stop()
        self.stop_flag = True
        self.thread.join(timeout=2)
        nv = self.nv
        names = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown),
                 ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown),
                 ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap))

        def summarise(rows):
            reasons = sorted({n for _, _, bits in rows for n, m in names if bits & m})
            return (float(np.median([r[1] for r in rows])) if rows else None), reasons

        timed = [r for r in self.samples if self.t0 is not None and self.t1 is not None and self.t0 <= r[0] <= self.t1]
        sm_t, reasons_t = summarise(timed)
        sm_all, reasons_all = summarise(self.samples)
        return {"sm_mhz": sm_t if timed else sm_all, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(set(reasons_t) | set(reasons_all)), "samples": len(timed),
                "samples_measured_section": len(self.samples), "sm_mhz_measured_section": sm_all,
                "source": "NVML: polling thread (1 ms) + samples taken by the main thread while the timed steps "
                          "execute; `samples` lie inside the timed region"}


class _SmiSampler:
    """nvidia-smi fallback of ClockSampler."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=self.tmp, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": "nvidia-smi -lms 20"}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.tmp.read().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.tmp.name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                       samples=len(sm))
        return out


def cpu_model():
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def make_frames(n, width=IM_W, height=IM_H):
    from union_thesis_slam_b200 import synth
    room = synth.room_bounds(ROOM)
    frames, K = synth.room_stream(n, room, width=width, height=height, seed=0)
    return room, frames, K


# ----------------------------------------------------------------------------------------------
def cpu_port(frames, K, room, voxel, steps, warmup, stride=CPU_SHEET_STRIDE):
    """Times the CPU oracle (C/OpenMP restatement of the reference's numba path) on all host
    threads over a bounded sample: every `stride`-th x-sheet of the same grid."""
    from oracle import oracle
    dim, _ = oracle.volume_geometry(room.copy(), voxel)
    xs = np.arange(0, int(dim[0]), stride, dtype=np.int64)
    vol = oracle.OracleTSDFVolume(room.copy(), voxel, xs=xs, nthreads=0)
    cores = int(oracle.lib().oracle_max_threads())
    for i in range(warmup):
        c, d, p = frames[i % len(frames)]
        vol.integrate(c, d, K, p)
    t0 = time.perf_counter()
    upd = 0
    for i in range(steps):
        c, d, p = frames[(warmup + i) % len(frames)]
        upd += vol.integrate(c, d, K, p)
    dt = time.perf_counter() - t0
    sample = "every %dth x-sheet (%d of %d sheets, %.2e voxels) of the %.3g m room grid, %d frames" % (
        stride, len(xs), int(dim[0]), len(xs) * float(dim[1]) * float(dim[2]), voxel, steps)
    return dict(value=upd / dt, unit="voxel-updates/s", cores=cores, kind="port", sample=sample,
                frames_per_s_sample=steps / dt, seconds=dt, cpu_model=cpu_model(), host_cpus=os.cpu_count())


REF_SLABS = 8           # the reference arm integrates REF_SLABS slabs of REF_SLAB_SHEETS x-sheets each,
REF_SLAB_SHEETS = 8     # spread evenly over the room: 8 x 8 x 1000 x 400 = 2.56e7 voxels (1/15.6 of the grid)


def cpu_reference(frames, K, room, voxel, steps, warmup, slabs=REF_SLABS, sheets=REF_SLAB_SHEETS):
    """Times the UNMODIFIED reference (oracle/_ref: grid_fusion.TSDFVolume, numba CPU path) on this
    box's host cores with the protocol of grid_demo1.py:28-41,76-87 (perf_counter around every
    integrate call) on a bounded sample of the workload: `slabs` x-slabs of the room, each built by
    the reference's own constructor on its own bounds and fed the same frames (the reference cannot
    hold the whole 4.0e8-voxel grid: ~138 B/voxel of host memory).  Its cost scales with
    the voxels VISITED (every voxel of the bounds, every frame), which is why the sample is spread
    over the room: the sample's update density then matches the whole grid's."""
    import contextlib
    import io
    from oracle import build_ref
    gf, _ = build_ref.load()
    import numba
    numba.set_num_threads(min(numba.config.NUMBA_NUM_THREADS, os.cpu_count() or 1))
    room = np.asarray(room, float)
    width = (sheets - 0.5) * voxel   # ceil() in the reference constructor makes it `sheets` sheets
    x_lo = np.linspace(room[0, 0], room[0, 1] - width, slabs + 2)[1:-1]
    x_lo = np.round(x_lo / (8 * voxel)) * 8 * voxel
    vols = []
    with contextlib.redirect_stdout(io.StringIO()):
        for x0 in x_lo:
            b = room.copy()
            b[0] = (x0, x0 + width)
            vols.append(gf.TSDFVolume(b, voxel, use_gpu=False))
    visits = float(sum(int(np.prod(v._vol_dim)) for v in vols))

    def total_w():
        return float(sum(np.sum(v._weight_vol_cpu, dtype=np.float64) for v in vols))

    for i in range(max(1, warmup)):     # the first call JIT-compiles the numba helpers
        c, d, p = frames[i % len(frames)]
        for v in vols:
            v.integrate(c, d, K, p, obs_weight=1.)
    w0 = total_w()
    dt = 0.0
    for i in range(steps):
        c, d, p = frames[(max(1, warmup) + i) % len(frames)]
        t0 = time.perf_counter()
        for v in vols:
            v.integrate(c, d, K, p, obs_weight=1.)
        dt += time.perf_counter() - t0
    upd = total_w() - w0                 # obs_weight 1: updates == increase of the weight sum
    sample = ("%d x-slabs of %d sheets spread over the %.3g m room grid (%.2e of %.2e voxels), %d frames, "
              "reference TSDFVolume.integrate (numba CPU path) called per slab" % (
                  slabs, sheets, voxel, visits, float(np.prod(np.ceil((room[:, 1] - room[:, 0]) / voxel))), steps))
    return dict(value=upd / dt, unit="voxel-updates/s", cores=int(numba.get_num_threads()), kind="reference",
                sample=sample, frames_per_s_sample=steps / dt, seconds=dt, cpu_model=cpu_model(),
                host_cpus=os.cpu_count(), voxel_visits_per_s=visits * steps / dt,
                threading_layer=numba.threading_layer(), numba=numba.__version__, numpy=np.__version__)


def reference_built():
    try:
        from oracle import build_ref
        return build_ref.built()
    except Exception:
        return False


def run_reference(args):
    """--impl reference: the reference's own CPU path timed on this box's host cores: the unmodified
    reference byte-compiled into oracle/_ref (oracle/build_ref.py) -- kind "reference" -- or, when
    oracle/_ref was never built, its C/OpenMP restatement oracle/tsdf_oracle.c -- kind "port"."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    room, frames, K = make_frames(min(N_FRAMES, max(1, args.warmup) + args.steps))
    if reference_built() and not args.port:
        r = cpu_reference(frames, K, room, VOXEL, args.steps, args.warmup)
    else:
        r = cpu_port(frames, K, room, VOXEL, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "tsdf_voxel_updates_per_s", "value": r["value"], "unit": "voxel-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * r["seconds"] / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        # the same workload description as the product arm prints at this N (the driver compares the two)
        "config": dict(workload_config(int(os.environ.get("WORLD_SIZE", "1"))), frames_per_brick_visit=args.batch),
        "cpu_baseline": {k: v for k, v in r.items() if k not in ("seconds", "frames_per_s_sample")},
        "e2e": {"value": r["value"], "unit": "voxel-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "frames_per_s": r["frames_per_s_sample"], "gpu_launches": 0,
    }
    print(json.dumps(line), file=_OUT, flush=True)
    return 0


def workload_config(world):
    return {"workload": "BASELINE.json configs[3]: synthetic 10x10x4 m room, dense grid, 1 cm voxels "
                        "(1000x1000x400 = 4.0e8 voxels, 4.8 GB), 640x480 frames, seeded trajectory",
            "voxel_size_m": VOXEL, "grid": [1000, 1000, 400], "image": [IM_W, IM_H],
            "frames_resident": N_FRAMES, "obs_weight": 1.0,
            "partition": "brick-thick x-sheets dealt to %d rank(s) in snake order" % world,
            "l2_policy": "working set (4.8 GB voxel state, 2-4 GB touched per frame) exceeds the 126 MB L2"}


# ----------------------------------------------------------------------------------------------
class Dist:
    """World / rank plumbing shared by the dense and the hash sections."""

    def __init__(self):
        import torch
        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise RuntimeError("bench.py needs a CUDA device: the product has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist:
            self.dist.barrier()

    def _reduce(self, x, op):
        if not self.dist:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x):
        return self._reduce(x, self.dist.ReduceOp.MAX) if self.dist else x

    def sum(self, x):
        return self._reduce(x, self.dist.ReduceOp.SUM) if self.dist else x

    def close(self):
        if self.dist:
            self.dist.barrier()
            self.dist.destroy_process_group()


def timed_device(D, vol, step, n, first, sampler=None):
    """Device time of n frames: CUDA events on the handle's own stream bracketed by barrier + synchronize,
    max over ranks.  Returns (ms, updates summed over ranks)."""
    torch = D.torch
    stream = torch.cuda.ExternalStream(vol.cuda_stream(), device=D.dev)
    u0 = vol.stats()["updated_total"]
    D.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(n):
        step(first + i)
    vol.flush()  # integrate kernels run on sibling streams: join them (on the device) before the end event
    e1.record(stream)
    if sampler:
        for _ in range(3):   # the host is ahead of the device here: these samples see the timed steps executing
            sampler.sample_now()
    vol.sync()
    D.barrier()
    ms = D.max(e0.elapsed_time(e1))
    upd = D.sum(float(vol.stats()["updated_total"] - u0))
    return ms, upd


E2E_DEPTH = 6  # steps whose result read-back may be outstanding while the host issues the next one


def e2e_loop(D, vol, call, n, first, after=None):
    """End to end through the public API with HOST frames: every step uploads its frame (inside the
    integrate call) and reads a step result back.  The read-back is pipelined: step k posts its
    counters (stats_post: a 64-byte device->host copy into pinned memory behind the step's kernels) and
    the host consumes step k-E2E_DEPTH's, so the upload and preparation of the next frames overlap the
    integration of frame k; every step's result has been read when the clock stops.  Returns (seconds max
    over ranks, updates)."""
    vol.sync()
    D.barrier()
    t0 = time.perf_counter()
    upd = 0
    posted = []
    for i in range(n):
        call(first + i)
        posted.append(vol.stats_post())
        if len(posted) > E2E_DEPTH:
            upd += vol.stats_wait(posted.pop(0))["updated_last"]
    for t in posted:
        upd += vol.stats_wait(t)["updated_last"]
    vol.sync()
    if after:
        after()
    D.barrier()
    return D.max(time.perf_counter() - t0), D.sum(float(upd))


def run_b200(args):
    from union_thesis_slam_b200 import grid_fusion
    D = Dist()
    torch, world, rank, local, dev = D.torch, D.world, D.rank, D.local, D.dev

    room, frames, K = make_frames(N_FRAMES)
    npix = IM_W * IM_H
    d_frames = [(torch.from_numpy(c).to(dev), torch.from_numpy(d).to(dev), np.linalg.inv(p)) for c, d, p in frames]
    vol = grid_fusion.TSDFVolume(room.copy(), VOXEL, rank=rank, world_size=world, device=local)
    vol.set_batch(args.batch)

    def step(i):
        dc, dd, tinv = d_frames[i % N_FRAMES]
        vol.integrate_device(dc, dd, IM_H, IM_W, K, tinv, 1.0)

    for i in range(args.warmup):
        step(i)
    vol.sync()

    # ---- timed region: frames resident in HBM, device-timed on the launching stream ----------
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.mark()
    ms, upd = timed_device(D, vol, step, args.steps, args.warmup, sampler)
    if sampler:
        sampler.unmark()
    value = upd / (ms * 1e-3)
    # per frame: prep_frame, cull_bricks, 2 x integrate_dense (+ one small memset); with temporal blocking
    # cull and integrate are launched once per batch
    launches_per_step = 1 + 3.0 / args.batch

    # ---- roofline of the dominant kernel (k_integrate_dense) -----------------------------------
    # Each frame launches it twice (two independent halves of the visible bricks on sibling streams,
    # so that one launch's tail is filled by the next); in the timed region the launches run back
    # to back without gaps, so `achieved` divides the algorithmic bytes of the timed frames by the
    # duration of the timed region itself (which also contains the overlapped prepare/cull kernels:
    # conservative).  A second pass times the two launches of every frame alone and serialised
    # (per-frame sync, nothing overlapped) for the per-launch figures.
    peak, peak_src = peaks()
    step_bytes = 24.0 * upd + 12.0 * npix * args.steps * world      # every rank reads the whole frame
    achieved = step_bytes / world / (ms * 1e-3) / 1e9                # per GPU
    vol.set_batch(1)
    vol.set_profiling(True)
    k_ms, k_bytes = 0.0, 0.0
    for i in range(args.steps):
        step(args.warmup + i)
        st = vol.last_stage_ms()
        u = vol.stats()["updated_last"]
        k_ms += st[-1]
        k_bytes += 24.0 * u + 12.0 * npix
    vol.set_profiling(False)
    serial = k_bytes / (k_ms * 1e-3) / 1e9
    launches = 2.0 / args.batch
    roofline = {"bound": "hbm", "kernel": "k_integrate_dense", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "peak_source": peak_src,
                "definition": "algorithmic bytes of the timed frames / duration of the timed region, per GPU "
                              "(the kernel's launches run back to back; overlapped prepare/cull time included)",
                "launches_per_frame": launches,
                "algorithmic_bytes_per_launch": step_bytes / world / args.steps / launches,
                "bytes_model": "24 B x voxels updated + 12 B x pixels (SURVEY 8d)",
                "serialised": {"achieved": serial, "frac": serial / peak,
                               "kernel_ms_per_frame": k_ms / args.steps,
                               "note": "the frame's two launches alone on one stream with a sync per frame: "
                                       "launch gap and both tails exposed"},
                "traffic": None}
    traffic_file = os.path.join(ROOT, "profiles", "dense_traffic.json")
    if os.path.exists(traffic_file):
        with open(traffic_file) as fh:
            tj = json.load(fh)
        # ncu dram__bytes_read+write of one captured launch, scaled to this run's mean launch by the
        # measured traffic/algorithmic ratio (launches differ only in how many bricks they touch)
        roofline["traffic"] = tj["ratio_traffic_over_algorithmic"] * roofline["algorithmic_bytes_per_launch"]
        roofline["traffic_source"] = "profiles/dense_traffic.json: %.2f x algorithmic bytes (ncu --set full, %s)" % (
            tj["ratio_traffic_over_algorithmic"], tj["kernel"])

    # ---- end to end through the drop-in API: pinned host frames in, a result out, every step ----
    h_frames = []
    for c, d, p in frames:
        pc, pd = torch.from_numpy(c).pin_memory(), torch.from_numpy(d).pin_memory()
        raw = torch.from_numpy(np.round(d * 1000.0).astype(np.uint16)).pin_memory()  # the sensor's own millimetres
        h_frames.append((pc.numpy(), pd.numpy(), p, raw.numpy(), pc, pd, raw))

    # Each end-to-end variant runs on a FRESH volume with the same warm-up frames and the same timed frames as
    # the device-timed region above, so its figure is directly comparable with `value` (a frame costs more
    # once its voxels have been observed before: more of them sit in the truncation band).
    # From two ranks on, the frame crosses PCIe once (on rank step % world) and is fanned out over NVLink
    # peer memory (dist.FrameFanout); "per_rank_upload" keeps the figure of every rank uploading every
    # frame itself, which is bounded by the host's PCIe complex.
    from union_thesis_slam_b200 import dist as tdist
    fan_error = []

    def e2e_variant(kind):
        v = grid_fusion.TSDFVolume(room.copy(), VOXEL, rank=rank, world_size=world, device=local)
        fan = None
        if world > 1 and kind in ("f64", "u16") and not fan_error:
            try:
                fan = tdist.FrameFanout(v, IM_H, IM_W, raw_depth=(kind == "u16"), depth_scale=1000., invalid_raw=65535)
            except Exception as exc:  # e.g. CUDA IPC not permitted in this container: every rank uploads for itself
                fan_error.append(repr(exc))

        def call(i):
            c, d, p, raw = h_frames[i % N_FRAMES][:4]
            if fan is not None:
                fan.integrate(c, raw if kind == "u16" else d, K, p, obs_weight=1.)
            elif kind == "u16":
                v.integrate_raw(c, raw, K, p, obs_weight=1., depth_scale=1000., invalid_raw=65535)
            else:
                v.integrate(c, d, K, p, obs_weight=1.)

        for i in range(args.warmup):
            call(i)
        if kind == "blocking":  # integrate() + stats() per step: the host waits for the device every step
            v.sync()
            D.barrier()
            t0 = time.perf_counter()
            n_upd = 0
            for i in range(args.steps):
                call(args.warmup + i)
                n_upd += v.stats()["updated_last"]
            v.sync()
            D.barrier()
            res = D.max(time.perf_counter() - t0), D.sum(float(n_upd))
        else:
            res = e2e_loop(D, v, call, args.steps, args.warmup, after=fan.drain if fan else None)
        if fan is not None:
            fan.close()
        v.close()
        return res

    vol.close()
    e2e_s, e2e_upd = e2e_variant("f64")
    raw_s, raw_upd = e2e_variant("u16")
    blk_s, blk_upd = e2e_variant("blocking")
    fan_note = "" if world == 1 else (
        " -- through dist.FrameFanout: frame k is uploaded by rank k %% %d alone and pulled by every peer's copy engine into "
        "its own ring slot over NVLink (%d x the frame's bytes per step), every rank integrates its own sheets" % (world, world - 1))
    e2e = {"value": e2e_upd / e2e_s, "unit": "voxel-updates/s", "h2d_bytes_per_step": npix * (8 + 3),
           "d2h_bytes_per_step": 64 * world, "frames_per_s": args.steps / e2e_s,
           "api": "TSDFVolume.integrate(color, depth f64, K, pose) every step from pinned host buffers (3.38 MB H2D "
                  "inside the call) + stats_post(); the 64-byte counter block of step k-%d is read (stats_wait) "
                  "while step k runs, all of them before the clock stops%s" % (E2E_DEPTH, fan_note),
           "raw_u16": {"value": raw_upd / raw_s, "frames_per_s": args.steps / raw_s,
                       "h2d_bytes_per_step": npix * (2 + 3),
                       "api": "TSDFVolume.integrate_raw: the sensor's uint16 depth, divided and masked on the device"},
           "blocking_read": {"value": blk_upd / blk_s, "frames_per_s": args.steps / blk_s,
                             "h2d_bytes_per_step": npix * (8 + 3) * world,
                             "api": "integrate() + stats() per step: the host waits for the device every step"
                                    + (" (every rank uploads the frame itself)" if world > 1 else "")}}
    if fan_error:
        e2e["fanout_error"] = fan_error[0]
        e2e["api"] = e2e["api"].replace(fan_note, " -- every rank uploads the frame itself (fan-out unavailable)")
        e2e["h2d_bytes_per_step"] = npix * (8 + 3) * world
    elif world > 1:
        e2e["nvlink_bytes_per_step"] = npix * (8 + 3) * (world - 1)
    if world > 1:
        pr_s, pr_upd = e2e_variant("per_rank")
        e2e["per_rank_upload"] = {"value": pr_upd / pr_s, "frames_per_s": args.steps / pr_s,
                                  "h2d_bytes_per_step": npix * (8 + 3) * world,
                                  "api": "every rank calls TSDFVolume.integrate with the host frame itself (no fan-out)"}
    clocks = sampler.stop() if sampler else None

    line = {
        "metric": "tsdf_voxel_updates_per_s", "value": value, "unit": "voxel-updates/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(workload_config(world), frames_per_brick_visit=args.batch),
        "frames_per_s": args.steps / (ms * 1e-3),
        "voxel_updates_per_step": upd / args.steps, "roofline": roofline, "e2e": e2e,
        "gpu_launches": int(round(launches_per_step * args.steps)), "clocks": clocks,
    }
    del d_frames, vol
    torch.cuda.empty_cache()

    if not args.no_configs:
        try:
            line["configs"] = bench_configs(args, D)
        except Exception as exc:  # the extra configurations must never hide the headline
            line["configs"] = {"error": repr(exc)}
    if not args.no_hash:
        try:
            line["hash"] = bench_hash(args, D)
        except Exception as exc:
            line["hash"] = {"error": repr(exc)}
    if rank == 0 and world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baselines(frames, K, room)
    if rank == 0:
        print(json.dumps(line), file=_OUT, flush=True)
    D.close()
    return 0


def cpu_baselines(frames, K, room):
    """The reference's own numba CPU path (oracle/_ref, kind "reference") on a bounded sample of the headline
    workload, with its C/OpenMP restatement ("port") timed beside it; only the port when oracle/_ref is absent."""
    keep = ("value", "unit", "cores", "kind", "sample", "cpu_model", "host_cpus", "voxel_visits_per_s",
            "threading_layer", "numba", "numpy", "frames_per_s_sample")
    port = cpu_port(frames, K, room, VOXEL, steps=4, warmup=1)
    out = {k: port[k] for k in keep if k in port}
    if reference_built():
        try:
            ref = cpu_reference(frames, K, room, VOXEL, steps=3, warmup=1, slabs=4)
            out = {k: ref[k] for k in keep if k in ref}
            out["port"] = {k: port[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as exc:
            out["reference_error"] = repr(exc)
    return out


def fixture_frames():
    """The first 100 frames of the bundled sequence kept under tests/golden/frames (in order, as the reference's demo
    integrates them), decoded as grid_demo1.py:80-84."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import frames as fx
    import scenes
    idx = [i for i in range(100) if os.path.exists(os.path.join(fx.FRAMES_DIR, "frame-%06d.depth.png" % i))]
    return [fx.load_frame(i, check=False) for i in idx], fx.intrinsics(), np.array(scenes.CONFIG1_BOUNDS)


def bench_configs(args, D):
    """The other dense configurations of BASELINE.json, device-timed with frames resident, each with its
    roofline figure: configs[0] (bundled frames, 2 cm; launch-bound, so also with 4 frames per brick
    visit), configs[2] (bundled frames, 5 mm) and the 48 GB reading of configs[3] (4.64 mm room)."""
    from union_thesis_slam_b200 import grid_fusion, synth
    torch, dev = D.torch, D.dev
    peak, _ = peaks()
    out = {}

    def run(name, bounds, voxel, frs, Kc, batch, passes, note):
        H, W = frs[0][1].shape
        dfr = [(torch.from_numpy(np.ascontiguousarray(c)).to(dev), torch.from_numpy(d).to(dev), np.linalg.inv(p))
               for c, d, p in frs]
        vol = grid_fusion.TSDFVolume(bounds.copy(), voxel, rank=D.rank, world_size=D.world, device=D.local)
        vol.set_batch(batch)

        def step(i):
            dc, dd, ti = dfr[i % len(dfr)]
            vol.integrate_device(dc, dd, H, W, Kc, ti, 1.0)

        for i in range(len(dfr)):       # first pass: every voxel of the trajectory seen once
            step(i)
        n = passes * len(dfr)
        ms, upd = timed_device(D, vol, step, n, 0)
        dim = [int(x) for x in vol._vol_dim]
        vol.close()
        del vol, dfr
        torch.cuda.empty_cache()
        gbs = (24.0 * upd + 12.0 * H * W * n * D.world) / D.world / (ms * 1e-3) / 1e9
        out[name] = {"workload": note, "grid": dim, "voxel_size_m": voxel, "frames_per_brick_visit": batch,
                     "steps": n, "ms_per_step": ms / n, "frames_per_s": n / (ms * 1e-3), "value": upd / (ms * 1e-3),
                     "unit": "voxel-updates/s", "voxel_updates_per_step": upd / n, "n_gpus": D.world,
                     "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak}}

    if D.world == 1:
        frs, Kf, bounds = fixture_frames()
        for batch in (1, 4):
            run("config0_2cm_batch%d" % batch, bounds, 0.02, frs, Kf, batch, 2,
                "BASELINE.json configs[0]: the first %d frames of the bundled 640x480 sequence (tests/golden/frames) in "
                "order, replayed, dense grid at 2 cm over the sequence's bounds" % len(frs))
        run("config2_5mm", bounds, 0.005, frs, Kf, 1, 1,
            "BASELINE.json configs[2]: the same frames replayed into a 5 mm grid over the same bounds (23.6 GB)")
    room = synth.room_bounds(ROOM)
    frs, Kr = synth.room_stream(8, room, seed=0)
    run("config3_48GB_reading", room, 0.00464, frs, Kr, 1, 1,
        "BASELINE.json configs[3] read as '~4e9 voxels, ~48 GB': the same 10x10x4 m room at 4.64 mm, "
        "first 8 frames of the headline trajectory")
    return out


def bench_hash(args, D):
    """Voxel-block hash map, BASELINE.json configs[4]-shaped: 1280x720 synthetic depth stream on a
    seeded trajectory through a 40 x 40 x 6 m room at 2 cm voxels (8^3 blocks; the bounded block grid
    holds 2.4 M blocks), blocks sharded over the ranks by owner hash (every rank walks every frame, no
    per-frame collective).  Frames resident in HBM; whole-pipeline device time; then end to end."""
    from union_thesis_slam_b200 import hash_fusion, synth
    torch, dev, world, rank, local = D.torch, D.dev, D.world, D.rank, D.local
    peak, _ = peaks()
    room = synth.room_bounds(HASH_ROOM)
    frames, K = synth.room_stream(N_FRAMES, room, width=HASH_W, height=HASH_H, seed=1)
    nblocks = int(np.prod(np.ceil(np.array(HASH_ROOM) / HASH_VOXEL / 8)))
    per_rank = int(nblocks / world * 1.25) + 1024  # owner hash balances to a few percent
    ht = hash_fusion.HashTable(room.copy(), HASH_VOXEL, map_size=4 * per_rank, heap_blocks=per_rank, device=local,
                               rank=rank, world_size=world)
    d_frames = [(torch.from_numpy(c).to(dev), torch.from_numpy(d).to(dev), np.linalg.inv(p)) for c, d, p in frames]
    torch.cuda.synchronize()
    steps = min(args.steps, 256)

    def step(i):
        dc, dd, ti = d_frames[i % N_FRAMES]
        ht.integrate_device(dc, dd, HASH_H, HASH_W, K, ti, 1.0)

    for i in range(max(args.warmup, N_FRAMES)):  # one full pass allocates the blocks of the trajectory
        step(i)
    ht.sync()
    s0 = ht.stats()
    stream = torch.cuda.ExternalStream(ht.cuda_stream(), device=dev)
    D.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(steps):
        step(i)
    e1.record(stream)  # the frame's integrate kernel is the last kernel of its pipeline
    ht.sync()
    D.barrier()
    wall = D.max(e0.elapsed_time(e1) * 1e-3)
    s1 = ht.stats()
    upd = D.sum(float(s1["updated_total"] - s0["updated_total"]))
    # algorithmic bytes of the same frames (second pass, per-frame counters)
    tot_bytes = 0.0
    for i in range(steps):
        step(i)
        st = ht.stats()
        tot_bytes += 24.0 * st["updated_last"] + 16.0 * (st["blocks_visited_last"] + st["keys_inserted_last"]) \
            + 12.0 * HASH_W * HASH_H
    tot_bytes = D.sum(tot_bytes)
    blocks_all = D.sum(float(s1["blocks"]))
    achieved = tot_bytes / wall / 1e9 / world  # per GPU, against one GPU's HBM peak

    # end to end: HashTable.integrate with pinned host frames, a pipelined per-step read-back (see e2e_loop)
    h_frames = []
    for c, d, p in frames:
        pc, pd = torch.from_numpy(c).pin_memory(), torch.from_numpy(d).pin_memory()
        h_frames.append((pc.numpy(), pd.numpy(), p, pc, pd))

    fan = None
    if world > 1:  # the frame is uploaded once and fanned out over NVLink (dist.FrameFanout)
        from union_thesis_slam_b200 import dist as tdist
        try:
            fan = tdist.FrameFanout(ht, HASH_H, HASH_W)
        except Exception:
            fan = None

    def call(i):
        c, d, p = h_frames[i % N_FRAMES][:3]
        if fan is not None:
            fan.integrate(c, d, K, p, obs_weight=1.)
        else:
            ht.integrate(c, d, K, p, obs_weight=1.)

    for i in range(2):
        call(i)
    e2e_s, e2e_upd = e2e_loop(D, ht, call, steps, 2, after=fan.drain if fan else None)
    if fan is not None:
        fan.close()
    out = {"metric": "tsdf_voxel_updates_per_s", "value": upd / wall, "unit": "voxel-updates/s", "n_gpus": world,
           "blocks_allocated_all_ranks": int(blocks_all),
           "frames_per_s": steps / wall, "ms_per_step": 1e3 * wall / steps, "steps": steps,
           "workload": "BASELINE.json configs[4]: 1280x720 stream, 40x40x6 m room, 2 cm voxels, 8^3 blocks",
           "voxel_size_m": HASH_VOXEL, "blocks_allocated": s1["blocks"], "table_keys": s1["table_keys"],
           "table_slots": s1["table_slots"], "heap_capacity": s1["heap_capacity"],
           "voxel_updates_per_step": upd / steps, "scaling": "strong",
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "bytes_model": "24 B x voxels updated + 16 B x unique keys (visible + inserted) + 12 B x "
                                       "pixels, over the device time (CUDA events, max over ranks) of the whole "
                                       "pipeline: prepare, allocation (super-block scan, refine + lock-free insert), "
                                       "compaction, integrate",
                        "traffic": None},
           "e2e": {"value": e2e_upd / e2e_s, "unit": "voxel-updates/s", "frames_per_s": steps / e2e_s,
                   "h2d_bytes_per_step": HASH_W * HASH_H * (8 + 3) * (1 if fan is not None else world),
                   "d2h_bytes_per_step": 96 * world,
                   "nvlink_bytes_per_step": HASH_W * HASH_H * (8 + 3) * (world - 1) if fan is not None else 0,
                   "api": "HashTable.integrate(color, depth f64, K, pose) from pinned host buffers + stats_post(); "
                          "step k-%d's counters are read while step k runs%s" % (
                              E2E_DEPTH, "; every rank uploads the frame itself" if fan is None else
                              "; frame k is uploaded by rank k % world alone and pulled by the peers' copy engines "
                              "over NVLink (dist.FrameFanout)")}}
    tf = os.path.join(ROOT, "profiles", "hash_traffic.json")
    if os.path.exists(tf):
        with open(tf) as fh:
            tj = json.load(fh)
        out["roofline"]["traffic"] = tj["ratio_traffic_over_algorithmic"] * tot_bytes / world / steps
        out["roofline"]["traffic_source"] = "profiles/hash_traffic.json: %.2f x algorithmic bytes per frame (ncu, %s)" % (
            tj["ratio_traffic_over_algorithmic"], tj["kernel"])
    ht.close()
    del d_frames
    torch.cuda.empty_cache()
    if rank == 0 and world == 1 and not args.no_cpu:
        out["cpu_baseline"] = cpu_hash_port(frames[:3], K)
    return out


def cpu_hash_port(frames, K, room_size=(8.0, 8.0, 3.0)):
    """CPU side of the hash line: the reference HashTable restated per voxel (oracle.OracleHashVox, C/OpenMP) on a
    room the host can hold -- the reference's own HashTable.integrate is a pure-Python loop at ~2e4 voxel-updates/s
    (BASELINE.md), timed here on one frame of a small room when oracle/_ref is present."""
    from oracle import oracle
    from union_thesis_slam_b200 import synth
    room = synth.room_bounds(room_size)
    fr, Kc = synth.room_stream(3, room, width=HASH_W, height=HASH_H, seed=5)
    orh = oracle.OracleHashVox(room.copy(), HASH_VOXEL)
    orh.integrate(fr[0][0], fr[0][1], Kc, fr[0][2])
    t0 = time.perf_counter()
    upd = 0
    for c, d, p in fr[1:]:
        upd += orh.integrate(c, d, Kc, p)
    dt = time.perf_counter() - t0
    out = {"value": upd / dt, "unit": "voxel-updates/s", "cores": int(oracle.lib().oracle_max_threads()), "kind": "port",
           "sample": "%gx%gx%g m room at 2 cm, 1280x720, 2 frames, per-voxel float64 entries (oracle.OracleHashVox)" % room_size}
    if reference_built():
        try:
            import contextlib
            import io
            from oracle import build_ref
            _, hf = build_ref.load()
            small = synth.room_bounds((2.0, 2.0, 2.0))
            frs, Ks = synth.room_stream(2, small, width=HASH_W, height=HASH_H, seed=5)
            with contextlib.redirect_stdout(io.StringIO()):
                ref = hf.HashTable(small.copy(), HASH_VOXEL, map_size=2000000)
                t0 = time.perf_counter()
                ref.integrate(frs[0][0], frs[0][1], Ks, frs[0][2])
                dt = time.perf_counter() - t0
                n = ref.count_num_hash_entries()
            out = {"value": n / dt, "unit": "voxel-updates/s", "cores": 1, "kind": "reference",
                   "sample": "reference HashTable.integrate (pure-Python per-voxel loop), one 1280x720 frame into a "
                             "2x2x2 m room at 2 cm: %d voxel entries in %.1f s" % (n, dt),
                   "port": out}
        except Exception as exc:
            out["reference_error"] = repr(exc)
    return out


def main():
    # the drop-in constructors print the reference's banners; keep stdout for the one JSON line
    real_stdout = sys.stdout
    sys.stdout = sys.stderr
    try:
        return _main(real_stdout)
    finally:
        sys.stdout = real_stdout


def _main(out):
    global _OUT
    _OUT = out
    sys.setswitchinterval(0.0005)  # let the clock-sampling thread run inside millisecond-long timed regions
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1024)
    ap.add_argument("--warmup", type=int, default=32)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-hash", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the extra dense configurations (configs 0, 2, 48 GB)")
    ap.add_argument("--batch", type=int, default=1, choices=[1, 2, 3, 4],
                    help="frames applied per brick visit in the device-timed region (temporal blocking)")
    ap.add_argument("--port", action="store_true", help="--impl reference: time the C/OpenMP port instead of oracle/_ref")
    ap.add_argument("--workload", default="dense", choices=["dense", "hash"],
                    help="hash: print the hash-map line alone (works under torchrun: blocks sharded by owner hash)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "hash":
        return run_hash_only(args)
    return run_b200(args)


def run_hash_only(args):
    """Hash map alone, optionally sharded over the ranks of a torchrun launch."""
    D = Dist()
    line = bench_hash(args, D)
    line.update({"higher_is_better": True, "data": "synthetic", "dtype": "f64"})
    if D.rank == 0:
        print(json.dumps(line), file=_OUT, flush=True)
    D.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
