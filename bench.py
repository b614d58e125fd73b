#!/usr/bin/env python
"""bench.py -- RAISR 2x output Mpix/s on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 3|2|5] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of synthetic luma frames (u8 in, u8 out).  The default workload
is the size north_star quotes its target on, BASELINE.json configs[2]: 4K -> 8K, 32 frames per GPU per step
(`--config 2` = configs[1], 1080p -> 4K x 64 frames; `--config 5` = configs[4], 720p x 256 frames).  With N GPUs every
rank processes its own frames (frames are independent, no collective: weak scaling); `value` is the whole-job
output Mpix/s.

  value      device-resident throughput (inputs already in HBM), CUDA events, max over ranks
  e2e        same metric through the C-ABI with HOST (pinned) buffers: H2D and D2H inside the call; next to it the
             raw copy ceiling of the box (the same bytes moved by bare cudaMemcpyAsync, all ranks at once)
  roofline   FP32-FFMA roofline of SURVEY.md 8(d): 412 algorithmic FLOP per output pixel over the summed kernel
             time of the step (CUDA events around every launch, on the launch stream), against 2*128*SMs*max_clock and
             a measured register-only FFMA kernel; ncu-derived figures are read from profiles/ncu_kernel_metrics.json
             (with the capture they come from), never hard-coded
  variants   the same step with fp32 and fp16 tap records (the default is "auto" -> 24-bit records), with each
             variant's maximum output deviation from the fp32-tap result
  band       BASELINE.json configs[3]: one 16384x16384 image, 3x, row-banded over the N ranks with the halo rows
             read from the neighbour's memory over NVLink; device ms, Gpix/s, halo bytes, output checksum
  next_rows  the byte-stream kernels of SURVEY.md 8(f) against the HBM roofline (N = 1 only)
  cpu_baseline  the oracle's C port (oracle/raisr_oracle.c) on the host cores of this box -- the reference has no CPU
             RAISR path and its OpenCL path cannot run
`--impl reference` times that CPU port alone (there is nothing else of the reference to time).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SCALE = 2
CONFIGS = {   # BASELINE.json config number (1-based) -> source size, frames per GPU per step
    2: dict(sw=1920, sh=1080, frames=64, name="1080p->4K", baseline="configs[1]"),
    3: dict(sw=3840, sh=2160, frames=32, name="4K->8K", baseline="configs[2] (north_star target size; 32 of its 512 frames per GPU per step)"),
    5: dict(sw=1280, sh=720, frames=256, name="720p->1440p", baseline="configs[4] (256 of its 2048 frames per GPU per step)"),
}
FLOP_PER_PX = 412.0                             # SURVEY.md 8(d), minimal form: whole path
FLOP_PER_PX_FILTER = 244.0                      # 121 FMA + store: the dominant kernel's share of the 412
BYTES_PER_PX = 1.0 / (SCALE * SCALE) + 1.0      # u8 in -> u8 out


def make_inputs(cfg, n_frames, rank):
    from oclcomputervision_b200 import synth
    pool = min(4, n_frames)
    frames = synth.synthetic_batch(pool, cfg["sh"], cfg["sw"], pool=pool, seed=1000 + 16 * rank)
    reps = (n_frames + pool - 1) // pool
    return np.ascontiguousarray(np.tile(frames, (reps, 1, 1))[:n_frames])


def cpu_oracle_rate(cfg, budget_s, steps=1, warm=0):
    """C oracle on all host threads over a bounded sample of the workload; returns (Mpix/s, threads, frames per step, seconds)."""
    from oracle import raisr_oracle as O
    from oclcomputervision_b200 import synth
    sw, sh = cfg["sw"], cfg["sh"]
    F = synth.random_filters(SCALE)
    frame = synth.synthetic_frame(sh, sw, 1000)
    threads = len(os.sched_getaffinity(0))   # torchrun exports OMP_NUM_THREADS=1; the oracle sets its own count
    t0 = time.perf_counter()
    O.raisr_ref_c(frame, F, SCALE, nthreads=threads, want=("out_u8",))
    one = time.perf_counter() - t0
    n = int(max(1, min(64, budget_s / max(one, 1e-3))))

    def step():
        for _ in range(n):
            O.raisr_ref_c(frame, F, SCALE, nthreads=threads, want=("out_u8",))
    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return steps * n * sw * sh * SCALE * SCALE / dt / 1e6, threads, n, dt


def cpu_baseline(cfg, budget_s=12.0):
    mpix, threads, n, dt = cpu_oracle_rate(cfg, budget_s)
    return dict(value=round(mpix, 3), unit="Mpix/s", cores=threads, kind="port",
                sample="%d frame(s) of the %s workload (%dx%d source), C oracle (fp32, OpenMP), %.1f s" %
                       (n, cfg["name"], cfg["sw"], cfg["sh"], dt))


class ClockSampler:
    """SM clock / throttle reasons of one GPU DURING the timed region.  NVML in a thread (a sample every
    few ms, so even a 50 ms region is covered); falls back to an `nvidia-smi -lms` child process."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, uuid=None, period_s=0.004):
        self.rows, self.p, self.thread, self.nv = [], None, None, None
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            if uuid:
                try:
                    h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
                except Exception:
                    h = None
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(int(index))
            self.nv, self.h = pynvml, h
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._loop, args=(period_s,), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nv = None
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def _loop(self, period_s):
        nv, h = self.nv, self.h
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
        while not self._stop.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                r = int(get_reasons(h))
                self.rows.append((sm, pw, [k for k, b in bits.items() if r & b]))
            except Exception:
                pass
            self._stop.wait(period_s)

    def stop(self):
        if self.nv is not None:
            self._stop.set()
            self.thread.join(timeout=2)
            if not self.rows:
                return dict(sm_mhz=None, sm_max_mhz=self.mx, reasons=["no samples"])
            sm = [r[0] for r in self.rows]; pw = [r[1] for r in self.rows]
            reasons = sorted({k for r in self.rows for k in r[2]})
            busy = [s_ for s_, p_ in zip(sm, pw) if p_ > 0.5 * max(pw)] or sm
            return dict(sm_mhz=float(np.median(busy)), sm_max_mhz=self.mx, samples=len(sm), power_w_max=float(max(pw)),
                        reasons=reasons, source="nvml")
        if self.p is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, mx, pw, reasons = [], [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        busy = [s_ for s_, p_ in zip(sm, pw) if p_ > 0.5 * max(pw)] or sm
        return dict(sm_mhz=float(np.median(busy)), sm_max_mhz=float(max(mx)), samples=len(sm),
                    power_w_max=float(max(pw)), reasons=sorted(reasons), source="nvidia-smi")


def bind_to_gpu_numa_node(index):
    """Pin this rank's host threads (and thereby its pinned staging buffers, first touch) to the CPUs NVML
    lists as local to its GPU, restricted to the CPUs the container allows.  Returns a short description."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(index))
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        local = {64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        pick = sorted(local & allowed)
        if pick and len(pick) < len(allowed):
            os.sched_setaffinity(0, pick)
            return "%d of %d allowed CPUs (GPU-local)" % (len(pick), len(allowed))
        return "no change (%d GPU-local CPUs among %d allowed)" % (len(pick), len(allowed))
    except Exception as e:  # affinity is an optimisation, never a requirement
        return "unavailable: %s" % type(e).__name__


def workload_text(cfg, n, note=""):
    return ("RAISR 2x %s luma u8->u8, batch of %d synthetic frames per GPU per step, random-init 24x3x3x4x121 table "
            "(BASELINE %s)%s" % (cfg["name"], n, cfg["baseline"], note))


def ref_kernel_via_shim():
    """The reference's own kernel text, compiled for the CPU by oracle/build_ref.py (oracle/_ref), timed on a small frame.
    It is a work-item EMULATOR (256 host threads per work-group, a pthread barrier per barrier()): a checker, not a CPU
    implementation -- reported next to the baseline for information, never used as the baseline."""
    try:
        import time
        from oracle import raisr_cl_ref as R
        from oclcomputervision_b200 import synth
        if not R.available():
            return None
        src = synth.synthetic_frame(128, 128, seed=11)
        flt = synth.random_filters(2)
        R.run(src[:16, :16], flt, 2, kind="intended", prec="f32")
        t0 = time.perf_counter()
        R.run(src, flt, 2, kind="intended", prec="f32")
        dt = time.perf_counter() - t0
        return dict(value=round(256 * 256 / dt / 1e6, 4), unit="Mpix/s", sample="one 128x128 -> 256x256 frame, %.2f s" % dt,
                    what="raisr.cl (early return off, its three slips corrected) through oracle/ref_shim: work-item emulation")
    except Exception as e:      # informational only
        return dict(error=str(e)[:200])


def run_reference(args, cfg, rank, world):
    """Reference arm: the reference has no CPU RAISR path; its OpenCL kernel runs here only through the work-item
    emulation of oracle/_ref (a checker, ~0.1 Mpix/s), so this times the oracle's C port of raisr.cl on all host threads
    -- the FASTER of the two, hence the conservative baseline.  One step = a bounded sample of the bench workload (as
    many frames as fit ~2 s); W warm-up steps, K timed steps."""
    if rank != 0:
        return
    steps, warm = max(1, args.steps), max(0, args.warmup)
    step_budget = min(2.0, 150.0 / (steps + warm))
    mpix, threads, n, dt = cpu_oracle_rate(cfg, step_budget, steps, warm)
    cb = dict(value=round(mpix, 3), unit="Mpix/s", cores=threads, kind="port",
              sample="%d step(s) of %d frame(s) of the %s workload (%dx%d source), C oracle (fp32, OpenMP), %.1f s" %
                     (steps, n, cfg["name"], cfg["sw"], cfg["sh"], dt))
    line = dict(metric="RAISR 2x output Mpix/s", value=cb["value"], unit="Mpix/s", impl="reference",
                n_gpus=args.gpus, steps=steps, warmup=warm, ms_per_step=round(dt / steps * 1e3, 3),
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload=workload_text(cfg, n, "; bounded sample: %d frame(s) per step" % n),
                            note="the reference has no CPU RAISR path; its OpenCL kernel text runs on a CPU only through the "
                                 "work-item emulation of oracle/_ref (ref_kernel_via_shim below, a checker); timed here is the "
                                 "oracle's C port of raisr.cl on the host cores, which is pinned to that kernel's outputs"),
                cpu_baseline=cb, ref_kernel_via_shim=ref_kernel_via_shim(), gpu_launches=0,
                e2e=dict(value=cb["value"], unit="Mpix/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def ncu_metrics(kernel, taps):
    """Figures that only a profiler can give (shared-memory wavefronts, DRAM bytes), from the committed capture."""
    try:
        table = json.load(open(os.path.join(ROOT, "profiles", "ncu_kernel_metrics.json")))
        return table.get("%s/%s" % (kernel, taps))
    except Exception:
        return None


def copy_ceiling(torch, h2d_bytes, d2h_bytes, barrier, reps=3):
    """The same bytes as one e2e step moved by bare cudaMemcpyAsync from / to pinned memory on two streams, every rank
    at once: what the box's host links give without any kernel in the way.  Returns seconds per step (this rank)."""
    hs = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory()
    hd = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
    ds = torch.empty(h2d_bytes, dtype=torch.uint8, device="cuda")
    dd = torch.empty(d2h_bytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def once():
        with torch.cuda.stream(s1):
            ds.copy_(hs, non_blocking=True)
        with torch.cuda.stream(s2):
            hd.copy_(dd, non_blocking=True)
    once()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def band_record(torch, dist, r, rank, world, local_rank):
    """BASELINE.json configs[3]: 16384x16384 -> 49152x49152 (3x, nine pixel types), one image, row bands over the ranks,
    halo rows read from the neighbours over NVLink P2P (multi_gpu.BandedUpscaler, C-ABI only)."""
    from oclcomputervision_b200 import synth, _cabi
    from oclcomputervision_b200 import multi_gpu as mg
    lib = _cabi.load()
    sw = sh = 16384
    s = 3
    r.filters_x3 = synth.random_filters(3)
    tile = synth.synthetic_frame(2048, 2048, 4242)          # the image is this tile repeated 8 x 8: cheap and rank-independent
    r.set_stream(0)
    up = mg.BandedUpscaler(r, sw, sh, s)
    me = up.me
    own = me.own_last - me.own_first + 1
    hsrc, hdst = ctypes.c_void_p(), ctypes.c_void_p()
    _cabi.check(lib.raisr_host_alloc(ctypes.byref(hsrc), max(1, own) * sw))
    _cabi.check(lib.raisr_host_alloc(ctypes.byref(hdst), max(1, me.dst_rows) * sw * s))
    src = np.ctypeslib.as_array((ctypes.c_ubyte * (own * sw)).from_address(hsrc.value)).reshape(own, sw)
    rows = (np.arange(me.own_first, me.own_last + 1) % 2048)
    src[:] = np.tile(tile[rows], (1, 8))
    out = np.ctypeslib.as_array((ctypes.c_ubyte * (me.dst_rows * sw * s)).from_address(hdst.value)).reshape(me.dst_rows, sw * s)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    # end to end once (H2D of the owned rows, halo, kernels, D2H), then device-resident repeats
    up.enqueue(hsrc.value, sw, hdst.value, sw * s)
    first = up.finish()
    times = []
    for _ in range(3):
        barrier()
        up.enqueue(0, 0, 0, 0)
        t = up.finish()
        times.append(t["halo"] + t["kernels"])
    dev_ms = float(np.median(times))
    barrier()
    t0 = time.perf_counter()
    up.enqueue(hsrc.value, sw, hdst.value, sw * s)
    e2e = up.finish()
    e2e_wall = time.perf_counter() - t0
    # position-sensitive checksum of the whole output image, independent of how it was banded
    grow = np.arange(me.dst_row0, me.dst_row0 + me.dst_rows, dtype=np.uint64) + np.uint64(1)
    col = np.arange(sw * s, dtype=np.uint64) + np.uint64(1)
    row_sums = out.sum(axis=1, dtype=np.uint64)
    col_sums = out.sum(axis=0, dtype=np.uint64)
    with np.errstate(over="ignore"):
        c1 = int((row_sums * grow).sum(dtype=np.uint64)) & 0xFFFFFFFFFFFF
        c2 = int((col_sums * col).sum(dtype=np.uint64)) & 0xFFFFFFFFFFFF
    stats = torch.tensor([dev_ms, e2e["total"], e2e_wall * 1e3, first["total"]], device="cuda", dtype=torch.float64)
    sums = torch.tensor([c1, c2, up.halo_bytes], device="cuda", dtype=torch.int64)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    up.close()
    lib.raisr_host_free(hsrc); lib.raisr_host_free(hdst)
    px = (sw * s) * (sh * s)
    dev_ms, e2e_ms, e2e_wall_ms, first_ms = [float(x) for x in stats.tolist()]
    c1, c2, halo = [int(x) for x in sums.tolist()]
    return dict(workload="RAISR 3x 16384x16384 -> 49152x49152 (BASELINE configs[3]), one image row-banded over %d rank(s)" % world,
                bands=world, band_rows_out=me.dst_rows, halo_rows_per_side_max=max([b.halo_above for b in up.bands] + [b.halo_below for b in up.bands]),
                halo_bytes=halo, halo_transport="cudaMemcpy2DAsync from the neighbour's IPC-mapped window (NVLink P2P), ordered by device flags; no collective",
                device_ms=round(dev_ms, 3), device_gpix_s=round(px / dev_ms / 1e6, 2), device_timing="halo wait + halo copies + kernels, CUDA events on the stream, median of 3, max over ranks",
                e2e_ms=round(e2e_ms, 3), e2e_gpix_s=round(px / e2e_ms / 1e6, 2), e2e_wall_ms=round(e2e_wall_ms, 3),
                e2e_note="H2D of the owned source rows + halo + kernels + D2H of the band, pinned host memory, max over ranks",
                roofline_frac=round(FLOP_PER_PX * px / (dev_ms * 1e-3) / 1e12 / (world * 74.45), 4),
                checksum="%012x%012x" % (c1 & 0xFFFFFFFFFFFF, c2 & 0xFFFFFFFFFFFF),
                checksum_note="sum over the whole output of byte*(row+1) and byte*(col+1), mod 2^48 per rank then summed: equal for every N")


def next_rows_record(torch, lib, h, peak):
    """SURVEY.md 8(f) byte-stream kernels against the HBM roofline: algorithmic bytes / median kernel time."""
    from oclcomputervision_b200 import _cabi
    out = {}
    ms = (ctypes.c_float * 3)()
    rng = np.random.default_rng(0)

    def timed(fn, reps=10):
        ts = []
        for i in range(reps + 3):
            _cabi.check(fn())
            if i >= 3:
                ts.append(ms[1])
        return float(np.median(ts))
    sw, sh, frames = 1920, 1080, 16
    dw, dh = 2 * sw, 2 * sh
    for ch in (1, 4):
        src = torch.from_numpy(rng.integers(0, 256, (frames, sh, sw * ch), dtype=np.uint8)).cuda()
        dst = torch.empty((frames, dh, dw * ch), dtype=torch.uint8, device="cuda")
        for mode, name in ((0, "bilinear_lds"), (1, "bicubic")):
            t = timed(lambda: lib.raisr_resize_u8(h, ctypes.c_void_p(src.data_ptr()), sw, sh, sw * ch, ch, ctypes.c_void_p(dst.data_ptr()),
                                                  dw, dh, dw * ch, mode, frames, _cabi.RAISR_DEVICE, ms))
            nbytes = frames * (sw * sh + dw * dh) * ch
            out["resize_%s_%s" % (name, "gray" if ch == 1 else "bgra")] = dict(
                ms=round(t, 4), out_gpix_s=round(frames * dw * dh / t / 1e6, 1), achieved_gbs=round(nbytes / t / 1e6, 1),
                frac=round(nbytes / t / 1e6 / peak, 3))
        del src, dst
    n = 16384
    img = torch.from_numpy(rng.integers(0, 256, (n, n), dtype=np.uint8)).cuda()
    dst = torch.empty((n, n), dtype=torch.uint8, device="cuda")
    nx = ny = n // 256
    hist = torch.empty((n // 32, nx, 256), dtype=torch.int32, device="cuda")
    maps = torch.from_numpy((rng.random((ny, nx, 256)) * 255).astype(np.float32)).cuda()
    lut = torch.from_numpy(rng.permutation(256).astype(np.uint8)).cuda()
    P = lambda t_: ctypes.c_void_p(t_.data_ptr())   # noqa: E731
    for name, fn, nbytes in (
            ("hist_tiles", lambda: lib.ocv_hist_grid_u8(h, P(img), n, n, n, P(hist), _cabi.RAISR_DEVICE, ms), n * n + n // 32 * nx * 1024),
            ("lut_apply", lambda: lib.ocv_histeq_global_u8(h, P(img), n, n, n, P(dst), n, P(lut), _cabi.RAISR_DEVICE, ms), 2 * n * n),
            ("lut_blend", lambda: lib.ocv_histeq_local_block_u8(h, P(img), n, n, n, P(dst), n, P(maps), nx, ny, 256, 256, _cabi.RAISR_DEVICE, ms), 2 * n * n)):
        t = timed(fn)
        out[name] = dict(ms=round(t, 4), achieved_gbs=round(nbytes / t / 1e6, 1), frac=round(nbytes / t / 1e6 / peak, 3))
    out["note"] = ("algorithmic bytes (source + destination once) / median kernel time of 10, device-resident; resize: 16 frames 1080p->4K; "
                   "histeq: one 16384x16384 image; frac = of the measured HBM copy bandwidth %.0f GB/s" % peak)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=sorted(CONFIGS), help="BASELINE.json config number: 3 = 4K->8K (default), 2 = 1080p->4K, 5 = 720p")
    ap.add_argument("--frames", type=int, default=None, help="frames per GPU per step (default: per config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the variants / band / next_rows sub-records")
    ap.add_argument("--filter-impl", type=int, default=None, help="1 octet (default), 0 block")
    ap.add_argument("--prep-ctas", type=int, default=None, help="persistent prep grid: CTAs per SM (0 = one CTA per tile)")
    ap.add_argument("--eig", type=int, default=None, help="1: eigen-solve / hash inside the filter kernel (s = 2, b24)")
    ap.add_argument("--duo", type=int, default=None, help="1: two pixel types per CTA for s = 2 with b24 records, 0: one type per CTA")
    ap.add_argument("--overlap", type=int, default=None, help="1: overlapped prep/filter pipeline, 0: serial")
    ap.add_argument("--taps", default="auto", choices=["auto", "fp32", "fp16", "b24"],
                    help="tap records in shared memory (auto = 24-bit when provably within 5e-5 of fp32 taps; fp16 is NOT a headline configuration)")
    ap.add_argument("--chunk-mb", type=int, default=None, help="scratch budget per kernel launch in MiB (library default 208)")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    SW, SH = cfg["sw"], cfg["sh"]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args, cfg, rank, world)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the RAISR path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from oclcomputervision_b200 import ClRaisr, synth, _cabi
    F = synth.random_filters(SCALE)
    r = ClRaisr(1, filters=F, device=local_rank, taps=args.taps)
    if args.filter_impl is not None:
        r.set_option("filter_impl", args.filter_impl)
    if args.overlap is not None:
        r.set_option("overlap", args.overlap)
    if args.duo is not None:
        r.set_option("filter_duo", args.duo)
    if args.eig is not None:
        r.set_option("eigen_in_filter", args.eig)
    if args.prep_ctas is not None:
        r.set_option("prep_ctas_per_sm", args.prep_ctas)
    if args.chunk_mb is not None:
        r.set_option("chunk_budget_bytes", args.chunk_mb << 20)
    info = r.device_info()
    n = args.frames or cfg["frames"]
    dw, dh = SW * SCALE, SH * SCALE
    host_src = make_inputs(cfg, n, rank)
    src = torch.from_numpy(host_src).cuda()
    dst = torch.empty((n, dh, dw), dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream()
    r.set_stream(stream.cuda_stream)
    px_per_step = n * dw * dh

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(timed):
        return r.upsample_device(src.data_ptr(), SW, SH, SW, dst.data_ptr(), dw, SCALE, n, np.uint8, timed=timed)

    def timed_steps(k):
        """k steps with per-kernel CUDA events inside; returns (elapsed ms, prep ms, filter ms) of this rank."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pm = fm = 0.0
        barrier()
        e0.record(stream)
        for _ in range(k):
            step(True)                       # per-kernel CUDA events on the launch stream, inside the region
            a, b = r.last_kernel_ms()
            pm += a; fm += b
        e1.record(stream)
        barrier()
        return e0.elapsed_time(e1), pm, fm

    for _ in range(max(args.warmup, 3)):
        step(False)
    barrier()
    launches0 = r.launch_count()
    sampler = ClockSampler(local_rank, uuid=getattr(torch.cuda.get_device_properties(local_rank), "uuid", None))
    elapsed_ms, prep_ms, filt_ms = timed_steps(args.steps)
    clocks = sampler.stop()
    launches = r.launch_count() - launches0
    t = torch.tensor([elapsed_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    value = world * px_per_step * args.steps / (elapsed_ms * 1e-3) / 1e6
    Feff, tap_format, b24_bound = r.effective_filters(SCALE)

    # ---- end to end through the C-ABI with pinned host buffers (H2D + kernels + D2H per step)
    lib = _cabi.load()
    hp_src, hp_dst = ctypes.c_void_p(), ctypes.c_void_p()
    _cabi.check(lib.raisr_host_alloc(ctypes.byref(hp_src), n * SW * SH))
    _cabi.check(lib.raisr_host_alloc(ctypes.byref(hp_dst), n * dw * dh))
    ctypes.memmove(hp_src.value, host_src.ctypes.data, n * SW * SH)
    r.set_stream(0)
    ms3 = (ctypes.c_float * 3)()

    def e2e_step():
        _cabi.check(lib.raisr_upsample_u8(r._h, hp_src.value, SW, SH, SW, hp_dst.value, dw, dh, dw, SCALE, n,
                                          _cabi.RAISR_HOST, ms3))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    ceil_s = copy_ceiling(torch, n * SW * SH, n * dw * dh, barrier)
    te = torch.tensor([e2e_s, ceil_s], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s, ceil_s = [float(x) for x in te.tolist()]
    e2e_value = world * px_per_step * e2e_steps / e2e_s / 1e6
    ceil_value = world * px_per_step / ceil_s / 1e6
    out_host = np.ctypeslib.as_array((ctypes.c_ubyte * (dw * dh)).from_address(hp_dst.value)).reshape(dh, dw)
    same_as_device = bool(np.array_equal(out_host, dst[0].cpu().numpy()))
    e2e_ms3 = [float(x) for x in ms3]
    # one frame through the same call, as the reference's own timing print does (raisr.py:135,182)
    one3 = (ctypes.c_float * 3)()
    for _ in range(3):
        _cabi.check(lib.raisr_upsample_u8(r._h, hp_src.value, SW, SH, SW, hp_dst.value, dw, dh, dw, SCALE, 1, _cabi.RAISR_HOST, one3))
    one_ms3 = [float(x) for x in one3]
    lib.raisr_host_free(hp_src); lib.raisr_host_free(hp_dst)

    # ---- the same step with the other tap records (device-resident), and what each does to the output
    variants = None
    if not args.no_extras:
        variants = {}
        probe = synth.synthetic_frame(1080, 1920, 1000)
        r.set_option("taps", 0)
        base_out = r.upsample_f32(probe, SCALE)
        r.set_stream(stream.cuda_stream)
        for name, mode in (("fp32", 0), ("b24", 2), ("fp16", 1)):
            r.set_option("taps", mode)
            for _ in range(2):
                step(False)
            k = max(1, min(args.steps, 5))
            el, pm, fm = timed_steps(k)
            tv = torch.tensor([el], device="cuda", dtype=torch.float64)
            if world > 1:
                dist.all_reduce(tv, op=dist.ReduceOp.MAX)
            el = float(tv.item())
            r.set_stream(0)
            dev = float(np.abs(r.upsample_f32(probe, SCALE) - base_out).max())
            r.set_stream(stream.cuda_stream)
            variants[name] = dict(value=round(world * px_per_step * k / (el * 1e-3) / 1e6, 1), unit="Mpix/s",
                                  prep_ms_per_step=round(pm / k, 3), filter_ms_per_step=round(fm / k, 3),
                                  max_abs_dev_vs_fp32_taps=dev,
                                  roofline_frac=round(FLOP_PER_PX * px_per_step * k / ((pm + fm) * 1e-3) / 1e12 / (2.0 * 128 * info["sm_count"] * info["sm_clock_khz"] * 1e3 / 1e12), 4))
        variants["note"] = ("fp32 = reference-precision taps; b24 = sign+exponent+15 mantissa bits (what 'auto' selects when the output provably stays "
                            "within 5e-5 of fp32 taps); fp16 = the reference's own (half)pf[...] rounding (raisr.cl:328), outside the 1e-4 bar against fp32 taps, "
                            "never the headline.  Deviation measured on one 1080p->4K frame, float output.")
        r.set_option("taps", {"fp32": 0, "fp16": 1, "b24": 2, "auto": 3}[args.taps])
        r.set_stream(0)

    band = None
    if not args.no_extras:
        try:
            band = band_record(torch, dist, r, rank, world, local_rank)
        except Exception as e:       # the headline must survive a failure of the side record, which then says so
            band = dict(error="%s: %s" % (type(e).__name__, e))

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        ffma_meas = r.measure_ffma_tflops()
        peak_nominal = 2.0 * 128 * info["sm_count"] * info["sm_clock_khz"] * 1e3 / 1e12
        kern_s = (prep_ms + filt_ms) * 1e-3
        achieved_tf = FLOP_PER_PX * px_per_step * args.steps / kern_s / 1e12
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        clk_hz = info["sm_clock_khz"] * 1e3
        filt_s = filt_ms * 1e-3
        px_total = px_per_step * args.steps
        chunk_mb = args.chunk_mb or 208
        per_frame_bytes = ((dh + 18 + 3) // 4 * 4) * (dw + 10) * 4
        chunk_frames = (chunk_mb << 20) // per_frame_bytes                      # frames per kernel launch (frames_per_launch() of raisr_api.cu)
        if chunk_frames < 3 and chunk_mb >= 200:
            chunk_frames = min(3, (2 << 30) // per_frame_bytes)
        chunk_frames = max(1, min(n, int(chunk_frames)))
        launches_per_step = 2 * ((n + chunk_frames - 1) // chunk_frames)
        # s = 2 with 24-bit records runs the two-types-per-CTA kernel unless --duo 0
        kname = "filter_duo_kernel" if (tap_format == "b24" and args.duo in (None, 1)) else "filter_octet_kernel"
        prof = ncu_metrics(kname, tap_format)
        dom = dict(name="%s<%s taps>" % (kname, tap_format), flop_per_px=FLOP_PER_PX_FILTER,
                   achieved_tflops=round(FLOP_PER_PX_FILTER * px_total / filt_s / 1e12, 3),
                   avg_launch_ms=round(filt_ms / (args.steps * launches_per_step / 2), 4),
                   binding_resource="shared-memory data pipe, 1 wavefront (128 B) per clock per SM: each pixel needs its own 121 taps")
        traffic = None
        if prof:
            frac = prof["smem_wavefronts_per_px"] * px_total / (filt_s * info["sm_count"] * clk_hz)
            dom.update(smem_wavefronts_per_px=prof["smem_wavefronts_per_px"], smem_pipe_frac=round(min(frac, 1.0), 4),
                       ncu_source=prof.get("source"))
            if frac > 1.0:
                dom["smem_pipe_frac_note"] = "profile figure and live time disagree (computed %.3f): re-capture the profile" % frac
            traffic = int(prof["dram_bytes_per_px"] * chunk_frames * dw * dh)
        roofline = dict(bound="fp32_ffma", achieved=round(achieved_tf, 3), peak=round(peak_nominal, 2), unit="TFLOP/s",
                        frac=round(achieved_tf / peak_nominal, 4), traffic=traffic,
                        traffic_note=("DRAM read+write bytes per launch of the dominant kernel from the ncu capture named in dominant_kernel.ncu_source; "
                                      if traffic else "no ncu capture of this kernel variant committed; ") + "algorithmic %.0f" % (BYTES_PER_PX * chunk_frames * dw * dh),
                        peak_source="2*128*SMs*clocks.max.sm (SURVEY 8(d)); measured register-only FFMA kernel: %.1f TFLOP/s" % ffma_meas,
                        frac_of_measured_ffma=round(achieved_tf / ffma_meas, 4),
                        flop_per_px=FLOP_PER_PX,
                        kernels=dict(prep_ms_per_step=round(prep_ms / args.steps, 3), filter_ms_per_step=round(filt_ms / args.steps, 3),
                                     filter_share=round(filt_ms / (prep_ms + filt_ms), 3), launches_per_step=launches_per_step,
                                     timing="CUDA events around every launch on the launch stream, inside the timed region"),
                        dominant_kernel=dom,
                        hbm=dict(algorithmic_bytes_per_px=BYTES_PER_PX,
                                 achieved_gbs=round(BYTES_PER_PX * px_per_step * args.steps / kern_s / 1e9, 1),
                                 peak_gbs=hbm_peak, peak_source="MEASURED_PEAKS.json" if peaks else "fallback 6650"))
        cb = None if (args.no_cpu_baseline or world > 1) else cpu_baseline(cfg)
        next_rows = None
        if not args.no_extras and world == 1:
            try:
                next_rows = next_rows_record(torch, lib, r._h, hbm_peak)
            except Exception as e:
                next_rows = dict(error="%s: %s" % (type(e).__name__, e))
        line = dict(metric="RAISR 2x output Mpix/s", value=round(value, 1), unit="Mpix/s", n_gpus=world, steps=args.steps,
                    warmup=max(args.warmup, 3), ms_per_step=round(elapsed_ms / args.steps, 3), higher_is_better=True,
                    scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                    config=dict(workload=workload_text(cfg, n, "" if args.taps != "fp16" else "; NON-DEFAULT: taps rounded to fp16 (raisr.cl:328), fp32 arithmetic"),
                                frames_per_gpu=n, src="%dx%d" % (SW, SH), dst="%dx%d" % (dw, dh), parallelism="frames sharded, no collective",
                                taps="%s (requested %s; b24 output bound %.2e vs fp32 taps; arithmetic fp32)" % (tap_format, args.taps, b24_bound),
                                l2="per-step working set (%.0f MB in + %.0f MB out) exceeds the 126 MB L2" % (n * SW * SH / 1e6, n * dw * dh / 1e6),
                                device=info["name"]),
                    clocks=clocks, gpu_launches=int(launches),
                    e2e=dict(value=round(e2e_value, 1), unit="Mpix/s", h2d_bytes_per_step=n * SW * SH, d2h_bytes_per_step=n * dw * dh,
                             h2d_kernel_d2h_ms=[round(x, 3) for x in e2e_ms3], one_frame_h2d_kernel_d2h_ms=[round(x, 3) for x in one_ms3],
                             matches_device_path=same_as_device,
                             copy_ceiling=dict(value=round(ceil_value, 1), unit="Mpix/s", gbs=round(world * (n * SW * SH + n * dw * dh) / ceil_s / 1e9, 1),
                                               note="the step's H2D and D2H bytes moved by bare cudaMemcpyAsync (pinned, two streams), all ranks at once, no kernels"),
                             frac_of_copy_ceiling=round(e2e_value / ceil_value, 3),
                             bound=dict(value=round(min(ceil_value, value), 1), unit="Mpix/s",
                                        by="kernels (device-resident value)" if value <= ceil_value else "host copies (copy_ceiling)"),
                             frac_of_bound=round(e2e_value / min(ceil_value, value), 3),
                             api="raisr_upsample_u8(where=RAISR_HOST), pinned buffers", host_affinity=numa),
                    roofline=roofline, cpu_baseline=cb, variants=variants, band=band, next_rows=next_rows)
        print(json.dumps(line), flush=True)
    r.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
