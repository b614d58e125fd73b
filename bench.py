#!/usr/bin/env python
"""bench.py -- RAISR 2x output Mpix/s on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch: BASELINE.json configs[1], 64 synthetic 1080p
luma frames -> 4K (u8 in, u8 out) per GPU.  With N GPUs every rank processes its own 64 frames
(frames are independent, no collective: weak scaling); `value` is the whole-job output Mpix/s.

  value      device-resident throughput (inputs already in HBM), CUDA events, max over ranks
  e2e        same metric through the C-ABI with HOST (pinned) buffers: H2D and D2H inside the call
  roofline   FP32-FFMA roofline of SURVEY.md 8(d): 412 algorithmic FLOP per output pixel over the
             summed kernel time of the step, against 2*128*SMs*max_clock and the measured FFMA peak
  cpu_baseline  the oracle's C port (oracle/raisr_oracle.c) on the host cores of this box -- the
             reference has no CPU RAISR path and its OpenCL path cannot run
`--impl reference` times that CPU port alone (there is nothing else of the reference to time).
"""
import argparse
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

SW, SH, SCALE, FRAMES = 1920, 1080, 2, 64       # BASELINE.json configs[1]
FLOP_PER_PX = 412.0                             # SURVEY.md 8(d), minimal form: whole path
FLOP_PER_PX_FILTER = 244.0                      # 121 FMA + store: the dominant kernel's share of the 412
BYTES_PER_PX = 1.0 / (SCALE * SCALE) + 1.0      # u8 in -> u8 out
# From the committed ncu capture of this command's kernels (profiles/r1i_ncu_summary.txt):
NCU_FILTER_WAVEFRONTS_PER_PX = 4.915            # l1tex__data_pipe_lsu_wavefronts_mem_shared.sum / output pixels
NCU_FILTER_DRAM_BYTES_PER_PX = 5.876            # dram__bytes_read.sum + dram__bytes_write.sum, per output pixel


def make_inputs(n_frames, rank):
    from oclcomputervision_b200 import synth
    pool = synth.synthetic_batch(8, SH, SW, pool=8, seed=1000 + 16 * rank)
    reps = (n_frames + 7) // 8
    return np.ascontiguousarray(np.tile(pool, (reps, 1, 1))[:n_frames])


def cpu_baseline(budget_s=12.0):
    """C oracle on all host threads, bounded sample of the bench workload."""
    from oracle import raisr_oracle as O
    from oclcomputervision_b200 import synth
    F = synth.random_filters(SCALE)
    frame = synth.synthetic_frame(SH, SW, 1000)
    threads = len(os.sched_getaffinity(0))   # torchrun exports OMP_NUM_THREADS=1; the oracle sets its own count
    t0 = time.perf_counter()
    O.raisr_ref_c(frame, F, SCALE, nthreads=threads, want=("out_u8",))
    one = time.perf_counter() - t0
    n = int(max(1, min(64, budget_s / max(one, 1e-3))))
    t0 = time.perf_counter()
    for k in range(n):
        O.raisr_ref_c(frame, F, SCALE, nthreads=threads, want=("out_u8",))
    dt = time.perf_counter() - t0
    mpix = n * SW * SH * SCALE * SCALE / dt / 1e6
    return dict(value=round(mpix, 3), unit="Mpix/s", cores=threads, kind="port",
                sample="%d frame(s) of the %dx%d->%dx%d workload, C oracle (fp32, OpenMP), %.1f s" %
                       (n, SW, SH, SW * SCALE, SH * SCALE, dt))


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


def run_reference(args, rank, world):
    """Reference arm: the reference has no CPU RAISR path and its OpenCL kernel cannot run here, so this times the
    oracle's C port of raisr.cl on all host threads.  One step = a bounded sample of the bench workload (as many
    1080p->4K frames as fit ~2 s); W warm-up steps, K timed steps."""
    if rank != 0:
        return
    from oracle import raisr_oracle as O
    from oclcomputervision_b200 import synth
    F = synth.random_filters(SCALE)
    frame = synth.synthetic_frame(SH, SW, 1000)
    threads = len(os.sched_getaffinity(0))
    t0 = time.perf_counter()
    O.raisr_ref_c(frame, F, SCALE, nthreads=threads, want=("out_u8",))
    one = time.perf_counter() - t0
    steps, warm = max(1, args.steps), max(0, args.warmup)
    step_budget = min(2.0, 150.0 / (steps + warm))
    n = int(max(1, min(64, step_budget / max(one, 1e-3))))

    def step():
        for _ in range(n):
            O.raisr_ref_c(frame, F, SCALE, nthreads=threads, want=("out_u8",))
    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    mpix = steps * n * SW * SH * SCALE * SCALE / dt / 1e6
    cb = dict(value=round(mpix, 3), unit="Mpix/s", cores=threads, kind="port",
              sample="%d step(s) of %d frame(s) of the %dx%d->%dx%d workload, C oracle (fp32, OpenMP), %.1f s" %
                     (steps, n, SW, SH, SW * SCALE, SH * SCALE, dt))
    line = dict(metric="RAISR 2x output Mpix/s", value=cb["value"], unit="Mpix/s", impl="reference",
                n_gpus=args.gpus, steps=steps, warmup=warm, ms_per_step=round(dt / steps * 1e3, 3),
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload="RAISR 2x 1080p->4K luma u8->u8, random-init 24x3x3x4x121 fp32 table (BASELINE configs[1]); "
                                     "bounded sample: %d frame(s) per step" % n,
                            note="the reference has no CPU RAISR path and its OpenCL kernel cannot run here; "
                                 "this is the oracle's C port of raisr.cl on the host cores"),
                cpu_baseline=cb, gpu_launches=0,
                e2e=dict(value=cb["value"], unit="Mpix/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=FRAMES, help="frames per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--filter-impl", type=int, default=None, help="1 octet (default), 0 block")
    ap.add_argument("--overlap", type=int, default=None, help="1: overlapped prep/filter pipeline, 0: serial")
    ap.add_argument("--taps", default="fp32", choices=["fp32", "fp16"],
                    help="fp16: taps rounded to half precision like the reference's (half)pf[...] (NOT the headline configuration)")
    ap.add_argument("--chunk-mb", type=int, default=208, help="scratch budget per kernel launch in MiB (library default 208)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

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
    import ctypes
    F = synth.random_filters(SCALE)
    r = ClRaisr(1, filters=F, device=local_rank, taps=args.taps)
    if args.filter_impl is not None:
        r.set_option("filter_impl", args.filter_impl)
    if args.overlap is not None:
        r.set_option("overlap", args.overlap)
    if args.chunk_mb is not None:
        r.set_option("chunk_budget_bytes", args.chunk_mb << 20)
    info = r.device_info()
    n = args.frames
    dw, dh = SW * SCALE, SH * SCALE
    host_src = make_inputs(n, rank)
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

    for _ in range(max(args.warmup, 3)):
        step(False)
    barrier()
    launches0 = r.launch_count()
    sampler = ClockSampler(local_rank, uuid=getattr(torch.cuda.get_device_properties(local_rank), "uuid", None))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    prep_ms = filt_ms = 0.0
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        step(True)                       # per-kernel CUDA events on the launch stream, inside the region
        a, b = r.last_kernel_ms()
        prep_ms += a; filt_ms += b
    ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    launches = r.launch_count() - launches0
    elapsed_ms = ev0.elapsed_time(ev1)
    t = torch.tensor([elapsed_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    value = world * px_per_step * args.steps / (elapsed_ms * 1e-3) / 1e6

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
    te = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * px_per_step * e2e_steps / float(te.item()) / 1e6
    out_host = np.ctypeslib.as_array((ctypes.c_ubyte * (dw * dh)).from_address(hp_dst.value)).reshape(dh, dw)
    same_as_device = bool(np.array_equal(out_host, dst[0].cpu().numpy()))
    e2e_ms3 = [float(x) for x in ms3]
    # one frame through the same call, as the reference's own timing print does (raisr.py:135,182)
    one3 = (ctypes.c_float * 3)()
    for _ in range(3):
        _cabi.check(lib.raisr_upsample_u8(r._h, hp_src.value, SW, SH, SW, hp_dst.value, dw, dh, dw, SCALE, 1, _cabi.RAISR_HOST, one3))
    one_ms3 = [float(x) for x in one3]
    lib.raisr_host_free(hp_src); lib.raisr_host_free(hp_dst)

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
        chunk_frames = max(1, min(n, int((args.chunk_mb << 20) // (((dh + 18 + 3) // 4 * 4) * (dw + 10) * 4))))   # frames per kernel launch (as in raisr_api.cu)
        roofline = dict(bound="fp32_ffma", achieved=round(achieved_tf, 3), peak=round(peak_nominal, 2), unit="TFLOP/s",
                        frac=round(achieved_tf / peak_nominal, 4),
                        traffic=int(NCU_FILTER_DRAM_BYTES_PER_PX * chunk_frames * dw * dh),
                        traffic_note="DRAM bytes per launch of the dominant kernel (ncu, profiles/); algorithmic %.0f" % (BYTES_PER_PX * chunk_frames * dw * dh),
                        peak_source="2*128*SMs*clocks.max.sm (SURVEY 8(d)); measured register-only FFMA kernel: %.1f TFLOP/s" % ffma_meas,
                        frac_of_measured_ffma=round(achieved_tf / ffma_meas, 4),
                        flop_per_px=FLOP_PER_PX,
                        kernels=dict(prep_ms_per_step=round(prep_ms / args.steps, 3), filter_ms_per_step=round(filt_ms / args.steps, 3),
                                     filter_share=round(filt_ms / (prep_ms + filt_ms), 3)),
                        dominant_kernel=dict(
                            name="filter_octet_kernel", flop_per_px=FLOP_PER_PX_FILTER,
                            achieved_tflops=round(FLOP_PER_PX_FILTER * px_total / filt_s / 1e12, 3),
                            binding_resource="shared-memory data pipe, 1 wavefront (128 B) per clock per SM: each pixel needs its own 484 B of fp32 taps",
                            wavefronts_per_px=NCU_FILTER_WAVEFRONTS_PER_PX,
                            smem_pipe_frac=round(NCU_FILTER_WAVEFRONTS_PER_PX * px_total / (filt_s * info["sm_count"] * clk_hz), 4)),
                        hbm=dict(algorithmic_bytes_per_px=BYTES_PER_PX,
                                 achieved_gbs=round(BYTES_PER_PX * px_per_step * args.steps / kern_s / 1e9, 1),
                                 peak_gbs=hbm_peak, peak_source="MEASURED_PEAKS.json" if peaks else "fallback 6650"))
        cb = None if (args.no_cpu_baseline or world > 1) else cpu_baseline()
        line = dict(metric="RAISR 2x output Mpix/s", value=round(value, 1), unit="Mpix/s", n_gpus=world, steps=args.steps,
                    warmup=max(args.warmup, 3), ms_per_step=round(elapsed_ms / args.steps, 3), higher_is_better=True,
                    scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                    config=dict(workload="RAISR 2x 1080p->4K luma u8->u8, batch of %d synthetic frames per GPU, "
                                         "random-init 24x3x3x4x121 fp32 table (BASELINE configs[1])%s" % (n, "" if args.taps == "fp32" else "; NON-DEFAULT: taps rounded to fp16 (raisr.cl:328), fp32 arithmetic"),
                                frames_per_gpu=n, src="%dx%d" % (SW, SH), dst="%dx%d" % (dw, dh), parallelism="frames sharded, no collective",
                                l2="per-step working set (%.0f MB in + %.0f MB out) exceeds the 126 MB L2" % (n * SW * SH / 1e6, n * dw * dh / 1e6),
                                device=info["name"]),
                    clocks=clocks, gpu_launches=int(launches),
                    e2e=dict(value=round(e2e_value, 1), unit="Mpix/s", h2d_bytes_per_step=n * SW * SH, d2h_bytes_per_step=n * dw * dh,
                             h2d_kernel_d2h_ms=[round(x, 3) for x in e2e_ms3], one_frame_h2d_kernel_d2h_ms=[round(x, 3) for x in one_ms3],
                             matches_device_path=same_as_device,
                             api="raisr_upsample_u8(where=RAISR_HOST), pinned buffers", host_affinity=numa),
                    roofline=roofline, cpu_baseline=cb)
        print(json.dumps(line), flush=True)
    r.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
