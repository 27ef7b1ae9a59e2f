#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 operator path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--extra]

Workload (BASELINE.json configs[1], "C2"): Sobel-X + Sobel-Y + Laplace 3x3 local operators on a float
8192 x 8192 image with MIRROR boundary handling.  One step = the three operators over the image (three
launches of the TMA-staged local-operator kernel).  value = operator-pixels per second: 3 * 8192 * 8192
pixels per step / step time, whole job.  N > 1 (torchrun, one rank per GPU): weak scaling -- every rank
owns an 8192 x 8192 row strip of an 8192 x (8192*N) image, ghost rows are exchanged with the
neighbouring ranks inside the timed step (NCCL send/recv), MIRROR is applied only at the global edges.

Timed on the device with CUDA events on the stream the kernels run on, after W >= 3 warm-up steps,
barrier + synchronize on both sides, max over ranks.  Inputs (256 MiB) and outputs (768 MiB) are
larger than the 126 MB L2, so no explicit flush is needed (stated in `config`).

Every run also reports, under "operators", the two sharded configs BASELINE.json names: C4 (fused Harris on a
32768 x 32768 uchar image) and C5 (8-level Gaussian / Laplacian pyramid of a 16384 x 16384 float image).  At N > 1 the
named image is cut into N row strips (STRONG scaling; peer-to-peer halo exchange / all-gather inside the timed step),
every rank also runs the unsharded operator on the whole image on its own GPU, checks its strip bit for bit against it
("sharded_parity") and times it ("strong_efficiency" = T(1 GPU) / (N * T(N GPUs)), both measured in this run).

--impl reference: the reference's CPU path for the same workload -- the -emit-cpu shaped loops in
oracle/emit_cpu_fast.cpp (constexpr masks, interior / border split, AVX2; bit-identical to the generic checker
oracle/emit_cpu.cpp; the Hipacc compiler itself cannot be built here -- DESIGN.md), all host threads, the FULL
8192 x 8192 image per step.  Rank 0 only: under torchrun the same host CPU is timed at every N.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

W = H = 8192
OPS = ("sobel_x", "sobel_y", "laplace")
ALG_BYTES_PER_PX = 8            # 4 B read + 4 B write per pixel per operator
METRIC = "Gpixels/s per operator (local operators Sobel-X + Sobel-Y + Laplace 3x3, float 8192x8192, MIRROR)"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons DURING the timed region (the profiling recipe's clocks line), sampled
    through NVML (the library behind nvidia-smi) every ~2 ms so that even a short timed region is covered."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_ev = index, [], threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.max_sm = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self._stop_ev.is_set():
            try:
                self.rows.append((nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM),
                                  nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)))
            except Exception:
                pass
            self._stop_ev.wait(0.002)

    def stop(self):
        self._stop_ev.set()
        self.join(timeout=6)
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": "nvml unavailable"}
        nv = self.nv
        sm = sorted(r[0] for r in self.rows)
        bits = 0
        for r in self.rows:
            bits |= r[1]
        names = [("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap)]
        reasons = [n for n, b in names if bits & b]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_sm, "reasons": reasons, "samples": len(self.rows),
                "source": "nvml"}


def specs_for_workload():
    import numpy as np
    from hipacc_b200 import _abi as A, masks as M, specs as S
    return [S.domain_reduce_f32(m.astype(np.float32), A.MIRROR) for m in (M.SOBEL3_X, M.SOBEL3_Y, M.LAPLACE3)]


# --------------------------------------------------------------------------------------- CPU legs
def _cpu_step_fn():
    """One CPU step = the three operators over the FULL 8192 x 8192 image (the GPU arm's per-GPU workload), through the
    specialised -emit-cpu shaped loops (oracle/emit_cpu_fast.cpp, pinned bit-for-bit against the generic oracle)."""
    import numpy as np
    from hipacc_b200 import synth
    from oracle import oracle as O
    O.set_num_threads(len(os.sched_getaffinity(0)))   # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every host core
    img = synth.image_np("float32", W, H, seed=2)
    specs = specs_for_workload()
    outs = [np.empty_like(img) for _ in specs]

    def step():
        for s, o in zip(specs, outs):
            O.local_op_fast(s, img, out=o)
    return step, O.num_threads(), len(specs)


CPU_SAMPLE = (f"3 operators on the full {W}x{H} float image per step (same config as the GPU arm at N = 1), -emit-cpu shaped loops: "
              "constexpr masks, interior / border split, g++ -O3 AVX2 without FMA contraction, OpenMP over rows")


def cpu_baseline(repeats=3):
    """-> (Gpx/s, cores, sample description); best of `repeats` after one warm-up step"""
    step, cores, nops = _cpu_step_fn()
    step()   # warm-up (page faults, OpenMP pool)
    best = float("inf")
    for _ in range(repeats):
        t = time.perf_counter()
        step()
        best = min(best, time.perf_counter() - t)
    return nops * W * H / best / 1e9, cores, CPU_SAMPLE + f", best of {repeats}"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = max(1, args.steps), max(1, args.warmup)
    step, cores, nops = _cpu_step_fn()
    for _ in range(warm):
        step()
    t = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t) / steps
    val = nops * W * H / dt / 1e9
    sample = CPU_SAMPLE + f", OpenMP {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "Gpixels/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C2: Sobel-X + Sobel-Y + Laplace 3x3 local operators, float 8192x8192, MIRROR boundary (CPU arm)", "sample": sample,
                   "note": "the CPU arm does not scale with --gpus: rank 0 times the same host cores on one 8192x8192 image at every N"},
        "cpu_baseline": {"value": val, "unit": "Gpixels/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "Gpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def measured_traffic(kernel_substr, grid_x=None):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernel whose name contains `kernel_substr`, from the
    newest committed ncu launch list (profiles/*launches*.csv, `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,
    dram__bytes_write.sum`); None when no list carries the metric.  grid_x: only launches with that many CTAs (the list
    also holds the same kernel on smaller images).  -> (bytes per launch, source file)"""
    import csv
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*launches*.csv")), key=lambda f: (os.path.basename(f).split("_")[0], os.path.getmtime(f)))
    for path in reversed(files):
        try:
            per_id = {}
            with open(path, newline="") as fh:
                rows = [r for r in csv.reader(l for l in fh if l.startswith('"'))]
            hdr = rows[0]
            iK, iM, iV, iI = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
            iG = hdr.index("Grid Size") if "Grid Size" in hdr else None
            want_grid = None if grid_x is None or iG is None else f"({int(grid_x)}, 1, 1)"
            for r in rows[1:]:
                if want_grid is not None and r[iG] != want_grid:
                    continue
                if kernel_substr in r[iK] and r[iM] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    per_id[r[iI]] = per_id.get(r[iI], 0.0) + float(r[iV].replace(",", ""))
            if per_id:
                return sum(per_id.values()) / len(per_id), os.path.relpath(path, ROOT)
        except Exception:  # noqa: BLE001
            continue
    return None, None


# --------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--extra", action="store_true", help="also time the other BASELINE configs (C1, C3, C4 strip, C5) into 'operators'")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end leg (tuning sweeps only)")
    ap.add_argument("--halo", default="p2p", choices=["p2p", "nccl"], help="N > 1: halo exchange by peer-to-peer push kernel (default) or NCCL send/recv")
    ap.add_argument("--e2e-blocking", action="store_true", help="time the end-to-end leg with the blocking hb_image_write / hb_image_read calls (the reference's API shape) instead of the pipelined async region copies")
    ap.add_argument("--no-overlap", action="store_true", help="N > 1: run the halo kernel in stream order instead of overlapping it with the first operator's interior rows")
    ap.add_argument("--no-graph", action="store_true", help="launch every operator from the host instead of replaying a CUDA graph of one step")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the sharded-vs-unsharded bit comparison (tuning sweeps only)")
    ap.add_argument("--no-named", action="store_true", help="skip the C4 / C5 entries of 'operators' (tuning sweeps only)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import hipacc_b200 as hb
    from hipacc_b200 import _abi as A, strips, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    steps, warm = max(1, args.steps), max(3, args.warmup)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    hb.init(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- data: this rank's strip of the global 8192 x (8192*world) image, resident in HBM
    plan = strips.StripPlan(W, H * world, world, rank, radius=1, boundary=A.MIRROR)
    plan.validate()
    pad = int(os.environ.get("HB_BENCH_ROW_PAD", "0"))   # experiment: extra floats per row (row pitch not a power of two)
    stride = (W + 63) // 64 * 64 + pad
    buf = hb.alloc_image(A.F32, stride, plan.buffer_rows, device=dev)   # a whole CUDA allocation: exportable through CUDA IPC
    strips.owned(buf, plan)[:, :W] = synth.image_torch("float32", W, plan.rows, seed=2, y0=plan.y0, device=dev)
    src = buf[:, :W]
    outs = [hb.empty_image(A.F32, W + pad, plan.buffer_rows, device=dev)[:, :W] for _ in OPS]
    specs = specs_for_workload()
    roi, ghost = plan.roi(), plan.ghost()
    stream = torch.cuda.Stream(device=dev)       # the stream every kernel of the timed region runs on
    torch.cuda.set_stream(stream)

    halo = None
    if world > 1 and args.halo == "p2p":
        try:
            halo = strips.P2PHalo(hb, buf, plan)   # raises on every rank together if CUDA IPC / peer access is unavailable
        except RuntimeError as e:
            sys.stderr.write(f"[rank {rank}] peer-to-peer halo exchange unavailable ({e}); using NCCL send/recv\n")

    skip_exchange = bool(int(os.environ.get("HB_BENCH_NO_EXCHANGE", "0")))   # diagnosis only: kernels without the halo exchange

    # N > 1 with the peer-to-peer exchange: the halo kernel runs on a side stream while the first operator works on the
    # rows that do not touch the ghost rows; its two 32-row edge strips and the other operators follow the join.
    overlap = halo is not None and not args.no_overlap and plan.rows > 4 * 32
    side = torch.cuda.Stream(device=dev) if overlap else None
    ev_fork, ev_join = torch.cuda.Event(), torch.cuda.Event()
    E = 32
    gt_, gb_ = plan.ghost_top, plan.ghost_bottom
    roi_mid, ghost_mid = (W, plan.rows - 2 * E, 0, gt_ + E), (gt_ + E, gb_ + E)
    roi_top, ghost_top_ = (W, E, 0, gt_), (gt_, plan.rows - E + gb_)
    roi_bot, ghost_bot = (W, E, 0, gt_ + plan.rows - E), (gt_ + plan.rows - E, gb_)

    def step_direct():
        if overlap and not skip_exchange:
            ev_fork.record(stream)
            side.wait_event(ev_fork)
            halo.exchange(side)
            ev_join.record(side)
            hb.local_op(specs[0], src, dst=outs[0], roi_in=roi_mid, roi_out=roi_mid, ghost=ghost_mid, stream=stream)
            stream.wait_event(ev_join)
            hb.local_op(specs[0], src, dst=outs[0], roi_in=roi_top, roi_out=roi_top, ghost=ghost_top_, stream=stream)
            hb.local_op(specs[0], src, dst=outs[0], roi_in=roi_bot, roi_out=roi_bot, ghost=ghost_bot, stream=stream)
            for s, o in zip(specs[1:], outs[1:]):
                hb.local_op(s, src, dst=o, roi_in=roi, roi_out=roi, ghost=ghost, stream=stream)
            return
        if skip_exchange:
            pass
        elif halo is not None:
            halo.exchange(stream)                   # one kernel: push edge rows into the neighbours' ghost rows over NVLink
        else:
            strips.exchange_halos(buf, plan)        # NCCL send/recv (no-op at N = 1)
        for s, o in zip(specs, outs):
            hb.local_op(s, src, dst=o, roi_in=roi, roi_out=roi, ghost=ghost, stream=stream)

    # One step = (halo exchange +) three operator launches.  The step is captured once into a CUDA graph and
    # replayed (the reference's own -use-graph mode, runtime/hipacc_cu_standalone.hpp:331-356), so the timed region
    # is not bounded by the Python host; at N > 1 the NCCL send/recv pair of the halo exchange is part of the graph.
    use_graph = not args.no_graph and not (world > 1 and halo is None)   # NCCL send/recv stays outside graphs
    launches_per_step = len(OPS) + (1 if halo is not None else 0) + (2 if overlap else 0)
    step, graph_note = step_direct, "direct launches through hb_local_op"
    if use_graph:
        step_direct()            # also creates the NCCL P2P communicators before capture
        torch.cuda.synchronize()
        ok = 1
        try:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=stream, capture_error_mode="thread_local"):
                step_direct()
        except Exception as e:   # noqa: BLE001 -- fall back to direct launches, on every rank
            ok = 0
            sys.stderr.write(f"[rank {rank}] CUDA graph capture failed ({type(e).__name__}: {e}); using direct launches\n")
        if world > 1:
            t_ok = torch.tensor([ok], dtype=torch.int32, device=dev)
            dist.all_reduce(t_ok, op=dist.ReduceOp.MIN)
            ok = int(t_ok.item())
        use_graph = bool(ok)
        if use_graph:
            step = graph.replay
            graph_note = "CUDA graph replay of the step (3 operator kernels" + ((" + 1 peer-to-peer halo kernel" + (" overlapped with the first operator's interior rows, 2 edge-strip launches)" if overlap else ")")) if world > 1 else ")")

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warm):
        step()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()   # before the barrier: its start-up cost must not skew rank 0 against the ranks that wait for its halo rows
    sync_all()
    n0 = hb.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    sync_all()
    ms = e0.elapsed_time(e1)
    if os.environ.get("HB_BENCH_VERBOSE"):
        sys.stderr.write(f"[rank {rank}] {ms / steps:.4f} ms per step on this rank\n")
    launches = launches_per_step * steps if use_graph else hb.launch_count() - n0
    clocks = sampler.stop() if sampler else None
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms = float(t_ms.item())
    ms_per_step = ms / steps
    px_per_step = len(OPS) * W * plan.rows * world      # whole job
    value = px_per_step / (ms_per_step * 1e-3) / 1e9

    # ---- roofline of the dominant kernel (local_tma_f32_kernel<3,3,...>, hb_local_tma.cu): per-launch average
    per_launch_ms = ms / (steps * len(OPS))
    peak, peak_src = peaks()
    achieved = ALG_BYTES_PER_PX * W * plan.rows / (per_launch_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "kernel": "local_tma_f32_kernel<3,3,Mask*> (TMA-staged, one 128x32 tile per CTA)",
                "note": "frac can slightly exceed 1: consecutive operators re-read the same 256 MiB input and a part of it still sits in the 126 MB L2; isolated single-operator launches (--extra: operators.C2_*) measure 0.98-1.02", "peak_source": peak_src,
                "algorithmic_bytes_per_launch": ALG_BYTES_PER_PX * W * plan.rows, "avg_launch_ms": per_launch_ms}
    roofline["traffic"], roofline["traffic_source"] = measured_traffic("local_tma_f32_kernel<3, 3", grid_x=((W + 127) // 128) * ((plan.rows + 31) // 32))

    # ---- e2e: the same step through the C-ABI memory calls with HOST (pinned) buffers, copies inside the timed region
    L = hb.lib()
    import ctypes as C
    h_in = torch.empty((plan.rows, W), dtype=torch.float32, pin_memory=True)
    h_in.copy_(strips.owned(buf, plan)[:, :W])
    h_out = [torch.empty((plan.rows, W), dtype=torch.float32, pin_memory=True) for _ in OPS]
    own_in = hb.view(strips.owned(buf, plan)[:, :W])
    own_out = [hb.view(o[plan.ghost_top:plan.ghost_top + plan.rows]) for o in outs]
    sp = hb.stream_ptr(stream)

    def e2e_step():
        L.hb_image_write(C.byref(own_in), C.c_void_p(h_in.data_ptr()), sp)       # host -> HBM (blocking, like hipaccWriteMemory)
        step_direct()
        for v, h in zip(own_out, h_out):
            L.hb_image_read(C.byref(v), C.c_void_p(h.data_ptr()), sp)            # HBM -> host (blocking, like hipaccReadMemory)
    # Pipelined form of the same step (default): the image is cut into K row strips; strip k's host->device copy, its three
    # operators and its device->host copies run on three streams, so PCIe transfers in both directions overlap each other
    # and the kernels (hb_image_write_region_async / hb_image_read_region_async; the blocking calls above are the
    # reference-shaped API, timed with --e2e-blocking).
    K = 8
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    bounds = [(plan.rows * k // K, plan.rows * (k + 1) // K) for k in range(K)]
    gt = plan.ghost_top
    ev_in = [torch.cuda.Event() for _ in range(K)]
    ev_k = [torch.cuda.Event() for _ in range(K)]
    ev_start, ev_done = torch.cuda.Event(), torch.cuda.Event()
    row_b = 4 * W

    def strip_view(t, y0, y1):
        return hb.view(t, roi=(W, y1 - y0, 0, gt + y0))

    def e2e_step_pipelined():
        ev_start.record(stream)
        s_in.wait_event(ev_start)
        s_out.wait_event(ev_start)
        for k in ([0, K - 1] + list(range(1, K - 1))):   # the first and the last strip first: they hold the rows the neighbours need
            y0, y1 = bounds[k]
            L.hb_image_write_region_async(C.byref(strip_view(src, y0, y1)), C.c_void_p(h_in.data_ptr() + y0 * row_b), row_b, hb.stream_ptr(s_in))
            ev_in[k].record(s_in)
        if world > 1:
            stream.wait_event(ev_in[0])
            stream.wait_event(ev_in[K - 1])
            if halo is not None:
                halo.exchange(stream)
            else:
                strips.exchange_halos(buf, plan)
        for k, (y0, y1) in enumerate(bounds):
            stream.wait_event(ev_in[min(k + 1, K - 1)])          # the strip below holds this strip's bottom halo row
            roi_k = (W, y1 - y0, 0, gt + y0)
            ghost_k = (gt + y0, plan.buffer_rows - (gt + y1))     # every other row of the buffer is real neighbour data
            for s_, o in zip(specs, outs):
                hb.local_op(s_, src, dst=o, roi_in=roi_k, roi_out=roi_k, ghost=ghost_k, stream=stream)
            ev_k[k].record(stream)
            s_out.wait_event(ev_k[k])
            for o, h in zip(outs, h_out):
                L.hb_image_read_region_async(C.byref(strip_view(o, y0, y1)), C.c_void_p(h.data_ptr() + y0 * row_b), row_b, hb.stream_ptr(s_out))
        ev_done.record(s_out)
        stream.wait_event(ev_done)

    if not args.e2e_blocking:
        e2e_step = e2e_step_pipelined
    e2e_steps = 0 if args.no_e2e else max(2, min(steps, 5))
    e2e_step()
    sync_all()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record(stream)
    for _ in range(e2e_steps):
        e2e_step()
    e3.record(stream)
    sync_all()
    if e2e_steps:   # the host buffers must hold exactly what the resident-data step computed (outside the timed region)
        step_direct()
        torch.cuda.synchronize()
        for o, h in zip(outs, h_out):
            assert torch.equal(o[plan.ghost_top:plan.ghost_top + plan.rows].cpu(), h), "e2e leg: host result differs from the device-resident step"
    e2e_s = e2.elapsed_time(e3) * 1e-3 / max(e2e_steps, 1)
    t_e = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e = {"value": px_per_step / float(t_e.item()) / 1e9, "unit": "Gpixels/s", "h2d_bytes_per_step": 4 * W * plan.rows * world,
           "d2h_bytes_per_step": 4 * W * plan.rows * len(OPS) * world, "ms_per_step": float(t_e.item()) * 1e3,
           "api": ("hb_image_write + 3 x hb_local_op + 3 x hb_image_read (blocking calls, pinned host buffers)" if args.e2e_blocking else
                   "8 row strips: hb_image_write_region_async -> 3 x hb_local_op -> 3 x hb_image_read_region_async on three streams (pinned host buffers; every byte crosses PCIe inside the timed region)")}

    # ---- sharded parity of the headline step (outside the timed region): every rank recomputes the three operators
    # UNSHARDED on the whole global image on its own GPU and compares its strip bit for bit
    parity = {}
    if world > 1 and not args.no_parity:
        step_direct()
        torch.cuda.synchronize()
        whole = hb.empty_image(A.F32, W, H * world, device=dev)
        for r in range(world):   # counter-based generator: any rank can produce any rows of the global image
            whole[r * H:(r + 1) * H].copy_(synth.image_torch("float32", W, H, seed=2, y0=r * H, device=dev))
        ref_out = hb.empty_image(A.F32, W, H * world, device=dev)
        ok = True
        for s_, o in zip(specs, outs):
            hb.local_op(s_, whole, dst=ref_out, stream=stream)
            torch.cuda.synchronize()
            ok = ok and torch.equal(o[plan.ghost_top:plan.ghost_top + plan.rows], ref_out[plan.y0:plan.y1])
        parity["C2"] = ok
        del whole, ref_out

    operators = named_operators(hb, dev, world, rank, stream, halo is not None, peak, parity, args)
    if args.extra:
        operators.update(extra_operators(hb, dev, peak, use_graph=not args.no_graph) if world == 1 else extra_sharded(hb, dev, world, rank, stream, halo is not None))

    sharded_parity = None
    if world > 1 and parity:
        flags = torch.tensor([1 if all(parity.values()) else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        mine = ", ".join(f"{k}:{'ok' if v else 'MISMATCH'}" for k, v in parity.items())
        if not all(parity.values()):
            sys.stderr.write(f"[rank {rank}] sharded parity FAILED: {mine}\n")
        sharded_parity = "ok" if int(flags.item()) else "MISMATCH"

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, cores, sample = cpu_baseline()
        cpu = {"value": v, "unit": "Gpixels/s", "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "Gpixels/s", "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "C2: Sobel-X + Sobel-Y + Laplace 3x3 local operators, float 8192x8192 per GPU, MIRROR boundary",
                       "pixels_per_step": px_per_step, "operators_per_step": len(OPS), "image": f"{W}x{plan.rows} per rank, {W}x{H * world} global",
                       "l2": "inputs 256 MiB + outputs 768 MiB per step exceed the 126 MB L2; no explicit flush",
                       "halo_exchange": "none (N=1)" if world == 1 else ("1 ghost row per side per step, pushed peer to peer over NVLink by hb_halo_exchange (CUDA IPC, device-side flags) inside the timed region"
                                         if halo is not None else "1 ghost row per side per step via NCCL send/recv inside the timed region"),
                       "launch": graph_note},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        }
        if operators:
            line["operators"] = operators
        if world > 1:
            line["sharded_parity"] = sharded_parity if sharded_parity else "not run"
            line["sharded_parity_detail"] = ("every rank ran the unsharded operator on the whole global image on its own GPU and compared its strip bit for bit: "
                                             + ", ".join(sorted(parity)) if parity else "skipped (--no-parity)")
        print(json.dumps(line))
    if world > 1:
        if halo is not None:   # every rank: a timed-out exchange poisons its control block
            n_ex, timed_out = halo.status()
            if timed_out or n_ex < 0:
                sys.stderr.write(f"[rank {rank}] ERROR: a halo exchange timed out waiting for a neighbour; results of this run are invalid\n")
        sys.stdout.flush()
        torch.cuda.synchronize()
        dist.barrier()
        os._exit(0)   # skip interpreter teardown: CUDA graphs / IPC mappings and the NCCL communicator do not need an orderly exit


def named_operators(hb, dev, world, rank, stream, p2p, peak, parity, args):
    """The two sharded configs BASELINE.json names, at every N: C4 = fused Harris on a 32768 x 32768 uchar image, C5 =
    8-level Gaussian / Laplacian pyramid of a 16384 x 16384 float image.  N = 1: the whole image on the GPU.  N > 1:
    STRONG scaling -- the same image cut into N row strips (halo exchange / all-gather peer to peer inside the timed
    step); every rank additionally runs the unsharded operator on the whole image on its own GPU, compares its strip
    bit for bit (parity[...]) and times it, so strong_efficiency = T(1) / (N * T(N)) comes from one run."""
    import torch
    import torch.distributed as dist
    from hipacc_b200 import _abi as A, masks as M, strips, synth
    res = {}
    if args.no_named:
        return res

    def timeit(fn, reps, warm):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def graphed(fn):
        """capture fn into a CUDA graph (every rank falls back together) -> (callable, how)"""
        fn()
        torch.cuda.synchronize()
        ok = 1
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream, capture_error_mode="thread_local"):
                fn()
        except Exception as e:  # noqa: BLE001
            ok = 0
            sys.stderr.write(f"[rank {rank}] graph capture failed: {e}\n")
        if world > 1:
            t_ok = torch.tensor([ok], dtype=torch.int32, device=dev)
            dist.all_reduce(t_ok, op=dist.ReduceOp.MIN)
            ok = int(t_ok.item())
        return (g.replay, "CUDA graph replay") if ok else (fn, "direct launches")

    # ------------------------------------------------------------------ C4: Harris 32768 x 32768 uchar
    Wc = Hc = 32768
    whole = hb.empty_image(A.U8, Wc, Hc, device=dev)
    for y in range(0, Hc, 4096):
        whole[y:y + 4096].copy_(synth.image_torch("uint8", Wc, 4096, seed=4, y0=y, device=dev))
    whole_out = hb.empty_image(A.U8, Wc, Hc, device=dev)
    t1 = timeit(lambda: hb.harris(whole, dst=whole_out, stream=stream), reps=5, warm=2)
    entry = {"Gpx_s": Wc * Hc / (t1 * 1e-3) / 1e9, "ms": t1, "n_gpus": world, "alg_GB_s": 2 * Wc * Hc / (t1 * 1e-3) / 1e9,
             "hbm_frac": 2 * Wc * Hc / (t1 * 1e-3) / 1e9 / peak, "note": "fused 9-kernel pipeline, one launch; integer-issue bound"}
    if world > 1:
        plan = strips.StripPlan(Wc, Hc, world, rank, radius=2, boundary=A.CLAMP)
        buf = hb.alloc_image(A.U8, Wc, plan.buffer_rows, device=dev)
        strips.owned(buf, plan).copy_(whole[plan.y0:plan.y1])
        out = torch.zeros_like(buf)
        halo = strips.P2PHalo(hb, buf, plan) if p2p else None

        # The exchange runs in stream order ahead of the fused kernel.  Overlapping it with the interior rows (halo kernel on a
        # side stream, two 32-row edge bands after the join -- what the C2 step does) was measured SLOWER here: 305 vs 295 us per
        # step at N = 8 (profiles/r4n_bench_n8_c4overlap.json): two extra launches of 256 CTAs cost more than the ~10 us exchange.
        def harris_step():
            if halo is not None:
                halo.exchange(stream)
            else:
                strips.exchange_halos(buf, plan)
            hb.harris(buf, dst=out, roi=plan.roi(), ghost=plan.ghost(), stream=stream)
        harris_step()
        torch.cuda.synchronize()
        if not args.no_parity:
            parity["C4"] = bool(torch.equal(strips.owned(out, plan), whole_out[plan.y0:plan.y1]))
        fn, how = graphed(harris_step) if p2p else (harris_step, "direct launches")
        tn = max_over_ranks(timeit(fn, reps=10, warm=3))
        entry = {"Gpx_s": Wc * Hc / (tn * 1e-3) / 1e9, "ms": tn, "n_gpus": world, "launch": how, "single_gpu_ms": t1,
                 "strong_efficiency": t1 / (world * tn),
                 "note": f"strong scaling: {plan.rows} rows per rank + 2 ghost rows pushed per step (" + ("peer-to-peer kernel" if p2p else "NCCL send/recv")
                         + "); single_gpu_ms = the unsharded 32768^2 image on this rank's GPU in the same run"}
        if halo is not None:
            halo.check()
        del buf, out
    # the two side legs of C4 must never cost the line its headline (2 GiB of pinned host memory, a host with few cores ...)
    if world == 1 and not args.no_e2e:
        try:
            entry["e2e"] = c4_e2e(hb, dev, stream, whole, whole_out, Wc, Hc)
        except Exception as e:  # noqa: BLE001
            entry["e2e"] = {"error": f"{type(e).__name__}: {e}"[:300]}
            torch.cuda.synchronize()
    if world == 1 and rank == 0 and not args.no_cpu:
        try:
            entry["cpu"] = c4_cpu()
        except Exception as e:  # noqa: BLE001
            entry["cpu"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    res["C4_harris_32768x32768"] = entry
    del whole, whole_out

    # ------------------------------------------------------------------ C5: pyramid, 8 levels, 16384 x 16384 float
    Wp = Hp = 16384
    depth = 8
    n = Wp * Hp
    img = hb.empty_image(A.F32, Wp, Hp, device=dev)
    for y in range(0, Hp, 2048):
        img[y:y + 2048].copy_(synth.image_torch("float32", Wp, 2048, seed=5, y0=y, device=dev))
    pg = hb.Pyramid(img, depth)
    pl = hb.Pyramid(hb.empty_image(A.F32, Wp, Hp, device=dev).zero_(), depth)
    sp = None
    if world > 1:   # sharded traversal first (its parity reference is the FIRST unsharded traversal of the same input)
        # HB_C5_SPLIT_DOG=1: the DifferenceOfGaussian of the sharded levels runs on a second stream, off the latency-bound chain
        plan = strips.PyramidShardPlan(Wp, Hp, depth, world, rank, 5, split_dog=bool(int(os.environ.get("HB_C5_SPLIT_DOG", "0"))))
        sp = strips.ShardedPyramid(plan, dev, hb=hb if p2p else None)
        if p2p:
            sp.enable_p2p(hb)
        sp.owned(sp.gaus, 0).copy_(img[plan.y0(0):plan.y1(0)])

        def traverse():
            if p2p:
                sp.traverse(hb, M.GAUSS5, stream=stream)
            else:   # NCCL fallback: send/recv for the level-0 halo, all-gather for level G
                strips.exchange_halos(sp.gaus[0], sp.strip0_plan())
                sp.down_sharded(hb, M.GAUSS5, stream)
                G = plan.G
                dist.all_gather_into_tensor(sp.gaus[G], sp.gaus[G][plan.y0(G):plan.y1(G)].contiguous())
                sp.coarse_and_up(hb, M.GAUSS5, stream)
        traverse()
        torch.cuda.synchronize()
    hb.pyramid_traverse(pg, pl, M.GAUSS5, stream=stream)
    torch.cuda.synchronize()
    if sp is not None and not args.no_parity:
        ok = True
        for l in range(depth):
            y0, y1 = plan.y0(l), plan.y1(l)
            ok = ok and torch.equal(sp.owned(sp.gaus, l), pg.levels[l][y0:y1]) and torch.equal(sp.owned(sp.lap, l), pl.levels[l][y0:y1])
        parity["C5"] = bool(ok)
    with hb.Graph(stream) as g:                      # hb_graph_begin / hb_graph_end: the 14 level kernels as one launch
        hb.pyramid_traverse(pg, pl, M.GAUSS5, stream=stream)
    t1 = timeit(lambda: g.launch(), reps=5, warm=2)
    g.destroy()
    alg = int(23 * n * 4 / 3)
    entry = {"Gpx_s": n / (t1 * 1e-3) / 1e9, "ms": t1, "n_gpus": world, "alg_GB_s": alg / (t1 * 1e-3) / 1e9, "hbm_frac": alg / (t1 * 1e-3) / 1e9 / peak,
             "note": "fused down (blur+subsample+DoG) 9n + fused up (Restore+Blend) 14n bytes per transition; 14 kernels replayed as one hb_graph launch"}
    if sp is not None:
        fn, how = graphed(traverse) if p2p else (traverse, "direct launches")
        tn = max_over_ranks(timeit(fn, reps=5, warm=2))
        entry = {"Gpx_s": n / (tn * 1e-3) / 1e9, "ms": tn, "n_gpus": world, "launch": how, "single_gpu_ms": t1, "strong_efficiency": t1 / (world * tn),
                 "note": f"strong scaling: ONE level-0 halo exchange ({plan.E0} rows per neighbour) + ONE all-gather of level {plan.G} "
                         f"({Wp >> plan.G}x{Hp >> plan.G}) per traversal, levels >= {plan.G} replicated, extension rows recomputed instead of exchanged; "
                         "single_gpu_ms = the unsharded traversal on this rank's GPU in the same run"}
        if os.environ.get("HB_BENCH_PHASES") and p2p:   # diagnosis: GPU-timer marks between the kernels of the captured traversal
            for mode in (False, True):
                marks = torch.zeros(16, dtype=torch.int64, device=dev)
                gfn, _ = graphed(lambda: sp.traverse(hb, M.GAUSS5, stream=stream, marks=marks, overlap=mode))
                acc = None
                for _ in range(8):
                    dist.barrier()
                    gfn()
                    torch.cuda.synchronize()
                    m = marks.cpu().numpy()[:9].astype("float64")
                    d_ = (m - m[0]) * 1e-3
                    acc = d_ if acc is None else acc + d_
                acc /= 8
                sys.stderr.write(f"[rank {rank}] C5 timeline overlap={mode} (us after start): " + ", ".join(f"{n_}: {a_:.1f}" for n_, a_ in zip(strips.ShardedPyramid.MARK_NAMES[1:], acc[1:])) + "\n")
        if sp.halo0 is not None:
            sp.halo0.check()
        if sp.gatherG is not None:
            sp.gatherG.check()
    res["C5_pyramid8_16384"] = entry
    return res


def c4_e2e(hb, dev, stream, whole, whole_out, Wc, Hc, K=16, reps=3):
    """C4 end to end with HOST buffers: the 1 GiB uchar image goes host -> HBM in K row strips, the fused Harris kernel runs
    per strip as soon as the strip and the first rows of the next one have landed, the 1 GiB result goes HBM -> host --
    three streams, both PCIe directions and the kernels overlap (hb_image_write_region_async / hb_harris /
    hb_image_read_region_async).  One byte in, one byte out per pixel and nine kernels' worth of work in between: the
    pipeline shape a host round trip suits, where C2 (1 plane in, 3 planes out, 0.25 ms of kernels) is the worst case."""
    import ctypes as C
    import torch
    L = hb.lib()
    h_in = torch.empty((Hc, Wc), dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty((Hc, Wc), dtype=torch.uint8, pin_memory=True)
    for y in range(0, Hc, 4096):
        h_in[y:y + 4096].copy_(whole[y:y + 4096])
    want = whole_out.clone()          # the device-resident run's result
    whole.zero_(); whole_out.zero_()
    torch.cuda.synchronize()
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    bounds = [(Hc * k // K, Hc * (k + 1) // K) for k in range(K)]
    ev_in = [torch.cuda.Event() for _ in range(K)]
    ev_k = [torch.cuda.Event() for _ in range(K)]
    ev_start, ev_done = torch.cuda.Event(), torch.cuda.Event()

    def step():
        ev_start.record(stream)
        s_in.wait_event(ev_start)
        s_out.wait_event(ev_start)
        for k, (y0, y1) in enumerate(bounds):
            L.hb_image_write_region_async(C.byref(hb.view(whole, roi=(Wc, y1 - y0, 0, y0))), C.c_void_p(h_in.data_ptr() + y0 * Wc), Wc, hb.stream_ptr(s_in))
            ev_in[k].record(s_in)
        for k, (y0, y1) in enumerate(bounds):
            stream.wait_event(ev_in[min(k + 1, K - 1)])        # the strip below holds this strip's two bottom halo rows
            hb.harris(whole, dst=whole_out, roi=(Wc, y1 - y0, 0, y0), ghost=(y0, Hc - y1), stream=stream)
            ev_k[k].record(stream)
            s_out.wait_event(ev_k[k])
            L.hb_image_read_region_async(C.byref(hb.view(whole_out, roi=(Wc, y1 - y0, 0, y0))), C.c_void_p(h_out.data_ptr() + y0 * Wc), Wc, hb.stream_ptr(s_out))
        ev_done.record(s_out)
        stream.wait_event(ev_done)

    step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    ok = all(torch.equal(h_out[y:y + 4096], want[y:y + 4096].cpu()) for y in range(0, Hc, 4096))
    assert ok, "C4 e2e leg: host result differs from the device-resident run"
    return {"Gpx_s": Wc * Hc / (ms * 1e-3) / 1e9, "ms": ms, "h2d_bytes_per_step": Wc * Hc, "d2h_bytes_per_step": Wc * Hc,
            "api": f"{K} row strips: hb_image_write_region_async -> hb_harris (strip ROI, neighbour rows as ghost rows) -> hb_image_read_region_async on three streams, "
                   "pinned host buffers; host result compared with the device-resident run"}


def c4_cpu(rows=4096, repeats=3):
    """The CPU arm of C4 on a bounded sample: the sample's nine kernels as specialised -emit-cpu shaped loops
    (oracle/emit_cpu_fast.cpp::ocf_harris, pinned bit for bit against the generic oracle) on a 32768 x `rows` strip."""
    import numpy as np
    from hipacc_b200 import synth
    from oracle import oracle as O
    O.set_num_threads(len(os.sched_getaffinity(0)))
    img = synth.image_np("uint8", 32768, rows, seed=4)
    out = np.empty_like(img)
    O.harris_fast(img, out=out)      # warm-up: page faults of the eight intermediate images, OpenMP pool
    best = float("inf")
    for _ in range(repeats):
        t = time.perf_counter()
        O.harris_fast(img, out=out)
        best = min(best, time.perf_counter() - t)
    return {"Gpx_s": img.size / best / 1e9, "cores": O.num_threads(), "kind": "port",
            "sample": f"32768x{rows} strip of the same synthetic image, nine unfused kernels, g++ -O3 AVX2, OpenMP over rows, best of {repeats}"}


def extra_operators(hb, dev, peak, use_graph=True):
    """Gpixels/s and HBM fraction of the other BASELINE.json configs on one GPU (kernel-only, CUDA events)."""
    import numpy as np
    import torch
    from hipacc_b200 import _abi as A, masks as M, specs as S, synth
    stream = torch.cuda.Stream(device=dev)   # a capturable stream: the timed launches replay as one CUDA graph
    res = {}

    def timeit(fn, reps=10, warm=3, graph=True):
        """average device time of fn.  graph=True: `reps` launches are captured once and replayed as ONE CUDA graph, so a
        short kernel (C1: 40 us) is not timed at the rate the Python host can enqueue it; blocking calls pass graph=False."""
        torch.cuda.synchronize()   # the inputs were written on the default stream
        with torch.cuda.stream(stream):
            for _ in range(warm):
                fn()
            torch.cuda.synchronize()
            run, n = fn, reps
            graph = graph and use_graph   # --no-graph (profiling passes): plain launches
            if graph:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=stream, capture_error_mode="thread_local"):
                    for _ in range(reps):
                        fn()
                g.replay()
                torch.cuda.synchronize()
                run, n = g.replay, 3
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(n):
                run()
            e1.record(stream)
            torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (n * reps if graph else n)

    def entry(name, px, alg_bytes, ms, note=""):
        gbs = alg_bytes / (ms * 1e-3) / 1e9
        res[name] = {"Gpx_s": px / (ms * 1e-3) / 1e9, "ms": ms, "alg_GB_s": gbs, "hbm_frac": gbs / peak, "note": note}

    # C1 Gaussian 5x5 uchar 4096^2 CLAMP
    u = hb.empty_image(A.U8, 4096, 4096, device=dev)
    u.copy_(synth.image_torch("uint8", 4096, 4096, seed=1, device=dev))
    uo = hb.empty_image(A.U8, 4096, 4096, device=dev)
    g5 = S.gaussian_blur(M.GAUSS5, A.CLAMP)
    entry("C1_gaussian5x5_u8_4096", 4096 * 4096, 2 * 4096 * 4096, timeit(lambda: hb.local_op(g5, u, dst=uo, stream=stream)),
          "bit-exact float mask: FP32-lane bound (25 separately rounded multiplies + 24 adds per pixel = 744 Gpx/s at 128 lanes/clk/SM); 16 MiB image, 3 waves of CTAs")
    del u, uo
    # the same operator on 16x the pixels: what the kernel sustains once the launch ramp and the last partial wave no longer weigh
    u = hb.empty_image(A.U8, 16384, 16384, device=dev)
    u.copy_(synth.image_torch("uint8", 16384, 16384, seed=1, device=dev))
    uo = hb.empty_image(A.U8, 16384, 16384, device=dev)
    entry("gaussian5x5_u8_16384", 16384 * 16384, 2 * 16384 * 16384, timeit(lambda: hb.local_op(g5, u, dst=uo, stream=stream)),
          "C1's operator on a 16384^2 image (steady state)")
    del u, uo
    # vector pixels: Gaussian_Blur_RGBA's own size, uchar4 (4 B read + 4 B written per pixel)
    rgba = torch.empty((3024, 4032, 4), dtype=torch.uint8, device=dev)
    rgba.view(3024, 4032 * 4).copy_(synth.image_torch("uint8", 4032 * 4, 3024, seed=6, device=dev))
    rgba_o = torch.empty_like(rgba)
    entry("gaussian5x5_rgba_u8x4_4032x3024", 4032 * 3024, 8 * 4032 * 3024, timeit(lambda: hb.local_op(g5, rgba, dst=rgba_o, stream=stream)),
          "uchar4 pixels, float4 accumulate per channel: FP32-issue bound (4 channels x 25 taps per pixel)")
    del rgba, rgba_o
    # C3 bilateral 13x13 float 8192^2 + fused min/max/sum
    f = hb.empty_image(A.F32, 8192, 8192, device=dev)
    f.copy_(synth.image_torch("float32", 8192, 8192, seed=3, scale=255.0, device=dev))
    fo = hb.empty_image(A.F32, 8192, 8192, device=dev)
    cm = M.bilateral_mask(13)
    entry("C3_bilateral13x13_f32_8192", 8192 * 8192, 8 * 8192 * 8192, timeit(lambda: hb.bilateral(f, 13, cm, 16, A.MIRROR, dst=fo, stream=stream), reps=3, warm=1),
          "169 ex2 per pixel: MUFU/FP32-issue bound")
    part = torch.zeros(4, dtype=torch.float32, device=dev)
    hbins = torch.zeros(256, dtype=torch.int32, device=dev)
    entry("C3_reduce_minmaxsum_f32_8192", 8192 * 8192, 4 * 8192 * 8192, timeit(lambda: hb.reduce_minmaxsum_async(fo, part, stream=stream)),
          "fused min+max+sum, one pass")
    entry("hist256_f32_8192", 8192 * 8192, 4 * 8192 * 8192, timeit(lambda: hb.binning_async(f, hbins, stream=stream)),
          "binning(): 256-bin histogram, per-warp shared-memory bins, one pass")
    # 1 read + 1 write references at the same size and timing method: what "HBM roofline" means in this loop
    c_src, c_dst = f, fo
    entry("ref_copy_torch_f32_8192", 8192 * 8192, 8 * 8192 * 8192, timeit(lambda: c_dst.copy_(c_src)), "torch copy_ (library kernel), measurement reference only")
    entry("point_copy_f32_8192", 8192 * 8192, 8 * 8192 * 8192, timeit(lambda: hb.point_op(A.POINT_COPY, [c_src], A.F32, dst=c_dst, stream=stream)), "hb_point_op COPY")
    # C2 single operators
    for nm, m in (("sobel_x", M.SOBEL3_X), ("laplace", M.LAPLACE3)):
        sp = S.domain_reduce_f32(m.astype(np.float32), A.MIRROR)
        entry(f"C2_{nm}_f32_8192", 8192 * 8192, 8 * 8192 * 8192, timeit(lambda: hb.local_op(sp, f, dst=fo, stream=stream)))
    del f, fo
    # C4 Harris fused, one 32768 x 4096 strip (the per-GPU share at 8 GPUs)
    hs = hb.empty_image(A.U8, 32768, 4096, device=dev)
    hs.copy_(synth.image_torch("uint8", 32768, 4096, seed=4, device=dev))
    ho = hb.empty_image(A.U8, 32768, 4096, device=dev)
    entry("C4_harris_fused_u8_32768x4096", 32768 * 4096, 2 * 32768 * 4096, timeit(lambda: hb.harris(hs, dst=ho, stream=stream), reps=5),
          "fused 9-kernel pipeline, integer-issue bound")
    del hs, ho
    # Reduction_Sum sample: int 4096 x 4096 (vectorised integer kernel); hb_reduce_async leaves the scalar in HBM, so the
    # kernel is timed like every other entry (graph replays); the blocking call adds the 16-byte read-back + a host sync
    it = torch.randint(-100, 100, (4096, 4096), dtype=torch.int32, device=dev)
    r8 = torch.zeros(2, dtype=torch.int32, device=dev)
    entry("reduce_sum_s32_4096", 4096 * 4096, 4 * 4096 * 4096, timeit(lambda: hb.reduce_async(it, A.SUM, r8, stream=stream)), "hb_reduce_async (Reduction_Sum's shape)")
    entry("reduce_sum_s32_4096_blocking", 4096 * 4096, 4 * 4096 * 4096, timeit(lambda: hb.reduce(it, A.SUM, stream=stream), graph=False), "blocking hb_reduce incl. the result read-back")
    u8r = torch.randint(0, 255, (8192, 8192), dtype=torch.uint8, device=dev)
    entry("reduce_sum_u8_8192", 8192 * 8192, 8192 * 8192, timeit(lambda: hb.reduce_async(u8r, A.SUM, r8, stream=stream)), "hb_reduce_async, IDP.4A byte sums")
    i16 = torch.randint(-100, 100, (16384, 16384), dtype=torch.int32, device=dev)
    entry("reduce_sum_s32_16384", 16384 * 16384, 4 * 16384 * 16384, timeit(lambda: hb.reduce_async(i16, A.SUM, r8, stream=stream)), "hb_reduce_async, 1 GiB image")
    return res


def extra_sharded(hb, dev, world, rank, stream, p2p=True):
    """N > 1 extras: the FIRST sharded-pyramid design for comparison (one halo exchange per level transition: 14 exchange
    launches per traversal; the default line times the one-exchange + one-all-gather design) and the C3 reductions
    (per-rank fused min/max/sum + one all-gather of the partials).  Device-timed, max over ranks."""
    import torch
    import torch.distributed as dist
    from hipacc_b200 import _abi as A, masks as M, strips, synth
    res = {}

    def graphed(fn):
        """capture fn (kernels + NCCL halo exchanges) into a CUDA graph; every rank falls back together"""
        fn()
        torch.cuda.synchronize()
        ok = 1
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream, capture_error_mode="thread_local"):
                fn()
        except Exception as e:  # noqa: BLE001
            ok = 0
            sys.stderr.write(f"[rank {rank}] graph capture failed: {e}\n")
        t_ok = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(t_ok, op=dist.ReduceOp.MIN)
        return (g.replay, "CUDA graph replay") if int(t_ok.item()) else (fn, "direct launches")

    def timeit(fn, reps=5, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # C5: 8-level pyramid of a 16384 x 16384 float image on row strips
    Wp = Hp = 16384
    pg = strips.StripPyramid(Wp, Hp, 8, world, rank, radius=4, device=dev, hb=hb)
    pl = strips.StripPyramid(Wp, Hp, 8, world, rank, radius=4, device=dev, hb=hb)
    if p2p:
        pg.enable_p2p(hb)
        pl.enable_p2p(hb)
    pg.owned(0).copy_(synth.image_torch("float32", Wp, pg.plans[0].rows, seed=5, y0=pg.plans[0].y0, device=dev))
    traverse = lambda: strips.pyramid_traverse_strips(hb, pg, pl, M.GAUSS5, stream=stream)  # noqa: E731
    fn, how = graphed(traverse) if p2p else (traverse, "direct launches")
    ms = timeit(fn, reps=3, warm=1)
    res["C5_pyramid8_f32_16384_sharded_exchange_per_level"] = {"Gpx_s": Wp * Hp / (ms * 1e-3) / 1e9, "ms": ms, "n_gpus": world, "launch": how,
                                            "note": "strong scaling: 14 halo exchange launches (4 rows per neighbour, " + ("peer-to-peer push kernels" if p2p else "NCCL send/recv") + ") + 14 fused level kernels per traversal"}
    del pg, pl
    # C3 reductions: per-rank fused min/max/sum partials + all-reduce (weak: 8192 x 8192 per rank)
    f = hb.empty_image(A.F32, 8192, 8192, device=dev)
    f.copy_(synth.image_torch("float32", 8192, 8192, seed=3, scale=255.0, y0=8192 * rank, device=dev))
    part = torch.zeros(4, dtype=torch.float32, device=dev)

    def reduce_step():
        hb.reduce_minmaxsum_async(f, part, stream=stream)
        strips.allgather_minmaxsum(part)        # one 16-byte all-gather + local fold
    fn, how = graphed(reduce_step)
    ms = timeit(fn)
    res["C3_reduce_minmaxsum_f32_8192_per_rank_allreduce"] = {"Gpx_s": 8192 * 8192 * world / (ms * 1e-3) / 1e9, "ms": ms, "n_gpus": world, "launch": how,
                                                              "note": "weak scaling: one pass over HBM per rank + one 16-byte NCCL all-gather of the {min, max, sum} partials"}
    return res


if __name__ == "__main__":
    main()
