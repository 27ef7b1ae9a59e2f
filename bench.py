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

--impl reference: the reference's CPU path for the same workload (the restated -emit-cpu code in
oracle/, all host threads; the Hipacc compiler itself cannot be built here -- DESIGN.md) on a bounded
sample.  Rank 0 only.
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
def cpu_baseline(rows, repeats=3):
    """The oracle ("port" of -emit-cpu, oracle/emit_cpu.cpp) on the host cores: the three operators on a
    W x rows sample of the same synthetic image.  Returns (Gpx/s, cores, sample description)."""
    import numpy as np
    from hipacc_b200 import synth
    from oracle import oracle as O
    O.set_num_threads(len(os.sched_getaffinity(0)))
    img = synth.image_np("float32", W, rows, seed=2)
    specs = specs_for_workload()
    outs = [np.empty_like(img) for _ in specs]
    for s, o in zip(specs, outs):   # warm-up (page faults, OpenMP pool)
        O.local_op(s, img, out=o)
    best = float("inf")
    for _ in range(repeats):
        t = time.perf_counter()
        for s, o in zip(specs, outs):
            O.local_op(s, img, out=o)
        best = min(best, time.perf_counter() - t)
    return len(specs) * W * rows / best / 1e9, O.num_threads(), f"3 operators on {W}x{rows} float rows of the same image, best of {repeats}"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = max(1, args.steps), max(1, args.warmup)
    import numpy as np
    from hipacc_b200 import synth
    from oracle import oracle as O
    O.set_num_threads(len(os.sched_getaffinity(0)))   # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every host core
    rows = 4096   # bounded sample: half of the image per step
    img = synth.image_np("float32", W, rows, seed=2)
    specs = specs_for_workload()
    outs = [np.empty_like(img) for _ in specs]

    def step():
        for s, o in zip(specs, outs):
            O.local_op(s, img, out=o)
    for _ in range(warm):
        step()
    t = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t) / steps
    val = len(specs) * W * rows / dt / 1e9
    cores = O.num_threads()
    sample = f"{W}x{rows} float rows per step (half of the 8192x8192 image), restated -emit-cpu code, OpenMP {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "Gpixels/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C2: Sobel-X + Sobel-Y + Laplace 3x3 float 8192x8192 MIRROR (CPU sample)", "sample": sample},
        "cpu_baseline": {"value": val, "unit": "Gpixels/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "Gpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


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
                "note": "frac can slightly exceed 1: consecutive operators re-read the same 256 MiB input and a part of it still sits in the 126 MB L2; single-operator launches reach 0.95 (operators.C2_*)", "peak_source": peak_src,
                "algorithmic_bytes_per_launch": ALG_BYTES_PER_PX * W * plan.rows, "avg_launch_ms": per_launch_ms}
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        try:
            roofline["traffic"] = json.load(open(traffic_file)).get("local_tiled_f32_3x3_bytes_per_launch")
        except Exception:
            pass

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

    if args.extra:
        operators = extra_operators(hb, dev, peak) if world == 1 else extra_sharded(hb, dev, world, rank, stream, halo is not None)
    else:
        operators = None

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, cores, sample = cpu_baseline(rows=4096)
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
        print(json.dumps(line))
    if world > 1:
        if halo is not None and rank == 0:
            n_ex, timed_out = halo.status()
            if timed_out:
                sys.stderr.write("WARNING: a halo exchange timed out waiting for a neighbour\n")
        sys.stdout.flush()
        torch.cuda.synchronize()
        dist.barrier()
        os._exit(0)   # skip interpreter teardown: CUDA graphs / IPC mappings and the NCCL communicator do not need an orderly exit


def extra_operators(hb, dev, peak):
    """Gpixels/s and HBM fraction of the other BASELINE.json configs on one GPU (kernel-only, CUDA events)."""
    import numpy as np
    import torch
    from hipacc_b200 import _abi as A, masks as M, specs as S, synth
    stream = torch.cuda.current_stream()
    res = {}

    def timeit(fn, reps=10, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    def entry(name, px, alg_bytes, ms, note=""):
        gbs = alg_bytes / (ms * 1e-3) / 1e9
        res[name] = {"Gpx_s": px / (ms * 1e-3) / 1e9, "ms": ms, "alg_GB_s": gbs, "hbm_frac": gbs / peak, "note": note}

    # C1 Gaussian 5x5 uchar 4096^2 CLAMP
    u = hb.empty_image(A.U8, 4096, 4096, device=dev)
    u.copy_(synth.image_torch("uint8", 4096, 4096, seed=1, device=dev))
    uo = hb.empty_image(A.U8, 4096, 4096, device=dev)
    g5 = S.gaussian_blur(M.GAUSS5, A.CLAMP)
    entry("C1_gaussian5x5_u8_4096", 4096 * 4096, 2 * 4096 * 4096, timeit(lambda: hb.local_op(g5, u, dst=uo, stream=stream)),
          "bit-exact float mask: FP32-issue bound, 16 MiB image fits L2")
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
    # C5 pyramid 8 levels float 16384^2
    base = hb.empty_image(A.F32, 16384, 16384, device=dev)
    base.copy_(synth.image_torch("float32", 16384, 16384, seed=5, device=dev))
    pg = hb.Pyramid(base, 8)
    pl = hb.Pyramid(hb.empty_image(A.F32, 16384, 16384, device=dev).zero_(), 8)
    hb.pyramid_traverse(pg, pl, M.GAUSS5, stream=stream)
    torch.cuda.synchronize()
    with hb.Graph(stream) as g:                      # hb_graph_begin / hb_graph_end: the 14 level kernels as one launch
        hb.pyramid_traverse(pg, pl, M.GAUSS5, stream=stream)
    ms = timeit(lambda: g.launch(), reps=5, warm=2)
    g.destroy()
    n = 16384 * 16384
    entry("C5_pyramid8_f32_16384", n, int(23 * n * 4 / 3), ms,
          "fused down (blur+subsample+DoG) 9n + fused up (Restore+Blend) 14n bytes per transition; 14 kernels replayed as one hb_graph launch")
    return res


def extra_sharded(hb, dev, world, rank, stream, p2p=True):
    """N > 1: the sharded BASELINE configs (strong scaling: the named global image cut into `world` row strips).
    C4 Harris 32768^2 uchar (halo exchange + fused kernel per step), C5 pyramid 16384^2 float, 8 levels (one halo
    exchange per level transition), C3 fused min/max/sum + one all-reduce per scalar.  Device-timed, max over ranks."""
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

    # C4: Harris on a 32768 x 32768 uchar image, `world` row strips with 2 ghost rows
    Wc, Hc = 32768, 32768
    plan = strips.StripPlan(Wc, Hc, world, rank, radius=2, boundary=A.CLAMP)
    buf = hb.alloc_image(A.U8, Wc, plan.buffer_rows, device=dev)
    strips.owned(buf, plan).copy_(synth.image_torch("uint8", Wc, plan.rows, seed=4, y0=plan.y0, device=dev))
    out = torch.empty_like(buf)
    halo = strips.P2PHalo(hb, buf, plan) if p2p else None

    def harris_step():
        if halo is not None:
            halo.exchange(stream)
        else:
            strips.exchange_halos(buf, plan)
        hb.harris(buf, dst=out, roi=plan.roi(), ghost=plan.ghost(), stream=stream)
    fn, how = graphed(harris_step) if p2p else (harris_step, "direct launches")
    ms = timeit(fn)
    res["C4_harris_fused_u8_32768x32768_sharded"] = {"Gpx_s": Wc * Hc / (ms * 1e-3) / 1e9, "ms": ms, "n_gpus": world, "launch": how,
                                                      "note": f"strong scaling: {plan.rows} rows per rank + 2 ghost rows exchanged per step (" + ("peer-to-peer push kernel" if p2p else "NCCL send/recv") + ")"}
    del buf, out
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
    res["C5_pyramid8_f32_16384_sharded"] = {"Gpx_s": Wp * Hp / (ms * 1e-3) / 1e9, "ms": ms, "n_gpus": world, "launch": how,
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
