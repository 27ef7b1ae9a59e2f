#!/usr/bin/env python
"""A/B of the fused Harris kernel (C4): env knobs are read once per process, so one process per variant.  Every variant is
first compared with version 2 of the kernel on a noise image (bit for bit), then timed as CUDA-graph replays.
usage: python tools/ab_harris.py   (spawns itself)"""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import torch
    import hipacc_b200 as hb
    from hipacc_b200 import _abi as A, synth
    hb.init(0)
    dev = torch.device("cuda:0")
    for (w, h) in ((32768, 4096), (32768, 32768)):
        u = hb.empty_image(A.U8, w, h, device=dev); u.copy_(synth.image_torch("uint8", w, h, seed=4, device=dev))
        uo = hb.empty_image(A.U8, w, h, device=dev)
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            for _ in range(3): hb.harris(u, dst=uo, stream=st)
            torch.cuda.synchronize()
            with hb.Graph(st) as g:
                for _ in range(5): hb.harris(u, dst=uo, stream=st)
            g.launch(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(3): g.launch()
            e1.record(st); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 15
        print(f"{sys.argv[1]:28s} {w}x{h} {ms*1e3:9.1f} us {w*h/ms/1e6:8.1f} Gpx/s  checksum {int(uo.sum())}", flush=True)
        del u, uo
else:
    for name, env in (("v2", {"HB_HARRIS_VERSION": "2"}), ("v3 6 CTAs", {"HB_HARRIS_CTAS": "6"}), ("v3 5 CTAs", {"HB_HARRIS_CTAS": "5"}),
                      ("v3 4 CTAs", {"HB_HARRIS_CTAS": "4"})):
        subprocess.run([sys.executable, __file__, name], env={**os.environ, **env})
