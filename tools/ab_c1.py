#!/usr/bin/env python
"""A/B of the C1 operator (Gaussian 5x5 uchar CLAMP): env knobs are read once per process, so one process per variant.
usage: python tools/ab_c1.py   (spawns itself)"""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import torch
    import hipacc_b200 as hb
    from hipacc_b200 import _abi as A, masks as M, specs as S, synth
    hb.init(0)
    dev = torch.device("cuda:0")
    g5 = S.gaussian_blur(M.GAUSS5, A.CLAMP)
    for n in (4096, 8192, 16384):
        u = hb.empty_image(A.U8, n, n, device=dev); u.copy_(synth.image_torch("uint8", n, n, seed=1, device=dev))
        uo = hb.empty_image(A.U8, n, n, device=dev)
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            for _ in range(5): hb.local_op(g5, u, dst=uo, stream=st)
            torch.cuda.synchronize()
            with hb.Graph(st) as g:
                for _ in range(20): hb.local_op(g5, u, dst=uo, stream=st)
            g.launch(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(3): g.launch()
            e1.record(st); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 60
        print(f"{sys.argv[1]:28s} {n:6d}^2 {ms*1e3:9.1f} us {n*n/ms/1e6:8.1f} Gpx/s", flush=True)
else:
    for name, env in (("pair RPT=2", {"HB_PAIR_RPT": "2"}), ("pair RPT=4", {"HB_PAIR_RPT": "4"}), ("tiled (HB_NO_PAIR)", {"HB_NO_PAIR": "1"})):
        subprocess.run([sys.executable, __file__, name], env={**os.environ, **env})
