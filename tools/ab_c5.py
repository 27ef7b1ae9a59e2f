#!/usr/bin/env python
"""A/B of the pyramid kernels across library builds (HIPACC_B200_LIB): the level-0 down / up steps alone and the 8-level
traversal of a 16384^2 float image, CUDA-graph replays.   usage: python tools/ab_c5.py lib1.so lib2.so ..."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if os.environ.get("AB_CHILD"):
    import torch
    import hipacc_b200 as hb
    from hipacc_b200 import _abi as A, masks as M, synth
    hb.init(0)
    dev = torch.device("cuda:0")
    n = 16384
    st = torch.cuda.Stream()
    img = hb.empty_image(A.F32, n, n, device=dev)
    for y in range(0, n, 2048):
        img[y:y + 2048].copy_(synth.image_torch("float32", n, 2048, seed=5, y0=y, device=dev))
    g1 = hb.empty_image(A.F32, n // 2, n // 2, device=dev); l0 = hb.empty_image(A.F32, n, n, device=dev)
    pg = hb.Pyramid(img, 8); pl = hb.Pyramid(hb.empty_image(A.F32, n, n, device=dev).zero_(), 8)
    def timeit(fn, reps):
        with torch.cuda.stream(st):
            fn(); torch.cuda.synchronize()
            with hb.Graph(st) as g:
                for _ in range(reps): fn()
            g.launch(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(3): g.launch()
            e1.record(st); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (3 * reps) * 1e3
    t_down = timeit(lambda: hb.pyr_down(img, g1, M.GAUSS5, lap_fine=l0, stream=st), 5)
    t_trav = timeit(lambda: hb.pyramid_traverse(pg, pl, M.GAUSS5, stream=st), 3)
    print(f"{os.path.basename(hb.LIB_PATH):22s} {os.environ.get('AB_TAG', ''):10s} down L0 {t_down:8.1f} us   traversal {t_trav:8.1f} us = {n * n / t_trav / 1e3:6.1f} Gpx/s", flush=True)
else:
    # every library is run with and without TMA staging of the fused down step (HB_PYR_NO_TMA, read once per process)
    for lib in (sys.argv[1:] or [os.path.join(ROOT, "hipacc_b200", "lib", "libhipacc_b200.so")]):
        for tag, env in (("tma", {}), ("no-tma", {"HB_PYR_NO_TMA": "1"})):
            subprocess.run([sys.executable, __file__], env={**os.environ, **env, "AB_CHILD": "1", "AB_TAG": tag, "HIPACC_B200_LIB": os.path.abspath(lib)})
