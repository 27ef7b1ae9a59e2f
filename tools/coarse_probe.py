"""coarse end of the C5 pyramid (levels 4..7: 1024^2 .. 128^2): per-level kernels vs the one-launch cooperative kernel, graph replays"""
import os, sys; sys.path.insert(0, '.')
import torch
import hipacc_b200 as hb
from hipacc_b200 import _abi as A, masks as M, synth
hb.init(0); dev = torch.device('cuda:0')
st = torch.cuda.Stream(device=dev); torch.cuda.set_stream(st)
base = hb.empty_image(A.F32, 1024, 1024, device=dev); base.copy_(synth.image_torch('float32', 1024, 1024, seed=5, device=dev))
pg = hb.Pyramid(base, 4); pl = hb.Pyramid(hb.empty_image(A.F32, 1024, 1024, device=dev).zero_(), 4)
def t(fn, reps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps): fn()
    e1.record(st); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps * 1e3
for name, f in (("per-level", lambda: hb.pyramid_traverse(pg, pl, M.GAUSS5, stream=st, fuse_coarse=False)),
                ("one launch", lambda: hb.pyr_traverse_coarse(pg.levels, pl.levels, M.GAUSS5, stream=st))):
    f(); torch.cuda.synchronize()
    with hb.Graph(st) as g: f()
    print(name, "HB_COARSE_CTAS_PER_SM=" + os.environ.get("HB_COARSE_CTAS_PER_SM", "default"), round(t(lambda: g.launch()), 1), "us per traversal (graph replay)")
    g.destroy()
