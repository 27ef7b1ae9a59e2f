"""separable integer masks: two-pass variant vs the 2-D fold (HB_NO_SEPARABLE=1), uchar 8192^2 -> int"""
import os, sys; sys.path.insert(0, '.')
import numpy as np, torch
import hipacc_b200 as hb
from hipacc_b200 import _abi as A, specs as S, masks as M, synth
hb.init(0); dev = torch.device('cuda:0')
u = hb.empty_image(A.U8, 8192, 8192, device=dev); u.copy_(synth.image_torch('uint8', 8192, 8192, seed=1, device=dev))
st = torch.cuda.current_stream()
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps): fn()
    e1.record(st); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps
for name, m in (('sobel3x', M.SOBEL3_X), ('sobel5x', M.SOBEL5_X), ('sobel7x', getattr(M, 'SOBEL7_X', None)), ('laplace5 (not separable)', M.LAPLACE5)):
    if m is None: continue
    sp = S.sobel_u8(m, A.CLAMP)
    o = torch.zeros((8192, 8192), dtype=torch.int32, device=dev)
    ms = t(lambda: hb.local_op(sp, u, dst=o, stream=st))
    print(name, 'HB_NO_SEPARABLE=' + os.environ.get('HB_NO_SEPARABLE', '0'), round(8192 * 8192 / ms / 1e6, 1), 'Gpx/s', int(o.sum().item()))
