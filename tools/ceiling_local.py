import sys; sys.path.insert(0, '.')
import numpy as np, torch
import hipacc_b200 as hb
from hipacc_b200 import _abi as A, specs as S, masks as M, synth
hb.init(0); dev = torch.device('cuda:0')
f = hb.empty_image(A.F32, 8192, 8192, device=dev); f.copy_(synth.image_torch('float32', 8192, 8192, seed=2, device=dev))
o = hb.empty_image(A.F32, 8192, 8192, device=dev)
st = torch.cuda.current_stream()
def t(fn, reps=30):
    for _ in range(5): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps): fn()
    e1.record(st); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps
ident = np.zeros((3, 3), np.float32); ident[1, 1] = 1
for name, m in (('identity', ident), ('sobel_x', M.SOBEL3_X.astype(np.float32)), ('laplace', M.LAPLACE3.astype(np.float32))):
    sp = S.domain_reduce_f32(m, A.MIRROR)
    ms = t(lambda: hb.local_op(sp, f, dst=o, stream=st))
    print(name, round(8192 * 8192 / ms / 1e6, 1), 'Gpx/s', round(ms * 1e3, 1), 'us')
ms = t(lambda: o.copy_(f)); print('torch copy', round(8192 * 8192 / ms / 1e6, 1), 'Gpx/s')
ms = t(lambda: hb.point_op(A.POINT_COPY, [f], A.F32, dst=o, stream=st)); print('point copy', round(8192 * 8192 / ms / 1e6, 1), 'Gpx/s')
