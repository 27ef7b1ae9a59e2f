"""Small repro / smoke of the float TMA local-operator kernel (used under compute-sanitizer)."""
import sys
import numpy as np
import torch
import hipacc_b200 as hb
from hipacc_b200 import _abi as A, masks as M, specs as S, synth
from oracle import oracle as O

hb.init(0)
dev = torch.device("cuda:0")
w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (517, 391)
img = synth.image_np("float32", w, h, seed=21)
d = hb.empty_image(A.F32, w, h, device=dev)
d.copy_(torch.from_numpy(img))
for name, m, b in (("sobelx", M.SOBEL3_X, A.CLAMP), ("lap", M.LAPLACE3, A.MIRROR), ("g5", M.GAUSS5, A.REPEAT), ("g7", M.GAUSS7, A.CONSTANT)):
    spec = S.domain_reduce_f32(m.astype(np.float32), b) if m.dtype != np.float32 else S.convolve_f32(m, b)
    got = hb.local_op(spec, d)
    torch.cuda.synchronize()
    ok = np.array_equal(got.cpu().numpy(), O.local_op(spec, img))
    print(name, "ok" if ok else "MISMATCH")
