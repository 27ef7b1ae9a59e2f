"""Peer-to-peer halo exchange (hb_halo_exchange, CUDA IPC + device-side flags) across real GPUs.

Needs >= 2 GPUs, so it is skipped on single-GPU boxes; run it on a multi-GPU box with
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/test_p2p_halo.py
(the pytest entry below spawns exactly that).  Every rank owns a strip of a synthetic global image whose pixels
change every round; after each exchange the ghost rows must equal the neighbours' rows of THAT round (a stale or
early push would be caught), for CLAMP (no wrap) and REPEAT (cyclic neighbours) layouts, uchar and float."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def worker():
    import torch
    import torch.distributed as dist
    import hipacc_b200 as hb
    from hipacc_b200 import _abi as A, strips, synth

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    hb.init(local)
    dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ok = True
    for dtype, tname, W, H, R, boundary in ((A.F32, "float32", 1000, 64 * world, 3, A.CLAMP), (A.U8, "uint8", 4099, 40 * world, 2, A.REPEAT),
                                            (A.F32, "float32", 8192, 256 * world, 1, A.MIRROR)):
        plan = strips.StripPlan(W, H, world, rank, R, boundary)
        buf = hb.alloc_image(dtype, W, plan.buffer_rows, device=dev)
        halo = strips.P2PHalo(hb, buf, plan)
        for rnd in range(6):
            strips.owned(buf, plan)[:, :W] = synth.image_torch(tname, W, plan.rows, seed=100 + rnd, y0=plan.y0, device=dev)
            halo.exchange(stream)
            # what the ghost rows must hold: the global image's rows around the strip (cyclic for REPEAT)
            for g0, n, ys in ((0, plan.ghost_top, plan.y0 - plan.ghost_top), (plan.ghost_top + plan.rows, plan.ghost_bottom, plan.y1)):
                for k in range(n):
                    want = synth.image_torch(tname, W, 1, seed=100 + rnd, y0=(ys + k) % H, device=dev)[0]
                    if not torch.equal(buf[g0 + k, :W], want):
                        ok = False
                        print(f"[rank {rank}] {tname} round {rnd}: ghost row {g0 + k} differs", flush=True)
        n_ex, timed_out = halo.status()
        ok = ok and n_ex == 6 and timed_out == 0
    t = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("P2P_HALO_OK" if int(t.item()) else "P2P_HALO_FAILED", flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    os._exit(0 if int(t.item()) else 1)


@pytest.mark.gpu
def test_p2p_halo_exchange_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    n = min(torch.cuda.device_count(), 4)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.abspath(__file__)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "P2P_HALO_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


if __name__ == "__main__":
    worker()
