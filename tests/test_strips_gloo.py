"""CPU tests of the multi-GPU host logic (world_size 2 and 3, gloo): strip partition + halo exchange
reproduce the boundary-extended global image, and the oracle applied per strip with ghost rows equals
the oracle on the whole image (the property the sharded CUDA path relies on)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hipacc_b200 import _abi as A, masks as M, specs as S, strips, synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, boundary, radius, H, W, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        plan = strips.StripPlan(W, H, world, rank, radius, boundary)
        plan.validate()
        stride = W + 5  # padded rows: whole-row messages include the padding
        buf = torch.full((plan.buffer_rows, stride), -1.0, dtype=torch.float32)
        strips.owned(buf, plan)[:, :W] = torch.from_numpy(synth.image_np("float32", W, plan.rows, seed=3, y0=plan.y0))
        strips.exchange_halos(buf, plan)
        mn, mx, sm = strips.allreduce_minmaxsum(float(strips.owned(buf, plan)[:, :W].min()), float(strips.owned(buf, plan)[:, :W].max()),
                                                float(strips.owned(buf, plan)[:, :W].double().sum()))
        # the single-collective combine of the device partial records {float min, float max, double sum}
        own = strips.owned(buf, plan)[:, :W]
        rec = torch.zeros(4, dtype=torch.float32)
        rec[0], rec[1] = own.min(), own.max()
        rec[2:4].view(torch.float64)[0] = own.double().sum()
        g3 = strips.allgather_minmaxsum(rec)
        assert float(g3[0]) == mn and float(g3[1]) == mx and abs(float(g3[2]) - sm) <= 1e-12 * abs(sm)
        q.put((rank, plan.y0, plan.y1, plan.ghost_top, plan.ghost_bottom, buf[:, :W].numpy().copy(), (mn, mx, sm)))
    finally:
        dist.destroy_process_group()


def _run(world, boundary, radius, H=37, W=19):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, boundary, radius, H, W, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("boundary", [A.MIRROR, A.CLAMP, A.REPEAT])
def test_halo_exchange_and_strip_parity(oracle, world, boundary):
    H, W, R = 37, 19, 2
    full = synth.image_np("float32", W, H, seed=3)
    res = _run(world, boundary, R, H, W)
    spec = S.domain_reduce_f32(M.LAPLACE5.astype(np.float32), boundary)
    want = oracle.local_op(spec, full)
    rows = 0
    for rank, y0, y1, gt, gb, buf, red in res:
        # ghost rows hold the neighbours' real rows (cyclic neighbours for REPEAT)
        ext = np.concatenate([full[(np.arange(y0 - gt, y0)) % H], full[y0:y1], full[(np.arange(y1, y1 + gb)) % H]])
        np.testing.assert_array_equal(buf, ext)
        roi = (W, y1 - y0, 0, gt)
        got = oracle.local_op(spec, np.ascontiguousarray(buf), roi_in=roi, roi_out=roi, ghost=(gt, gb))
        np.testing.assert_array_equal(got[gt:gt + y1 - y0], want[y0:y1])   # strip result == rows of the global result
        rows += y1 - y0
        assert np.float32(red[0]) == full.min() and np.float32(red[1]) == full.max()
        assert abs(red[2] - full.astype(np.float64).sum()) < 1e-9 * full.sum()
    assert rows == H


def test_plan_partitions_every_row_once():
    for H, world in [(37, 3), (8192, 8), (4097, 4), (5, 5)]:
        plans = [strips.StripPlan(16, H, world, r, 1) for r in range(world)]
        assert plans[0].y0 == 0 and plans[-1].y1 == H
        assert all(plans[i].y1 == plans[i + 1].y0 for i in range(world - 1))
        assert plans[0].ghost_top == 0 and plans[-1].ghost_bottom == 0
    with pytest.raises(AssertionError):
        strips.StripPlan(16, 8, 8, 0, 2).validate()


def test_strip_pyramid_levels_align_across_ranks():
    """every level is cut at the same relative rows: strip boundaries halve exactly from level to level, so a strip's
    local row parity equals the global one (the condition the fused level kernels rely on)"""
    for world, H, W, depth, R in [(2, 256, 392, 3, 4), (8, 16384, 16384, 8, 4), (4, 512, 264, 4, 3), (3, 384, 520, 3, 5)]:
        pyrs = [strips.StripPyramid(W, H, depth, world, r, R, "cpu") for r in range(world)]
        for l in range(depth):
            plans = [p.plans[l] for p in pyrs]
            assert plans[0].y0 == 0 and plans[-1].y1 == H >> l
            assert all(plans[i].y1 == plans[i + 1].y0 for i in range(world - 1))
            assert all(pl.y0 == pyrs[i].plans[0].y0 >> l and pl.rows == pyrs[i].plans[0].rows >> l for i, pl in enumerate(plans))
            assert all(pl.rows % 2 == 0 or l == depth - 1 for pl in plans)
            assert plans[0].ghost_top == 0 and plans[-1].ghost_bottom == 0 and all(pl.ghost_top == R for pl in plans[1:])
            assert all(tuple(p.bufs[l].shape)[0] == p.plans[l].buffer_rows for p in pyrs)
    with pytest.raises(AssertionError):
        strips.StripPyramid(256, 100, 3, 2, 0, 4, "cpu")   # 100 rows cannot be cut into 2 strips of multiples of 4


def test_pyramid_shard_plan_geometry():
    """PyramidShardPlan (one exchange + one all-gather): every region a level transition touches lies inside the rows
    the rank holds, fine regions are twice their coarse regions and start on even global rows, and the fused down
    kernel always finds its K ghost rows beyond the fine region on interior sides."""
    for world, H, W, depth, sz in [(8, 16384, 16384, 8, 5), (2, 16384, 16384, 8, 5), (4, 1024, 264, 4, 3), (3, 768, 520, 3, 7), (1, 256, 256, 4, 5)]:
        plans = [strips.PyramidShardPlan(W, H, depth, world, r, sz) for r in range(world)]
        for p in plans:
            assert 1 <= p.G <= max(depth - 1, 1)
            if world == 1:
                assert p.E0 == 0
                continue
            assert p.E0 == p.V[0] and p.E0 <= p.rows(0)
            for l in range(1, p.G + 1):      # way down
                c = p.span(l, p.e[l])
                f = (2 * c[0], 2 * c[1])
                roi, ghost = p.view_args(l - 1, f, p.K)
                assert f[0] % 2 == 0 and roi[1] == 2 * (c[1] - c[0])
                assert ghost[0] == (p.K if p.rank > 0 else 0) and ghost[1] == (p.K if p.rank < world - 1 else 0)
                p.view_args(l, c, 0)         # asserts containment
            for l in range(p.G - 1, -1, -1):  # way up
                f = p.span(l, p.f[l])
                c = (f[0] // 2, f[1] // 2)
                assert f[0] % 2 == 0 and f[1] % 2 == 0
                _, ghost = p.view_args(l + 1, c, 1)
                assert ghost[0] == (1 if p.rank > 0 else 0) and ghost[1] == (1 if p.rank < world - 1 else 0)
                # the Laplacian level written on the way down covers what the up step reads and rewrites
                d = p.span(l + 1, p.e[l + 1])
                assert 2 * d[0] <= f[0] and f[1] <= 2 * d[1]
        # all ranks agree on G and the strips tile every level
        assert len({p.G for p in plans}) == 1
        for l in range(depth):
            assert plans[0].y0(l) == 0 and plans[-1].y1(l) == H >> l
            assert all(plans[i].y1(l) == plans[i + 1].y0(l) for i in range(world - 1))
    with pytest.raises(AssertionError):
        strips.PyramidShardPlan(256, 256, 6, 8, 1, 5, gather_level=5)   # 32-row strips cannot hold the extension rows
