"""world_size-2 tests of the multi-GPU host logic (vdb_mapping_b200/dist.py). On CPU (gloo) the engine is an
oracle-backed test double; with >= 2 GPUs the same worker runs the product CUDA engine over NCCL."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "dist_worker.py")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(world, backend, mode, tmp_path, extra=()):
    out = tmp_path / f"result_{backend}_{mode}.txt"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), WORKER, "--backend", backend, "--out", str(out), "--mode", mode, *extra]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    res = out.read_text()
    assert res.startswith("OK"), res
    return res


def test_owner_function_matches_the_library():
    from vdb_mapping_b200 import dist as vdist
    from vdb_mapping_b200.mapping import leaf_owner
    rng = np.random.default_rng(5)
    for _ in range(500):
        o = (rng.integers(-5000, 5000, 3) * 8).astype(np.int32)
        for w in (2, 3, 4, 8):
            assert vdist.leaf_owner_py(o, w) == leaf_owner(o, w)


def test_split_points_covers_the_cloud():
    from vdb_mapping_b200.dist import split_points
    for n in (0, 1, 7, 262144, 1000001):
        for w in (1, 2, 4, 8):
            spans = [split_points(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


@pytest.mark.parametrize("mode", ["own_cloud", "split"])
def test_exchange_logic_world2_gloo(mode, tmp_path):
    res = _run(2, "gloo", mode, tmp_path)
    sizes = eval(res.split("shards=")[1].split(" total")[0])
    assert len(sizes) == 2 and min(sizes) > 0


def test_sector_ownership_world2_gloo(tmp_path):
    """rays split by azimuth sector, map owned by the same sectors (ShardPlan mode 1): union of shards == oracle"""
    res = _run(2, "gloo", "split", tmp_path, extra=("--plan", "sectors"))
    sizes = eval(res.split("shards=")[1].split(" total")[0])
    assert len(sizes) == 2 and min(sizes) > 0


def test_sector_planner_properties():
    from vdb_mapping_b200 import dist as vdist, scans
    pts, origin = scans.make_scan(1, 0)
    for world in (2, 3, 8, 16):
        plan = vdist.plan_sectors(pts, origin, 0.1, 50.0, world)
        b = plan.bounds
        assert len(b) == world and b[0] == 0.0 and all(x < y for x, y in zip(b, b[1:])) and b[-1] < 4.0
        parts = [vdist.sector_rays(pts, origin, plan, r) for r in range(world)]
        assert sorted(np.concatenate(parts).tolist()) == list(range(len(pts)))  # every ray on exactly one rank
        v = vdist.ray_visits(pts, origin, 0.1, 50.0)
        share = np.array([v[p].sum() for p in parts]) / v.sum()
        assert share.max() < 1.25 / world, share  # balanced by voxel visits
    # diamond angle is monotone in the true angle
    rng = np.random.default_rng(3)
    d = rng.normal(size=(4000, 2))
    assert (np.argsort(vdist.diamond_angle(d[:, 0], d[:, 1])) == np.argsort(np.mod(np.arctan2(d[:, 1], d[:, 0]), 2 * np.pi))).all()
    # a leaf with the planned owner r lies in sector r
    plan = vdist.plan_sectors(pts, origin, 0.1, 50.0, 4)
    for o in (rng.integers(-300, 300, (200, 3)) * 8):
        r = vdist.leaf_owner_planned_py(o, plan)
        a = float(vdist.diamond_angle((int(o[0]) >> 3) - plan.cx, (int(o[1]) >> 3) - plan.cy))
        lo, hi = plan.bounds[r], (plan.bounds[r + 1] if r + 1 < 4 else 4.0)
        assert lo <= a < hi or (r == 3 and a < plan.bounds[0])


@pytest.mark.gpu
@pytest.mark.parametrize("exchange", ["nccl", "p2p"])
@pytest.mark.parametrize("mode", ["own_cloud", "split"])
def test_sharded_map_world2_gpu(mode, exchange, tmp_path):
    """Union of the per-rank map shards == oracle map, with the NCCL all-to-all exchange and with the fused
    peer-memory (CUDA IPC / NVLink) exchange."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    _run(2, "nccl", mode, tmp_path, extra=("--points", "20000", "--exchange", exchange))


@pytest.mark.gpu
@pytest.mark.parametrize("exchange", ["nccl", "p2p"])
def test_sector_ownership_world2_gpu(exchange, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    _run(2, "nccl", "split", tmp_path, extra=("--points", "20000", "--exchange", exchange, "--plan", "sectors"))
