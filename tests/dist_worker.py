"""Worker for the multi-rank tests (launched with torch.distributed.run).

  backend gloo (CPU, this container): the engine is a TEST DOUBLE built on the CPU oracle; what is under test is
      the host-side exchange logic of vdb_mapping_b200/dist.py (owner function, split sizes, all-to-all, import).
  backend nccl (GPU box, >= 2 GPUs): the engine is the product CudaEngine; the union of the per-rank map shards must
      be bit-identical to the single-process oracle map.
Rank 0 writes "OK" or the failure to the file given by --out.
"""
import argparse
import os
import sys
import traceback

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from vdb_mapping_b200 import dist as vdist, scans  # noqa: E402
from oracle.oracle import OracleOccupancyVDBMapping  # noqa: E402


class OracleEngine:
    """Test double with the CudaEngine interface, backed by the CPU oracle (CPU tests only)."""

    def __init__(self, m, src, plan=None):
        self.m, self.src, self.plan = m, src, plan

    def accumulate(self, points, origin):
        return self.m.accumulateUpdate(points, origin, self.src)

    def partition(self, world):
        ls = self.m.exportUpdateGrid(self.src)
        self.m.clearUpdateGrid(self.src)
        plan = self.plan or vdist.ShardPlan(0, world)
        owners = np.array([vdist.leaf_owner_planned_py(o, plan) for o in ls.origins], dtype=np.int64)
        order = np.argsort(owners, kind="stable")
        counts = np.bincount(owners, minlength=world).astype(np.int64)
        rec = np.zeros((len(ls), vdist.RECORD_WORDS), dtype=np.uint64)
        if len(ls):
            rec[:, 0] = [vdist.pack_leaf_key(o) for o in ls.origins]
            rec[:, 1:9] = ls.active
            rec[:, 9:17] = ls.valmask
        return counts, torch.from_numpy(rec[order].view(np.int64).reshape(-1).copy())

    def new_recv(self, n):
        return torch.empty(n * vdist.RECORD_WORDS, dtype=torch.int64)

    def import_records(self, buf, n):
        if not n:
            return
        rec = buf.numpy().view(np.uint64).reshape(n, vdist.RECORD_WORDS)
        keys = rec[:, 0]
        mask21 = np.uint64((1 << 21) - 1)
        lx = ((keys >> np.uint64(42)) & mask21).astype(np.int64) - (1 << 20)
        ly = ((keys >> np.uint64(21)) & mask21).astype(np.int64) - (1 << 20)
        lz = (keys & mask21).astype(np.int64) - (1 << 20)
        origins = (np.stack([lx, ly, lz], axis=1) * 8).astype(np.int32)
        self.m.importUpdate(self.src, origins, rec[:, 1:9].copy(), rec[:, 9:17].copy())

    def integrate(self):
        self.m.integrateUpdate()

    def counts_tensor(self, counts):
        return torch.as_tensor(np.asarray(counts, dtype=np.int64))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="gloo")
    ap.add_argument("--out", required=True)
    ap.add_argument("--mode", default="own_cloud")
    ap.add_argument("--scans", type=int, default=3)
    ap.add_argument("--points", type=int, default=3000)
    ap.add_argument("--exchange", default="nccl", choices=["nccl", "p2p"])
    ap.add_argument("--plan", default="hash", choices=["hash", "sectors"],
                    help="map ownership: the leaf hash, or azimuth sectors planned from the first scan (rays split by the same sectors)")
    args = ap.parse_args()
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    if args.backend == "nccl":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    else:
        dist.init_process_group("gloo")
    status = "OK"
    try:
        res, rng, cfg = 0.1, 4.0, (0.9, 0.1, 0.49, 0.51)
        if args.backend == "nccl":
            from vdb_mapping_b200.mapping import OccupancyVDBMapping
            m = OccupancyVDBMapping(res, device=torch.cuda.current_device())
        else:
            m = OracleOccupancyVDBMapping(res)
        m.setConfig(rng, *cfg)
        m.addInputSource("s", rng)
        plan = vdist.ShardPlan(0, world)
        if args.plan == "sectors":
            first_pts, first_origin = scans.small_scan(700, n=args.points, scale=2.5)
            plan = vdist.plan_sectors(first_pts, first_origin, res, rng, world)
            if args.backend == "nccl":
                m.setShardPlan(plan)
        eng = vdist.CudaEngine(m, "s") if args.backend == "nccl" else OracleEngine(m, "s", plan)
        p2p = args.backend == "nccl" and args.exchange == "p2p"
        if p2p:
            vdist.connect_peers(m, dist, capacity_records_per_sender=1 << 16)

        def insert(points, origin, mode):
            if not p2p:
                return vdist.sharded_insert(eng, points, origin, world, dist, mode=mode)
            if mode == "split":
                if args.plan == "sectors":
                    points = points[vdist.sector_rays(points, origin, plan, rank)]
                else:
                    lo, hi = vdist.split_points(points.shape[0], rank, world)
                    points = points[lo:hi]
            eng.accumulate(points, origin)
            vdist.push_pull_and_integrate(eng)
        ref = OracleOccupancyVDBMapping(res) if rank == 0 else None
        if ref:
            ref.setConfig(rng, *cfg)
            ref.addInputSource("s", rng)
        for k in range(args.scans):
            if args.mode == "own_cloud":
                clouds = [scans.small_scan(500 + 10 * k + r, n=args.points, scale=2.5) for r in range(world)]
                origin = clouds[0][1]
                pts = clouds[rank][0]  # every sensor's cloud is already in map coordinates; all share clouds[0]'s origin
                insert(pts, origin, "own_cloud")
                if ref:
                    for r in range(world):
                        ref.accumulateUpdate(clouds[r][0], origin, "s")
                    ref.integrateUpdate()
            else:
                pts, origin = scans.small_scan(700 + k, n=args.points, scale=2.5)
                if args.plan == "sectors" and not p2p:
                    # sector split of the rays on the generic (all-to-all) path: done here, the exchange then sees "own clouds"
                    sel = vdist.sector_rays(pts, origin, plan, rank)
                    eng.accumulate(pts[sel], origin)
                    vdist.exchange_and_integrate(eng, world, dist)
                else:
                    insert(pts, origin, "split")
                if ref:
                    ref.insertPointCloud(pts, origin, "s")
        # gather the shards on rank 0 and compare the union with the single-process oracle map
        shard = m.exportMap()
        owners_ok = all(vdist.leaf_owner_planned_py(o, plan) == rank for o in shard.origins)
        if args.backend == "nccl":
            owners_ok = owners_ok and all(m.leafOwnerPlanned(o, world) == rank for o in shard.origins[:2000])
            chk = m.mapChecksum()
            sums = [None] * world
            dist.all_gather_object(sums, chk)
        payload = [shard.origins, shard.active, shard.values, owners_ok]
        gathered = [None] * world
        dist.all_gather_object(gathered, payload)
        if rank == 0:
            assert all(g[3] for g in gathered), "a rank holds leaves it does not own"
            origins = np.concatenate([g[0] for g in gathered])
            active = np.concatenate([g[1] for g in gathered])
            values = np.concatenate([g[2] for g in gathered])
            keys = np.array([vdist.pack_leaf_key(o) for o in origins], dtype=np.uint64)
            assert len(np.unique(keys)) == len(keys), "a leaf lives on two ranks"
            order = np.argsort(keys)
            want = ref.exportMap()
            assert np.array_equal(origins[order], want.origins), "leaf sets differ"
            assert np.array_equal(active[order], want.active), "active masks differ"
            assert np.array_equal(values[order].view(np.uint32), want.values.view(np.uint32)), "values differ"
            sizes = [len(g[0]) for g in gathered]
            if args.backend == "nccl":
                # the checksum witness bench.py relies on: sum over the shards == one map holding the union (rebuilt here
                # from the oracle's leaves through applyMapSectionGrid)
                one = OccupancyVDBMapping(res, device=torch.cuda.current_device())
                one.setConfig(rng, *cfg)
                one.applyMapSectionGrid(want, tile_quirk=False)
                assert (sum(x[0] for x in sums) & ((1 << 64) - 1), sum(x[1] for x in sums)) == one.mapChecksum(), "checksum witness differs"
                one.close()
            status = f"OK shards={sizes} total={len(keys)}"
    except Exception:
        status = "FAIL\n" + traceback.format_exc()
    if rank == 0:
        open(args.out, "w").write(status)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
