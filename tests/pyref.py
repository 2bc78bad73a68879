"""Second, independent restatement of the reference path in pure Python (dict of voxels, no tree).

Test-only. Used to cross-check the C++ oracle (oracle/vdbm_oracle.cpp) on small cases, so that a
bug in the oracle's tree/accessor/iteration code cannot hide behind "GPU == oracle". Follows
/root/reference/include/vdb_mapping/VDBMapping.hpp:466-566,612-631,731-792 and
OccupancyVDBMapping.hpp:59-117; OpenVDB DDA/MinIndex semantics per SURVEY.md Appendix A.
Python floats are IEEE binary64 with no FMA contraction; float32 arithmetic uses numpy scalars.
"""
import math

import numpy as np

DBL_MAX = float(np.finfo(np.float64).max)
MIN_INDEX = [2, 1, 9, 1, 2, 9, 0, 0]


def world_to_index(c, res):
    inv = 1.0 / res
    out = []
    for v in c:
        v = float(v)
        if math.fmod(v, res) != 0:
            v = v + res / 2.0
        out.append(int(math.floor(v * inv)))
    return tuple(out)


def dda_voxels(o, e):
    """Voxels marked by castRayIntoGrid (VDBMapping.hpp:550-566)."""
    if o == e:
        return []
    voxel = list(o)
    nxt, delta, step = [0.0] * 3, [0.0] * 3, [0] * 3
    for a in range(3):
        d = float(e[a]) - float(o[a])
        pos = float(o[a]) + 0.5
        if d == 0.0:
            step[a], nxt[a], delta[a] = 0, DBL_MAX, DBL_MAX
        else:
            inv = 1.0 / d
            if inv > 0:
                step[a] = 1
                nxt[a] = 0.0 + (float(voxel[a] + 1) - pos) * inv
            else:
                step[a] = -1
                nxt[a] = 0.0 + (float(voxel[a]) - pos) * inv
            delta[a] = float(step[a]) * inv
    out = []
    while True:
        out.append(tuple(voxel))
        key = (int(nxt[0] < nxt[1]) << 2) + (int(nxt[0] < nxt[2]) << 1) + int(nxt[1] < nxt[2])
        ax = MIN_INDEX[key]
        t = nxt[ax]
        nxt[ax] += delta[ax]
        voxel[ax] += step[ax]
        if not (t <= 1.0):
            break
    return out


def raycast(points, origin, res, rng):
    """Returns dict voxel -> hit flag (the update grid: key present = active, value = bool)."""
    upd = {}
    origin = [float(x) for x in origin]
    if any(math.isnan(x) for x in origin):
        return upd
    o_idx = world_to_index(origin, res)
    for p in points:
        end = [float(np.float32(x)) for x in p[:3]]
        if any(math.isnan(x) for x in end):
            continue
        clipped = False
        if rng > 0.0:
            d = [end[a] - origin[a] for a in range(3)]
            ln = math.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])
            if ln > rng:
                end = [origin[a] + (d[a] / ln) * rng for a in range(3)]
                clipped = True
        e_idx = world_to_index(end, res)
        for v in dda_voxels(o_idx, e_idx):
            upd.setdefault(v, False)
        if not clipped:
            upd[e_idx] = True
    return upd


def logodds(p):
    return np.float32(math.log(p) - math.log(1 - p))


class PyMap:
    def __init__(self, res, max_range, prob_hit, prob_miss, thres_min, thres_max, quirk=True):
        self.res, self.max_range = res, max_range
        self.hit, self.miss = logodds(prob_hit), logodds(prob_miss)
        self.tmin, self.tmax = logodds(thres_min), logodds(thres_max)
        self.maxlo = np.float32(math.log(0.99) - math.log(0.01))
        self.minlo = np.float32(math.log(0.01) - math.log(0.99))
        self.vox = {}        # voxel -> (float32 value, active)
        self.leaves = set()  # existing map leaves (origin >> 3)
        self.quirk = quirk

    def _op(self, v, a, is_hit):
        if is_hit:
            v = np.float32(v + self.hit)
            if v > self.tmax:
                a = True
                if v > self.maxlo:
                    v = self.maxlo
        else:
            v = np.float32(v + self.miss)
            if v < self.tmin:
                a = False
                if v < self.minlo:
                    v = self.minlo
        return v, a

    def update(self, upd):
        """updateMap; returns change dict voxel -> value flag."""
        change = {}
        # reference iteration order only matters per leaf: ascending offset (x&7,y&7,z&7) x-major
        by_leaf = {}
        for v in upd:
            by_leaf.setdefault((v[0] >> 3, v[1] >> 3, v[2] >> 3), []).append(v)
        for lk, voxels in by_leaf.items():
            voxels.sort(key=lambda v: ((v[0] & 7) << 6) | ((v[1] & 7) << 3) | (v[2] & 7))
            for v in voxels:
                is_hit = upd[v]
                changed = False
                if lk not in self.leaves:
                    # tile probe with inverted state (inactive background tile -> probe state True)
                    pv, pa = self._op(np.float32(0.0), True, is_hit)
                    if pa is not True:
                        changed = True
                    create = (pa is not False) or (pv != np.float32(0.0))
                    if create:
                        self.leaves.add(lk)
                    else:
                        if changed and self.quirk:
                            change[v] = is_hit
                        continue
                val, act = self.vox.get(v, (np.float32(0.0), False))
                nv, na = self._op(val, act, is_hit)
                if na != act:
                    changed = True
                self.vox[v] = (nv, na)
                if not self.quirk:
                    changed = (na != act)
                if changed:
                    change[v] = is_hit
        return change

    def insert(self, points, origin, rng=None):
        rng = self.max_range if rng is None else rng
        if not rng > 0:
            return {}, {}
        upd = raycast(points, origin, self.res, rng)
        return upd, self.update(upd)


def leafset_to_voxels(ls, with_values=False):
    """LeafSet -> dict voxel -> (active, value/flag) for every voxel that is active or non-background."""
    out = {}
    for i in range(len(ls)):
        ox, oy, oz = (int(x) for x in ls.origins[i])
        for w in range(8):
            aw = int(ls.active[i, w])
            vw = int(ls.valmask[i, w]) if ls.valmask is not None else 0
            for b in range(64):
                n = w * 64 + b
                act = (aw >> b) & 1
                if ls.values is not None:
                    val = ls.values[i, n]
                    if act or val != 0:
                        out[(ox + w, oy + (b >> 3), oz + (b & 7))] = (bool(act), np.float32(val))
                else:
                    flag = (vw >> b) & 1
                    if act or flag:
                        out[(ox + w, oy + (b >> 3), oz + (b & 7))] = (bool(act), bool(flag))
    return out
