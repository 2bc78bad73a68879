// vdbm_kernels.cu — hand-written sm_100a kernels of the scan-integration hot path.
//
//   K0 prep_rays_kernel      raycastPointCloud per-point set-up   (VDBMapping.hpp:499-519, worldToIndex :612-631)
//   K1 raycast_dda_kernel    castRayIntoGrid 3D-DDA + endpoint    (VDBMapping.hpp:550-566, :533-536)
//   K2 apply_update_kernel   updateMap + Occupancy node ops       (VDBMapping.hpp:731-792, OccupancyVDBMapping.hpp:92-117)
//   K3 section_kernel        getMapSection / extract*Leaf         (VDBMapping.hpp:921-1011)
//   K4 gather_* kernels      getGrid()/update-grid materialisation
//
// Compiled with -fmad=false; every fp64 operation that decides a voxel path additionally uses an explicit
// round-to-nearest intrinsic (__dadd_rn/__dmul_rn/__ddiv_rn/__dsqrt_rn) so no contraction or re-association
// can change the sequence of roundings the reference executes on an x86-64 (no-FMA) build.
#include "vdbm_device.cuh"

#include <atomic>
#include <cfloat>
#include <cub/device/device_radix_sort.cuh>

namespace vdbm {

namespace {

constexpr unsigned kFull = 0xFFFFFFFFu;

__device__ __forceinline__ uint64_t ldcg64(const uint64_t* p)
{
  uint64_t v;
  asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void redOr64(uint64_t* p, uint64_t v)
{
  asm volatile("red.global.or.b64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src)
{
  uint32_t lo = __shfl_sync(kFull, uint32_t(v), src);
  uint32_t hi = __shfl_sync(kFull, uint32_t(v >> 32), src);
  return (uint64_t(hi) << 32) | lo;
}
__device__ __forceinline__ uint64_t shflXor64(uint64_t v, int m)
{
  uint32_t lo = __shfl_xor_sync(kFull, uint32_t(v), m);
  uint32_t hi = __shfl_xor_sync(kFull, uint32_t(v >> 32), m);
  return (uint64_t(hi) << 32) | lo;
}
// 256-bit global load/store (Blackwell sm_100+: LDG.E.ENL2.256 / STG.E.ENL2.256), one 32-byte sector per lane
__device__ __forceinline__ void ld256(const float* p, float (&v)[8])
{
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}
__device__ __forceinline__ void st256(float* p, const float (&v)[8])
{
  asm volatile("st.global.v8.f32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};" ::"f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]),
               "f"(v[5]), "f"(v[6]), "f"(v[7]), "l"(p)
               : "memory");
}

// ---- update-grid hash: find or insert a leaf slot (slot == storage; masks are zero for empty slots) ----
__device__ __forceinline__ uint32_t updFindOrInsert(const UpdateTable& t, uint64_t key, Counters* ctr)
{
  uint32_t h = uint32_t(mix64(key)) & t.cap_mask;
  const uint32_t max_probe = min(t.cap_mask, 4096u);
  for (uint32_t probe = 0; probe <= max_probe; ++probe)
  {
    uint64_t k = ldcg64(t.keys + h);
    if (k == key) return h;
    if (k == kEmptyKey)
    {
      unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long*>(t.keys + h), kEmptyKey, key);
      if (old == kEmptyKey)
      {
        uint32_t idx   = atomicAdd(t.n_touched, 1u);
        t.touched[idx] = h; // idx < cap always: at most cap successful claims
        return h;
      }
      if (old == key) return h;
    }
    h = (h + 1) & t.cap_mask;
  }
  atomicOr(&ctr->flags, kFlagUpdateOverflow);
  return kInvalid;
}

} // namespace

// ====================================================================================================
// K0: per-point set-up. One thread per point; fully convergent, so the expensive fp64 div/sqrt/fmod run at
// full SIMD efficiency here instead of inside the divergent DDA kernel.
// ====================================================================================================
__device__ __forceinline__ int32_t worldToIndex1(double c, double res, double half_res, double inv_res)
{
  // VDBMapping.hpp:612-631: +res/2 iff fmod(c,res) != 0; Transform::worldToIndex = multiply by 1/res; Coord::floor
  if (fmod(c, res) != 0.0) c = __dadd_rn(c, half_res);
  return int32_t(floor(__dmul_rn(c, inv_res)));
}

__global__ void __launch_bounds__(256) prep_rays_kernel(RaycastArgs a, Counters* ctr)
{
  const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  unsigned long long visits = 0;
  unsigned nan_skipped = 0, clipped = 0, range_err = 0;
  if (i < a.n)
  {
    const float* p = reinterpret_cast<const float*>(a.points + i * a.stride);
    double e[3]    = {double(p[0]), double(p[1]), double(p[2])}; // VDBMapping.hpp:501
    RayRec r;
    r.flags = 0;
    r.end[0] = r.end[1] = r.end[2] = 0;
    r.delta[0] = r.delta[1] = r.delta[2] = DBL_MAX;
    // VDBMapping.hpp:505-510 skips NaN; +-inf is undefined behaviour in the reference and is dropped here too
    const bool finite = isfinite(e[0]) && isfinite(e[1]) && isfinite(e[2]);
    if (!finite) nan_skipped = 1;
    else
    {
      bool max_range_ray = false;
      if (a.range > 0.0)
      {
        const double dx = __dsub_rn(e[0], a.origin[0]), dy = __dsub_rn(e[1], a.origin[1]), dz = __dsub_rn(e[2], a.origin[2]);
        // openvdb Vec3::length(): sqrt(x*x + y*y + z*z), left-associated
        const double len = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
        if (len > a.range) // VDBMapping.hpp:512-517: origin + (d.unit() * range), unit() = d / len
        {
          e[0]          = __dadd_rn(a.origin[0], __dmul_rn(__ddiv_rn(dx, len), a.range));
          e[1]          = __dadd_rn(a.origin[1], __dmul_rn(__ddiv_rn(dy, len), a.range));
          e[2]          = __dadd_rn(a.origin[2], __dmul_rn(__ddiv_rn(dz, len), a.range));
          max_range_ray = true;
          clipped       = 1;
        }
      }
      bool in_range = true;
      long long l1  = 0;
      bool zero     = true;
#pragma unroll
      for (int k = 0; k < 3; ++k)
      {
        const double fl = floor(__dmul_rn((fmod(e[k], a.resolution) != 0.0) ? __dadd_rn(e[k], a.half_res) : e[k], a.inv_res));
        if (!(fabs(fl) < double(kVoxelLimit))) in_range = false;
        const int32_t ei = in_range ? int32_t(fl) : 0;
        r.end[k]         = ei;
        const double dir = __dsub_rn(double(ei), double(a.origin_idx[k])); // exact
        if (dir != 0.0)
        {
          zero       = false;
          r.delta[k] = fabs(__ddiv_rn(1.0, dir)); // Ray::mInvDir = 1/dir; DDA::mDelta = step * inv = |inv|
          l1 += (long long)fabs(dir);
        }
      }
      if (!in_range) range_err = 1;
      else
      {
        r.flags = kRayValid | (max_range_ray ? kRayClipped : 0u) | (zero ? kRayZeroLen : 0u);
        visits  = zero ? 0ull : (unsigned long long)(1 + l1); // castRayIntoGrid marks 1 + |dx|+|dy|+|dz| voxels
      }
    }
    a.rays[i] = r;
  }
  // warp-aggregated statistics
  visits      = __reduce_add_sync(kFull, unsigned(visits)); // per-lane visits <= 1 + 3*2^24, 32 lanes fit in 32 bits
  nan_skipped = __reduce_add_sync(kFull, nan_skipped);
  clipped     = __reduce_add_sync(kFull, clipped);
  range_err   = __reduce_add_sync(kFull, range_err);
  if ((threadIdx.x & 31) == 0)
  {
    if (visits) atomicAdd(&ctr->visits, visits);
    if (nan_skipped) atomicAdd(&ctr->nan_skipped, (unsigned long long)nan_skipped);
    if (clipped) atomicAdd(&ctr->clipped, (unsigned long long)clipped);
    if (range_err) atomicOr(&ctr->flags, kFlagCoordRange);
  }
}

// ====================================================================================================
// K1: 3D-DDA. Persistent warps; every lane owns one ray at a time and refills itself from a global cursor
// the moment its ray ends, so ray-length variance (10..2000+ visits) costs no idle lanes.
// Voxel stepping replays openvdb::math::DDA<Ray<double>,0> bit for bit:
//   next[a] = 0.5*|1/dir[a]| (exact), then repeated  next[axis] += delta[axis]  in fp64,
//   axis = MinIndex(next) with its tie table {2,1,9,1,2,9,0,0}, continue while t <= 1.0.
// Marking: bits of consecutive visits that fall in the same (leaf, x-slice) 64-bit mask word are merged in a
// register and flushed with ONE red.global.or.b64; x is monotonic along a ray, so every (leaf, word) pair is
// flushed at most once per ray. Leaf slots come from the update hash on leaf change only.
// ====================================================================================================
__global__ void __launch_bounds__(256) raycast_dda_kernel(RaycastArgs a, UpdateTable ut, Counters* ctr)
{
  const int lane = threadIdx.x & 31;
  const int ox = a.origin_idx[0], oy = a.origin_idx[1], oz = a.origin_idx[2];

  bool busy = false, done = false;
  double n0 = 0, n1 = 0, n2 = 0, d0 = 0, d1 = 0, d2 = 0;
  int x = 0, y = 0, z = 0, sx = 0, sy = 0, sz = 0;
  int ex = 0, ey = 0, ez = 0;
  uint32_t rflags = 0;
  int cur_x = INT_MIN, cur_ly = 0, cur_lz = 0;
  uint32_t cur_slot = kInvalid;
  uint64_t acc      = 0;

  for (;;)
  {
    // ---- refill idle lanes ----
    const unsigned need = __ballot_sync(kFull, !busy && !done);
    if (need)
    {
      unsigned base = 0;
      const int leader = __ffs(need) - 1;
      if (lane == leader) base = atomicAdd(&ctr->ray_cursor, (unsigned)__popc(need));
      base = __shfl_sync(kFull, base, leader);
      if (!busy && !done)
      {
        const uint64_t idx = uint64_t(base) + __popc(need & ((1u << lane) - 1u));
        if (idx >= a.n) done = true;
        else
        {
          const RayRec r = a.rays[idx];
          if (r.flags & kRayValid)
          {
            ex = r.end[0]; ey = r.end[1]; ez = r.end[2];
            rflags = r.flags;
            if (r.flags & kRayZeroLen)
            {
              // no DDA (VDBMapping.hpp:559); a non-clipped endpoint is still set on with value true (:533-536)
              if (!(r.flags & kRayClipped))
              {
                const uint32_t slot = updFindOrInsert(ut, packLeafKey(ex >> 3, ey >> 3, ez >> 3), ctr);
                if (slot != kInvalid)
                {
                  const uint64_t bit = uint64_t(1) << (((ey & 7) << 3) | (ez & 7));
                  redOr64(ut.active + size_t(slot) * 8 + (ex & 7), bit);
                  redOr64(ut.value + size_t(slot) * 8 + (ex & 7), bit);
                }
              }
            }
            else
            {
              d0 = r.delta[0]; d1 = r.delta[1]; d2 = r.delta[2];
              // DDA::init: next = t0 + (voxel + {1|0} - pos) * inv = 0.5 * |inv| exactly; DBL_MAX if dir == 0
              n0 = (d0 == DBL_MAX) ? DBL_MAX : __dmul_rn(0.5, d0);
              n1 = (d1 == DBL_MAX) ? DBL_MAX : __dmul_rn(0.5, d1);
              n2 = (d2 == DBL_MAX) ? DBL_MAX : __dmul_rn(0.5, d2);
              sx = (ex > ox) - (ex < ox); sy = (ey > oy) - (ey < oy); sz = (ez > oz) - (ez < oz);
              x = ox; y = oy; z = oz;
              cur_x = INT_MIN; cur_slot = kInvalid; acc = 0;
              busy = true;
            }
          }
        }
      }
    }
    if (__all_sync(kFull, done && !busy)) break;

    if (busy)
    {
      // ---- mark current voxel (setActiveState(dda.voxel(), true), VDBMapping.hpp:563) ----
      const int ly = y >> 3, lz = z >> 3;
      if (x != cur_x || ly != cur_ly || lz != cur_lz)
      {
        if (acc != 0 && cur_slot != kInvalid) redOr64(ut.active + size_t(cur_slot) * 8 + (cur_x & 7), acc);
        const bool same_leaf = (cur_x != INT_MIN) && ((x >> 3) == (cur_x >> 3)) && ly == cur_ly && lz == cur_lz;
        if (!same_leaf) cur_slot = updFindOrInsert(ut, packLeafKey(x >> 3, ly, lz), ctr);
        cur_x = x; cur_ly = ly; cur_lz = lz;
        acc = 0;
      }
      acc |= uint64_t(1) << (((y & 7) << 3) | (z & 7));

      // ---- DDA::step(): MinIndex, t = next[axis], next[axis] += delta[axis], voxel[axis] += step[axis] ----
      // MinIndex table {2,1,9,1,2,9,0,0} on key ((n0<n1)<<2)+((n0<n2)<<1)+(n1<n2):
      //   (n0<n1 && n0<n2) -> 0 ; else (n1<n2) -> 1 ; else 2   (keys 2 and 5 are unreachable)
      const bool c01 = n0 < n1, c02 = n0 < n2, c12 = n1 < n2;
      double t;
      if (c01 && c02) { t = n0; n0 = __dadd_rn(n0, d0); x += sx; }
      else if (c12)   { t = n1; n1 = __dadd_rn(n1, d1); y += sy; }
      else            { t = n2; n2 = __dadd_rn(n2, d2); z += sz; }
      if (!(t <= 1.0))
      {
        // ray finished: the voxel just stepped to is NOT marked. Flush, then the endpoint hit (:533-536).
        if (cur_slot != kInvalid)
        {
          redOr64(ut.active + size_t(cur_slot) * 8 + (cur_x & 7), acc);
          if (!(rflags & kRayClipped))
            redOr64(ut.value + size_t(cur_slot) * 8 + (ex & 7), uint64_t(1) << (((ey & 7) << 3) | (ez & 7)));
        }
        busy = false;
      }
    }
  }
}

// ====================================================================================================
// K2: updateMap. One warp per touched update leaf; lane L owns the 16 consecutive voxels [16L, 16L+16)
// (= 64 contiguous bytes of leaf values, moved with two 256-bit loads/stores). Streaming RMW over the map leaf.
// ====================================================================================================
__global__ void __launch_bounds__(256) apply_update_kernel(UpdateTable ut, MapTable mt, LogOdds lo, LeafRecord* change_out,
                                                          uint32_t change_cap, Counters* ctr)
{
  const int lane         = threadIdx.x & 31;
  const uint32_t warp    = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t n       = *ut.n_touched;
  unsigned long long upd_total = 0, chg_total = 0, new_total = 0;

  for (uint32_t i = warp; i < n; i += n_warps)
  {
    const uint32_t slot = ut.touched[i];
    const uint64_t key  = ut.keys[slot];
    uint64_t A = 0, V = 0;
    if (lane < 8)
    {
      A = ut.active[size_t(slot) * 8 + lane];
      V = ut.value[size_t(slot) * 8 + lane];
      // consume: leave the slot clean for the next accumulation period (VDBMapping.hpp:384)
      ut.active[size_t(slot) * 8 + lane] = 0;
      ut.value[size_t(slot) * 8 + lane]  = 0;
    }
    if (lane == 0) ut.keys[slot] = kEmptyKey;

    const unsigned nz_words = __ballot_sync(kFull, A != 0) & 0xFFu;
    if (nz_words == 0) continue; // nothing active (can only happen for imported empty records)
    const unsigned hit_words = __ballot_sync(kFull, V != 0) & 0xFFu;

    // ---- find or create the map leaf (lane 0 probes; keys are unique per launch, so no same-key races) ----
    // OpenVDB tile probe: on a missing leaf a miss whose probe result is (0.0f, inactive) does not create it.
    const bool create_ok = !lo.miss_probe_no_create || hit_words != 0;
    uint32_t leaf = kInvalid;
    int is_new    = 0;
    if (lane == 0)
    {
      uint32_t h = uint32_t(mix64(key)) & mt.hcap_mask;
      for (uint32_t probe = 0; probe <= mt.hcap_mask; ++probe)
      {
        const uint64_t k = ldcg64(mt.hkeys + h);
        if (k == key) { leaf = mt.hvals[h]; break; }
        if (k == kEmptyKey)
        {
          if (!create_ok) break;
          unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long*>(mt.hkeys + h), kEmptyKey, key);
          if (old == kEmptyKey)
          {
            const uint32_t li = atomicAdd(mt.n_leaves, 1u);
            if (li >= mt.pool_cap) { atomicOr(&ctr->flags, kFlagMapOverflow); break; }
            mt.hvals[h]      = li;
            mt.leaf_keys[li] = key;
            leaf             = li;
            is_new           = 1;
            break;
          }
        }
        h = (h + 1) & mt.hcap_mask;
      }
      if (leaf != kInvalid && atomicExch(mt.leaf_dirty + leaf, 1u) == 0u) mt.dirty_list[atomicAdd(mt.n_dirty, 1u)] = leaf;
    }
    leaf   = __shfl_sync(kFull, leaf, 0);
    is_new = __shfl_sync(kFull, is_new, 0);

    const int w  = lane >> 2;        // mask word of this lane's 16 voxels
    const int sh = (lane & 3) << 4;  // bit offset inside the word
    const uint32_t ua = uint32_t(shfl64(A, w) >> sh) & 0xFFFFu; // active update bits
    const uint32_t uv = uint32_t(shfl64(V, w) >> sh) & 0xFFFFu; // hit bits

    // lowest-offset active update voxel of the leaf (first one the reference visits): offset n_first
    const int w_first      = __ffs(nz_words) - 1;
    const uint64_t a_first = shfl64(A, w_first), v_first = shfl64(V, w_first);
    const int b_first      = __ffsll((long long)a_first) - 1;
    const int n_first      = (w_first << 6) | b_first;
    const bool first_is_hit = (v_first >> b_first) & 1;

    uint32_t ca = 0, cv = 0; // change-grid bits of this lane's 16 voxels
    if (leaf == kInvalid)
    {
      // no leaf and none may be created: every (miss) voxel only runs the tile probe. With the quirk each of
      // them is reported when the probe flipped the inverted state (VDBMapping.hpp:743-750 via the lambda).
      if (lo.replicate_quirk && lo.miss_probe_flips) ca = ua & ~uv;
    }
    else
    {
      uint64_t M = 0;
      if (lane < 8 && !is_new) M = mt.leaf_mask[size_t(leaf) * 8 + lane];
      const uint32_t oa = uint32_t(shfl64(M, w) >> sh) & 0xFFFFu;
      float* vp = mt.leaf_vals + size_t(leaf) * 512 + lane * 16;
      float v0[8], v1[8];
      const bool touch = (ua != 0);
      if (!is_new && touch) { ld256(vp, v0); ld256(vp + 8, v1); }
      else
      {
#pragma unroll
        for (int j = 0; j < 8; ++j) { v0[j] = 0.0f; v1[j] = 0.0f; }
      }
      uint32_t na = oa;
#pragma unroll
      for (int j = 0; j < 16; ++j)
      {
        const bool a_ = (ua >> j) & 1, h_ = (uv >> j) & 1, o_ = (oa >> j) & 1;
        float& ref = (j < 8) ? v0[j & 7] : v1[j & 7];
        // OccupancyVDBMapping.hpp:92-117 (clamping only inside the threshold branch)
        float nv  = __fadd_rn(ref, h_ ? lo.hit : lo.miss);
        bool act  = o_;
        if (h_) { if (nv > lo.thres_max) { act = true;  if (nv > lo.max_lo) nv = lo.max_lo; } }
        else    { if (nv < lo.thres_min) { act = false; if (nv < lo.min_lo) nv = lo.min_lo; } }
        if (a_)
        {
          ref = nv;
          na  = (na & ~(1u << j)) | (uint32_t(act) << j);
          const bool changed = (act != o_);
          ca |= uint32_t(changed) << j;
          cv |= uint32_t(changed && h_) << j;
        }
      }
      if (is_new || touch) { st256(vp, v0); st256(vp + 8, v1); }
      // assemble the new 64-bit active word from the 4 lanes that share it
      uint64_t piece = uint64_t(na) << sh;
      piece |= shflXor64(piece, 1);
      piece |= shflXor64(piece, 2);
      if ((lane & 3) == 0) mt.leaf_mask[size_t(leaf) * 8 + w] = piece;

      // tile-probe quirk: first visited voxel of a leaf that did not exist, if it is a miss whose
      // probe flips the inverted tile state, is reported as changed although its flag did not change.
      if (is_new && lo.replicate_quirk && lo.miss_probe_flips)
      {
        if (!lo.miss_probe_no_create)
        {
          if (!first_is_hit && (n_first >> 4) == lane) ca |= 1u << (n_first & 15);
        }
        else
        {
          // degenerate config: misses do not create the leaf, so every miss BEFORE the first hit was probed
          const int wh       = __ffs(hit_words) - 1;
          const uint64_t vh  = shfl64(V, wh);
          const int n_hit    = (wh << 6) | (__ffsll((long long)vh) - 1);
          const int lo_n     = lane << 4;
          uint32_t before    = 0;
          if (n_hit >= lo_n + 16) before = 0xFFFFu;
          else if (n_hit > lo_n) before = (1u << (n_hit - lo_n)) - 1u;
          ca |= ua & ~uv & before;
        }
      }
    }

    upd_total += __popc(ua);
    chg_total += __popc(ca);
    new_total += (lane == 0 && is_new) ? 1 : 0;

    if (change_out != nullptr)
    {
      uint64_t pa = uint64_t(ca) << sh, pv = uint64_t(cv) << sh;
      pa |= shflXor64(pa, 1); pa |= shflXor64(pa, 2);
      pv |= shflXor64(pv, 1); pv |= shflXor64(pv, 2);
      const unsigned any = __ballot_sync(kFull, ca != 0);
      if (any)
      {
        uint32_t ci = 0;
        if (lane == 0) ci = atomicAdd(&ctr->n_change, 1u);
        ci = __shfl_sync(kFull, ci, 0);
        if (ci < change_cap)
        {
          LeafRecord* r = change_out + ci;
          if (lane == 0) r->key = key;
          if ((lane & 3) == 0) { r->active[w] = pa; r->value[w] = pv; }
        }
      }
    }
  }
  // per-warp totals -> global counters
  upd_total = __reduce_add_sync(kFull, unsigned(upd_total));
  chg_total = __reduce_add_sync(kFull, unsigned(chg_total));
  new_total = __reduce_add_sync(kFull, unsigned(new_total));
  if (lane == 0)
  {
    if (upd_total) atomicAdd(&ctr->voxel_updates, upd_total);
    if (chg_total) atomicAdd(&ctr->state_changes, chg_total);
    if (new_total) atomicAdd(&ctr->new_leaves, new_total);
  }
}

// ====================================================================================================
// growth / import / export helpers
// ====================================================================================================
__global__ void rehash_update_kernel(UpdateTable old_t, uint32_t old_n, UpdateTable new_t, Counters* ctr)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= old_n) return;
  const uint32_t s  = old_t.touched[i];
  const uint32_t ns = updFindOrInsert(new_t, old_t.keys[s], ctr);
  if (ns == kInvalid) return;
#pragma unroll
  for (int w = 0; w < 8; ++w)
  {
    new_t.active[size_t(ns) * 8 + w] = old_t.active[size_t(s) * 8 + w];
    new_t.value[size_t(ns) * 8 + w]  = old_t.value[size_t(s) * 8 + w];
  }
}

__global__ void rehash_map_kernel(MapTable mt, uint32_t n_leaves, Counters* ctr)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_leaves) return;
  const uint64_t key = mt.leaf_keys[i];
  uint32_t h         = uint32_t(mix64(key)) & mt.hcap_mask;
  for (uint32_t probe = 0; probe <= mt.hcap_mask; ++probe)
  {
    unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long*>(mt.hkeys + h), kEmptyKey, key);
    if (old == kEmptyKey) { mt.hvals[h] = i; return; }
    h = (h + 1) & mt.hcap_mask;
  }
  atomicOr(&ctr->flags, kFlagMapOverflow);
}

__global__ void import_update_kernel(UpdateTable ut, const LeafRecord* recs, uint64_t n, Counters* ctr)
{
  // 16 lanes per record: lane j<8 ORs active[j], lane 8..15 ORs value[j-8]
  const uint64_t t   = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const uint64_t rec = t >> 4;
  const int j        = int(t & 15);
  const bool valid   = rec < n;
  uint64_t word = 0, key = 0;
  if (valid)
  {
    key  = recs[rec].key;
    word = (j < 8) ? recs[rec].active[j] : recs[rec].value[j - 8];
  }
  // the 16 lanes of a record agree on whether it has any active bit
  const unsigned grp   = 0xFFFFu << (threadIdx.x & 16);
  const unsigned nz    = __ballot_sync(kFull, valid && j < 8 && word != 0) & grp;
  uint32_t slot        = kInvalid;
  if (valid && nz && j == 0) slot = updFindOrInsert(ut, key, ctr);
  slot = __shfl_sync(kFull, slot, (threadIdx.x & 16));
  if (slot == kInvalid || word == 0) return;
  uint64_t* dst = (j < 8) ? ut.active + size_t(slot) * 8 + j : ut.value + size_t(slot) * 8 + (j - 8);
  redOr64(dst, word);
}

__global__ void clear_update_kernel(UpdateTable ut, uint32_t n)
{
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t i = t >> 4;
  const int j      = t & 15;
  if (i >= n) return;
  const uint32_t s = ut.touched[i];
  if (j < 8) ut.active[size_t(s) * 8 + j] = 0;
  else ut.value[size_t(s) * 8 + (j - 8)] = 0;
  if (j == 0) ut.keys[s] = kEmptyKey;
}

__global__ void keys_from_slots_kernel(const uint64_t* keys, const uint32_t* slots, uint32_t n, uint64_t* out_keys, uint32_t* out_idx)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t s = slots ? slots[i] : i;
  out_keys[i]      = keys[s];
  out_idx[i]       = s;
}

__global__ void iota_kernel(uint32_t* out, uint32_t n)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = i;
}

__global__ void unpack_origins_kernel(const uint64_t* keys, uint32_t n, int32_t* origins)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int32_t x, y, z;
  unpackLeafOrigin(keys[i], x, y, z);
  origins[3 * size_t(i) + 0] = x;
  origins[3 * size_t(i) + 1] = y;
  origins[3 * size_t(i) + 2] = z;
}

// update-grid leaves in `order` (slot indices) -> records
__global__ void gather_update_kernel(UpdateTable ut, uint32_t n, const uint32_t* order, LeafRecord* out)
{
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t i = t >> 4;
  const int j      = t & 15;
  if (i >= n) return;
  const uint32_t s = order[i];
  if (j < 8) out[i].active[j] = ut.active[size_t(s) * 8 + j];
  else out[i].value[j - 8] = ut.value[size_t(s) * 8 + (j - 8)];
  if (j == 0) out[i].key = ut.keys[s];
}

__global__ void split_records_kernel(const LeafRecord* recs, uint32_t n, int32_t* origins, uint64_t* active, uint64_t* value)
{
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t i = t >> 4;
  const int j      = t & 15;
  if (i >= n) return;
  if (j < 8) active[size_t(i) * 8 + j] = recs[i].active[j];
  else value[size_t(i) * 8 + (j - 8)] = recs[i].value[j - 8];
  if (j == 0)
  {
    int32_t x, y, z;
    unpackLeafOrigin(recs[i].key, x, y, z);
    origins[3 * size_t(i)] = x; origins[3 * size_t(i) + 1] = y; origins[3 * size_t(i) + 2] = z;
  }
}

// map leaves listed in leaf_idx -> SoA staging (one warp per leaf, 2 KB values with 256-bit accesses)
__global__ void __launch_bounds__(256) gather_map_kernel(MapTable mt, uint32_t n, const uint32_t* leaf_idx, int32_t* origins,
                                                        uint64_t* mask, float* vals, int clear_dirty)
{
  const int lane         = threadIdx.x & 31;
  const uint32_t warp    = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t i = warp; i < n; i += n_warps)
  {
    const uint32_t l = leaf_idx[i];
    float a[8], b[8];
    ld256(mt.leaf_vals + size_t(l) * 512 + lane * 16, a);
    ld256(mt.leaf_vals + size_t(l) * 512 + lane * 16 + 8, b);
    st256(vals + size_t(i) * 512 + lane * 16, a);
    st256(vals + size_t(i) * 512 + lane * 16 + 8, b);
    if (lane < 8) mask[size_t(i) * 8 + lane] = mt.leaf_mask[size_t(l) * 8 + lane];
    if (lane == 8)
    {
      int32_t x, y, z;
      unpackLeafOrigin(mt.leaf_keys[l], x, y, z);
      origins[3 * size_t(i)] = x; origins[3 * size_t(i) + 1] = y; origins[3 * size_t(i) + 2] = z;
      if (clear_dirty) mt.leaf_dirty[l] = 0;
    }
  }
}

// K3: getMapSection. One warp per map leaf; overlapping leaves with content inside the box are appended
// (unsorted; the host sorts the small result by key).
__global__ void __launch_bounds__(256) section_kernel(MapTable mt, uint32_t n_leaves, int bx0, int by0, int bz0, int bx1, int by1, int bz1,
                                                     int full, int result_float, uint64_t* out_keys, uint64_t* out_active,
                                                     uint64_t* out_valmask, float* out_vals, uint32_t out_cap, Counters* ctr)
{
  const int lane         = threadIdx.x & 31;
  const uint32_t warp    = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t l = warp; l < n_leaves; l += n_warps)
  {
    const uint64_t key = mt.leaf_keys[l];
    int32_t ox, oy, oz;
    unpackLeafOrigin(key, ox, oy, oz);
    // CoordBBox::hasOverlap of [origin, origin+7] with the inclusive box
    if (ox + 7 < bx0 || ox > bx1 || oy + 7 < by0 || oy > by1 || oz + 7 < bz0 || oz > bz1) continue;
    // inside-box bits of this lane's 16 voxels: x = ox + (lane>>2), y = oy + 2*(lane&3) + {0,1}, z = oz + 0..7
    const int xx = ox + (lane >> 2);
    uint32_t in16 = 0;
    if (xx >= bx0 && xx <= bx1)
    {
      uint32_t zb = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) zb |= uint32_t(oz + k >= bz0 && oz + k <= bz1) << k;
      const int yy = oy + 2 * (lane & 3);
      if (yy >= by0 && yy <= by1) in16 |= zb;
      if (yy + 1 >= by0 && yy + 1 <= by1) in16 |= zb << 8;
    }
    const int w  = lane >> 2;
    const int sh = (lane & 3) << 4;
    uint64_t M   = 0;
    if (lane < 8) M = mt.leaf_mask[size_t(l) * 8 + lane];
    const uint32_t on16 = uint32_t(shfl64(M, w) >> sh) & 0xFFFFu & in16;
    float a[8], b[8];
    uint32_t nz16 = 0; // value != 0 bits
    if (full)
    {
      ld256(mt.leaf_vals + size_t(l) * 512 + lane * 16, a);
      ld256(mt.leaf_vals + size_t(l) * 512 + lane * 16 + 8, b);
#pragma unroll
      for (int j = 0; j < 8; ++j)
      {
        if (!((in16 >> j) & 1)) a[j] = 0.0f;
        if (!((in16 >> (j + 8)) & 1)) b[j] = 0.0f;
        nz16 |= uint32_t(a[j] != 0.0f) << j;
        nz16 |= uint32_t(b[j] != 0.0f) << (j + 8);
      }
    }
    else
    {
      // extractSparseLeaf: setValueOn(coord, true) -> value true / 1.0f
#pragma unroll
      for (int j = 0; j < 8; ++j)
      {
        a[j] = ((on16 >> j) & 1) ? 1.0f : 0.0f;
        b[j] = ((on16 >> (j + 8)) & 1) ? 1.0f : 0.0f;
      }
      nz16 = on16;
    }
    // result leaf exists iff some voxel in the box is active or differs from the background
    const unsigned exists = __ballot_sync(kFull, (on16 | nz16) != 0);
    if (!exists) continue;
    uint32_t oi = 0;
    if (lane == 0) oi = atomicAdd(&ctr->n_out, 1u);
    oi = __shfl_sync(kFull, oi, 0);
    if (oi >= out_cap) continue;
    uint64_t pa = uint64_t(on16) << sh, pv = uint64_t(nz16) << sh;
    pa |= shflXor64(pa, 1); pa |= shflXor64(pa, 2);
    pv |= shflXor64(pv, 1); pv |= shflXor64(pv, 2);
    if (lane == 0) out_keys[oi] = key;
    if ((lane & 3) == 0)
    {
      out_active[size_t(oi) * 8 + w] = pa;
      if (!result_float) out_valmask[size_t(oi) * 8 + w] = pv;
    }
    if (result_float)
    {
      st256(out_vals + size_t(oi) * 512 + lane * 16, a);
      st256(out_vals + size_t(oi) * 512 + lane * 16 + 8, b);
    }
  }
}

__global__ void probe_kernel(MapTable mt, int32_t x, int32_t y, int32_t z, float* out_val, int32_t* out_active)
{
  const uint64_t key = packLeafKey(x >> 3, y >> 3, z >> 3);
  uint32_t h         = uint32_t(mix64(key)) & mt.hcap_mask;
  *out_val    = 0.0f;
  *out_active = 0;
  for (uint32_t probe = 0; probe <= mt.hcap_mask; ++probe)
  {
    const uint64_t k = mt.hkeys[h];
    if (k == kEmptyKey) return;
    if (k == key)
    {
      const uint32_t l = mt.hvals[h];
      const uint32_t n = (uint32_t(x & 7) << 6) | (uint32_t(y & 7) << 3) | uint32_t(z & 7);
      *out_val         = mt.leaf_vals[size_t(l) * 512 + n];
      *out_active      = int32_t((mt.leaf_mask[size_t(l) * 8 + (n >> 6)] >> (n & 63)) & 1);
      return;
    }
    h = (h + 1) & mt.hcap_mask;
  }
}

// multi-GPU: bin touched update leaves by owner rank. pass 0 counts, pass 1 scatters (and clears the slots).
__global__ void partition_kernel(UpdateTable ut, uint32_t n, int32_t n_ranks, uint32_t* rank_counts, uint32_t* rank_cursor,
                                 LeafRecord* out, int pass)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t s   = ut.touched[i];
  const uint64_t key = ut.keys[s];
  const int32_t r    = leafOwner(key, n_ranks);
  if (pass == 0) { atomicAdd(rank_counts + r, 1u); return; }
  const uint32_t dst = atomicAdd(rank_cursor + r, 1u);
  out[dst].key       = key;
#pragma unroll
  for (int w = 0; w < 8; ++w)
  {
    out[dst].active[w] = ut.active[size_t(s) * 8 + w];
    out[dst].value[w]  = ut.value[size_t(s) * 8 + w];
    ut.active[size_t(s) * 8 + w] = 0;
    ut.value[size_t(s) * 8 + w]  = 0;
  }
  ut.keys[s] = kEmptyKey;
}

// ====================================================================================================
// launch wrappers
// ====================================================================================================
static inline unsigned blocksFor(uint64_t n, unsigned per_block) { return unsigned((n + per_block - 1) / per_block); }

static std::atomic<uint32_t> g_launches{0};
uint32_t launchCount() { return g_launches.load(); }
#define VDBM_COUNT_LAUNCH() g_launches.fetch_add(1, std::memory_order_relaxed)

void launchPrepRays(const RaycastArgs& a, Counters* ctr, cudaStream_t s)
{
  if (a.n == 0) return;
  { prep_rays_kernel<<<blocksFor(a.n, 256), 256, 0, s>>>(a, ctr); VDBM_COUNT_LAUNCH(); }
}

int raycastDDAGrid(int device)
{
  int sms = 148, per_sm = 4;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, raycast_dda_kernel, 256, 0);
  if (per_sm < 1) per_sm = 1;
  return sms * per_sm; // persistent: one wave exactly
}

void launchRaycastDDA(const RaycastArgs& a, UpdateTable ut, Counters* ctr, int grid, cudaStream_t s)
{
  if (a.n == 0) return;
  const uint64_t warps_needed = (a.n + 31) / 32;
  const uint64_t blocks       = (warps_needed + 7) / 8;
  if (uint64_t(grid) > blocks) grid = int(blocks);
  { raycast_dda_kernel<<<grid, 256, 0, s>>>(a, ut, ctr); VDBM_COUNT_LAUNCH(); }
}

void launchApplyUpdate(UpdateTable ut, MapTable mt, LogOdds lo, LeafRecord* change_out, uint32_t change_cap, Counters* ctr,
                       uint32_t n_touched_hint, cudaStream_t s)
{
  if (n_touched_hint == 0) return;
  static int grid_cap = 0;
  if (grid_cap == 0)
  {
    int dev = 0, sms = 148, per_sm = 4;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, apply_update_kernel, 256, 0);
    grid_cap = sms * (per_sm < 1 ? 1 : per_sm);
  }
  unsigned grid = blocksFor(uint64_t(n_touched_hint) * 32, 256);
  if (grid > unsigned(grid_cap)) grid = unsigned(grid_cap);
  { apply_update_kernel<<<grid, 256, 0, s>>>(ut, mt, lo, change_out, change_cap, ctr); VDBM_COUNT_LAUNCH(); }
}

void launchRehashUpdate(UpdateTable old_t, uint32_t old_n, UpdateTable new_t, Counters* ctr, cudaStream_t s)
{
  if (old_n) { rehash_update_kernel<<<blocksFor(old_n, 256), 256, 0, s>>>(old_t, old_n, new_t, ctr); VDBM_COUNT_LAUNCH(); }
}
void launchRehashMap(MapTable mt, uint32_t n_leaves, Counters* ctr, cudaStream_t s)
{
  if (n_leaves) { rehash_map_kernel<<<blocksFor(n_leaves, 256), 256, 0, s>>>(mt, n_leaves, ctr); VDBM_COUNT_LAUNCH(); }
}
void launchGatherUpdate(UpdateTable ut, uint32_t n, const uint32_t* order, LeafRecord* out, cudaStream_t s)
{
  if (n) { gather_update_kernel<<<blocksFor(uint64_t(n) * 16, 256), 256, 0, s>>>(ut, n, order, out); VDBM_COUNT_LAUNCH(); }
}
void launchImportUpdate(UpdateTable ut, const LeafRecord* recs, uint64_t n, Counters* ctr, cudaStream_t s)
{
  if (n) { import_update_kernel<<<blocksFor(n * 16, 256), 256, 0, s>>>(ut, recs, n, ctr); VDBM_COUNT_LAUNCH(); }
}
void launchClearUpdate(UpdateTable ut, uint32_t n, cudaStream_t s)
{
  if (n) { clear_update_kernel<<<blocksFor(uint64_t(n) * 16, 256), 256, 0, s>>>(ut, n); VDBM_COUNT_LAUNCH(); }
}
void launchGatherMap(MapTable mt, uint32_t n, const uint32_t* leaf_idx, int32_t* origins, uint64_t* mask, float* vals,
                     int clear_dirty, cudaStream_t s)
{
  if (!n) return;
  unsigned grid = blocksFor(uint64_t(n) * 32, 256);
  if (grid > 148u * 8u) grid = 148u * 8u;
  { gather_map_kernel<<<grid, 256, 0, s>>>(mt, n, leaf_idx, origins, mask, vals, clear_dirty); VDBM_COUNT_LAUNCH(); }
}
void launchSection(MapTable mt, uint32_t n_leaves, const int32_t bbmin[3], const int32_t bbmax[3], int full, int result_float,
                   uint64_t* out_keys, uint64_t* out_active, uint64_t* out_valmask, float* out_vals, uint32_t out_cap,
                   Counters* ctr, cudaStream_t s)
{
  if (!n_leaves) return;
  unsigned grid = blocksFor(uint64_t(n_leaves) * 32, 256);
  if (grid > 148u * 8u) grid = 148u * 8u;
  section_kernel<<<grid, 256, 0, s>>>(mt, n_leaves, bbmin[0], bbmin[1], bbmin[2], bbmax[0], bbmax[1], bbmax[2], full, result_float,
                                      out_keys, out_active, out_valmask, out_vals, out_cap, ctr);
}
void launchProbe(MapTable mt, int32_t x, int32_t y, int32_t z, float* out_val, int32_t* out_active, cudaStream_t s)
{
  { probe_kernel<<<1, 1, 0, s>>>(mt, x, y, z, out_val, out_active); VDBM_COUNT_LAUNCH(); }
}
void launchPartition(UpdateTable ut, uint32_t n, int32_t n_ranks, uint32_t* rank_counts, uint32_t* rank_offsets, LeafRecord* out,
                     int pass, cudaStream_t s)
{
  if (n) { partition_kernel<<<blocksFor(n, 256), 256, 0, s>>>(ut, n, n_ranks, rank_counts, rank_offsets, out, pass); VDBM_COUNT_LAUNCH(); }
}
void launchKeysFromSlots(const uint64_t* keys, const uint32_t* slots, uint32_t n, uint64_t* out_keys, uint32_t* out_idx, cudaStream_t s)
{
  if (n) { keys_from_slots_kernel<<<blocksFor(n, 256), 256, 0, s>>>(keys, slots, n, out_keys, out_idx); VDBM_COUNT_LAUNCH(); }
}
void launchIota(uint32_t* out, uint32_t n, cudaStream_t s)
{
  if (n) { iota_kernel<<<blocksFor(n, 256), 256, 0, s>>>(out, n); VDBM_COUNT_LAUNCH(); }
}
void launchUnpackOrigins(const uint64_t* keys, uint32_t n, int32_t* origins, cudaStream_t s)
{
  if (n) { unpack_origins_kernel<<<blocksFor(n, 256), 256, 0, s>>>(keys, n, origins); VDBM_COUNT_LAUNCH(); }
}
void launchSplitRecords(const LeafRecord* recs, uint32_t n, int32_t* origins, uint64_t* active, uint64_t* value, cudaStream_t s)
{
  if (n) { split_records_kernel<<<blocksFor(uint64_t(n) * 16, 256), 256, 0, s>>>(recs, n, origins, active, value); VDBM_COUNT_LAUNCH(); }
}

size_t sortPairs(void* d_temp, size_t temp_bytes, const uint64_t* keys_in, uint64_t* keys_out, const uint32_t* idx_in,
                 uint32_t* idx_out, uint32_t n, cudaStream_t s)
{
  size_t bytes = temp_bytes;
  if (d_temp == nullptr) bytes = 0;
  cub::DeviceRadixSort::SortPairs(d_temp, bytes, keys_in, keys_out, idx_in, idx_out, int(n), 0, 63, s);
  return bytes;
}

} // namespace vdbm
