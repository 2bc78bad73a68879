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

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cstdlib>
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

// ---- update-grid brick hash: find or insert a brick slot (slot == storage; masks are zero for empty slots) ----
__device__ __forceinline__ uint32_t brickFindOrInsert(const UpdateGrid& g, uint64_t bkey, Counters* ctr)
{
  uint32_t h = uint32_t(mix64(bkey)) & g.cap_mask;
  const uint32_t max_probe = min(g.cap_mask, 1024u);
  for (uint32_t probe = 0; probe <= max_probe; ++probe)
  {
    uint64_t k = ldcg64(g.bkeys + h);
    if (k == bkey) return h;
    if (k == kEmptyKey)
    {
      unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long*>(g.bkeys + h), kEmptyKey, bkey);
      if (old == kEmptyKey)
      {
        uint32_t idx    = atomicAdd(g.counters, 1u);
        g.btouched[idx] = h; // idx < cap always: at most cap successful claims
        return h;
      }
      if (old == bkey) return h;
    }
    h = (h + 1) & g.cap_mask;
  }
  atomicOr(&ctr->flags, kFlagUpdateOverflow);
  return kInvalid;
}

// Same, but a full table yields the TRASH brick (slot == capacity; the mask arrays hold capacity + 1 bricks): marks
// go somewhere harmless, the overflow flag makes the host grow the table and replay the scan. Lets the DDA kernel
// mark without validity checks.
__device__ __forceinline__ uint32_t brickFindOrInsertOrTrash(const UpdateGrid& g, uint64_t bkey, Counters* ctr)
{
  const uint32_t s = brickFindOrInsert(g, bkey, ctr);
  return s == kInvalid ? g.cap_mask + 1u : s;
}

// word offset (in uint64 units) of voxel (x,y,z)'s mask word inside its brick: leaf_in_brick * 8 + (x & 7)
__device__ __forceinline__ uint32_t brickWordOffset(int x, int y, int z)
{
  return (uint32_t(x & 56) << 6) | (uint32_t(y & 56) << 3) | uint32_t(z & 56) | uint32_t(x & 7);
}

} // namespace

// ====================================================================================================
// K0: per-point set-up. One thread per point; fully convergent, so the expensive fp64 div/sqrt/fmod run at
// full SIMD efficiency here instead of inside the divergent DDA kernel.
// ====================================================================================================
// sort keys of the work items (rays / segments): min(marks, 2047), sorted on bits [3, 11) only
constexpr int kSortLoBit = 3, kSortHiBit = 11;
constexpr uint32_t kSortMinKey = 1u << kSortLoBit, kSortMaxKey = (1u << kSortHiBit) - 1u;

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
    RayRec r;
    r.flags  = 0;
    r.visits = 0;
    r.delta[0] = r.delta[1] = r.delta[2] = DBL_MAX;
    uint32_t sign_bits = 0;
    int4 end_rec = make_int4(0, 0, 0, 0); // end voxel + (bit0 valid, bit1 hit): the "reduced" update of this scan
    double e[3] = {0.0, 0.0, 0.0};
    bool finite, index_ray = false, max_range_ray = false, mine = true;
    int32_t ei_given[3] = {0, 0, 0};
    if (a.index_mode)
    {
      // receiver side of a reduced update: the record IS the end voxel (x, y, z, flags), no world arithmetic
      const int4 q = *reinterpret_cast<const int4*>(a.points + i * a.stride);
      finite        = (q.w & 1) != 0;
      index_ray     = true;
      max_range_ray = (q.w & 2) == 0;
      ei_given[0] = q.x; ei_given[1] = q.y; ei_given[2] = q.z;
    }
    else
    {
      const float* p = reinterpret_cast<const float*>(a.points + i * a.stride);
      e[0] = double(p[0]); e[1] = double(p[1]); e[2] = double(p[2]); // VDBMapping.hpp:501
      // VDBMapping.hpp:505-510 skips NaN; +-inf is undefined behaviour in the reference and is dropped here too
      finite = isfinite(e[0]) && isfinite(e[1]) && isfinite(e[2]);
      if (a.sector_n > 1)
      {
        // one scan split across the GPUs of a box: every rank sees the whole cloud and keeps the rays of its azimuth sector
        // (invalid points count as angle 0, so exactly one rank reports them)
        double ang = finite ? diamondAngle(__dsub_rn(e[0], a.origin[0]), __dsub_rn(e[1], a.origin[1])) : 0.0;
        if (!(ang >= 0.0 && ang < 4.0)) ang = 0.0;
        int32_t sector = a.sector_n - 1;
        for (int32_t r = 0; r < a.sector_n; ++r)
          if (ang >= a.sector_bounds[r]) sector = r;
        mine = sector == a.sector_rank;
      }
    }
    if (!mine) {}
    else if (!finite) nan_skipped = index_ray ? 0 : 1;
    else
    {
      if (!index_ray && a.range > 0.0)
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
        const double fl = index_ray ? double(ei_given[k])
                                    : floor(__dmul_rn((fmod(e[k], a.resolution) != 0.0) ? __dadd_rn(e[k], a.half_res) : e[k], a.inv_res));
        if (!(fabs(fl) < double(kVoxelLimit))) in_range = false;
        const int32_t ei = in_range ? int32_t(fl) : 0;
        (&end_rec.x)[k]  = ei;
        const double dir = __dsub_rn(double(ei), double(a.origin_idx[k])); // exact
        if (dir != 0.0)
        {
          zero       = false;
          sign_bits |= (dir > 0.0 ? 1u : 2u) << (4 + 2 * k);
          r.delta[k] = fabs(__ddiv_rn(1.0, dir)); // Ray::mInvDir = 1/dir; DDA::mDelta = step * inv = |inv|
          l1 += (long long)fabs(dir);
        }
      }
      if (!in_range) range_err = 1;
      else
      {
        r.flags  = kRayValid | (max_range_ray ? kRayClipped : 0u) | (zero ? kRayZeroLen : 0u) | sign_bits;
        visits   = zero ? 0ull : (unsigned long long)(1 + l1); // castRayIntoGrid marks 1 + |dx|+|dy|+|dz| voxels
        r.visits = uint32_t(visits);
        end_rec.w = 1 | (max_range_ray ? 0 : 2);
      }
    }
    a.rays[i] = r;
    if (a.ends) a.ends[i] = end_rec;
    // segment i = this ray's first (or only) segment, starting at the sensor voxel
    SegRec sg;
    sg.next[0] = (r.delta[0] == DBL_MAX) ? DBL_MAX : __dmul_rn(0.5, r.delta[0]); // DDA::init: 0.5 * |inv| exactly
    sg.next[1] = (r.delta[1] == DBL_MAX) ? DBL_MAX : __dmul_rn(0.5, r.delta[1]);
    sg.next[2] = (r.delta[2] == DBL_MAX) ? DBL_MAX : __dmul_rn(0.5, r.delta[2]);
    sg.m[0] = sg.m[1] = sg.m[2] = 0;
    sg.count = r.visits;
    sg.ray   = uint32_t(i);
    sg.last  = 1;
    // zero-length rays with a hit still need one visit by the DDA kernel (count 0, flagged in the ray record)
    uint32_t key = uint32_t(visits > kSortMaxKey ? kSortMaxKey : visits);
    if ((r.flags & kRayValid) && (r.flags & kRayZeroLen) && !(r.flags & kRayClipped)) key = 1;
    if (a.seg_len != 0 && r.visits > a.seg_len + a.seg_len / 2)
    {
      // long ray: reserve P - 1 extra segment slots; if the pool is exhausted the ray simply stays whole (the
      // reservation is not given back: every later one fails too, and the unused slots keep a zero sort key)
      const uint32_t P    = (r.visits + a.seg_len - 1) / a.seg_len;
      const uint32_t base = atomicAdd(&ctr->n_extra, P - 1);
      if (uint64_t(a.n) + base + (P - 1) <= a.seg_cap)
      {
        a.seg_base[i] = uint32_t(a.n) + base;
        a.long_rays[atomicAdd(&ctr->n_long, 1u)] = uint32_t(i);
        key = 0; // finalised (with the extra segments) by long_ray_segments_kernel
      }
    }
    a.segs[i]      = sg;
    // The sort only looks at key bits [3, 11) (ONE radix pass: longest-first scheduling needs no finer order than 8 marks,
    // and everything above 2047 marks simply goes first): a work item with 1..7 marks would tie with the empty items
    // (key 0) and could end up behind more zero keys than there are lanes to retire on them. Lift every real item above the zeros.
    a.sort_keys[i] = (key != 0u && key < kSortMinKey) ? kSortMinKey : key;
    a.sort_idx[i]  = uint32_t(i);
  }
  if (a.fast_mode) visits = 0; // fast_mode counts the marks castRayIntoGridFast really makes (raycast_fast_kernel)
  // warp-aggregated statistics
  const unsigned vmax = __reduce_max_sync(kFull, unsigned(visits));
  visits      = __reduce_add_sync(kFull, unsigned(visits)); // per-lane visits <= 1 + 3*2^24, 32 lanes fit in 32 bits
  nan_skipped = __reduce_add_sync(kFull, nan_skipped);
  clipped     = __reduce_add_sync(kFull, clipped);
  range_err   = __reduce_add_sync(kFull, range_err);
  if ((threadIdx.x & 31) == 0)
  {
    if (visits) atomicAdd(&ctr->visits, visits);
    if (vmax) atomicMax(&ctr->max_visits, vmax);
    if (nan_skipped) atomicAdd(&ctr->nan_skipped, (unsigned long long)nan_skipped);
    if (clipped) atomicAdd(&ctr->clipped, (unsigned long long)clipped);
    if (range_err) atomicOr(&ctr->flags, kFlagCoordRange);
  }
}

// ====================================================================================================
// K0b: segment boundaries of long rays. Three adjacent lanes per ray, one per axis: each walks ITS axis' crossing-time
// sequence v_0 = d/2, v_{k+1} = fl(v_k + d) (the same fp64 additions the DDA performs, in the same order) and records,
// for every threshold t_j = j / P, how many steps precede it and the value of `next` there. No marking, no memory
// traffic inside the loop: ~30x cheaper per step than the DDA itself, and it removes the serial floor of one ray.
// ====================================================================================================
__global__ void __launch_bounds__(128) long_ray_segments_kernel(RaycastArgs a, uint32_t n_long)
{
  // a warp takes 10 rays: lanes 0..29 = (ray, axis) pairs, lanes 30 and 31 idle, so a ray never straddles warps
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane      = threadIdx.x & 31;
  const uint32_t li   = warp * 10u + uint32_t(lane / 3);
  const int axis      = lane % 3;
  if (lane >= 30 || li >= n_long) return;
  const uint32_t ray = a.long_rays[li];
  const RayRec r     = a.rays[ray];
  const uint32_t P   = (r.visits + a.seg_len - 1) / a.seg_len;
  const uint32_t xb  = a.seg_base[ray];           // extra segments j = 1..P-1 live at xb + j - 1
  const double d     = r.delta[axis];
  const double invP  = 1.0 / double(P);
  double v   = (d == DBL_MAX) ? DBL_MAX : __dmul_rn(0.5, d);
  uint32_t m = 0;
  for (uint32_t j = 1; j < P; ++j)
  {
    const double tj = double(j) * invP;           // any threshold works as long as the three axes use the same value
    while (v < tj) { v = __dadd_rn(v, d); ++m; }
    SegRec* sg    = a.segs + (xb + j - 1);
    sg->next[axis] = v;
    sg->m[axis]    = m;
  }
  // the three lanes of this ray are in the same warp: make their writes visible to the finalising lane
  __syncwarp(__activemask());
  if (axis != 0) return;
  // finalise: marks per segment = steps between consecutive boundaries; the last one runs to the end voxel
  uint32_t s_prev = 0; // steps before segment j
  for (uint32_t j = 0; j < P; ++j)
  {
    SegRec* sg = (j == 0) ? a.segs + ray : a.segs + (xb + j - 1);
    uint32_t s_next;
    if (j + 1 < P)
    {
      const SegRec* nx = a.segs + (xb + j);
      s_next           = nx->m[0] + nx->m[1] + nx->m[2];
    }
    else s_next = r.visits; // marks V_{s_prev} .. V_T, T = visits - 1
    const uint32_t cnt = s_next - s_prev;
    sg->count = cnt;
    sg->ray   = ray;
    sg->last  = (j + 1 == P) ? 1u : 0u;
    const uint32_t slot = (j == 0) ? ray : (xb + j - 1);
    a.sort_keys[slot]   = cnt > kSortMaxKey ? kSortMaxKey : ((cnt != 0u && cnt < kSortMinKey) ? kSortMinKey : cnt); // see prep_rays_kernel
    a.sort_idx[slot]    = slot;
    s_prev              = s_next;
  }
}

// ====================================================================================================
// K1: 3D-DDA. Persistent warps; every lane owns one ray at a time and refills itself from a global cursor
// when its ray ends, so ray-length variance (10..2000+ visits) costs no idle lanes.
// Voxel stepping replays openvdb::math::DDA<Ray<double>,0> bit for bit:
//   next[a] = 0.5*|1/dir[a]| (exact), then repeated  next[axis] += delta[axis]  in fp64,
//   axis = MinIndex(next) with its tie table {2,1,9,1,2,9,0,0}, continue while t <= 1.0.
// The step is branch-free (selects), so the 32 lanes stay converged whatever axis each one takes.
//
// Marking (round 2: the instruction diet). While this kernel runs, the masks of the update grid are Z-SLICE words: word
// (z & 7) of a leaf holds its 8x8 x-y tile, bit (y&7)<<3 | (x&7). LiDAR rays are close to horizontal, so a ray stays 4-5
// steps inside one such word (tools/word_orientation_study.py: 2.2-2.8x fewer flushes than with OpenVDB's x-slice
// words). Bits of consecutive visits that fall in the same word are merged in a register and flushed with ONE
// fire-and-forget red.global.or.b64. compact_leaves_kernel<true>, which reads every touched leaf anyway, turns the
// touched leaves into OpenVDB's x-slice layout in place (an 8x8 bit transpose per byte lane), so every consumer of the
// grid keeps seeing the reference's layout.
// The voxel position inside the current 64^3 brick is ONE integer, the "voxel code":
//   bits 0-5 x&63 | bits 6-7 guard | bits 8-13 y&63 | bits 14-15 guard | bits 16-21 z&63 | bits 22-23 guard
// A DDA step is one add of +-1 / +-2^8 / +-2^16. Guards rest at binary 10: leaving the brick upwards carries into them
// (11), downwards borrows from them (01), never into the neighbouring field. "Did the mask word change" and "did the
// ray leave the brick" are each one masked compare; word index and bit index are a few shifts of the code. The brick
// hash is consulted only when a ray enters a new brick, and those lookups (like the ray refills) are batched at
// warp-uniform points every kBatch steps so that their L2 round trip is paid once per batch and not once per lane event.
// ====================================================================================================
#ifndef VDBM_DDA_CTAS
#define VDBM_DDA_CTAS 4 // resident CTAs per SM the DDA kernel is compiled for: 52 registers at 4 (measured on B200: 0.42 ms per cfg2 scan; 5 CTAs x 48 registers 0.47 ms, 6 x 40 with spills 0.52 ms)
#endif
#ifndef VDBM_DDA_BLOCK
#define VDBM_DDA_BLOCK 256 // threads per DDA CTA (a multiple of 32)
#endif
#ifndef VDBM_KBATCH
#define VDBM_KBATCH 8
#endif
constexpr int kBatch = VDBM_KBATCH;

constexpr uint32_t kCodeGuardRest = 0x00808080u; // guards at rest (binary 10 per field)
constexpr uint32_t kCodeGuardMask = 0x00C0C0C0u;
constexpr uint32_t kCodeWordMask  = 0x00FFF8F8u; // x>>3, y>>3, z and the guards: any change = another mask word
__host__ __device__ __forceinline__ uint32_t voxelCode(int x, int y, int z)
{
  return uint32_t(x & 63) | (uint32_t(y & 63) << 8) | (uint32_t(z & 63) << 16) | kCodeGuardRest;
}
// word offset inside the brick: leaf_in_brick * 8 + (z & 7) = (x>>3)<<9 | (y>>3)<<6 | z     (brick-local x, y, z)
__host__ __device__ __forceinline__ uint32_t codeWord(uint32_t c) { return ((c & 0x38u) << 6) | ((c >> 5) & 0x1C0u) | ((c >> 16) & 0x3Fu); }
// bit inside a z-slice word: (y&7)<<3 | (x&7)
__host__ __device__ __forceinline__ uint32_t codeBit(uint32_t c) { return (c & 7u) | ((c >> 5) & 0x38u); }

// MODE is a profiling hook: 0 = product path; 1 = no mask writes (pure traversal cost); 2 = plain 64-bit stores instead of
// REDs (LSU path without the L2 atomic unit). Modes 1/2 give WRONG maps and only exist in builds made with
// -DVDBM_EXPERIMENTS (then selected with env VDBM_DDA_MODE); the product library always runs mode 0.
// MODE 4 (product; chosen per scan by the host for scans whose rays overlap heavily, e.g. a depth camera: 47 visits per
// distinct voxel in BASELINE config 3): test before set. Bits are only ever SET while this kernel runs, so a (possibly
// stale, L1-cached) load that already shows all bits of `v` proves the RED redundant; a stale zero merely costs the RED
// it would have cost anyway. There ~98 % of the REDs - same-address atomics that serialise in L2 - become L1/L2 load hits.
// For LiDAR scans (2-3 visits per voxel) the extra dependent load would only add latency, hence a separate instantiation.
template <int MODE>
__device__ __forceinline__ void markWord(uint64_t* p, uint64_t v)
{
  if (MODE == 4)
  {
    const uint64_t old = __ldca(reinterpret_cast<const unsigned long long*>(p));
    if ((old & v) != v) redOr64(p, v);
    return;
  }
  if (MODE == 0) redOr64(p, v);
  else if (MODE == 2) asm volatile("st.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// same under a predicate (one predicated RED, no branch)
template <int MODE>
__device__ __forceinline__ void markWordIf(uint64_t* p, uint64_t v, bool pred)
{
  if (MODE == 0)
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %2, 0; @q red.global.or.b64 [%0], %1; }" ::"l"(p), "l"(v), "r"((unsigned)pred) : "memory");
  else if (pred) markWord<MODE>(p, v);
}

// Near field: every ray starts at the sensor, so the mask words within a few leaves of it receive an update
// from (almost) every ray: measured on B200 those same-address REDs serialise in a handful of L2 slices and cost
// a third of the kernel. The 2x2x2 bricks around the sensor (a 128^3-voxel region with the sensor at least 32
// voxels inside) are therefore PRIVATISED: kNearCopies zero-initialised copies of their active masks live in
// global memory, a CTA marks copy (blockIdx % kNearCopies), and merge_near_kernel ORs the copies into the real
// bricks afterwards. The hot addresses are spread kNearCopies-fold and the inner loop keeps ONE uniform mark path
// (a brick is just a base pointer).
constexpr int kNearCopies = 32;
constexpr int kNearBricks = 8;
constexpr uint64_t kNearKeyBits = (uint64_t(1) << 42) | (uint64_t(1) << 21) | uint64_t(1);

// first brick (per axis) of the 2x2x2 near region: the brick below the sensor's if the sensor sits in the lower half
__host__ __device__ __forceinline__ int nearBrick0(int o) { return (o >> 6) - (((o & 63) < 32) ? 1 : 0); }

// in-place predicated add (one instruction; keeps the three DDA axes branch-free without select chains)
__device__ __forceinline__ void addIf(double& n, double d, bool p)
{
  asm("{ .reg .pred q; setp.ne.u32 q, %2, 0; @q add.rn.f64 %0, %0, %1; }" : "+d"(n) : "d"(d), "r"((unsigned)p));
}

// Pipe budget (ncu, B200): this loop is bound by the 16-lane ALU pipe (LOP3 / SHF / SEL / ISETP: 2 cycles per warp
// instruction; 70 % busy, 84 % on the busiest SM), not by issue slots, fp64 or memory. Hence the odd-looking choices
// below, each of which moves work OFF that pipe:
//  * "next[axis] += delta[axis]" is a DFMA with a multiplier that is exactly 1.0 or 0.0 (one SEL for the high word; the
//    low word of both constants is zero): fma(d, 1, n) = rn(n + d) and fma(d, 0, n) = n exactly (d is finite), so the
//    rounding sequence is the reference's; the compiler turns a predicated DADD into DADD + 2 FSEL instead.
//  * the code is stepped with three predicated integer adds (FMA pipe) instead of two SELs and an add;
//  * 1 << bit comes from a shared-memory table indexed with the masked code (one LOP3 + LDS instead of 5 ALU ops);
//  * "lane is running" is one test of the code word itself: bit 31 = no ray, low guard bits = waiting for a brick lookup.
constexpr uint32_t kCodeLowGuards = 0x00404040u; // set (whatever the direction) by a step that leaves the brick
constexpr uint32_t kCodeIdle      = 0x80000000u; // lane has no ray
constexpr uint32_t kBitTableSize  = 0x708u;      // indexed by code & 0x707 (x&7 | (y&7) << 8)

__device__ __forceinline__ void addIfInt(uint32_t& n, int d, bool p)
{
  asm("{ .reg .pred q; setp.ne.u32 q, %2, 0; @q add.s32 %0, %0, %1; }" : "+r"(n) : "r"(d), "r"((unsigned)p));
}
// n + (p ? d : 0) with the reference's single rounding, see above
__device__ __forceinline__ double addSel(double n, double d, bool p)
{
  return __fma_rn(d, __hiloint2double(p ? 0x3FF00000 : 0, 0), n);
}

template <int MODE>
__global__ void __launch_bounds__(VDBM_DDA_BLOCK, VDBM_DDA_CTAS) raycast_dda_kernel(RaycastArgs a, UpdateGrid g, uint64_t* near_act, Counters* ctr)
{
  __shared__ uint32_t s_near_slot[kNearBricks];
  __shared__ uint64_t s_bit[kBitTableSize];
  const int lane = threadIdx.x & 31;
  const int ox = a.origin_idx[0], oy = a.origin_idx[1], oz = a.origin_idx[2];
  const int nbx = nearBrick0(ox), nby = nearBrick0(oy), nbz = nearBrick0(oz);
  // resolve the real slots of the 8 near bricks once per CTA (endpoint hits and the merge kernel use them)
  if (threadIdx.x < kNearBricks)
    s_near_slot[threadIdx.x] = brickFindOrInsertOrTrash(g, packLeafKey(nbx + (threadIdx.x >> 2), nby + ((threadIdx.x >> 1) & 1), nbz + (threadIdx.x & 1)), ctr);
  for (uint32_t i = threadIdx.x; i < kBitTableSize; i += blockDim.x) s_bit[i] = uint64_t(1) << ((i & 7u) | ((i >> 5) & 0x38u));
  __syncthreads();
  uint64_t* const my_near  = near_act + size_t(blockIdx.x % kNearCopies) * (kBrickLeaves * 8 * kNearBricks);
  const uint64_t near_key0 = packLeafKey(nbx, nby, nbz);
  uint32_t bit_table; // shared-window address of s_bit, opaque to the compiler so that it stays in ONE register
  asm("mov.b32 %0, %1;" : "=r"(bit_table) : "r"(uint32_t(__cvta_generic_to_shared(s_bit))));

  bool done = false;
  double n0 = 0, n1 = 0, n2 = 0, d0 = 0, d1 = 0, d2 = 0;
  uint64_t bit = 0;              // mask bit of the current voxel, fetched one step ahead (shared-memory latency)
  uint32_t c = kCodeIdle | kCodeGuardRest; // voxel code of the current voxel inside its brick (+ lane state, see above)
  int ix = 0, iy = 0, iz = 0;    // code increments of one step along x / y / z
  uint64_t bkey = 0;             // key of the current brick
  uint32_t remaining = 0; // voxels still to mark on the current ray (exact: 1 + |dx|+|dy|+|dz|, see prep_rays_kernel)
  uint32_t clipped   = 0;
  uint32_t slot      = kInvalid;  // real brick slot (value masks, validity)
  uint64_t* act_base = my_near;   // where this brick's active mask words go: a private near copy or the real brick
  uint64_t acc       = 0;         // bits collected for the current mask word

  for (;;)
  {
    // ================= batch point (warp-uniform): refills + brick lookups, every kBatch voxel steps =================
    {
      // ---- refill idle lanes (warp-aggregated fetch from the global ray cursor) ----
      const bool idle     = (c & kCodeIdle) != 0;
      const unsigned want = __ballot_sync(kFull, idle && !done);
      if (want)
      {
        unsigned base    = 0;
        const int leader = __ffs(want) - 1;
        if (lane == leader) base = atomicAdd(&ctr->ray_cursor, (unsigned)__popc(want));
        base = __shfl_sync(kFull, base, leader);
        if (idle && !done)
        {
          const uint64_t idx = uint64_t(base) + __popc(want & ((1u << lane) - 1u));
          // keys are sorted descending: the first zero key means only empty work items (NaN / zero-length rays without a
          // hit, unused segment slots) follow
          if (idx >= a.n_segs || a.sorted_keys[idx] == 0u) done = true;
          else
          {
            const SegRec sg = a.segs[a.order[idx]]; // longest segments first: the tail of the kernel is made of short ones
            const RayRec r  = a.rays[sg.ray];
            if (r.flags & kRayValid)
            {
              if (r.flags & kRayZeroLen)
              {
                // no DDA (VDBMapping.hpp:559); a non-clipped endpoint is still set on with value true (:533-536)
                if (!(r.flags & kRayClipped))
                {
                  const int origin_ni = (((ox >> 6) - nbx) << 2) | (((oy >> 6) - nby) << 1) | ((oz >> 6) - nbz);
                  const uint32_t co   = voxelCode(ox, oy, oz);
                  const size_t w      = size_t(s_near_slot[origin_ni]) * (kBrickLeaves * 8) + codeWord(co);
                  const uint64_t bit  = uint64_t(1) << codeBit(co);
                  redOr64(g.act + w, bit);
                  redOr64(g.val + w, bit);
                }
              }
              else if (sg.count != 0)
              {
                d0 = r.delta[0]; d1 = r.delta[1]; d2 = r.delta[2];
                n0 = sg.next[0]; n1 = sg.next[1]; n2 = sg.next[2];
                const int sx = rayStep(r.flags, 0), sy = rayStep(r.flags, 1), sz = rayStep(r.flags, 2);
                const int x = ox + sx * int(sg.m[0]), y = oy + sy * int(sg.m[1]), z = oz + sz * int(sg.m[2]);
                // low guards set = "look my brick up" (right below; a first segment starts in the near field); with all
                // high guards clear the lookup recognises a fresh segment whose bkey is already right
                c    = (voxelCode(x, y, z) & ~kCodeGuardMask) | kCodeLowGuards;
                bkey = packLeafKey(x >> 6, y >> 6, z >> 6);
                ix = sx; iy = sy * 256; iz = sz * 65536;
                remaining = sg.count;
                // only the segment that reaches the end voxel delivers the hit; a clipped ray never does
                clipped = (r.flags & kRayClipped) | (sg.last ? 0u : 1u);
                acc     = 0;
              }
            }
          }
        }
      }
      // ---- batched brick lookups for lanes that entered a new brick ----
      const bool lookup = !(c & kCodeIdle) && (c & kCodeLowGuards) != 0;
      if (__any_sync(kFull, lookup))
      {
        if (lookup)
        {
          // a step that left the brick shows in the guard bits of that axis: 11 = one brick up, 01 = one brick down, the
          // other two axes stay at rest (10). All three at 01 can only be the refill's "fresh segment" mark.
          const int gx = int((c >> 6) & 3u), gy = int((c >> 14) & 3u), gz = int((c >> 22) & 3u);
          const bool fresh = (c & kCodeGuardMask) == kCodeLowGuards; // straight from the refill: bkey is already right
          const long long dkey = fresh ? 0ll
                                       : ((long long)((gx & 1) * (gx - 2)) << 42) + ((long long)((gy & 1) * (gy - 2)) << 21) +
                                           (long long)((gz & 1) * (gz - 2));
          bkey += (uint64_t)dkey;
          c = (c & ~kCodeGuardMask) | kCodeGuardRest;
          // near field? (per-axis brick offset 0 or 1 from the first near brick: no borrow can fake that)
          const uint64_t diff = bkey - near_key0;
          if ((diff & ~kNearKeyBits) == 0)
          {
            const int ni = int(((diff >> 42) & 1) << 2) | int(((diff >> 21) & 1) << 1) | int(diff & 1);
            slot         = s_near_slot[ni];
            act_base     = my_near + size_t(ni) * (kBrickLeaves * 8);
          }
          else
          {
            slot     = brickFindOrInsertOrTrash(g, bkey, ctr);
            act_base = g.act + size_t(slot) * (kBrickLeaves * 8);
          }
        }
      }
      if (__all_sync(kFull, done && (c & kCodeIdle) != 0)) break;
      asm("ld.shared.u64 %0, [%1];" : "=l"(bit) : "r"((c & 0x707u) * 8u + bit_table));
    }

    // ================= kBatch voxel steps =================
#pragma unroll
    for (int u = 0; u < kBatch; ++u)
    {
      if ((c & (kCodeIdle | kCodeLowGuards)) == 0)
      {
        // ---- mark current voxel (setActiveState(dda.voxel(), true), VDBMapping.hpp:563) ----
        acc |= bit;
        if (--remaining == 0)
        {
          // Last voxel of the ray (= the end voxel). OpenVDB's loop ends when the NEXT crossing time exceeds t1 = 1;
          // that happens exactly after 1 + |dx|+|dy|+|dz| marks (the crossing times of axis a are (m + 0.5)/|d_a| up to
          // fp64 rounding, m < |d_a| <=> t < 1 with a margin of 0.5/|d_a| >> accumulated rounding for |d_a| < 2^24).
          // Flush; the end voxel also receives the hit unless the ray was clipped (VDBMapping.hpp:533-536).
          const uint32_t w = codeWord(c);
          markWord<MODE>(act_base + w, acc);
          if (!clipped) markWord<MODE>(g.val + size_t(slot) * (kBrickLeaves * 8) + w, bit);
          c |= kCodeIdle;
        }
        else
        {
          // ---- DDA::step(): axis = MinIndex(next); next[axis] += delta[axis]; voxel[axis] += step[axis] ----
          // MinIndex table {2,1,9,1,2,9,0,0} on key ((n0<n1)<<2)+((n0<n2)<<1)+(n1<n2):
          //   (n0<n1 && n0<n2) -> x ; else (n1<n2) -> y ; else z      (keys 2 and 5 are unreachable)
          const bool ax = (n0 < n1) && (n0 < n2);
          const bool ay = !ax && (n1 < n2);
          const bool az = !ax && !ay;
          n0 = addSel(n0, d0, ax); n1 = addSel(n1, d1, ay); n2 = addSel(n2, d2, az);
          uint32_t c2 = c;
          addIfInt(c2, ix, ax); addIfInt(c2, iy, ay); addIfInt(c2, iz, az);
          {
            // The step left the mask word (or the brick, which shows in the guards): flush what was collected. Written as a
            // predicated RED with the word address computed by every lane: some lane of a warp leaves its word in almost
            // every step, so a divergent flush block would issue its address arithmetic (at a quarter of the lanes) nearly
            // every time anyway; this way it schedules between the fp64 instructions (measured on B200: -1.6 %).
            const bool left = ((c2 ^ c) & kCodeWordMask) != 0;
            markWordIf<MODE>(act_base + codeWord(c), acc, left);
            acc = left ? 0 : acc;
          }
          c = c2; // a step out of the brick has set a low guard bit: the lane now waits for the next batch point
          asm("ld.shared.u64 %0, [%1];" : "=l"(bit) : "r"((c & 0x707u) * 8u + bit_table));
        }
      }
    }
  }
}

// OR the privatised near-brick copies into the real bricks and leave the copies zeroed for the next scan. Eight
// threads per mask word, four copies each (word index fastest: coalesced); most words are zero, the few that are not are
// OR-ed into the real brick with a RED.
constexpr int kMergeSplit = 8;
__global__ void __launch_bounds__(256) merge_near_kernel(UpdateGrid g, uint64_t* near_act, int nbx, int nby, int nbz, Counters* ctr)
{
  __shared__ uint32_t s_slot[kNearBricks];
  if (threadIdx.x < kNearBricks)
    s_slot[threadIdx.x] = brickFindOrInsert(g, packLeafKey(nbx + (threadIdx.x >> 2), nby + ((threadIdx.x >> 1) & 1), nbz + (threadIdx.x & 1)), ctr);
  __syncthreads();
  constexpr uint32_t kWords = kNearBricks * kBrickLeaves * 8;
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t i = t % kWords, part = t / kWords; // word index in [0, 8 * 4096), copy group
  if (part >= kMergeSplit) return;
  uint64_t w[kNearCopies / kMergeSplit];
#pragma unroll
  for (int k = 0; k < kNearCopies / kMergeSplit; ++k) w[k] = near_act[size_t(part * (kNearCopies / kMergeSplit) + k) * kWords + i];
  uint64_t v = 0;
#pragma unroll
  for (int k = 0; k < kNearCopies / kMergeSplit; ++k)
    if (w[k]) { v |= w[k]; near_act[size_t(part * (kNearCopies / kMergeSplit) + k) * kWords + i] = 0; }
  const uint32_t slot = s_slot[i / (kBrickLeaves * 8)];
  if (v && slot != kInvalid) redOr64(g.act + size_t(slot) * (kBrickLeaves * 8) + (i % (kBrickLeaves * 8)), v);
}

// ====================================================================================================
// K1b: rebuild the compact list of touched leaves from the occupied bricks. One block per brick, one thread per
// leaf (512); a leaf is touched iff its 64-byte active mask is non-zero.
// COOK (after the DDA kernel, which marks z-slice words): the thread also turns its leaf's two masks into OpenVDB's
// x-slice layout, in place: x-slice word[x] bit (y<<3|z)  <->  z-slice word[z] bit (y<<3|x), i.e. for every byte lane y
// an 8x8 bit transpose between word index and bit-in-byte. The transform is an involution: uncook_leaves_kernel applies
// it again to the listed leaves before another scan is marked into a grid that already holds (cooked) data.
// ====================================================================================================
__device__ __forceinline__ void transposeLeafWords(uint64_t (&w)[8])
{
  uint32_t lo[8], hi[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { lo[i] = uint32_t(w[i]); hi[i] = uint32_t(w[i] >> 32); }
#pragma unroll
  for (int j = 4; j >= 1; j >>= 1)
  {
    const uint32_t m = (j == 4) ? 0x0F0F0F0Fu : (j == 2 ? 0x33333333u : 0x55555555u);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (!(i & j))
      {
        // swap the bits of word i whose in-byte index has bit j set with the bits of word i|j that have it clear
        uint32_t t = ((lo[i] >> j) ^ lo[i | j]) & m; lo[i | j] ^= t; lo[i] ^= t << j;
        t          = ((hi[i] >> j) ^ hi[i | j]) & m; hi[i | j] ^= t; hi[i] ^= t << j;
      }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) w[i] = (uint64_t(hi[i]) << 32) | lo[i];
}

__device__ __forceinline__ void loadLeafWords(const uint64_t* p, uint64_t (&w)[8])
{
  const ulonglong2* q = reinterpret_cast<const ulonglong2*>(p);
#pragma unroll
  for (int k = 0; k < 4; ++k) { const ulonglong2 v = q[k]; w[2 * k] = v.x; w[2 * k + 1] = v.y; }
}
__device__ __forceinline__ void storeLeafWords(uint64_t* p, const uint64_t (&w)[8])
{
  ulonglong2* q = reinterpret_cast<ulonglong2*>(p);
#pragma unroll
  for (int k = 0; k < 4; ++k) q[k] = make_ulonglong2(w[2 * k], w[2 * k + 1]);
}

template <bool COOK>
__global__ void __launch_bounds__(512) compact_leaves_kernel(UpdateGrid g)
{
  const int lane          = threadIdx.x & 31;
  const uint32_t n_bricks = min(g.counters[0], g.cap_mask + 1u); // device-side count: no host round trip needed
  for (uint32_t b = blockIdx.x; b < n_bricks; b += gridDim.x)
  {
    const uint32_t slot = g.btouched[b];
    const uint32_t e    = slot * kBrickLeaves + threadIdx.x;
    uint64_t a[8];
    loadLeafWords(g.act + size_t(e) * 8, a);
    const bool nz    = (a[0] | a[1] | a[2] | a[3] | a[4] | a[5] | a[6] | a[7]) != 0;
    const unsigned m = __ballot_sync(kFull, nz);
    if (m)
    {
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(g.counters + 1, (uint32_t)__popc(m));
      base = __shfl_sync(kFull, base, 0);
      if (nz) g.entries[base + __popc(m & ((1u << lane) - 1u))] = e;
    }
    if (COOK && nz)
    {
      transposeLeafWords(a);
      storeLeafWords(g.act + size_t(e) * 8, a);
      uint64_t v[8];
      loadLeafWords(g.val + size_t(e) * 8, v);
      if ((v[0] | v[1] | v[2] | v[3] | v[4] | v[5] | v[6] | v[7]) != 0)
      {
        transposeLeafWords(v);
        storeLeafWords(g.val + size_t(e) * 8, v);
      }
    }
  }
}

// the listed (cooked, x-slice) leaves back to z-slice words, before the DDA kernel marks into a non-empty grid
__global__ void __launch_bounds__(256) uncook_leaves_kernel(UpdateGrid g, uint32_t n_entries)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_entries) return;
  const uint32_t e = g.entries[i];
  uint64_t w[8];
  loadLeafWords(g.act + size_t(e) * 8, w);
  transposeLeafWords(w);
  storeLeafWords(g.act + size_t(e) * 8, w);
  loadLeafWords(g.val + size_t(e) * 8, w);
  if ((w[0] | w[1] | w[2] | w[3] | w[4] | w[5] | w[6] | w[7]) != 0)
  {
    transposeLeafWords(w);
    storeLeafWords(g.val + size_t(e) * 8, w);
  }
}

// ====================================================================================================
// K2: updateMap, in two kernels.
//  K2a resolve_leaves_kernel (one THREAD per touched leaf): entry -> leaf key -> map hash probe / insert. New leaves
//      are allocated with one warp-aggregated atomicAdd. Output: resolved[i] = map leaf index | kNewLeafBit, or
//      kInvalid when no leaf exists and none may be created. Hundreds of thousands of independent dependent-load
//      chains are in flight, so the probe latency is hidden by parallelism, not paid per warp.
//  K2b apply_update_kernel (one WARP per touched leaf): lane L owns the 16 consecutive voxels [16L, 16L+16) (= 64
//      contiguous bytes of leaf values, moved with two 256-bit loads/stores): a pure streaming read-modify-write of
//      the 2 KB map leaf; entry/resolved/mask words of the NEXT leaf are prefetched while the current one is processed.
// The update masks are zeroed as they are consumed (the grid is empty again afterwards, VDBMapping.hpp:384).
// ====================================================================================================
constexpr uint32_t kNewLeafBit = 0x80000000u;

// n_dev != nullptr: the number of entries is read on the device (deferred launch: the host never saw it), n is then
// only the bound the grid was sized for. All threads of a block run the same number of iterations (warp ballots inside).
__global__ void __launch_bounds__(256) resolve_leaves_kernel(UpdateGrid g, MapTable mt, LogOdds lo, uint32_t* resolved, Counters* ctr,
                                                            uint32_t n, const uint32_t* n_dev)
{
  if (n_dev) n = *n_dev;
  const int lane = threadIdx.x & 31;
  for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x)
  {
  const uint32_t i = base + threadIdx.x;
  uint32_t leaf = kInvalid, hslot = 0;
  int is_new    = 0;
  uint64_t key  = 0;
  if (i < n)
  {
    const uint32_t e = g.entries[i];
    key              = leafKeyOfEntry(g.bkeys[e >> 9], e & 511u);
    // OpenVDB tile probe: on a missing leaf a miss whose probe result is (0.0f, inactive) does not create it
    bool create_ok = true;
    if (lo.miss_probe_no_create)
    {
      uint64_t any_hit = 0;
#pragma unroll
      for (int w = 0; w < 8; ++w) any_hit |= g.val[size_t(e) * 8 + w];
      create_ok = any_hit != 0;
    }
    uint32_t h = uint32_t(mix64(key)) & mt.hcap_mask;
    for (uint32_t probe = 0; probe <= mt.hcap_mask; ++probe)
    {
      const uint64_t k = ldcg64(mt.hkeys + h);
      if (k == key) { leaf = mt.hvals[h]; break; }
      if (k == kEmptyKey)
      {
        if (!create_ok) break;
        // keys are unique within a launch, so nobody else inserts THIS key; another key may win this slot
        unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long*>(mt.hkeys + h), kEmptyKey, key);
        if (old == kEmptyKey) { is_new = 1; hslot = h; break; }
      }
      h = (h + 1) & mt.hcap_mask;
    }
  }
  const unsigned new_mask = __ballot_sync(kFull, is_new);
  if (new_mask)
  {
    uint32_t first = 0;
    if (lane == 0) first = atomicAdd(mt.n_leaves, (uint32_t)__popc(new_mask));
    first = __shfl_sync(kFull, first, 0);
    if (is_new)
    {
      const uint32_t li = first + __popc(new_mask & ((1u << lane) - 1u));
      if (li < mt.pool_cap)
      {
        mt.hvals[hslot]  = li;
        mt.leaf_keys[li] = key;
        leaf             = li;
      }
      else
      {
        atomicOr(&ctr->flags, kFlagMapOverflow); // host sizes the pool before the launch; cannot happen
        is_new = 0;
      }
    }
    if (lane == 0) atomicAdd(&ctr->new_leaves, (unsigned long long)__popc(new_mask));
  }
  if (i < n)
  {
    if (leaf != kInvalid) mt.leaf_dirty[leaf] = 1u;
    resolved[i] = (leaf == kInvalid) ? kInvalid : (leaf | (is_new ? kNewLeafBit : 0u));
  }
  }
}

struct LeafWork
{
  uint32_t e, r;    // entry, resolved leaf (uniform across the warp)
  uint64_t A, V, M; // update active word, update value word, old map active word (held by lanes 0..7)
};

__device__ __forceinline__ LeafWork loadLeafWork(const UpdateGrid& g, const MapTable& mt, const uint32_t* resolved, uint32_t i, int lane)
{
  LeafWork w;
  w.e = g.entries[i];
  w.r = resolved[i];
  w.A = w.V = w.M = 0;
  if (lane < 8)
  {
    w.A = g.act[size_t(w.e) * 8 + lane];
    w.V = g.val[size_t(w.e) * 8 + lane];
    if (w.r != kInvalid && !(w.r & kNewLeafBit)) w.M = mt.leaf_mask[size_t(w.r) * 8 + lane];
  }
  return w;
}

__global__ void __launch_bounds__(256, 4) apply_update_kernel(UpdateGrid g, MapTable mt, LogOdds lo, const uint32_t* resolved,
                                                             LeafRecord* change_out, uint32_t change_cap, Counters* ctr, uint32_t n,
                                                             const uint32_t* n_dev)
{
  if (n_dev) n = *n_dev;
  const int lane         = threadIdx.x & 31;
  const uint32_t warp    = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
  const int w  = lane >> 2;       // mask word of this lane's 16 voxels
  const int sh = (lane & 3) << 4; // bit offset inside the word
  unsigned upd_total = 0, chg_total = 0;

  LeafWork nxt{};
  if (warp < n) nxt = loadLeafWork(g, mt, resolved, warp, lane);
  for (uint32_t i = warp; i < n; i += n_warps)
  {
    const LeafWork cur = nxt;
    const uint32_t leaf = (cur.r == kInvalid) ? kInvalid : (cur.r & ~kNewLeafBit);
    const bool is_new   = (cur.r != kInvalid) && (cur.r & kNewLeafBit);
    float* vp = mt.leaf_vals + size_t(leaf == kInvalid ? 0 : leaf) * 512 + lane * 16;
    float v0[8], v1[8];
    // The mask words of this leaf were prefetched one iteration ago, so the update bits of my 16 voxels are known
    // without waiting: issue the leaf read first, and only for the 64-byte pieces that actually get an update
    // (a piece without update bits is neither read nor written), then prefetch the next leaf's small words.
    const uint32_t ua = uint32_t(shfl64(cur.A, w) >> sh) & 0xFFFFu; // active update bits of my 16 voxels
    if (leaf != kInvalid && !is_new && ua != 0) { ld256(vp, v0); ld256(vp + 8, v1); }
    else
    {
#pragma unroll
      for (int q = 0; q < 8; ++q) { v0[q] = 0.0f; v1[q] = 0.0f; }
    }
    if (i + n_warps < n) nxt = loadLeafWork(g, mt, resolved, i + n_warps, lane);
    // consume: leave the update masks clean for the next accumulation period
    if (lane < 8)
    {
      g.act[size_t(cur.e) * 8 + lane] = 0;
      g.val[size_t(cur.e) * 8 + lane] = 0;
    }
    const uint32_t uv = uint32_t(shfl64(cur.V, w) >> sh) & 0xFFFFu; // hit bits
    const unsigned nz_words  = __ballot_sync(kFull, cur.A != 0) & 0xFFu;
    const unsigned hit_words = __ballot_sync(kFull, cur.V != 0) & 0xFFu;

    uint32_t ca = 0, cv = 0; // change-grid bits of this lane's 16 voxels
    if (leaf == kInvalid)
    {
      // no leaf and none may be created: every (miss) voxel only runs the tile probe. With the quirk each of
      // them is reported when the probe flipped the inverted state (VDBMapping.hpp:743-750 via the lambda).
      if (lo.replicate_quirk && lo.miss_probe_flips) ca = ua & ~uv;
    }
    else
    {
      const uint32_t oa = uint32_t(shfl64(cur.M, w) >> sh) & 0xFFFFu;
      // OccupancyVDBMapping.hpp:92-117, branch-free (all 32 lanes stay converged whatever mix of hits/misses):
      //   nv = v + (hit ? logodds_hit : logodds_miss)
      //   P  = hit ? (nv > thres_max) : (nv < thres_min)              threshold crossed
      //   nv = P ? (hit ? min(nv, max_logodds) : max(nv, min_logodds)) : nv   (clamp only inside the threshold branch)
      //   active = P ? hit : old_active
      uint32_t pm = 0; // P bits of my 16 voxels
#pragma unroll
      for (int q = 0; q < 16; ++q)
      {
        const bool a_ = (ua >> q) & 1, h_ = (uv >> q) & 1;
        float& ref    = (q < 8) ? v0[q & 7] : v1[q & 7];
        const float nv  = __fadd_rn(ref, h_ ? lo.hit : lo.miss);
        const bool P    = h_ ? (nv > lo.thres_max) : (nv < lo.thres_min);
        const float cl  = h_ ? fminf(nv, lo.max_lo) : fmaxf(nv, lo.min_lo);
        const float nv2 = P ? cl : nv;
        ref = a_ ? nv2 : ref;
        pm |= uint32_t(P) << q;
      }
      // bit-parallel mask algebra on the 16 voxels
      const uint32_t act = (pm & uv) | (~pm & oa); // P ? hit : old
      const uint32_t na  = (ua & act) | (~ua & oa);
      const uint32_t chg = ua & (act ^ oa);        // VDBMapping.hpp:746-749: flag flipped
      ca |= chg;
      cv |= chg & uv;                              // :772 setValueOn(true) for hits, :780 setActiveState for misses
      if (is_new || ua != 0) { st256(vp, v0); st256(vp + 8, v1); } // untouched 64-byte pieces are not rewritten
      // assemble the new 64-bit active word from the 4 lanes that share it
      uint64_t piece = uint64_t(na) << sh;
      piece |= shflXor64(piece, 1);
      piece |= shflXor64(piece, 2);
      if ((lane & 3) == 0) mt.leaf_mask[size_t(leaf) * 8 + w] = piece;

      // tile-probe quirk: the first visited voxel (lowest offset) of a leaf that did not exist, if it
      // is a miss whose probe flips the inverted tile state, is reported as changed although its flag did not.
      if (is_new && lo.replicate_quirk && lo.miss_probe_flips)
      {
        if (!lo.miss_probe_no_create)
        {
          const int w_first       = __ffs(nz_words) - 1;
          const uint64_t a_first  = shfl64(cur.A, w_first), v_first = shfl64(cur.V, w_first);
          const int b_first       = __ffsll((long long)a_first) - 1;
          const int n_first       = (w_first << 6) | b_first;
          const bool first_is_hit = (v_first >> b_first) & 1;
          if (!first_is_hit && (n_first >> 4) == lane) ca |= 1u << (n_first & 15);
        }
        else
        {
          // degenerate config: misses do not create the leaf, so every miss BEFORE the first hit was probed
          const int wh      = __ffs(hit_words) - 1;
          const uint64_t vh = shfl64(cur.V, wh);
          const int n_hit   = (wh << 6) | (__ffsll((long long)vh) - 1);
          const int lo_n    = lane << 4;
          uint32_t before   = 0;
          if (n_hit >= lo_n + 16) before = 0xFFFFu;
          else if (n_hit > lo_n) before = (1u << (n_hit - lo_n)) - 1u;
          ca |= ua & ~uv & before;
        }
      }
    }
    upd_total += __popc(ua);
    chg_total += __popc(ca);

    if (change_out != nullptr)
    {
      const unsigned any = __ballot_sync(kFull, ca != 0);
      if (any)
      {
        uint64_t pa = uint64_t(ca) << sh, pv = uint64_t(cv) << sh;
        pa |= shflXor64(pa, 1); pa |= shflXor64(pa, 2);
        pv |= shflXor64(pv, 1); pv |= shflXor64(pv, 2);
        uint32_t ci = 0;
        if (lane == 0) ci = atomicAdd(&ctr->n_change, 1u);
        ci = __shfl_sync(kFull, ci, 0);
        if (ci < change_cap)
        {
          LeafRecord* rec = change_out + ci;
          if (lane == 0) rec->key = leafKeyOfEntry(g.bkeys[cur.e >> 9], cur.e & 511u);
          if ((lane & 3) == 0) { rec->active[w] = pa; rec->value[w] = pv; }
        }
      }
    }
  }
  // per-warp totals -> global counters
  upd_total = __reduce_add_sync(kFull, upd_total);
  chg_total = __reduce_add_sync(kFull, chg_total);
  if (lane == 0)
  {
    if (upd_total) atomicAdd(&ctr->voxel_updates, (unsigned long long)upd_total);
    if (chg_total) atomicAdd(&ctr->state_changes, (unsigned long long)chg_total);
  }
}

// ====================================================================================================
// growth / import / export helpers
// ====================================================================================================
// after the entries were consumed (their masks are zero again): forget the bricks
__global__ void reset_bricks_kernel(UpdateGrid g, uint32_t n_bricks, const uint32_t* n_dev)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (n_dev)
  {
    n_bricks = *n_dev; // deferred launch: 0 = the update was skipped, leave the grid alone
    if (n_bricks == 0) return;
  }
  if (i < n_bricks) g.bkeys[g.btouched[i]] = kEmptyKey;
  if (i == 0) { g.counters[0] = 0; g.counters[1] = 0; }
}

// Deferred updateMap (vdbm_insert_async): the host queued the update kernels without knowing how many leaves the raycast
// touched. This one-thread kernel decides ON THE DEVICE whether they may run: no status flag raised, the brick hash not
// crowded, and room for every touched leaf in `resolved`, the leaf pool and the map hash. It publishes the counts the
// following kernels read; zero counts turn them into no-ops and the host redoes the scan synchronously.
__global__ void update_guard_kernel(UpdateGrid g, MapTable mt, Counters* ctr, uint32_t resolved_cap)
{
  const uint32_t n_bricks = g.counters[0], n_entries = g.counters[1];
  const uint64_t need = uint64_t(*mt.n_leaves) + n_entries;
  const bool ok = ctr->flags == 0 && // any pending condition (overflow, out-of-range points, exchange trouble) goes to the host first
                  uint64_t(n_bricks) * 10 <= (uint64_t(g.cap_mask) + 1) * 7 && n_entries <= resolved_cap && need <= mt.pool_cap &&
                  need * 2 <= uint64_t(mt.hcap_mask) + 1;
  ctr->deferred_entries = ok ? n_entries : 0u;
  ctr->deferred_bricks  = ok ? n_bricks : 0u;
  ctr->deferred_skip    = ok ? 0u : 1u;
  ctr->deferred_seen_entries = n_entries;
}

__global__ void clear_entries_kernel(UpdateGrid g, uint32_t n_entries)
{
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t i = t >> 4;
  const int j      = t & 15;
  if (i >= n_entries) return;
  const uint32_t e = g.entries[i];
  if (j < 8) g.act[size_t(e) * 8 + j] = 0;
  else g.val[size_t(e) * 8 + (j - 8)] = 0;
}

// copy every occupied brick of the old grid into the (bigger) new grid; one block per brick
__global__ void __launch_bounds__(256) rehash_update_kernel(UpdateGrid old_g, uint32_t old_bricks, UpdateGrid new_g, Counters* ctr)
{
  __shared__ uint32_t s_slot;
  for (uint32_t b = blockIdx.x; b < old_bricks; b += gridDim.x)
  {
    const uint32_t os = old_g.btouched[b];
    if (threadIdx.x == 0) s_slot = brickFindOrInsert(new_g, old_g.bkeys[os], ctr);
    __syncthreads();
    const uint32_t ns = s_slot;
    if (ns != kInvalid)
    {
      const ulonglong2* sa = reinterpret_cast<const ulonglong2*>(old_g.act + size_t(os) * (kBrickLeaves * 8));
      const ulonglong2* sv = reinterpret_cast<const ulonglong2*>(old_g.val + size_t(os) * (kBrickLeaves * 8));
      ulonglong2* da       = reinterpret_cast<ulonglong2*>(new_g.act + size_t(ns) * (kBrickLeaves * 8));
      ulonglong2* dv       = reinterpret_cast<ulonglong2*>(new_g.val + size_t(ns) * (kBrickLeaves * 8));
      for (int k = threadIdx.x; k < kBrickLeaves * 4; k += blockDim.x) { da[k] = sa[k]; dv[k] = sv[k]; }
    }
    __syncthreads();
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

__global__ void import_update_kernel(UpdateGrid g, const LeafRecord* recs, uint64_t n, Counters* ctr)
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
  const unsigned grp = 0xFFFFu << (threadIdx.x & 16);
  const unsigned nz  = __ballot_sync(kFull, valid && j < 8 && word != 0) & grp;
  uint32_t slot = kInvalid, lib = 0;
  if (valid && nz && j == 0)
  {
    const uint64_t bkey = brickKeyOfLeaf(key, lib);
    slot                = brickFindOrInsert(g, bkey, ctr);
  }
  slot = __shfl_sync(kFull, slot, (threadIdx.x & 16));
  lib  = __shfl_sync(kFull, lib, (threadIdx.x & 16));
  if (slot == kInvalid || word == 0) return;
  const size_t e = size_t(slot) * kBrickLeaves + lib;
  uint64_t* dst  = (j < 8) ? g.act + e * 8 + j : g.val + e * 8 + (j - 8);
  redOr64(dst, word);
}

// same, from host-layout leaf arrays (origins [n][3], active [n][8], value [n][8]) copied to the device as they are
__global__ void import_update_soa_kernel(UpdateGrid g, const int32_t* origins, const uint64_t* active, const uint64_t* value, uint64_t n,
                                         Counters* ctr)
{
  const uint64_t t   = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const uint64_t rec = t >> 4;
  const int j        = int(t & 15);
  const bool valid   = rec < n;
  uint64_t word = 0;
  if (valid) word = (j < 8) ? active[rec * 8 + j] : (value ? value[rec * 8 + (j - 8)] : 0ull);
  const unsigned grp = 0xFFFFu << (threadIdx.x & 16);
  const unsigned nz  = __ballot_sync(kFull, valid && j < 8 && word != 0) & grp;
  uint32_t slot = kInvalid, lib = 0;
  if (valid && nz && j == 0)
  {
    const int32_t x = origins[rec * 3], y = origins[rec * 3 + 1], z = origins[rec * 3 + 2];
    if (max(abs(x), max(abs(y), abs(z))) >= kVoxelLimit) atomicOr(&ctr->flags, kFlagCoordRange);
    else
    {
      const uint64_t bkey = brickKeyOfLeaf(packLeafKey(x >> 3, y >> 3, z >> 3), lib);
      slot                = brickFindOrInsert(g, bkey, ctr);
    }
  }
  slot = __shfl_sync(kFull, slot, (threadIdx.x & 16));
  lib  = __shfl_sync(kFull, lib, (threadIdx.x & 16));
  if (slot == kInvalid || word == 0) return;
  const size_t e = size_t(slot) * kBrickLeaves + lib;
  uint64_t* dst  = (j < 8) ? g.act + e * 8 + j : g.val + e * 8 + (j - 8);
  redOr64(dst, word);
}

// leaf keys of the listed entries (+ the entry ids, for sorting)
__global__ void entry_keys_kernel(UpdateGrid g, uint32_t n, uint64_t* out_keys, uint32_t* out_entries)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t e = g.entries[i];
  out_keys[i]      = leafKeyOfEntry(g.bkeys[e >> 9], e & 511u);
  out_entries[i]   = e;
}

__global__ void keys_from_idx_kernel(const uint64_t* keys, const uint32_t* idx, uint32_t n, uint64_t* out_keys, uint32_t* out_idx)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t s = idx ? idx[i] : i;
  out_keys[i]      = keys[s];
  out_idx[i]       = s;
}

// update-grid leaves (sorted keys + their entries) -> records
__global__ void gather_update_kernel(UpdateGrid g, uint32_t n, const uint64_t* keys, const uint32_t* entries, LeafRecord* out)
{
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t i = t >> 4;
  const int j      = t & 15;
  if (i >= n) return;
  const uint32_t e = entries[i];
  if (j < 8) out[i].active[j] = g.act[size_t(e) * 8 + j];
  else out[i].value[j - 8] = g.val[size_t(e) * 8 + (j - 8)];
  if (j == 0) out[i].key = keys[i];
}

__global__ void split_records_kernel(const LeafRecord* recs, const uint32_t* perm, uint32_t n, int32_t* origins, uint64_t* active,
                                     uint64_t* value)
{
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t i = t >> 4;
  const int j      = t & 15;
  if (i >= n) return;
  const LeafRecord& r = recs[perm ? perm[i] : i];
  if (j < 8) active[size_t(i) * 8 + j] = r.active[j];
  else value[size_t(i) * 8 + (j - 8)] = r.value[j - 8];
  if (j == 0)
  {
    int32_t x, y, z;
    unpackLeafOrigin(r.key, x, y, z);
    origins[3 * size_t(i)] = x; origins[3 * size_t(i) + 1] = y; origins[3 * size_t(i) + 2] = z;
  }
}

__global__ void record_keys_kernel(const LeafRecord* recs, uint32_t n, uint64_t* keys, uint32_t* idx)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  keys[i] = recs[i].key;
  idx[i]  = i;
}

// section rows (written in arrival order by section_kernel) -> rows in sorted-key order + leaf origins. Warp per leaf.
__global__ void __launch_bounds__(256) permute_section_kernel(uint32_t n, const uint64_t* sorted_keys, const uint32_t* perm, const uint64_t* in_active,
                                                             const uint64_t* in_valmask, const float* in_vals, int32_t* origins,
                                                             uint64_t* out_active, uint64_t* out_valmask, float* out_vals)
{
  const int lane         = threadIdx.x & 31;
  const uint32_t warp    = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t i = warp; i < n; i += n_warps)
  {
    const uint32_t j = perm[i];
    if (lane < 8)
    {
      out_active[size_t(i) * 8 + lane] = in_active[size_t(j) * 8 + lane];
      if (in_valmask) out_valmask[size_t(i) * 8 + lane] = in_valmask[size_t(j) * 8 + lane];
    }
    if (in_vals)
    {
      float a[8], b[8];
      ld256(in_vals + size_t(j) * 512 + lane * 16, a);
      ld256(in_vals + size_t(j) * 512 + lane * 16 + 8, b);
      st256(out_vals + size_t(i) * 512 + lane * 16, a);
      st256(out_vals + size_t(i) * 512 + lane * 16 + 8, b);
    }
    if (lane == 0)
    {
      int32_t x, y, z;
      unpackLeafOrigin(sorted_keys[i], x, y, z);
      origins[3 * size_t(i)] = x; origins[3 * size_t(i) + 1] = y; origins[3 * size_t(i) + 2] = z;
    }
  }
}

__global__ void __launch_bounds__(256) gather_map_kernel(MapTable mt, uint32_t n, const uint32_t* leaf_idx, int32_t* origins,
                                                        uint64_t* mask, float* vals)
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
    }
  }
}

// dirty leaves -> index list (warp-aggregated append through ctr->n_out); flags are cleared
__global__ void collect_dirty_kernel(MapTable mt, uint32_t n_leaves, uint32_t* out_idx, Counters* ctr)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane   = threadIdx.x & 31;
  bool d = false;
  if (i < n_leaves)
  {
    d = mt.leaf_dirty[i] != 0;
    if (d) mt.leaf_dirty[i] = 0;
  }
  const unsigned m = __ballot_sync(kFull, d);
  if (!m) return;
  uint32_t base = 0;
  if (lane == 0) base = atomicAdd(&ctr->n_out, (uint32_t)__popc(m));
  base = __shfl_sync(kFull, base, 0);
  if (d) out_idx[base + __popc(m & ((1u << lane) - 1u))] = i;
}

// vdbm_map_mirror aborted by its consumer: the leaves that were not delivered become dirty again
__global__ void mark_dirty_kernel(MapTable mt, const uint32_t* idx, uint32_t n)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) mt.leaf_dirty[idx[i]] = 1u;
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

// multi-GPU: bin touched update leaves by owner rank. pass 0 counts, pass 1 scatters (and zeroes the entry masks).
__global__ void partition_kernel(UpdateGrid g, uint32_t n, ShardPlan plan, uint32_t* rank_counts, uint32_t* rank_cursor,
                                 LeafRecord* out, int pass)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t e   = g.entries[i];
  const uint64_t key = leafKeyOfEntry(g.bkeys[e >> 9], e & 511u);
  const int32_t r    = leafOwnerPlanned(key, plan);
  if (pass == 0) { atomicAdd(rank_counts + r, 1u); return; }
  const uint32_t dst = atomicAdd(rank_cursor + r, 1u);
  out[dst].key       = key;
#pragma unroll
  for (int w = 0; w < 8; ++w)
  {
    out[dst].active[w] = g.act[size_t(e) * 8 + w];
    out[dst].value[w]  = g.val[size_t(e) * 8 + w];
    g.act[size_t(e) * 8 + w] = 0;
    g.val[size_t(e) * 8 + w] = 0;
  }
}

// ====================================================================================================
// Receiver side of remote mapping: applyMapSectionUpdateGrid / applyMapSectionGrid
// ====================================================================================================
// 64-bit in-box mask of mask word w (x = ox + w) of the leaf at (ox, oy, oz) for the inclusive box
__device__ __forceinline__ uint64_t inBoxWord(int xx, int oy, int oz, int bx0, int by0, int bz0, int bx1, int by1, int bz1)
{
  if (xx < bx0 || xx > bx1) return 0;
  uint64_t zb = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) zb |= uint64_t(oz + k >= bz0 && oz + k <= bz1) << k;
  uint64_t m = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (oy + j >= by0 && oy + j <= by1) m |= zb << (8 * j);
  return m;
}

// VDBMapping.hpp:1073-1079: every active map voxel inside the box is deactivated (values stay). Thread per (leaf, word).
__global__ void section_deactivate_kernel(MapTable mt, uint32_t n_leaves, int bx0, int by0, int bz0, int bx1, int by1, int bz1)
{
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t l = t >> 3;
  const int w      = t & 7;
  if (l >= n_leaves) return;
  int32_t ox, oy, oz;
  unpackLeafOrigin(mt.leaf_keys[l], ox, oy, oz);
  if (ox + 7 < bx0 || ox > bx1 || oy + 7 < by0 || oy > by1 || oz + 7 < bz0 || oz > bz1) return;
  const uint64_t in  = inBoxWord(ox + w, oy, oz, bx0, by0, bz0, bx1, by1, bz1);
  const uint64_t old = mt.leaf_mask[size_t(l) * 8 + w];
  if (old & in)
  {
    mt.leaf_mask[size_t(l) * 8 + w] = old & ~in;
    mt.leaf_dirty[l]                = 1u;
  }
}

// find (or create, if `create`) the map leaf of `key`; lane 0 of the warp does the probe. Returns leaf index / kInvalid.
__device__ __forceinline__ uint32_t warpFindOrCreateLeaf(MapTable& mt, uint64_t key, bool create, int lane, int& is_new, Counters* ctr)
{
  uint32_t leaf = kInvalid;
  is_new        = 0;
  if (lane == 0)
  {
    uint32_t h = uint32_t(mix64(key)) & mt.hcap_mask;
    for (uint32_t probe = 0; probe <= mt.hcap_mask; ++probe)
    {
      const uint64_t k = ldcg64(mt.hkeys + h);
      if (k == key) { leaf = mt.hvals[h]; break; }
      if (k == kEmptyKey)
      {
        if (!create) break;
        unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long*>(mt.hkeys + h), kEmptyKey, key);
        if (old == kEmptyKey)
        {
          const uint32_t li = atomicAdd(mt.n_leaves, 1u);
          if (li >= mt.pool_cap) { atomicOr(&ctr->flags, kFlagMapOverflow); break; }
          mt.hvals[h]      = li;
          mt.leaf_keys[li] = key;
          leaf             = li;
          is_new           = 1;
          atomicAdd(&ctr->new_leaves, 1ull);
          break;
        }
      }
      h = (h + 1) & mt.hcap_mask;
    }
  }
  leaf   = __shfl_sync(kFull, leaf, 0);
  is_new = __shfl_sync(kFull, is_new, 0);
  return leaf;
}

// VDBMapping.hpp:1080-1083: setActiveState(coord, true) for every active section voxel (a missing leaf is created with
// background values). Warp per section leaf (keys are unique).
__global__ void __launch_bounds__(256) section_activate_kernel(MapTable mt, const uint64_t* keys, const uint64_t* active, uint32_t n, Counters* ctr)
{
  const int lane         = threadIdx.x & 31;
  const uint32_t warp    = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t i = warp; i < n; i += n_warps)
  {
    uint64_t a = (lane < 8) ? active[size_t(i) * 8 + lane] : 0;
    if (!(__ballot_sync(kFull, a != 0) & 0xFFu)) continue;
    int is_new;
    const uint32_t leaf = warpFindOrCreateLeaf(mt, keys[i], true, lane, is_new, ctr);
    if (leaf == kInvalid) continue;
    if (is_new)
    {
      float z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      st256(mt.leaf_vals + size_t(leaf) * 512 + lane * 16, z);
      st256(mt.leaf_vals + size_t(leaf) * 512 + lane * 16 + 8, z);
    }
    if (lane < 8) mt.leaf_mask[size_t(leaf) * 8 + lane] = (is_new ? 0ull : mt.leaf_mask[size_t(leaf) * 8 + lane]) | a;
    if (lane == 0) mt.leaf_dirty[leaf] = 1u;
  }
}

// VDBMapping.hpp:1034-1045 for the voxels of the section LEAVES: the section leaf replaces the map leaf voxel for voxel
// (value and state). A missing map leaf is only created when the section leaf holds an active voxel or a non-background value.
__global__ void __launch_bounds__(256) section_apply_grid_kernel(MapTable mt, const uint64_t* keys, const uint64_t* active, const float* values,
                                                                uint32_t n, Counters* ctr)
{
  const int lane         = threadIdx.x & 31;
  const uint32_t warp    = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t i = warp; i < n; i += n_warps)
  {
    float a[8], b[8];
    ld256(values + size_t(i) * 512 + lane * 16, a);
    ld256(values + size_t(i) * 512 + lane * 16 + 8, b);
    const uint64_t m = (lane < 8) ? active[size_t(i) * 8 + lane] : 0;
    bool content = (m != 0);
#pragma unroll
    for (int q = 0; q < 8; ++q) content = content || a[q] != 0.0f || b[q] != 0.0f;
    const bool any_content = __any_sync(kFull, content);
    int is_new;
    const uint32_t leaf = warpFindOrCreateLeaf(mt, keys[i], any_content, lane, is_new, ctr);
    if (leaf == kInvalid) continue;
    st256(mt.leaf_vals + size_t(leaf) * 512 + lane * 16, a);
    st256(mt.leaf_vals + size_t(leaf) * 512 + lane * 16 + 8, b);
    if (lane < 8) mt.leaf_mask[size_t(leaf) * 8 + lane] = m;
    if (lane == 0) mt.leaf_dirty[leaf] = 1u;
  }
}

__device__ __forceinline__ bool sortedContains(const uint64_t* a, uint32_t n, uint64_t key)
{
  uint32_t lo = 0, hi = n;
  while (lo < hi)
  {
    const uint32_t mid = (lo + hi) >> 1;
    const uint64_t v   = a[mid];
    if (v == key) return true;
    if (v < key) lo = mid + 1;
    else hi = mid;
  }
  return false;
}

__device__ __forceinline__ void clearVoxel0IfLeafExists(MapTable& mt, uint64_t key)
{
  uint32_t h = uint32_t(mix64(key)) & mt.hcap_mask;
  for (uint32_t probe = 0; probe <= mt.hcap_mask; ++probe)
  {
    const uint64_t k = mt.hkeys[h];
    if (k == kEmptyKey) return;
    if (k == key)
    {
      const uint32_t l = mt.hvals[h];
      mt.leaf_vals[size_t(l) * 512] = 0.0f;                // setValueOff(tile origin, background): value 0 ...
      atomicAnd(reinterpret_cast<unsigned long long*>(mt.leaf_mask + size_t(l) * 8), ~1ull); // ... and inactive
      mt.leaf_dirty[l] = 1u;
      return;
    }
    h = (h + 1) & mt.hcap_mask;
  }
}

// The value-all iteration of the reference also visits the inactive background TILES of the section tree's internal
// nodes (VDBMapping.hpp:1034 cbeginValueAll): map.setValueOff(tile origin, 0) clears that one voxel wherever the map
// has a leaf. level 0: the 16^3 leaf slots of every 128^3 block (I1 node) that holds a section leaf;
// level 1: the 32^3 I1 slots of every 4096^3 block (I2 node). `blocks` = sorted unique block keys (leaf-key format).
__global__ void section_tile_quirk_kernel(MapTable mt, const uint64_t* blocks, uint32_t n_blocks, int level, const uint64_t* present,
                                          uint32_t n_present)
{
  const uint32_t per   = level == 0 ? 4096u : 32768u;
  const uint64_t t     = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= uint64_t(n_blocks) * per) return;
  const uint32_t bi = uint32_t(t / per), slot = uint32_t(t % per);
  int32_t ox, oy, oz;
  unpackLeafOrigin(blocks[bi], ox, oy, oz);
  int32_t x, y, z;
  if (level == 0) { x = ox + int32_t(slot >> 8) * 8; y = oy + int32_t((slot >> 4) & 15) * 8; z = oz + int32_t(slot & 15) * 8; }
  else            { x = ox + int32_t(slot >> 10) * 128; y = oy + int32_t((slot >> 5) & 31) * 128; z = oz + int32_t(slot & 31) * 128; }
  const uint64_t key = packLeafKey(x >> 3, y >> 3, z >> 3);
  // level 0: slots holding a section leaf are not tiles; level 1: slots holding an I1 node are not tiles
  if (sortedContains(present, n_present, key)) return;
  clearVoxel0IfLeafExists(mt, key);
}

// ====================================================================================================
// Multi-GPU: fused bin-and-send over peer memory (NVLink). A block takes 256 touched update leaves:
//   A  every thread finds the owner rank of its leaf and reserves a slot in a per-block, per-owner histogram (smem)
//   B  one global atomicAdd per (block, owner) reserves the block's range in that owner's sender region
//   C  16 lanes per record gather the 17 words from the brick masks and store them into the OWNER'S inbox
//      (a mapped peer pointer: the stores go over NVLink); the update masks are zeroed as they are read
// publish_counts_kernel then releases (epoch << 32 | count) to every peer with system scope.
// ====================================================================================================
__global__ void __launch_bounds__(256) push_update_kernel(UpdateGrid g, uint32_t n, ExchangePeers px, ShardPlan plan, uint32_t parity,
                                                         uint32_t* cursors, Counters* ctr)
{
  __shared__ uint32_t s_count[kMaxRanks], s_base[kMaxRanks];
  __shared__ uint32_t s_entry[256], s_dst[256]; // entry id, (owner << 24 | local slot); kInvalid = the leaf is this rank's own
  __shared__ uint64_t s_key[256];
  const uint32_t i0 = blockIdx.x * 256u;
  if (threadIdx.x < kMaxRanks) s_count[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t i = i0 + threadIdx.x;
  if (i < n)
  {
    const uint32_t e   = g.entries[i];
    const uint64_t key = leafKeyOfEntry(g.bkeys[e >> 9], e & 511u);
    const uint32_t r   = uint32_t(leafOwnerPlanned(key, plan));
    s_entry[threadIdx.x] = e;
    s_key[threadIdx.x]   = key;
    // a leaf this rank owns STAYS in its update grid (the pull ORs the peers' records on top): with sector ownership
    // that is nearly every touched leaf, and nothing of it is copied
    s_dst[threadIdx.x] = (r == uint32_t(px.rank)) ? kInvalid : ((r << 24) | atomicAdd(&s_count[r], 1u));
  }
  __syncthreads();
  if (threadIdx.x < px.n_ranks) s_base[threadIdx.x] = s_count[threadIdx.x] ? atomicAdd(cursors + threadIdx.x, s_count[threadIdx.x]) : 0u;
  __syncthreads();
  const uint32_t n_here = min(256u, n - i0);
  // 16 lanes per record, 16 records per pass, 16 passes. All loads of a thread are issued before its first store
  // (otherwise every pass exposes one full load latency: the stores may alias the next loads for the compiler).
  const uint32_t j = threadIdx.x & 15u, k0 = threadIdx.x >> 4;
  const uint32_t region = parity * uint32_t(px.n_ranks) + uint32_t(px.rank);
  uint64_t w[16];
#pragma unroll
  for (int it = 0; it < 16; ++it)
  {
    const uint32_t k = k0 + 16u * it;
    w[it] = 0;
    if (k < n_here && s_dst[k] != kInvalid)
    {
      const uint32_t e = s_entry[k];
      w[it] = (j < 8) ? g.act[size_t(e) * 8 + j] : g.val[size_t(e) * 8 + (j - 8)];
    }
  }
#pragma unroll
  for (int it = 0; it < 16; ++it)
  {
    const uint32_t k = k0 + 16u * it;
    if (k >= n_here || s_dst[k] == kInvalid) continue;
    const uint32_t e = s_entry[k], r = s_dst[k] >> 24, pos = s_base[r] + (s_dst[k] & 0xFFFFFFu);
    if (j < 8) g.act[size_t(e) * 8 + j] = 0;
    else g.val[size_t(e) * 8 + (j - 8)] = 0;
    if (pos >= px.cap) { if (j == 0) atomicOr(&ctr->flags, kFlagExchangeOverflow); continue; }
    // 16 lanes x 8 bytes = one aligned 128-byte line in the owner's inbox
    px.inbox[r][inboxMaskWord(region, px.cap, pos) + j] = w[it];
  }
  // keys: thread t writes the key of record t (records of one owner are contiguous per block -> coalesced runs)
  if (threadIdx.x < n_here && s_dst[threadIdx.x] != kInvalid)
  {
    const uint32_t r = s_dst[threadIdx.x] >> 24, pos = s_base[r] + (s_dst[threadIdx.x] & 0xFFFFFFu);
    if (pos < px.cap)
      px.inbox[r][inboxKeyWord(2u * uint32_t(px.n_ranks), parity * uint32_t(px.n_ranks) + uint32_t(px.rank), px.cap, pos)] = s_key[threadIdx.x];
  }
}

// bit 31 of the published count: this sender could not fit all its records into the receiver's region. The records that
// did not fit are gone (their masks were consumed), so EVERY receiver of this epoch must fail loudly, not only the sender.
constexpr uint32_t kExchangeOverflowBit = 0x80000000u;

__global__ void publish_counts_kernel(ExchangePeers px, uint32_t parity, uint32_t epoch, const uint32_t* cursors)
{
  const int r = threadIdx.x;
  if (r >= px.n_ranks) return;
  const uint32_t sent = cursors[r];
  const unsigned long long word = ((unsigned long long)epoch << 32) | min(sent, px.cap) | (sent > px.cap ? kExchangeOverflowBit : 0u);
  __threadfence_system(); // the records written by push_update_kernel (previous launch on this stream) come first
  unsigned long long* p = px.ctrl[r] + size_t(parity) * px.n_ranks + px.rank;
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(word) : "memory");
}

// single warp: wait (bounded by timeout_ns of the global timer) until every sender has published this epoch, then leave the
// counts in device memory. A sender that ran out of inbox space raises kFlagExchangeOverflow HERE as well.
__global__ void wait_peers_kernel(const unsigned long long* ctrl, int32_t n_ranks, uint32_t parity, uint32_t epoch, uint32_t* counts_out,
                                  Counters* ctr, unsigned long long timeout_ns)
{
  const int s = threadIdx.x;
  if (s >= n_ranks) return;
  const unsigned long long* p = ctrl + size_t(parity) * n_ranks + s;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  unsigned long long w = 0;
  for (;;)
  {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    if (uint32_t(w >> 32) == epoch) break;
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > timeout_ns) { atomicOr(&ctr->flags, kFlagExchangeTimeout); w = 0; break; }
    __nanosleep(200);
  }
  if (uint32_t(w) & kExchangeOverflowBit) atomicOr(&ctr->flags, kFlagExchangeOverflow);
  counts_out[s] = uint32_t(w) & ~kExchangeOverflowBit;
}

__global__ void __launch_bounds__(256) pull_update_kernel(UpdateGrid g, const uint64_t* inbox, uint32_t cap, int32_t n_ranks, uint32_t parity,
                                                         const uint32_t* counts, Counters* ctr)
{
  // 16 lanes per record, kU records per pass and group (all loads of a pass are issued before the first RED, so a pass
  // exposes one load latency, not kU), grid-stride over all sender regions of this parity
  constexpr int kU = 4;
  const uint32_t group = (blockIdx.x * blockDim.x + threadIdx.x) >> 4, n_groups = (gridDim.x * blockDim.x) >> 4;
  const int j = threadIdx.x & 15;
  for (int s = 0; s < n_ranks; ++s)
  {
    const uint32_t cnt    = counts[s];
    const uint32_t region = parity * uint32_t(n_ranks) + uint32_t(s);
    // all 32 lanes of a warp run the same number of passes (the ballot below needs the full warp)
    const uint32_t passes = (cnt + n_groups * kU - 1) / (n_groups * kU);
    for (uint32_t it = 0; it < passes; ++it)
    {
      uint64_t word[kU], key[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u)
      {
        const uint32_t rec = (it * kU + u) * n_groups + group;
        word[u] = 0; key[u] = 0;
        if (rec < cnt)
        {
          word[u] = inbox[inboxMaskWord(region, cap, rec) + j];
          if (j == 0) key[u] = inbox[inboxKeyWord(2u * uint32_t(n_ranks), region, cap, rec)];
        }
      }
#pragma unroll
      for (int u = 0; u < kU; ++u)
      {
        const uint32_t rec = (it * kU + u) * n_groups + group;
        const bool valid   = rec < cnt;
        const unsigned grp = 0xFFFFu << (threadIdx.x & 16);
        const unsigned nz  = __ballot_sync(kFull, valid && j < 8 && word[u] != 0) & grp;
        uint32_t slot = kInvalid, lib = 0;
        if (valid && nz && j == 0)
        {
          const uint64_t bkey = brickKeyOfLeaf(key[u], lib);
          slot                = brickFindOrInsert(g, bkey, ctr);
        }
        slot = __shfl_sync(kFull, slot, (threadIdx.x & 16));
        lib  = __shfl_sync(kFull, lib, (threadIdx.x & 16));
        if (slot == kInvalid || word[u] == 0) continue;
        const size_t e = size_t(slot) * kBrickLeaves + lib;
        redOr64((j < 8) ? g.act + e * 8 + j : g.val + e * 8 + (j - 8), word[u]);
      }
    }
  }
}

// ====================================================================================================
// Remote-mapping deltas (createUpdate / applyUpdate) and direct map edits (8f N4)
// ====================================================================================================
// end-voxel records (x, y, z, bit0 valid | bit1 hit) -> bits of a bool grid: active, and value where hit
__global__ void __launch_bounds__(256) mark_ends_kernel(const int4* ends, uint64_t n, UpdateGrid g, Counters* ctr)
{
  const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int4 q = ends[i];
  if (!(q.w & 1)) return;
  const uint64_t bkey = packLeafKey(q.x >> 6, q.y >> 6, q.z >> 6);
  const uint32_t slot = brickFindOrInsert(g, bkey, ctr);
  if (slot == kInvalid) return;
  const size_t w     = size_t(slot) * (kBrickLeaves * 8) + brickWordOffset(q.x, q.y, q.z);
  const uint64_t bit = uint64_t(1) << (((q.y & 7) << 3) | (q.z & 7));
  redOr64(g.act + w, bit);
  if (q.w & 2) redOr64(g.val + w, bit);
}

// the active voxels of a bool grid -> end-voxel records (any order). Warp per touched leaf, lane w < 8 owns mask word w.
__global__ void __launch_bounds__(256) expand_ends_kernel(UpdateGrid g, uint32_t n_entries, int4* out, uint32_t out_cap, Counters* ctr)
{
  const int lane         = threadIdx.x & 31;
  const uint32_t warp    = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t i = warp; i < n_entries; i += n_warps)
  {
    const uint32_t e = g.entries[i];
    uint64_t a = 0, v = 0;
    if (lane < 8) { a = g.act[size_t(e) * 8 + lane]; v = g.val[size_t(e) * 8 + lane]; }
    const uint32_t cnt = uint32_t(__popcll(a));
    uint32_t incl = cnt; // inclusive prefix sum over the warp
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
      const uint32_t t = __shfl_up_sync(kFull, incl, d);
      if (lane >= d) incl += t;
    }
    const uint32_t total = __shfl_sync(kFull, incl, 31);
    uint32_t base = 0;
    if (lane == 0 && total) base = atomicAdd(&ctr->n_out, total);
    base = __shfl_sync(kFull, base, 0);
    if (!cnt) continue;
    int32_t ox, oy, oz;
    unpackLeafOrigin(leafKeyOfEntry(g.bkeys[e >> 9], e & 511u), ox, oy, oz);
    uint32_t o = base + incl - cnt;
    while (a)
    {
      const int b = __ffsll((long long)a) - 1;
      a &= a - 1;
      if (o < out_cap) out[o] = make_int4(ox + lane, oy + (b >> 3), oz + (b & 7), 1 | (((v >> b) & 1ull) ? 2 : 0));
      ++o;
    }
  }
}

// pcl points -> end-voxel records for addPointsToGrid / removePointsFromGrid (VDBMapping.hpp:413-447):
// Coord::floor(grid->worldToIndex(pt)) = floor(pt * (1/res)) per component, NOT the fmod variant of V:612-631
__global__ void __launch_bounds__(256) points_to_ends_kernel(const uint8_t* points, uint64_t n, uint32_t stride, double inv_res, int occupied,
                                                            int4* out, Counters* ctr)
{
  const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = reinterpret_cast<const float*>(points + i * stride);
  int4 q         = make_int4(0, 0, 0, 0);
  bool ok        = isfinite(p[0]) && isfinite(p[1]) && isfinite(p[2]); // non-finite points are undefined in the reference
  bool in_range  = true;
#pragma unroll
  for (int k = 0; k < 3; ++k)
  {
    const double fl = floor(__dmul_rn(double(p[k]), inv_res));
    if (!(fabs(fl) < double(kVoxelLimit))) in_range = false;
    (&q.x)[k] = in_range ? int32_t(fl) : 0;
  }
  if (ok && !in_range) atomicOr(&ctr->flags, kFlagCoordRange);
  q.w    = (ok && in_range) ? (1 | (occupied ? 2 : 0)) : 0;
  out[i] = q;
}

// overwriteMap / setNodeToOccupied / setNodeToFree (OccupancyVDBMapping.hpp:118-129) for every active voxel of a bool
// grid: value bit -> (max_logodds, active), else (min_logodds, inactive). Both ops change the value of a background
// tile, so OpenVDB's tile probe always creates the leaf. Warp per touched leaf, lane = 16 consecutive voxels.
__global__ void __launch_bounds__(256) overwrite_kernel(UpdateGrid g, uint32_t n_entries, MapTable mt, LogOdds lo, Counters* ctr)
{
  const int lane         = threadIdx.x & 31;
  const uint32_t warp    = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t i = warp; i < n_entries; i += n_warps)
  {
    const uint32_t e = g.entries[i];
    uint64_t a = 0, v = 0;
    if (lane < 8) { a = g.act[size_t(e) * 8 + lane]; v = g.val[size_t(e) * 8 + lane]; }
    if (!(__ballot_sync(kFull, a != 0) & 0xFFu)) continue;
    int is_new;
    const uint32_t leaf = warpFindOrCreateLeaf(mt, leafKeyOfEntry(g.bkeys[e >> 9], e & 511u), true, lane, is_new, ctr);
    if (leaf == kInvalid) continue;
    float x[8] = {0, 0, 0, 0, 0, 0, 0, 0}, y[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float* dst = mt.leaf_vals + size_t(leaf) * 512 + lane * 16;
    if (!is_new) { ld256(dst, x); ld256(dst + 8, y); }
    // voxels lane*16 .. lane*16+15 live in mask word lane>>2, bits (lane&3)*16 ..
    const uint32_t a16 = uint32_t(shfl64(a, lane >> 2) >> ((lane & 3) * 16)) & 0xFFFFu;
    const uint32_t v16 = uint32_t(shfl64(v, lane >> 2) >> ((lane & 3) * 16)) & 0xFFFFu;
#pragma unroll
    for (int q = 0; q < 8; ++q)
    {
      if ((a16 >> q) & 1u) x[q] = ((v16 >> q) & 1u) ? lo.max_lo : lo.min_lo;
      if ((a16 >> (q + 8)) & 1u) y[q] = ((v16 >> (q + 8)) & 1u) ? lo.max_lo : lo.min_lo;
    }
    st256(dst, x);
    st256(dst + 8, y);
    if (lane < 8)
    {
      const uint64_t old = is_new ? 0ull : mt.leaf_mask[size_t(leaf) * 8 + lane];
      mt.leaf_mask[size_t(leaf) * 8 + lane] = (old & ~a) | (a & v);
    }
    if (lane == 0) mt.leaf_dirty[leaf] = 1u;
  }
}

// ---- artificial areas (VDBMapping.hpp:1152-1236 and the tail of updateMap V:785-789) ----
// addArtificialWall: thread (wall, level) runs castRayIntoGrid(start + (0,0,i), end + (0,0,i)) with the same fp64 sequence
// as the scan DDA and marks straight into the artificial-area grid. Walls are short and rare: no tuning here.
__global__ void __launch_bounds__(128) wall_dda_kernel(const int32_t* walls /*[n][6] start xyz, end xyz*/, uint32_t n_walls, int32_t neg_index,
                                                      int32_t pos_index, UpdateGrid g, Counters* ctr)
{
  const int32_t levels = pos_index - neg_index;
  const uint64_t t     = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (levels <= 0 || t >= uint64_t(n_walls) * uint64_t(levels)) return;
  const uint32_t w = uint32_t(t / uint64_t(levels));
  const int32_t lv = neg_index + int32_t(t % uint64_t(levels));
  int32_t vox[3]   = {walls[6 * w], walls[6 * w + 1], walls[6 * w + 2] + lv};
  const int32_t en[3] = {walls[6 * w + 3], walls[6 * w + 4], walls[6 * w + 5] + lv};
  if (vox[0] == en[0] && vox[1] == en[1] && vox[2] == en[2]) return; // V:559
  double next[3], delta[3];
  int32_t step[3];
  for (int k = 0; k < 3; ++k)
  {
    const double dir = __dsub_rn(double(en[k]), double(vox[k]));
    if (dir == 0.0) { step[k] = 0; next[k] = DBL_MAX; delta[k] = DBL_MAX; }
    else
    {
      delta[k] = fabs(__ddiv_rn(1.0, dir));
      next[k]  = __dmul_rn(0.5, delta[k]);
      step[k]  = dir > 0.0 ? 1 : -1;
    }
  }
  uint64_t cur_bkey = kEmptyKey;
  uint32_t slot     = kInvalid;
  bool more;
  do
  {
    if (max(abs(vox[0]), max(abs(vox[1]), abs(vox[2]))) >= kVoxelLimit) { atomicOr(&ctr->flags, kFlagCoordRange); return; }
    const uint64_t bkey = packLeafKey(vox[0] >> 6, vox[1] >> 6, vox[2] >> 6);
    if (bkey != cur_bkey) { cur_bkey = bkey; slot = brickFindOrInsert(g, bkey, ctr); }
    if (slot != kInvalid)
      redOr64(g.act + size_t(slot) * (kBrickLeaves * 8) + brickWordOffset(vox[0], vox[1], vox[2]), uint64_t(1) << (((vox[1] & 7) << 3) | (vox[2] & 7)));
    // DDA::step with math::MinIndex: x iff n0 < n1 && n0 < n2; else y iff n1 < n2; else z
    const int axis = (next[0] < next[1] && next[0] < next[2]) ? 0 : ((next[1] < next[2]) ? 1 : 2);
    const double tt = next[axis];
    next[axis]      = __dadd_rn(next[axis], delta[axis]);
    vox[axis] += step[axis];
    more = (tt <= 1.0);
  } while (more);
}

// V:785-789: acc.setActiveState(coord, true) for every voxel of the artificial-area grid (a missing leaf is created with
// background values). Warp per artificial leaf.
__global__ void __launch_bounds__(256) grid_activate_kernel(UpdateGrid g, uint32_t n_entries, MapTable mt, Counters* ctr)
{
  const int lane         = threadIdx.x & 31;
  const uint32_t warp    = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t i = warp; i < n_entries; i += n_warps)
  {
    const uint32_t e = g.entries[i];
    const uint64_t a = (lane < 8) ? g.act[size_t(e) * 8 + lane] : 0;
    if (!(__ballot_sync(kFull, a != 0) & 0xFFu)) continue;
    int is_new;
    const uint32_t leaf = warpFindOrCreateLeaf(mt, leafKeyOfEntry(g.bkeys[e >> 9], e & 511u), true, lane, is_new, ctr);
    if (leaf == kInvalid) continue;
    if (is_new)
    {
      float z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      st256(mt.leaf_vals + size_t(leaf) * 512 + lane * 16, z);
      st256(mt.leaf_vals + size_t(leaf) * 512 + lane * 16 + 8, z);
    }
    if (lane < 8) mt.leaf_mask[size_t(leaf) * 8 + lane] = (is_new ? 0ull : mt.leaf_mask[size_t(leaf) * 8 + lane]) | a;
    if (lane == 0) mt.leaf_dirty[leaf] = 1u;
  }
}

// restoreMapIntegrity V:1152-1166: setNodeState (active = value > thres_max, O:130-135) through
// modifyValueAndActiveState for every artificial-area voxel. On a missing leaf OpenVDB probes the (0, inactive) tile:
// the leaf is created only if 0 > thres_max (then the voxel becomes active with value 0).
__global__ void __launch_bounds__(256) restore_state_kernel(UpdateGrid g, uint32_t n_entries, MapTable mt, LogOdds lo, Counters* ctr)
{
  const int lane         = threadIdx.x & 31;
  const uint32_t warp    = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
  const bool bg_active   = 0.0f > lo.thres_max;
  for (uint32_t i = warp; i < n_entries; i += n_warps)
  {
    const uint32_t e = g.entries[i];
    const uint64_t a = (lane < 8) ? g.act[size_t(e) * 8 + lane] : 0;
    if (!(__ballot_sync(kFull, a != 0) & 0xFFu)) continue;
    int is_new;
    const uint32_t leaf = warpFindOrCreateLeaf(mt, leafKeyOfEntry(g.bkeys[e >> 9], e & 511u), bg_active, lane, is_new, ctr);
    if (leaf == kInvalid) continue;
    float x[8] = {0, 0, 0, 0, 0, 0, 0, 0}, y[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float* src = mt.leaf_vals + size_t(leaf) * 512 + lane * 16;
    if (is_new) { st256(src, x); st256(src + 8, y); }
    else { ld256(src, x); ld256(src + 8, y); }
    // "value > thres_max" bits of this lane's 16 voxels, gathered into the 8 mask words
    uint32_t on16 = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q)
    {
      on16 |= (x[q] > lo.thres_max ? 1u : 0u) << q;
      on16 |= (y[q] > lo.thres_max ? 1u : 0u) << (q + 8);
    }
    uint64_t on = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) on |= uint64_t(__shfl_sync(kFull, on16, (lane & 7) * 4 + k)) << (16 * k);
    if (lane < 8)
    {
      const uint64_t old = is_new ? 0ull : mt.leaf_mask[size_t(leaf) * 8 + lane];
      mt.leaf_mask[size_t(leaf) * 8 + lane] = (old & ~a) | (a & on);
    }
    if (lane == 0) mt.leaf_dirty[leaf] = 1u;
  }
}


// ====================================================================================================
// fast_mode (castRayIntoGridFast, VDBMapping.hpp:577-602) and raytrace (VDBMapping.hpp:675-721): SURVEY.md 8f N3 / N4.
// Both are built on openvdb::tools::VolumeRayIntersector<FloatGrid>, restated here on the flat map: the node levels of the
// 5-4-3 tree are the 4096^3 / 128^3 / 8^3 blocks that hold at least one map leaf (nodes of this path are only ever created
// together with a leaf and never pruned; the device map holds no tiles), kept as two small hash sets of coarse block keys
// next to the leaf hash. Arithmetic follows math/Ray.h, math/DDA.h and tools/RayIntersector.h operation for operation
// (explicit _rn intrinsics, -fmad=false). Parity is unpinned at the
// OpenVDB boundary like everything else that has no vector in the reference.
// ====================================================================================================
namespace {

struct HRay
{
  double eye[3], dir[3], inv[3];
};
struct DdaState
{
  double t0, t1, next[3], delta[3];
  int32_t vox[3], step[3];
};

// DDA<Ray, LOG2>::init(ray, startTime, maxTime)
template <int LOG2>
__device__ __forceinline__ void ddaInit(DdaState& d, const HRay& r, double start, double maxt)
{
  constexpr int32_t DIM = int32_t(1) << LOG2;
  d.t0 = start;
  d.t1 = maxt;
#pragma unroll
  for (int k = 0; k < 3; ++k)
  {
    const double pos = __dadd_rn(r.eye[k], __dmul_rn(r.dir[k], start)); // ray(t) = eye + dir * t
    const int32_t v  = int32_t(floor(pos)) & ~(DIM - 1);
    d.vox[k]         = v;
    if (r.dir[k] == 0.0)
    {
      d.step[k]  = 0;
      d.next[k]  = DBL_MAX;
      d.delta[k] = DBL_MAX;
    }
    else if (r.inv[k] > 0.0)
    {
      d.step[k]  = DIM;
      d.next[k]  = __dadd_rn(start, __dmul_rn(__dsub_rn(double(v + DIM), pos), r.inv[k]));
      d.delta[k] = __dmul_rn(double(DIM), r.inv[k]);
    }
    else
    {
      d.step[k]  = -DIM;
      d.next[k]  = __dadd_rn(start, __dmul_rn(__dsub_rn(double(v), pos), r.inv[k]));
      d.delta[k] = __dmul_rn(double(-DIM), r.inv[k]);
    }
  }
}
// DDA::step(): axis = math::MinIndex(next) (x iff n0 < n1 && n0 < n2; else y iff n1 < n2; else z)
__device__ __forceinline__ bool ddaStep(DdaState& d)
{
  if (d.next[0] < d.next[1] && d.next[0] < d.next[2])
  {
    d.t0 = d.next[0];
    d.next[0] = __dadd_rn(d.next[0], d.delta[0]);
    d.vox[0] += d.step[0];
  }
  else if (d.next[1] < d.next[2])
  {
    d.t0 = d.next[1];
    d.next[1] = __dadd_rn(d.next[1], d.delta[1]);
    d.vox[1] += d.step[1];
  }
  else
  {
    d.t0 = d.next[2];
    d.next[2] = __dadd_rn(d.next[2], d.delta[2]);
    d.vox[2] += d.step[2];
  }
  return d.t0 <= d.t1;
}
// DDA::next() = math::Min(t1, next[0], next[1], next[2])
__device__ __forceinline__ double ddaNext(const DdaState& d)
{
  const double a = d.next[0] < d.t1 ? d.next[0] : d.t1;
  const double b = d.next[2] < d.next[1] ? d.next[2] : d.next[1];
  return b < a ? b : a;
}

__device__ __forceinline__ bool coarseContains(const uint64_t* keys, uint32_t mask, uint64_t key)
{
  uint32_t h = uint32_t(mix64(key)) & mask;
  for (uint32_t probe = 0; probe <= mask; ++probe)
  {
    const uint64_t k = keys[h];
    if (k == key) return true;
    if (k == kEmptyKey) return false;
    h = (h + 1) & mask;
  }
  return false;
}
__device__ __forceinline__ uint32_t mapFindLeaf(const MapTable& mt, uint64_t key)
{
  uint32_t h = uint32_t(mix64(key)) & mt.hcap_mask;
  for (uint32_t probe = 0; probe <= mt.hcap_mask; ++probe)
  {
    const uint64_t k = mt.hkeys[h];
    if (k == key) return mt.hvals[h];
    if (k == kEmptyKey) return kInvalid;
    h = (h + 1) & mt.hcap_mask;
  }
  return kInvalid;
}

// Ray::clip(bbox) (Ray::intersects): slab test; the span is only replaced on a hit
__device__ __forceinline__ bool rayClip(const HRay& r, const int32_t* bb /*min xyz, max xyz*/, double& t0, double& t1)
{
  double a0 = t0, a1 = t1;
#pragma unroll
  for (int k = 0; k < 3; ++k)
  {
    double a = __dmul_rn(__dsub_rn(double(bb[k]), r.eye[k]), r.inv[k]);
    double b = __dmul_rn(__dsub_rn(double(bb[3 + k]), r.eye[k]), r.inv[k]);
    if (a > b) { const double t = a; a = b; b = t; }
    if (a > a0) a0 = a;
    if (b < a1) a1 = b;
    if (a0 > a1) return false;
  }
  t0 = a0;
  t1 = a1;
  return true;
}

constexpr double kTimeDelta = 1e-9; // math::Delta<double>::value()

// VolumeHDDA<Tree, Ray, 2> over the ray span [t0, t1]: fn(span_t0, span_t1) is called for every valid span of consecutive
// leaf nodes; it returns true to terminate (march) or false to go on (hits). The last span (closed by the end of the ray)
// is delivered like the others.
template <typename Fn>
__device__ __forceinline__ void volumeHdda(const HRay& r, double t0, double t1, const MapTable& mt, const CoarseSets& cs, Fn&& fn)
{
  double s0 = -1.0, s1 = -1.0;
  DdaState d2;
  ddaInit<12>(d2, r, t0, t1);
  do
  {
    if (coarseContains(cs.k2, cs.mask2, packLeafKey(d2.vox[0] >> 12, d2.vox[1] >> 12, d2.vox[2] >> 12)))
    {
      DdaState d1;
      ddaInit<7>(d1, r, d2.t0, ddaNext(d2));
      do
      {
        if (coarseContains(cs.k1, cs.mask1, packLeafKey(d1.vox[0] >> 7, d1.vox[1] >> 7, d1.vox[2] >> 7)))
        {
          DdaState d0;
          ddaInit<3>(d0, r, d1.t0, ddaNext(d1));
          do
          {
            if (mapFindLeaf(mt, packLeafKey(d0.vox[0] >> 3, d0.vox[1] >> 3, d0.vox[2] >> 3)) != kInvalid)
            {
              if (s0 < 0.0) s0 = d0.t0;
            }
            else if (s0 >= 0.0)
            {
              s1 = d0.t0;
              if (__dsub_rn(s1, s0) > kTimeDelta && fn(s0, s1)) return;
              s0 = s1 = -1.0;
            }
          } while (ddaStep(d0));
          if (s0 >= 0.0) s1 = d0.t1;
        }
        else if (s0 >= 0.0)
        {
          s1 = d1.t0;
          if (__dsub_rn(s1, s0) > kTimeDelta && fn(s0, s1)) return;
          s0 = s1 = -1.0;
        }
      } while (ddaStep(d1));
      if (s0 >= 0.0) s1 = d1.t1;
    }
    else if (s0 >= 0.0)
    {
      s1 = d2.t0;
      if (__dsub_rn(s1, s0) > kTimeDelta && fn(s0, s1)) return;
      s0 = s1 = -1.0;
    }
  } while (ddaStep(d2));
  if (s0 >= 0.0) s1 = d2.t1;
  if (__dsub_rn(s1, s0) > kTimeDelta) fn(s0, s1);
}

// GridT::Accessor::isValueOn(voxel) on the flat map with a one-leaf cache
struct LeafCache
{
  uint64_t key  = kEmptyKey;
  uint32_t leaf = kInvalid;
};
__device__ __forceinline__ bool mapIsValueOn(const MapTable& mt, LeafCache& c, const int32_t v[3])
{
  const uint64_t key = packLeafKey(v[0] >> 3, v[1] >> 3, v[2] >> 3);
  if (key != c.key)
  {
    c.key  = key;
    c.leaf = mapFindLeaf(mt, key);
  }
  if (c.leaf == kInvalid) return false;
  return (mt.leaf_mask[size_t(c.leaf) * 8 + (v[0] & 7)] >> (((v[1] & 7) << 3) | (v[2] & 7))) & 1ull;
}

} // namespace

// the 128^3 and 4096^3 block keys of map leaves [from, to) -> the coarse sets (leaves are only ever appended to the pool)
__global__ void __launch_bounds__(256) coarse_insert_kernel(MapTable mt, uint32_t from, uint32_t to, CoarseSets cs, Counters* ctr)
{
  const uint32_t i = from + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= to) return;
  const uint64_t key = mt.leaf_keys[i];
  const int32_t lx = int32_t(uint32_t(key >> 42) & 0x1FFFFFu) - kLeafBias, ly = int32_t(uint32_t(key >> 21) & 0x1FFFFFu) - kLeafBias,
                lz = int32_t(uint32_t(key) & 0x1FFFFFu) - kLeafBias;
#pragma unroll
  for (int level = 0; level < 2; ++level)
  {
    const int sh         = level ? 9 : 4; // leaf coordinate -> 128^3 block (>> 4), 4096^3 block (>> 9)
    uint64_t* keys       = level ? cs.k2 : cs.k1;
    const uint32_t mask  = level ? cs.mask2 : cs.mask1;
    const uint64_t ckey  = packLeafKey(lx >> sh, ly >> sh, lz >> sh);
    uint32_t h           = uint32_t(mix64(ckey)) & mask;
    bool done            = false;
    for (uint32_t probe = 0; probe <= mask && !done; ++probe)
    {
      const uint64_t k = keys[h];
      if (k == ckey) done = true;
      else if (k == kEmptyKey)
      {
        const unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long*>(keys + h), kEmptyKey, ckey);
        if (old == kEmptyKey)
        {
          atomicAdd(level ? &cs.counts[1] : &cs.counts[0], 1u);
          done = true;
        }
        else if (old == ckey) done = true;
      }
      h = (h + 1) & mask;
    }
    if (!done) atomicOr(&ctr->flags, kFlagUpdateOverflow);
  }
}

// RootNode::evalActiveBoundingBox(bbox, visit_voxels = false): union of the node boxes of all leaves that hold an active voxel.
// out6 = min xyz, max xyz (voxel coordinates, max inclusive), initialised to INT_MAX / INT_MIN by the host.
__global__ void __launch_bounds__(256) active_bbox_kernel(MapTable mt, uint32_t n_leaves, int32_t* out6)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  int32_t mn[3] = {INT32_MAX, INT32_MAX, INT32_MAX}, mx[3] = {INT32_MIN, INT32_MIN, INT32_MIN};
  if (i < n_leaves)
  {
    const uint64_t* w = mt.leaf_mask + size_t(i) * 8;
    if (w[0] | w[1] | w[2] | w[3] | w[4] | w[5] | w[6] | w[7])
    {
      int32_t o[3];
      unpackLeafOrigin(mt.leaf_keys[i], o[0], o[1], o[2]);
      for (int k = 0; k < 3; ++k) { mn[k] = o[k]; mx[k] = o[k] + 7; }
    }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k)
  {
    mn[k] = __reduce_min_sync(kFull, mn[k]);
    mx[k] = __reduce_max_sync(kFull, mx[k]);
  }
  if ((threadIdx.x & 31) == 0 && mn[0] != INT32_MAX)
  {
    for (int k = 0; k < 3; ++k)
    {
      atomicMin(out6 + k, mn[k]);
      atomicMax(out6 + 3 + k, mx[k]);
    }
  }
}

// castRayIntoGridFast for every prepared ray (thread per ray; the rays of a fast-mode scan are cheap: ~len/8 node probes plus
// voxel probes inside the spans). bb = intersector bbox (max already offset by 1). Marks x-slice words like every
// non-DDA marker. Followed by the end point rule V:533-536.
__global__ void __launch_bounds__(128) raycast_fast_kernel(RaycastArgs a, UpdateGrid g, MapTable mt, CoarseSets cs, const int32_t* bbox6,
                                                          uint32_t map_empty, Counters* ctr)
{
  const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  unsigned marks = 0;
  if (i < a.n)
  {
    const int4 q = a.ends[i];
    if (q.w & 1)
    {
      uint64_t cur_bkey = kEmptyKey;
      uint32_t slot     = kInvalid;
      auto mark = [&](const int32_t v[3], bool value) {
        const uint64_t bkey = packLeafKey(v[0] >> 6, v[1] >> 6, v[2] >> 6);
        if (bkey != cur_bkey) { cur_bkey = bkey; slot = brickFindOrInsert(g, bkey, ctr); }
        if (slot == kInvalid) return;
        const size_t w     = size_t(slot) * (kBrickLeaves * 8) + brickWordOffset(v[0], v[1], v[2]);
        const uint64_t bit = uint64_t(1) << (((v[1] & 7) << 3) | (v[2] & 7));
        redOr64(g.act + w, bit);
        if (value) redOr64(g.val + w, bit);
      };
      const int32_t e[3] = {q.x, q.y, q.z};
      if (!map_empty) // V:522
      {
        HRay r;
#pragma unroll
        for (int k = 0; k < 3; ++k)
        {
          r.dir[k] = __dsub_rn(double(e[k]), double(a.origin_idx[k])); // V:586
          r.eye[k] = __dadd_rn(double(a.origin_idx[k]), 0.5);          // V:587
          r.inv[k] = __ddiv_rn(1.0, r.dir[k]);
        }
        int32_t bb[6];
        for (int k = 0; k < 3; ++k) { bb[k] = bbox6[k]; bb[3 + k] = int32_t(uint32_t(bbox6[3 + k]) + 1u); }
        double t0 = 0.0, t1 = 1.0;
        rayClip(r, bb, t0, t1); // setIndexRay V:588; a miss leaves [0, 1]
        LeafCache cache;
        volumeHdda(r, t0, t1, mt, cs, [&](double h0, double h1) {
          DdaState d;
          ddaInit<0>(d, r, h0, h1); // fine_ray V:593-594
          do
          {
            if (mapIsValueOn(mt, cache, d.vox)) // V:597
            {
              mark(d.vox, false);
              ++marks;
            }
          } while (ddaStep(d));
          return false;
        });
      }
      if (q.w & 2) mark(e, true); // V:533-536
    }
  }
  marks = __reduce_add_sync(kFull, marks);
  if ((threadIdx.x & 31) == 0 && marks) atomicAdd(&ctr->visits, (unsigned long long)marks);
}

// raytrace V:675-721, thread per ray. success: 1 hit, 0 miss; a ray whose index-space end points leave the +-2^23 voxel range
// is reported as a miss and raises kFlagCoordRange.
__global__ void __launch_bounds__(128) raytrace_kernel(uint64_t n, const double* origins, const double* directions, const double* max_lengths,
                                                      double res, double inv_res, MapTable mt, CoarseSets cs, const int32_t* bbox6,
                                                      uint32_t map_empty, int32_t* success, double* end_points, Counters* ctr)
{
  const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  HRay r;
  const double dx = directions[3 * i], dy = directions[3 * i + 1], dz = directions[3 * i + 2];
  const double len = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
  const double il  = __ddiv_rn(1.0, len); // Vec3::normalize: *= 1 / length
  const double d3[3] = {dx, dy, dz};
  bool in_range = true;
#pragma unroll
  for (int k = 0; k < 3; ++k)
  {
    const double dn = __dmul_rn(__dmul_rn(d3[k], il), max_lengths[i]);  // V:691-692
    r.eye[k]        = __dmul_rn(origins[3 * i + k], inv_res);           // V:694
    r.dir[k]        = __dmul_rn(dn, inv_res);                           // V:695
    r.inv[k]        = __ddiv_rn(1.0, r.dir[k]);
    if (!(fabs(r.eye[k]) < double(kVoxelLimit)) || !(fabs(__dadd_rn(r.eye[k], r.dir[k])) < double(kVoxelLimit))) in_range = false;
  }
  bool hit = false;
  double h0 = -1.0, h1 = -1.0;
  if (!in_range) atomicOr(&ctr->flags, kFlagCoordRange);
  else if (!map_empty)
  {
    int32_t bb[6];
    for (int k = 0; k < 3; ++k) { bb[k] = bbox6[k]; bb[3 + k] = int32_t(uint32_t(bbox6[3 + k]) + 1u); }
    double t0 = 0.0, t1 = 1.0;
    rayClip(r, bb, t0, t1);
    if (__dsub_rn(t1, t0) > kTimeDelta) // VolumeHDDA::march: only a valid ray is marched
      volumeHdda(r, t0, t1, mt, cs, [&](double a0, double a1) {
        if (hit) return true;
        h0  = a0;
        h1  = a1;
        hit = true;
        return true;
      });
  }
  if (hit)
  {
    DdaState d;
    ddaInit<0>(d, r, h0, h1);
    LeafCache cache;
    while (ddaStep(d) && !mapIsValueOn(mt, cache, d.vox)) {} // V:709-713
    for (int k = 0; k < 3; ++k) end_points[3 * i + k] = __dmul_rn(double(d.vox[k]), res); // V:714
    success[i] = 1;
  }
  else
  {
    for (int k = 0; k < 3; ++k) end_points[3 * i + k] = __dmul_rn(__dadd_rn(r.eye[k], r.dir[k]), res); // V:719
    success[i] = 0;
  }
}


// Order-independent checksum of the map (multi-GPU parity witness: the sum over all shards must equal the single-GPU
// map's). Warp per leaf; a leaf's hash mixes its key, its 8 mask words and its 512 value bit patterns position by position.
__global__ void __launch_bounds__(256) map_checksum_kernel(MapTable mt, uint32_t n_leaves, unsigned long long* out2)
{
  const int lane         = threadIdx.x & 31;
  const uint32_t warp    = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
  unsigned long long sum = 0, cnt = 0;
  for (uint32_t l = warp; l < n_leaves; l += n_warps)
  {
    float a[8], b[8];
    ld256(mt.leaf_vals + size_t(l) * 512 + lane * 16, a);
    ld256(mt.leaf_vals + size_t(l) * 512 + lane * 16 + 8, b);
    uint64_t h = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q)
    {
      h = mix64(h ^ (uint64_t(__float_as_uint(a[q])) | (uint64_t(lane * 16 + q + 1) << 32)));
      h = mix64(h ^ (uint64_t(__float_as_uint(b[q])) | (uint64_t(lane * 16 + q + 9) << 32)));
    }
    if (lane < 8) h ^= mix64(mt.leaf_mask[size_t(l) * 8 + lane] + 0x9E3779B97F4A7C15ULL * uint64_t(lane + 1));
    // combine the lanes position-dependently (lane index is already mixed into every term), then bind to the key
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) h += shflXor64(h, d);
    if (lane == 0) { sum += mix64(h ^ mix64(mt.leaf_keys[l])); ++cnt; }
  }
  if (lane == 0 && cnt)
  {
    atomicAdd(out2, sum);
    atomicAdd(out2 + 1, cnt);
  }
}

// ====================================================================================================
// launch wrappers
// ====================================================================================================
static inline unsigned blocksFor(uint64_t n, unsigned per_block) { return unsigned((n + per_block - 1) / per_block); }

static std::atomic<uint32_t> g_launches{0};
uint32_t launchCount() { return g_launches.load(); }
#define VDBM_LAUNCH(kernel, grid, block, stream, ...)                 \
  do                                                                  \
  {                                                                   \
    kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__);            \
    g_launches.fetch_add(1, std::memory_order_relaxed);               \
  } while (0)

static int smCount()
{
  static int sms = 0;
  if (!sms)
  {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

void launchPrepRays(const RaycastArgs& a, Counters* ctr, cudaStream_t s)
{
  if (a.n == 0) return;
  VDBM_LAUNCH(prep_rays_kernel, blocksFor(a.n, 256), 256, s, a, ctr);
}

void launchLongRaySegments(const RaycastArgs& a, uint32_t n_long, cudaStream_t s)
{
  if (n_long) VDBM_LAUNCH(long_ray_segments_kernel, blocksFor(blocksFor(n_long, 10), 4), 128, s, a, n_long);
}

int raycastDDAGrid(int device)
{
  int sms = 148, per_sm = 4;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, raycast_dda_kernel<0>, VDBM_DDA_BLOCK, 0);
  if (per_sm < 1) per_sm = 1;
  if (const char* e = getenv("VDBM_DDA_CTAS_PER_SM")) per_sm = std::max(1, std::min(per_sm, atoi(e))); // experiments
  return sms * per_sm; // persistent: exactly one resident wave
}
int raycastDDABlock() { return VDBM_DDA_BLOCK; }

size_t nearCopiesBytes() { return size_t(kNearCopies) * kNearBricks * kBrickLeaves * 8 * sizeof(uint64_t); }

void launchRaycastDDA(const RaycastArgs& a, UpdateGrid g, uint64_t* near_act, Counters* ctr, int grid, cudaStream_t s, bool test_before_set)
{
  if (a.n == 0) return;
  constexpr uint64_t kWarps = VDBM_DDA_BLOCK / 32;
  const uint64_t blocks     = (((uint64_t(a.n_segs) + 31) / 32) + kWarps - 1) / kWarps;
  if (uint64_t(grid) > blocks) grid = int(blocks);
#ifdef VDBM_EXPERIMENTS
  static const int mode = [] { const char* e = getenv("VDBM_DDA_MODE"); return e ? atoi(e) : 0; }();
  if (mode == 1) VDBM_LAUNCH(raycast_dda_kernel<1>, grid, VDBM_DDA_BLOCK, s, a, g, near_act, ctr);
  else if (mode == 2) VDBM_LAUNCH(raycast_dda_kernel<2>, grid, VDBM_DDA_BLOCK, s, a, g, near_act, ctr);
  else
#endif
  if (test_before_set) VDBM_LAUNCH(raycast_dda_kernel<4>, grid, VDBM_DDA_BLOCK, s, a, g, near_act, ctr);
  else VDBM_LAUNCH(raycast_dda_kernel<0>, grid, VDBM_DDA_BLOCK, s, a, g, near_act, ctr);
  VDBM_LAUNCH(merge_near_kernel, (kNearBricks * kBrickLeaves * 8 * kMergeSplit) / 256, 256, s, g, near_act, nearBrick0(a.origin_idx[0]),
              nearBrick0(a.origin_idx[1]), nearBrick0(a.origin_idx[2]), ctr);
}

void launchCompactLeaves(UpdateGrid g, cudaStream_t s, bool cook)
{
  cudaMemsetAsync(g.counters + 1, 0, 4, s);
  const unsigned grid = std::min<unsigned>(g.cap_mask + 1u, unsigned(smCount()) * 4u);
  if (cook) VDBM_LAUNCH(compact_leaves_kernel<true>, grid, 512, s, g);
  else VDBM_LAUNCH(compact_leaves_kernel<false>, grid, 512, s, g);
}

void launchUncookLeaves(UpdateGrid g, uint32_t n_entries, cudaStream_t s)
{
  if (n_entries) VDBM_LAUNCH(uncook_leaves_kernel, blocksFor(n_entries, 256), 256, s, g, n_entries);
}

static int applyUpdateBlocksPerSM()
{
  static int per_sm = 0;
  if (per_sm == 0)
  {
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, apply_update_kernel, 256, 0);
    if (per_sm < 1) per_sm = 1;
  }
  return per_sm;
}

void launchApplyUpdate(UpdateGrid g, MapTable mt, LogOdds lo, uint32_t* resolved, LeafRecord* change_out, uint32_t change_cap,
                       Counters* ctr, uint32_t n_entries, cudaStream_t s)
{
  if (n_entries == 0) return;
  VDBM_LAUNCH(resolve_leaves_kernel, blocksFor(n_entries, 256), 256, s, g, mt, lo, resolved, ctr, n_entries, (const uint32_t*)nullptr);
  // one warp per leaf, grid-stride; exactly one resident wave (multiple of the SM count)
  unsigned grid = std::min<unsigned>(blocksFor(n_entries, 8), unsigned(smCount() * applyUpdateBlocksPerSM()));
  VDBM_LAUNCH(apply_update_kernel, grid, 256, s, g, mt, lo, resolved, change_out, change_cap, ctr, n_entries, (const uint32_t*)nullptr);
}

void launchApplyUpdateDeferred(UpdateGrid g, MapTable mt, LogOdds lo, uint32_t* resolved, uint32_t resolved_cap, Counters* ctr,
                               uint32_t expected_entries, cudaStream_t s)
{
  VDBM_LAUNCH(update_guard_kernel, 1, 1, s, g, mt, ctr, resolved_cap);
  const unsigned rgrid = std::max(1u, std::min<unsigned>(blocksFor(std::max(expected_entries, 1u), 256), unsigned(smCount()) * 8u));
  VDBM_LAUNCH(resolve_leaves_kernel, rgrid, 256, s, g, mt, lo, resolved, ctr, 0u, (const uint32_t*)&ctr->deferred_entries);
  int per_sm = applyUpdateBlocksPerSM();
  if (const char* e = getenv("VDBM_APPLY_CTAS_PER_SM")) per_sm = std::max(1, std::min(per_sm, atoi(e))); // experiments (co-residency with the DDA)
  VDBM_LAUNCH(apply_update_kernel, unsigned(smCount() * per_sm), 256, s, g, mt, lo, resolved, (LeafRecord*)nullptr, 0u, ctr, 0u,
              (const uint32_t*)&ctr->deferred_entries);
  VDBM_LAUNCH(reset_bricks_kernel, blocksFor(uint64_t(g.cap_mask) + 1, 256), 256, s, g, 0u, (const uint32_t*)&ctr->deferred_bricks);
}

void launchResetBricks(UpdateGrid g, uint32_t n_bricks, cudaStream_t s)
{
  VDBM_LAUNCH(reset_bricks_kernel, std::max(1u, blocksFor(n_bricks, 256)), 256, s, g, n_bricks, (const uint32_t*)nullptr);
}
void launchClearEntries(UpdateGrid g, uint32_t n_entries, cudaStream_t s)
{
  if (n_entries) VDBM_LAUNCH(clear_entries_kernel, blocksFor(uint64_t(n_entries) * 16, 256), 256, s, g, n_entries);
}
void launchRehashUpdate(UpdateGrid old_g, uint32_t old_bricks, UpdateGrid new_g, Counters* ctr, cudaStream_t s)
{
  if (old_bricks) VDBM_LAUNCH(rehash_update_kernel, std::min<unsigned>(old_bricks, unsigned(smCount()) * 8u), 256, s, old_g, old_bricks, new_g, ctr);
}
void launchRehashMap(MapTable mt, uint32_t n_leaves, Counters* ctr, cudaStream_t s)
{
  if (n_leaves) VDBM_LAUNCH(rehash_map_kernel, blocksFor(n_leaves, 256), 256, s, mt, n_leaves, ctr);
}
void launchEntryKeys(UpdateGrid g, uint32_t n_entries, uint64_t* out_keys, uint32_t* out_entries, cudaStream_t s)
{
  if (n_entries) VDBM_LAUNCH(entry_keys_kernel, blocksFor(n_entries, 256), 256, s, g, n_entries, out_keys, out_entries);
}
void launchGatherUpdate(UpdateGrid g, uint32_t n, const uint64_t* keys, const uint32_t* entries, LeafRecord* out, cudaStream_t s)
{
  if (n) VDBM_LAUNCH(gather_update_kernel, blocksFor(uint64_t(n) * 16, 256), 256, s, g, n, keys, entries, out);
}
void launchImportUpdateSoA(UpdateGrid g, const int32_t* origins, const uint64_t* active, const uint64_t* value, uint64_t n, Counters* ctr,
                           cudaStream_t s)
{
  if (n) VDBM_LAUNCH(import_update_soa_kernel, blocksFor(n * 16, 256), 256, s, g, origins, active, value, n, ctr);
}
void launchImportUpdate(UpdateGrid g, const LeafRecord* recs, uint64_t n, Counters* ctr, cudaStream_t s)
{
  if (n) VDBM_LAUNCH(import_update_kernel, blocksFor(n * 16, 256), 256, s, g, recs, n, ctr);
}
void launchGatherMap(MapTable mt, uint32_t n, const uint32_t* leaf_idx, int32_t* origins, uint64_t* mask, float* vals, cudaStream_t s)
{
  if (!n) return;
  const unsigned grid = std::min<unsigned>(blocksFor(uint64_t(n) * 32, 256), unsigned(smCount()) * 8u);
  VDBM_LAUNCH(gather_map_kernel, grid, 256, s, mt, n, leaf_idx, origins, mask, vals);
}
void launchCollectDirty(MapTable mt, uint32_t n_leaves, uint32_t* out_idx, Counters* ctr, cudaStream_t s)
{
  if (n_leaves) VDBM_LAUNCH(collect_dirty_kernel, blocksFor(n_leaves, 256), 256, s, mt, n_leaves, out_idx, ctr);
}
void launchMarkDirty(MapTable mt, const uint32_t* idx, uint32_t n, cudaStream_t s)
{
  if (n) VDBM_LAUNCH(mark_dirty_kernel, blocksFor(n, 256), 256, s, mt, idx, n);
}
void launchSection(MapTable mt, uint32_t n_leaves, const int32_t bbmin[3], const int32_t bbmax[3], int full, int result_float,
                   uint64_t* out_keys, uint64_t* out_active, uint64_t* out_valmask, float* out_vals, uint32_t out_cap,
                   Counters* ctr, cudaStream_t s)
{
  if (!n_leaves) return;
  const unsigned grid = std::min<unsigned>(blocksFor(uint64_t(n_leaves) * 32, 256), unsigned(smCount()) * 8u);
  VDBM_LAUNCH(section_kernel, grid, 256, s, mt, n_leaves, bbmin[0], bbmin[1], bbmin[2], bbmax[0], bbmax[1], bbmax[2], full, result_float,
              out_keys, out_active, out_valmask, out_vals, out_cap, ctr);
}
void launchProbe(MapTable mt, int32_t x, int32_t y, int32_t z, float* out_val, int32_t* out_active, cudaStream_t s)
{
  VDBM_LAUNCH(probe_kernel, 1, 1, s, mt, x, y, z, out_val, out_active);
}
void launchPartition(UpdateGrid g, uint32_t n, ShardPlan plan, uint32_t* rank_counts, uint32_t* rank_cursor, LeafRecord* out, int pass,
                     cudaStream_t s)
{
  if (n) VDBM_LAUNCH(partition_kernel, blocksFor(n, 256), 256, s, g, n, plan, rank_counts, rank_cursor, out, pass);
}
void launchKeysFromIdx(const uint64_t* keys, const uint32_t* idx, uint32_t n, uint64_t* out_keys, uint32_t* out_idx, cudaStream_t s)
{
  if (n) VDBM_LAUNCH(keys_from_idx_kernel, blocksFor(n, 256), 256, s, keys, idx, n, out_keys, out_idx);
}
void launchSplitRecords(const LeafRecord* recs, const uint32_t* perm, uint32_t n, int32_t* origins, uint64_t* active, uint64_t* value, cudaStream_t s)
{
  if (n == 0) return;
  VDBM_LAUNCH(split_records_kernel, blocksFor(uint64_t(n) * 16, 256), 256, s, recs, perm, n, origins, active, value);
}
void launchRecordKeys(const LeafRecord* recs, uint32_t n, uint64_t* keys, uint32_t* idx, cudaStream_t s)
{
  if (n == 0) return;
  VDBM_LAUNCH(record_keys_kernel, blocksFor(n, 256), 256, s, recs, n, keys, idx);
}
void launchPermuteSection(uint32_t n, const uint64_t* sorted_keys, const uint32_t* perm, const uint64_t* in_active, const uint64_t* in_valmask,
                          const float* in_vals, int32_t* origins, uint64_t* out_active, uint64_t* out_valmask, float* out_vals, cudaStream_t s)
{
  if (n == 0) return;
  const unsigned blocks = unsigned(std::min<uint64_t>((uint64_t(n) * 32 + 255) / 256, 148ull * 16));
  VDBM_LAUNCH(permute_section_kernel, dim3(blocks), dim3(256), s, n, sorted_keys, perm, in_active, in_valmask, in_vals, origins, out_active,
              out_valmask, out_vals);
}

void launchSectionDeactivate(MapTable mt, uint32_t n_leaves, const int32_t bbmin[3], const int32_t bbmax[3], cudaStream_t s)
{
  if (n_leaves) VDBM_LAUNCH(section_deactivate_kernel, blocksFor(uint64_t(n_leaves) * 8, 256), 256, s, mt, n_leaves, bbmin[0], bbmin[1], bbmin[2], bbmax[0], bbmax[1], bbmax[2]);
}
void launchSectionActivate(MapTable mt, const uint64_t* keys, const uint64_t* active, uint32_t n, Counters* ctr, cudaStream_t s)
{
  if (n) VDBM_LAUNCH(section_activate_kernel, std::min<unsigned>(blocksFor(uint64_t(n) * 32, 256), unsigned(smCount()) * 8u), 256, s, mt, keys, active, n, ctr);
}
void launchSectionApplyGrid(MapTable mt, const uint64_t* keys, const uint64_t* active, const float* values, uint32_t n, Counters* ctr, cudaStream_t s)
{
  if (n) VDBM_LAUNCH(section_apply_grid_kernel, std::min<unsigned>(blocksFor(uint64_t(n) * 32, 256), unsigned(smCount()) * 8u), 256, s, mt, keys, active, values, n, ctr);
}
void launchSectionTileQuirk(MapTable mt, const uint64_t* blocks, uint32_t n_blocks, int level, const uint64_t* present, uint32_t n_present, cudaStream_t s)
{
  if (!n_blocks) return;
  const uint64_t threads = uint64_t(n_blocks) * (level == 0 ? 4096u : 32768u);
  VDBM_LAUNCH(section_tile_quirk_kernel, blocksFor(threads, 256), 256, s, mt, blocks, n_blocks, level, present, n_present);
}

void launchPushUpdate(UpdateGrid g, uint32_t n_entries, ExchangePeers px, ShardPlan plan, uint32_t parity, uint32_t epoch, uint32_t* cursors,
                      Counters* ctr, cudaStream_t s)
{
  cudaMemsetAsync(cursors, 0, kMaxRanks * sizeof(uint32_t), s);
  if (n_entries) VDBM_LAUNCH(push_update_kernel, blocksFor(n_entries, 256), 256, s, g, n_entries, px, plan, parity, cursors, ctr);
  VDBM_LAUNCH(publish_counts_kernel, 1, 32, s, px, parity, epoch, cursors);
}

void launchWaitPeers(const unsigned long long* ctrl, int32_t n_ranks, uint32_t parity, uint32_t epoch, uint32_t* counts_out, Counters* ctr,
                     cudaStream_t s)
{
  // a peer may legitimately be late by a reallocation or a table growth (hundreds of ms): default 20 s, VDBM_EXCHANGE_TIMEOUT_MS overrides
  static const unsigned long long timeout_ns = [] {
    const char* e = getenv("VDBM_EXCHANGE_TIMEOUT_MS");
    const double ms = e ? atof(e) : 20000.0;
    return (unsigned long long)((ms > 0 ? ms : 20000.0) * 1e6);
  }();
  VDBM_LAUNCH(wait_peers_kernel, 1, 32, s, ctrl, n_ranks, parity, epoch, counts_out, ctr, timeout_ns);
}
void launchPullUpdate(UpdateGrid g, const uint64_t* inbox, const unsigned long long* ctrl, uint32_t cap, int32_t n_ranks, uint32_t parity,
                      uint32_t epoch, uint32_t* counts_out, Counters* ctr, cudaStream_t s)
{
  (void)ctrl; (void)epoch;
  VDBM_LAUNCH(pull_update_kernel, unsigned(smCount()) * 4u, 256, s, g, inbox, cap, n_ranks, parity, counts_out, ctr);
}

void launchCoarseInsert(MapTable mt, uint32_t from, uint32_t to, CoarseSets cs, Counters* ctr, cudaStream_t s)
{
  if (to <= from) return;
  VDBM_LAUNCH(coarse_insert_kernel, blocksFor(to - from, 256), 256, s, mt, from, to, cs, ctr);
}
void launchActiveBBox(MapTable mt, uint32_t n_leaves, int32_t* out6, cudaStream_t s)
{
  if (!n_leaves) return;
  VDBM_LAUNCH(active_bbox_kernel, blocksFor(n_leaves, 256), 256, s, mt, n_leaves, out6);
}
void launchRaycastFast(const RaycastArgs& a, UpdateGrid g, MapTable mt, CoarseSets cs, const int32_t* bbox6, uint32_t map_empty, Counters* ctr,
                       cudaStream_t s)
{
  if (!a.n) return;
  VDBM_LAUNCH(raycast_fast_kernel, blocksFor(a.n, 128), 128, s, a, g, mt, cs, bbox6, map_empty, ctr);
}
void launchRaytrace(uint64_t n, const double* origins, const double* directions, const double* max_lengths, double res, double inv_res, MapTable mt,
                    CoarseSets cs, const int32_t* bbox6, uint32_t map_empty, int32_t* success, double* end_points, Counters* ctr, cudaStream_t s)
{
  if (!n) return;
  VDBM_LAUNCH(raytrace_kernel, blocksFor(n, 128), 128, s, n, origins, directions, max_lengths, res, inv_res, mt, cs, bbox6, map_empty, success,
         end_points, ctr);
}
void launchMapChecksum(MapTable mt, uint32_t n_leaves, unsigned long long* out2, cudaStream_t s)
{
  if (!n_leaves) return;
  const unsigned grid = std::min<unsigned>(blocksFor(uint64_t(n_leaves) * 32, 256), unsigned(smCount()) * 8u);
  VDBM_LAUNCH(map_checksum_kernel, grid, 256, s, mt, n_leaves, out2);
}

size_t sortRaysByLength(void* d_temp, size_t temp_bytes, const uint32_t* keys_in, uint32_t* keys_out, const uint32_t* idx_in,
                        uint32_t* idx_out, uint32_t n, cudaStream_t s)
{
  size_t bytes = temp_bytes;
  if (d_temp == nullptr) bytes = 0;
  cub::DeviceRadixSort::SortPairsDescending(d_temp, bytes, keys_in, keys_out, idx_in, idx_out, int(n), kSortLoBit, kSortHiBit, s);
  return bytes;
}

size_t sortKeys32(void* d_temp, size_t temp_bytes, const uint32_t* keys_in, uint32_t* keys_out, uint32_t n, cudaStream_t s)
{
  size_t bytes = temp_bytes;
  if (d_temp == nullptr) bytes = 0;
  cub::DeviceRadixSort::SortKeys(d_temp, bytes, keys_in, keys_out, int(n), 0, 32, s);
  return bytes;
}

size_t sortPairs(void* d_temp, size_t temp_bytes, const uint64_t* keys_in, uint64_t* keys_out, const uint32_t* idx_in,
                 uint32_t* idx_out, uint32_t n, cudaStream_t s)
{
  size_t bytes = temp_bytes;
  if (d_temp == nullptr) bytes = 0;
  cub::DeviceRadixSort::SortPairs(d_temp, bytes, keys_in, keys_out, idx_in, idx_out, int(n), 0, 63, s);
  return bytes;
}

void launchMarkEnds(const int4* ends, uint64_t n, UpdateGrid g, Counters* ctr, cudaStream_t s)
{
  if (n == 0) return;
  VDBM_LAUNCH(mark_ends_kernel, dim3(unsigned((n + 255) / 256)), dim3(256), s, ends, n, g, ctr);
}
void launchExpandEnds(UpdateGrid g, uint32_t n_entries, int4* out, uint32_t out_cap, Counters* ctr, cudaStream_t s)
{
  if (n_entries == 0) return;
  const unsigned blocks = unsigned(std::min<uint64_t>((uint64_t(n_entries) * 32 + 255) / 256, 148ull * 16));
  VDBM_LAUNCH(expand_ends_kernel, dim3(blocks), dim3(256), s, g, n_entries, out, out_cap, ctr);
}
void launchPointsToEnds(const uint8_t* points, uint64_t n, uint32_t stride, double inv_res, int occupied, int4* out, Counters* ctr, cudaStream_t s)
{
  if (n == 0) return;
  VDBM_LAUNCH(points_to_ends_kernel, dim3(unsigned((n + 255) / 256)), dim3(256), s, points, n, stride, inv_res, occupied, out, ctr);
}
void launchOverwrite(UpdateGrid g, uint32_t n_entries, MapTable mt, LogOdds lo, Counters* ctr, cudaStream_t s)
{
  if (n_entries == 0) return;
  const unsigned blocks = unsigned(std::min<uint64_t>((uint64_t(n_entries) * 32 + 255) / 256, 148ull * 16));
  VDBM_LAUNCH(overwrite_kernel, dim3(blocks), dim3(256), s, g, n_entries, mt, lo, ctr);
}
void launchWallDDA(const int32_t* d_walls, uint32_t n_walls, int32_t neg_index, int32_t pos_index, UpdateGrid g, Counters* ctr, cudaStream_t s)
{
  const int64_t levels = int64_t(pos_index) - int64_t(neg_index);
  if (n_walls == 0 || levels <= 0) return;
  const uint64_t threads = uint64_t(n_walls) * uint64_t(levels);
  VDBM_LAUNCH(wall_dda_kernel, dim3(unsigned((threads + 127) / 128)), dim3(128), s, d_walls, n_walls, neg_index, pos_index, g, ctr);
}
void launchGridActivate(UpdateGrid g, uint32_t n_entries, MapTable mt, Counters* ctr, cudaStream_t s)
{
  if (n_entries == 0) return;
  const unsigned blocks = unsigned(std::min<uint64_t>((uint64_t(n_entries) * 32 + 255) / 256, 148ull * 16));
  VDBM_LAUNCH(grid_activate_kernel, dim3(blocks), dim3(256), s, g, n_entries, mt, ctr);
}
void launchRestoreState(UpdateGrid g, uint32_t n_entries, MapTable mt, LogOdds lo, Counters* ctr, cudaStream_t s)
{
  if (n_entries == 0) return;
  const unsigned blocks = unsigned(std::min<uint64_t>((uint64_t(n_entries) * 32 + 255) / 256, 148ull * 16));
  VDBM_LAUNCH(restore_state_kernel, dim3(blocks), dim3(256), s, g, n_entries, mt, lo, ctr);
}

} // namespace vdbm
