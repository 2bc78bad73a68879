// vdbm_device.cuh — shared device/host types of libvdbm_b200 (sm_100a only).
//
// Data layout in HBM:
//   update grid (one per input source; replaces openvdb Tree4<bool,1,4,3>, VDBMapping.hpp:89,97)
//     open-addressing hash of 64^3-voxel BRICKS, slot == storage:
//       bkeys[C] u64 | act[C][512 leaves][8] u64 | val[C][512][8] u64   (64 KB per brick slot)
//     + btouched[C] (occupied slots) + entries[] (compact list of touched leaves, rebuilt after each raycast)
//   map (replaces openvdb Tree4<float,5,4,3> = FloatTree, VDBMapping.hpp:88)
//     open-addressing hash  hkeys[H] u64 -> hvals[H] u32 (leaf index)
//     leaf pool (SoA):  leaf_keys[P] u64 | leaf_mask[P][8] u64 | leaf_vals[P][512] f32 | leaf_dirty[P] u32 (flag)
//   in-leaf layout mirrors openvdb::tree::LeafNode<T,3>: offset n = (x&7)<<6 | (y&7)<<3 | (z&7),
//   mask word n>>6, bit n&63.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace vdbm {

constexpr uint64_t kEmptyKey   = ~uint64_t(0);
constexpr uint32_t kInvalid    = 0xFFFFFFFFu;
constexpr int32_t kLeafBias    = 1 << 20; // leaf coordinate bias: 21 bits per axis
constexpr int32_t kVoxelLimit  = 1 << 23; // |voxel coordinate| must stay below this

// status flag bits written by kernels
constexpr uint32_t kFlagUpdateOverflow = 1u; // update hash full / probe limit hit
constexpr uint32_t kFlagMapOverflow    = 2u; // map hash or leaf pool full
constexpr uint32_t kFlagCoordRange     = 4u; // voxel coordinate outside +-2^23
constexpr uint32_t kFlagExchangeOverflow = 8u; // a sender region of a peer inbox was too small
constexpr uint32_t kFlagExchangeTimeout  = 16u; // a peer did not publish its epoch in time

// 63-bit leaf key from LEAF coordinates (voxel >> 3). Sorting keys ascending == sorting leaf origins
// lexicographically by (x, y, z) (the canonical export order).
__host__ __device__ __forceinline__ uint64_t packLeafKey(int32_t lx, int32_t ly, int32_t lz)
{
  return (uint64_t(uint32_t(lx + kLeafBias) & 0x1FFFFFu) << 42) | (uint64_t(uint32_t(ly + kLeafBias) & 0x1FFFFFu) << 21) |
         uint64_t(uint32_t(lz + kLeafBias) & 0x1FFFFFu);
}
__host__ __device__ __forceinline__ void unpackLeafOrigin(uint64_t key, int32_t& x, int32_t& y, int32_t& z)
{
  x = (int32_t(uint32_t(key >> 42) & 0x1FFFFFu) - kLeafBias) * 8;
  y = (int32_t(uint32_t(key >> 21) & 0x1FFFFFu) - kLeafBias) * 8;
  z = (int32_t(uint32_t(key) & 0x1FFFFFu) - kLeafBias) * 8;
}

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x)
{
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return x;
}

// Owner rank of a leaf for the multi-GPU sharding: leaves are grouped in 2x2x2 bricks so
// that consecutive DDA leaves of a ray mostly share an owner, then hashed.
__host__ __device__ __forceinline__ int32_t leafOwner(uint64_t key, int32_t n_ranks)
{
  const uint64_t brick = key & ~((uint64_t(1) << 42) | (uint64_t(1) << 21) | uint64_t(1));
  return int32_t(mix64(brick ^ 0x9E3779B97F4A7C15ULL) % uint64_t(n_ranks));
}

// Sharding plan of the map (multi-GPU). mode 0: the hash above (nothing is local on purpose: any sensor layout balances).
// mode 1: AZIMUTH SECTORS around a fixed leaf column (cx, cy): rank r owns the leaves whose centre direction falls into
// [bounds[r], bounds[r+1]) (the last sector wraps), measured as a "diamond angle" in [0, 4) - a monotone function of the
// true angle that needs one IEEE division of small integers, so host, device and the Python twin agree bit for bit.
// A ray cast from a sensor near the centre stays inside its azimuth sector: when the rays are split by the same bounds
// nearly every touched leaf is owned by the rank that touched it, and only the leaves around the sensor and along the
// sector borders cross NVLink.
constexpr int kMaxRanks = 16;
struct ShardPlan
{
  int32_t mode, n_ranks;
  int32_t cx, cy;           // centre, in LEAF coordinates (voxel >> 3)
  double bounds[kMaxRanks]; // ascending diamond angles, one per rank
};
__host__ __device__ __forceinline__ double diamondAngle(double dx, double dy)
{
  if (dx == 0.0 && dy == 0.0) return 0.0;
  if (dy >= 0.0) return dx >= 0.0 ? dy / (dx + dy) : 1.0 - dx / (dy - dx);
  return dx < 0.0 ? 2.0 - dy / (-dx - dy) : 3.0 + dx / (dx - dy);
}
__host__ __device__ __forceinline__ int32_t leafOwnerPlanned(uint64_t key, const ShardPlan& sp)
{
  if (sp.mode == 0) return leafOwner(key, sp.n_ranks);
  const int32_t lx = int32_t(uint32_t(key >> 42) & 0x1FFFFFu) - kLeafBias, ly = int32_t(uint32_t(key >> 21) & 0x1FFFFFu) - kLeafBias;
  const double a = diamondAngle(double(lx - sp.cx), double(ly - sp.cy));
  int32_t owner = sp.n_ranks - 1; // below the first bound: the last sector wraps around
  for (int32_t r = 0; r < sp.n_ranks; ++r)
    if (a >= sp.bounds[r]) owner = r;
  return owner;
}

// ------------------------------------------------------------------------------------------------
// Update grid of one input source: an open-addressing hash of BRICKS (8^3 leaves = 64^3 voxels), slot ==
// storage. Inside a brick every leaf mask word has a fixed address, so the DDA kernel needs a hash lookup only
// when a ray enters a new brick (~every 50 visits) and addresses everything else arithmetically:
//   word offset in brick = leaf_in_brick * 8 + (x & 7),  leaf_in_brick = ((x>>3)&7)<<6 | ((y>>3)&7)<<3 | ((z>>3)&7)
// A "leaf entry" e = slot * 512 + leaf_in_brick indexes act/val (8 words each) directly.
constexpr int kBrickLeaves = 512;
struct UpdateGrid
{
  uint64_t* bkeys;    // [cap]          brick key = packLeafKey(voxel >> 6 per axis)
  uint64_t* act;      // [cap][512][8]  active masks of the brick's leaves (32 KB per brick)
  uint64_t* val;      // [cap][512][8]  value (hit) masks
  uint32_t* btouched; // [cap]          occupied brick slots in insertion order
  uint32_t* entries;  // [cap*512]      compact list of touched leaf entries (built by compact_leaves_kernel)
  uint32_t* counters; // [0] occupied bricks, [1] leaf entries
  uint32_t cap_mask;  // cap - 1 (cap is a power of two)
};

// leaf key of entry e given its brick key
__host__ __device__ __forceinline__ uint64_t leafKeyOfEntry(uint64_t bkey, uint32_t leaf_in_brick)
{
  const int32_t bx = int32_t(uint32_t(bkey >> 42) & 0x1FFFFFu) - kLeafBias;
  const int32_t by = int32_t(uint32_t(bkey >> 21) & 0x1FFFFFu) - kLeafBias;
  const int32_t bz = int32_t(uint32_t(bkey) & 0x1FFFFFu) - kLeafBias;
  return packLeafKey(bx * 8 + int32_t(leaf_in_brick >> 6), by * 8 + int32_t((leaf_in_brick >> 3) & 7), bz * 8 + int32_t(leaf_in_brick & 7));
}
// brick key and leaf_in_brick of a LEAF key
__host__ __device__ __forceinline__ uint64_t brickKeyOfLeaf(uint64_t lkey, uint32_t& leaf_in_brick)
{
  const int32_t lx = int32_t(uint32_t(lkey >> 42) & 0x1FFFFFu) - kLeafBias;
  const int32_t ly = int32_t(uint32_t(lkey >> 21) & 0x1FFFFFu) - kLeafBias;
  const int32_t lz = int32_t(uint32_t(lkey) & 0x1FFFFFu) - kLeafBias;
  leaf_in_brick = (uint32_t(lx & 7) << 6) | (uint32_t(ly & 7) << 3) | uint32_t(lz & 7);
  return packLeafKey(lx >> 3, ly >> 3, lz >> 3);
}

struct MapTable
{
  uint64_t* hkeys;     // [hcap]
  uint32_t* hvals;     // [hcap]
  uint32_t hcap_mask;
  uint64_t* leaf_keys; // [pool_cap]
  uint64_t* leaf_mask; // [pool_cap][8]
  float* leaf_vals;    // [pool_cap][512]
  uint32_t* leaf_dirty; // [pool_cap] 1 = modified since the last export
  uint32_t pool_cap;
  uint32_t* n_leaves; // device counter
};

// Node levels of the reference's 5-4-3 tree above the leaves, for fast_mode / raytrace (VolumeRayIntersector): open-addressing
// SETS of the 128^3-voxel (k1) and 4096^3-voxel (k2) blocks that hold at least one map leaf; key = packLeafKey(voxel >> 7)
// resp. (voxel >> 12). Built lazily and incrementally from the (append-only) leaf pool.
struct CoarseSets
{
  uint64_t* k1;
  uint64_t* k2;
  uint32_t mask1, mask2;
  uint32_t* counts; // device: [0] keys in k1, [1] keys in k2
};

struct LogOdds
{
  float hit, miss, thres_min, thres_max, max_lo, min_lo;
  // derived (host-computed once per setConfig, see OpenVDB tile-probe notes in DESIGN.md):
  uint32_t miss_probe_flips;     // (0.0f + miss) < thres_min  -> spurious change bit on new leaves (quirk)
  uint32_t miss_probe_no_create; // probe result == (0.0f, false): a miss alone does not create a leaf
  uint32_t replicate_quirk;
};

// device-side counters, one block per map handle
struct Counters
{
  unsigned long long rays, nan_skipped, clipped, visits;          // raycast
  unsigned long long voxel_updates, state_changes, new_leaves;    // update
  unsigned int flags;                                             // kFlag*
  unsigned int ray_cursor;                                        // work-fetch cursor of the DDA kernel
  unsigned int n_change;                                          // change records appended
  unsigned int n_out;                                             // generic output counter (sections, partition)
  unsigned int n_extra;                                           // extra segment slots handed out by prep_rays
  unsigned int n_long;                                            // rays split into segments
  unsigned int max_visits;                                        // longest ray of the last prep_rays launch
  // deferred updateMap (vdbm_insert_async): written by update_guard_kernel, read by the update kernels and the host
  unsigned int deferred_entries;      // touched leaves the queued update kernels process (0 = skipped)
  unsigned int deferred_bricks;       // occupied bricks the queued reset kernel forgets (0 = skipped)
  unsigned int deferred_skip;         // 1: the guard refused (overflow / capacity): the host redoes the scan synchronously
  unsigned int deferred_seen_entries; // touched leaves the guard saw, whatever it decided
};

// One prepared ray (written by prep_rays_kernel, consumed by raycast_dda_kernel). 32 bytes.
struct __align__(16) RayRec
{
  double delta[3];  // |1/dir| per axis, DBL_MAX when dir == 0     (DDA::mDelta)
  uint32_t visits;  // voxels castRayIntoGrid marks: 1 + |dx| + |dy| + |dz| (0 for a zero-length ray)
  uint32_t flags;   // bit0 valid, bit1 clipped (no hit at the end voxel), bit2 zero-length, bits 4..9 step signs
};
constexpr uint32_t kRayValid = 1u, kRayClipped = 2u, kRayZeroLen = 4u;
// step sign of axis a in bits (4+2a, 5+2a): 0 = none, 1 = +1, 2 = -1
__host__ __device__ __forceinline__ int rayStep(uint32_t flags, int axis)
{
  const uint32_t f = (flags >> (4 + 2 * axis)) & 3u;
  return f == 1u ? 1 : (f == 2u ? -1 : 0);
}

// One unit of DDA work: a whole ray, or one SEGMENT of a long ray. A segment starts from the exact DDA state
// (next[3], steps taken per axis) that the sequential traversal passes through at a time threshold t = j/P: all
// crossings with time < t happen before any crossing with time >= t whatever the tie rules, so that state is visited
// by the reference's loop too, and the three axes' crossing-time sequences (v_0 = d/2, v_{k+1} = fl(v_k + d)) are
// independent of each other -> long_ray_segments_kernel can compute them per axis without marking anything. 48 bytes.
struct __align__(16) SegRec
{
  double next[3];   // DDA::mNext at the segment start
  uint32_t m[3];    // steps already taken along each axis (start voxel = origin + step * m)
  uint32_t count;   // voxels this segment marks (0 = nothing to do)
  uint32_t ray;     // index into RaycastArgs::rays
  uint32_t last;    // 1: this segment ends at the ray's end voxel (it alone delivers the hit)
};
// (segment length is a per-scan runtime choice, RaycastArgs::seg_len; 0 = do not split)

// 136-byte exchange / export record of an update-grid leaf
struct LeafRecord
{
  uint64_t key;
  uint64_t active[8];
  uint64_t value[8];
};

struct RaycastArgs
{
  const uint8_t* points; // device, pcl::PointXYZ layout
  uint64_t n;
  uint32_t stride;
  double origin[3];
  int32_t origin_idx[3];
  double range;      // raycast_range
  double resolution;
  double half_res;   // resolution / 2.0
  double inv_res;    // 1.0 / resolution
  RayRec* rays;      // [n]
  SegRec* segs;      // [seg_cap]: segment i < n is ray i's first (or only) segment, the rest are extra segments of long rays
  uint32_t seg_cap;
  uint32_t seg_len;  // target voxel marks per segment; rays with more than 1.5 x seg_len marks are split. 0 = never split
  uint32_t n_segs;   // segments to traverse (n + extra segments in use); set by the host before the DDA kernel
  uint32_t* long_rays; // [n] indices of rays that were split (count in Counters::n_long)
  uint32_t* seg_base;  // [n] first extra segment slot of a split ray
  uint32_t* sort_keys; // [seg_cap] min(segment marks, 2047); 0 = nothing to do
  uint32_t* sort_idx;  // [seg_cap] identity
  const uint32_t* order; // [n_segs] segment indices, longest first (LPT schedule for the DDA kernel)
  const uint32_t* sorted_keys; // [n_segs] the keys in that order (descending)
  int4* ends;          // [n] or nullptr: end voxel of every ray + (bit0 valid | bit1 hit), the scan's "reduced" update
  uint32_t index_mode; // 1: `points` holds such int4 end-voxel records (16-byte stride) instead of world points
  // multi-GPU ray split on the device: only the points whose direction around the scan origin (diamond angle of
  // (x - origin.x, y - origin.y)) falls into sector `sector_rank` of `sector_bounds` are cast; the others are not this rank's
  int32_t sector_n, sector_rank; // sector_n <= 1: no filter
  double sector_bounds[kMaxRanks];
  uint32_t fast_mode;  // 1: castRayIntoGridFast follows (prep_rays only produces the end-voxel records and the ray statistics)
};

// Peer-memory exchange state (device-visible part). inbox layout on every rank, SoA so that every record's 16 mask
// words form ONE aligned 128-byte line (full-line NVLink stores): for region R = parity * n_ranks + sender:
//   masks: uint64 [R][cap][16]  (active[8] | value[8])      keys: uint64 [2 * n_ranks * cap + R * cap + i] after the masks
// ctrl layout: [parity 2][sender n_ranks] uint64 = (epoch << 32) | record count, written by the sender.
struct ExchangePeers
{
  uint64_t* inbox[kMaxRanks];          // peer r's inbox base (mapped into this process), [rank] = own
  unsigned long long* ctrl[kMaxRanks]; // peer r's ctrl base
  uint32_t cap;                        // records per sender region
  int32_t n_ranks, rank;
};
__host__ __device__ __forceinline__ size_t inboxMaskWord(uint32_t region, uint32_t cap, uint32_t i) { return (size_t(region) * cap + i) * 16; }
__host__ __device__ __forceinline__ size_t inboxKeyWord(uint32_t n_regions, uint32_t region, uint32_t cap, uint32_t i)
{
  return size_t(n_regions) * cap * 16 + size_t(region) * cap + i;
}
__host__ __device__ __forceinline__ size_t inboxBytes(uint32_t n_regions, uint32_t cap) { return size_t(n_regions) * cap * 17 * sizeof(uint64_t); }

// ---- launch wrappers (vdbm_kernels.cu) ----------------------------------------------------------------
void launchPrepRays(const RaycastArgs& a, Counters* ctr, cudaStream_t s);
void launchLongRaySegments(const RaycastArgs& a, uint32_t n_long, cudaStream_t s);
// near_act: zero-initialised scratch of nearCopiesBytes() bytes (privatised near-field bricks), left zeroed again
// test_before_set: load the mask word and skip the RED when its bits are already there (scans with heavily overlapping rays)
void launchRaycastDDA(const RaycastArgs& a, UpdateGrid ug, uint64_t* near_act, Counters* ctr, int grid, cudaStream_t s,
                      bool test_before_set = false);
size_t nearCopiesBytes();
int raycastDDAGrid(int device);
int raycastDDABlock(); // threads per DDA CTA
// rebuild ug.entries / counters[1] from the occupied bricks (brick count read on the device).
// cook = true after launchRaycastDDA: the DDA kernel marks Z-SLICE mask words (word z&7 = the 8x8 x-y tile); the compaction
// turns every touched leaf into OpenVDB's x-slice layout in place, so every other kernel sees x-slice words only.
void launchCompactLeaves(UpdateGrid ug, cudaStream_t s, bool cook = false);
// the inverse (same transform) for the listed leaves: needed before the DDA kernel marks into a grid that holds data
void launchUncookLeaves(UpdateGrid ug, uint32_t n_entries, cudaStream_t s);
// resolve (K2a) + apply (K2b); `resolved` is a device scratch array of >= n_entries u32
void launchApplyUpdate(UpdateGrid ug, MapTable mt, LogOdds lo, uint32_t* resolved, LeafRecord* change_out, uint32_t change_cap,
                       Counters* ctr, uint32_t n_entries, cudaStream_t s);
// updateMap queued WITHOUT the host knowing the number of touched leaves: a one-thread guard kernel checks capacities and
// flags on the device and publishes the counts; resolve / apply / reset then run (or turn into no-ops). expected_entries
// only sizes the resolve grid (it is grid-stride).
void launchApplyUpdateDeferred(UpdateGrid ug, MapTable mt, LogOdds lo, uint32_t* resolved, uint32_t resolved_cap, Counters* ctr,
                               uint32_t expected_entries, cudaStream_t s);
// empty the grid after its entries were consumed (entry masks are zeroed by the consumer): reset brick keys + counters
void launchResetBricks(UpdateGrid ug, uint32_t n_bricks, cudaStream_t s);
// zero the masks of all listed entries (used by reset / source re-add; consumers zero masks themselves)
void launchClearEntries(UpdateGrid ug, uint32_t n_entries, cudaStream_t s);
void launchRehashUpdate(UpdateGrid old_g, uint32_t old_bricks, UpdateGrid new_g, Counters* ctr, cudaStream_t s);
void launchRehashMap(MapTable mt, uint32_t n_leaves, Counters* ctr, cudaStream_t s);
void launchEntryKeys(UpdateGrid ug, uint32_t n_entries, uint64_t* out_keys, uint32_t* out_entries, cudaStream_t s);
void launchGatherUpdate(UpdateGrid ug, uint32_t n, const uint64_t* keys, const uint32_t* entries, LeafRecord* out, cudaStream_t s);
void launchImportUpdate(UpdateGrid ug, const LeafRecord* recs, uint64_t n, Counters* ctr, cudaStream_t s);
// same from leaf arrays in the ABI's host layout (value may be nullptr = all false); leaf origins outside the voxel
// range raise kFlagCoordRange and are skipped
void launchImportUpdateSoA(UpdateGrid ug, const int32_t* origins, const uint64_t* active, const uint64_t* value, uint64_t n, Counters* ctr,
                           cudaStream_t s);
void launchGatherMap(MapTable mt, uint32_t n, const uint32_t* leaf_idx, int32_t* origins, uint64_t* mask, float* vals, cudaStream_t s);
void launchCollectDirty(MapTable mt, uint32_t n_leaves, uint32_t* out_idx, Counters* ctr, cudaStream_t s); // appends via ctr->n_out, clears flags
void launchMarkDirty(MapTable mt, const uint32_t* idx, uint32_t n, cudaStream_t s); // sets the flags of the listed leaves again
void launchSection(MapTable mt, uint32_t n_leaves, const int32_t bbmin[3], const int32_t bbmax[3], int full, int result_float,
                   uint64_t* out_keys, uint64_t* out_active, uint64_t* out_valmask, float* out_vals, uint32_t out_cap,
                   Counters* ctr, cudaStream_t s);
void launchProbe(MapTable mt, int32_t x, int32_t y, int32_t z, float* out_val, int32_t* out_active, cudaStream_t s);
// pass 0: count entries per owner rank; pass 1: scatter records grouped by rank and zero the entry masks
void launchPartition(UpdateGrid ug, uint32_t n_entries, ShardPlan plan, uint32_t* rank_counts, uint32_t* rank_cursor,
                     LeafRecord* out, int pass, cudaStream_t s);
void launchKeysFromIdx(const uint64_t* keys, const uint32_t* idx, uint32_t n, uint64_t* out_keys, uint32_t* out_idx, cudaStream_t s);
// perm: optional permutation (output row i = record perm[i])
void launchSplitRecords(const LeafRecord* recs, const uint32_t* perm, uint32_t n, int32_t* origins, uint64_t* active, uint64_t* value, cudaStream_t s);
void launchRecordKeys(const LeafRecord* recs, uint32_t n, uint64_t* keys, uint32_t* idx, cudaStream_t s);
void launchPermuteSection(uint32_t n, const uint64_t* sorted_keys, const uint32_t* perm, const uint64_t* in_active, const uint64_t* in_valmask,
                          const float* in_vals, int32_t* origins, uint64_t* out_active, uint64_t* out_valmask, float* out_vals, cudaStream_t s);
// receiver side of remote mapping (applyMapSection*)
void launchSectionDeactivate(MapTable mt, uint32_t n_leaves, const int32_t bbmin[3], const int32_t bbmax[3], cudaStream_t s);
void launchSectionActivate(MapTable mt, const uint64_t* keys, const uint64_t* active, uint32_t n, Counters* ctr, cudaStream_t s);
void launchSectionApplyGrid(MapTable mt, const uint64_t* keys, const uint64_t* active, const float* values, uint32_t n, Counters* ctr, cudaStream_t s);
void launchSectionTileQuirk(MapTable mt, const uint64_t* blocks, uint32_t n_blocks, int level, const uint64_t* present, uint32_t n_present, cudaStream_t s);
// fused bin + send over peer memory; cursors = device scratch [kMaxRanks] (zeroed by the wrapper)
void launchPushUpdate(UpdateGrid ug, uint32_t n_entries, ExchangePeers px, ShardPlan plan, uint32_t parity, uint32_t epoch, uint32_t* cursors,
                      Counters* ctr, cudaStream_t s);
// order-independent 64-bit checksum of the map: sum over leaves of hash(key, active mask, value bits); out[0] += sum, out[1] += leaves
void launchMapChecksum(MapTable mt, uint32_t n_leaves, unsigned long long* out2, cudaStream_t s);
void launchWaitPeers(const unsigned long long* ctrl, int32_t n_ranks, uint32_t parity, uint32_t epoch, uint32_t* counts_out, Counters* ctr,
                     cudaStream_t s);
// after launchWaitPeers: OR all inbox records of this parity into the grid
void launchPullUpdate(UpdateGrid ug, const uint64_t* inbox, const unsigned long long* ctrl, uint32_t cap, int32_t n_ranks,
                      uint32_t parity, uint32_t epoch, uint32_t* counts_out, Counters* ctr, cudaStream_t s);
// remote-mapping deltas, direct edits, artificial areas
void launchMarkEnds(const int4* ends, uint64_t n, UpdateGrid g, Counters* ctr, cudaStream_t s);
void launchExpandEnds(UpdateGrid g, uint32_t n_entries, int4* out, uint32_t out_cap, Counters* ctr, cudaStream_t s); // appends via ctr->n_out
void launchPointsToEnds(const uint8_t* points, uint64_t n, uint32_t stride, double inv_res, int occupied, int4* out, Counters* ctr, cudaStream_t s);
void launchOverwrite(UpdateGrid g, uint32_t n_entries, MapTable mt, LogOdds lo, Counters* ctr, cudaStream_t s);
void launchWallDDA(const int32_t* d_walls, uint32_t n_walls, int32_t neg_index, int32_t pos_index, UpdateGrid g, Counters* ctr, cudaStream_t s);
void launchGridActivate(UpdateGrid g, uint32_t n_entries, MapTable mt, Counters* ctr, cudaStream_t s);
void launchRestoreState(UpdateGrid g, uint32_t n_entries, MapTable mt, LogOdds lo, Counters* ctr, cudaStream_t s);
// fast_mode / raytrace
void launchCoarseInsert(MapTable mt, uint32_t from, uint32_t to, CoarseSets cs, Counters* ctr, cudaStream_t s);
void launchActiveBBox(MapTable mt, uint32_t n_leaves, int32_t* out6, cudaStream_t s); // out6 pre-set to INT_MAX x3, INT_MIN x3
void launchRaycastFast(const RaycastArgs& a, UpdateGrid g, MapTable mt, CoarseSets cs, const int32_t* bbox6, uint32_t map_empty, Counters* ctr,
                       cudaStream_t s);
void launchRaytrace(uint64_t n, const double* origins, const double* directions, const double* max_lengths, double res, double inv_res, MapTable mt,
                    CoarseSets cs, const int32_t* bbox6, uint32_t map_empty, int32_t* success, double* end_points, Counters* ctr, cudaStream_t s);
uint32_t launchCount(); // kernels of this library launched by this process
// CUB radix sort (descending) of (visit count, ray index) on key bits [3, 11) (one radix pass); returns temp bytes when d_temp == nullptr
size_t sortRaysByLength(void* d_temp, size_t temp_bytes, const uint32_t* keys_in, uint32_t* keys_out, const uint32_t* idx_in,
                        uint32_t* idx_out, uint32_t n, cudaStream_t s);
// CUB radix sort of 32-bit keys (ascending); returns bytes of temp storage needed when d_temp == nullptr
size_t sortKeys32(void* d_temp, size_t temp_bytes, const uint32_t* keys_in, uint32_t* keys_out, uint32_t n, cudaStream_t s);
// CUB radix sort of (key, idx) pairs; returns bytes of temp storage needed when d_temp == nullptr
size_t sortPairs(void* d_temp, size_t temp_bytes, const uint64_t* keys_in, uint64_t* keys_out, const uint32_t* idx_in,
                 uint32_t* idx_out, uint32_t n, cudaStream_t s);

} // namespace vdbm
