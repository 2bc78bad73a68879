// vdbm_device.cuh — shared device/host types of libvdbm_b200 (sm_100a only).
//
// Data layout in HBM:
//   update grid (one per input source; replaces openvdb Tree4<bool,1,4,3>, VDBMapping.hpp:89,97)
//     open-addressing hash, slot == storage:  keys[C] u64 | active[C][8] u64 | value[C][8] u64
//     + touched[C] u32 : compact list of occupied slots in insertion order (drives the update kernel)
//   map (replaces openvdb Tree4<float,5,4,3> = FloatTree, VDBMapping.hpp:88)
//     open-addressing hash  hkeys[H] u64 -> hvals[H] u32 (leaf index)
//     leaf pool (SoA):  leaf_keys[P] u64 | leaf_mask[P][8] u64 | leaf_vals[P][512] f32 | leaf_dirty[P] u32
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

// ------------------------------------------------------------------------------------------------
struct UpdateTable
{
  uint64_t* keys;    // [cap]
  uint64_t* active;  // [cap][8]
  uint64_t* value;   // [cap][8]
  uint32_t* touched; // [cap]
  uint32_t* n_touched; // device counter
  uint32_t cap_mask; // cap - 1 (cap is a power of two)
};

struct MapTable
{
  uint64_t* hkeys;     // [hcap]
  uint32_t* hvals;     // [hcap]
  uint32_t hcap_mask;
  uint64_t* leaf_keys; // [pool_cap]
  uint64_t* leaf_mask; // [pool_cap][8]
  float* leaf_vals;    // [pool_cap][512]
  uint32_t* leaf_dirty; // [pool_cap] 1 = in the dirty list
  uint32_t* dirty_list; // [pool_cap]
  uint32_t pool_cap;
  uint32_t* n_leaves; // device counter
  uint32_t* n_dirty;  // device counter
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
};

// One prepared ray (written by prep_rays_kernel, consumed by raycast_dda_kernel). 48 bytes.
struct __align__(16) RayRec
{
  double delta[3];  // |1/dir| per axis, DBL_MAX when dir == 0     (DDA::mDelta)
  int32_t end[3];   // end voxel
  uint32_t flags;   // bit0 valid, bit1 clipped (no hit at the end voxel), bit2 zero-length
};
constexpr uint32_t kRayValid = 1u, kRayClipped = 2u, kRayZeroLen = 4u;

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
};

// ---- launch wrappers (vdbm_kernels.cu) ----------------------------------------------------------------
void launchPrepRays(const RaycastArgs& a, Counters* ctr, cudaStream_t s);
void launchRaycastDDA(const RaycastArgs& a, UpdateTable ut, Counters* ctr, int grid, cudaStream_t s);
int raycastDDAGrid(int device);
void launchApplyUpdate(UpdateTable ut, MapTable mt, LogOdds lo, LeafRecord* change_out, uint32_t change_cap, Counters* ctr,
                       uint32_t n_touched_hint, cudaStream_t s);
void launchRehashUpdate(UpdateTable old_t, uint32_t old_n, UpdateTable new_t, Counters* ctr, cudaStream_t s);
void launchRehashMap(MapTable mt, uint32_t n_leaves, Counters* ctr, cudaStream_t s);
void launchGatherUpdate(UpdateTable ut, uint32_t n, const uint32_t* order, LeafRecord* out, cudaStream_t s);
void launchImportUpdate(UpdateTable ut, const LeafRecord* recs, uint64_t n, Counters* ctr, cudaStream_t s);
void launchClearUpdate(UpdateTable ut, uint32_t n, cudaStream_t s);
void launchGatherMap(MapTable mt, uint32_t n, const uint32_t* leaf_idx, int32_t* origins, uint64_t* mask, float* vals,
                     int clear_dirty, cudaStream_t s);
void launchSection(MapTable mt, uint32_t n_leaves, const int32_t bbmin[3], const int32_t bbmax[3], int full, int result_float,
                   uint64_t* out_keys, uint64_t* out_active, uint64_t* out_valmask, float* out_vals, uint32_t out_cap,
                   Counters* ctr, cudaStream_t s);
void launchProbe(MapTable mt, int32_t x, int32_t y, int32_t z, float* out_val, int32_t* out_active, cudaStream_t s);
void launchPartition(UpdateTable ut, uint32_t n, int32_t n_ranks, uint32_t* rank_counts, uint32_t* rank_offsets,
                     LeafRecord* out, int pass, cudaStream_t s);
void launchKeysFromSlots(const uint64_t* keys, const uint32_t* slots, uint32_t n, uint64_t* out_keys, uint32_t* out_idx,
                         cudaStream_t s);
void launchIota(uint32_t* out, uint32_t n, cudaStream_t s);
void launchUnpackOrigins(const uint64_t* keys, uint32_t n, int32_t* origins, cudaStream_t s);
void launchSplitRecords(const LeafRecord* recs, uint32_t n, int32_t* origins, uint64_t* active, uint64_t* value, cudaStream_t s);
uint32_t launchCount(); // kernels of this library launched by this process
// CUB radix sort of (key, idx) pairs; returns bytes of temp storage needed when d_temp == nullptr
size_t sortPairs(void* d_temp, size_t temp_bytes, const uint64_t* keys_in, uint64_t* keys_out, const uint32_t* idx_in,
                 uint32_t* idx_out, uint32_t n, cudaStream_t s);

} // namespace vdbm
