// index_prep.cuh -- integer preparation (I1): stable sorts of the observations by point and
// by camera, CSR offsets, the (a, b) observation-pair lists that feed every off-diagonal
// block of the reduced camera system, and its BSR pattern.  Replaces torch.unique / bae's
// index tracing (bundle_adjustment.py:108-109).  Everything here is bit-exact testable
// against numpy (stable argsort, searchsorted).
#pragma once
#include "common.cuh"

namespace isfm {

#ifndef ISFM_SPMV_CHUNK
#define ISFM_SPMV_CHUNK 48
#endif
constexpr int SPMV_CHUNK = ISFM_SPMV_CHUNK;   // largest mat-vec work unit in slots; multiple of 4

struct ObsIndex {
  int64_t n_cam = 0, n_pt = 0, n_obs = 0;
  // point-major order ("sorted position" a = 0..n_obs-1)
  DeviceBuffer<int32_t> obs_perm;   // [n_obs] sorted position -> original observation
  DeviceBuffer<int32_t> pt_of;      // [n_obs] point of sorted position (non-decreasing)
  DeviceBuffer<int32_t> cam_of;     // [n_obs] camera of sorted position
  DeviceBuffer<int32_t> pt_off;     // [n_pt + 1]
  // camera-major view
  DeviceBuffer<int32_t> cam_perm;   // [n_obs] camera-major rank -> sorted position
  DeviceBuffer<int32_t> cam_off;    // [n_cam + 1]
};

// Reduced camera system pattern.  E = sum_p Hcp Hpp^-1 Hcp^T is symmetric: only its upper
// triangle (block (i, j), i <= j, every diagonal block included) is stored, as BSR sorted by
// (i, j).  The lists of observation pairs feeding each block are grouped in the same order.
// For the symmetric mat-vec every strictly-upper entry also has a position in the row-major
// ordering of the LOWER triangle (`tpos`), where its transposed product is deposited.
struct SchurPattern {
  int64_t n_pairs = 0;   // total (a, b) pairs
  int64_t n_lists = 0;   // unique (i, j), i <= j, with at least one pair
  int64_t nnzu = 0;      // stored slots: strictly-upper blocks + n_cam diagonal blocks + row padding
                         // (each row padded to a multiple of 4 slots; padding: col = row, zeros)
  int64_t n_blocks = 0;  // real blocks among them
  int64_t n_off = 0;     // strictly-upper blocks (= lower-triangle entries)
  DeviceBuffer<uint64_t> pairs;      // [n_pairs] (a << 32) | b, cam(a) <= cam(b), grouped by list
  DeviceBuffer<int64_t> list_off;    // [n_lists + 1]
  DeviceBuffer<int32_t> list_slot;   // [n_lists] upper slot of block (i, j)
  DeviceBuffer<uint8_t> list_diag;   // [n_lists] 1 if i == j (accumulate onto the camera pass)
  DeviceBuffer<int32_t> list_order;  // [n_lists] list ids by decreasing length (stable)
  DeviceBuffer<int32_t> urow_ptr;    // [n_cam + 1] upper BSR
  DeviceBuffer<int32_t> ucol;        // [nnzu]
  DeviceBuffer<int32_t> tpos;        // [nnzu] position in the lower ordering, -1 for diagonal blocks
  DeviceBuffer<int32_t> lrow_ptr;    // [n_cam + 1] lower triangle, row-major
  DeviceBuffer<int32_t> diag_slot;   // [n_cam]
  // mat-vec work units: every upper row is cut into chunks of <= SPMV_CHUNK blocks so that the
  // long rows of a dense system do not serialise on one CTA
  int64_t n_chunks = 0;
  DeviceBuffer<int32_t> chunk_row;   // [n_chunks]
  DeviceBuffer<int32_t> chunk_beg;   // [n_chunks] first upper slot of the chunk
  DeviceBuffer<int32_t> chunk_ptr;   // [n_cam + 1] first chunk of each row
  std::vector<int32_t> h_chunk_beg;  // host copy of chunk_beg (splitting the units across ranks)
  int unit_slots = SPMV_CHUNK;       // slots per mat-vec work unit of THIS pattern (<= SPMV_CHUNK, see build_schur_pattern)
  // Stage table of the persistent PCG kernel: every unit cut into stages of <= stage_blocks slots,
  // in unit order; one int4 per stage = {row, first slot, slots | first-of-unit << 8 | last-of-unit << 9, unit}.
  // A warp of that kernel walks a contiguous range of this table as ONE continuous cp.async stream.
  int stage_blocks = 0;
  int64_t n_stages = 0;
  DeviceBuffer<int4> stages;            // [n_stages]
  DeviceBuffer<int32_t> unit_stage_ptr; // [n_chunks + 1] first stage of each unit
  std::vector<int32_t> h_unit_stage_ptr;
};

// cam_idx / pt_idx: device int32 [n_obs] in the caller's order.
void build_obs_index(ObsIndex& ix, int64_t n_cam, int64_t n_pt, int64_t n_obs, const int32_t* cam_idx,
                     const int32_t* pt_idx, cudaStream_t stream, KernelTimers& kt);
// Optional hook called once the local pair lists are known: may return extra block keys
// (i * n_cam + j; anything with i >= j or ~0 is ignored, duplicates are fine) that must exist in
// the pattern even without local pairs -- the multi-rank solver passes the other ranks' keys so
// that every rank builds the UNION pattern (zero blocks where it has no pairs).
struct PatternKeyHook {
  virtual ~PatternKeyHook() {}
  int matvec_share = 1;   // set by extra_keys: ranks that will share the mat-vec of this pattern (union pattern: world)
  virtual int64_t extra_keys(const uint64_t* list_key, int64_t n_lists, int64_t n_cam, DeviceBuffer<uint64_t>& out,
                             cudaStream_t stream) = 0;
};
// stage_blocks > 0: also build the stage table (slots per stage of the caller's mat-vec kernel).
// grid_warps > 0 (warps of the persistent PCG grid): the work unit shrinks from SPMV_CHUNK to 24 or
// 12 slots until every warp has >= 8 units per mat-vec -- small per-rank shares (8-way strong
// scaling) are then balanced at stage granularity instead of losing 20 % to the last unit.
void build_schur_pattern(SchurPattern& sp, const ObsIndex& ix, cudaStream_t stream, KernelTimers& kt,
                         PatternKeyHook* hook = nullptr, int stage_blocks = 0, int grid_warps = 0);
// number of distinct valid strictly-upper keys (i < j) in `keys` (device, n entries; sorts a copy)
int64_t count_unique_upper_keys(const uint64_t* keys, int64_t n, int64_t n_cam, cudaStream_t stream);
// Two 48-bit order-independent hashes of the upper BSR pattern (row pointers and columns): equal
// on two ranks <=> (with overwhelming probability) identical patterns.  Synchronises the stream.
void schur_pattern_signature(const SchurPattern& sp, int64_t n_cam, cudaStream_t stream, uint64_t sig_out[2]);

}  // namespace isfm
