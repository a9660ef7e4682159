// index_prep.cu -- see index_prep.cuh.
#include <cub/cub.cuh>
#include <vector>

#include "index_prep.cuh"

namespace isfm {
namespace {

constexpr int TPB = 256;

__global__ void iota_kernel(int32_t* v, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) v[i] = (int32_t)i;
}

__global__ void validate_kernel(const int32_t* cam_idx, const int32_t* pt_idx, int64_t n, int32_t n_cam,
                                int32_t n_pt, int* bad) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int32_t c = cam_idx[i], p = pt_idx[i];
  if (c < 0 || c >= n_cam || p < 0 || p >= n_pt) *bad = 1;
}

__global__ void gather_i32(const int32_t* src, const int32_t* idx, int32_t* dst, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}

// off[k] = first position whose sorted key >= k, for k = 0..n_keys  (searchsorted, side=left)
__global__ void lower_bound_kernel(const int32_t* __restrict__ sorted, int64_t n, int32_t* off, int64_t n_keys) {
  int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k > n_keys) return;
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (sorted[mid] < (int32_t)k) lo = mid + 1; else hi = mid;
  }
  off[k] = (int32_t)lo;
}

// pairs emitted by sorted position a: every b > a of the same point; twice if same camera.
__global__ void pair_count_kernel(const int32_t* __restrict__ pt_of, const int32_t* __restrict__ cam_of,
                                  const int32_t* __restrict__ pt_off, int64_t n, int64_t* cnt) {
  int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (a >= n) return;
  int32_t end = pt_off[pt_of[a] + 1];
  int32_t ca = cam_of[a];
  int64_t c = 0;
  for (int32_t b = (int32_t)a + 1; b < end; ++b) c += (cam_of[b] == ca) ? 2 : 1;
  cnt[a] = c;
}

__global__ void pair_fill_kernel(const int32_t* __restrict__ pt_of, const int32_t* __restrict__ cam_of,
                                 const int32_t* __restrict__ pt_off, int64_t n, const int64_t* __restrict__ off,
                                 int64_t n_cam, uint64_t* keys, uint64_t* vals) {
  int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (a >= n) return;
  int32_t end = pt_off[pt_of[a] + 1];
  int64_t ca = cam_of[a];
  int64_t o = off[a];
  for (int32_t b = (int32_t)a + 1; b < end; ++b) {
    int64_t cb = cam_of[b];
    uint64_t ab = ((uint64_t)a << 32) | (uint32_t)b, ba = ((uint64_t)(uint32_t)b << 32) | (uint32_t)a;
    if (ca < cb) { keys[o] = (uint64_t)(ca * n_cam + cb); vals[o] = ab; ++o; }
    else if (ca > cb) { keys[o] = (uint64_t)(cb * n_cam + ca); vals[o] = ba; ++o; }
    else { keys[o] = (uint64_t)(ca * n_cam + cb); vals[o] = ab; ++o; keys[o] = (uint64_t)(ca * n_cam + cb); vals[o] = ba; ++o; }
  }
}

// upper-BSR entry keys: every list (i <= j) and every camera's diagonal (i, i); tag = list id
// or -(i + 1).  Diagonal lists and diagonal entries share a key: a later pass keeps one.
__global__ void upper_keys_kernel(const uint64_t* __restrict__ list_key, int64_t n_lists, int64_t n_cam,
                                  uint64_t* keys, int64_t* tags) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t < n_lists) {
    uint64_t k = list_key[t];
    bool diag = (k / (uint64_t)n_cam) == (k % (uint64_t)n_cam);
    keys[t] = diag ? ~0ull : k;          // diagonal lists own no entry: dropped after the sort
    tags[t] = t;
  } else if (t < n_lists + n_cam) {
    int64_t i = t - n_lists;
    keys[t] = (uint64_t)i * (uint64_t)n_cam + (uint64_t)i;
    tags[t] = -(i + 1);
  }
}

constexpr int64_t EXTRA_TAG = INT64_MIN;   // entry that exists only because another rank has pairs for it

// extra (foreign) keys appended behind the local entries: strictly-upper ones survive
__global__ void extra_keys_kernel(const uint64_t* __restrict__ extra, int64_t n_extra, int64_t n_cam, uint64_t* keys, int64_t* tags) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= n_extra) return;
  const uint64_t k = extra[t];
  const bool ok = k != ~0ull && k < (uint64_t)n_cam * (uint64_t)n_cam && (k / (uint64_t)n_cam) < (k % (uint64_t)n_cam);
  keys[t] = ok ? k : ~0ull;
  tags[t] = EXTRA_TAG;
}

// after a STABLE sort by key: every entry equal to its predecessor is a duplicate (local entries
// were ahead of the extras in the input, so the survivor of a run is the local one)
__global__ void dedupe_keys_kernel(const uint64_t* __restrict__ sorted, int64_t n, uint64_t* out) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= n) return;
  out[t] = (t > 0 && sorted[t] == sorted[t - 1]) ? ~0ull : sorted[t];
}

__global__ void count_valid_kernel(const uint64_t* __restrict__ keys, int64_t n, unsigned long long* count) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t < n && keys[t] != ~0ull) atomicAdd(count, 1ull);  // set-up only; integer, order-free
}

__global__ void upper_scatter_kernel(const uint64_t* __restrict__ keys, const int64_t* __restrict__ tags, int64_t nnzu,
                                     int64_t n_cam, int32_t* ucol, int32_t* row_of, int32_t* list_slot, int32_t* diag_slot) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= nnzu) return;
  uint64_t k = keys[e];
  ucol[e] = (int32_t)(k % (uint64_t)n_cam);
  row_of[e] = (int32_t)(k / (uint64_t)n_cam);
  int64_t tag = tags[e];
  if (tag == EXTRA_TAG) return;   // block of the union pattern without local pairs
  if (tag < 0) diag_slot[-tag - 1] = (int32_t)e;
  else list_slot[tag] = (int32_t)e;
}

__global__ void diag_list_slot_kernel(const uint64_t* __restrict__ list_key, int64_t n_lists, int64_t n_cam,
                                      const int32_t* __restrict__ diag_slot, int32_t* list_slot, uint8_t* list_diag) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= n_lists) return;
  uint64_t k = list_key[t];
  uint64_t i = k / (uint64_t)n_cam, j = k % (uint64_t)n_cam;
  list_diag[t] = (i == j) ? 1 : 0;
  if (i == j) list_slot[t] = diag_slot[i];
}

// lower ordering: key (j, i) for each strictly-upper entry e = (i, j); diagonal entries get ~0
__global__ void lower_keys_kernel(const int32_t* __restrict__ row_of, const int32_t* __restrict__ ucol, int64_t nnzu,
                                  int64_t n_cam, uint64_t* keys, int32_t* vals) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= nnzu) return;
  int32_t i = row_of[e], j = ucol[e];
  keys[e] = (i == j) ? ~0ull : (uint64_t)j * (uint64_t)n_cam + (uint64_t)i;
  vals[e] = (int32_t)e;
}

__global__ void lower_scatter_kernel(const uint64_t* __restrict__ keys_sorted, const int32_t* __restrict__ vals_sorted,
                                     int64_t nnzu, int64_t n_off, int64_t n_cam, int32_t* tpos, int32_t* lrow_of) {
  int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= nnzu) return;
  int32_t e = vals_sorted[k];
  if (k < n_off) { tpos[e] = (int32_t)k; lrow_of[k] = (int32_t)(keys_sorted[k] / (uint64_t)n_cam); }
  else tpos[e] = -1;
}

// Row padding: every upper row starts at a slot that is a multiple of 4 blocks, so that any
// run of 4 blocks (4 D^2 elements) is 16-byte aligned for 128-bit streaming.
__global__ void pad_remap_kernel(int64_t nnzu, const int32_t* __restrict__ row_of, const int32_t* __restrict__ up,
                                 const int32_t* __restrict__ upp, const int32_t* __restrict__ ucol,
                                 const int32_t* __restrict__ tpos, int32_t* ucol_p, int32_t* tpos_p, int32_t* slot_of) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= nnzu) return;
  const int32_t row = row_of[e];
  const int32_t s = upp[row] + ((int32_t)e - up[row]);
  ucol_p[s] = ucol[e];
  tpos_p[s] = tpos[e];
  slot_of[e] = s;
}

__global__ void pad_fill_kernel(int64_t n_cam, const int32_t* __restrict__ up, const int32_t* __restrict__ upp,
                                int32_t* ucol_p, int32_t* tpos_p) {
  int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (row >= n_cam) return;
  for (int32_t s = upp[row] + (up[row + 1] - up[row]); s < upp[row + 1]; ++s) { ucol_p[s] = (int32_t)row; tpos_p[s] = -1; }
}

__global__ void remap_slots_kernel(int64_t n, int32_t* slots, const int32_t* __restrict__ slot_of) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t < n) slots[t] = slot_of[slots[t]];
}

// Visiting order of the pair lists: row-major windows of LIST_WINDOW lists (locality of the
// a-side records), decreasing length inside a window (equal trip counts inside a warp).
constexpr int LIST_WINDOW = 4096;   // ~2 block rows of a 1.8 k-camera dense system
__global__ void list_len_key_kernel(const int64_t* __restrict__ list_off, int64_t n_lists, uint64_t* key, int32_t* id) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= n_lists) return;
  const uint64_t len = (uint64_t)(list_off[t + 1] - list_off[t]);
  key[t] = ((uint64_t)(t / LIST_WINDOW) << 32) | (0xffffffffull - (len < 0xffffffffull ? len : 0xffffffffull));
  id[t] = (int32_t)t;
}

// stage table: flag (bit 10) the stages whose columns are consecutive
__global__ void stage_contiguity_kernel(int4* stages, int64_t n_stages, const int32_t* __restrict__ ucol) {
  int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (g >= n_stages) return;
  int4 m = stages[g];
  const int nb = m.z & 0xff;
  bool contig = nb > 0;
  const int j0 = nb > 0 ? ucol[m.y] : 0;
  for (int b = 1; b < nb && contig; ++b) contig = ucol[m.y + b] == j0 + b;
  if (contig) { m.z |= 1 << 10; stages[g] = m; }
}

int bits_for(uint64_t max_value) {
  int b = 1;
  while (b < 64 && (max_value >> b) != 0) ++b;
  return b;
}

template <typename K, typename V>
void sort_pairs(const K* k_in, K* k_out, const V* v_in, V* v_out, int64_t n, int end_bit, cudaStream_t s) {
  size_t bytes = 0;
  ISFM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, k_in, k_out, v_in, v_out, n, 0, end_bit, s));
  DeviceBuffer<uint8_t> tmp;
  tmp.alloc(bytes);
  ISFM_CUDA(cub::DeviceRadixSort::SortPairs(tmp.get(), bytes, k_in, k_out, v_in, v_out, n, 0, end_bit, s));
  ISFM_CUDA(cudaStreamSynchronize(s));  // tmp is freed at scope exit
}

}  // namespace

void build_obs_index(ObsIndex& ix, int64_t n_cam, int64_t n_pt, int64_t n_obs, const int32_t* cam_idx,
                     const int32_t* pt_idx, cudaStream_t s, KernelTimers& kt) {
  ISFM_REQUIRE(n_cam > 0 && n_pt > 0 && n_obs > 0, ISFM_EINVAL, "empty problem");
  ISFM_REQUIRE(n_obs < (1ll << 31) && n_pt < (1ll << 31) && n_cam < (1ll << 31), ISFM_EINVAL, "sizes must fit int32");
  ix.n_cam = n_cam; ix.n_pt = n_pt; ix.n_obs = n_obs;
  TimerScope ts(kt, T_INDEX_PREP);
  const int g = div_up(n_obs, TPB);
  DeviceBuffer<int> bad; bad.alloc(1); bad.zero(s);
  validate_kernel<<<g, TPB, 0, s>>>(cam_idx, pt_idx, n_obs, (int32_t)n_cam, (int32_t)n_pt, bad.get());
  int h_bad = 0;
  ISFM_CUDA(cudaMemcpyAsync(&h_bad, bad.get(), sizeof(int), cudaMemcpyDeviceToHost, s));
  ISFM_CUDA(cudaStreamSynchronize(s));
  ISFM_REQUIRE(h_bad == 0, ISFM_EINVAL, "camera / point index out of range");

  DeviceBuffer<int32_t> iota, tmp_keys;
  iota.alloc(n_obs); tmp_keys.alloc(n_obs);
  iota_kernel<<<g, TPB, 0, s>>>(iota.get(), n_obs);
  // stable sort by point
  ix.obs_perm.alloc(n_obs); ix.pt_of.alloc(n_obs); ix.cam_of.alloc(n_obs);
  sort_pairs(pt_idx, ix.pt_of.get(), iota.get(), ix.obs_perm.get(), n_obs, bits_for((uint64_t)n_pt), s);
  gather_i32<<<g, TPB, 0, s>>>(cam_idx, ix.obs_perm.get(), ix.cam_of.get(), n_obs);
  ix.pt_off.alloc(n_pt + 1);
  lower_bound_kernel<<<div_up(n_pt + 1, TPB), TPB, 0, s>>>(ix.pt_of.get(), n_obs, ix.pt_off.get(), n_pt);
  // stable sort of the point-major positions by camera
  ix.cam_perm.alloc(n_obs);
  sort_pairs(ix.cam_of.get(), tmp_keys.get(), iota.get(), ix.cam_perm.get(), n_obs, bits_for((uint64_t)n_cam), s);
  ix.cam_off.alloc(n_cam + 1);
  lower_bound_kernel<<<div_up(n_cam + 1, TPB), TPB, 0, s>>>(tmp_keys.get(), n_obs, ix.cam_off.get(), n_cam);
  ISFM_CUDA(cudaGetLastError());
  ISFM_CUDA(cudaStreamSynchronize(s));
}

int64_t count_unique_upper_keys(const uint64_t* keys, int64_t n, int64_t n_cam, cudaStream_t s) {
  if (n <= 0) return 0;
  DeviceBuffer<uint64_t> a, b; DeviceBuffer<int64_t> ta, tb;
  a.alloc(n); b.alloc(n); ta.alloc(n); tb.alloc(n);
  extra_keys_kernel<<<div_up(n, TPB), TPB, 0, s>>>(keys, n, n_cam, a.get(), ta.get());
  sort_pairs(a.get(), b.get(), ta.get(), tb.get(), n, 64, s);
  dedupe_keys_kernel<<<div_up(n, TPB), TPB, 0, s>>>(b.get(), n, a.get());
  DeviceBuffer<unsigned long long> valid; valid.alloc(1); valid.zero(s);
  count_valid_kernel<<<div_up(n, TPB), TPB, 0, s>>>(a.get(), n, valid.get());
  unsigned long long c = 0;
  ISFM_CUDA(cudaMemcpyAsync(&c, valid.get(), sizeof c, cudaMemcpyDeviceToHost, s));
  ISFM_CUDA(cudaStreamSynchronize(s));
  return (int64_t)c;
}

void build_schur_pattern(SchurPattern& sp, const ObsIndex& ix, cudaStream_t s, KernelTimers& kt, PatternKeyHook* hook, int stage_blocks, int grid_warps) {
  TimerScope ts(kt, T_INDEX_PREP);
  const int64_t n = ix.n_obs, n_cam = ix.n_cam;
  const int g = div_up(n, TPB);
  // 1. pair counts -> offsets
  DeviceBuffer<int64_t> cnt, off;
  cnt.alloc(n + 1); off.alloc(n + 1);
  ISFM_CUDA(cudaMemsetAsync(cnt.get(), 0, (n + 1) * sizeof(int64_t), s));
  pair_count_kernel<<<g, TPB, 0, s>>>(ix.pt_of.get(), ix.cam_of.get(), ix.pt_off.get(), n, cnt.get());
  {
    size_t bytes = 0;
    ISFM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, cnt.get(), off.get(), n + 1, s));
    DeviceBuffer<uint8_t> tmp; tmp.alloc(bytes);
    ISFM_CUDA(cub::DeviceScan::ExclusiveSum(tmp.get(), bytes, cnt.get(), off.get(), n + 1, s));
    ISFM_CUDA(cudaStreamSynchronize(s));
  }
  int64_t n_pairs = 0;
  ISFM_CUDA(cudaMemcpyAsync(&n_pairs, off.get() + n, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  ISFM_CUDA(cudaStreamSynchronize(s));
  sp.n_pairs = n_pairs;
  // 2. fill + sort by block key
  DeviceBuffer<uint64_t> keys, vals, keys_s;
  keys.alloc(n_pairs); vals.alloc(n_pairs); keys_s.alloc(n_pairs); sp.pairs.alloc(n_pairs);
  pair_fill_kernel<<<g, TPB, 0, s>>>(ix.pt_of.get(), ix.cam_of.get(), ix.pt_off.get(), n, off.get(), n_cam,
                                     keys.get(), vals.get());
  cnt.release(); off.release();
  int64_t n_lists = 0;
  DeviceBuffer<uint64_t> list_key;
  if (n_pairs > 0) {
    sort_pairs(keys.get(), keys_s.get(), vals.get(), sp.pairs.get(), n_pairs,
               bits_for((uint64_t)n_cam * (uint64_t)n_cam), s);
    keys.release(); vals.release();
    // 3. run-length encode -> lists
    DeviceBuffer<int64_t> run_len; DeviceBuffer<int64_t> n_runs;
    list_key.alloc(n_pairs); run_len.alloc(n_pairs + 1); n_runs.alloc(1);
    size_t bytes = 0;
    ISFM_CUDA(cub::DeviceRunLengthEncode::Encode(nullptr, bytes, keys_s.get(), list_key.get(), run_len.get(),
                                                 n_runs.get(), n_pairs, s));
    {
      DeviceBuffer<uint8_t> tmp; tmp.alloc(bytes);
      ISFM_CUDA(cub::DeviceRunLengthEncode::Encode(tmp.get(), bytes, keys_s.get(), list_key.get(), run_len.get(),
                                                   n_runs.get(), n_pairs, s));
      ISFM_CUDA(cudaMemcpyAsync(&n_lists, n_runs.get(), sizeof(int64_t), cudaMemcpyDeviceToHost, s));
      ISFM_CUDA(cudaStreamSynchronize(s));
    }
    keys_s.release();
    sp.list_off.alloc(n_lists + 1);
    ISFM_CUDA(cudaMemsetAsync(run_len.get() + n_lists, 0, sizeof(int64_t), s));
    ISFM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, run_len.get(), sp.list_off.get(), n_lists + 1, s));
    DeviceBuffer<uint8_t> tmp; tmp.alloc(bytes);
    ISFM_CUDA(cub::DeviceScan::ExclusiveSum(tmp.get(), bytes, run_len.get(), sp.list_off.get(), n_lists + 1, s));
    ISFM_CUDA(cudaStreamSynchronize(s));
  } else {
    sp.list_off.alloc(1);
    ISFM_CUDA(cudaMemsetAsync(sp.list_off.get(), 0, sizeof(int64_t), s));
  }
  sp.n_lists = n_lists;
  // 4. upper BSR pattern: strictly-upper lists + every diagonal block
  DeviceBuffer<uint64_t> extra;
  const int64_t n_extra = hook ? hook->extra_keys(list_key.get(), n_lists, n_cam, extra, s) : 0;
  const int64_t n_own = n_lists + n_cam, n_ent = n_own + n_extra;
  DeviceBuffer<uint64_t> ekeys, ekeys_s; DeviceBuffer<int64_t> etags, etags_s;
  ekeys.alloc(n_ent); ekeys_s.alloc(n_ent); etags.alloc(n_ent); etags_s.alloc(n_ent);
  upper_keys_kernel<<<div_up(n_own, TPB), TPB, 0, s>>>(list_key.get(), n_lists, n_cam, ekeys.get(), etags.get());
  if (n_extra > 0)
    extra_keys_kernel<<<div_up(n_extra, TPB), TPB, 0, s>>>(extra.get(), n_extra, n_cam, ekeys.get() + n_own, etags.get() + n_own);
  sort_pairs(ekeys.get(), ekeys_s.get(), etags.get(), etags_s.get(), n_ent, 64, s);
  if (n_extra > 0) {
    // drop the duplicates (stable sort: the local entry of a run comes first), then sort the
    // survivors to the front again
    dedupe_keys_kernel<<<div_up(n_ent, TPB), TPB, 0, s>>>(ekeys_s.get(), n_ent, ekeys.get());
    etags.swap(etags_s);
    sort_pairs(ekeys.get(), ekeys_s.get(), etags.get(), etags_s.get(), n_ent, 64, s);
    extra.release();
  }
  DeviceBuffer<unsigned long long> valid; valid.alloc(1); valid.zero(s);
  count_valid_kernel<<<div_up(n_ent, TPB), TPB, 0, s>>>(ekeys_s.get(), n_ent, valid.get());
  unsigned long long nnzu = 0;
  ISFM_CUDA(cudaMemcpyAsync(&nnzu, valid.get(), sizeof(nnzu), cudaMemcpyDeviceToHost, s));
  ISFM_CUDA(cudaStreamSynchronize(s));
  ISFM_REQUIRE(nnzu < (1ull << 31), ISFM_EINVAL, "reduced camera system has too many blocks for int32 slots");
  sp.nnzu = (int64_t)nnzu;
  sp.n_off = sp.nnzu - n_cam;
  sp.ucol.alloc(nnzu); sp.urow_ptr.alloc(n_cam + 1); sp.diag_slot.alloc(n_cam); sp.tpos.alloc(nnzu);
  sp.lrow_ptr.alloc(n_cam + 1);
  sp.list_slot.alloc(n_lists); sp.list_diag.alloc(n_lists);
  DeviceBuffer<int32_t> row_of; row_of.alloc(nnzu);
  upper_scatter_kernel<<<div_up(nnzu, TPB), TPB, 0, s>>>(ekeys_s.get(), etags_s.get(), (int64_t)nnzu, n_cam, sp.ucol.get(),
                                                         row_of.get(), sp.list_slot.get(), sp.diag_slot.get());
  lower_bound_kernel<<<div_up(n_cam + 1, TPB), TPB, 0, s>>>(row_of.get(), (int64_t)nnzu, sp.urow_ptr.get(), n_cam);
  if (n_lists > 0)
    diag_list_slot_kernel<<<div_up(n_lists, TPB), TPB, 0, s>>>(list_key.get(), n_lists, n_cam, sp.diag_slot.get(),
                                                               sp.list_slot.get(), sp.list_diag.get());
  // 5. lower ordering of the strictly-upper entries
  {
    DeviceBuffer<uint64_t> lkeys, lkeys_s; DeviceBuffer<int32_t> lvals, lvals_s, lrow_of;
    lkeys.alloc(nnzu); lkeys_s.alloc(nnzu); lvals.alloc(nnzu); lvals_s.alloc(nnzu); lrow_of.alloc(sp.n_off);
    lower_keys_kernel<<<div_up(nnzu, TPB), TPB, 0, s>>>(row_of.get(), sp.ucol.get(), (int64_t)nnzu, n_cam, lkeys.get(), lvals.get());
    sort_pairs(lkeys.get(), lkeys_s.get(), lvals.get(), lvals_s.get(), (int64_t)nnzu, 64, s);
    lower_scatter_kernel<<<div_up(nnzu, TPB), TPB, 0, s>>>(lkeys_s.get(), lvals_s.get(), (int64_t)nnzu, sp.n_off, n_cam,
                                                           sp.tpos.get(), lrow_of.get());
    lower_bound_kernel<<<div_up(n_cam + 1, TPB), TPB, 0, s>>>(lrow_of.get(), sp.n_off, sp.lrow_ptr.get(), n_cam);
    ISFM_CUDA(cudaGetLastError());
    ISFM_CUDA(cudaStreamSynchronize(s));
  }
  // 5b. list visiting order (stable radix sort)
  if (n_lists > 0) {
    ISFM_REQUIRE(n_lists < (1ll << 31), ISFM_EINVAL, "too many pair lists");
    DeviceBuffer<uint64_t> lkey, lkey_s; DeviceBuffer<int32_t> lid;
    lkey.alloc(n_lists); lkey_s.alloc(n_lists); lid.alloc(n_lists); sp.list_order.alloc(n_lists);
    list_len_key_kernel<<<div_up(n_lists, TPB), TPB, 0, s>>>(sp.list_off.get(), n_lists, lkey.get(), lid.get());
    sort_pairs(lkey.get(), lkey_s.get(), lid.get(), sp.list_order.get(), n_lists, 64, s);
  }
  // 6. pad every row to a multiple of 4 slots (padding slots: col = row, zero values, no deposit)
  //    and cut the rows into mat-vec chunks (host: n_cam + 1 integers)
  {
    std::vector<int32_t> up((size_t)n_cam + 1), upp((size_t)n_cam + 1), crow, cbeg, cptr((size_t)n_cam + 1);
    ISFM_CUDA(cudaMemcpyAsync(up.data(), sp.urow_ptr.get(), up.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    ISFM_CUDA(cudaStreamSynchronize(s));
    upp[0] = 0;
    for (int64_t i = 0; i < n_cam; ++i) upp[i + 1] = upp[i] + (up[i + 1] - up[i] + 3) / 4 * 4;
    const int64_t nnzp = upp[n_cam];
    DeviceBuffer<int32_t> d_up, d_upp, ucol_p, tpos_p, slot_of;
    d_up.alloc(n_cam + 1); d_upp.alloc(n_cam + 1); ucol_p.alloc(nnzp); tpos_p.alloc(nnzp); slot_of.alloc(nnzu);
    ISFM_CUDA(cudaMemcpyAsync(d_up.get(), up.data(), up.size() * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    ISFM_CUDA(cudaMemcpyAsync(d_upp.get(), upp.data(), upp.size() * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    pad_remap_kernel<<<div_up(nnzu, TPB), TPB, 0, s>>>((int64_t)nnzu, row_of.get(), d_up.get(), d_upp.get(), sp.ucol.get(),
                                                       sp.tpos.get(), ucol_p.get(), tpos_p.get(), slot_of.get());
    pad_fill_kernel<<<div_up(n_cam, TPB), TPB, 0, s>>>(n_cam, d_up.get(), d_upp.get(), ucol_p.get(), tpos_p.get());
    if (n_lists > 0) remap_slots_kernel<<<div_up(n_lists, TPB), TPB, 0, s>>>(n_lists, sp.list_slot.get(), slot_of.get());
    remap_slots_kernel<<<div_up(n_cam, TPB), TPB, 0, s>>>(n_cam, sp.diag_slot.get(), slot_of.get());
    ISFM_CUDA(cudaGetLastError());
    ISFM_CUDA(cudaStreamSynchronize(s));
    sp.ucol.swap(ucol_p);
    sp.tpos.swap(tpos_p);
    sp.urow_ptr.swap(d_upp);
    sp.n_blocks = sp.nnzu;
    sp.nnzu = nnzp;
    up = upp;
    int unit = SPMV_CHUNK;
    if (grid_warps > 0 && !getenv("ISFM_FIXED_UNITS")) {
      const int64_t slots_per_warp = nnzp / ((int64_t)grid_warps * std::max(hook ? hook->matvec_share : 1, 1));
      int min_units = 8;   // units per warp below which the unit is halved (tail balance of the static deal)
      if (const char* e = getenv("ISFM_UNITS_PER_WARP")) min_units = std::max(1, atoi(e));
      while (unit > 12 && slots_per_warp / unit < min_units) unit /= 2;
    }
    sp.unit_slots = unit;
    for (int64_t i = 0; i < n_cam; ++i) {
      cptr[i] = (int32_t)crow.size();
      for (int32_t b = up[i]; b < up[i + 1]; b += unit) { crow.push_back((int32_t)i); cbeg.push_back(b); }
    }
    cptr[n_cam] = (int32_t)crow.size();
    sp.n_chunks = (int64_t)crow.size();
    sp.h_chunk_beg = cbeg;
    sp.chunk_row.alloc(crow.size()); sp.chunk_beg.alloc(cbeg.size()); sp.chunk_ptr.alloc(cptr.size());
    ISFM_CUDA(cudaMemcpyAsync(sp.chunk_row.get(), crow.data(), crow.size() * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    ISFM_CUDA(cudaMemcpyAsync(sp.chunk_beg.get(), cbeg.data(), cbeg.size() * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    ISFM_CUDA(cudaMemcpyAsync(sp.chunk_ptr.get(), cptr.data(), cptr.size() * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    // stage table (persistent PCG kernel)
    sp.stage_blocks = stage_blocks; sp.n_stages = 0;
    std::vector<int4> stg;
    sp.h_unit_stage_ptr.assign(crow.size() + 1, 0);
    if (stage_blocks > 0) {
      stg.reserve(crow.size() * ((unit + stage_blocks - 1) / stage_blocks));
      for (size_t u = 0; u < crow.size(); ++u) {
        const int32_t row = crow[u], beg = cbeg[u], end = std::min<int32_t>(beg + unit, up[row + 1]);
        sp.h_unit_stage_ptr[u] = (int32_t)stg.size();
        for (int32_t b = beg; b < end; b += stage_blocks) {
          const int32_t nb = std::min<int32_t>(stage_blocks, end - b);
          stg.push_back(make_int4(row, b, nb | (b == beg ? 1 << 8 : 0) | (b + stage_blocks >= end ? 1 << 9 : 0), (int32_t)u));
        }
      }
      sp.h_unit_stage_ptr[crow.size()] = (int32_t)stg.size();
      sp.n_stages = (int64_t)stg.size();
      sp.stages.alloc(std::max<size_t>(stg.size(), 1)); sp.unit_stage_ptr.alloc(sp.h_unit_stage_ptr.size());
      ISFM_CUDA(cudaMemcpyAsync(sp.stages.get(), stg.data(), stg.size() * sizeof(int4), cudaMemcpyHostToDevice, s));
      ISFM_CUDA(cudaMemcpyAsync(sp.unit_stage_ptr.get(), sp.h_unit_stage_ptr.data(), sp.h_unit_stage_ptr.size() * sizeof(int32_t),
                                cudaMemcpyHostToDevice, s));
      if (sp.n_stages > 0 && !getenv("ISFM_PCG_NO_CONTIG"))
        stage_contiguity_kernel<<<div_up(sp.n_stages, TPB), TPB, 0, s>>>(sp.stages.get(), sp.n_stages, sp.ucol.get());
    }
    ISFM_CUDA(cudaStreamSynchronize(s));
  }
  ISFM_CUDA(cudaGetLastError());
  ISFM_CUDA(cudaStreamSynchronize(s));
}

namespace {
__global__ void pattern_hash_kernel(const int32_t* __restrict__ v, int64_t n, uint64_t salt, unsigned long long* out) {
  uint64_t a = 0, b = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t x = (uint64_t)(uint32_t)v[i] + 1u, k = (uint64_t)i + salt;
    a += x * (k * 0x9E3779B97F4A7C15ull + 0x7F4A7C15ull);
    b += (x ^ (x << 17)) * (k * 0xC2B2AE3D27D4EB4Full + 0x165667B1ull);
  }
  // integer atomics: the sum does not depend on the order
  atomicAdd(out, (unsigned long long)a);
  atomicAdd(out + 1, (unsigned long long)b);
}
}  // namespace

void schur_pattern_signature(const SchurPattern& sp, int64_t n_cam, cudaStream_t s, uint64_t sig_out[2]) {
  DeviceBuffer<unsigned long long> acc;
  acc.alloc(2); acc.zero(s);
  if (sp.nnzu > 0) pattern_hash_kernel<<<div_up(sp.nnzu, 256 * 8), 256, 0, s>>>(sp.ucol.get(), sp.nnzu, 1u, acc.get());
  pattern_hash_kernel<<<div_up(n_cam + 1, 256), 256, 0, s>>>(sp.urow_ptr.get(), n_cam + 1, 0x100000000ull, acc.get());
  unsigned long long h[2];
  ISFM_CUDA(cudaMemcpyAsync(h, acc.get(), sizeof h, cudaMemcpyDeviceToHost, s));
  ISFM_CUDA(cudaStreamSynchronize(s));
  sig_out[0] = h[0] >> 16; sig_out[1] = h[1] >> 16;
}

}  // namespace isfm
