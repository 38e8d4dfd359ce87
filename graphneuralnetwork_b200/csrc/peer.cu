// Multi-GPU halo exchange over NVLink peer memory.
// One process per GPU.  Halo buffers are plain cudaMalloc allocations exported with CUDA
// IPC so that every rank can map every peer's buffer.  gnn_halo_push_f32 is the fused
// pack + transfer: each CTA copies feature rows of the local X straight into the owning
// peer's halo buffer with 128-bit stores that travel over NVLink/NVSwitch — there is no
// staging buffer and no separate copy engine step.  The reference has no multi-GPU path
// for this (only nn.DataParallel, HAN/train_utils/train_eval.py:46); this is new.
#include "common.cuh"

using namespace gnn;

namespace {

constexpr int kMaxPeers = 16;

struct PushArgs {
  const float* X;
  int64_t ldx;
  int32_t F;
  const int32_t* send_rows;
  int64_t send_off[kMaxPeers + 1];
  float* halo[kMaxPeers];
  int64_t dst_off[kMaxPeers];
  int64_t ld_halo;
  int32_t n_peers;
  int64_t rot;  // first peer served (warp slots are dealt to peers rot, rot+1, ... mod n_peers)
};

// One warp per sent row, UNROLL rows in flight per warp (all index loads, then all row loads,
// then the stores).  The grid is deliberately small (one CTA per SM): NVLink needs ~1.5 MB in
// flight, and the remaining thread slots of every SM stay free for the local-column SpMM that
// runs concurrently on the main stream.
// All-to-all schedules (both avoid the ingress hot spot of "everybody pushes to peer 0 first",
// which cost 2.5x at 8 GPUs):
//   SCHED 0  rotated: rank r walks its segments in the order r+1, r+2, ... (one receiver at a time)
//   SCHED 1  interleaved: warp w serves peer first_peer + w % n_peers (all receivers at once)
//
// SM partition ("halo.dedicated_sms" = N > 0): instead of one small CTA on every SM — where the 8
// push warps compete with 24 SpMM warps for the same load/store unit and the overlapped exchange
// ran at ~420 GB/s instead of 610 — the push runs as N CTAs of 1024 threads that each claim
// "halo.exclusion_smem_kb" of shared memory.  One such CTA fills an SM's shared memory, and the
// concurrently launched local-column SpMM asks for a token amount of dynamic shared memory
// ("spmm.exclusion_smem_kb") that no longer fits next to it: the block scheduler itself keeps the
// two kernels on disjoint SMs (N for the NVLink stores, 148 - N for the HBM-bound SpMM).
template <int VEC, int UNROLL, int SCHED, int THREADS>
__global__ void __launch_bounds__(THREADS) halo_push_kernel(const PushArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t total = a.send_off[a.n_peers];
  int64_t k_begin, k_step, k_end, seg0 = 0;
  float* halo_q = nullptr;
  if (SCHED == 1) {
    const int slot = (int)(w % a.n_peers);
    const int q = (int)((a.rot + slot) % a.n_peers);
    const int64_t nwq = (nw - slot + a.n_peers - 1) / a.n_peers;  // warps dealt to this peer
    seg0 = a.send_off[q];
    k_begin = (w / a.n_peers) * UNROLL;
    k_step = nwq * UNROLL;
    k_end = a.send_off[q + 1] - seg0;
    halo_q = a.halo[q] + a.dst_off[q] * a.ld_halo;
  } else {
    k_begin = w * UNROLL;
    k_step = nw * UNROLL;
    k_end = total;
  }
  const int64_t rot_rows = a.send_off[a.rot];
  // position k of this warp's walk -> index into send_rows (SCHED 0 walks the segments rotated by rot)
  auto slot_of = [&](int64_t k) -> int64_t {
    if (SCHED == 1) return seg0 + k;
    k += rot_rows;
    return k >= total ? k - total : k;
  };
  // The sent-row ids of the NEXT iteration are loaded before the rows of this one are stored: without
  // this prefetch every iteration was a chain of two dependent DRAM round trips (id, then row), which
  // held a dedicated push SM to ~15 GB/s.
  int32_t rid[UNROLL], rid_next[UNROLL];
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    const int64_t k = k_begin + u;
    rid[u] = (k < k_end) ? __ldg(a.send_rows + slot_of(k)) : -1;
  }
  for (int64_t k0 = k_begin; k0 < k_end; k0 += k_step) {
    const float* src[UNROLL];
    float* dst[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int64_t kn = k0 + k_step + u;
      rid_next[u] = (kn < k_end) ? __ldg(a.send_rows + slot_of(kn)) : -1;
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int64_t k = k0 + u;
      src[u] = nullptr;
      dst[u] = nullptr;
      if (k < k_end) {
        src[u] = a.X + (int64_t)rid[u] * a.ldx;
        if (SCHED == 1) {
          dst[u] = halo_q + k * a.ld_halo;
        } else {
          const int64_t ks = slot_of(k);
          int q = 0;
          while (q + 1 < a.n_peers && ks >= a.send_off[q + 1]) ++q;
          dst[u] = a.halo[q] + (a.dst_off[q] + (ks - a.send_off[q])) * a.ld_halo;
        }
      }
    }
    for (int c = lane * VEC; c < a.F; c += 32 * VEC) {
      float v[UNROLL][VEC];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
        if (src[u]) VecIO<float, VEC>::load(src[u] + c, v[u]);
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
        if (dst[u]) VecIO<float, VEC>::store(dst[u] + c, v[u]);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) rid[u] = rid_next[u];
  }
}

// ---- TMA variant of the push (experimental, "halo.tma" = 1; not the default) -------------------
// The vector-store push sustains only ~16-20 GB/s of remote stores per SM (r01 8-GPU sweep), so NVLink
// needs >= 32 SMs.  Here no feature row passes through registers or the load/store unit: every warp is
// an independent mover with a ring of kTmaStages shared-memory stages of 32 rows.  32 lanes issue one
// bulk gather (cp.async.bulk global -> shared) each; when the stage's mbarrier completes, ONE lane issues
// ONE bulk store (cp.async.bulk shared -> peer global) for the whole stage: the rows a peer receives are
// consecutive in its halo buffer, so a stage leaves as a single 16 KB NVLink write.  The host deals whole
// stages ("chunks") that never straddle a peer segment, in the rotated order of SCHED 0.
constexpr int kTmaStages = 3;
constexpr int kTmaWarps = 4;

struct PushTmaArgs {
  const float* X;
  int64_t ldx;
  int32_t row_bytes;            // F * 4, multiple of 16, == ld_halo * 4 (received rows are contiguous)
  const int32_t* send_rows;
  int64_t seg_begin[kMaxPeers];   // rotated slot s: first entry of send_rows
  int64_t seg_rows[kMaxPeers];    // rotated slot s: rows to send
  int64_t chunk_off[kMaxPeers + 1];  // rotated slot s: first chunk id (32 rows per chunk)
  float* dst[kMaxPeers];          // rotated slot s: peer halo base + dst_off rows
  int32_t n_slots;
};

__global__ void __launch_bounds__(kTmaWarps * 32) halo_push_tma_kernel(const PushTmaArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int stage_bytes = 32 * a.row_bytes;
  unsigned char* ring = smem + (size_t)warp * kTmaStages * stage_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)kTmaWarps * kTmaStages * stage_bytes) + warp * kTmaStages;
  if (lane == 0) {
    for (int s = 0; s < kTmaStages; ++s) mbar_init(full + s, 1);
    mbar_fence_init();
  }
  __syncwarp();
  const int64_t n_chunks = a.chunk_off[a.n_slots];
  const int64_t w = (int64_t)blockIdx.x * kTmaWarps + warp;
  const int64_t nw = (int64_t)gridDim.x * kTmaWarps;
  // chunk id -> (slot, first row of the chunk within the slot's segment)
  auto locate = [&](int64_t c, int& slot, int64_t& row0) {
    slot = 0;
    while (slot + 1 < a.n_slots && c >= a.chunk_off[slot + 1]) ++slot;
    row0 = (c - a.chunk_off[slot]) * 32;
  };
  auto load_id = [&](int64_t c) -> int32_t {
    if (c >= n_chunks) return -1;
    int slot;
    int64_t row0;
    locate(c, slot, row0);
    const int64_t r = row0 + lane;
    return r < a.seg_rows[slot] ? __ldg(a.send_rows + a.seg_begin[slot] + r) : -1;
  };
  // gathers of chunk c into ring stage `stg` (ids already in `rid_c`)
  auto issue = [&](int64_t c, int32_t rid_c, int stg) {
    int slot;
    int64_t row0;
    locate(c, slot, row0);
    const int64_t left = a.seg_rows[slot] - row0;
    const int rows = (int)(left < 32 ? left : 32);
    if (lane == 0) mbar_arrive_expect_tx(full + stg, (uint32_t)rows * (uint32_t)a.row_bytes);
    __syncwarp();
    if (lane < rows)
      bulk_g2s(ring + (size_t)stg * stage_bytes + (size_t)lane * a.row_bytes, a.X + (int64_t)rid_c * a.ldx,
               (uint32_t)a.row_bytes, full + stg);
  };
  // Software pipeline: the gathers of the next kTmaStages-1 chunks are in flight while chunk k is stored.
  constexpr int P = kTmaStages - 1;
  int32_t rid_ahead = load_id(w);
  int fill_stage = 0;
#pragma unroll
  for (int k = 0; k < P; ++k) {
    const int64_t c = w + (int64_t)k * nw;
    if (c < n_chunks) issue(c, rid_ahead, fill_stage);
    rid_ahead = load_id(c + nw);
    fill_stage = (fill_stage + 1 == kTmaStages) ? 0 : fill_stage + 1;
  }
  // here: rid_ahead = ids of chunk w + P*nw, fill_stage = P % kTmaStages
  int stage = 0;
  uint32_t parity = 0;
  for (int64_t c = w; c < n_chunks; c += nw) {
    int slot;
    int64_t row0;
    locate(c, slot, row0);
    const int64_t left = a.seg_rows[slot] - row0;
    const int rows = (int)(left < 32 ? left : 32);
    mbar_wait(full + stage, parity);
    if (lane == 0) {
      fence_proxy_async_smem();
      bulk_s2g(reinterpret_cast<unsigned char*>(a.dst[slot]) + row0 * a.row_bytes, ring + (size_t)stage * stage_bytes,
               (uint32_t)rows * (uint32_t)a.row_bytes);
      bulk_commit_group();
    }
    // refill the stage chunk c - nw was stored from: that store must have read it (at most the store just
    // committed may still be reading)
    const int64_t cf = c + (int64_t)P * nw;
    if (cf < n_chunks) {
      if (lane == 0) bulk_wait_group_read<1>();
      __syncwarp();
      issue(cf, rid_ahead, fill_stage);
    }
    rid_ahead = load_id(cf + nw);
    fill_stage = (fill_stage + 1 == kTmaStages) ? 0 : fill_stage + 1;
    if (++stage == kTmaStages) {
      stage = 0;
      parity ^= 1u;
    }
  }
  if (lane == 0) bulk_wait_group<0>();  // every store performed before the kernel (and the barrier after it) ends
  __syncwarp();
}

}  // namespace

extern "C" {

int gnn_peer_alloc(size_t bytes, void** dev_ptr, void* ipc_handle_64B_host) {
  GNN_REQUIRE(dev_ptr && ipc_handle_64B_host && bytes > 0, GNN_ERR_BAD_ARG, "bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
  GNN_CUDA(cudaMalloc(dev_ptr, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, *dev_ptr);
  if (e != cudaSuccess) {
    cudaFree(*dev_ptr);
    *dev_ptr = nullptr;
    set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    return GNN_ERR_CUDA;
  }
  memcpy(ipc_handle_64B_host, &h, 64);
  return GNN_OK;
}

int gnn_peer_open(const void* ipc_handle_64B_host, void** dev_ptr) {
  GNN_REQUIRE(dev_ptr && ipc_handle_64B_host, GNN_ERR_BAD_ARG, "bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle_64B_host, 64);
  GNN_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return GNN_OK;
}

int gnn_peer_close(void* dev_ptr) {
  if (!dev_ptr) return GNN_OK;
  GNN_CUDA(cudaIpcCloseMemHandle(dev_ptr));
  return GNN_OK;
}

int gnn_peer_free(void* dev_ptr) {
  if (!dev_ptr) return GNN_OK;
  GNN_CUDA(cudaFree(dev_ptr));
  return GNN_OK;
}

int gnn_peer_copy_async(void* dst, const void* src, size_t bytes, gnn_stream_t stream) {
  if (bytes == 0) return GNN_OK;
  GNN_REQUIRE(dst && src, GNN_ERR_BAD_ARG, "null pointer");
  GNN_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return GNN_OK;
}

int gnn_halo_push_f32(const float* X, int64_t ldx, int32_t F, const int32_t* send_rows, const int64_t* send_off_host,
                      float* const* peer_halo_host, const int64_t* dst_off_host, int64_t ld_halo, int32_t n_peers,
                      int32_t first_peer, gnn_stream_t stream) {
  GNN_REQUIRE(n_peers >= 0 && n_peers <= kMaxPeers, GNN_ERR_UNSUPPORTED, "n_peers=%d exceeds %d", n_peers, kMaxPeers);
  if (n_peers == 0) return GNN_OK;
  GNN_REQUIRE(X && send_off_host && peer_halo_host && dst_off_host, GNN_ERR_BAD_ARG, "null pointer");
  GNN_REQUIRE(F > 0 && ldx >= F && ld_halo >= F, GNN_ERR_BAD_ARG, "bad feature width / leading dimension");
  PushArgs a{};
  a.X = X;
  a.ldx = ldx;
  a.F = F;
  a.send_rows = send_rows;
  a.ld_halo = ld_halo;
  a.n_peers = n_peers;
  bool vec4 = aligned_to(X, 16) && ldx % 4 == 0 && ld_halo % 4 == 0 && F % 4 == 0;
  for (int q = 0; q <= n_peers; ++q) a.send_off[q] = send_off_host[q];
  for (int q = 0; q < n_peers; ++q) {
    a.halo[q] = peer_halo_host[q];
    a.dst_off[q] = dst_off_host[q];
    GNN_REQUIRE(a.send_off[q + 1] >= a.send_off[q], GNN_ERR_BAD_ARG, "send_off not monotone");
    GNN_REQUIRE(a.send_off[q + 1] == a.send_off[q] || a.halo[q] != nullptr, GNN_ERR_BAD_ARG, "null peer halo %d", q);
    vec4 = vec4 && aligned_to(a.halo[q], 16);
  }
  const int64_t total = a.send_off[n_peers];
  if (total == 0) return GNN_OK;
  GNN_REQUIRE(first_peer >= 0 && first_peer < n_peers, GNN_ERR_BAD_ARG, "first_peer out of range");
  a.rot = first_peer;
  GNN_REQUIRE(send_rows != nullptr, GNN_ERR_BAD_ARG, "null send_rows");
  // the grid must deal at least one warp to every peer: never fewer than n_peers warps
  int64_t grid = (total * 32 + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * tuning("halo.ctas_per_sm", 1);
  grid = grid > cap ? cap : grid;
  const int64_t min_grid = (n_peers + 7) / 8;
  grid = grid < min_grid ? min_grid : grid;
  const int sched = tuning("halo.schedule", 0);
  const int unroll = tuning("halo.unroll", 4);
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned g = (unsigned)grid;
  const int dedicated = tuning("halo.dedicated_sms", 0);
  if (tuning("halo.tma", 0) && vec4 && ld_halo == F && ((int64_t)F * 4) % 16 == 0 && (int64_t)F * 4 * 32 * kTmaStages <= 56 * 1024) {
    // experimental TMA mover (see halo_push_tma_kernel); needs contiguous received rows (ld_halo == F)
    PushTmaArgs t{};
    t.X = X;
    t.ldx = ldx;
    t.row_bytes = F * 4;
    t.send_rows = send_rows;
    int ns = 0;
    t.chunk_off[0] = 0;
    for (int s = 0; s < n_peers; ++s) {
      const int q = (first_peer + s) % n_peers;
      const int64_t rows = a.send_off[q + 1] - a.send_off[q];
      if (rows == 0) continue;
      t.seg_begin[ns] = a.send_off[q];
      t.seg_rows[ns] = rows;
      t.dst[ns] = a.halo[q] + a.dst_off[q] * ld_halo;
      t.chunk_off[ns + 1] = t.chunk_off[ns] + (rows + 31) / 32;
      ++ns;
    }
    t.n_slots = ns;
    const size_t smem = (size_t)kTmaWarps * kTmaStages * 32 * t.row_bytes + (size_t)kTmaWarps * kTmaStages * 8;
    static size_t configured = 0;
    if (smem > configured) {
      GNN_CUDA(cudaFuncSetAttribute(halo_push_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured = smem;
    }
    int64_t gt = dedicated > 0 ? dedicated : num_sms();
    const int64_t need = (t.chunk_off[ns] + kTmaWarps - 1) / kTmaWarps;
    gt = gt > need ? need : gt;
    halo_push_tma_kernel<<<(unsigned)(gt < 1 ? 1 : gt), kTmaWarps * 32, smem, st>>>(t);
  } else if (dedicated > 0 && vec4) {
    const size_t smem = (size_t)tuning("halo.exclusion_smem_kb", 200) * 1024;
    static size_t configured = 0;
    if (smem > configured) {
      GNN_CUDA(cudaFuncSetAttribute(halo_push_kernel<4, 4, 0, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      GNN_CUDA(cudaFuncSetAttribute(halo_push_kernel<4, 8, 0, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured = smem;
    }
    const unsigned gd = (unsigned)(dedicated < num_sms() ? dedicated : num_sms());
    // 64 KB of rows in flight per SM either way: 32 warps x 4 rows, or 16 warps x 8 rows (8 rows per
    // warp need more than the 64 registers a 1024-thread CTA leaves each thread)
    if (unroll >= 8) halo_push_kernel<4, 8, 0, 512><<<gd, 512, smem, st>>>(a);
    else halo_push_kernel<4, 4, 0, 1024><<<gd, 1024, smem, st>>>(a);
  } else if (vec4) {
    if (sched == 1 && unroll >= 8) halo_push_kernel<4, 8, 1, 256><<<g, 256, 0, st>>>(a);
    else if (sched == 1) halo_push_kernel<4, 4, 1, 256><<<g, 256, 0, st>>>(a);
    else if (unroll >= 8) halo_push_kernel<4, 8, 0, 256><<<g, 256, 0, st>>>(a);
    else halo_push_kernel<4, 4, 0, 256><<<g, 256, 0, st>>>(a);
  } else {
    if (sched == 1) halo_push_kernel<1, 4, 1, 256><<<g, 256, 0, st>>>(a);
    else halo_push_kernel<1, 4, 0, 256><<<g, 256, 0, st>>>(a);
  }
  GNN_LAUNCH_CHECK();
  return GNN_OK;
}

}  // extern "C"
