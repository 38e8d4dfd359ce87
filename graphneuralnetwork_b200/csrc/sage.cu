// GraphSAGE fixed-fanout neighbour gather + reduce, fused.
// Replaces, at /root/reference:
//   GraphSAGE_Pytorch/data_utils.py:64          CPU feature gather of the sampled ids
//   GraphSAGE_Pytorch/models/Aggregator.py:19-24 mean / sum / max over the fanout axis
//   GraphSAGE/graph_utils.py:6 + GraphSAGE.py:47-49  torch.embedding + torch.mean
//
// TMA path (sm_100a): a persistent CTA owns source nodes blockIdx.x, +grid, ...; one
// producer warp reads the sampled ids and issues one cp.async.bulk (TMA 1-D bulk copy)
// per gathered feature row into a shared-memory ring, completing on an mbarrier; five
// consumer warps add the staged rows from shared memory (LDS.128) in fanout order and
// write the reduced row once.  No feature row touches a register before it is reduced,
// and ~200 KB of row fetches are in flight per SM.
// Vector-load path: the row-parallel kernel of rowreduce.cuh with an implicit CSR
// (rowptr = i*fanout), used when rows are not 16-byte aligned or too short for TMA.
#include "rowreduce.cuh"

using namespace gnn;

namespace {

constexpr int kConsumerWarps = 5;
constexpr int kConsumerThreads = kConsumerWarps * 32;
constexpr int kSageThreads = 32 + kConsumerThreads;

template <typename T>
struct SageArgs {
  const T* table;
  int64_t ld;
  const int32_t* idx32;
  const int64_t* idx64;
  int64_t n_src;
  int32_t fanout;
  int32_t F;
  float scale;
  T* out;
  int64_t ldo;
  int32_t* argmax;
  int32_t row_bytes;  // bytes copied per row (multiple of 16)
  int32_t kc;         // rows per ring stage
  int32_t stages;
  int32_t nvec;       // 16-byte vectors per row
  int32_t out_vec16;  // output rows can take 16-byte stores
};

template <typename T>
struct Vec16;
template <>
struct Vec16<float> {
  static constexpr int E = 4;
  static __device__ __forceinline__ void unpack(const uint4& w, float (&o)[4]) {
    o[0] = __uint_as_float(w.x); o[1] = __uint_as_float(w.y); o[2] = __uint_as_float(w.z); o[3] = __uint_as_float(w.w);
  }
};
template <>
struct Vec16<__nv_bfloat16> {
  static constexpr int E = 8;
  static __device__ __forceinline__ void unpack(const uint4& w, float (&o)[8]) {
    bf16x2_to_f32(w.x, o[0], o[1]); bf16x2_to_f32(w.y, o[2], o[3]);
    bf16x2_to_f32(w.z, o[4], o[5]); bf16x2_to_f32(w.w, o[6], o[7]);
  }
};

template <typename T, int NV, int OP>
__global__ void __launch_bounds__(kSageThreads) sage_tma_kernel(const SageArgs<T> a) {
  constexpr int E = Vec16<T>::E;
  extern __shared__ __align__(128) unsigned char smem[];
  const int S = a.stages;
  const int stage_bytes = a.kc * a.row_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)S * stage_bytes);
  uint64_t* empty = full + S;
  uint32_t* mask = reinterpret_cast<uint32_t*>(empty + S);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(full + i, 1);
      mbar_init(empty + i, kConsumerWarps);
    }
    mbar_fence_init();
  }
  __syncthreads();

  const int nchunk = (a.fanout + a.kc - 1) / a.kc;
  const int64_t my_src = (a.n_src > (int64_t)blockIdx.x) ? (a.n_src - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int64_t total = my_src * nchunk;

  if (warp == 0) {
    // ===== producer: sampled ids -> one bulk copy per gathered row =====
    auto load_id = [&](int64_t it) -> int64_t {
      if (it >= total) return -1;
      const int64_t src = blockIdx.x + (it / nchunk) * (int64_t)gridDim.x;
      const int c0 = (int)(it % nchunk) * a.kc;
      const int rows = min(a.kc, a.fanout - c0);
      if (lane >= rows) return -1;
      const int64_t p = src * a.fanout + c0 + lane;
      if (a.idx32) return (int64_t)__ldg(a.idx32 + p);
      if (a.idx64) return __ldg(a.idx64 + p);
      return p;
    };
    int64_t r_next = load_id(0);
    for (int64_t it = 0; it < total; ++it) {
      const int64_t r = r_next;
      r_next = load_id(it + 1);  // prefetch the next chunk's ids while this one is issued
      const int stage = (int)(it % S);
      const uint32_t par = (uint32_t)((it / S) & 1);
      mbar_wait(empty + stage, par ^ 1u);
      const bool valid = r >= 0;
      const unsigned m = __ballot_sync(0xffffffffu, valid);
      if (lane == 0) {
        mask[stage] = m;
        mbar_arrive_expect_tx(full + stage, (uint32_t)__popc(m) * (uint32_t)a.row_bytes);
      }
      __syncwarp();
      if (valid)
        bulk_g2s(smem + (size_t)stage * stage_bytes + (size_t)lane * a.row_bytes, a.table + r * a.ld,
                 (uint32_t)a.row_bytes, full + stage);
    }
  } else {
    // ===== consumers: reduce the staged rows in fanout order =====
    const int t = threadIdx.x - 32;
    float acc[NV][E];
    int best[NV][E];
    auto reset = [&]() {
#pragma unroll
      for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int i = 0; i < E; ++i) {
          acc[v][i] = (OP == GNN_REDUCE_MAX) ? -INFINITY : 0.f;
          best[v][i] = 0;
        }
    };
    reset();
    for (int64_t it = 0; it < total; ++it) {
      const int stage = (int)(it % S);
      const uint32_t par = (uint32_t)((it / S) & 1);
      const int chunk = (int)(it % nchunk);
      const int c0 = chunk * a.kc;
      const int rows = min(a.kc, a.fanout - c0);
      mbar_wait(full + stage, par);
      const unsigned m = mask[stage];
      const unsigned char* sb = smem + (size_t)stage * stage_bytes;
#pragma unroll 4
      for (int k = 0; k < rows; ++k) {
        if (!((m >> k) & 1u)) continue;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const int vi = t + v * kConsumerThreads;
          if (vi < a.nvec) {
            const uint4 w = *reinterpret_cast<const uint4*>(sb + (size_t)k * a.row_bytes + (size_t)vi * 16);
            float x[E];
            Vec16<T>::unpack(w, x);
#pragma unroll
            for (int i = 0; i < E; ++i) {
              if (OP == GNN_REDUCE_MAX) {
                if (x[i] > acc[v][i]) {
                  acc[v][i] = x[i];
                  best[v][i] = c0 + k;
                }
              } else {
                acc[v][i] += x[i];
              }
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + stage);
      if (chunk == nchunk - 1) {
        const int64_t src = blockIdx.x + (it / nchunk) * (int64_t)gridDim.x;
        T* orow = a.out + src * a.ldo;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const int vi = t + v * kConsumerThreads;
          if (vi < a.nvec) {
            float o[E];
#pragma unroll
            for (int i = 0; i < E; ++i) o[i] = (OP == GNN_REDUCE_MAX) ? acc[v][i] : acc[v][i] * a.scale;
            const int col0 = vi * E;
            if (a.out_vec16) {
              VecIO<T, E>::store(orow + col0, o);
            } else {
#pragma unroll
              for (int i = 0; i < E; ++i)
                if (col0 + i < a.F) {
                  float o1[1] = {o[i]};
                  VecIO<T, 1>::store(orow + col0 + i, o1);
                }
            }
            if (OP == GNN_REDUCE_MAX && a.argmax) {
#pragma unroll
              for (int i = 0; i < E; ++i)
                if (col0 + i < a.F) a.argmax[src * a.ldo + col0 + i] = best[v][i];
            }
          }
        }
        reset();
      }
    }
  }
}

template <typename T, int NV, int OP>
int launch_tma_inst(const SageArgs<T>& a, size_t smem_bytes, int grid, cudaStream_t st) {
  static size_t configured = 0;  // benign race: attribute set is idempotent
  if (smem_bytes > configured) {
    GNN_CUDA(cudaFuncSetAttribute(sage_tma_kernel<T, NV, OP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem_bytes));
    configured = smem_bytes;
  }
  sage_tma_kernel<T, NV, OP><<<grid, kSageThreads, smem_bytes, st>>>(a);
  GNN_LAUNCH_CHECK();
  return GNN_OK;
}

template <typename T, int OP>
int launch_tma(const SageArgs<T>& a, size_t smem_bytes, int grid, cudaStream_t st) {
  const int nv = (a.nvec + kConsumerThreads - 1) / kConsumerThreads;
  switch (nv) {
    case 1: return launch_tma_inst<T, 1, OP>(a, smem_bytes, grid, st);
    case 2: return launch_tma_inst<T, 2, OP>(a, smem_bytes, grid, st);
    case 3: return launch_tma_inst<T, 3, OP>(a, smem_bytes, grid, st);
    case 4: return launch_tma_inst<T, 4, OP>(a, smem_bytes, grid, st);
    default: break;
  }
  set_error("TMA gather: row too wide (nvec=%d)", a.nvec);
  return GNN_ERR_UNSUPPORTED;
}

template <typename T>
int gather_reduce_impl(const T* table, int64_t ld, int64_t n_table_rows, const void* idx, int idx_bits, int64_t n_src,
                       int32_t fanout, int32_t F, int reduce, T* out, int64_t ldo, int32_t* argmax, cudaStream_t st) {
  GNN_REQUIRE(n_src >= 0 && fanout >= 0 && F >= 0, GNN_ERR_BAD_ARG, "negative size");
  GNN_REQUIRE(reduce == GNN_REDUCE_MEAN || reduce == GNN_REDUCE_SUM || reduce == GNN_REDUCE_MAX, GNN_ERR_BAD_ARG,
              "unknown reduce %d (GraphSAGE_Pytorch/models/Aggregator.py:26 raises ValueError)", reduce);
  if (n_src == 0 || F == 0) return GNN_OK;
  GNN_REQUIRE(fanout > 0, GNN_ERR_BAD_ARG, "fanout must be positive");
  GNN_REQUIRE(table && out, GNN_ERR_BAD_ARG, "null table/out");
  GNN_REQUIRE(idx == nullptr || idx_bits == 32 || idx_bits == 64, GNN_ERR_BAD_ARG, "idx_bits must be 32 or 64");
  GNN_REQUIRE(ld >= F && ldo >= F, GNN_ERR_BAD_ARG, "leading dimension smaller than F");
  GNN_REQUIRE(n_table_rows < 0x7fffffffLL, GNN_ERR_UNSUPPORTED, "table rows do not fit int32");
  GNN_REQUIRE(idx != nullptr || n_src * (int64_t)fanout <= n_table_rows, GNN_ERR_BAD_ARG,
              "identity block needs n_src*fanout <= n_table_rows");
  const float scale = (reduce == GNN_REDUCE_MEAN) ? 1.0f / (float)fanout : 1.0f;

  const int esz = (int)sizeof(T);
  const int row_bytes = (int)round_up((size_t)F * esz, 16);
  const bool tma_ok = !tuning("sage.force_ldg", 0) && aligned_to(table, 16) && ((ld * esz) % 16 == 0) &&
                      (int64_t)row_bytes <= ld * esz && row_bytes >= 256 && (row_bytes / 16) <= 4 * kConsumerThreads;
  if (tma_ok) {
    SageArgs<T> a{};
    a.table = table;
    a.ld = ld;
    a.idx32 = (idx && idx_bits == 32) ? (const int32_t*)idx : nullptr;
    a.idx64 = (idx && idx_bits == 64) ? (const int64_t*)idx : nullptr;
    a.n_src = n_src;
    a.fanout = fanout;
    a.F = F;
    a.scale = scale;
    a.out = out;
    a.ldo = ldo;
    a.argmax = (reduce == GNN_REDUCE_MAX) ? argmax : nullptr;
    a.row_bytes = row_bytes;
    a.nvec = row_bytes / 16;
    const int E = 16 / esz;
    a.out_vec16 = aligned_to(out, 16) && ((ldo * esz) % 16 == 0) && ((int64_t)a.nvec * E <= ldo);
    int kc = tuning("sage.chunk_bytes", 12288) / row_bytes;
    kc = kc < 1 ? 1 : kc;
    kc = kc > 32 ? 32 : kc;
    kc = kc > fanout ? fanout : kc;
    const int nchunk = (fanout + kc - 1) / kc;
    kc = (fanout + nchunk - 1) / nchunk;  // balance the chunks of one source
    const size_t budget = (size_t)tuning("sage.smem_kb", 48) * 1024;
    int stages = (int)(budget / ((size_t)kc * row_bytes));
    while (stages < 2 && kc > 1) {
      kc = (kc + 1) / 2;
      stages = (int)(budget / ((size_t)kc * row_bytes));
    }
    stages = stages > 8 ? 8 : stages;
    GNN_REQUIRE(stages >= 2, GNN_ERR_UNSUPPORTED, "row of %d bytes does not fit the shared-memory ring", row_bytes);
    a.kc = kc;
    a.stages = stages;
    const size_t smem_bytes = (size_t)stages * kc * row_bytes + (size_t)stages * (16 + 4) + 16;
    int64_t grid = (int64_t)num_sms() * tuning("sage.ctas_per_sm", 4);
    grid = grid > n_src ? n_src : grid;
    switch (reduce) {
      case GNN_REDUCE_MAX: return launch_tma<T, GNN_REDUCE_MAX>(a, smem_bytes, (int)grid, st);
      default: return launch_tma<T, GNN_REDUCE_SUM>(a, smem_bytes, (int)grid, st);
    }
  }

  RowArgs<T> r{};
  r.rowptr = nullptr;
  r.fanout = fanout;
  r.col32 = (idx && idx_bits == 32) ? (const int32_t*)idx : nullptr;
  r.col64 = (idx && idx_bits == 64) ? (const int64_t*)idx : nullptr;
  r.val = nullptr;
  r.src_div = 0;
  r.scale = scale;
  r.X = table;
  r.ldx = ld;
  r.Y = out;
  r.ldy = ldo;
  r.n_rows = n_src;
  r.F = F;
  r.skip_deg_gt = 0;
  r.argmax = (reduce == GNN_REDUCE_MAX) ? argmax : nullptr;
  if (reduce == GNN_REDUCE_MAX) return launch_row_reduce<T, 1>(r, st);
  return launch_row_reduce<T, 0>(r, st);
}

__global__ void __launch_bounds__(256) bwd_dense_kernel(const float* __restrict__ d_out, int64_t ld,
                                                        const int32_t* __restrict__ argmax, int64_t n_src, int fanout,
                                                        int F, float scale, float* __restrict__ d_neigh) {
  const int64_t total = n_src * (int64_t)fanout * F;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
    const int f = (int)(p % F);
    const int64_t ik = p / F;
    const int k = (int)(ik % fanout);
    const int64_t i = ik / fanout;
    const float g = __ldg(d_out + i * ld + f);
    float v;
    if (argmax) v = (__ldg(argmax + i * ld + f) == k) ? g : 0.f;
    else v = g * scale;
    d_neigh[p] = v;
  }
}

}  // namespace

extern "C" {

int gnn_gather_reduce_f32(const float* table, int64_t ld_table, int64_t n_table_rows, const void* idx, int idx_bits,
                          int64_t n_src, int32_t fanout, int32_t F, int reduce, float* out, int64_t ld_out,
                          int32_t* argmax, gnn_stream_t stream) {
  return gather_reduce_impl<float>(table, ld_table, n_table_rows, idx, idx_bits, n_src, fanout, F, reduce, out, ld_out,
                                   argmax, (cudaStream_t)stream);
}

int gnn_gather_reduce_bf16(const void* table, int64_t ld_table, int64_t n_table_rows, const void* idx, int idx_bits,
                           int64_t n_src, int32_t fanout, int32_t F, int reduce, void* out, int64_t ld_out,
                           int32_t* argmax, gnn_stream_t stream) {
  return gather_reduce_impl<__nv_bfloat16>((const __nv_bfloat16*)table, ld_table, n_table_rows, idx, idx_bits, n_src,
                                           fanout, F, reduce, (__nv_bfloat16*)out, ld_out, argmax,
                                           (cudaStream_t)stream);
}

int gnn_gather_reduce_bwd_f32(const int64_t* rowptr_t, const int32_t* pos_t, int64_t n_table_rows, int32_t fanout,
                              float scale, const float* d_out, int64_t ld_dout, float* d_table, int64_t ld_dtable,
                              int32_t F, gnn_stream_t stream) {
  GNN_REQUIRE(n_table_rows >= 0 && F >= 0 && fanout > 0, GNN_ERR_BAD_ARG, "bad size");
  if (n_table_rows == 0 || F == 0) return GNN_OK;
  GNN_REQUIRE(rowptr_t && pos_t && d_out && d_table, GNN_ERR_BAD_ARG, "null pointer");
  GNN_REQUIRE(ld_dout >= F && ld_dtable >= F, GNN_ERR_BAD_ARG, "leading dimension smaller than F");
  RowArgs<float> r{};
  r.rowptr = rowptr_t;
  r.fanout = 0;
  r.col32 = pos_t;
  r.col64 = nullptr;
  r.val = nullptr;
  r.src_div = fanout;
  r.scale = scale;
  r.X = d_out;
  r.ldx = ld_dout;
  r.Y = d_table;
  r.ldy = ld_dtable;
  r.n_rows = n_table_rows;
  r.F = F;
  r.skip_deg_gt = 0;
  r.argmax = nullptr;
  return launch_row_reduce<float, 0>(r, (cudaStream_t)stream);
}

int gnn_gather_reduce_bwd_dense_f32(const float* d_out, int64_t ld_dout, const int32_t* argmax, int64_t n_src,
                                    int32_t fanout, int32_t F, float scale, float* d_neigh, gnn_stream_t stream) {
  GNN_REQUIRE(n_src >= 0 && fanout > 0 && F >= 0, GNN_ERR_BAD_ARG, "bad size");
  if (n_src == 0 || F == 0) return GNN_OK;
  GNN_REQUIRE(d_out && d_neigh, GNN_ERR_BAD_ARG, "null pointer");
  const int64_t total = n_src * (int64_t)fanout * F;
  int64_t grid = (total + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 16;
  grid = grid > cap ? cap : grid;
  bwd_dense_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(d_out, ld_dout, argmax, n_src, fanout, F, scale,
                                                                   d_neigh);
  GNN_LAUNCH_CHECK();
  return GNN_OK;
}

}  // extern "C"
