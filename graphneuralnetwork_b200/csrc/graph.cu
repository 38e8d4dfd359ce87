// Graph construction on the device (bit-exact integer work) and the synthetic
// benchmark-shape generators.
// Reference behaviour restated here (/root/reference):
//   GCN/data_utils.py:63-70   scipy COO (row-major sorted after .astype) -> torch COO
//   GAT/models/layers.py:29,98  `adj > 0` mask / adj.nonzero() row-major edge order
//   HAN/utils/data_utils.py:85-89  float64 0/1 metapath adjacency
//   GraphSAGE_Pytorch/sample_utils.py:16  src-major fixed-fanout blocks
#include "common.cuh"
#include <cub/cub.cuh>

using namespace gnn;

namespace {

struct WsCursor {
  char* base;
  size_t off;
  size_t cap;
  template <typename T>
  T* take(size_t n) {
    off = round_up(off, 256);
    T* p = reinterpret_cast<T*>(base + off);
    off += n * sizeof(T);
    return p;
  }
};

inline int grid_for(int64_t n, int block = 256) {
  int64_t g = (n + block - 1) / block;
  const int64_t cap = (int64_t)num_sms() * 32;
  g = g < 1 ? 1 : g;
  return (int)(g > cap ? cap : g);
}

__global__ void iota64_kernel(int64_t* p, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = i;
}

template <typename K>
__global__ void narrow_keys_kernel(const int64_t* in, K* out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (K)in[i];
}

// sorted keys -> rowptr[r] = first slot whose key >= r  (keys < 0 are not present)
__global__ void rowptr_from_sorted_kernel(const int32_t* keys, int64_t n, int64_t n_rows, int64_t* rowptr) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t prev = (i == 0) ? -1 : (int64_t)keys[i - 1];
    const int64_t cur = (i == n) ? n_rows : (int64_t)keys[i];
    for (int64_t r = prev + 1; r <= cur; ++r) rowptr[r] = i;
  }
}

__global__ void gather_coo_kernel(const int64_t* perm, const int64_t* coo_col, const float* coo_val, int64_t n,
                                  int32_t* col, float* val, int64_t* perm_out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = perm[i];
    col[i] = (int32_t)coo_col[p];
    if (val) val[i] = coo_val ? coo_val[p] : 1.f;
    if (perm_out) perm_out[i] = p;
  }
}

// one warp per row: expand rowptr into per-edge row ids
__global__ void expand_rows_kernel(const int64_t* rowptr, int64_t n_rows, int32_t* rows) {
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = w; r < n_rows; r += nw) {
    const int64_t s = rowptr[r], e = rowptr[r + 1];
    for (int64_t k = s + lane; k < e; k += 32) rows[k] = (int32_t)r;
  }
}

__global__ void gather_transpose_kernel(const int64_t* perm, const int32_t* rows, const float* val, int64_t n,
                                        int32_t* col_t, float* val_t, int64_t* perm_t) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = perm[i];
    col_t[i] = rows[p];
    if (val_t) val_t[i] = val ? val[p] : 1.f;
    if (perm_t) perm_t[i] = p;
  }
}

template <typename A>
__global__ void mask_count_kernel(const A* adj, int64_t n_rows, int64_t n_cols, int64_t ld, int64_t* counts) {
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = w; r < n_rows; r += nw) {
    const A* row = adj + r * ld;
    int cnt = 0;
    for (int64_t c = lane; c < n_cols; c += 32) cnt += (row[c] > (A)0) ? 1 : 0;
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) counts[r] = cnt;
  }
}

template <typename A>
__global__ void mask_fill_kernel(const A* adj, int64_t n_rows, int64_t n_cols, int64_t ld, const int64_t* rowptr,
                                 int32_t* col) {
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = w; r < n_rows; r += nw) {
    const A* row = adj + r * ld;
    int64_t out = rowptr[r];
    for (int64_t c0 = 0; c0 < n_cols; c0 += 32) {
      const int64_t c = c0 + lane;
      const bool on = (c < n_cols) && (row[c] > (A)0);
      const unsigned b = __ballot_sync(0xffffffffu, on);
      if (on) col[out + __popc(b & ((1u << lane) - 1u))] = (int32_t)c;
      out += __popc(b);
    }
  }
}

template <typename I>
__global__ void idx_keys_kernel(const I* idx, int64_t n, int64_t n_table_rows, int32_t* keys, int32_t* pos) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = (int64_t)idx[i];
    // negative ids sort past the last table row and are cut off by rowptr[n_table_rows]
    keys[i] = (r < 0 || r >= n_table_rows) ? (int32_t)n_table_rows : (int32_t)r;
    pos[i] = (int32_t)i;
  }
}

inline int bits_for(int64_t n) {
  int b = 1;
  while (b < 63 && (1LL << b) < n) ++b;
  return b;
}

// ---- counter-based hash RNG (splitmix64) for the synthetic graphs ---------------
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9e3779b97f4a7c15ULL;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}
__device__ __forceinline__ double u01(uint64_t h) { return (double)(h >> 11) * (1.0 / 9007199254740992.0); }

__global__ void synth_degrees_kernel(int64_t n_rows, int64_t row_offset, double xmin, double exponent,
                                     int64_t max_degree, uint64_t seed, int64_t* deg) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_rows; i += (int64_t)gridDim.x * blockDim.x) {
    const double u = u01(mix64(seed ^ mix64((uint64_t)(i + row_offset))));
    // Pareto(xmin, exponent-1) degree, plus the self-loop
    double d = xmin * pow(1.0 - u, -1.0 / (exponent - 1.0));
    int64_t di = (int64_t)d;
    di = di < 0 ? 0 : di;
    di = di > max_degree ? max_degree : di;
    deg[i] = di + 1;
  }
}

// One warp per row.  Slot 0 is the self-loop (the +I of GCN/data_utils.py:78); neighbour
// k of a row with d other edges is drawn from stratum [k/d,(k+1)/d) of [0,1) and mapped
// through x -> n_cols * x^skew: ascending within the row and skewed to low ids (hubs).
__global__ void synth_fill_kernel(int64_t n_rows, int64_t row_offset, int64_t n_cols, const int64_t* rowptr,
                                  double skew, double p_local, int64_t window, uint64_t seed, int32_t* col) {
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = w; r < n_rows; r += nw) {
    const int64_t s = rowptr[r], e = rowptr[r + 1];
    const int64_t d = e - s - 1;
    const int64_t self = r + row_offset;
    const uint64_t rs = mix64(seed ^ mix64((uint64_t)self * 0x100000001b3ULL + 7));
    if (lane == 0) col[s] = (int32_t)self;
    for (int64_t k = lane; k < d; k += 32) {
      const uint64_t h = mix64(rs + (uint64_t)k);
      const double u = u01(h);
      int64_t c;
      if (p_local > 0.0 && u01(mix64(h ^ 0xa5a5a5a5a5a5a5a5ULL)) < p_local) {
        // community edge: within +-window of the row (a locality-preserving node ordering)
        const double w = u01(mix64(h + 0x1234567ULL));
        const int64_t wabs = window < 0 ? -window : window;
        const int64_t off = 1 + (int64_t)((double)wabs * w * w);
        c = (h & 1) ? self + off : self - off;
        c = c < 0 ? -c : c;
        c = c >= n_cols ? 2 * (n_cols - 1) - c : c;
        c = c < 0 ? 0 : c;
      } else {
        const double x = ((double)k + u) / (double)d;
        c = (int64_t)((double)n_cols * pow(x, skew));
        c = c >= n_cols ? n_cols - 1 : c;
        // window < 0: scatter the popular ids over the whole id range with a fixed bijection
        // (hubs of a real graph are not sorted by id); popularity itself is unchanged
        if (window < 0) c = (int64_t)(((uint64_t)c * 2654435761ULL) % (uint64_t)n_cols);  // c < 2^31: no overflow
      }
      c = c >= n_cols ? n_cols - 1 : c;
      col[s + 1 + k] = (int32_t)c;
    }
  }
}

__global__ void synth_values_kernel(int64_t n_rows, int64_t row_offset, const int64_t* rowptr, const int32_t* col,
                                    const int64_t* deg_all, float* val) {
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = w; r < n_rows; r += nw) {
    const int64_t s = rowptr[r], e = rowptr[r + 1];
    const double di = 1.0 / sqrt((double)deg_all[r + row_offset]);
    for (int64_t k = s + lane; k < e; k += 32) {
      // float64 product cast to fp32, as GCN/data_utils.py:54-60,65 does
      val[k] = (float)((1.0 / sqrt((double)deg_all[col[k]])) * di);
    }
  }
}

}  // namespace

extern "C" {

size_t gnn_build_csr_from_coo_workspace_size(int64_t nnz, int64_t n_rows) {
  size_t temp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, temp, (const int32_t*)nullptr, (int32_t*)nullptr, (const int64_t*)nullptr,
                                  (int64_t*)nullptr, nnz > 0 ? nnz : 1, 0, bits_for(n_rows + 1));
  const size_t n = (size_t)(nnz > 0 ? nnz : 1);
  return round_up(temp, 256) + 2 * round_up(n * sizeof(int32_t), 256) + 2 * round_up(n * sizeof(int64_t), 256) + 1024;
}

int gnn_build_csr_from_coo(const int64_t* coo_row, const int64_t* coo_col, const float* coo_val, int64_t nnz,
                           int64_t n_rows, int64_t n_cols, int64_t* rowptr, int32_t* col, float* val,
                           int64_t* perm_out, void* workspace, size_t workspace_bytes, gnn_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  GNN_REQUIRE(nnz >= 0 && n_rows >= 0 && n_cols >= 0, GNN_ERR_BAD_ARG, "negative size");
  GNN_REQUIRE(n_rows < 0x7fffffffLL && n_cols < 0x7fffffffLL, GNN_ERR_UNSUPPORTED, "dimension does not fit int32");
  GNN_REQUIRE(rowptr != nullptr, GNN_ERR_BAD_ARG, "null rowptr");
  if (nnz == 0) {
    GNN_CUDA(cudaMemsetAsync(rowptr, 0, (size_t)(n_rows + 1) * sizeof(int64_t), st));
    return GNN_OK;
  }
  GNN_REQUIRE(coo_row && coo_col && col, GNN_ERR_BAD_ARG, "null pointer");
  GNN_REQUIRE(workspace && workspace_bytes >= gnn_build_csr_from_coo_workspace_size(nnz, n_rows), GNN_ERR_WORKSPACE,
              "workspace too small (%zu bytes)", workspace_bytes);
  WsCursor ws{(char*)workspace, 0, workspace_bytes};
  int32_t* keys_in = ws.take<int32_t>(nnz);
  int32_t* keys_out = ws.take<int32_t>(nnz);
  int64_t* pos_in = ws.take<int64_t>(nnz);
  int64_t* pos_out = ws.take<int64_t>(nnz);
  size_t temp = 0;
  const int bits = bits_for(n_rows + 1);
  cub::DeviceRadixSort::SortPairs(nullptr, temp, keys_in, keys_out, pos_in, pos_out, nnz, 0, bits, st);
  void* temp_ptr = ws.take<char>(temp);
  narrow_keys_kernel<int32_t><<<grid_for(nnz), 256, 0, st>>>(coo_row, keys_in, nnz);
  GNN_LAUNCH_CHECK();
  iota64_kernel<<<grid_for(nnz), 256, 0, st>>>(pos_in, nnz);
  GNN_LAUNCH_CHECK();
  GNN_CUDA(cub::DeviceRadixSort::SortPairs(temp_ptr, temp, keys_in, keys_out, pos_in, pos_out, nnz, 0, bits, st));
  count_launch(2);
  rowptr_from_sorted_kernel<<<grid_for(nnz + 1), 256, 0, st>>>(keys_out, nnz, n_rows, rowptr);
  GNN_LAUNCH_CHECK();
  gather_coo_kernel<<<grid_for(nnz), 256, 0, st>>>(pos_out, coo_col, coo_val, nnz, col, val, perm_out);
  GNN_LAUNCH_CHECK();
  return GNN_OK;
}

size_t gnn_dense_mask_count_workspace_size(int64_t n_rows) {
  size_t temp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, temp, (const int64_t*)nullptr, (int64_t*)nullptr, n_rows + 1);
  return round_up(temp, 256) + round_up((size_t)(n_rows + 1) * sizeof(int64_t), 256) + 512;
}

int gnn_dense_mask_count(const void* adj, int adj_dtype, int64_t n_rows, int64_t n_cols, int64_t ld, int64_t* rowptr,
                         void* workspace, size_t workspace_bytes, gnn_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  GNN_REQUIRE(n_rows >= 0 && n_cols >= 0 && ld >= n_cols, GNN_ERR_BAD_ARG, "bad size");
  GNN_REQUIRE(adj_dtype == 0 || adj_dtype == 1, GNN_ERR_BAD_ARG, "adj_dtype must be 0 (fp32) or 1 (fp64)");
  GNN_REQUIRE(n_cols < 0x7fffffffLL, GNN_ERR_UNSUPPORTED, "n_cols does not fit int32");
  GNN_REQUIRE(rowptr && (adj || n_rows * n_cols == 0), GNN_ERR_BAD_ARG, "null pointer");
  GNN_REQUIRE(workspace && workspace_bytes >= gnn_dense_mask_count_workspace_size(n_rows), GNN_ERR_WORKSPACE,
              "workspace too small");
  WsCursor ws{(char*)workspace, 0, workspace_bytes};
  int64_t* counts = ws.take<int64_t>(n_rows + 1);
  size_t temp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, temp, counts, rowptr, n_rows + 1, st);
  void* temp_ptr = ws.take<char>(temp);
  GNN_CUDA(cudaMemsetAsync(counts, 0, (size_t)(n_rows + 1) * sizeof(int64_t), st));
  if (n_rows > 0 && n_cols > 0) {
    const int grid = grid_for(n_rows * 32);
    if (adj_dtype == 0) mask_count_kernel<float><<<grid, 256, 0, st>>>((const float*)adj, n_rows, n_cols, ld, counts);
    else mask_count_kernel<double><<<grid, 256, 0, st>>>((const double*)adj, n_rows, n_cols, ld, counts);
    GNN_LAUNCH_CHECK();
  }
  GNN_CUDA(cub::DeviceScan::ExclusiveSum(temp_ptr, temp, counts, rowptr, n_rows + 1, st));
  count_launch(1);
  return GNN_OK;
}

int gnn_dense_mask_fill(const void* adj, int adj_dtype, int64_t n_rows, int64_t n_cols, int64_t ld,
                        const int64_t* rowptr, int32_t* col, gnn_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  GNN_REQUIRE(n_rows >= 0 && n_cols >= 0 && ld >= n_cols, GNN_ERR_BAD_ARG, "bad size");
  GNN_REQUIRE(adj_dtype == 0 || adj_dtype == 1, GNN_ERR_BAD_ARG, "adj_dtype must be 0 (fp32) or 1 (fp64)");
  if (n_rows == 0 || n_cols == 0) return GNN_OK;
  GNN_REQUIRE(adj && rowptr && col, GNN_ERR_BAD_ARG, "null pointer");
  const int grid = grid_for(n_rows * 32);
  if (adj_dtype == 0) mask_fill_kernel<float><<<grid, 256, 0, st>>>((const float*)adj, n_rows, n_cols, ld, rowptr, col);
  else mask_fill_kernel<double><<<grid, 256, 0, st>>>((const double*)adj, n_rows, n_cols, ld, rowptr, col);
  GNN_LAUNCH_CHECK();
  return GNN_OK;
}

size_t gnn_csr_transpose_workspace_size(int64_t nnz, int64_t /*n_rows*/, int64_t n_cols) {
  size_t temp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, temp, (const int32_t*)nullptr, (int32_t*)nullptr, (const int64_t*)nullptr,
                                  (int64_t*)nullptr, nnz > 0 ? nnz : 1, 0, bits_for(n_cols + 1));
  const size_t n = (size_t)(nnz > 0 ? nnz : 1);
  return round_up(temp, 256) + 2 * round_up(n * sizeof(int32_t), 256) + 2 * round_up(n * sizeof(int64_t), 256) + 1280;
}

int gnn_csr_transpose(const int64_t* rowptr, const int32_t* col, const float* val, int64_t n_rows, int64_t n_cols,
                      int64_t nnz, int64_t* rowptr_t, int32_t* col_t, float* val_t, int64_t* perm_t, void* workspace,
                      size_t workspace_bytes, gnn_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  GNN_REQUIRE(nnz >= 0 && n_rows >= 0 && n_cols >= 0, GNN_ERR_BAD_ARG, "negative size");
  GNN_REQUIRE(n_rows < 0x7fffffffLL && n_cols < 0x7fffffffLL, GNN_ERR_UNSUPPORTED, "dimension does not fit int32");
  GNN_REQUIRE(rowptr_t != nullptr, GNN_ERR_BAD_ARG, "null rowptr_t");
  if (nnz == 0) {
    GNN_CUDA(cudaMemsetAsync(rowptr_t, 0, (size_t)(n_cols + 1) * sizeof(int64_t), st));
    return GNN_OK;
  }
  GNN_REQUIRE(rowptr && col && col_t, GNN_ERR_BAD_ARG, "null pointer");
  GNN_REQUIRE(workspace && workspace_bytes >= gnn_csr_transpose_workspace_size(nnz, n_rows, n_cols), GNN_ERR_WORKSPACE,
              "workspace too small");
  WsCursor ws{(char*)workspace, 0, workspace_bytes};
  int32_t* rows = ws.take<int32_t>(nnz);
  int32_t* keys_out = ws.take<int32_t>(nnz);
  int64_t* pos_in = ws.take<int64_t>(nnz);
  int64_t* pos_out = ws.take<int64_t>(nnz);
  size_t temp = 0;
  const int bits = bits_for(n_cols + 1);
  cub::DeviceRadixSort::SortPairs(nullptr, temp, col, keys_out, pos_in, pos_out, nnz, 0, bits, st);
  void* temp_ptr = ws.take<char>(temp);
  expand_rows_kernel<<<grid_for(n_rows * 32), 256, 0, st>>>(rowptr, n_rows, rows);
  GNN_LAUNCH_CHECK();
  iota64_kernel<<<grid_for(nnz), 256, 0, st>>>(pos_in, nnz);
  GNN_LAUNCH_CHECK();
  // stable LSD radix sort on the column id: equal columns keep ascending original slot,
  // i.e. ascending original row.
  GNN_CUDA(cub::DeviceRadixSort::SortPairs(temp_ptr, temp, col, keys_out, pos_in, pos_out, nnz, 0, bits, st));
  count_launch(2);
  rowptr_from_sorted_kernel<<<grid_for(nnz + 1), 256, 0, st>>>(keys_out, nnz, n_cols, rowptr_t);
  GNN_LAUNCH_CHECK();
  gather_transpose_kernel<<<grid_for(nnz), 256, 0, st>>>(pos_out, rows, val, nnz, col_t, val_t, perm_t);
  GNN_LAUNCH_CHECK();
  return GNN_OK;
}

size_t gnn_index_block_transpose_workspace_size(int64_t n_idx, int64_t n_table_rows) {
  size_t temp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, temp, (const int32_t*)nullptr, (int32_t*)nullptr, (const int32_t*)nullptr,
                                  (int32_t*)nullptr, n_idx > 0 ? n_idx : 1, 0, bits_for(n_table_rows + 2));
  const size_t n = (size_t)(n_idx > 0 ? n_idx : 1);
  return round_up(temp, 256) + 3 * round_up(n * sizeof(int32_t), 256) + 1280;
}

int gnn_index_block_transpose(const void* idx, int idx_bits, int64_t n_idx, int64_t n_table_rows, int64_t* rowptr_t,
                              int32_t* pos_t, void* workspace, size_t workspace_bytes, gnn_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  GNN_REQUIRE(n_idx >= 0 && n_table_rows >= 0, GNN_ERR_BAD_ARG, "negative size");
  GNN_REQUIRE(idx_bits == 32 || idx_bits == 64, GNN_ERR_BAD_ARG, "idx_bits must be 32 or 64");
  GNN_REQUIRE(n_idx < 0x7fffffffLL && n_table_rows < 0x7ffffffeLL, GNN_ERR_UNSUPPORTED, "size does not fit int32");
  GNN_REQUIRE(rowptr_t != nullptr, GNN_ERR_BAD_ARG, "null rowptr_t");
  if (n_idx == 0) {
    GNN_CUDA(cudaMemsetAsync(rowptr_t, 0, (size_t)(n_table_rows + 1) * sizeof(int64_t), st));
    return GNN_OK;
  }
  GNN_REQUIRE(idx && pos_t, GNN_ERR_BAD_ARG, "null pointer");
  GNN_REQUIRE(workspace && workspace_bytes >= gnn_index_block_transpose_workspace_size(n_idx, n_table_rows),
              GNN_ERR_WORKSPACE, "workspace too small");
  WsCursor ws{(char*)workspace, 0, workspace_bytes};
  int32_t* keys_in = ws.take<int32_t>(n_idx);
  int32_t* keys_out = ws.take<int32_t>(n_idx);
  int32_t* pos_in = ws.take<int32_t>(n_idx);
  size_t temp = 0;
  const int bits = bits_for(n_table_rows + 2);
  cub::DeviceRadixSort::SortPairs(nullptr, temp, keys_in, keys_out, pos_in, pos_t, n_idx, 0, bits, st);
  void* temp_ptr = ws.take<char>(temp);
  if (idx_bits == 32)
    idx_keys_kernel<int32_t><<<grid_for(n_idx), 256, 0, st>>>((const int32_t*)idx, n_idx, n_table_rows, keys_in, pos_in);
  else
    idx_keys_kernel<int64_t><<<grid_for(n_idx), 256, 0, st>>>((const int64_t*)idx, n_idx, n_table_rows, keys_in, pos_in);
  GNN_LAUNCH_CHECK();
  GNN_CUDA(cub::DeviceRadixSort::SortPairs(temp_ptr, temp, keys_in, keys_out, pos_in, pos_t, n_idx, 0, bits, st));
  count_launch(2);
  // rowptr over n_table_rows+1 "rows": the extra row collects the skipped ids; only the
  // first n_table_rows+1 offsets are written to the caller.
  rowptr_from_sorted_kernel<<<grid_for(n_idx + 1), 256, 0, st>>>(keys_out, n_idx, n_table_rows, rowptr_t);
  GNN_LAUNCH_CHECK();
  return GNN_OK;
}

int gnn_synth_powerlaw_degrees(int64_t n_rows, int64_t row_offset, double mean_degree, double exponent,
                               int64_t max_degree, uint64_t seed, int64_t* deg, gnn_stream_t stream) {
  GNN_REQUIRE(n_rows >= 0 && deg && exponent > 2.0 && mean_degree > 0, GNN_ERR_BAD_ARG,
              "bad argument (exponent must exceed 2 for a finite mean)");
  if (n_rows == 0) return GNN_OK;
  // E[Pareto] = xmin*(a)/(a-1) with a = exponent-1; floor() removes ~0.5 on average
  const double a = exponent - 1.0;
  const double xmin = (mean_degree + 0.5) * (a - 1.0) / a;
  synth_degrees_kernel<<<grid_for(n_rows), 256, 0, (cudaStream_t)stream>>>(n_rows, row_offset, xmin, exponent,
                                                                           max_degree, seed, deg);
  GNN_LAUNCH_CHECK();
  return GNN_OK;
}

int gnn_synth_powerlaw_fill(int64_t n_rows, int64_t row_offset, int64_t n_cols, const int64_t* rowptr, double skew,
                            double p_local, int64_t window, uint64_t seed, int32_t* col, gnn_stream_t stream) {
  GNN_REQUIRE(n_rows >= 0 && n_cols > 0 && n_cols < 0x7fffffffLL && rowptr && col && skew >= 1.0 && p_local >= 0.0 &&
                  p_local <= 1.0,
              GNN_ERR_BAD_ARG, "bad argument");
  if (n_rows == 0) return GNN_OK;
  synth_fill_kernel<<<grid_for(n_rows * 32), 256, 0, (cudaStream_t)stream>>>(n_rows, row_offset, n_cols, rowptr, skew,
                                                                             p_local, window, seed, col);
  GNN_LAUNCH_CHECK();
  return GNN_OK;
}

int gnn_synth_gcn_values(int64_t n_rows, int64_t row_offset, const int64_t* rowptr, const int32_t* col,
                         const int64_t* deg_all, float* val, gnn_stream_t stream) {
  GNN_REQUIRE(n_rows >= 0 && rowptr && col && deg_all && val, GNN_ERR_BAD_ARG, "bad argument");
  if (n_rows == 0) return GNN_OK;
  synth_values_kernel<<<grid_for(n_rows * 32), 256, 0, (cudaStream_t)stream>>>(n_rows, row_offset, rowptr, col,
                                                                               deg_all, val);
  GNN_LAUNCH_CHECK();
  return GNN_OK;
}

}  // extern "C"
