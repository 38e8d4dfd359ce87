// GCN aggregation Y = Â·X: C ABI over the row-parallel CSR kernel.
// Replaces torch.spmm(adj, support) at /root/reference GCN/GCN.py:43 and, on the
// transposed CSR, its autograd backward Âᵀ·dY.
#include "rowreduce.cuh"

using namespace gnn;

namespace {

template <typename T>
int spmm_impl(const int64_t* rowptr, const int32_t* col, const float* val, const T* X, T* Y, int64_t n_rows,
              int64_t n_cols, int32_t F, int64_t ldx, int64_t ldy, const int64_t* long_rows, int64_t n_long,
              cudaStream_t st) {
  GNN_REQUIRE(n_rows >= 0 && n_cols >= 0 && F >= 0, GNN_ERR_BAD_ARG, "negative size");
  if (n_rows == 0 || F == 0) return GNN_OK;
  // col may be null only for a graph without edges (nnz lives on the device; rowptr decides)
  GNN_REQUIRE(rowptr && X && Y, GNN_ERR_BAD_ARG, "null pointer (rowptr/X/Y)");
  GNN_REQUIRE(ldx >= F && ldy >= F, GNN_ERR_BAD_ARG, "leading dimension smaller than F (ldx=%lld ldy=%lld F=%d)",
              (long long)ldx, (long long)ldy, F);
  GNN_REQUIRE(n_cols < 0x7fffffffLL, GNN_ERR_UNSUPPORTED, "n_cols does not fit int32 column ids");
  GNN_REQUIRE(aligned_to(X, sizeof(T)) && aligned_to(Y, sizeof(T)), GNN_ERR_MISALIGNED, "X/Y not element aligned");
  RowArgs<T> a{};
  a.rowptr = rowptr;
  a.fanout = 0;
  a.col32 = col;
  a.col64 = nullptr;
  a.val = val;
  a.src_div = 0;
  a.scale = 1.f;
  a.X = X;
  a.ldx = ldx;
  a.Y = Y;
  a.ldy = ldy;
  a.n_rows = n_rows;
  a.F = F;
  a.skip_deg_gt = (n_long > 0) ? (int64_t)tuning("spmm.long_row", 2048) : 0;
  a.argmax = nullptr;
  int rc = launch_row_reduce<T, 0>(a, st);
  if (rc != GNN_OK) return rc;
  if (n_long > 0) {
    GNN_REQUIRE(long_rows != nullptr, GNN_ERR_BAD_ARG, "n_long > 0 but long_rows is null");
    rc = launch_row_reduce_long<T>(a, long_rows, n_long, st);
  }
  return rc;
}

}  // namespace

extern "C" {

size_t gnn_spmm_csr_workspace_size(int64_t, int64_t, int32_t) { return 0; }

int gnn_spmm_csr_f32(const int64_t* rowptr, const int32_t* col, const float* val, const float* X, float* Y,
                     int64_t n_rows, int64_t n_cols, int32_t F, int64_t ldx, int64_t ldy, gnn_stream_t stream) {
  return spmm_impl<float>(rowptr, col, val, X, Y, n_rows, n_cols, F, ldx, ldy, nullptr, 0, (cudaStream_t)stream);
}

int gnn_spmm_csr_bf16(const int64_t* rowptr, const int32_t* col, const float* val, const void* X, void* Y,
                      int64_t n_rows, int64_t n_cols, int32_t F, int64_t ldx, int64_t ldy, gnn_stream_t stream) {
  return spmm_impl<__nv_bfloat16>(rowptr, col, val, (const __nv_bfloat16*)X, (__nv_bfloat16*)Y, n_rows, n_cols, F, ldx,
                                  ldy, nullptr, 0, (cudaStream_t)stream);
}

int gnn_spmm_csr_planned_f32(const int64_t* rowptr, const int32_t* col, const float* val, const float* X, float* Y,
                             int64_t n_rows, int64_t n_cols, int32_t F, int64_t ldx, int64_t ldy,
                             const int64_t* long_rows, int64_t n_long, void* /*workspace*/, size_t /*workspace_bytes*/,
                             gnn_stream_t stream) {
  return spmm_impl<float>(rowptr, col, val, X, Y, n_rows, n_cols, F, ldx, ldy, long_rows, n_long,
                          (cudaStream_t)stream);
}

}  // extern "C"
