// Edge-gradient SDDMM: out[e] = <A[row[e], :], B[col[e], :]> over the edges of a sparse pattern.
//
// This is the backward of a sparse-times-dense product with respect to the sparse VALUES:
//   /root/reference GAT/models/layers.py:55-61 (SpecialSpmmFunction.backward) computes it as
//   `grad_output.matmul(b.t())` — a dense N x N matrix — and then picks the nnz entries
//   `[row * N + col]`; here only the nnz dot products are formed.
// The same contraction is what a GCN over a LEARNED adjacency needs (GTN/models/GTN.py:49-52
// `gcn_conv` on the composed metapath matrix) once that matrix is held as a sparse pattern.
//
// Edge-parallel, so perfectly balanced on power-law graphs: a team of GROUP lanes (chosen from F)
// owns one edge slot, a warp holds 32/GROUP slots and U edges per slot are in flight.  For edges
// sorted by row (CSR order, `adj.nonzero()`) consecutive teams read the same A row (L1 hits) and
// the bytes that matter are the gathered B rows: 8 (ids) + F*4 (B row) + 4 (out) per edge.
// Deterministic: fixed shuffle tree per edge, no atomics.
#include "common.cuh"

using namespace gnn;

namespace {

struct SddmmArgs {
  const int32_t* row32;
  const int64_t* row64;
  const int32_t* col32;
  const int64_t* col64;
  const float* A;
  int64_t lda;
  const float* B;
  int64_t ldb;
  int32_t F;
  int64_t nnz;
  float* out;
};

template <int VEC, int GROUP, int U, bool SINGLE>
__global__ void __launch_bounds__(256) sddmm_kernel(const SddmmArgs a) {
  constexpr int SLOTS = 32 / GROUP;
  constexpr int EPW = SLOTS * U;  // edges per warp per iteration
  const int lane = threadIdx.x & 31;
  const int gl = lane % GROUP;
  const int slot = lane / GROUP;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t base = warp * EPW; base < a.nnz; base += n_warps * EPW) {  // warp-uniform trip count
    const float* ar[U];
    const float* br[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t e = base + (int64_t)u * SLOTS + slot;
      ok[u] = e < a.nnz;
      int64_t r = 0, c = 0;
      if (ok[u]) {
        r = a.row32 ? (int64_t)__ldg(a.row32 + e) : __ldg(a.row64 + e);
        c = a.col32 ? (int64_t)__ldg(a.col32 + e) : __ldg(a.col64 + e);
      }
      ar[u] = a.A + r * a.lda;
      br[u] = a.B + c * a.ldb;
    }
    float part[U];
    if (SINGLE) {
      // F <= GROUP*VEC: one vector per lane and operand; all 2U loads are issued before the math
      float av[U][VEC], bv[U][VEC];
      const int f0 = gl * VEC;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool live = ok[u] && f0 < a.F;
        if (live) {
          VecIO<float, VEC>::load(ar[u] + f0, av[u]);
          VecIO<float, VEC>::load(br[u] + f0, bv[u]);
        } else {
#pragma unroll
          for (int i = 0; i < VEC; ++i) av[u][i] = bv[u][i] = 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < VEC; ++i) s = fmaf(av[u][i], bv[u][i], s);
        part[u] = s;
      }
    } else {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float s = 0.f;
        if (ok[u]) {
          for (int f0 = gl * VEC; f0 < a.F; f0 += GROUP * VEC) {
            float av[VEC], bv[VEC];
            VecIO<float, VEC>::load(ar[u] + f0, av);
            VecIO<float, VEC>::load(br[u] + f0, bv);
#pragma unroll
            for (int i = 0; i < VEC; ++i) s = fmaf(av[i], bv[i], s);
          }
        }
        part[u] = s;
      }
    }
    // combine the GROUP lanes of every team (fixed butterfly; offsets stay inside the team)
#pragma unroll
    for (int off = GROUP >> 1; off > 0; off >>= 1)
#pragma unroll
      for (int u = 0; u < U; ++u) part[u] += __shfl_xor_sync(0xffffffffu, part[u], off);
    if (gl == 0) {
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (ok[u]) a.out[base + (int64_t)u * SLOTS + slot] = part[u];
    }
  }
}

template <int VEC, int GROUP, bool SINGLE>
int sddmm_launch(const SddmmArgs& a, cudaStream_t st) {
  constexpr int U = 4;
  constexpr int EPW = (32 / GROUP) * U;
  const int64_t warps = (a.nnz + EPW - 1) / EPW;
  int64_t grid = (warps + 7) / 8;
  const int64_t cap = (int64_t)num_sms() * 32;  // grid-stride beyond 8 CTAs x 148 SMs x 4 waves
  grid = grid > cap ? cap : grid;
  sddmm_kernel<VEC, GROUP, U, SINGLE><<<(unsigned)grid, 256, 0, st>>>(a);
  GNN_LAUNCH_CHECK();
  return GNN_OK;
}

template <int VEC>
int sddmm_dispatch(const SddmmArgs& a, cudaStream_t st) {
  const int nvec = (a.F + VEC - 1) / VEC;
  if (nvec <= 1) return sddmm_launch<VEC, 1, true>(a, st);
  if (nvec <= 2) return sddmm_launch<VEC, 2, true>(a, st);
  if (nvec <= 4) return sddmm_launch<VEC, 4, true>(a, st);
  if (nvec <= 8) return sddmm_launch<VEC, 8, true>(a, st);
  if (nvec <= 16) return sddmm_launch<VEC, 16, true>(a, st);
  if (nvec <= 32) return sddmm_launch<VEC, 32, true>(a, st);
  return sddmm_launch<VEC, 32, false>(a, st);
}

}  // namespace

extern "C" {

int gnn_sddmm_coo_f32(const void* row, const void* col, int idx_bits, int64_t nnz, const float* A, int64_t lda,
                      const float* B, int64_t ldb, int32_t F, float* out, gnn_stream_t stream) {
  GNN_REQUIRE(nnz >= 0 && F >= 0, GNN_ERR_BAD_ARG, "negative size");
  if (nnz == 0) return GNN_OK;
  GNN_REQUIRE(row && col && out, GNN_ERR_BAD_ARG, "null pointer (row/col/out)");
  GNN_REQUIRE(idx_bits == 32 || idx_bits == 64, GNN_ERR_BAD_ARG, "idx_bits must be 32 or 64");
  GNN_REQUIRE(F == 0 || (A && B), GNN_ERR_BAD_ARG, "null operand");
  GNN_REQUIRE(lda >= F && ldb >= F, GNN_ERR_BAD_ARG, "leading dimension smaller than F");
  SddmmArgs a{};
  a.row32 = idx_bits == 32 ? (const int32_t*)row : nullptr;
  a.row64 = idx_bits == 64 ? (const int64_t*)row : nullptr;
  a.col32 = idx_bits == 32 ? (const int32_t*)col : nullptr;
  a.col64 = idx_bits == 64 ? (const int64_t*)col : nullptr;
  a.A = A;
  a.lda = lda;
  a.B = B;
  a.ldb = ldb;
  a.F = F;
  a.nnz = nnz;
  a.out = out;
  const bool vec4 = F % 4 == 0 && lda % 4 == 0 && ldb % 4 == 0 && aligned_to(A, 16) && aligned_to(B, 16);
  return vec4 ? sddmm_dispatch<4>(a, (cudaStream_t)stream) : sddmm_dispatch<1>(a, (cudaStream_t)stream);
}

}  // extern "C"
