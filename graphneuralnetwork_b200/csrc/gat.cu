// Fused multi-head graph attention aggregation (GAT / HAN node-level attention).
// Replaces, at /root/reference:
//   GAT/models/layers.py:25-32  (== HAN/models/NodeAttention.py:25-33)
//       the [N,N,2F'] pair tensor, masked softmax and dense attention·Wh matmul
//   GAT/models/layers.py:108-122  the edge-list variant exp(-LeakyReLU) / rowsum
//   GAT/models/layers.py:55-64    SpecialSpmmFunction.backward (dense N×N edge gradient)
//
// Forward, one warp per destination row, all H heads in one pass over the CSR row:
//   A  lane-per-edge:  logit[k,h] = ±LeakyReLU(s[i,h] + t[col_k,h]) staged in shared memory
//   B  per-head chunk max by warp shuffle, running max/scale (softmax mode)
//   B2 logit -> p = exp(logit - max) in place (one exp per edge and head)
//   C  lane-per-column: acc += p[k,head(col)] * Wh[col_k, column]  (coalesced row gathers)
// Backward: kernel A (CSR rows) computes the per-edge attention weight and the edge
// gradient dz (SDDMM dOut_i·Wh_j), stashes both per edge and reduces d_s by row; kernel B
// (transposed CSR rows) reduces d_Wh and d_t in source order.  Ordered sums only.
#include "common.cuh"

using namespace gnn;

namespace {

constexpr int kGatWarps = 4;

struct GatArgs {
  const int64_t* rowptr;
  const int32_t* col;
  const int64_t* perm;  // backward B: transposed slot -> forward edge slot
  const float* Wh;
  int64_t ldw;
  const float* s;
  const float* t;
  int64_t n;
  int H, Hp, Fp, HF;
  float alpha;
  int mode;
  int elu;
  const float* col_mean;
  const float* keep;
  float* out;
  int64_t ldo;
  float* row_max;
  float* row_sum;
  int SE;
  // backward
  const float* out_pre;
  const float* d_out;
  const float* rowdot;
  float* edge_w;
  float* edge_dz;
  float* d_Wh;
  int64_t ld_dwh;
  float* d_s;
  float* d_t;
};

__device__ __forceinline__ float act_elu(float x, int elu) {
  if (elu >= 1) x = elu1(x);
  if (elu >= 2) x = elu1(x);
  return x;
}

template <int CPL>
__global__ void __launch_bounds__(kGatWarps * 32) gat_fwd_kernel(const GatArgs a) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * kGatWarps + warp;
  if (i >= a.n) return;  // warp-uniform; no block-wide barrier below
  const int H = a.H, Hp = a.Hp, SE = a.SE;
  float* logit = sm + (size_t)warp * (SE * H + SE + 64);
  int* cols = reinterpret_cast<int*>(logit + SE * H);
  float* mh = reinterpret_cast<float*>(cols + SE);
  float* sc = mh + 32;

  int hc[CPL];
  bool cv[CPL];
  float acc[CPL], l[CPL];
#pragma unroll
  for (int c = 0; c < CPL; ++c) {
    const int ci = lane + 32 * c;
    cv[c] = ci < a.HF;
    hc[c] = cv[c] ? ci / a.Fp : 0;
    acc[c] = 0.f;
    l[c] = 0.f;
  }
  const int64_t e0 = __ldg(a.rowptr + i), e1 = __ldg(a.rowptr + i + 1);
  const int64_t d = e1 - e0;
  float* orow = a.out + i * a.ldo;
  if (d == 0) {
    // GAT/models/layers.py:28-30: an all -9e15 row soft-maxes to the uniform 1/N over ALL nodes
#pragma unroll
    for (int c = 0; c < CPL; ++c)
      if (cv[c]) {
        const int ci = lane + 32 * c;
        orow[ci] = act_elu(a.col_mean ? a.col_mean[ci] : 0.f, a.elu);
        if (ci % a.Fp == 0) {
          if (a.row_max) a.row_max[i * H + hc[c]] = 0.f;
          if (a.row_sum) a.row_sum[i * H + hc[c]] = 0.f;
        }
      }
    return;
  }
  if (lane < Hp) {
    mh[lane] = (a.mode == GNN_GAT_SOFTMAX) ? -INFINITY : 0.f;
    sc[lane] = 1.f;
  }
  __syncwarp();
  const int ngrp = 32 / Hp;
  const int hsub = lane % Hp, g = lane / Hp;

  for (int64_t c0 = 0; c0 < d; c0 += SE) {
    const int ne = (int)((d - c0) < SE ? (d - c0) : SE);
    // A: logits, one lane per edge
    for (int k = lane; k < ne; k += 32) {
      const int j = __ldg(a.col + e0 + c0 + k);
      cols[k] = j;
      for (int h = 0; h < H; ++h) {
        const float z = __ldg(a.s + i * H + h) + __ldg(a.t + (int64_t)j * H + h);
        float e = z > 0.f ? z : a.alpha * z;
        if (a.mode == GNN_GAT_EXPNEG) e = -e;
        logit[k * H + h] = e;
      }
    }
    __syncwarp();
    // B: running per-head max
    if (a.mode == GNN_GAT_SOFTMAX) {
      float mx = -INFINITY;
      if (hsub < H)
        for (int k = g; k < ne; k += ngrp) mx = fmaxf(mx, logit[k * H + hsub]);
      for (int o = Hp; o < 32; o <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      if (lane < H) {
        const float mo = mh[lane];
        const float mn = fmaxf(mo, mx);
        sc[lane] = (mo == -INFINITY) ? 0.f : expf(mo - mn);
        mh[lane] = mn;
      }
      __syncwarp();
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        const float f = sc[hc[c]];
        acc[c] *= f;
        l[c] *= f;
      }
    }
    // B2: logits -> un-normalised probabilities, one exp per (edge, head)
    for (int idx = lane; idx < ne * H; idx += 32) logit[idx] = expf(logit[idx] - mh[idx % H]);
    __syncwarp();
    // C: weighted accumulation, one lane per output column
#pragma unroll 4
    for (int k = 0; k < ne; ++k) {
      const int j = cols[k];
      const float* wr = a.Wh + (int64_t)j * a.ldw;
#pragma unroll
      for (int c = 0; c < CPL; ++c)
        if (cv[c]) {
          const float p = logit[k * H + hc[c]];
          l[c] += p;
          const float w = a.keep ? p * __ldg(a.keep + (e0 + c0 + k) * H + hc[c]) : p;
          acc[c] = fmaf(w, __ldg(wr + lane + 32 * c), acc[c]);
        }
    }
    __syncwarp();
  }
#pragma unroll
  for (int c = 0; c < CPL; ++c)
    if (cv[c]) {
      const int ci = lane + 32 * c;
      orow[ci] = act_elu(acc[c] / l[c], a.elu);
      if (ci % a.Fp == 0) {
        if (a.row_max) a.row_max[i * H + hc[c]] = mh[hc[c]];
        if (a.row_sum) a.row_sum[i * H + hc[c]] = l[c];
      }
    }
}

__global__ void __launch_bounds__(256) gat_scores_kernel(const float* __restrict__ Wh, int64_t ldw,
                                                         const float* __restrict__ a_src,
                                                         const float* __restrict__ a_dst, int64_t n, int H, int Fp,
                                                         float* __restrict__ s, float* __restrict__ t) {
  const int64_t total = n * H;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
    const int h = (int)(p % H);
    const int64_t i = p / H;
    const float* w = Wh + i * ldw + h * Fp;
    float ss = 0.f, tt = 0.f;
    for (int f = 0; f < Fp; ++f) {
      const float x = __ldg(w + f);
      ss = fmaf(x, __ldg(a_src + h * Fp + f), ss);
      tt = fmaf(x, __ldg(a_dst + h * Fp + f), tt);
    }
    s[p] = ss;
    t[p] = tt;
  }
}

__global__ void __launch_bounds__(256) gat_rowdot_kernel(const float* __restrict__ d_out,
                                                         const float* __restrict__ out_pre, int64_t ldo, int64_t n,
                                                         int H, int Fp, float* __restrict__ rowdot) {
  const int64_t total = n * H;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
    const int h = (int)(p % H);
    const int64_t i = p / H;
    const float* a = d_out + i * ldo + h * Fp;
    const float* b = out_pre + i * ldo + h * Fp;
    float acc = 0.f;
    for (int f = 0; f < Fp; ++f) acc = fmaf(__ldg(a + f), __ldg(b + f), acc);
    rowdot[p] = acc;
  }
}

// Backward A: one warp per destination row i (forward CSR).
__global__ void __launch_bounds__(kGatWarps * 32) gat_bwd_rows_kernel(const GatArgs a) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * kGatWarps + warp;
  if (i >= a.n) return;
  const int H = a.H, Hp = a.Hp, SE = a.SE, HF = a.HF, Fp = a.Fp;
  float* dz_s = sm + (size_t)warp * (SE * H + HF + 128);
  float* dout_s = dz_s + SE * H;
  float* cst = dout_s + HF;  // [4][32]: s_i, m_i, 1/l_i, D_i
  for (int c = lane; c < HF; c += 32) dout_s[c] = __ldg(a.d_out + i * a.ldo + c);
  if (lane < H) {
    cst[lane] = __ldg(a.s + i * H + lane);
    cst[32 + lane] = __ldg(a.row_max + i * H + lane);
    const float l = __ldg(a.row_sum + i * H + lane);
    cst[64 + lane] = l > 0.f ? 1.f / l : 0.f;
    cst[96 + lane] = __ldg(a.rowdot + i * H + lane);
  }
  __syncwarp();
  const int64_t e0 = __ldg(a.rowptr + i), e1 = __ldg(a.rowptr + i + 1);
  const int64_t d = e1 - e0;
  const int ngrp = 32 / Hp;
  const int hsub = lane % Hp, g = lane / Hp;
  float dsum = 0.f;
  for (int64_t c0 = 0; c0 < d; c0 += SE) {
    const int ne = (int)((d - c0) < SE ? (d - c0) : SE);
    for (int k = lane; k < ne; k += 32) {
      const int64_t e = e0 + c0 + k;
      const int j = __ldg(a.col + e);
      const float* wr = a.Wh + (int64_t)j * a.ldw;
      for (int h = 0; h < H; ++h) {
        float dot = 0.f;
        for (int f = 0; f < Fp; ++f) dot = fmaf(dout_s[h * Fp + f], __ldg(wr + h * Fp + f), dot);
        const float z = cst[h] + __ldg(a.t + (int64_t)j * H + h);
        float slope = z > 0.f ? 1.f : a.alpha;
        float ee = z * slope;
        if (a.mode == GNN_GAT_EXPNEG) {
          ee = -ee;
          slope = -slope;
        }
        const float al = expf(ee - cst[32 + h]) * cst[64 + h];
        const float kp = a.keep ? __ldg(a.keep + e * H + h) : 1.f;
        const float dz = al * (kp * dot - cst[96 + h]) * slope;
        a.edge_w[e * H + h] = kp * al;
        a.edge_dz[e * H + h] = dz;
        dz_s[k * H + h] = dz;
      }
    }
    __syncwarp();
    if (hsub < H)
      for (int k = g; k < ne; k += ngrp) dsum += dz_s[k * H + hsub];
    __syncwarp();
  }
  for (int o = Hp; o < 32; o <<= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
  if (lane < H) a.d_s[i * H + lane] = dsum;
}

// Backward B: one warp per source node j (transposed CSR), sources ascending.
template <int CPL>
__global__ void __launch_bounds__(kGatWarps * 32) gat_bwd_cols_kernel(const GatArgs a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t j = (int64_t)blockIdx.x * kGatWarps + warp;
  if (j >= a.n) return;
  const int H = a.H, Hp = a.Hp;
  int hc[CPL];
  bool cv[CPL];
  float acc[CPL];
#pragma unroll
  for (int c = 0; c < CPL; ++c) {
    const int ci = lane + 32 * c;
    cv[c] = ci < a.HF;
    hc[c] = cv[c] ? ci / a.Fp : 0;
    acc[c] = 0.f;
  }
  const int64_t e0 = __ldg(a.rowptr + j), e1 = __ldg(a.rowptr + j + 1);
  const int ngrp = 32 / Hp;
  const int hsub = lane % Hp, g = lane / Hp;
  float dtsum = 0.f;
  for (int64_t c0 = e0; c0 < e1; c0 += 32) {
    const int ne = (int)((e1 - c0) < 32 ? (e1 - c0) : 32);
    int il = 0;
    int64_t pl = 0;
    if (lane < ne) {
      il = __ldg(a.col + c0 + lane);
      pl = __ldg(a.perm + c0 + lane);
    }
    for (int k0 = 0; k0 < ne; k0 += ngrp) {
      const int kk = k0 + g;
      const int64_t p = __shfl_sync(0xffffffffu, pl, kk & 31);
      if (kk < ne && hsub < H) dtsum += __ldg(a.edge_dz + p * H + hsub);
    }
#pragma unroll 4
    for (int k = 0; k < ne; ++k) {
      const int i = __shfl_sync(0xffffffffu, il, k);
      const int64_t p = __shfl_sync(0xffffffffu, pl, k);
      const float* dr = a.d_out + (int64_t)i * a.ldo;
#pragma unroll
      for (int c = 0; c < CPL; ++c)
        if (cv[c]) acc[c] = fmaf(__ldg(a.edge_w + p * H + hc[c]), __ldg(dr + lane + 32 * c), acc[c]);
    }
  }
#pragma unroll
  for (int c = 0; c < CPL; ++c)
    if (cv[c]) a.d_Wh[j * a.ld_dwh + lane + 32 * c] = acc[c];
  for (int o = Hp; o < 32; o <<= 1) dtsum += __shfl_xor_sync(0xffffffffu, dtsum, o);
  if (lane < H) a.d_t[j * H + lane] = dtsum;
}

inline int next_pow2(int x) {
  int p = 1;
  while (p < x) p <<= 1;
  return p;
}

int check_common(int64_t n, int H, int Fp) {
  GNN_REQUIRE(n >= 0 && H > 0 && Fp > 0, GNN_ERR_BAD_ARG, "bad size (n=%lld H=%d Fp=%d)", (long long)n, H, Fp);
  GNN_REQUIRE(H <= 32, GNN_ERR_UNSUPPORTED, "more than 32 heads per call (H=%d): split the heads", H);
  GNN_REQUIRE(H * Fp <= 256, GNN_ERR_UNSUPPORTED, "H*Fp=%d exceeds 256 columns per call: split the heads", H * Fp);
  GNN_REQUIRE(n < 0x7fffffffLL, GNN_ERR_UNSUPPORTED, "n does not fit int32");
  return GNN_OK;
}

int stage_edges(int H) {
  int se = tuning("gat.stage_edges", 128);
  while (se > 32 && se * H > 1024) se >>= 1;
  return se < 32 ? 32 : se;
}

}  // namespace

extern "C" {

int gnn_gat_scores_f32(const float* Wh, int64_t ldw, const float* a_src, const float* a_dst, int64_t n, int32_t H,
                       int32_t Fp, float* s, float* t, gnn_stream_t stream) {
  int rc = check_common(n, H, Fp);
  if (rc != GNN_OK) return rc;
  if (n == 0) return GNN_OK;
  GNN_REQUIRE(Wh && a_src && a_dst && s && t, GNN_ERR_BAD_ARG, "null pointer");
  GNN_REQUIRE(ldw >= (int64_t)H * Fp, GNN_ERR_BAD_ARG, "ldw smaller than H*Fp");
  int64_t grid = (n * H + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 16;
  grid = grid > cap ? cap : grid;
  gat_scores_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(Wh, ldw, a_src, a_dst, n, H, Fp, s, t);
  GNN_LAUNCH_CHECK();
  return GNN_OK;
}

int gnn_gat_fused_fwd_f32(const int64_t* rowptr, const int32_t* col, const float* Wh, int64_t ldw, const float* s,
                          const float* t, int64_t n, int32_t H, int32_t Fp, float alpha, int mode, int apply_elu,
                          const float* col_mean, const float* edge_keep, float* out, int64_t ldo, float* row_max,
                          float* row_sum, gnn_stream_t stream) {
  int rc = check_common(n, H, Fp);
  if (rc != GNN_OK) return rc;
  if (n == 0) return GNN_OK;
  GNN_REQUIRE(rowptr && Wh && s && t && out, GNN_ERR_BAD_ARG, "null pointer");
  GNN_REQUIRE(mode == GNN_GAT_SOFTMAX || mode == GNN_GAT_EXPNEG, GNN_ERR_BAD_ARG, "unknown mode %d", mode);
  GNN_REQUIRE(apply_elu >= 0 && apply_elu <= 2, GNN_ERR_BAD_ARG, "apply_elu must be 0, 1 or 2");
  const int HF = H * Fp;
  GNN_REQUIRE(ldw >= HF && ldo >= HF, GNN_ERR_BAD_ARG, "leading dimension smaller than H*Fp");
  GatArgs a{};
  a.rowptr = rowptr;
  a.col = col;
  a.Wh = Wh;
  a.ldw = ldw;
  a.s = s;
  a.t = t;
  a.n = n;
  a.H = H;
  a.Hp = next_pow2(H);
  a.Fp = Fp;
  a.HF = HF;
  a.alpha = alpha;
  a.mode = mode;
  a.elu = apply_elu;
  a.col_mean = col_mean;
  a.keep = edge_keep;
  a.out = out;
  a.ldo = ldo;
  a.row_max = row_max;
  a.row_sum = row_sum;
  a.SE = stage_edges(H);
  const size_t smem = (size_t)kGatWarps * (a.SE * H + a.SE + 64) * sizeof(float);
  const unsigned grid = (unsigned)((n + kGatWarps - 1) / kGatWarps);
  cudaStream_t st = (cudaStream_t)stream;
  const int cpl = (HF + 31) / 32;
  if (cpl <= 1) gat_fwd_kernel<1><<<grid, kGatWarps * 32, smem, st>>>(a);
  else if (cpl <= 2) gat_fwd_kernel<2><<<grid, kGatWarps * 32, smem, st>>>(a);
  else if (cpl <= 4) gat_fwd_kernel<4><<<grid, kGatWarps * 32, smem, st>>>(a);
  else gat_fwd_kernel<8><<<grid, kGatWarps * 32, smem, st>>>(a);
  GNN_LAUNCH_CHECK();
  return GNN_OK;
}

int gnn_gat_fused_bwd_f32(const int64_t* rowptr, const int32_t* col, const int64_t* rowptr_t, const int32_t* col_t,
                          const int64_t* perm_t, const float* Wh, int64_t ldw, const float* s, const float* t,
                          const float* row_max, const float* row_sum, const float* out_pre, const float* d_out,
                          int64_t ldo, int64_t n, int32_t H, int32_t Fp, float alpha, int mode,
                          const float* edge_keep, float* d_Wh, int64_t ld_dwh, float* d_s, float* d_t,
                          float* d_rowdot, float* edge_scratch, int64_t nnz, gnn_stream_t stream) {
  int rc = check_common(n, H, Fp);
  if (rc != GNN_OK) return rc;
  if (n == 0) return GNN_OK;
  GNN_REQUIRE(rowptr && rowptr_t && Wh && s && t && row_max && row_sum && out_pre && d_out && d_Wh && d_s && d_t &&
                  d_rowdot,
              GNN_ERR_BAD_ARG, "null pointer");
  GNN_REQUIRE(nnz >= 0 && (nnz == 0 || (col && col_t && perm_t && edge_scratch)), GNN_ERR_BAD_ARG,
              "null edge pointer (col/col_t/perm_t/edge_scratch)");
  GNN_REQUIRE(mode == GNN_GAT_SOFTMAX || mode == GNN_GAT_EXPNEG, GNN_ERR_BAD_ARG, "unknown mode %d", mode);
  const int HF = H * Fp;
  GNN_REQUIRE(ldw >= HF && ldo >= HF && ld_dwh >= HF, GNN_ERR_BAD_ARG, "leading dimension smaller than H*Fp");
  cudaStream_t st = (cudaStream_t)stream;
  GatArgs a{};
  a.rowptr = rowptr;
  a.col = col;
  a.Wh = Wh;
  a.ldw = ldw;
  a.s = s;
  a.t = t;
  a.n = n;
  a.H = H;
  a.Hp = next_pow2(H);
  a.Fp = Fp;
  a.HF = HF;
  a.alpha = alpha;
  a.mode = mode;
  a.keep = edge_keep;
  a.ldo = ldo;
  a.row_max = const_cast<float*>(row_max);
  a.row_sum = const_cast<float*>(row_sum);
  a.SE = stage_edges(H);
  a.out_pre = out_pre;
  a.d_out = d_out;
  a.rowdot = d_rowdot;
  a.d_Wh = d_Wh;
  a.ld_dwh = ld_dwh;
  a.d_s = d_s;
  a.d_t = d_t;
  {
    int64_t grid = (n * H + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 16;
    grid = grid > cap ? cap : grid;
    gat_rowdot_kernel<<<(unsigned)grid, 256, 0, st>>>(d_out, out_pre, ldo, n, H, Fp, d_rowdot);
    GNN_LAUNCH_CHECK();
  }
  a.edge_w = edge_scratch;
  a.edge_dz = edge_scratch + nnz * H;
  const unsigned grid = (unsigned)((n + kGatWarps - 1) / kGatWarps);
  {
    a.rowptr = rowptr;
    a.col = col;
    const size_t smem = (size_t)kGatWarps * (a.SE * H + HF + 128) * sizeof(float);
    gat_bwd_rows_kernel<<<grid, kGatWarps * 32, smem, st>>>(a);
    GNN_LAUNCH_CHECK();
  }
  {
    a.rowptr = rowptr_t;
    a.col = col_t;
    a.perm = perm_t;
    const int cpl = (HF + 31) / 32;
    if (cpl <= 1) gat_bwd_cols_kernel<1><<<grid, kGatWarps * 32, 0, st>>>(a);
    else if (cpl <= 2) gat_bwd_cols_kernel<2><<<grid, kGatWarps * 32, 0, st>>>(a);
    else if (cpl <= 4) gat_bwd_cols_kernel<4><<<grid, kGatWarps * 32, 0, st>>>(a);
    else gat_bwd_cols_kernel<8><<<grid, kGatWarps * 32, 0, st>>>(a);
    GNN_LAUNCH_CHECK();
  }
  return GNN_OK;
}

}  // extern "C"
