// Fused multi-head graph attention aggregation (GAT / HAN node-level attention).
// Replaces, at /root/reference:
//   GAT/models/layers.py:25-32  (== HAN/models/NodeAttention.py:25-33)
//       the [N,N,2F'] pair tensor, masked softmax and dense attention·Wh matmul
//   GAT/models/layers.py:108-122  the edge-list variant exp(-LeakyReLU) / rowsum
//   GAT/models/layers.py:55-64    SpecialSpmmFunction.backward (dense N×N edge gradient)
//
// Forward, all H heads in one pass over a CSR row, staged through shared memory:
//   A  lane-per-edge:  logit[k,h] = ±LeakyReLU(s[i,h] + t[col_k,h])
//   B  per-head chunk max by warp shuffle, running max / rescale (softmax mode)
//   B2 logit -> p = exp(logit - max) in place: ONE exp per (edge, head)
//   C  lane-per-column: acc += p[k,head(col)] * Wh[col_k, column], 8 coalesced row gathers
//      in flight per lane
// Scheduling: W = 1 -> one warp per row (low-degree graphs: Cora); W = 4 -> one CTA per row,
// the warps take alternate 128-edge chunks and their (max, sum, acc) partials are merged in
// warp order (dense metapath / Reddit-like graphs).  Both are deterministic.
// Backward: kernel A (CSR rows) computes the edge gradient dz (SDDMM dOut_i·Wh_j reduced per
// head by segmented warp shuffle), stashes attention weight and dz per edge and reduces d_s;
// kernel B (transposed CSR rows) reduces d_Wh and d_t in source order.  Ordered sums only.
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

// floats of shared memory per warp for the staged chunk
__host__ __device__ inline int gat_warp_floats(int SE, int H, int HF) { return SE * H + SE + 128 + HF; }

template <int CPL, int W>
__global__ void __launch_bounds__(kGatWarps * 32) gat_fwd_kernel(const GatArgs a) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t i = (W == 1) ? (int64_t)blockIdx.x * kGatWarps + warp : (int64_t)blockIdx.x;
  const int wsub = (W == 1) ? 0 : warp;
  if (W == 1 && i >= a.n) return;  // warp-uniform; the W == 1 schedule has no block-wide barrier
  const int H = a.H, Hp = a.Hp, SE = a.SE;
  const int PW = gat_warp_floats(SE, H, a.HF);
  float* logit = sm + (size_t)warp * PW;
  int* cols = reinterpret_cast<int*>(logit + SE * H);
  float* mh = reinterpret_cast<float*>(cols + SE);
  float* sc = mh + 32;
  float* marr = sm + (size_t)kGatWarps * PW;      // [W][32]
  float* larr = marr + kGatWarps * 32;            // [W][CPL*32]
  float* aarr = larr + kGatWarps * CPL * 32;      // [W][CPL*32]

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
    if (wsub == 0) {
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
    }
    return;  // uniform over the CTA when W > 1 (one row per CTA)
  }
  if (lane < Hp) {
    mh[lane] = (a.mode == GNN_GAT_SOFTMAX) ? -INFINITY : 0.f;
    sc[lane] = 1.f;
  }
  __syncwarp();
  const int ngrp = 32 / Hp;
  const int hsub = lane % Hp, g = lane / Hp;

  for (int64_t c0 = (int64_t)wsub * SE; c0 < d; c0 += (int64_t)W * SE) {
    const int ne = (int)((d - c0) < SE ? (d - c0) : SE);
    // A: logits, one lane per edge
    for (int k = lane; k < ne; k += 32) {
      const int j = __ldg(a.col + e0 + c0 + k);
      cols[k] = j;
      const float* tj = a.t + (int64_t)j * H;
      const float* si = a.s + i * H;
#pragma unroll 4
      for (int h = 0; h < H; ++h) {
        const float z = __ldg(si + h) + __ldg(tj + h);
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
    // C: weighted accumulation, one lane per output column, 8 row gathers in flight
    for (int k0 = 0; k0 < ne; k0 += 8) {
      float x[8][CPL];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const bool ok = k0 + u < ne;
        const float* wr = a.Wh + (int64_t)(ok ? cols[k0 + u] : 0) * a.ldw;
#pragma unroll
        for (int c = 0; c < CPL; ++c) x[u][c] = (ok && cv[c]) ? __ldg(wr + lane + 32 * c) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (k0 + u < ne) {
#pragma unroll
          for (int c = 0; c < CPL; ++c)
            if (cv[c]) {
              const float p = logit[(k0 + u) * H + hc[c]];
              l[c] += p;
              const float w = a.keep ? p * __ldg(a.keep + (e0 + c0 + k0 + u) * H + hc[c]) : p;
              acc[c] = fmaf(w, x[u][c], acc[c]);
            }
        }
      }
    }
    __syncwarp();
  }

  if (W == 1) {
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
    return;
  }
  // merge the warps' partial (max, sum, acc) in warp order
  if (lane < Hp) marr[warp * 32 + lane] = mh[lane];
#pragma unroll
  for (int c = 0; c < CPL; ++c) {
    larr[(warp * CPL + c) * 32 + lane] = l[c];
    aarr[(warp * CPL + c) * 32 + lane] = acc[c];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int c = 0; c < CPL; ++c)
      if (cv[c]) {
        const int ci = lane + 32 * c;
        float M = -INFINITY;
        for (int w = 0; w < kGatWarps; ++w) M = fmaxf(M, marr[w * 32 + hc[c]]);
        float L = 0.f, A = 0.f;
        for (int w = 0; w < kGatWarps; ++w) {
          const float mw = marr[w * 32 + hc[c]];
          const float f = (mw == -INFINITY) ? 0.f : expf(mw - M);
          L = fmaf(larr[(w * CPL + c) * 32 + lane], f, L);
          A = fmaf(aarr[(w * CPL + c) * 32 + lane], f, A);
        }
        orow[ci] = act_elu(A / L, a.elu);
        if (ci % a.Fp == 0) {
          if (a.row_max) a.row_max[i * H + hc[c]] = M;
          if (a.row_sum) a.row_sum[i * H + hc[c]] = L;
        }
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

// Backward A (forward CSR rows): per-edge attention weight w = keep*alpha_ij, edge gradient dz,
// and d_s[i,h] = sum_j dz.  SEG == true: Fp is a power of two <= 32 (the head groups of the
// lane->column map are aligned lane groups, reduced by segmented shuffle) or H == 1 (full-warp
// reduce).  SEG == false: generic lane-per-edge dot (any H, Fp).
template <int CPL, int W, bool SEG>
__global__ void __launch_bounds__(kGatWarps * 32) gat_bwd_rows_kernel(const GatArgs a) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t i = (W == 1) ? (int64_t)blockIdx.x * kGatWarps + warp : (int64_t)blockIdx.x;
  const int wsub = (W == 1) ? 0 : warp;
  if (W == 1 && i >= a.n) return;
  const int H = a.H, Hp = a.Hp, SE = a.SE, HF = a.HF, Fp = a.Fp;
  const int PW = gat_warp_floats(SE, H, HF);
  float* dot_s = sm + (size_t)warp * PW;       // [SE*H]: head dots, then dz in place
  int* cols = reinterpret_cast<int*>(dot_s + SE * H);
  float* cst = reinterpret_cast<float*>(cols + SE);  // [4][32]: s_i, m_i, 1/l_i, D_i
  float* dout_s = cst + 128;                    // [HF] (generic path)
  float* dsarr = sm + (size_t)kGatWarps * PW;   // [W][32]

  int hc[CPL];
  bool cv[CPL];
  float dcol[CPL];
#pragma unroll
  for (int c = 0; c < CPL; ++c) {
    const int ci = lane + 32 * c;
    cv[c] = ci < HF;
    hc[c] = cv[c] ? ci / Fp : 0;
    dcol[c] = cv[c] ? __ldg(a.d_out + i * a.ldo + ci) : 0.f;
    if (!SEG && cv[c]) dout_s[ci] = dcol[c];
  }
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
  for (int64_t c0 = (int64_t)wsub * SE; c0 < d; c0 += (int64_t)W * SE) {
    const int ne = (int)((d - c0) < SE ? (d - c0) : SE);
    for (int k = lane; k < ne; k += 32) cols[k] = __ldg(a.col + e0 + c0 + k);
    __syncwarp();
    // phase 1: per-head dots dOut_i . Wh_j
    if (SEG) {
      for (int k0 = 0; k0 < ne; k0 += 4) {
        float x[4][CPL];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const bool ok = k0 + u < ne;
          const float* wr = a.Wh + (int64_t)(ok ? cols[k0 + u] : 0) * a.ldw;
#pragma unroll
          for (int c = 0; c < CPL; ++c) x[u][c] = (ok && cv[c]) ? __ldg(wr + lane + 32 * c) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (k0 + u < ne) {  // uniform
            if (H == 1) {
              float pr = 0.f;
#pragma unroll
              for (int c = 0; c < CPL; ++c) pr = fmaf(dcol[c], x[u][c], pr);
              for (int o = 16; o > 0; o >>= 1) pr += __shfl_xor_sync(0xffffffffu, pr, o);
              if (lane == 0) dot_s[k0 + u] = pr;
            } else {
#pragma unroll
              for (int c = 0; c < CPL; ++c) {
                float pr = dcol[c] * x[u][c];
                for (int o = 1; o < Fp; o <<= 1) pr += __shfl_xor_sync(0xffffffffu, pr, o);
                if (cv[c] && ((lane + 32 * c) % Fp) == 0) dot_s[(k0 + u) * H + hc[c]] = pr;
              }
            }
          }
        }
      }
    } else {
      for (int k = lane; k < ne; k += 32) {
        const float* wr = a.Wh + (int64_t)cols[k] * a.ldw;
        for (int h = 0; h < H; ++h) {
          float dot = 0.f;
          for (int f = 0; f < Fp; ++f) dot = fmaf(dout_s[h * Fp + f], __ldg(wr + h * Fp + f), dot);
          dot_s[k * H + h] = dot;
        }
      }
    }
    __syncwarp();
    // phase 2: one lane per (edge, head): attention weight and dz, stashed per edge (coalesced)
    for (int idx = lane; idx < ne * H; idx += 32) {
      const int k = idx / H, h = idx - k * H;
      const int64_t e = e0 + c0 + k;
      const float z = cst[h] + __ldg(a.t + (int64_t)cols[k] * H + h);
      float slope = z > 0.f ? 1.f : a.alpha;
      float ee = z * slope;
      if (a.mode == GNN_GAT_EXPNEG) {
        ee = -ee;
        slope = -slope;
      }
      const float al = expf(ee - cst[32 + h]) * cst[64 + h];
      const float kp = a.keep ? __ldg(a.keep + e * H + h) : 1.f;
      const float dz = al * (kp * dot_s[idx] - cst[96 + h]) * slope;
      a.edge_w[e * H + h] = kp * al;
      a.edge_dz[e * H + h] = dz;
      dot_s[idx] = dz;
    }
    __syncwarp();
    if (hsub < H)
      for (int k = g; k < ne; k += ngrp) dsum += dot_s[k * H + hsub];
    __syncwarp();
  }
  for (int o = Hp; o < 32; o <<= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
  if (W == 1) {
    if (lane < H) a.d_s[i * H + lane] = dsum;
    return;
  }
  if (lane < H) dsarr[warp * 32 + lane] = dsum;
  __syncthreads();
  if (warp == 0 && lane < H) {
    float tot = 0.f;
    for (int w = 0; w < kGatWarps; ++w) tot += dsarr[w * 32 + lane];
    a.d_s[i * H + lane] = tot;
  }
}

// Backward B: transposed CSR rows (source node j), sources of the forward edges ascending.
template <int CPL, int W>
__global__ void __launch_bounds__(kGatWarps * 32) gat_bwd_cols_kernel(const GatArgs a) {
  __shared__ float accarr[(W > 1) ? kGatWarps * CPL * 32 : 1];
  __shared__ float dtarr[(W > 1) ? kGatWarps * 32 : 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t j = (W == 1) ? (int64_t)blockIdx.x * kGatWarps + warp : (int64_t)blockIdx.x;
  const int wsub = (W == 1) ? 0 : warp;
  if (W == 1 && j >= a.n) return;
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
  for (int64_t c0 = e0 + (int64_t)wsub * 32; c0 < e1; c0 += (int64_t)W * 32) {
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
    for (int k0 = 0; k0 < ne; k0 += 8) {
      float w[8][CPL], x[8][CPL];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = __shfl_sync(0xffffffffu, il, (k0 + u) & 31);
        const int64_t p = __shfl_sync(0xffffffffu, pl, (k0 + u) & 31);
        const bool ok = k0 + u < ne;
        const float* dr = a.d_out + (int64_t)i * a.ldo;
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
          const bool on = ok && cv[c];
          w[u][c] = on ? __ldg(a.edge_w + p * H + hc[c]) : 0.f;
          x[u][c] = on ? __ldg(dr + lane + 32 * c) : 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int c = 0; c < CPL; ++c) acc[c] = fmaf(w[u][c], x[u][c], acc[c]);
    }
  }
  for (int o = Hp; o < 32; o <<= 1) dtsum += __shfl_xor_sync(0xffffffffu, dtsum, o);
  if (W == 1) {
#pragma unroll
    for (int c = 0; c < CPL; ++c)
      if (cv[c]) a.d_Wh[j * a.ld_dwh + lane + 32 * c] = acc[c];
    if (lane < H) a.d_t[j * H + lane] = dtsum;
    return;
  }
#pragma unroll
  for (int c = 0; c < CPL; ++c) accarr[(warp * CPL + c) * 32 + lane] = acc[c];
  if (lane < H) dtarr[warp * 32 + lane] = dtsum;
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int c = 0; c < CPL; ++c)
      if (cv[c]) {
        float tot = 0.f;
        for (int w = 0; w < kGatWarps; ++w) tot += accarr[(w * CPL + c) * 32 + lane];
        a.d_Wh[j * a.ld_dwh + lane + 32 * c] = tot;
      }
    if (lane < H) {
      float tot = 0.f;
      for (int w = 0; w < kGatWarps; ++w) tot += dtarr[w * 32 + lane];
      a.d_t[j * H + lane] = tot;
    }
  }
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

// one CTA per row when rows are long on average
bool cooperative(int64_t n, int64_t nnz) {
  const int thr = tuning("gat.coop_min_avg_deg", 48);
  return n > 0 && nnz > 0 && nnz / n >= thr;
}

template <int W>
size_t gat_smem_bytes(int SE, int H, int HF, int cpl) {
  size_t f = (size_t)kGatWarps * gat_warp_floats(SE, H, HF);
  if (W > 1) f += (size_t)kGatWarps * 32 + 2 * (size_t)kGatWarps * cpl * 32;
  return f * sizeof(float);
}

#define GNN_GAT_CPL_DISPATCH(KERNEL, ...)               \
  do {                                                  \
    if (cpl <= 1) KERNEL<1, __VA_ARGS__;                \
    else if (cpl <= 2) KERNEL<2, __VA_ARGS__;           \
    else if (cpl <= 4) KERNEL<4, __VA_ARGS__;           \
    else KERNEL<8, __VA_ARGS__;                         \
  } while (0)

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
                          const float* t, int64_t n, int64_t nnz, int32_t H, int32_t Fp, float alpha, int mode,
                          int apply_elu, const float* col_mean, const float* edge_keep, float* out, int64_t ldo,
                          float* row_max, float* row_sum, gnn_stream_t stream) {
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
  cudaStream_t st = (cudaStream_t)stream;
  const int cpl = (HF + 31) / 32;
  const int thr = kGatWarps * 32;
  if (cooperative(n, nnz)) {
    const size_t smem = gat_smem_bytes<kGatWarps>(a.SE, H, HF, cpl <= 1 ? 1 : cpl <= 2 ? 2 : cpl <= 4 ? 4 : 8);
    const unsigned grid = (unsigned)n;
    GNN_GAT_CPL_DISPATCH(gat_fwd_kernel, kGatWarps><<<grid, thr, smem, st>>>(a));
  } else {
    const size_t smem = gat_smem_bytes<1>(a.SE, H, HF, 0);
    const unsigned grid = (unsigned)((n + kGatWarps - 1) / kGatWarps);
    GNN_GAT_CPL_DISPATCH(gat_fwd_kernel, 1><<<grid, thr, smem, st>>>(a));
  }
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
  a.edge_w = edge_scratch;
  a.edge_dz = edge_scratch + nnz * H;
  {
    int64_t grid = (n * H + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 16;
    grid = grid > cap ? cap : grid;
    gat_rowdot_kernel<<<(unsigned)grid, 256, 0, st>>>(d_out, out_pre, ldo, n, H, Fp, d_rowdot);
    GNN_LAUNCH_CHECK();
  }
  const int cpl = (HF + 31) / 32;
  const int cplr = cpl <= 1 ? 1 : cpl <= 2 ? 2 : cpl <= 4 ? 4 : 8;
  const int thr = kGatWarps * 32;
  const bool coop = cooperative(n, nnz);
  const bool seg = (H == 1) || (Fp <= 32 && (Fp & (Fp - 1)) == 0);
  const unsigned grid1 = (unsigned)((n + kGatWarps - 1) / kGatWarps);
  {
    a.rowptr = rowptr;
    a.col = col;
    if (coop) {
      const size_t smem = gat_smem_bytes<kGatWarps>(a.SE, H, HF, cplr);
      if (seg) GNN_GAT_CPL_DISPATCH(gat_bwd_rows_kernel, kGatWarps, true><<<(unsigned)n, thr, smem, st>>>(a));
      else GNN_GAT_CPL_DISPATCH(gat_bwd_rows_kernel, kGatWarps, false><<<(unsigned)n, thr, smem, st>>>(a));
    } else {
      const size_t smem = gat_smem_bytes<1>(a.SE, H, HF, 0);
      if (seg) GNN_GAT_CPL_DISPATCH(gat_bwd_rows_kernel, 1, true><<<grid1, thr, smem, st>>>(a));
      else GNN_GAT_CPL_DISPATCH(gat_bwd_rows_kernel, 1, false><<<grid1, thr, smem, st>>>(a));
    }
    GNN_LAUNCH_CHECK();
  }
  {
    a.rowptr = rowptr_t;
    a.col = col_t;
    a.perm = perm_t;
    if (coop) GNN_GAT_CPL_DISPATCH(gat_bwd_cols_kernel, kGatWarps><<<(unsigned)n, thr, 0, st>>>(a));
    else GNN_GAT_CPL_DISPATCH(gat_bwd_cols_kernel, 1><<<grid1, thr, 0, st>>>(a));
    GNN_LAUNCH_CHECK();
  }
  return GNN_OK;
}

}  // extern "C"
