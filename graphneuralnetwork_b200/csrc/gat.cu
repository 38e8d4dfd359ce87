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
  const int64_t* perm;   // backward B: transposed slot -> forward edge slot (nullptr: the stash is in transposed order)
  const int64_t* tslot;  // backward A: forward edge slot -> transposed slot (nullptr: stash in forward order)
  const void* Wh;        // T [n, ldw]: fp32 or bf16 (softmax and accumulation are fp32 either way)
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
  void* out;             // T [n, ldo]
  int64_t ldo;
  float* row_max;
  float* row_sum;
  int SE;
  // backward
  const void* out_pre;   // T
  const void* d_out;     // T
  const float* rowdot;
  float* edge_st;        // [nnz][2][H] fp32: per edge (keep*alpha, dz), in transposed slot order when tslot != nullptr
  void* d_Wh;            // T
  int64_t ld_dwh;
  float* d_s;
  float* d_t;
  // schedule: the warp-per-row launch leaves rows longer than skip_deg_gt to the CTA-per-row
  // launch, which walks row_list (nullable: row = blockIdx.x)
  const int64_t* row_list;
  int64_t skip_deg_gt;
};

__device__ __forceinline__ bool aligned_to_dev(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

template <typename T>
__device__ __forceinline__ float ldv(const T* p);
template <>
__device__ __forceinline__ float ldv<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float ldv<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __uint_as_float(((uint32_t)__ldg(reinterpret_cast<const unsigned short*>(p))) << 16);
}
__device__ __forceinline__ void stv(float* p, float v) { *p = v; }
__device__ __forceinline__ void stv(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

__device__ __forceinline__ float act_elu(float x, int elu) {
  if (elu >= 1) x = elu1(x);
  if (elu >= 2) x = elu1(x);
  return x;
}

// floats of shared memory per warp for the staged chunk
__host__ __device__ inline int gat_warp_floats(int SE, int H, int HF) { return SE * H + SE + 160 + HF; }
// backward A additionally stages the int64 transposed slot of every edge of the chunk (first, 8-byte aligned)
__host__ __device__ inline int gat_warp_floats_bwd(int SE, int H, int HF) {
  return 2 * SE + SE * H + SE + 160 + ((HF + 1) & ~1);
}

// W  = warps that share one row: 1 -> one warp per row (4 rows per CTA); 4 / 16 -> one CTA of W
//      warps per row, the warps take alternate SE-edge chunks and merge (max, sum, acc) in warp order.
// HT = compile-time head count (1, 8; must divide 32) or 0 for the generic run-time H.  With HT known
//      the per-lane head of phase B/B2 is fixed (lane % HT), t rows are loaded as vectors, the row sum
//      is reduced once per chunk instead of once per edge, and all index arithmetic folds away
//      (the generic path executed 79 warp instructions per edge, ncu profiles/r01_prof_gat_fwd_raw.csv).
template <int W>
struct GatBlock {
  static constexpr int kWarps = (W == 1) ? kGatWarps : W;
};

template <typename T, int CPL, int W, int HT>
__global__ void __launch_bounds__(GatBlock<W>::kWarps * 32) gat_fwd_kernel(const GatArgs a) {
  constexpr int BW = GatBlock<W>::kWarps;
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t i = (W == 1) ? (int64_t)blockIdx.x * kGatWarps + warp
                             : (a.row_list ? __ldg(a.row_list + blockIdx.x) : (int64_t)blockIdx.x);
  const int wsub = (W == 1) ? 0 : warp;
  if (W == 1 && i >= a.n) return;  // warp-uniform; the W == 1 schedule has no block-wide barrier
  const int H = HT ? HT : a.H, Hp = HT ? HT : a.Hp, SE = a.SE;
  const int PW = gat_warp_floats(SE, H, a.HF);
  float* logit = sm + (size_t)warp * PW;
  int* cols = reinterpret_cast<int*>(logit + SE * H);
  float* mh = reinterpret_cast<float*>(cols + SE);  // [32] running max per head
  float* sc = mh + 32;                              // [32] rescale factor of this chunk
  float* lh = sc + 32;                              // [32] running sum per head (HT path)
  float* marr = sm + (size_t)BW * PW;               // [BW][32]
  float* larr = marr + BW * 32;                     // [BW][CPL*32]
  float* aarr = larr + BW * CPL * 32;               // [BW][CPL*32]

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
  if (W == 1 && a.skip_deg_gt > 0 && d > a.skip_deg_gt) return;  // a long row: the CTA-per-row launch owns it
  T* orow = reinterpret_cast<T*>(a.out) + i * a.ldo;
  if (d == 0) {
    // GAT/models/layers.py:28-30: an all -9e15 row soft-maxes to the uniform 1/N over ALL nodes
    if (wsub == 0) {
#pragma unroll
      for (int c = 0; c < CPL; ++c)
        if (cv[c]) {
          const int ci = lane + 32 * c;
          stv(orow + ci, act_elu(a.col_mean ? a.col_mean[ci] : 0.f, a.elu));
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
    lh[lane] = 0.f;
  }
  __syncwarp();
  const int ngrp = 32 / Hp;
  const int hsub = lane % Hp, g = lane / Hp;
  float si_reg[HT ? HT : 1];
  if (HT) {
#pragma unroll
    for (int h = 0; h < (HT ? HT : 1); ++h) si_reg[h] = __ldg(a.s + i * H + h);
  }
  const bool t_vec4 = HT && (HT % 4 == 0) && aligned_to_dev(a.t, 16);

  for (int64_t c0 = (int64_t)wsub * SE; c0 < d; c0 += (int64_t)W * SE) {
    const int ne = (int)((d - c0) < SE ? (d - c0) : SE);
    // A: logits, one lane per edge
    for (int k = lane; k < ne; k += 32) {
      const int j = __ldg(a.col + e0 + c0 + k);
      cols[k] = j;
      const float* tj = a.t + (int64_t)j * H;
      if (HT) {
        float tv[HT ? HT : 1];
        if (t_vec4) {
#pragma unroll
          for (int q = 0; q < (HT ? HT : 1) / 4; ++q) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(tj) + q);
            tv[4 * q] = v.x; tv[4 * q + 1] = v.y; tv[4 * q + 2] = v.z; tv[4 * q + 3] = v.w;
          }
        } else {
#pragma unroll
          for (int h = 0; h < (HT ? HT : 1); ++h) tv[h] = __ldg(tj + h);
        }
#pragma unroll
        for (int h = 0; h < (HT ? HT : 1); ++h) {
          const float z = si_reg[h] + tv[h];
          float e = fmaxf(z, 0.f) + a.alpha * fminf(z, 0.f);  // LeakyReLU without a branch
          if (a.mode == GNN_GAT_EXPNEG) e = -e;
          logit[k * H + h] = e;
        }
      } else {
        const float* si = a.s + i * H;
#pragma unroll 4
        for (int h = 0; h < H; ++h) {
          const float z = __ldg(si + h) + __ldg(tj + h);
          float e = z > 0.f ? z : a.alpha * z;
          if (a.mode == GNN_GAT_EXPNEG) e = -e;
          logit[k * H + h] = e;
        }
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
        const float f = (mo == -INFINITY) ? 0.f : __expf(mo - mn);
        sc[lane] = f;
        mh[lane] = mn;
        lh[lane] *= f;
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
    if (HT) {
      // idx = lane + 32*j  =>  head = lane % HT for every j (HT divides 32)
      const float mymax = mh[hsub];
      float ps = 0.f;
      for (int idx = lane; idx < ne * H; idx += 32) {
        const float p = __expf(logit[idx] - mymax);
        logit[idx] = p;
        ps += p;
      }
      for (int o = Hp; o < 32; o <<= 1) ps += __shfl_xor_sync(0xffffffffu, ps, o);
      if (lane < H) lh[lane] += ps;
    } else {
      for (int idx = lane; idx < ne * H; idx += 32) logit[idx] = __expf(logit[idx] - mh[idx % H]);
    }
    __syncwarp();
    // C: weighted accumulation, one lane per output column, 8 row gathers in flight
    for (int k0 = 0; k0 < ne; k0 += 8) {
      float x[8][CPL];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const bool ok = k0 + u < ne;
        const T* wr = reinterpret_cast<const T*>(a.Wh) + (int64_t)(ok ? cols[k0 + u] : 0) * a.ldw;
#pragma unroll
        for (int c = 0; c < CPL; ++c) x[u][c] = (ok && cv[c]) ? ldv<T>(wr + lane + 32 * c) : 0.f;
      }
      const float* pk = logit + k0 * H;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (k0 + u < ne) {
#pragma unroll
          for (int c = 0; c < CPL; ++c)
            if (cv[c]) {
              const float p = pk[u * H + hc[c]];
              if (!HT) l[c] += p;
              const float w = a.keep ? p * __ldg(a.keep + (e0 + c0 + k0 + u) * H + hc[c]) : p;
              acc[c] = fmaf(w, x[u][c], acc[c]);
            }
        }
      }
    }
    __syncwarp();
  }
  if (HT) {
#pragma unroll
    for (int c = 0; c < CPL; ++c) l[c] = lh[hc[c]];
  }

  if (W == 1) {
#pragma unroll
    for (int c = 0; c < CPL; ++c)
      if (cv[c]) {
        const int ci = lane + 32 * c;
        stv(orow + ci, act_elu(acc[c] / l[c], a.elu));
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
        for (int w = 0; w < BW; ++w) M = fmaxf(M, marr[w * 32 + hc[c]]);
        float L = 0.f, A = 0.f;
        for (int w = 0; w < BW; ++w) {
          const float mw = marr[w * 32 + hc[c]];
          const float f = (mw == -INFINITY) ? 0.f : __expf(mw - M);
          L = fmaf(larr[(w * CPL + c) * 32 + lane], f, L);
          A = fmaf(aarr[(w * CPL + c) * 32 + lane], f, A);
        }
        stv(orow + ci, act_elu(A / L, a.elu));
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

template <typename T>
__global__ void __launch_bounds__(256) gat_rowdot_kernel(const T* __restrict__ d_out,
                                                         const T* __restrict__ out_pre, int64_t ldo, int64_t n,
                                                         int H, int Fp, float* __restrict__ rowdot) {
  const int64_t total = n * H;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
    const int h = (int)(p % H);
    const int64_t i = p / H;
    const T* a = d_out + i * ldo + h * Fp;
    const T* b = out_pre + i * ldo + h * Fp;
    float acc = 0.f;
    for (int f = 0; f < Fp; ++f) acc = fmaf(ldv<T>(a + f), ldv<T>(b + f), acc);
    rowdot[p] = acc;
  }
}

// Phase 1 of backward A for one staged chunk: dots dOut_i . Wh_j per head.  FPT = head width
// (power of two <= 32: aligned lane groups, segmented xor-shuffle) or 0 (single head: full warp).
template <typename T, int CPL, int FPT>
__device__ __forceinline__ void bwd_head_dots(const GatArgs& a, const int* cols, int ne, const float (&dcol)[CPL],
                                              const bool (&cv)[CPL], const bool (&lead)[CPL], const int (&hc)[CPL],
                                              float* dot_s, int H, int lane) {
  for (int k0 = 0; k0 < ne; k0 += 4) {
    float x[4][CPL];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const bool ok = k0 + u < ne;
      const T* wr = reinterpret_cast<const T*>(a.Wh) + (int64_t)(ok ? cols[k0 + u] : 0) * a.ldw;
#pragma unroll
      for (int c = 0; c < CPL; ++c) x[u][c] = (ok && cv[c]) ? ldv<T>(wr + lane + 32 * c) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (k0 + u < ne) {  // uniform
        if (FPT == 0) {
          float pr = 0.f;
#pragma unroll
          for (int c = 0; c < CPL; ++c) pr = fmaf(dcol[c], x[u][c], pr);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) pr += __shfl_xor_sync(0xffffffffu, pr, o);
          if (lane == 0) dot_s[k0 + u] = pr;
        } else {
#pragma unroll
          for (int c = 0; c < CPL; ++c) {
            float pr = dcol[c] * x[u][c];
#pragma unroll
            for (int o = 1; o < FPT; o <<= 1) pr += __shfl_xor_sync(0xffffffffu, pr, o);
            if (lead[c]) dot_s[(k0 + u) * H + hc[c]] = pr;
          }
        }
      }
    }
  }
}

// Backward A (forward CSR rows): per-edge attention weight w = keep*alpha_ij, edge gradient dz,
// and d_s[i,h] = sum_j dz.  SEG == true: Fp is a power of two <= 32 (the head groups of the
// lane->column map are aligned lane groups, reduced by segmented shuffle) or H == 1 (full-warp
// reduce).  SEG == false: generic lane-per-edge dot (any H, Fp).
template <typename T, int CPL, int W, bool SEG, int HT>
__global__ void __launch_bounds__(GatBlock<W>::kWarps * 32) gat_bwd_rows_kernel(const GatArgs a) {
  constexpr int BW = GatBlock<W>::kWarps;
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t i = (W == 1) ? (int64_t)blockIdx.x * kGatWarps + warp
                             : (a.row_list ? __ldg(a.row_list + blockIdx.x) : (int64_t)blockIdx.x);
  const int wsub = (W == 1) ? 0 : warp;
  if (W == 1 && i >= a.n) return;
  if (W == 1 && a.skip_deg_gt > 0 && __ldg(a.rowptr + i + 1) - __ldg(a.rowptr + i) > a.skip_deg_gt) return;
  const int H = HT ? HT : a.H, Hp = HT ? HT : a.Hp, SE = a.SE, HF = a.HF, Fp = a.Fp;
  const int PW = gat_warp_floats_bwd(SE, H, HF);
  int64_t* tsl = reinterpret_cast<int64_t*>(sm + (size_t)warp * PW);  // [SE]: stash slot of every staged edge
  float* dot_s = sm + (size_t)warp * PW + 2 * SE;                     // [SE*H]: head dots, then dz in place
  int* cols = reinterpret_cast<int*>(dot_s + SE * H);
  float* cst = reinterpret_cast<float*>(cols + SE);  // [4][32]: s_i, m_i, 1/l_i, D_i
  float* dout_s = cst + 128;                    // [HF] (generic path)
  float* dsarr = sm + (size_t)BW * PW;          // [BW][32]

  int hc[CPL];
  bool cv[CPL], lead[CPL];
  float dcol[CPL];
#pragma unroll
  for (int c = 0; c < CPL; ++c) {
    const int ci = lane + 32 * c;
    cv[c] = ci < HF;
    hc[c] = cv[c] ? ci / Fp : 0;
    lead[c] = cv[c] && (ci % Fp) == 0;
    dcol[c] = cv[c] ? ldv<T>(reinterpret_cast<const T*>(a.d_out) + i * a.ldo + ci) : 0.f;
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
  (void)e1;
  for (int64_t c0 = (int64_t)wsub * SE; c0 < d; c0 += (int64_t)W * SE) {
    const int ne = (int)((d - c0) < SE ? (d - c0) : SE);
    for (int k = lane; k < ne; k += 32) {
      cols[k] = __ldg(a.col + e0 + c0 + k);
      tsl[k] = a.tslot ? __ldg(a.tslot + e0 + c0 + k) : (e0 + c0 + k);
    }
    __syncwarp();
    // phase 1: per-head dots dOut_i . Wh_j
    if (SEG) {
      // the head-group reduction is unrolled for the common widths; `lead` marks the lane that
      // owns a head's first column (no run-time modulo or loop in the per-edge path)
      switch (H == 1 ? 0 : Fp) {
        case 0: bwd_head_dots<T, CPL, 0>(a, cols, ne, dcol, cv, lead, hc, dot_s, H, lane); break;
        case 1: bwd_head_dots<T, CPL, 1>(a, cols, ne, dcol, cv, lead, hc, dot_s, H, lane); break;
        case 2: bwd_head_dots<T, CPL, 2>(a, cols, ne, dcol, cv, lead, hc, dot_s, H, lane); break;
        case 4: bwd_head_dots<T, CPL, 4>(a, cols, ne, dcol, cv, lead, hc, dot_s, H, lane); break;
        case 8: bwd_head_dots<T, CPL, 8>(a, cols, ne, dcol, cv, lead, hc, dot_s, H, lane); break;
        case 16: bwd_head_dots<T, CPL, 16>(a, cols, ne, dcol, cv, lead, hc, dot_s, H, lane); break;
        default: bwd_head_dots<T, CPL, 32>(a, cols, ne, dcol, cv, lead, hc, dot_s, H, lane); break;
      }
    } else {
      for (int k = lane; k < ne; k += 32) {
        const T* wr = reinterpret_cast<const T*>(a.Wh) + (int64_t)cols[k] * a.ldw;
        for (int h = 0; h < H; ++h) {
          float dot = 0.f;
          for (int f = 0; f < Fp; ++f) dot = fmaf(dout_s[h * Fp + f], ldv<T>(wr + h * Fp + f), dot);
          dot_s[k * H + h] = dot;
        }
      }
    }
    __syncwarp();
    // phase 2: one lane per (edge, head): attention weight and dz, stashed per edge (coalesced)
    for (int idx = lane; idx < ne * H; idx += 32) {
      const int k = idx / H, h = idx - k * H;  // H is a compile-time constant when HT != 0
      const int64_t e = e0 + c0 + k;
      const float z = cst[h] + __ldg(a.t + (int64_t)cols[k] * H + h);
      float slope = z > 0.f ? 1.f : a.alpha;
      float ee = z * slope;
      if (a.mode == GNN_GAT_EXPNEG) {
        ee = -ee;
        slope = -slope;
      }
      const float al = __expf(ee - cst[32 + h]) * cst[64 + h];
      const float kp = a.keep ? __ldg(a.keep + e * H + h) : 1.f;
      const float dz = al * (kp * dot_s[idx] - cst[96 + h]) * slope;
      // the stash is written in the order kernel B walks it (scattered 32-byte stores here, no stall;
      // sequential reads there instead of a random 32-byte read per edge through the permutation)
      float* stp = a.edge_st + tsl[k] * (2 * H);
      stp[h] = kp * al;
      stp[H + h] = dz;
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
    for (int w = 0; w < BW; ++w) tot += dsarr[w * 32 + lane];
    a.d_s[i * H + lane] = tot;
  }
}

// Backward B: transposed CSR rows (source node j), sources of the forward edges ascending.
template <typename T, int CPL, int W>
__global__ void __launch_bounds__(GatBlock<W>::kWarps * 32) gat_bwd_cols_kernel(const GatArgs a) {
  constexpr int BW = GatBlock<W>::kWarps;
  __shared__ float accarr[(W > 1) ? BW * CPL * 32 : 1];
  __shared__ float dtarr[(W > 1) ? BW * 32 : 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t j = (W == 1) ? (int64_t)blockIdx.x * kGatWarps + warp
                             : (a.row_list ? __ldg(a.row_list + blockIdx.x) : (int64_t)blockIdx.x);
  const int wsub = (W == 1) ? 0 : warp;
  if (W == 1 && j >= a.n) return;
  if (W == 1 && a.skip_deg_gt > 0 && __ldg(a.rowptr + j + 1) - __ldg(a.rowptr + j) > a.skip_deg_gt) return;
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
      pl = a.perm ? __ldg(a.perm + c0 + lane) : (c0 + lane);
    }
    for (int k0 = 0; k0 < ne; k0 += ngrp) {
      const int kk = k0 + g;
      const int64_t p = __shfl_sync(0xffffffffu, pl, kk & 31);
      if (kk < ne && hsub < H) dtsum += __ldg(a.edge_st + p * (2 * H) + H + hsub);
    }
    for (int k0 = 0; k0 < ne; k0 += 8) {
      float w[8][CPL], x[8][CPL];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = __shfl_sync(0xffffffffu, il, (k0 + u) & 31);
        const int64_t p = __shfl_sync(0xffffffffu, pl, (k0 + u) & 31);
        const bool ok = k0 + u < ne;
        const T* dr = reinterpret_cast<const T*>(a.d_out) + (int64_t)i * a.ldo;
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
          const bool on = ok && cv[c];
          w[u][c] = on ? __ldg(a.edge_st + p * (2 * H) + hc[c]) : 0.f;
          x[u][c] = on ? ldv<T>(dr + lane + 32 * c) : 0.f;
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
      if (cv[c]) stv(reinterpret_cast<T*>(a.d_Wh) + j * a.ld_dwh + lane + 32 * c, acc[c]);
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
        for (int w = 0; w < BW; ++w) tot += accarr[(w * CPL + c) * 32 + lane];
        stv(reinterpret_cast<T*>(a.d_Wh) + j * a.ld_dwh + lane + 32 * c, tot);
      }
    if (lane < H) {
      float tot = 0.f;
      for (int w = 0; w < BW; ++w) tot += dtarr[w * 32 + lane];
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

// one CTA per row for every row when rows are long on average (dense metapath adjacencies);
// otherwise one warp per row, and only the rows of the caller's long-row list get a CTA
bool cooperative(int64_t n, int64_t nnz) {
  const int thr = tuning("gat.coop_min_avg_deg", 256);
  return n > 0 && nnz > 0 && nnz / n >= thr;
}

template <int W>
size_t gat_smem_bytes(int SE, int H, int HF, int cpl, bool bwd = false) {
  constexpr int BW = GatBlock<W>::kWarps;
  size_t f = (size_t)BW * (bwd ? gat_warp_floats_bwd(SE, H, HF) : gat_warp_floats(SE, H, HF));
  if (W > 1) f += (size_t)BW * 32 + 2 * (size_t)BW * cpl * 32;
  return f * sizeof(float);
}

// forward launch: picks the compile-time head count when there is one
template <typename T, int CPL, int W>
int gat_fwd_launch(const GatArgs& a, unsigned grid, size_t smem, cudaStream_t st) {
  constexpr int thr = GatBlock<W>::kWarps * 32;
#define GNN_GAT_FWD(HT)                                                                                         \
  do {                                                                                                          \
    if (smem > 48 * 1024)                                                                                       \
      GNN_CUDA(cudaFuncSetAttribute(gat_fwd_kernel<T, CPL, W, HT>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                    (int)smem));                                                                \
    gat_fwd_kernel<T, CPL, W, HT><<<grid, thr, smem, st>>>(a);                                                  \
  } while (0)
  if (a.H == 8) GNN_GAT_FWD(8);
  else if (a.H == 1) GNN_GAT_FWD(1);
  else GNN_GAT_FWD(0);
#undef GNN_GAT_FWD
  GNN_LAUNCH_CHECK();
  return GNN_OK;
}

template <typename T, int W>
int gat_fwd_dispatch(const GatArgs& a, int cpl, unsigned grid, cudaStream_t st) {
  const int cplr = cpl <= 1 ? 1 : cpl <= 2 ? 2 : cpl <= 4 ? 4 : 8;
  const size_t smem = gat_smem_bytes<W>(a.SE, a.H, a.HF, cplr);
  switch (cplr) {
    case 1: return gat_fwd_launch<T, 1, W>(a, grid, smem, st);
    case 2: return gat_fwd_launch<T, 2, W>(a, grid, smem, st);
    case 4: return gat_fwd_launch<T, 4, W>(a, grid, smem, st);
    default: return gat_fwd_launch<T, 8, W>(a, grid, smem, st);
  }
}

template <typename T>
int gat_fwd_impl(const int64_t* rowptr, const int32_t* col, const T* Wh, int64_t ldw, const float* s, const float* t,
                 int64_t n, int64_t nnz, int32_t H, int32_t Fp, float alpha, int mode, int apply_elu,
                 const float* col_mean, const float* edge_keep, T* out, int64_t ldo, float* row_max, float* row_sum,
                 const int64_t* long_rows, int64_t n_long, int64_t long_threshold, cudaStream_t st) {
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
  const int cpl = (HF + 31) / 32;
  GNN_REQUIRE(n_long == 0 || (long_rows && long_threshold > 0), GNN_ERR_BAD_ARG, "inconsistent long-row list");
  if (cooperative(n, nnz)) return gat_fwd_dispatch<T, kGatWarps>(a, cpl, (unsigned)n, st);
  if (n_long > 0) {  // hub rows: one 16-warp CTA each, launched first so the short rows fill in behind
    a.row_list = long_rows;
    rc = gat_fwd_dispatch<T, 16>(a, cpl, (unsigned)n_long, st);
    if (rc != GNN_OK) return rc;
    a.row_list = nullptr;
    a.skip_deg_gt = long_threshold;
  }
  return gat_fwd_dispatch<T, 1>(a, cpl, (unsigned)((n + kGatWarps - 1) / kGatWarps), st);
}

// rows kernel launch for one (CPL, W): picks SEG / HT
template <typename T, int CPLV, int WV>
int gat_bwd_rows_launch(const GatArgs& a, bool seg, unsigned grid, cudaStream_t st) {
  const size_t smem = gat_smem_bytes<WV>(a.SE, a.H, a.HF, CPLV, true);
  constexpr int thrv = GatBlock<WV>::kWarps * 32;
#define GNN_ROWS(SEGV, HTV)                                                                                 \
  do {                                                                                                      \
    if (smem > 48 * 1024)                                                                                   \
      GNN_CUDA(cudaFuncSetAttribute(gat_bwd_rows_kernel<T, CPLV, WV, SEGV, HTV>,                            \
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));               \
    gat_bwd_rows_kernel<T, CPLV, WV, SEGV, HTV><<<grid, thrv, smem, st>>>(a);                               \
  } while (0)
  if (!seg) GNN_ROWS(false, 0);
  else if (a.H == 8) GNN_ROWS(true, 8);
  else if (a.H == 1) GNN_ROWS(true, 1);
  else GNN_ROWS(true, 0);
#undef GNN_ROWS
  GNN_LAUNCH_CHECK();
  return GNN_OK;
}

template <typename T, int WV>
int gat_bwd_rows_dispatch(const GatArgs& a, int cplr, bool seg, unsigned grid, cudaStream_t st) {
  switch (cplr) {
    case 1: return gat_bwd_rows_launch<T, 1, WV>(a, seg, grid, st);
    case 2: return gat_bwd_rows_launch<T, 2, WV>(a, seg, grid, st);
    case 4: return gat_bwd_rows_launch<T, 4, WV>(a, seg, grid, st);
    default: return gat_bwd_rows_launch<T, 8, WV>(a, seg, grid, st);
  }
}

template <typename T, int WV>
int gat_bwd_cols_dispatch(const GatArgs& a, int cplr, unsigned grid, cudaStream_t st) {
  constexpr int thrv = GatBlock<WV>::kWarps * 32;
  switch (cplr) {
    case 1: gat_bwd_cols_kernel<T, 1, WV><<<grid, thrv, 0, st>>>(a); break;
    case 2: gat_bwd_cols_kernel<T, 2, WV><<<grid, thrv, 0, st>>>(a); break;
    case 4: gat_bwd_cols_kernel<T, 4, WV><<<grid, thrv, 0, st>>>(a); break;
    default: gat_bwd_cols_kernel<T, 8, WV><<<grid, thrv, 0, st>>>(a); break;
  }
  GNN_LAUNCH_CHECK();
  return GNN_OK;
}

template <typename T>
int gat_bwd_impl(const int64_t* rowptr, const int32_t* col, const int64_t* rowptr_t, const int32_t* col_t,
                 const int64_t* perm_t, const int64_t* edge_to_tslot, const T* Wh, int64_t ldw, const float* s,
                 const float* t, const float* row_max, const float* row_sum, const T* out_pre, const T* d_out,
                 int64_t ldo, int64_t n, int32_t H, int32_t Fp, float alpha, int mode, const float* edge_keep, T* d_Wh,
                 int64_t ld_dwh, float* d_s, float* d_t, float* d_rowdot, float* edge_scratch, int64_t nnz,
                 const int64_t* long_rows, int64_t n_long, const int64_t* long_rows_t, int64_t n_long_t,
                 int64_t long_threshold, cudaStream_t st) {
  int rc = check_common(n, H, Fp);
  if (rc != GNN_OK) return rc;
  if (n == 0) return GNN_OK;
  GNN_REQUIRE(rowptr && rowptr_t && Wh && s && t && row_max && row_sum && out_pre && d_out && d_Wh && d_s && d_t &&
                  d_rowdot,
              GNN_ERR_BAD_ARG, "null pointer");
  GNN_REQUIRE(nnz >= 0 && (nnz == 0 || (col && col_t && (perm_t || edge_to_tslot) && edge_scratch)), GNN_ERR_BAD_ARG,
              "null edge pointer (col/col_t/perm_t|edge_to_tslot/edge_scratch)");
  GNN_REQUIRE(mode == GNN_GAT_SOFTMAX || mode == GNN_GAT_EXPNEG, GNN_ERR_BAD_ARG, "unknown mode %d", mode);
  const int HF = H * Fp;
  GNN_REQUIRE(ldw >= HF && ldo >= HF && ld_dwh >= HF, GNN_ERR_BAD_ARG, "leading dimension smaller than H*Fp");
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
  a.edge_st = edge_scratch;
  a.tslot = edge_to_tslot;
  {
    int64_t grid = (n * H + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 16;
    grid = grid > cap ? cap : grid;
    gat_rowdot_kernel<T><<<(unsigned)grid, 256, 0, st>>>(d_out, out_pre, ldo, n, H, Fp, d_rowdot);
    GNN_LAUNCH_CHECK();
  }
  const int cpl = (HF + 31) / 32;
  const int cplr = cpl <= 1 ? 1 : cpl <= 2 ? 2 : cpl <= 4 ? 4 : 8;
  const bool coop = cooperative(n, nnz);
  const bool seg = (H == 1) || (Fp <= 32 && (Fp & (Fp - 1)) == 0);
  const unsigned grid1 = (unsigned)((n + kGatWarps - 1) / kGatWarps);
  GNN_REQUIRE((n_long == 0 || long_rows) && (n_long_t == 0 || long_rows_t) &&
                  ((n_long == 0 && n_long_t == 0) || long_threshold > 0),
              GNN_ERR_BAD_ARG, "inconsistent long-row lists");
  // kernel A over the forward CSR.  W: 1 warp per row, 4 = every row a CTA (dense graphs), 16 = the listed hub rows
  a.rowptr = rowptr;
  a.col = col;
  if (coop) {
    rc = gat_bwd_rows_dispatch<T, 4>(a, cplr, seg, (unsigned)n, st);
    if (rc != GNN_OK) return rc;
  } else {
    if (n_long > 0) {
      a.row_list = long_rows;
      rc = gat_bwd_rows_dispatch<T, 16>(a, cplr, seg, (unsigned)n_long, st);
      if (rc != GNN_OK) return rc;
      a.row_list = nullptr;
      a.skip_deg_gt = long_threshold;
    }
    rc = gat_bwd_rows_dispatch<T, 1>(a, cplr, seg, grid1, st);
    if (rc != GNN_OK) return rc;
  }
  // kernel B over the transposed CSR (the stash is in its own slot order when edge_to_tslot was given)
  a.rowptr = rowptr_t;
  a.col = col_t;
  a.perm = edge_to_tslot ? nullptr : perm_t;
  a.row_list = nullptr;
  a.skip_deg_gt = 0;
  if (coop) return gat_bwd_cols_dispatch<T, 4>(a, cplr, (unsigned)n, st);
  if (n_long_t > 0) {
    a.row_list = long_rows_t;
    rc = gat_bwd_cols_dispatch<T, 16>(a, cplr, (unsigned)n_long_t, st);
    if (rc != GNN_OK) return rc;
    a.row_list = nullptr;
    a.skip_deg_gt = long_threshold;
  }
  return gat_bwd_cols_dispatch<T, 1>(a, cplr, grid1, st);
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
                          const float* t, int64_t n, int64_t nnz, int32_t H, int32_t Fp, float alpha, int mode,
                          int apply_elu, const float* col_mean, const float* edge_keep, float* out, int64_t ldo,
                          float* row_max, float* row_sum, const int64_t* long_rows, int64_t n_long,
                          int64_t long_threshold, gnn_stream_t stream) {
  return gat_fwd_impl<float>(rowptr, col, Wh, ldw, s, t, n, nnz, H, Fp, alpha, mode, apply_elu, col_mean, edge_keep, out,
                             ldo, row_max, row_sum, long_rows, n_long, long_threshold, (cudaStream_t)stream);
}

int gnn_gat_fused_fwd_bf16(const int64_t* rowptr, const int32_t* col, const void* Wh, int64_t ldw, const float* s,
                           const float* t, int64_t n, int64_t nnz, int32_t H, int32_t Fp, float alpha, int mode,
                           int apply_elu, const float* col_mean, const float* edge_keep, void* out, int64_t ldo,
                           float* row_max, float* row_sum, const int64_t* long_rows, int64_t n_long,
                           int64_t long_threshold, gnn_stream_t stream) {
  return gat_fwd_impl<__nv_bfloat16>(rowptr, col, (const __nv_bfloat16*)Wh, ldw, s, t, n, nnz, H, Fp, alpha, mode,
                                     apply_elu, col_mean, edge_keep, (__nv_bfloat16*)out, ldo, row_max, row_sum,
                                     long_rows, n_long, long_threshold, (cudaStream_t)stream);
}

int gnn_gat_fused_bwd_f32(const int64_t* rowptr, const int32_t* col, const int64_t* rowptr_t, const int32_t* col_t,
                          const int64_t* perm_t, const int64_t* edge_to_tslot, const float* Wh, int64_t ldw,
                          const float* s, const float* t, const float* row_max, const float* row_sum,
                          const float* out_pre, const float* d_out, int64_t ldo, int64_t n, int32_t H, int32_t Fp,
                          float alpha, int mode, const float* edge_keep, float* d_Wh, int64_t ld_dwh, float* d_s,
                          float* d_t, float* d_rowdot, float* edge_scratch, int64_t nnz, const int64_t* long_rows,
                          int64_t n_long, const int64_t* long_rows_t, int64_t n_long_t, int64_t long_threshold,
                          gnn_stream_t stream) {
  return gat_bwd_impl<float>(rowptr, col, rowptr_t, col_t, perm_t, edge_to_tslot, Wh, ldw, s, t, row_max, row_sum,
                             out_pre, d_out, ldo, n, H, Fp, alpha, mode, edge_keep, d_Wh, ld_dwh, d_s, d_t, d_rowdot,
                             edge_scratch, nnz, long_rows, n_long, long_rows_t, n_long_t, long_threshold,
                             (cudaStream_t)stream);
}

int gnn_gat_fused_bwd_bf16(const int64_t* rowptr, const int32_t* col, const int64_t* rowptr_t, const int32_t* col_t,
                           const int64_t* perm_t, const int64_t* edge_to_tslot, const void* Wh, int64_t ldw,
                           const float* s, const float* t, const float* row_max, const float* row_sum,
                           const void* out_pre, const void* d_out, int64_t ldo, int64_t n, int32_t H, int32_t Fp,
                           float alpha, int mode, const float* edge_keep, void* d_Wh, int64_t ld_dwh, float* d_s,
                           float* d_t, float* d_rowdot, float* edge_scratch, int64_t nnz, const int64_t* long_rows,
                           int64_t n_long, const int64_t* long_rows_t, int64_t n_long_t, int64_t long_threshold,
                           gnn_stream_t stream) {
  return gat_bwd_impl<__nv_bfloat16>(rowptr, col, rowptr_t, col_t, perm_t, edge_to_tslot, (const __nv_bfloat16*)Wh, ldw,
                                     s, t, row_max, row_sum, (const __nv_bfloat16*)out_pre, (const __nv_bfloat16*)d_out,
                                     ldo, n, H, Fp, alpha, mode, edge_keep, (__nv_bfloat16*)d_Wh, ld_dwh, d_s, d_t,
                                     d_rowdot, edge_scratch, nnz, long_rows, n_long, long_rows_t, n_long_t,
                                     long_threshold, (cudaStream_t)stream);
}

}  // extern "C"
