// Fixed-fanout neighbour sampler on the device.
// Semantics of /root/reference GraphSAGE_Pytorch/sample_utils.py:4-17 (and the per-node step
// of GraphSAGE/data_utils.py:91-94): for every source node, k DISTINCT neighbours uniformly at
// random when the node has at least k (random.sample), otherwise k draws with replacement
// (random.choices); results flat, src-major (sample_utils.py:16).  It is semantically — not
// bit- — equal to the reference: Python's Mersenne-Twister stream is not reproduced (SURVEY.md
// §8 a9).  Every (source, draw) pair is an independent thread:
//   without replacement: draw j takes position P(j) of a keyed pseudo-random permutation P of
//     [0, deg) (bijective multiply / xor-shift rounds on ceil(log2 deg) bits + cycle walking),
//     so the k draws are distinct by construction — no rejection loop, no shared state;
//   with replacement:    hash(seed, i, j) mod deg.
#include "common.cuh"

using namespace gnn;

namespace {

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9e3779b97f4a7c15ULL;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}

// keyed bijection of [0, 2^bits): every step (odd multiply, add, xor-shift-right) is invertible mod 2^bits
__device__ __forceinline__ uint32_t permute_bits(uint32_t x, int bits, uint64_t key) {
  const uint32_t mask = (bits >= 32) ? 0xffffffffu : ((1u << bits) - 1u);
  const int sh = bits > 1 ? (bits + 1) / 2 : 1;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const uint32_t k1 = (uint32_t)(key >> (16 * r)) | 1u;
    const uint32_t k2 = (uint32_t)(key >> (8 * r + 3));
    x = (x * k1 + k2) & mask;
    x ^= x >> sh;
    x = (x * 0x9E3779B1u) & mask;
    x ^= x >> (sh > 1 ? sh - 1 : 1);
  }
  return x & mask;
}

template <typename SRC, typename OUT>
__global__ void __launch_bounds__(256) sample_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                     const SRC* __restrict__ src, int64_t n_src, int k, uint64_t seed,
                                                     const int64_t* __restrict__ seed_offset, OUT* __restrict__ out) {
  const int64_t total = n_src * k;
  if (seed_offset) seed += (uint64_t)__ldg(seed_offset) * 0x9E3779B97F4A7C15ULL;  // per-replay stream of a CUDA graph
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = p / k;
    const int j = (int)(p - i * k);
    const int64_t s = (int64_t)src[i];
    OUT res = (OUT)-1;
    if (s >= 0) {
      const int64_t e0 = __ldg(rowptr + s), e1 = __ldg(rowptr + s + 1);
      const uint32_t deg = (uint32_t)(e1 - e0);
      if (deg > 0) {
        const uint64_t key = mix64(seed ^ mix64((uint64_t)i * 0x100000001b3ULL + 0x51ed27ULL));
        uint32_t pos;
        if (deg >= (uint32_t)k) {
          int bits = 1;
          while ((1u << bits) < deg && bits < 31) ++bits;
          pos = permute_bits((uint32_t)j, bits, key);
          while (pos >= deg) pos = permute_bits(pos, bits, key);  // cycle walking keeps it a bijection on [0,deg)
        } else {
          pos = (uint32_t)(mix64(key + (uint64_t)j) % deg);
        }
        res = (OUT)__ldg(col + e0 + pos);
      }
    }
    out[p] = res;
  }
}

}  // namespace

extern "C" int gnn_sample_neighbors(const int64_t* rowptr, const int32_t* col, const void* src, int src_bits,
                                    int64_t n_src, int32_t k, uint64_t seed, const int64_t* seed_offset_dev, void* out,
                                    int out_bits, gnn_stream_t stream) {
  GNN_REQUIRE(n_src >= 0 && k > 0, GNN_ERR_BAD_ARG, "bad size (n_src=%lld, k=%d)", (long long)n_src, k);
  GNN_REQUIRE((src_bits == 32 || src_bits == 64) && (out_bits == 32 || out_bits == 64), GNN_ERR_BAD_ARG,
              "id widths must be 32 or 64");
  if (n_src == 0) return GNN_OK;
  GNN_REQUIRE(rowptr && src && out, GNN_ERR_BAD_ARG, "null pointer");
  int64_t grid = (n_src * k + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 16;
  grid = grid > cap ? cap : grid;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned g = (unsigned)grid;
  if (src_bits == 32 && out_bits == 32)
    sample_kernel<int32_t, int32_t><<<g, 256, 0, st>>>(rowptr, col, (const int32_t*)src, n_src, k, seed, seed_offset_dev,
                                                          (int32_t*)out);
  else if (src_bits == 32)
    sample_kernel<int32_t, int64_t><<<g, 256, 0, st>>>(rowptr, col, (const int32_t*)src, n_src, k, seed, seed_offset_dev,
                                                          (int64_t*)out);
  else if (out_bits == 32)
    sample_kernel<int64_t, int32_t><<<g, 256, 0, st>>>(rowptr, col, (const int64_t*)src, n_src, k, seed, seed_offset_dev,
                                                          (int32_t*)out);
  else
    sample_kernel<int64_t, int64_t><<<g, 256, 0, st>>>(rowptr, col, (const int64_t*)src, n_src, k, seed, seed_offset_dev,
                                                          (int64_t*)out);
  GNN_LAUNCH_CHECK();
  return GNN_OK;
}
