// pm_rmat.cuh — bit-exact reproduction of the reference's R-MAT input on the GPU.
//
// Reference: /root/reference/src/generate_rmat.cpp:197-205 (16 * 2^scale edges split
// evenly over the generating ranks, rank r seeded 5489 + 3r, a,b,c,d = .57,.19,.19,.05,
// scramble on, undirected on), include/havoqgt/rmat_edge_generator.hpp:126-139 (every
// generated (u,v) is followed by (v,u)), :218-259 (generate_edge) and
// include/havoqgt/detail/hash.hpp:65-143 (hash_nbits).
//
// The generator draws from boost::uniform_01<boost::mt19937> (Boost 1.57, not vendored
// by the reference): mt19937 output x mapped to x * 2^-32.  An edge consumes exactly
// 5 * scale draws (the mapping never rejects), so draw positions are known in advance:
// one CTA owns one generating rank's stream, advances the 624-word Mersenne-Twister
// state with a three-stage parallel twist in shared memory, stages the tempered
// outputs of 5*scale twists (= the draws of 624 edges) in an L2-resident scratch slab
// and then lets its threads evaluate those 624 edges independently.  All probability
// arithmetic uses explicit round-to-nearest double intrinsics in the reference's
// operation order, so no FMA contraction can change a comparison.
#pragma once

#include "pm_common.cuh"
#include "pm_graph.cuh"

namespace pm {

__device__ __forceinline__ uint32_t d_mix32(uint32_t a) {  // hash32, hash.hpp:65-74
  a = (a + 0x7ed55d16u) + (a << 12);
  a = (a ^ 0xc761c23cu) ^ (a >> 19);
  a = (a + 0x165667b1u) + (a << 5);
  a = (a + 0xd3a2646cu) ^ (a << 9);
  a = (a + 0xfd7046c5u) + (a << 3);
  a = (a ^ 0xb55a4f09u) ^ (a >> 16);
  return a;
}
__device__ __forceinline__ uint32_t d_mix16(uint32_t a) {  // hash16 with uint16_t truncation, hash.hpp:76-85
  a &= 0xffffu;
  a = ((a + 0x5d16u) + (a << 6)) & 0xffffu;
  a = ((a ^ 0xc23cu) ^ (a >> 9)) & 0xffffu;
  a = ((a + 0x67b1u) + (a << 5)) & 0xffffu;
  a = ((a + 0x646cu) ^ (a << 7)) & 0xffffu;
  a = ((a + 0x46c5u) + (a << 3)) & 0xffffu;
  a = ((a ^ 0x4f09u) ^ (a >> 8)) & 0xffffu;
  return a;
}
__device__ __forceinline__ unsigned long long d_hash_nbits(unsigned long long x, int n) {  // hash.hpp:115-143
  if (n == 32) return d_mix32((uint32_t)x);
  if (n > 32) {
    const int k = n - 32;
    for (int i = 0; i <= k; ++i) {
      const unsigned long long m = 0xffffffffull << i;
      x = (x & ~m) | ((unsigned long long)d_mix32((uint32_t)((x >> i) & 0xffffffffull)) << i);
    }
    for (int i = k; i >= 0; --i) {
      const unsigned long long m = 0xffffffffull << i;
      x = (x & ~m) | ((unsigned long long)d_mix32((uint32_t)((x >> i) & 0xffffffffull)) << i);
    }
    return x;
  }
  const int k = n - 16;
  for (int i = 0; i <= k; ++i) {
    const unsigned long long m = 0xffffull << i;
    x = (x & ~m) | ((unsigned long long)d_mix16((uint32_t)((x >> i) & 0xffffull)) << i);
  }
  for (int i = k; i >= 0; --i) {
    const unsigned long long m = 0xffffull << i;
    x = (x & ~m) | ((unsigned long long)d_mix16((uint32_t)((x >> i) & 0xffffull)) << i);
  }
  return x;
}

__device__ __forceinline__ double d_u01(uint32_t x) { return __dmul_rn((double)x, 1.0 / 4294967296.0); }

// rmat_edge_generator.hpp:218-259, draws[0 .. 5*scale)
__device__ __forceinline__ void d_rmat_edge(const uint32_t* __restrict__ draws, int scale, uint32_t& u_out,
                                            uint32_t& v_out) {
  double a = 0.57, b = 0.19, c = 0.19, d = 0.05;
  unsigned long long u = 0, v = 0, step = (1ull << scale) >> 1;
  for (int j = 0; j < scale; ++j) {
    const double p = d_u01(draws[5 * j]);
    const double ab = __dadd_rn(a, b);
    const double abc = __dadd_rn(ab, c);
    if (p < a) {
    } else if (p >= a && p < ab) {
      v += step;
    } else if (p >= ab && p < abc) {
      u += step;
    } else {
      u += step;
      v += step;
    }
    step >>= 1;
    a = __dmul_rn(a, __dadd_rn(0.9, __dmul_rn(0.2, d_u01(draws[5 * j + 1]))));
    b = __dmul_rn(b, __dadd_rn(0.9, __dmul_rn(0.2, d_u01(draws[5 * j + 2]))));
    c = __dmul_rn(c, __dadd_rn(0.9, __dmul_rn(0.2, d_u01(draws[5 * j + 3]))));
    d = __dmul_rn(d, __dadd_rn(0.9, __dmul_rn(0.2, d_u01(draws[5 * j + 4]))));
    const double S = __dadd_rn(__dadd_rn(__dadd_rn(a, b), c), d);
    a = __ddiv_rn(a, S);
    b = __ddiv_rn(b, S);
    c = __ddiv_rn(c, S);
    d = __dsub_rn(__dsub_rn(__dsub_rn(1.0, a), b), c);
  }
  u_out = (uint32_t)d_hash_nbits(u, scale);
  v_out = (uint32_t)d_hash_nbits(v, scale);
}

#define PM_MT_N 624
#define PM_MT_M 397

// one CTA per generating rank; scratch: gridDim.x slabs of 624 * 5 * scale words
// (with several GPUs, GPU g generates the streams of the generating ranks r = blockIdx.x * n_gpus + g)
__global__ void __launch_bounds__(256) k_rmat_stream(int scale, uint64_t per_rank, uint32_t first_rank,
                                                      uint32_t rank_stride, uint32_t* __restrict__ scratch,
                                                      uint32_t* __restrict__ src, uint32_t* __restrict__ dst) {
  __shared__ uint32_t xa[PM_MT_N], xb[PM_MT_N];
  const uint32_t r = first_rank + blockIdx.x * rank_stride;
  const int tid = threadIdx.x;
  const int dpe = 5 * scale;  // draws per edge
  uint32_t* slab = scratch + (uint64_t)blockIdx.x * PM_MT_N * dpe;
  if (tid == 0) {  // mt19937 seeding, seed = 5489 + 3 * rank (generate_rmat.cpp:202)
    uint32_t x = 5489u + 3u * r;
    xa[0] = x;
    for (int i = 1; i < PM_MT_N; ++i) {
      x = 1812433253u * (x ^ (x >> 30)) + (uint32_t)i;
      xa[i] = x;
    }
  }
  __syncthreads();
  uint32_t* cur = xa;
  uint32_t* nxt = xb;
  for (uint64_t e0 = 0; e0 < per_rank; e0 += PM_MT_N) {
    // ---- 5*scale twists: the draws of edges [e0, e0 + 624)
    for (int t = 0; t < dpe; ++t) {
      // new[k] = x[(k+397)%624] ^ twist(x[k], x[(k+1)%624]); entries the sequential
      // algorithm has already replaced when it reaches k are read from `nxt`
      for (int k = tid; k < 227; k += blockDim.x) {
        const uint32_t y = (cur[k] & 0x80000000u) | (cur[k + 1] & 0x7fffffffu);
        nxt[k] = cur[k + PM_MT_M] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      }
      __syncthreads();
      for (int k = 227 + tid; k < 454; k += blockDim.x) {
        const uint32_t y = (cur[k] & 0x80000000u) | (cur[k + 1] & 0x7fffffffu);
        nxt[k] = nxt[k - 227] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      }
      __syncthreads();
      for (int k = 454 + tid; k < PM_MT_N; k += blockDim.x) {
        const uint32_t up = (k == PM_MT_N - 1) ? nxt[0] : cur[k + 1];
        const uint32_t y = (cur[k] & 0x80000000u) | (up & 0x7fffffffu);
        nxt[k] = nxt[k - 227] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      }
      __syncthreads();
      for (int k = tid; k < PM_MT_N; k += blockDim.x) {  // tempering
        uint32_t y = nxt[k];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        slab[t * PM_MT_N + k] = y;
      }
      uint32_t* tmp = cur;
      cur = nxt;
      nxt = tmp;
      __syncthreads();
    }
    __threadfence_block();
    // ---- evaluate the 624 edges
    for (int i = tid; i < PM_MT_N; i += blockDim.x) {
      const uint64_t e = e0 + i;
      if (e < per_rank) {
        uint32_t u, v;
        d_rmat_edge(slab + (uint64_t)i * dpe, scale, u, v);
        const uint64_t o = 2 * ((uint64_t)blockIdx.x * per_rank + e);
        src[o] = u; dst[o] = v;          // generated edge
        src[o + 1] = v; dst[o + 1] = u;  // reversed copy (rmat_edge_generator.hpp:126-139)
      }
    }
    __syncthreads();
  }
}

inline int rmat_slots_device(pm_ctx* c, uint64_t scale, uint64_t gen_ranks, uint32_t** d_src, uint32_t** d_dst,
                             uint64_t* n_slots) {
  if (scale <= 16 || scale > 31) return fail(c, PM_ERR_ARG, "R-MAT scale must be in 17..31 (hash_nbits needs n > 16)");
  if (gen_ranks == 0 || gen_ranks > 65535) return fail(c, PM_ERR_ARG, "gen_ranks must be in 1..65535");
  const uint64_t V = 1ull << scale;
  const uint64_t per_rank = V * 16 / gen_ranks;  // generate_rmat.cpp:201
  // this GPU's share of the generating ranks
  const uint64_t G = c->n_ranks;
  const uint64_t mine = gen_ranks / G + ((uint64_t)c->rank < gen_ranks % G ? 1 : 0);
  const uint64_t n = 2 * per_rank * mine;
  int rc;
  uint32_t* scratch = nullptr;
  if ((rc = dev_alloc(c, d_src, n))) return rc;
  if ((rc = dev_alloc(c, d_dst, n))) { dev_free(*d_src); return rc; }
  if ((rc = dev_alloc(c, &scratch, mine * (uint64_t)PM_MT_N * 5 * scale))) { dev_free(*d_src); dev_free(*d_dst); return rc; }
  if (mine) {
    k_rmat_stream<<<(unsigned)mine, 256, 0, c->stream>>>((int)scale, per_rank, (uint32_t)c->rank, (uint32_t)G, scratch,
                                                         *d_src, *d_dst);
    c->launches++;
  }
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  dev_free(scratch);
  if (e != cudaSuccess) { dev_free(*d_src); dev_free(*d_dst); return fail(c, PM_ERR_CUDA, cudaGetErrorString(e)); }
  *n_slots = n;
  return 0;
}

inline int rmat_build(pm_ctx* c, uint64_t scale, uint64_t gen_ranks) {
  uint32_t *d_src = nullptr, *d_dst = nullptr;
  uint64_t n = 0;
  int rc = rmat_slots_device(c, scale, gen_ranks, &d_src, &d_dst, &n);
  if (rc) return rc;
  rc = graph_build_from_device_slots(c, 1ull << scale, n, d_src, d_dst, /*route=*/true);
  dev_free(d_src);
  dev_free(d_dst);
  return rc;
}

}  // namespace pm
