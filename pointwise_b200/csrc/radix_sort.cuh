// radix_sort.cuh -- a CTA-wide stable LSD radix sort of (key, payload) pairs in global memory with its digit
// histograms and per-warp rank tables in shared memory; used by the voxel sort of the neighbour plan (sort.cu) and by
// the xyz sort of the input pipeline (augment.cu).
#pragma once
#include "common.cuh"

namespace c3p {

constexpr int SORT_THREADS = 512;
constexpr int SORT_WARPS = SORT_THREADS / 32;

struct RadixTables {
  uint32_t* base;                    // [256]
  uint32_t* wsum;                    // [8]
  uint16_t (*wcount)[256];           // [SORT_WARPS][256], all zero on entry and on exit
  uint16_t (*wpre)[256];             // [SORT_WARPS][256]
};

// Stable LSD radix sort of (key, payload) pairs on bits [first_bit, first_bit + nbits) of the key, 8 bits per
// pass, by the whole CTA (SORT_THREADS threads).  The sorted pairs end up in (kin, iin) -- the buffers are
// swapped after every pass, by reference.
__device__ inline void radix_sort_pairs(uint32_t*& kin, uint32_t*& iin, uint32_t*& kout, uint32_t*& iout, int N,
                                 int first_bit, int nbits, const RadixTables& tb) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int shift = first_bit; shift < first_bit + nbits; shift += 8) {
    if (tid < 256) tb.base[tid] = 0;
    __syncthreads();
    for (int i = tid; i < N; i += SORT_THREADS) atomicAdd(&tb.base[(kin[i] >> shift) & 255u], 1u);
    __syncthreads();
    uint32_t h = 0, inc = 0;
    if (tid < 256) {  // exclusive scan of the 256 digit counts
      h = tb.base[tid];
      inc = h;
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t u = __shfl_up_sync(C3P_FULL_MASK, inc, o);
        if (lane >= o) inc += u;
      }
      if (lane == 31) tb.wsum[warp] = inc;
    }
    __syncthreads();
    if (tid < 256) {
      uint32_t off = 0;
      for (int w = 0; w < warp; ++w) off += tb.wsum[w];
      tb.base[tid] = off + inc - h;
    }
    __syncthreads();

    for (int c0 = 0; c0 < N; c0 += SORT_THREADS) {
      const int i = c0 + tid;
      const bool valid = i < N;
      uint32_t key = 0, idx = 0, d = 0xffffffffu;
      if (valid) {
        key = kin[i];
        idx = iin[i];
        d = (key >> shift) & 255u;
      }
      const unsigned peers = __match_any_sync(C3P_FULL_MASK, d);
      const int rank = __popc(peers & lanemask_lt());
      if (valid && rank == 0) tb.wcount[warp][d] = (uint16_t)__popc(peers);
      __syncthreads();
      uint32_t run = 0;
      if (tid < 256) {
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w) {
          tb.wpre[w][tid] = (uint16_t)run;
          run += tb.wcount[w][tid];
        }
      }
      __syncthreads();
      if (valid) {
        uint32_t pos = tb.base[d] + tb.wpre[warp][d] + rank;
        kout[pos] = key;
        iout[pos] = idx;
        if (rank == 0) tb.wcount[warp][d] = 0;  // leave the table clean for the next chunk
      }
      __syncthreads();
      if (tid < 256) tb.base[tid] += run;
    }
    __syncthreads();
    uint32_t* t0 = kin; kin = kout; kout = t0;
    uint32_t* t1 = iin; iin = iout; iout = t1;
  }

}

}  // namespace c3p
