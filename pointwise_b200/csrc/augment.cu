// augment.cu -- the per-step host work of the reference's data providers, on the GPU:
//   * rotate_point_cloud + jitter_point_cloud (modelnet_provider.py:23-41, 64-75): a random rotation about the up (y)
//     axis per cloud and a clipped Gaussian offset per coordinate.  The random draws (one angle per cloud, one
//     standard-normal sample per coordinate) are INPUTS, so that the same draws give the same batch as numpy;
//     arithmetic is done in double like numpy's (float32 data x float64 matrix / noise) and rounded to float32 once
//     per stage, as the provider's float32 arrays do.
//   * sort_point_cloud_xyz / sort_point_cloud_xyz2 (util.py:55-109): rows of every cloud sorted by x, then y, then z
//     (three chained stable sorts, least significant field first), attributes permuted along.
#include "common.cuh"
#include "radix_sort.cuh"

namespace c3p {

// out[b, i, :] = float32( clip(sigma * noise[b, i, :], -clip, clip) + float32( data[b, i, :] @ R(angle[b]) ) )
// R = [[c, 0, s], [0, 1, 0], [-s, 0, c]]  (modelnet_provider.py:36-40).  noise == nullptr: rotation only;
// angles == nullptr: jitter only.
__global__ void k_rotate_jitter(const float* __restrict__ data, const double* __restrict__ angles,
                                const double* __restrict__ noise, double sigma, double clip, long long B, int N,
                                float* __restrict__ out) {
  const long long total = B * N;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long b = e / N;
    const float* p = data + 3 * e;
    float r[3] = {p[0], p[1], p[2]};
    if (angles) {
      const double ang = angles[b], c = cos(ang), s = sin(ang);
      const double x = p[0], y = p[1], z = p[2];
      // row vector times matrix, summed in index order like a 3-term dot product
      r[0] = (float)(x * c + y * 0.0 + z * (-s));
      r[1] = (float)(x * 0.0 + y * 1.0 + z * 0.0);
      r[2] = (float)(x * s + y * 0.0 + z * c);
    }
    if (noise) {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        double j = sigma * noise[3 * e + a];
        j = j < -clip ? -clip : (j > clip ? clip : j);   // np.clip
        r[a] = (float)(j + (double)r[a]);
      }
    }
    out[3 * e + 0] = r[0];
    out[3 * e + 1] = r[1];
    out[3 * e + 2] = r[2];
  }
}

// Order-preserving map float -> uint32 (-0.0 == +0.0, as in a comparison sort).
__device__ __forceinline__ uint32_t float_key(float v) {
  const uint32_t u = __float_as_uint(v + 0.0f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// One CTA per cloud: order[b, s] = row of the cloud that comes s-th in (x, y, z) order, ties by original index.
__global__ void __launch_bounds__(SORT_THREADS)
k_xyz_order(const float* __restrict__ data, int N, int K, uint32_t* __restrict__ tmp, int* __restrict__ order) {
  __shared__ uint32_t base[256];
  __shared__ uint32_t wsum[8];
  __shared__ uint16_t wcount[SORT_WARPS][256];
  __shared__ uint16_t wpre[SORT_WARPS][256];
  const int b = blockIdx.x, tid = threadIdx.x;
  if (N == 0) return;
  const float* P = data + (size_t)b * N * K;
  uint32_t* kin = tmp + (size_t)b * 4 * N;
  uint32_t* iin = kin + N;
  uint32_t* kout = kin + 2 * (size_t)N;
  uint32_t* iout = kin + 3 * (size_t)N;
  for (int i = tid; i < N; i += SORT_THREADS) iin[i] = (uint32_t)i;
  for (int e = tid; e < SORT_WARPS * 256; e += SORT_THREADS) (&wcount[0][0])[e] = 0;
  __syncthreads();
  RadixTables tb{base, wsum, wcount, wpre};
  for (int field = 2; field >= 0; --field) {   // least significant field first: z, then y, then x
    for (int i = tid; i < N; i += SORT_THREADS) kin[i] = float_key(P[(size_t)iin[i] * K + field]);
    __syncthreads();
    radix_sort_pairs(kin, iin, kout, iout, N, 0, 32, tb);
  }
  for (int i = tid; i < N; i += SORT_THREADS) order[(size_t)b * N + i] = (int)iin[i];
}

// dst[b, s, :] = src[b, order[b, s], :]
__global__ void k_permute_rows(const float* __restrict__ src, const int* __restrict__ order, long long B, int N,
                               int C, float* __restrict__ dst) {
  const long long total = B * N * C;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long row = e / C;
    const int c = (int)(e - row * C);
    const long long b = row / N;
    dst[e] = src[(b * N + order[row]) * C + c];
  }
}

}  // namespace c3p

using namespace c3p;

extern "C" {

int conv3p_augment_rotate_jitter_f32(const float* data, const double* angles, const double* noise, double sigma,
                                     double clip, int B, int N, float* out, conv3p_stream_t stream) {
  if (B < 0 || N < 0 || !(clip > 0.0)) return CONV3P_ERR_INVALID_ARGUMENT;   // assert(clip > 0), :72
  const long long total = (long long)B * N;
  if (total == 0) return CONV3P_OK;
  if (!data || !out) return CONV3P_ERR_INVALID_ARGUMENT;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  {
    LaunchTimer timer_("k_rotate_jitter", stream);
    k_rotate_jitter<<<(unsigned)blocks, 256, 0, stream>>>(data, angles, noise, sigma, clip, B, N, out);
  }
  C3P_LAUNCH_CHECK("k_rotate_jitter");
  return CONV3P_OK;
}

size_t conv3p_xyz_sort_workspace_bytes(int B, int N) {
  if (B < 0 || N < 0) return 0;
  return align_up(sizeof(uint32_t) * 4 * (size_t)B * N) + 256;
}

int conv3p_xyz_sort_f32(const float* data, int K, const float* attributes, int M, int B, int N, int* order,
                        float* sorted_data, float* sorted_attributes, void* workspace, size_t workspace_bytes,
                        conv3p_stream_t stream) {
  if (B < 0 || N < 0 || K < 3 || M < 0 || N >= (1 << 27)) return CONV3P_ERR_INVALID_ARGUMENT;
  if ((long long)B * N == 0) return CONV3P_OK;
  if (!data || !order) return CONV3P_ERR_INVALID_ARGUMENT;
  if (!workspace || workspace_bytes < conv3p_xyz_sort_workspace_bytes(B, N)) return CONV3P_ERR_BUFFER_TOO_SMALL;
  {
    LaunchTimer timer_("k_xyz_order", stream);
    k_xyz_order<<<B, SORT_THREADS, 0, stream>>>(data, N, K, static_cast<uint32_t*>(workspace), order);
  }
  C3P_LAUNCH_CHECK("k_xyz_order");
  auto permute = [&](const float* src, int C, float* dst) -> int {
    const long long total = (long long)B * N * C;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    {
      LaunchTimer timer_("k_permute_rows", stream);
      k_permute_rows<<<(unsigned)blocks, 256, 0, stream>>>(src, order, B, N, C, dst);
    }
    C3P_LAUNCH_CHECK("k_permute_rows");
    return CONV3P_OK;
  };
  if (sorted_data) {
    const int st = permute(data, K, sorted_data);
    if (st) return st;
  }
  if (attributes && sorted_attributes && M > 0) {
    const int st = permute(attributes, M, sorted_attributes);
    if (st) return st;
  }
  return CONV3P_OK;
}

}  // extern "C"
