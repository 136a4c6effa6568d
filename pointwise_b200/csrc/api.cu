// api.cu -- the C ABI of libconv3p_b200.so (declared in include/conv3p_b200.h).
#include <algorithm>
#include <atomic>
#include <map>
#include <mutex>
#include <set>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>
#include <cstdio>
#include <cstring>

#include "common.cuh"

namespace c3p {

static thread_local char g_cuda_error[512] = "";
static std::atomic<long long> g_launches{0};   // process-wide: autograd runs the backward pass on its own thread
static std::atomic<int> g_engine{0};

int cuda_fail(cudaError_t e, const char* what) {
  snprintf(g_cuda_error, sizeof(g_cuda_error), "%s: %s (%s)", what, cudaGetErrorName(e),
           cudaGetErrorString(e));
  (void)cudaGetLastError();
  return CONV3P_ERR_CUDA;
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int engine() { return g_engine.load(std::memory_order_relaxed); }

// ---- per-device facts and per-kernel attributes, resolved once ------------------------------------------
int sm_count() {
  static std::atomic<int> cached[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
    (void)cudaGetLastError();
    return 148;
  }
  int n = cached[dev].load(std::memory_order_relaxed);
  if (n > 0) return n;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) {
    (void)cudaGetLastError();
    n = 148;
  }
  cached[dev].store(n, std::memory_order_relaxed);
  return n;
}

int ensure_dynamic_smem(const void* kernel, size_t bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, size_t> done;   // (kernel, device) -> bytes granted
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return cuda_fail(cudaGetLastError(), "cudaGetDevice");
  std::lock_guard<std::mutex> lock(mu);
  size_t& have = done[std::make_pair(kernel, dev)];   // largest request already served for this (kernel, device)
  if (have >= bytes) return CONV3P_OK;
  const size_t requested = bytes;
  // static + dynamic shared memory share the 227 KB a CTA may opt in to
  cudaFuncAttributes fa;
  cudaError_t e = cudaFuncGetAttributes(&fa, kernel);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncGetAttributes");
  const size_t limit = 227 * 1024 - fa.sharedSizeBytes;
  if (bytes > limit) bytes = limit;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
  have = requested;
  return CONV3P_OK;
}

// ---- which plans have backward lists (host-side mirror of header[H_HAS_BWD]) ------------------------------
// The entry points are called in stream order, so the state of a plan buffer can be tracked per address without
// reading the device: conv3p_plan_build_f32 marks it "forward lists only", conv3p_plan_build_backward "complete".
// Unknown addresses (a plan built elsewhere and copied) are given the benefit of the doubt.
static std::mutex g_plan_mutex;
static std::unordered_map<const void*, bool> g_plan_has_bwd;
static void plan_mark(const void* plan, bool has_bwd) {
  std::lock_guard<std::mutex> lock(g_plan_mutex);
  if (g_plan_has_bwd.size() > 4096) g_plan_has_bwd.clear();   // bounded: addresses recycle
  g_plan_has_bwd[plan] = has_bwd;
}
static bool plan_known_without_backward(const void* plan) {
  std::lock_guard<std::mutex> lock(g_plan_mutex);
  auto it = g_plan_has_bwd.find(plan);
  return it != g_plan_has_bwd.end() && !it->second;
}

// ---- ordered reduction of grad_filter partials, NaN when the plan's lists overflowed ------------------------
__global__ void k_reduce_partials(const float* __restrict__ partial, int S, long long nW, float* __restrict__ out,
                                  const long long* __restrict__ header) {
  const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nW) return;
  float s = 0.f;
  for (int i = 0; i < S; ++i) s += partial[(size_t)i * nW + w];  // fixed order: deterministic
  if (header && header[H_OVERFLOW] != 0) s = __int_as_float(0x7fc00000);
  out[w] = s;
}

// The plan's counters written straight into MAPPED pinned host memory by a one-warp kernel: no copy engine is
// involved, so the deferred overflow check never queues behind (or in front of) an application's bulk copies.
__global__ void k_publish_header(const long long* __restrict__ header, volatile long long* host_out) {
  if (threadIdx.x < H_SLOTS) host_out[threadIdx.x] = header[threadIdx.x];
}

int launch_reduce_partials(const float* partial, int S, long long nW, float* out, const long long* plan_header,
                           cudaStream_t stream) {
  {
    LaunchTimer timer_("k_reduce_partials", stream);
    k_reduce_partials<<<(unsigned)((nW + 255) / 256), 256, 0, stream>>>(partial, S, nW, out, plan_header);
  }
  C3P_LAUNCH_CHECK("k_reduce_partials");
  return CONV3P_OK;
}

// ---- optional per-kernel event timing ---------------------------------------------------------------
struct TimedLaunch {
  const char* name;
  cudaEvent_t e0, e1;
};
static std::mutex g_prof_mutex;
static std::vector<TimedLaunch> g_prof_records;
static std::vector<cudaEvent_t> g_prof_pool;
static std::atomic<int> g_prof_on{0};

static cudaEvent_t prof_event() {
  if (!g_prof_pool.empty()) {
    cudaEvent_t e = g_prof_pool.back();
    g_prof_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

LaunchTimer::LaunchTimer(const char* name, cudaStream_t s) : slot(-1), stream(s) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  TimedLaunch t{name, prof_event(), prof_event()};
  cudaEventRecord(t.e0, stream);
  g_prof_records.push_back(t);
  slot = (int)g_prof_records.size() - 1;
}

LaunchTimer::~LaunchTimer() {
  if (slot < 0) return;
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  if (slot < (int)g_prof_records.size()) cudaEventRecord(g_prof_records[slot].e1, stream);
}

int check_geom(const conv3p_geom_t* g) {
  if (!g) return CONV3P_ERR_INVALID_ARGUMENT;
  if (g->B < 0 || g->N < 0 || g->N >= (1 << 27)) return CONV3P_ERR_INVALID_ARGUMENT;
  if ((long long)g->B * g->N >= (1LL << 31)) return CONV3P_ERR_INVALID_ARGUMENT;
  for (int a = 0; a < 3; ++a)
    if (g->stride[a] < 1 || g->stride[a] > 4096) return CONV3P_ERR_INVALID_ARGUMENT;
  if (!(g->voxel_size > 0.0f) || !(g->voxel_size < 1e30f)) return CONV3P_ERR_INVALID_ARGUMENT;
  if (g->pair_capacity < 0) return CONV3P_ERR_INVALID_ARGUMENT;
  return CONV3P_OK;
}

int compute_layout(const conv3p_geom_t* g, conv3p_plan_layout_t* L) {
  int st = check_geom(g);
  if (st) return st;
  const size_t pts = (size_t)g->B * g->N;
  const size_t cap = (size_t)g->pair_capacity;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    size_t at = o;
    o += align_up(bytes);
    return at;
  };
  L->header = take(sizeof(long long) * H_SLOTS);
  L->cloud_meta = take(sizeof(float) * 8 * (size_t)g->B);
  L->sorted_key = take(sizeof(uint32_t) * pts);
  L->sorted_xyzi = take(sizeof(float4) * pts);
  L->count_table = take(sizeof(int) * pts * C3P_NCELL);
  L->pair_begin = take(sizeof(long long) * pts);
  L->pair_len = take(sizeof(int) * pts);
  L->pair_row = take(sizeof(int) * cap);
  L->bwd_count = take(sizeof(int) * pts * C3P_NCELL);
  L->bwd_row = take(sizeof(int) * cap);
  L->bwd_weight = take(sizeof(float) * cap);
  L->sort_tmp = take(sizeof(uint32_t) * 4 * pts);
  L->cell_start = take(sizeof(uint32_t) * (size_t)g->B * ((size_t)cell_cap(g->N) + 1));
  L->total_bytes = o;
  return CONV3P_OK;
}

int make_view(const conv3p_geom_t* g, const void* plan, size_t plan_bytes, PlanView* v) {
  conv3p_plan_layout_t L;
  int st = compute_layout(g, &L);
  if (st) return st;
  if (!plan || plan_bytes < L.total_bytes) return CONV3P_ERR_BUFFER_TOO_SMALL;
  if (reinterpret_cast<uintptr_t>(plan) % 16 != 0) return CONV3P_ERR_INVALID_ARGUMENT;
  char* p = static_cast<char*>(const_cast<void*>(plan));
  v->header = reinterpret_cast<long long*>(p + L.header);
  v->cloud_meta = reinterpret_cast<float*>(p + L.cloud_meta);
  v->sorted_key = reinterpret_cast<uint32_t*>(p + L.sorted_key);
  v->sorted_xyzi = reinterpret_cast<float4*>(p + L.sorted_xyzi);
  v->count_table = reinterpret_cast<int*>(p + L.count_table);
  v->pair_begin = reinterpret_cast<long long*>(p + L.pair_begin);
  v->pair_len = reinterpret_cast<int*>(p + L.pair_len);
  v->pair_row = reinterpret_cast<int*>(p + L.pair_row);
  v->bwd_count = reinterpret_cast<int*>(p + L.bwd_count);
  v->bwd_row = reinterpret_cast<int*>(p + L.bwd_row);
  v->bwd_weight = reinterpret_cast<float*>(p + L.bwd_weight);
  v->sort_tmp = reinterpret_cast<uint32_t*>(p + L.sort_tmp);
  v->cell_start = reinterpret_cast<uint32_t*>(p + L.cell_start);
  v->cell_cap = cell_cap(g->N);
  return CONV3P_OK;
}

// G store shared by the grad_input and grad_filter kernels of one backward call (gather_mma2.cu / backward_filter2.cu):
// one Cout-wide row per (sorted position, cell), last region of the scratch buffer; 0 when the pair of kernels that
// uses it does not apply to the shape.
static size_t g_store_bytes(const conv3p_geom_t* g, int Cin, int Cout) {
  if (!gather_mma2_supported(g->N, g->pair_capacity, Cout, Cin) ||
      !backward_filter2_supported(g->N, g->pair_capacity, Cin, Cout))
    return 0;
  const size_t pts = ((size_t)g->B * g->N + 127) / 128 * 128;
  return align_up(pts * C3P_NCELL * (size_t)Cout * sizeof(float));
}

static int check_channels(int Cin, int Cout) {
  if (Cin < 1 || Cout < 1 || Cin > (1 << 16) || Cout > (1 << 16)) return CONV3P_ERR_INVALID_ARGUMENT;
  return CONV3P_OK;
}

static conv3p_geom_t make_geom(int B, int N, const int stride[3], float voxel, long long cap) {
  conv3p_geom_t g;
  g.B = B; g.N = N;
  g.stride[0] = stride ? stride[0] : 0;
  g.stride[1] = stride ? stride[1] : 0;
  g.stride[2] = stride ? stride[2] : 0;
  g.voxel_size = voxel;
  g.pair_capacity = cap;
  return g;
}

// out[r, c] = grad[r, c] * selu'(x), from the activated value y = selu(x): scale if y >= 0 (x >= 0, selu.py:25) else
// y + scale * alpha
__global__ void k_selu_backward(const float* __restrict__ y, long long ys, const float* __restrict__ g, long long gs,
                                float* __restrict__ out, long long rows, int C) {
  const float alpha = 1.6732632423543772848170429916717f, scale = 1.0507009873554804934193349852946f;
  const long long total = rows * C;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / C;
    const int c = (int)(e - r * C);
    const float yv = __ldg(y + r * ys + c);
    out[e] = __ldg(g + r * gs + c) * (yv >= 0.f ? scale : yv + scale * alpha);   // selu.py:25: linear branch for x >= 0
  }
}

}  // namespace c3p

using namespace c3p;

extern "C" {

int conv3p_abi_version(void) { return 1; }

const char* conv3p_status_string(int s) {
  switch (s) {
    case CONV3P_OK: return "ok";
    case CONV3P_ERR_INVALID_ARGUMENT: return "invalid argument";
    case CONV3P_ERR_BUFFER_TOO_SMALL: return "plan/scratch/workspace buffer too small";
    case CONV3P_ERR_CUDA: return "CUDA runtime error";
    case CONV3P_ERR_UNSUPPORTED: return "unsupported configuration (filter with more than 512 cells)";
    case CONV3P_ERR_PAIR_OVERFLOW: return "pair_capacity too small for this batch";
    case CONV3P_ERR_NO_BACKWARD_LISTS: return "backward lists not built";
    default: return "unknown status";
  }
}

const char* conv3p_last_cuda_error(void) { return g_cuda_error; }

long long conv3p_launch_count(int reset) {
  return reset ? g_launches.exchange(0) : g_launches.load();
}

int conv3p_set_engine(int e) { return g_engine.exchange(e); }

int conv3p_profile_enable(int on) {
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  for (auto& r : g_prof_records) {
    g_prof_pool.push_back(r.e0);
    g_prof_pool.push_back(r.e1);
  }
  g_prof_records.clear();
  return g_prof_on.exchange(on ? 1 : 0);
}

long long conv3p_profile_read(char* buf, size_t cap) {
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  std::map<std::string, std::pair<long long, double> > agg;
  for (auto& r : g_prof_records) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) != cudaSuccess) {
      (void)cudaGetLastError();
      continue;
    }
    auto& a = agg[r.name];
    a.first += 1;
    a.second += ms;
  }
  std::string out;
  char line[256];
  for (auto& kv : agg) {
    snprintf(line, sizeof(line), "%s %lld %.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
    out += line;
  }
  if (buf && cap) {
    size_t n = out.size() < cap - 1 ? out.size() : cap - 1;
    memcpy(buf, out.data(), n);
    buf[n] = 0;
  }
  return (long long)g_prof_records.size();
}

size_t conv3p_plan_bytes(const conv3p_geom_t* geom) {
  conv3p_plan_layout_t L;
  if (compute_layout(geom, &L)) return 0;
  return L.total_bytes;
}

int conv3p_plan_layout(const conv3p_geom_t* geom, conv3p_plan_layout_t* out) {
  if (!out) return CONV3P_ERR_INVALID_ARGUMENT;
  return compute_layout(geom, out);
}

int conv3p_plan_build_f32(const conv3p_geom_t* geom, const float* points, void* plan,
                          size_t plan_bytes, conv3p_stream_t stream) {
  PlanView v;
  int st = make_view(geom, plan, plan_bytes, &v);
  if (st) return st;
  if (!points && (long long)geom->B * geom->N > 0) return CONV3P_ERR_INVALID_ARGUMENT;
  plan_mark(plan, false);
  st = launch_cloud_sort(geom, points, v, stream);
  if (st) return st;
  return launch_neighbor_search(geom, v, stream);
}

int conv3p_plan_build_backward(const conv3p_geom_t* geom, const float* points, void* plan,
                               size_t plan_bytes, conv3p_stream_t stream) {
  PlanView v;
  int st = make_view(geom, plan, plan_bytes, &v);
  if (st) return st;
  if (!points && (long long)geom->B * geom->N > 0) return CONV3P_ERR_INVALID_ARGUMENT;
  st = launch_backward_lists(geom, points, v, stream);
  if (st) return st;
  C3P_CUDA(cudaMemsetAsync(v.header + H_HAS_BWD, 1, sizeof(long long), stream));
  plan_mark(plan, true);
  return CONV3P_OK;
}

int conv3p_plan_stats(const conv3p_geom_t* geom, const void* plan, conv3p_plan_stats_t* out,
                      conv3p_stream_t stream) {
  if (!out) return CONV3P_ERR_INVALID_ARGUMENT;
  conv3p_plan_layout_t L;
  int st = compute_layout(geom, &L);
  if (st) return st;
  if (!plan) return CONV3P_ERR_INVALID_ARGUMENT;
  long long h[H_SLOTS];
  C3P_CUDA(cudaMemcpyAsync(h, static_cast<const char*>(plan) + L.header, sizeof(h),
                           cudaMemcpyDeviceToHost, stream));
  C3P_CUDA(cudaStreamSynchronize(stream));
  out->total_pairs = h[H_CURSOR];
  out->backward_pairs = h[H_BWD_PAIRS];
  out->overflow = (h[H_OVERFLOW] != 0 || h[H_CURSOR] > geom->pair_capacity) ? 1 : 0;
  out->has_backward = h[H_HAS_BWD] != 0;
  return CONV3P_OK;
}

int conv3p_plan_publish_stats(const conv3p_geom_t* geom, const void* plan, long long* host_mapped16,
                              conv3p_stream_t stream) {
  if (!host_mapped16) return CONV3P_ERR_INVALID_ARGUMENT;
  conv3p_plan_layout_t L;
  int st = compute_layout(geom, &L);
  if (st) return st;
  if (!plan) return CONV3P_ERR_INVALID_ARGUMENT;
  k_publish_header<<<1, 32, 0, stream>>>(reinterpret_cast<const long long*>(static_cast<const char*>(plan) + L.header),
                                         host_mapped16);
  C3P_LAUNCH_CHECK("k_publish_header");
  return CONV3P_OK;
}

static size_t scratch_bytes_plain(const conv3p_geom_t* geom, int Cin, int Cout) {
  if (check_geom(geom) || check_channels(Cin, Cout)) return 0;
  // [weight panel images | work-item lists of the tensor-core gather kernels | split-K partials of grad_filter]
  size_t filt = backward_filter_scratch_bytes(geom, Cin, Cout);
  if (small_backward_filter_supported(Cin, Cout)) {
    const size_t t = backward_filter_small_scratch_bytes(Cin, Cout);
    if (t > filt) filt = t;
  }
  if (backward_filter2_supported(geom->N, geom->pair_capacity, Cin, Cout)) {
    const size_t t = backward_filter2_scratch_bytes(geom, Cin, Cout);
    if (t > filt) filt = t;
  }
  return weight_panel_bytes(Cin, Cout) + tc_items_bytes(geom, Cin, Cout) + align_up(filt) + 256;
}

// ---- channel padding: shapes the tensor-core kernels do not take as they are (36->13 of the segmentation network,
// 48 or 100 channels, ...) run on them with the channels zero-padded -- Cin to a multiple of 32, Cout to a multiple of 16
// (forward) or to 32 / 64 / 128 / 256 (backward) -- in copies held in the scratch buffer: padded input rows and
// weights contribute exact zeros, padded outputs are dropped.  Not when the padded product is more than 6x the real
// one; the tiny shapes (both counts <= 16) only for large batches, see below.  Engine flag 2048 switches it off (A/B
// timing).
static bool pad_channels(const conv3p_geom_t* g, int Cin, int Cout, bool backward, int* Pi, int* Po) {
  if (!engine_allows_tc() || engine_flag(2048)) return false;
  // Tiny shapes (both counts <= 16, the 9 -> 9 layers of both networks): measured on one B200, fwd+bwd step of one 9 -> 9
  // layer on the warp-per-point kernels against padded to 32 x 16 / 32 x 32 -- 262,144 points 2.95 -> 2.19 ms (4096 per
  // cloud), 6.34 -> 4.62 ms (16384 per cloud), 1.91 -> 1.75 ms (1024 per cloud); 65,536 points 0.794 -> 0.717 ms, but
  // nine more launches per layer, which a whole network at that size pays for on the host (segmentation network 4.69
  // -> 4.97 ms); 32,768 points 0.324 -> 0.407 ms.  So only from 128k points up (CONV3P_PAD_TINY_MIN_POINTS).
  static const long long tiny_min_points = [] {
    const char* e = getenv("CONV3P_PAD_TINY_MIN_POINTS");
    return e ? atoll(e) : 131072LL;
  }();
  if (Cin > 256 || Cout > 256) return false;
  const bool tiny = Cin <= 16 && Cout <= 16;
  if (tiny && (long long)g->B * g->N < tiny_min_points) return false;
  const int pi = (Cin + 31) / 32 * 32;
  const int po = !backward ? (Cout + 15) / 16 * 16 : Cout <= 32 ? 32 : Cout <= 64 ? 64 : Cout <= 128 ? 128 : 256;
  if (pi == Cin && po == Cout) return false;
  if (!tiny && (long long)pi * po > 6LL * Cin * Cout) return false;
  if (!backward) {
    if (Cin % 4 == 0 && Cout % 4 == 0 && forward_tc_supported(g->N, g->pair_capacity, Cin, Cout)) return false;
    if (!forward_tc_supported(g->N, g->pair_capacity, pi, po)) return false;
  } else {
    if (Cin % 4 == 0 && Cout % 4 == 0 && backward_input_tc_supported(g->N, g->pair_capacity, Cin, Cout) &&
        backward_filter2_supported(g->N, g->pair_capacity, Cin, Cout))
      return false;
    if (!backward_input_tc_supported(g->N, g->pair_capacity, pi, po) ||
        !backward_filter2_supported(g->N, g->pair_capacity, pi, po))
      return false;
  }
  *Pi = pi;
  *Po = po;
  return true;
}

struct PadLayout {     // byte offsets into the scratch buffer
  size_t a, b, c, w, gf, inner, total;
};
// forward: [x_pad | y_pad | W_pad | inner scratch]; backward: [g_pad | x_pad | gi_pad | W_pad | gf_pad | inner (+ G store)]
static PadLayout pad_layout(const conv3p_geom_t* g, int Pi, int Po, bool backward, bool with_g_store) {
  const size_t pts = (size_t)g->B * g->N;
  PadLayout L{};
  size_t o = 0;
  if (!backward) {
    L.a = o; o += align_up(pts * Pi * sizeof(float));
    L.b = o; o += align_up(pts * Po * sizeof(float));
    L.c = 0;
    L.w = o; o += align_up((size_t)C3P_NCELL * Pi * Po * sizeof(float));
    L.gf = 0;
  } else {
    L.a = o; o += align_up(pts * Po * sizeof(float));
    L.b = o; o += align_up(pts * Pi * sizeof(float));
    L.c = o; o += align_up(pts * Pi * sizeof(float));
    L.w = o; o += align_up((size_t)C3P_NCELL * Pi * Po * sizeof(float));
    L.gf = o; o += align_up((size_t)C3P_NCELL * Pi * Po * sizeof(float));
  }
  L.inner = o;
  o += scratch_bytes_plain(g, Pi, Po);
  if (backward && with_g_store) o += g_store_bytes(g, Pi, Po);
  L.total = o;
  return L;
}

size_t conv3p_scratch_bytes(const conv3p_geom_t* geom, int Cin, int Cout) {
  size_t n = scratch_bytes_plain(geom, Cin, Cout);
  if (!n) return 0;
  int pi, po;
  if (pad_channels(geom, Cin, Cout, false, &pi, &po)) n = std::max(n, pad_layout(geom, pi, po, false, false).total);
  if (pad_channels(geom, Cin, Cout, true, &pi, &po)) n = std::max(n, pad_layout(geom, pi, po, true, false).total);
  return n;
}

size_t conv3p_backward_scratch_bytes(const conv3p_geom_t* geom, int Cin, int Cout) {
  const size_t base = conv3p_scratch_bytes(geom, Cin, Cout);
  if (!base) return 0;
  size_t n = std::max(base, scratch_bytes_plain(geom, Cin, Cout) + g_store_bytes(geom, Cin, Cout));
  int pi, po;
  if (pad_channels(geom, Cin, Cout, true, &pi, &po)) n = std::max(n, pad_layout(geom, pi, po, true, true).total);
  return n;
}

// rows [pts][C] (row stride ls) -> [pts][P] with zeros in the padding, and back (row stride lo)
__global__ void k_pad_channels(const float* __restrict__ src, long long ls, int C, float* __restrict__ dst, int P,
                               long long pts) {
  const long long total = pts * P;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / P;
    const int c = (int)(e - r * P);
    dst[e] = c < C ? src[r * ls + c] : 0.f;
  }
}
__global__ void k_unpad_channels(const float* __restrict__ src, int P, float* __restrict__ dst, long long lo, int C,
                                 long long pts) {
  const long long total = pts * C;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / C;
    const int c = (int)(e - r * C);
    dst[r * lo + c] = src[r * P + c];
  }
}
// filter [27][Cin][Cout] <-> [27][Pi][Po]
__global__ void k_pad_filter(const float* __restrict__ src, int Cin, int Cout, float* __restrict__ dst, int Pi, int Po,
                             int to_padded) {
  const int total = C3P_NCELL * (to_padded ? Pi * Po : Cin * Cout);
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    if (to_padded) {
      const int c = e % Po, k = (e / Po) % Pi, f = e / (Po * Pi);
      dst[e] = (k < Cin && c < Cout) ? src[((size_t)f * Cin + k) * Cout + c] : 0.f;
    } else {
      const int c = e % Cout, k = (e / Cout) % Cin, f = e / (Cout * Cin);
      dst[e] = src[((size_t)f * Pi + k) * Po + c];
    }
  }
}
static int launch_pad(const float* src, long long ls, int C, float* dst, int P, long long pts, cudaStream_t stream) {
  {
    LaunchTimer timer_("k_pad_channels", stream);
    const long long blocks = std::min<long long>((pts * P + 255) / 256, (long long)sm_count() * 16);
    k_pad_channels<<<(unsigned)blocks, 256, 0, stream>>>(src, ls, C, dst, P, pts);
  }
  C3P_LAUNCH_CHECK("k_pad_channels");
  return CONV3P_OK;
}
static int launch_unpad(const float* src, int P, float* dst, long long lo, int C, long long pts, cudaStream_t stream) {
  {
    LaunchTimer timer_("k_unpad_channels", stream);
    const long long blocks = std::min<long long>((pts * C + 255) / 256, (long long)sm_count() * 16);
    k_unpad_channels<<<(unsigned)blocks, 256, 0, stream>>>(src, P, dst, lo, C, pts);
  }
  C3P_LAUNCH_CHECK("k_unpad_channels");
  return CONV3P_OK;
}
static int launch_pad_filter(const float* src, int Cin, int Cout, float* dst, int Pi, int Po, bool to_padded,
                             cudaStream_t stream) {
  {
    LaunchTimer timer_("k_pad_filter", stream);
    const int total = C3P_NCELL * (to_padded ? Pi * Po : Cin * Cout);
    k_pad_filter<<<(total + 255) / 256, 256, 0, stream>>>(src, Cin, Cout, dst, Pi, Po, to_padded ? 1 : 0);
  }
  C3P_LAUNCH_CHECK("k_pad_filter");
  return CONV3P_OK;
}

int conv3p_forward_ex_f32(const conv3p_geom_t* geom, const void* plan, const float* input,
                          long long input_row_stride, const float* filter, int Cin, int Cout, float* output,
                          long long output_row_stride, int activation, void* scratch, size_t scratch_bytes,
                          conv3p_stream_t stream) {
  int st = check_channels(Cin, Cout);
  if (st) return st;
  if (activation != CONV3P_ACT_NONE && activation != CONV3P_ACT_SELU) return CONV3P_ERR_INVALID_ARGUMENT;
  if ((input_row_stride && input_row_stride < Cin) || (output_row_stride && output_row_stride < Cout))
    return CONV3P_ERR_INVALID_ARGUMENT;
  PlanView v;
  st = make_view(geom, plan, conv3p_plan_bytes(geom), &v);
  if (st) return st;
  if ((long long)geom->B * geom->N == 0) return CONV3P_OK;
  if (!input || !filter || !output) return CONV3P_ERR_INVALID_ARGUMENT;
  RowIO io;
  io.src_stride = input_row_stride == Cin ? 0 : input_row_stride;
  io.out_stride = output_row_stride == Cout ? 0 : output_row_stride;
  io.activation = activation;
  const long long ls = io.src_stride ? io.src_stride : Cin, lo = io.out_stride ? io.out_stride : Cout;
  {
    int pi, po;
    if (pad_channels(geom, Cin, Cout, false, &pi, &po)) {
      const PadLayout P = pad_layout(geom, pi, po, false, false);
      if (scratch && scratch_bytes >= P.total) {     // (a smaller scratch buffer: the fp32 engines below)
        char* sp = static_cast<char*>(scratch);
        float* x_pad = reinterpret_cast<float*>(sp + P.a);
        float* y_pad = reinterpret_cast<float*>(sp + P.b);
        float* w_pad = reinterpret_cast<float*>(sp + P.w);
        const long long pts = (long long)geom->B * geom->N;
        if ((st = launch_pad(input, ls, Cin, x_pad, pi, pts, stream))) return st;
        if ((st = launch_pad_filter(filter, Cin, Cout, w_pad, pi, po, true, stream))) return st;
        st = conv3p_forward_ex_f32(geom, plan, x_pad, 0, w_pad, pi, po, y_pad, 0, activation, sp + P.inner,
                                   scratch_bytes - P.inner, stream);
        if (st) return st;
        return launch_unpad(y_pad, po, output, lo, Cout, pts, stream);
      }
    }
  }
  const bool rows_aligned = ls % 4 == 0 && lo % 4 == 0 && reinterpret_cast<uintptr_t>(input) % 16 == 0 &&
                            reinterpret_cast<uintptr_t>(output) % 16 == 0;
  // tensor cores need 16-byte aligned rows; other layouts fall back to the fp32 engines
  if (engine_allows_tc() && rows_aligned && forward_tc_supported(geom->N, geom->pair_capacity, Cin, Cout)) {
    if (!scratch || scratch_bytes < weight_panel_bytes(Cin, Cout)) return CONV3P_ERR_BUFFER_TOO_SMALL;
    return launch_forward_tc(geom, v, input, filter, Cin, Cout, output, scratch, scratch_bytes, stream, io);
  }
  if (engine_allows_small() && small_forward_supported(Cin, Cout))
    return launch_forward_small(geom, v, input, filter, Cin, Cout, output, stream, io);
  return launch_forward_simt(geom, v, input, filter, Cin, Cout, output, stream, io);
}

int conv3p_forward_f32(const conv3p_geom_t* geom, const void* plan, const float* input,
                       const float* filter, int Cin, int Cout, float* output, void* scratch,
                       size_t scratch_bytes, conv3p_stream_t stream) {
  return conv3p_forward_ex_f32(geom, plan, input, 0, filter, Cin, Cout, output, 0, CONV3P_ACT_NONE, scratch,
                               scratch_bytes, stream);
}

int conv3p_selu_backward_f32(const float* y, long long y_row_stride, const float* grad, long long grad_row_stride,
                             float* out, long long rows, int C, conv3p_stream_t stream) {
  if (rows < 0 || C < 1) return CONV3P_ERR_INVALID_ARGUMENT;
  if (rows == 0) return CONV3P_OK;
  if (!y || !grad || !out) return CONV3P_ERR_INVALID_ARGUMENT;
  const long long ys = y_row_stride ? y_row_stride : C, gs = grad_row_stride ? grad_row_stride : C;
  if (ys < C || gs < C) return CONV3P_ERR_INVALID_ARGUMENT;
  const long long total = rows * C;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  {
    LaunchTimer timer_("k_selu_backward", stream);
    k_selu_backward<<<(unsigned)blocks, 256, 0, stream>>>(y, ys, grad, gs, out, rows, C);
  }
  C3P_LAUNCH_CHECK("k_selu_backward");
  return CONV3P_OK;
}

int conv3p_backward_f32(const conv3p_geom_t* geom, const void* plan, const float* grad_output,
                        const float* input, const float* filter, int Cin, int Cout,
                        float* grad_input, float* grad_filter, void* scratch, size_t scratch_bytes,
                        conv3p_stream_t stream) {
  int st = check_channels(Cin, Cout);
  if (st) return st;
  PlanView v;
  st = make_view(geom, plan, conv3p_plan_bytes(geom), &v);
  if (st) return st;
  if (!grad_output && (long long)geom->B * geom->N > 0) return CONV3P_ERR_INVALID_ARGUMENT;
  if (plan_known_without_backward(plan)) return CONV3P_ERR_NO_BACKWARD_LISTS;
  if ((long long)geom->B * geom->N == 0) {
    if (grad_filter) C3P_CUDA(cudaMemsetAsync(grad_filter, 0, sizeof(float) * C3P_NCELL * (size_t)Cin * Cout, stream));
    return CONV3P_OK;
  }
  {
    int pi, po;
    if (pad_channels(geom, Cin, Cout, true, &pi, &po)) {
      const bool share = grad_input && grad_filter && !engine_flag(256);
      PadLayout P = pad_layout(geom, pi, po, true, share);
      if (share && !(scratch && scratch_bytes >= P.total)) P = pad_layout(geom, pi, po, true, false);
      if (scratch && scratch_bytes >= P.total) {     // (a smaller scratch buffer: the fp32 engines below)
        char* sp = static_cast<char*>(scratch);
        float* g_pad = reinterpret_cast<float*>(sp + P.a);
        float* x_pad = reinterpret_cast<float*>(sp + P.b);
        float* gi_pad = reinterpret_cast<float*>(sp + P.c);
        float* w_pad = reinterpret_cast<float*>(sp + P.w);
        float* gf_pad = reinterpret_cast<float*>(sp + P.gf);
        const long long pts = (long long)geom->B * geom->N;
        if (grad_input && !filter) return CONV3P_ERR_INVALID_ARGUMENT;
        if (grad_filter && !input) return CONV3P_ERR_INVALID_ARGUMENT;
        if ((st = launch_pad(grad_output, Cout, Cout, g_pad, po, pts, stream))) return st;
        if (grad_filter && (st = launch_pad(input, Cin, Cin, x_pad, pi, pts, stream))) return st;
        if (grad_input && (st = launch_pad_filter(filter, Cin, Cout, w_pad, pi, po, true, stream))) return st;
        st = conv3p_backward_f32(geom, plan, g_pad, x_pad, w_pad, pi, po, grad_input ? gi_pad : nullptr,
                                 grad_filter ? gf_pad : nullptr, sp + P.inner, scratch_bytes - P.inner, stream);
        if (st) return st;
        if (grad_input && (st = launch_unpad(gi_pad, pi, grad_input, Cin, Cin, pts, stream))) return st;
        if (grad_filter && (st = launch_pad_filter(gf_pad, Cin, Cout, grad_filter, pi, po, false, stream))) return st;
        return CONV3P_OK;
      }
    }
  }
  // Engine per gradient.  The tensor-core kernels use 16-byte vector accesses; other layouts run on the fp32 engines.
  const bool aligned = reinterpret_cast<uintptr_t>(grad_output) % 16 == 0 && Cin % 4 == 0 && Cout % 4 == 0 &&
                       reinterpret_cast<uintptr_t>(grad_input) % 16 == 0 && reinterpret_cast<uintptr_t>(input) % 16 == 0;
  const bool gi_tc = grad_input && engine_allows_tc() && aligned &&
                     backward_input_tc_supported(geom->N, geom->pair_capacity, Cin, Cout);
  const bool gf_tc = grad_filter && engine_allows_tc() && aligned &&
                     backward_filter2_supported(geom->N, geom->pair_capacity, Cin, Cout);
  // Both gradients on the tensor-core kernels: the grad_input kernel leaves its aggregated rows in the G store (tail of
  // the scratch buffer) for the grad_filter kernel -- only when BOTH launches below really take that pair of kernels.
  // Engine bit 256 switches the sharing off (A/B timing).
  float* g_store = nullptr;
  if (gi_tc && gf_tc && !engine_flag(256)) {
    const size_t gsb = g_store_bytes(geom, Cin, Cout);
    const size_t base = scratch_bytes_plain(geom, Cin, Cout);
    if (gsb && scratch && scratch_bytes >= base + gsb) g_store = reinterpret_cast<float*>(static_cast<char*>(scratch) + base);
  }
  // scratch layout: [weight panel images | 128-row work-item lists | grad_filter scratch (its 64-row lists first)]
  const size_t wpb = weight_panel_bytes(Cin, Cout) + tc_items_bytes(geom, Cin, Cout);
  // Both gradients on tensor cores with 64-point tiles in the weight gradient: the pre-pass of the grad_input
  // kernel also emits the weight-gradient kernel's lists (one walk over the count table instead of two).
  bool items_shared = false;
  GroupItems gi64{};
  if (gi_tc && gf_tc && backward_filter2_tile_rows(geom->N, geom->pair_capacity, Cin, Cout) == 64 && scratch &&
      scratch_bytes >= wpb + backward_filter2_scratch_bytes(geom, Cin, Cout)) {
    gi64 = carve_group_items(static_cast<char*>(scratch) + wpb, (long long)geom->B * geom->N, 64);
    items_shared = true;
  }
  if (grad_input) {
    if (!filter) return CONV3P_ERR_INVALID_ARGUMENT;
    if (gi_tc) {
      if (!scratch || scratch_bytes < weight_panel_bytes(Cin, Cout)) return CONV3P_ERR_BUFFER_TOO_SMALL;
      st = launch_backward_input_tc(geom, v, grad_output, filter, Cin, Cout, grad_input, scratch,
                                    scratch_bytes, stream, g_store, items_shared ? &gi64 : nullptr);
    } else if (engine_allows_small() && small_backward_input_supported(Cin, Cout)) {
      st = launch_backward_input_small(geom, v, grad_output, filter, Cin, Cout, grad_input, stream);
    } else {
      st = launch_backward_input_simt(geom, v, grad_output, filter, Cin, Cout, grad_input, stream);
    }
    if (st) return st;
  }
  if (grad_filter) {
    if (!input) return CONV3P_ERR_INVALID_ARGUMENT;
    if (!scratch || scratch_bytes < wpb) return CONV3P_ERR_BUFFER_TOO_SMALL;
    if (gf_tc)
      st = launch_backward_filter2(geom, v, grad_output, input, Cin, Cout, grad_filter,
                                   static_cast<char*>(scratch) + wpb, scratch_bytes - wpb, stream, g_store, items_shared);
    else if (engine_allows_small() && small_backward_filter_supported(Cin, Cout))
      st = launch_backward_filter_small(geom, v, grad_output, input, Cin, Cout, grad_filter,
                                        static_cast<char*>(scratch) + wpb, scratch_bytes - wpb, stream);
    else
      st = launch_backward_filter_simt(geom, v, grad_output, input, Cin, Cout, grad_filter,
                                       static_cast<char*>(scratch) + wpb, scratch_bytes - wpb, stream);
    if (st) return st;
  }
  return CONV3P_OK;
}

// ---- one-shot calls ------------------------------------------------------------------------------

size_t conv3p_op_workspace_bytes(const conv3p_geom_t* geom, int Cin, int Cout) {
  size_t p = conv3p_plan_bytes(geom);
  if (!p) return 0;
  return p + conv3p_scratch_bytes(geom, Cin, Cout);
}

size_t conv3p_op_backward_workspace_bytes(const conv3p_geom_t* geom, int Cin, int Cout) {
  size_t p = conv3p_plan_bytes(geom);
  if (!p) return 0;
  return p + conv3p_backward_scratch_bytes(geom, Cin, Cout);
}

// 3x3x3 (every reference model) runs on the tuned engines; other shapes on the general path (generic_filter.cu)
static bool is_333(const int d[3]) { return d[0] == 3 && d[1] == 3 && d[2] == 3; }
static int check_filter_dims(const int filter_dims[3]) {
  if (!filter_dims) return CONV3P_ERR_INVALID_ARGUMENT;
  if (filter_dims[0] < 1 || filter_dims[1] < 1 || filter_dims[2] < 1) return CONV3P_ERR_INVALID_ARGUMENT;
  if (!is_333(filter_dims) && !generic_filter_supported(filter_dims)) return CONV3P_ERR_UNSUPPORTED;
  return CONV3P_OK;
}

size_t conv3p_op_workspace_bytes_ex(const conv3p_geom_t* geom, const int filter_dims[3], int Cin, int Cout,
                                    int backward) {
  if (!filter_dims || check_filter_dims(filter_dims)) return 0;
  if (is_333(filter_dims))
    return backward ? conv3p_op_backward_workspace_bytes(geom, Cin, Cout) : conv3p_op_workspace_bytes(geom, Cin, Cout);
  return generic_workspace_bytes(geom, filter_dims, Cin, Cout);
}

// ---- T = double (register_op.cpp:45, 64; tf_conv3p_atrous.cpp:516, 727): one-shot calls on the general path --------
static int check_f64(const int filter_dims[3], const int stride_xyz[3], double voxel_size, int B, int N, int Cin,
                     int Cout, long long pair_capacity, conv3p_geom_t* g) {
  if (!filter_dims || !stride_xyz) return CONV3P_ERR_INVALID_ARGUMENT;
  if (filter_dims[0] < 1 || filter_dims[1] < 1 || filter_dims[2] < 1) return CONV3P_ERR_INVALID_ARGUMENT;
  if (!generic_filter_supported(filter_dims)) return CONV3P_ERR_UNSUPPORTED;
  if (!(voxel_size > 0.0) || !((float)voxel_size > 0.f)) return CONV3P_ERR_INVALID_ARGUMENT;
  *g = make_geom(B, N, stride_xyz, (float)voxel_size, pair_capacity);   // float cell size of the candidate grid only
  int st = check_geom(g);
  if (st) return st;
  return check_channels(Cin, Cout);
}

size_t conv3p_op_workspace_bytes_f64(const conv3p_geom_t* geom, const int filter_dims[3], int Cin, int Cout) {
  if (!filter_dims) return 0;
  return generic_workspace_bytes(geom, filter_dims, Cin, Cout, 8);
}

int conv3p_op_forward_f64(const double* points, const double* input, const double* filter, const int filter_dims[3],
                          const int stride_xyz[3], double voxel_size, int B, int N, int Cin, int Cout,
                          long long pair_capacity, double* output, void* workspace, size_t workspace_bytes,
                          conv3p_stream_t stream) {
  conv3p_geom_t g;
  const int st = check_f64(filter_dims, stride_xyz, voxel_size, B, N, Cin, Cout, pair_capacity, &g);
  if (st) return st;
  return generic_forward_f64(&g, filter_dims, voxel_size, points, input, filter, Cin, Cout, output, workspace,
                             workspace_bytes, stream);
}

int conv3p_op_backward_f64(const double* grad_output, const double* points, const double* input, const double* filter,
                           const int filter_dims[3], const int stride_xyz[3], double voxel_size, int B, int N, int Cin,
                           int Cout, long long pair_capacity, double* grad_input, double* grad_filter, void* workspace,
                           size_t workspace_bytes, conv3p_stream_t stream) {
  conv3p_geom_t g;
  const int st = check_f64(filter_dims, stride_xyz, voxel_size, B, N, Cin, Cout, pair_capacity, &g);
  if (st) return st;
  return generic_backward_f64(&g, filter_dims, voxel_size, grad_output, points, input, filter, Cin, Cout, grad_input,
                              grad_filter, workspace, workspace_bytes, stream);
}

int conv3p_op_forward_f32(const float* points, const float* input, const float* filter,
                          const int filter_dims[3], const int stride_xyz[3], float voxel_size, int B,
                          int N, int Cin, int Cout, long long pair_capacity, float* output,
                          void* workspace, size_t workspace_bytes, conv3p_stream_t stream) {
  int st = check_filter_dims(filter_dims);
  if (st) return st;
  if (!stride_xyz) return CONV3P_ERR_INVALID_ARGUMENT;
  conv3p_geom_t g = make_geom(B, N, stride_xyz, voxel_size, pair_capacity);
  st = check_geom(&g);
  if (st) return st;
  if (!is_333(filter_dims)) {
    st = check_channels(Cin, Cout);
    if (st) return st;
    return generic_forward(&g, filter_dims, points, input, filter, Cin, Cout, output, workspace, workspace_bytes, stream);
  }
  const size_t pb = conv3p_plan_bytes(&g);
  if (!workspace || workspace_bytes < conv3p_op_workspace_bytes(&g, Cin, Cout))
    return CONV3P_ERR_BUFFER_TOO_SMALL;
  st = conv3p_plan_build_f32(&g, points, workspace, pb, stream);
  if (st) return st;
  return conv3p_forward_f32(&g, workspace, input, filter, Cin, Cout, output,
                            static_cast<char*>(workspace) + pb, workspace_bytes - pb, stream);
}

int conv3p_op_backward_f32(const float* grad_output, const float* points, const float* input,
                           const float* filter, const int filter_dims[3], const int stride_xyz[3],
                           float voxel_size, int B, int N, int Cin, int Cout,
                           long long pair_capacity, float* grad_input, float* grad_filter,
                           void* workspace, size_t workspace_bytes, conv3p_stream_t stream) {
  int st = check_filter_dims(filter_dims);
  if (st) return st;
  if (!stride_xyz) return CONV3P_ERR_INVALID_ARGUMENT;
  conv3p_geom_t g = make_geom(B, N, stride_xyz, voxel_size, pair_capacity);
  st = check_geom(&g);
  if (st) return st;
  if (!is_333(filter_dims)) {
    st = check_channels(Cin, Cout);
    if (st) return st;
    return generic_backward(&g, filter_dims, grad_output, points, input, filter, Cin, Cout, grad_input, grad_filter,
                            workspace, workspace_bytes, stream);
  }
  const size_t pb = conv3p_plan_bytes(&g);
  if (!workspace || workspace_bytes < conv3p_op_workspace_bytes(&g, Cin, Cout))
    return CONV3P_ERR_BUFFER_TOO_SMALL;
  st = conv3p_plan_build_f32(&g, points, workspace, pb, stream);
  if (st) return st;
  st = conv3p_plan_build_backward(&g, points, workspace, pb, stream);
  if (st) return st;
  return conv3p_backward_f32(&g, workspace, grad_output, input, filter, Cin, Cout, grad_input,
                             grad_filter, static_cast<char*>(workspace) + pb, workspace_bytes - pb,
                             stream);
}

// ---- host-buffer calls -----------------------------------------------------------------------------

static size_t host_io_bytes(const conv3p_geom_t* g, int Cin, int Cout) {
  const size_t pts = (size_t)g->B * g->N;
  const size_t nW = (size_t)C3P_NCELL * Cin * Cout;
  // points, input, filter, out/grad_out, grad_input, grad_filter
  return align_up(pts * 3 * 4) + align_up(pts * Cin * 4) + align_up(nW * 4) + align_up(pts * Cout * 4) +
         align_up(pts * Cin * 4) + align_up(nW * 4);
}

size_t conv3p_host_workspace_bytes(const conv3p_geom_t* geom, int Cin, int Cout) {
  size_t w = conv3p_op_workspace_bytes(geom, Cin, Cout);
  if (!w) return 0;
  return w + host_io_bytes(geom, Cin, Cout);
}

struct HostIO {
  float *points, *input, *filter, *outg, *grad_input, *grad_filter;
  char* op_ws;
  size_t op_ws_bytes;
};

static HostIO carve(const conv3p_geom_t* g, int Cin, int Cout, void* ws, size_t ws_bytes) {
  const size_t pts = (size_t)g->B * g->N;
  const size_t nW = (size_t)C3P_NCELL * Cin * Cout;
  char* p = static_cast<char*>(ws);
  HostIO io;
  io.points = reinterpret_cast<float*>(p); p += align_up(pts * 3 * 4);
  io.input = reinterpret_cast<float*>(p); p += align_up(pts * Cin * 4);
  io.filter = reinterpret_cast<float*>(p); p += align_up(nW * 4);
  io.outg = reinterpret_cast<float*>(p); p += align_up(pts * Cout * 4);
  io.grad_input = reinterpret_cast<float*>(p); p += align_up(pts * Cin * 4);
  io.grad_filter = reinterpret_cast<float*>(p); p += align_up(nW * 4);
  io.op_ws = p;
  io.op_ws_bytes = ws_bytes - (size_t)(p - static_cast<char*>(ws));
  return io;
}

int conv3p_host_forward_f32(const float* h_points, const float* h_input, const float* h_filter,
                            const int stride_xyz[3], float voxel_size, int B, int N, int Cin,
                            int Cout, long long pair_capacity, float* h_output, void* workspace,
                            size_t workspace_bytes, conv3p_stream_t stream) {
  if (!stride_xyz) return CONV3P_ERR_INVALID_ARGUMENT;
  conv3p_geom_t g = make_geom(B, N, stride_xyz, voxel_size, pair_capacity);
  int st = check_geom(&g);
  if (st) return st;
  st = check_channels(Cin, Cout);
  if (st) return st;
  if (!workspace || workspace_bytes < conv3p_host_workspace_bytes(&g, Cin, Cout))
    return CONV3P_ERR_BUFFER_TOO_SMALL;
  const size_t pts = (size_t)B * N, nW = (size_t)C3P_NCELL * Cin * Cout;
  HostIO io = carve(&g, Cin, Cout, workspace, workspace_bytes);
  C3P_CUDA(cudaMemcpyAsync(io.points, h_points, pts * 3 * 4, cudaMemcpyHostToDevice, stream));
  C3P_CUDA(cudaMemcpyAsync(io.input, h_input, pts * Cin * 4, cudaMemcpyHostToDevice, stream));
  C3P_CUDA(cudaMemcpyAsync(io.filter, h_filter, nW * 4, cudaMemcpyHostToDevice, stream));
  const int fd[3] = {3, 3, 3};
  st = conv3p_op_forward_f32(io.points, io.input, io.filter, fd, stride_xyz, voxel_size, B, N, Cin,
                             Cout, pair_capacity, io.outg, io.op_ws, io.op_ws_bytes, stream);
  if (st) return st;
  C3P_CUDA(cudaMemcpyAsync(h_output, io.outg, pts * Cout * 4, cudaMemcpyDeviceToHost, stream));
  C3P_CUDA(cudaStreamSynchronize(stream));
  conv3p_plan_stats_t stats;
  st = conv3p_plan_stats(&g, io.op_ws, &stats, stream);
  if (st) return st;
  return stats.overflow ? CONV3P_ERR_PAIR_OVERFLOW : CONV3P_OK;
}

int conv3p_host_backward_f32(const float* h_grad_output, const float* h_points,
                             const float* h_input, const float* h_filter, const int stride_xyz[3],
                             float voxel_size, int B, int N, int Cin, int Cout,
                             long long pair_capacity, float* h_grad_input, float* h_grad_filter,
                             void* workspace, size_t workspace_bytes, conv3p_stream_t stream) {
  if (!stride_xyz) return CONV3P_ERR_INVALID_ARGUMENT;
  conv3p_geom_t g = make_geom(B, N, stride_xyz, voxel_size, pair_capacity);
  int st = check_geom(&g);
  if (st) return st;
  st = check_channels(Cin, Cout);
  if (st) return st;
  if (!workspace || workspace_bytes < conv3p_host_workspace_bytes(&g, Cin, Cout))
    return CONV3P_ERR_BUFFER_TOO_SMALL;
  const size_t pts = (size_t)B * N, nW = (size_t)C3P_NCELL * Cin * Cout;
  HostIO io = carve(&g, Cin, Cout, workspace, workspace_bytes);
  C3P_CUDA(cudaMemcpyAsync(io.points, h_points, pts * 3 * 4, cudaMemcpyHostToDevice, stream));
  C3P_CUDA(cudaMemcpyAsync(io.input, h_input, pts * Cin * 4, cudaMemcpyHostToDevice, stream));
  C3P_CUDA(cudaMemcpyAsync(io.filter, h_filter, nW * 4, cudaMemcpyHostToDevice, stream));
  C3P_CUDA(cudaMemcpyAsync(io.outg, h_grad_output, pts * Cout * 4, cudaMemcpyHostToDevice, stream));
  const int fd[3] = {3, 3, 3};
  st = conv3p_op_backward_f32(io.outg, io.points, io.input, io.filter, fd, stride_xyz, voxel_size, B,
                              N, Cin, Cout, pair_capacity, h_grad_input ? io.grad_input : nullptr,
                              h_grad_filter ? io.grad_filter : nullptr, io.op_ws, io.op_ws_bytes,
                              stream);
  if (st) return st;
  if (h_grad_input)
    C3P_CUDA(cudaMemcpyAsync(h_grad_input, io.grad_input, pts * Cin * 4, cudaMemcpyDeviceToHost, stream));
  if (h_grad_filter)
    C3P_CUDA(cudaMemcpyAsync(h_grad_filter, io.grad_filter, nW * 4, cudaMemcpyDeviceToHost, stream));
  C3P_CUDA(cudaStreamSynchronize(stream));
  conv3p_plan_stats_t stats;
  st = conv3p_plan_stats(&g, io.op_ws, &stats, stream);
  if (st) return st;
  return stats.overflow ? CONV3P_ERR_PAIR_OVERFLOW : CONV3P_OK;
}

}  // extern "C"
