// forward_tc.cu -- Conv3p forward for GEMM-class channel counts on the 5th-gen tensor cores.
//
// For a tile of T*128 voxel-sorted points the (27*Cin) x Cout contraction of SURVEY section 0 is run as
// T accumulators of 128 x Cout fp32 in TMEM; K is walked as (kernel cell f, point sub-tile t, 32-channel
// panel kc).  The aggregated operand A_f (per-cell mean of the neighbour rows, tf_conv3p_atrous.cpp:
// 480-494 regrouped as mean-then-contract) is never written to HBM: producer warps gather neighbour
// rows with 16-byte loads (quarter-warp per point = one 128-byte row segment per instruction), reduce
// them in registers, split the mean into TF32 hi/lo parts and store them straight into the
// 128B-swizzled K-major operand panels of a shared-memory ring.  The weights are pre-split, pre-
// transposed and pre-swizzled once per call into panel images, so one thread streams them with
// 32 KB bulk async copies (UBLKCP) signalled on mbarriers.  One thread issues tcgen05.mma
// (kind::tf32, M=128, N=Cout, K=8): D += A_hi*W_hi + A_lo*W_hi + A_hi*W_lo, i.e. 3xTF32 -- every
// product keeps ~21 mantissa bits, accumulation is fp32 in TMEM, which holds the stated fp32-class
// tolerance (tests/test_gpu_tc.py, tests/test_gpu_parity.py).  tcgen05.commit releases ring slots; the
// epilogue reads TMEM with tcgen05.ld and writes each output row exactly once.
//
// Warp roles: warps [0, NPW) producers (also the epilogue), warp NPW = MMA issuer, warp NPW+1 =
// weight loader + TMEM allocator.
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_gather.cuh"

namespace c3p {

using namespace tc;

constexpr int FT_NPW = 16;                 // producer warps
constexpr int FT_NQ = FT_NPW * 4;          // quarter-warps: one point-row segment each
constexpr int FT_THREADS = (FT_NPW + 2) * 32;
constexpr int FT_NAS = 3;                  // A ring stages
constexpr int FT_A_STAGE = 2 * 128 * PANEL_ROW_BYTES;  // hi + lo panels of 128 rows
constexpr int FT_MAX_NKC = 4;

// One kernel serves the forward pass (src = input, lists = forward lists, per-cell mean) and the input
// gradient (src = grad_out, lists = backward lists with per-entry weights 1/count(ii,f'), panels = W
// itself): out[p, n] = sum_f sum_k A_f[p, k] * Wpanel_f[n, k].

struct FTArgs {
  const float* src;         // gathered rows [B*N, Csrc]
  const unsigned char* wp;  // weight panel images [27][Csrc/32][hi,lo][Nout][128 B]
  float* out;               // [B*N, Nout]
  const int* cnt;           // [B*N, 27] group sizes of the lists
  const long long* begin;
  const int* len;
  const int* rows;
  const float* weights;     // per-entry weights (WEIGHTED) or nullptr (per-cell mean)
  const float4* sorted_xyzi;
  long long total_points, capacity;
  int N, Csrc, Nout, nkb, T, NWS;  // nkb = K batches of NKC 32-channel panels
  int debug;
};

// weights [27][Cin][Cout] -> panel images: for (f, kc, hl) a [Cout rows x 32 k] K-major swizzled panel
__global__ void k_prep_weight_panels(const float* __restrict__ filter, unsigned char* __restrict__ wp,
                                     int Cin, int Cout, int transposed_out) {
  // transposed_out == 0: rows = Cout (n = c), K = Cin (forward B operand, W^T)
  // transposed_out == 1: rows = Cin  (n = k), K = Cout (input-gradient B operand, W)
  const int R = transposed_out ? Cin : Cout;   // panel rows
  const int KD = transposed_out ? Cout : Cin;  // contraction length
  const int nkc = KD / PANEL_K;
  const long long total = (long long)C3P_NCELL * nkc * R * PANEL_K;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(e % PANEL_K);
    const int r = (int)((e / PANEL_K) % R);
    const int kc = (int)((e / ((long long)PANEL_K * R)) % nkc);
    const int f = (int)(e / ((long long)PANEL_K * R * nkc));
    const int kk = kc * PANEL_K + k;
    const float w = transposed_out ? filter[((size_t)f * Cin + r) * Cout + kk]
                                   : filter[((size_t)f * Cin + kk) * Cout + r];
    const float h = tf32_hi(w);
    unsigned char* base = wp + ((size_t)(f * nkc + kc) * 2) * R * PANEL_ROW_BYTES;
    *reinterpret_cast<float*>(base + panel_offset(r, k)) = h;
    *reinterpret_cast<float*>(base + (size_t)R * PANEL_ROW_BYTES + panel_offset(r, k)) = w - h;
  }
}


template <int NKC, bool WEIGHTED>
__global__ void __launch_bounds__(FT_THREADS, 1) k_gather_mma_tc(const FTArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int T = a.T, Cout = a.Nout, NWS = a.NWS;
  const int PT = T * 128;
  const uint32_t w_slot_bytes = 2u * Cout * PANEL_ROW_BYTES;
  unsigned char* w_base = smem;                                  // NWS slots
  unsigned char* a_base = w_base + (size_t)NWS * w_slot_bytes;   // FT_NAS stages
  unsigned char* tab = a_base + (size_t)FT_NAS * FT_A_STAGE;
  uint32_t* beg = reinterpret_cast<uint32_t*>(tab);              // [PT] list start (capacity < 2^32)
  int* rowid = reinterpret_cast<int*>(beg + PT);                 // [PT]
  // exclusive prefix of the per-cell counts: list offset of cell f = pre16[f], members = pre16[f+1]-pre16[f]
  // (immutable, so any producer may serve any point; needs K_i <= 65535, guaranteed by N <= 65535)
  uint16_t* pre16 = reinterpret_cast<uint16_t*>(rowid + PT);     // [PT][28]
  __shared__ uint64_t a_full[FT_NAS], a_empty[FT_NAS], w_full[FT_MAX_NKC + 1], w_empty[FT_MAX_NKC + 1],
      acc_full;
  __shared__ uint32_t tmem_slot;
  __shared__ unsigned active[C3P_NCELL];  // bit t: sub-tile t has members in cell f

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long s0 = (long long)blockIdx.x * PT;

  if (tid < C3P_NCELL) active[tid] = 0;
  if (warp == FT_NPW && lane == 0) {
    for (int i = 0; i < FT_NAS; ++i) {
      mbar_init(&a_full[i], FT_NPW);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < NWS; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    mbar_init(&acc_full, 1);
    mbar_fence_init();
  }
  if (warp == FT_NPW + 1) tmem_alloc(&tmem_slot, 512);
  __syncthreads();
  for (int p = tid; p < PT; p += FT_THREADS) {
    const long long s = s0 + p;
    int row = -1;
    long long bg = 0;
    bool ok = false;
    if (s < a.total_points) {
      const int b = (int)(s / a.N);
      row = b * a.N + __float_as_int(a.sorted_xyzi[s].w);
      bg = a.begin[row];
      ok = bg + a.len[row] <= a.capacity;
      if (!ok) row = -2 - row;  // incomplete list: poison this point's output
    }
    const int t = p >> 7;
    int run = 0;
    for (int f = 0; f < C3P_NCELL; ++f) {
      const int c = ok ? __ldg(a.cnt + (size_t)row * C3P_NCELL + f) : 0;
      pre16[p * 28 + f] = (uint16_t)run;
      run += c;
      if (c) atomicOr(&active[f], 1u << t);
    }
    pre16[p * 28 + 27] = (uint16_t)run;
    beg[p] = (uint32_t)bg;
    rowid[p] = row;
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;

  if (warp < FT_NPW) {
    // =========================== producers: gather -> mean -> hi/lo -> operand panels ===============
    // Quarter-warp q serves rows (q + 32*g) mod NQ (+ NQ) of group g; lane l8 owns one 16-byte chunk of the
    // row segment.  The loop is software-pipelined over (group, repetition) slots: the list ids of the next
    // slot are loaded while the rows of the current slot are in flight (select on the ADDRESS, so nothing
    // waits on the prefetch), and up to eight members x NKC panels of a row are in flight at once.
    const int q = warp * 4 + (lane >> 3);
    const int l8 = lane & 7;
    struct Slot {
      int p;          // row in the sub-tile, or -1
      GatherSlot gs;  // its cell's list: members, position, prefetched ids / weights
    };
    int gf = -1, gkb = a.nkb - 1, gt = T - 1;   // group iterator (cell, K batch, sub-tile)
    unsigned gact = 0;
    auto advance = [&](int& f_, int& kb_, int& t_, unsigned& act_) -> bool {
      for (;;) {
        if (++t_ >= T) {
          t_ = 0;
          if (++kb_ >= a.nkb) {
            kb_ = 0;
            do {
              if (++f_ >= C3P_NCELL) return false;
              act_ = active[f_];
            } while (!act_);
          }
        }
        if ((act_ >> t_) & 1u) return true;
      }
    };
    auto fetch = [&](int f_, int t_, int g_, int rep) -> Slot {
      Slot d;
      const int p = (q + g_ * 32) % FT_NQ + rep * FT_NQ;
      d.p = p < 128 ? p : -1;
      const int pt = t_ * 128 + (p < 128 ? p : 0);
      const int off = pre16[pt * 28 + f_];
      const int n = p < 128 ? (int)pre16[pt * 28 + f_ + 1] - off : 0;
      d.gs = fetch_slot<WEIGHTED>(n, beg[pt] + (uint32_t)off, a.rows, a.weights, l8);
      return d;
    };
    auto gather = [&](const Slot& d, int col, float4 (&acc)[NKC]) {
      gather_slot<NKC, WEIGHTED>(acc, d.gs, a.src, a.Csrc, col, a.rows, a.weights, l8);
    };
    auto store = [&](const Slot& d, int s, const float4 (&acc)[NKC]) {
      if (d.p >= 0) {
        const uint32_t o = panel_chunk_offset(d.p, l8);
#pragma unroll
        for (int kc = 0; kc < NKC; ++kc)
          store_split(a_base + (size_t)((s + kc) % FT_NAS) * FT_A_STAGE + o, 128 * PANEL_ROW_BYTES, acc[kc]);
      }
    };
    const bool two_reps = FT_NQ < 128;
    int g = 0, s = 0;
    bool have = advance(gf, gkb, gt, gact);
    Slot d0;
    if (have) d0 = fetch(gf, gt, g, 0);
#define C3P_PHASE(i) do { } while (0)  /* phase timers moved to gather_mma2.cu */
    while (have) {
      const int col = gkb * NKC * PANEL_K;
      Slot d1;
      if (two_reps) d1 = fetch(gf, gt, g, 1);       // ids of the second row, behind nothing yet
      float4 acc[NKC];
      C3P_PHASE(0);
      gather(d0, col, acc);
      C3P_PHASE(1);
      // ring slots of this group must have been drained by the tensor core
#pragma unroll
      for (int kc = 0; kc < NKC; ++kc) {
        const int st = s + kc, use = st / FT_NAS;
        if (use >= 1) mbar_wait(&a_empty[st - use * FT_NAS], (uint32_t)((use - 1) & 1));
      }
      C3P_PHASE(2);
      store(d0, s, acc);
      C3P_PHASE(3);
      // next group's first row: fetch its ids now, they arrive while the second row is gathered
      int nf = gf, nkb = gkb, nt = gt;
      unsigned nact = gact;
      const bool have_next = advance(nf, nkb, nt, nact);
      Slot dn;
      if (have_next) dn = fetch(nf, nt, g + 1, 0);
      C3P_PHASE(4);
      if (two_reps && __any_sync(C3P_FULL_MASK, d1.p >= 0)) {
        gather(d1, col, acc);
        store(d1, s, acc);
      }
      C3P_PHASE(5);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
#pragma unroll
        for (int kc = 0; kc < NKC; ++kc) mbar_arrive(&a_full[(s + kc) % FT_NAS]);
      }
      s += NKC;
      ++g;
      gf = nf; gkb = nkb; gt = nt; gact = nact;
      have = have_next;
      d0 = dn;
      C3P_PHASE(6);
    }
    // =========================== epilogue: TMEM -> registers -> global ==============================
    mbar_wait(&acc_full, 0);
    tc_fence_after_sync();
    if (warp < 4 * T) {
      const int t = warp >> 2, sub = warp & 3;
      bool any = false;
      for (int f = 0; f < C3P_NCELL; ++f) any |= ((active[f] >> t) & 1u) != 0;
      const int pt = t * 128 + sub * 32 + lane;
      int row = rowid[pt];
      const bool poison = row < -1;
      if (poison) row = -2 - row;
      for (int c0 = 0; c0 < Cout; c0 += 32) {
        float v[32];
        if (any) {
          tmem_ld_32x32(tmem + ((uint32_t)(sub * 32) << 16) + (uint32_t)(t * Cout + c0), v);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
        }
        if (row >= 0) {
          float* o = a.out + (size_t)row * Cout + c0;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (c0 + j < Cout) {
              float4 w4 = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
              if (poison) w4 = make_float4(__int_as_float(0x7fc00000), __int_as_float(0x7fc00000),
                                           __int_as_float(0x7fc00000), __int_as_float(0x7fc00000));
              *reinterpret_cast<float4*>(o + j) = w4;
            }
          }
        }
      }
    }
  } else if (warp == FT_NPW) {
    // =========================== MMA issuer (one thread) ============================================
    // No divisions in this loop: ring positions are tracked incrementally and descriptors are derived
    // from per-ring bases with one add.
    if (lane == 0) {
      const uint32_t idesc = make_idesc_tf32(128, Cout);
      const uint64_t a_desc0 = make_smem_desc(smem_u32(a_base)), w_desc0 = make_smem_desc(smem_u32(w_base));
      const uint64_t a_lo_off = (uint64_t)((128 * PANEL_ROW_BYTES) >> 4);
      const uint64_t w_lo_off = (uint64_t)(((uint32_t)Cout * PANEL_ROW_BYTES) >> 4);
      const uint64_t a_step = (uint64_t)(FT_A_STAGE >> 4), w_step = (uint64_t)(w_slot_bytes >> 4);
      unsigned started = 0;
      int aslot = 0, wslot0 = -NKC;
      uint32_t aphase = 0, wphase0 = 0;
      for (int f = 0; f < C3P_NCELL; ++f) {
        const unsigned act = active[f];
        if (!act) continue;
        const int t_first = __ffs(act) - 1, t_last = 31 - __clz(act);
        for (int kb = 0; kb < a.nkb; ++kb) {
          wslot0 += NKC;
          if (wslot0 >= NWS) { wslot0 -= NWS; wphase0 ^= 1u; }
          for (int t = 0; t < T; ++t) {
            if (!((act >> t) & 1u)) continue;
            const uint32_t d = tmem + (uint32_t)(t * Cout);
            uint32_t acc_flag = (started >> t) & 1u;
            started |= 1u << t;
#pragma unroll
            for (int kc = 0; kc < NKC; ++kc) {
              int wslot = wslot0 + kc;
              uint32_t wphase = wphase0;
              if (wslot >= NWS) { wslot -= NWS; wphase ^= 1u; }
              if (t == t_first) mbar_wait(&w_full[wslot], wphase);
              mbar_wait(&a_full[aslot], aphase);
              tc_fence_after_sync();
              const uint64_t dah = a_desc0 + (uint64_t)aslot * a_step, dal = dah + a_lo_off;
              const uint64_t dwh = w_desc0 + (uint64_t)wslot * w_step, dwl = dwh + w_lo_off;
#pragma unroll
              for (int ks = 0; ks < PANEL_K / UMMA_K; ++ks) {
                const uint64_t adv = (uint64_t)((ks * UMMA_K * 4) >> 4);
                mma_tf32(d, dah + adv, dwh + adv, idesc, acc_flag);
                mma_tf32(d, dal + adv, dwh + adv, idesc, 1u);
                mma_tf32(d, dah + adv, dwl + adv, idesc, 1u);
                acc_flag = 1u;
              }
              mma_commit(&a_empty[aslot]);
              if (t == t_last) mma_commit(&w_empty[wslot]);
              if (++aslot == FT_NAS) { aslot = 0; aphase ^= 1u; }
            }
          }
        }
      }
      mma_commit(&acc_full);
    }
  } else {
    // =========================== weight loader (one thread) =========================================
    if (lane == 0) {
      int u = 0;
      for (int f = 0; f < C3P_NCELL; ++f) {
        if (!active[f]) continue;
        for (int kk = 0; kk < a.nkb * NKC; ++kk, ++u) {
          const int slot = u % NWS, use = u / NWS;
          if (use >= 1) mbar_wait(&w_empty[slot], (uint32_t)((use - 1) & 1));
          mbar_arrive_expect_tx(&w_full[slot], w_slot_bytes);
          bulk_copy_g2s(w_base + (size_t)slot * w_slot_bytes,
                        a.wp + (size_t)(f * a.nkb * NKC + kk) * w_slot_bytes, w_slot_bytes, &w_full[slot]);
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == FT_NPW + 1) tmem_dealloc(tmem, 512);
}

static size_t ft_smem_bytes(int Nout, int T, int NWS) {
  const size_t PT = (size_t)T * 128;
  return (size_t)NWS * 2 * Nout * PANEL_ROW_BYTES + (size_t)FT_NAS * FT_A_STAGE + PT * (4 + 4 + 56);
}

struct FTConfig {
  int NKC, nkb, T, NWS;
  size_t smem;
};

// Csrc = contraction width per cell (Cin forward, Cout backward), Nout = output width.
static bool ft_config(int N, long long capacity, int Csrc, int Nout, FTConfig* c, long long points = 0) {
  if (N > 65535) return false;                // per-point list offsets are kept as 16-bit prefixes
  if (capacity >= (1LL << 32)) return false;  // list starts are kept as 32-bit offsets
  if (Csrc % 32 || Nout % 16 || Csrc < 32 || Nout < 16 || Nout > 256) return false;
  c->NKC = (Csrc % 64 == 0) ? 2 : 1;
  c->nkb = Csrc / (32 * c->NKC);
  c->T = 512 / Nout >= 4 ? 4 : (512 / Nout);
  // Wave quantisation: with one CTA per SM the launch takes ceil(tiles/SMs) rounds of T sub-tiles each;
  // a smaller T can need fewer sub-tile rounds in total (e.g. 262,144 points on 148 SMs: T=4 -> 4x4 = 16,
  // T=3 -> 5x3 = 15).  Weight panels are re-streamed per tile, so only T >= 3 is considered.
  if (points > 0 && c->T == 4) {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    (void)cudaGetLastError();
    auto rounds = [&](int T) {
      const long long tiles = (points + (long long)T * 128 - 1) / ((long long)T * 128);
      return ((tiles + sms - 1) / sms) * T;
    };
    if (rounds(3) < rounds(4)) c->T = 3;
  }
  c->NWS = c->NKC + 1;
  c->smem = ft_smem_bytes(Nout, c->T, c->NWS);
  return c->smem <= 227 * 1024 - 1024;
}

// Shapes the tensor-core path takes; everything else stays on the fp32 SIMT engine.
bool forward_tc_supported(int N, long long capacity, int Cin, int Cout) {
  FTConfig c;
  return ft_config(N, capacity, Cin, Cout, &c);
}
bool backward_input_tc_supported(int N, long long capacity, int Cin, int Cout) {
  FTConfig c;
  return ft_config(N, capacity, Cout, Cin, &c);
}

size_t weight_panel_bytes(int Cin, int Cout) { return align_up((size_t)2 * C3P_NCELL * Cin * Cout * 4); }

size_t tc_items_bytes(const conv3p_geom_t* g, int Cin, int Cout) {
  if (gather_mma2_supported(g->N, g->pair_capacity, Cin, Cout) ||
      gather_mma2_supported(g->N, g->pair_capacity, Cout, Cin))
    return gather_mma2_scratch_bytes(g);
  return 0;
}

int launch_prep_weight_panels(const float* filter, void* wp, int Cin, int Cout, int transposed_out,
                              cudaStream_t stream) {
  const long long total = (long long)C3P_NCELL * Cin * Cout;
  const int threads = 256;
  const int blocks = (int)((total + threads - 1) / threads < 2048 ? (total + threads - 1) / threads : 2048);
  {
    LaunchTimer timer_("k_prep_weight_panels", stream);
    k_prep_weight_panels<<<blocks, threads, 0, stream>>>(filter, static_cast<unsigned char*>(wp), Cin, Cout,
                                                         transposed_out);
  }
  C3P_LAUNCH_CHECK("k_prep_weight_panels");
  return CONV3P_OK;
}

static int launch_gather_mma(FTArgs& a, const FTConfig& c, bool weighted, const char* name,
                             cudaStream_t stream) {
  a.nkb = c.nkb; a.T = c.T; a.NWS = c.NWS;
  a.debug = engine() >= 64 ? (engine() & ~(64 | 128)) : 0;
  const long long tiles = (a.total_points + (long long)a.T * 128 - 1) / ((long long)a.T * 128);
  if (tiles == 0) return CONV3P_OK;
  auto launch = [&](auto kern) -> int {
    C3P_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem));
    {
      LaunchTimer timer_(name, stream);
      kern<<<(unsigned)tiles, FT_THREADS, c.smem, stream>>>(a);
    }
    C3P_LAUNCH_CHECK(name);
    return CONV3P_OK;
  };
  if (c.NKC == 1) return weighted ? launch(k_gather_mma_tc<1, true>) : launch(k_gather_mma_tc<1, false>);
  return weighted ? launch(k_gather_mma_tc<2, true>) : launch(k_gather_mma_tc<2, false>);
}

int launch_forward_tc(const conv3p_geom_t* g, const PlanView& v, const float* input, const float* filter,
                      int Cin, int Cout, float* output, void* scratch, size_t scratch_bytes,
                      cudaStream_t stream, const RowIO& io) {
  FTConfig c;
  if (!ft_config(g->N, g->pair_capacity, Cin, Cout, &c, (long long)g->B * g->N)) return CONV3P_ERR_UNSUPPORTED;
  if (!scratch || scratch_bytes < weight_panel_bytes(Cin, Cout)) return CONV3P_ERR_BUFFER_TOO_SMALL;
  int st = launch_prep_weight_panels(filter, scratch, Cin, Cout, 0, stream);
  if (st) return st;
  if (!(engine() & 128) && gather_mma2_supported(g->N, g->pair_capacity, Cin, Cout)) {
    const size_t wpb = weight_panel_bytes(Cin, Cout);
    return launch_gather_mma2(g, v, input, scratch, Cin, Cout, output, false, static_cast<char*>(scratch) + wpb,
                              scratch_bytes - wpb, "k_forward_tc", stream, nullptr, io);
  }
  if (io.src_stride || io.out_stride || io.activation) return CONV3P_ERR_UNSUPPORTED;  // first-generation kernel: dense rows only
  FTArgs a{};
  a.src = input; a.wp = static_cast<const unsigned char*>(scratch); a.out = output;
  a.cnt = v.count_table; a.begin = v.pair_begin; a.len = v.pair_len; a.rows = v.pair_row;
  a.weights = nullptr; a.sorted_xyzi = v.sorted_xyzi;
  a.total_points = (long long)g->B * g->N; a.capacity = g->pair_capacity;
  a.N = g->N; a.Csrc = Cin; a.Nout = Cout;
  return launch_gather_mma(a, c, false, "k_forward_tc", stream);
}

int launch_backward_input_tc(const conv3p_geom_t* g, const PlanView& v, const float* grad_out,
                             const float* filter, int Cin, int Cout, float* grad_input, void* scratch,
                             size_t scratch_bytes, cudaStream_t stream, float* g_store) {
  FTConfig c;
  if (!ft_config(g->N, g->pair_capacity, Cout, Cin, &c, (long long)g->B * g->N)) return CONV3P_ERR_UNSUPPORTED;
  if (!scratch || scratch_bytes < weight_panel_bytes(Cin, Cout)) return CONV3P_ERR_BUFFER_TOO_SMALL;
  int st = launch_prep_weight_panels(filter, scratch, Cin, Cout, 1, stream);
  if (st) return st;
  if (!(engine() & 128) && gather_mma2_supported(g->N, g->pair_capacity, Cout, Cin)) {
    const size_t wpb = weight_panel_bytes(Cin, Cout);
    return launch_gather_mma2(g, v, grad_out, scratch, Cout, Cin, grad_input, true,
                              static_cast<char*>(scratch) + wpb, scratch_bytes - wpb, "k_backward_input_tc", stream,
                              g_store);
  }
  FTArgs a{};
  a.src = grad_out; a.wp = static_cast<const unsigned char*>(scratch); a.out = grad_input;
  a.cnt = v.bwd_count; a.begin = v.pair_begin; a.len = v.pair_len; a.rows = v.bwd_row;
  a.weights = v.bwd_weight; a.sorted_xyzi = v.sorted_xyzi;
  a.total_points = (long long)g->B * g->N; a.capacity = g->pair_capacity;
  a.N = g->N; a.Csrc = Cout; a.Nout = Cin;
  return launch_gather_mma(a, c, true, "k_backward_input_tc", stream);
}

}  // namespace c3p
