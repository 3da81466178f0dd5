// tcgen05 implicit-GEMM convolution for the ResBlock / AMPBlock convs (sm_100a).
//
// Mapping:  D[t, co] += sum_{tap j} A_j[t, ci] * W_j[co, ci]
//   M = time (128 rows per MMA, m M-tiles per CTA), N = C_out (<= 256), K = C_in per tap.
//   A (activations): fp16/bf16 in shared memory, no-swizzle K-major core-matrix
//     layout [ci/8][row][8]: a tap shift of j*d time steps is a +j*d*16 B move of
//     the descriptor start address, so one resident activation tile serves all k taps.
//   B (weights): pre-packed in global memory in exactly the shared-memory image
//     ([tap][ci/32][ (ci%32)/8 ][co][8]) and streamed stage by stage with
//     cp.async.bulk (TMA, 1-D) into an mbarrier ring.
//   D: fp32 in TMEM, m * Np columns; read back with tcgen05.ld for the epilogues.
// Phases per CTA tile: load+activate x -> smem A | conv1 MMAs | epilogue 1
// (TMEM -> +bias, lrelu -> A, aliased) | conv2 MMAs | epilogue 2 (TMEM -> +bias
// +residual (+branch sum, /nk) -> global).  Warp roles: 8 worker warps, 1 TMA
// producer warp, 1 MMA-issue warp.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include <vector>

#include "ab_tc.cuh"
#include "ab_tc_ptx.cuh"

namespace ab {

using namespace tcx;

// operand layout knob (debug): AB_TC_LAYOUT=0 -> no-swizzle interleaved, default SWIZZLE_32B
int tc_layout() {
  static const int v = [] { const char* e = getenv("AB_TC_LAYOUT"); return (e && e[0] == '0') ? 0 : 1; }();
  return v;
}

namespace {

constexpr int TC_WORKER_WARPS = 8;
constexpr int TC_WORKERS = TC_WORKER_WARPS * 32;
constexpr int TC_THREADS = TC_WORKERS + 64;  // + producer warp (8) + MMA warp (9)
constexpr int TC_MAX_STAGES = 8;
constexpr int TC_MAX_C = 256;
constexpr uint32_t TC_SMEM_LIMIT = 227 * 1024;

struct TcGeom {
  int Np;          // channels padded to 16 (N and K extent)
  int nkc;         // 32-channel K chunks per tap
  int m;           // M tiles (128 rows each) per CTA
  int V;           // valid output rows per CTA tile
  int rowsA;       // allocated activation rows (multiple of 8)
  int tiles;       // tiles per sequence
  int nstages;
  uint32_t stage_bytes;
  uint32_t off_w, off_bias, off_bar;  // byte offsets in dynamic smem (A at 0)
  uint32_t smem_bytes;
  int hh;          // time of A row 0 is T0 - hh
  int h2;          // time of the intermediate row 0 is T0 - h2 (pair mode)
  uint32_t idesc;
  int swap_lbo_sbo;  // debug knob
  int layout;        // 0 = no-swizzle interleaved core matrices, 1 = SWIZZLE_32B rows
  int base_off;      // debug knob: fill the descriptor base_offset field from the start address
  int stagger_groups;        // first-wave CTAs are delayed by (smid % groups) * stagger_cycles so that the
  long long stagger_cycles;  // HBM-bound phases of different SMs do not run in lock-step
  long long* dbg;    // debug: per-CTA phase timestamps (AB_TC_DEBUG_TIMING=1), else nullptr
};

constexpr int DBG_BLOCKS = 2048, DBG_SLOTS = 8;

// ---------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1) tc_conv_kernel(TcConvParams p, TcGeom g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / g.tiles, tile = blockIdx.x - b * g.tiles;
  const int T0 = tile * g.V;
  const int bf16 = p.precision == AB_PREC_TC_BF16;

  const uint32_t sA = smem_u32(smem);
  const uint32_t sW = sA + g.off_w;
  float* bias_s = reinterpret_cast<float*>(smem + g.off_bias);   // [2][Np]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + g.off_bar);
  const uint32_t bar0 = smem_u32(bars);
  // barrier slots: full[s] = s, empty[s] = MAX+s, a_ready = 2*MAX, acc_full = 2*MAX+1
  auto bar_full = [&](int s) { return bar0 + 8u * s; };
  auto bar_empty = [&](int s) { return bar0 + 8u * (TC_MAX_STAGES + s); };
  const uint32_t bar_aready = bar0 + 8u * (2 * TC_MAX_STAGES);
  const uint32_t bar_accfull = bar0 + 8u * (2 * TC_MAX_STAGES + 1);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_MAX_STAGES + 2);

  if (threadIdx.x == 0) {
    for (int s = 0; s < g.nstages; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1);
    }
    mbar_init(bar_aready, TC_WORKERS);
    mbar_init(bar_accfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
  }
  if (warp == TC_WORKER_WARPS + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  auto stamp = [&](int slot) {
    if (g.dbg != nullptr && threadIdx.x == 0 && blockIdx.x < DBG_BLOCKS)
      g.dbg[blockIdx.x * DBG_SLOTS + slot] = clock64();
  };
  if (g.stagger_groups > 1 && blockIdx.x < gridDim.x && blockIdx.x < 148u) {
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    const long long wait = (long long)(smid % (uint32_t)g.stagger_groups) * g.stagger_cycles;
    const long long t0 = clock64();
    while (clock64() - t0 < wait) __nanosleep(200);
  }
  stamp(0);

  const uint32_t lboA = (uint32_t)g.rowsA * 16u;   // next 8-channel group
  const uint32_t lboB = (uint32_t)g.Np * 16u;
  const int mrows = g.m * 128;

  if (warp < TC_WORKER_WARPS) {
    // ===================== worker warps =====================
    for (int i = threadIdx.x; i < 2 * g.Np; i += TC_WORKERS) {
      const int which = i / g.Np, c = i - which * g.Np;
      const float* src = which ? p.b2 : p.b1;
      bias_s[i] = (src != nullptr && c < p.C) ? __ldg(src + c) : 0.f;
    }
    // ---- prologue: A[row][ci] = lrelu(x[b, ci, T0 - hh + row]) (zero outside [0,T) and for ci >= C)
    // one item = 8 channels x 128 rows; each lane keeps 32 independent loads in flight
    {
      const int ngrp = (g.rowsA + 127) >> 7;
      const int c8n = g.Np >> 3;
      const float* xb = p.x + (int64_t)b * p.C * p.T;
      for (int item = warp; item < c8n * ngrp; item += TC_WORKER_WARPS) {
        const int c8 = item / ngrp, grp = item - c8 * ngrp;
        const int row0 = (grp << 7) + lane;
        float v[4][8];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int row = row0 + 32 * r;
          const int t = T0 - g.hh + row;
          const bool ok = row < g.rowsA && t >= 0 && t < p.T;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int c = c8 * 8 + e;
            v[r][e] = (ok && c < p.C) ? __ldg(xb + (int64_t)c * p.T + t) : 0.f;
          }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int row = row0 + 32 * r;
          if (row < g.rowsA) {
            uint4 q;
            q.x = pack2(lrelu(v[r][0], p.pre_slope), lrelu(v[r][1], p.pre_slope), bf16);
            q.y = pack2(lrelu(v[r][2], p.pre_slope), lrelu(v[r][3], p.pre_slope), bf16);
            q.z = pack2(lrelu(v[r][4], p.pre_slope), lrelu(v[r][5], p.pre_slope), bf16);
            q.w = pack2(lrelu(v[r][6], p.pre_slope), lrelu(v[r][7], p.pre_slope), bf16);
            *reinterpret_cast<uint4*>(smem + unit_offset(g.layout, g.rowsA, c8, row)) = q;
          }
        }
      }
    }
    fence_proxy_async();
    mbar_arrive(bar_aready);
    stamp(1);

    const int q4 = warp & 3, hsel = warp >> 2;
    const int nch = g.Np >> 4;  // 16-column chunks
    if (p.nconv == 2) {
      // ---- epilogue 1: intermediate = lrelu(conv1 + b1) -> A (aliased), zero outside [0,T)
      mbar_wait(bar_accfull, 0, 10);
      tc_fence_after();
      stamp(2);
      for (int i = 0; i < g.m; ++i) {
        const int row = i * 128 + q4 * 32 + lane;
        const int t = T0 - g.h2 + row;
        const bool ok = t >= 0 && t < p.T;
        for (int ch = hsel; ch < nch; ch += 2) {
          uint32_t r[16];
          tc_ld16(tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(i * g.Np + ch * 16), r);
          tc_wait_ld();
          float v[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float a = __uint_as_float(r[e]) + bias_s[ch * 16 + e];
            v[e] = ok ? lrelu(a, p.mid_slope) : 0.f;
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint4 q;
            q.x = pack2(v[8 * h + 0], v[8 * h + 1], bf16);
            q.y = pack2(v[8 * h + 2], v[8 * h + 3], bf16);
            q.z = pack2(v[8 * h + 4], v[8 * h + 5], bf16);
            q.w = pack2(v[8 * h + 6], v[8 * h + 7], bf16);
            *reinterpret_cast<uint4*>(smem + unit_offset(g.layout, g.rowsA, ch * 2 + h, row)) = q;
          }
        }
      }
      tc_fence_before();
      fence_proxy_async();
      mbar_arrive(bar_aready);
      stamp(3);
    }
    // ---- epilogue 2: y = ((acc + bias) + residual + acc_prev) / out_div
    mbar_wait(bar_accfull, (uint32_t)(p.nconv - 1), 11);
    tc_fence_after();
    stamp(4);
    const float* bias2 = bias_s + (p.nconv == 2 ? g.Np : 0);
    for (int i = 0; i < g.m; ++i) {
      const int row = i * 128 + q4 * 32 + lane;
      const int t = T0 + row;
      const bool ok = row < g.V && t < p.T;
      const float* rrow = p.residual ? p.residual + (int64_t)b * p.C * p.T + t : nullptr;
      const float* arow = p.acc_prev ? p.acc_prev + (int64_t)b * p.C * p.T + t : nullptr;
      float* yrow = p.y + (int64_t)b * p.C * p.T + t;
      // this warp's column chunks: hsel, hsel+2, ... ; two chunks (32 columns) per iteration
      for (int ch = hsel; ch < nch; ch += 4) {
        const bool two = ch + 2 < nch;
        uint32_t r0[16], r1[16];
        const uint32_t tbase = tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(i * g.Np);
        tc_ld16(tbase + (uint32_t)(ch * 16), r0);
        if (two) tc_ld16(tbase + (uint32_t)((ch + 2) * 16), r1);
        float res[32], acp[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int co = (e < 16 ? ch : ch + 2) * 16 + (e & 15);
          const bool w = ok && co < p.C && (e < 16 || two);
          res[e] = (w && rrow) ? __ldg(rrow + (int64_t)co * p.T) : 0.f;
          acp[e] = (w && arow) ? __ldg(arow + (int64_t)co * p.T) : 0.f;
        }
        tc_wait_ld();
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int co = (e < 16 ? ch : ch + 2) * 16 + (e & 15);
          if (ok && co < p.C && (e < 16 || two)) {
            float v = __uint_as_float(e < 16 ? r0[e & 15] : r1[e & 15]) + bias2[co];
            v += res[e];
            v += acp[e];
            if (p.out_div != 1.0f) v = v / p.out_div;
            yrow[(int64_t)co * p.T] = v;
          }
        }
      }
    }
    tc_fence_before();
    stamp(5);
  } else if (warp == TC_WORKER_WARPS) {
    // ===================== TMA weight producer =====================
    if (lane == 0) {
      const int per_conv = p.k * g.nkc;
      const int total = p.nconv * per_conv;
      for (int it = 0; it < total; ++it) {
        const int s = it % g.nstages;
        const uint32_t ph = (uint32_t)(it / g.nstages) & 1u;
        mbar_wait(bar_empty(s), ph ^ 1u, 20);
        mbar_arrive_expect_tx(bar_full(s), g.stage_bytes);
        const int conv = it / per_conv, local = it - conv * per_conv;
        const uint8_t* src = static_cast<const uint8_t*>(conv ? p.w2 : p.w1) + (size_t)local * g.stage_bytes;
        bulk_g2s(sW + (uint32_t)s * g.stage_bytes, src, g.stage_bytes, bar_full(s));
      }
    }
  } else {
    // ===================== MMA issuer =====================
    // All 32 lanes run the warp-uniform loops; one elected lane issues tcgen05.mma / commit.
    // Descriptors (cute::UMMA::SmemDescriptor, no swizzle, K-major): hi word is constant
    // (SBO = 128 B between 8-row groups, version 1); lo word = addr>>4 | (LBO>>4)<<16.
    const uint32_t elected = elect_one_sync();
    const int nks_total = g.Np >> 4;  // 16-channel K steps per tap
    const uint32_t rows16 = (uint32_t)g.rowsA;          // LBO of A in 16-byte units
    const uint32_t np16 = (uint32_t)g.Np;               // LBO of B in 16-byte units
    // layout 0: LBO = rows*16 B (next K core matrix), SBO = 128 B (next 8-row group), no swizzle.
    // layout 1: SWIZZLE_32B (layout_type 6): LBO field 1 (unused), SBO = 256 B; K step = next [c16] chunk.
    uint64_t hiA, hiB;
    uint32_t lboA_f, lboB_f, kstepA, kstepB, rowunit;
    if (g.layout == 0) {
      hiA = g.swap_lbo_sbo ? ((uint64_t)(rows16 | (1u << 14)) << 32) : ((uint64_t)(8u | (1u << 14)) << 32);
      hiB = g.swap_lbo_sbo ? ((uint64_t)(np16 | (1u << 14)) << 32) : ((uint64_t)(8u | (1u << 14)) << 32);
      lboA_f = (g.swap_lbo_sbo ? 8u : rows16) << 16;
      lboB_f = (g.swap_lbo_sbo ? 8u : np16) << 16;
      kstepA = 2u * rows16;   // two 8-channel units per K=16 step
      kstepB = 2u * np16;
      rowunit = 1u;           // 16 B per row
    } else {
      hiA = hiB = ((uint64_t)(16u | (1u << 14)) << 32) | (6ull << 61);
      lboA_f = lboB_f = 1u << 16;
      kstepA = 2u * rows16;   // one [c16] chunk = rows * 32 B
      kstepB = 2u * np16;
      rowunit = 2u;           // 32 B per row
    }
    const uint32_t a16 = sA >> 4, w16 = sW >> 4, stage16 = g.stage_bytes >> 4;
    int it = 0;
    for (int conv = 0; conv < p.nconv; ++conv) {
      mbar_wait(bar_aready, (uint32_t)conv & 1u, 30);
      tc_fence_after();
      const int dil = conv == 0 ? p.d1 : 1;
      for (int j = 0; j < p.k; ++j) {
        for (int kc = 0; kc < g.nkc; ++kc, ++it) {
          const int s = it % g.nstages;
          const uint32_t ph = (uint32_t)(it / g.nstages) & 1u;
          mbar_wait(bar_full(s), ph, 31);
          tc_fence_after();
          const bool two = nks_total - kc * 2 >= 2;
          const uint32_t astart = a16 + (uint32_t)(kc * 2) * kstepA + (uint32_t)(j * dil) * rowunit;
          uint32_t alo = astart | lboA_f;
          const uint32_t blo = (w16 + (uint32_t)s * stage16) | lboB_f;
          const uint32_t acc0 = (j | kc) != 0 ? 1u : 0u;
          uint32_t td = tmem;
          for (int i = 0; i < g.m; ++i) {
            if (elected) {
              // base_offset (bits 49-51) = (start address >> 7) & 7 when the debug knob asks for it
              const uint64_t bo0 = g.base_off ? ((uint64_t)((alo >> 3) & 7u) << 49) : 0ull;
              const uint64_t bo1 = g.base_off ? ((uint64_t)(((alo + kstepA) >> 3) & 7u) << 49) : 0ull;
              tc_mma_f16(td, hiA | bo0 | alo, hiB | blo, g.idesc, acc0);
              if (two) tc_mma_f16(td, hiA | bo1 | (alo + kstepA), hiB | (blo + kstepB), g.idesc, 1u);
            }
            alo += 128u * rowunit;
            td += (uint32_t)g.Np;
          }
          if (elected) tc_commit(bar_empty(s));
          __syncwarp();
        }
      }
      if (elected) tc_commit(bar_accfull);
      __syncwarp();
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == TC_WORKER_WARPS + 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
  (void)mrows;
}

// ---------------------------------------------------------------------------
// weight image:  [tap j][kc][c8l 0..3][co 0..Np)[8]  16-bit, zero padded
// ---------------------------------------------------------------------------
__global__ void tc_pack_weight_kernel(const float* __restrict__ w_t, uint16_t* __restrict__ img, int cin,
                                      int cout, int k, int Np, int nkc, int bf16, int layout) {
  // one thread per 16-bit element of the image; stage (j, kc) holds channels [32*kc, 32*kc+32)
  const int64_t total = (int64_t)k * nkc * 4 * Np * 8;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int e = (int)(idx & 7);
    int64_t r = idx >> 3;
    const int co = (int)(r % Np);
    r /= Np;
    const int c8l = (int)(r & 3);
    r >>= 2;
    const int kc = (int)(r % nkc);
    const int j = (int)(r / nkc);
    const int ci = kc * 32 + c8l * 8 + e;
    float v = 0.f;
    if (ci < cin && co < cout) v = w_t[((int64_t)ci * k + j) * cout + co];
    uint16_t bits;
    if (bf16) {
      __nv_bfloat16 h = __float2bfloat16_rn(v);
      bits = *reinterpret_cast<uint16_t*>(&h);
    } else {
      v = fminf(fmaxf(v, -65504.f), 65504.f);
      __half h = __float2half_rn(v);
      bits = *reinterpret_cast<uint16_t*>(&h);
    }
    const int64_t stage = (int64_t)j * nkc + kc;
    int64_t off;   // in 16-bit elements within the image
    if (layout == 0) {
      off = stage * (4 * Np * 8) + ((int64_t)c8l * Np + co) * 8 + e;
    } else {
      const int unit = (c8l & 1) ^ ((co >> 2) & 1);
      off = stage * (4 * Np * 8) + (int64_t)(c8l >> 1) * Np * 16 + (int64_t)co * 16 + unit * 8 + e;
    }
    img[off] = bits;
  }
}

int round_up(int x, int a) { return (x + a - 1) / a * a; }

int make_geom(const TcConvParams& p, TcGeom& g) {
  if (p.C <= 0 || p.C > TC_MAX_C) return fail(AB_ERR_UNSUPPORTED, "tc_conv: C=%d not in [1,%d]", p.C, TC_MAX_C);
  if (p.k <= 0 || !(p.k & 1) || p.d1 <= 0) return fail(AB_ERR_UNSUPPORTED, "tc_conv: need odd k and positive dilation");
  if (p.nconv != 1 && p.nconv != 2) return fail(AB_ERR_ARG, "tc_conv: nconv must be 1 or 2");
  g.Np = round_up(p.C, 16);
  g.nkc = (g.Np + 31) / 32;
  g.stage_bytes = (uint32_t)g.Np * 64u;
  const int h1 = (p.k - 1) * p.d1 / 2, h2 = (p.k - 1) / 2;
  g.h2 = h2;
  g.hh = p.nconv == 2 ? h1 + h2 : h1;
  const int lost = p.nconv == 2 ? (p.k - 1) : 0;   // rows of the tile that conv2 cannot produce
  const int halo = (p.k - 1) * p.d1;
  int m = 512 / g.Np;
  if (m > 16) m = 16;
  // do not tile far past the sequence
  while (m > 1 && (m - 1) * 128 - lost >= p.T) --m;
  const uint32_t misc = 2u * g.Np * 4u + 8u * (2 * TC_MAX_STAGES + 2) + 16u;
  for (;; --m) {
    if (m < 1) return fail(AB_ERR_UNSUPPORTED, "tc_conv: C=%d k=%d d=%d does not fit shared memory", p.C, p.k, p.d1);
    if (m * 128 - lost < 8) continue;
    g.rowsA = round_up(m * 128 + halo, 8);
    const uint32_t abytes = (uint32_t)g.rowsA * (uint32_t)g.Np * 2u;
    if (abytes + 2u * g.stage_bytes + misc + 1280u > TC_SMEM_LIMIT) continue;
    g.m = m;
    int ns = (int)((TC_SMEM_LIMIT - abytes - misc - 1280u) / g.stage_bytes);
    if (ns > TC_MAX_STAGES) ns = TC_MAX_STAGES;
    g.nstages = ns;
    g.off_w = (abytes + 1023u) & ~1023u;
    g.off_bias = g.off_w + (uint32_t)ns * g.stage_bytes;
    g.off_bar = (g.off_bias + 2u * g.Np * 4u + 15u) & ~15u;
    g.smem_bytes = g.off_bar + 8u * (2 * TC_MAX_STAGES + 2) + 16u;
    break;
  }
  g.V = ((g.m * 128 - lost) / 8) * 8;
  g.tiles = (p.T + g.V - 1) / g.V;
  const uint32_t fmt = p.precision == AB_PREC_TC_BF16 ? 1u : 0u;
  // cute::UMMA::InstrDescriptor: c_format F32 (1) @4, a_format @7, b_format @10, K-major A and B,
  // N>>3 @17, M>>4 @24
  g.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(g.Np >> 3) << 17) | ((128u >> 4) << 24);
  const char* sw = getenv("AB_TC_SWAP_LBO_SBO");
  g.swap_lbo_sbo = (sw && sw[0] == '1') ? 1 : 0;
  g.layout = tc_layout();
  const char* bo = getenv("AB_TC_BASE_OFFSET");
  g.base_off = (bo && bo[0] == '1') ? 1 : 0;
  g.dbg = nullptr;
  static const int stag = [] { const char* e = getenv("AB_TC_STAGGER"); return e ? atoi(e) : 0; }();
  g.stagger_groups = stag;
  const double mma = (double)p.nconv * p.k * g.nkc * g.m * 2.0 * (64.0 + g.Np / 2.0);
  const double mem = (double)g.m * 128.0 * g.Np * 12.0 / 11.0;
  g.stagger_cycles = stag > 1 ? (long long)((mma + mem + 15000.0) / stag) : 0;
  return AB_OK;
}

}  // namespace

int tc_max_channels() { return TC_MAX_C; }

bool tc_conv_supported(int C, int k) { return C > 0 && C <= TC_MAX_C && (k & 1) && k <= 31; }

size_t tc_weight_image_bytes(int cin, int cout, int k) {
  if (cin != cout || !tc_conv_supported(cin, k)) return 0;
  const int Np = round_up(cout, 16), nkc = (Np + 31) / 32;
  return (size_t)k * nkc * Np * 64;
}

int launch_tc_pack_weight(const float* w_t, void* image, int cin, int cout, int k, int precision,
                          cudaStream_t s) {
  const size_t bytes = tc_weight_image_bytes(cin, cout, k);
  if (bytes == 0) return AB_OK;  // shape not served by the tensor-core path
  const int Np = round_up(cout, 16), nkc = (Np + 31) / 32;
  const int64_t total = (int64_t)bytes / 2;
  const int blocks = (int)std::min<int64_t>((total + 255) / 256, 148 * 8);
  tc_pack_weight_kernel<<<blocks, 256, 0, s>>>(w_t, static_cast<uint16_t*>(image), cin, cout, k, Np, nkc,
                                               precision == AB_PREC_TC_BF16 ? 1 : 0, tc_layout());
  AB_LAUNCH_CHECK("tc_pack_weight_kernel");
  return AB_OK;
}

int launch_tc_conv(const TcConvParams& p, cudaStream_t s) {
  if (!p.x || !p.y || !p.w1 || (p.nconv == 2 && !p.w2)) return fail(AB_ERR_ARG, "tc_conv: null argument");
  if (p.B <= 0 || p.T <= 0) return fail(AB_ERR_ARG, "tc_conv: bad shape");
  if (p.precision != AB_PREC_TC_F16 && p.precision != AB_PREC_TC_BF16) return fail(AB_ERR_ARG, "tc_conv: bad precision");
  TcGeom g;
  int rc = make_geom(p, g);
  if (rc != AB_OK) return rc;
  static bool configured = false;
  if (!configured) {
    AB_CUDA_TRY(cudaFuncSetAttribute(tc_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_LIMIT));
    configured = true;
  }
  const int64_t grid = (int64_t)p.B * g.tiles;
  if (grid > 0x7fffffffll) return fail(AB_ERR_UNSUPPORTED, "tc_conv: grid too large");
  // request > half of the SM's shared memory so exactly one CTA (512 TMEM columns) is resident
  const uint32_t smem = std::max<uint32_t>(g.smem_bytes, 120u * 1024u);
  static const bool dbg_on = [] { const char* e = getenv("AB_TC_DEBUG_TIMING"); return e && e[0] == '1'; }();
  static long long* dbg_buf = nullptr;
  if (dbg_on) {  // debug only: the one place the library allocates, never on the product path
    if (!dbg_buf) AB_CUDA_TRY(cudaMalloc(&dbg_buf, sizeof(long long) * DBG_BLOCKS * DBG_SLOTS));
    AB_CUDA_TRY(cudaMemsetAsync(dbg_buf, 0, sizeof(long long) * DBG_BLOCKS * DBG_SLOTS, s));
    g.dbg = dbg_buf;
  }
  tc_conv_kernel<<<(unsigned)grid, TC_THREADS, smem, s>>>(p, g);
  AB_LAUNCH_CHECK("tc_conv_kernel");
  if (dbg_on) {
    AB_CUDA_TRY(cudaStreamSynchronize(s));
    static std::vector<long long> h(DBG_BLOCKS * DBG_SLOTS);
    AB_CUDA_TRY(cudaMemcpy(h.data(), dbg_buf, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost));
    const int nb = (int)std::min<int64_t>(grid, DBG_BLOCKS);
    double ph[5] = {0, 0, 0, 0, 0};
    int cnt = 0;
    for (int i = 0; i < nb; ++i) {
      const long long* r = &h[(size_t)i * DBG_SLOTS];
      if (r[5] == 0) continue;
      ++cnt;
      ph[0] += (double)(r[1] - r[0]);                       // prologue
      if (p.nconv == 2) {
        ph[1] += (double)(r[2] - r[1]);                     // conv1 MMA (wait)
        ph[2] += (double)(r[3] - r[2]);                     // epilogue 1
        ph[3] += (double)(r[4] - r[3]);                     // conv2 MMA (wait)
      } else {
        ph[1] += (double)(r[4] - r[1]);
      }
      ph[4] += (double)(r[5] - r[4]);                       // epilogue 2
    }
    if (cnt) {
      const double ideal = (double)g.m * (g.Np / 2.0) * (g.Np / 16.0) * p.k;   // cycles per conv at 8192 flop/clk/SM
      fprintf(stderr,
              "[tc_timing] C=%d k=%d d=%d nconv=%d m=%d V=%d tiles=%lld stages=%d | cycles: prologue %.0f conv1 %.0f epi1 %.0f "
              "conv2 %.0f epi2 %.0f | ideal MMA/conv %.0f\n",
              p.C, p.k, p.d1, p.nconv, g.m, g.V, (long long)grid, g.nstages, ph[0] / cnt, ph[1] / cnt, ph[2] / cnt,
              ph[3] / cnt, ph[4] / cnt, ideal);
    }
  }
  return AB_OK;
}

}  // namespace ab
