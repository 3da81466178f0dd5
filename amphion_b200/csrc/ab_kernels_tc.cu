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

#include "ab_tc.cuh"

namespace ab {

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
};

// ---------------------------------------------------------------------------
// PTX helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// Bounded wait: a protocol bug traps (-> cudaErrorLaunchFailure) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int tag) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 22)) {
      printf("amphion_b200: mbarrier timeout tag=%d block=%d thread=%d parity=%u\n", tag, blockIdx.x,
             threadIdx.x, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// no-swizzle K-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   [0,14) start>>4 | [16,30) LBO>>4 (next core matrix along K) | [32,46) SBO>>4
//   (next 8-row group along M/N) | [46,48) version=1 | [61,64) layout=0 (none)
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

__device__ __forceinline__ uint32_t pack2(float a, float b, int bf16) {
  if (bf16) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  // saturating: fp16 operands must never become inf
  a = fminf(fmaxf(a, -65504.f), 65504.f);
  b = fminf(fmaxf(b, -65504.f), 65504.f);
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ float lrelu(float v, float slope) { return v >= 0.f ? v : v * slope; }

// ---------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1) tc_conv_kernel(TcConvParams p, TcGeom g) {
  extern __shared__ __align__(128) uint8_t smem[];
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
    {
      const int nrb = g.rowsA >> 5, rem = g.rowsA & 31;
      const int nblocks = nrb + (rem ? 1 : 0);
      const int c8n = g.Np >> 3;
      const float* xb = p.x + (int64_t)b * p.C * p.T;
      for (int item = warp; item < c8n * nblocks; item += TC_WORKER_WARPS) {
        const int c8 = item / nblocks, rb = item - c8 * nblocks;
        const int row = (rb << 5) + lane;
        const int t = T0 - g.hh + row;
        const bool ok = row < g.rowsA && t >= 0 && t < p.T;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int c = c8 * 8 + e;
          v[e] = (ok && c < p.C) ? __ldg(xb + (int64_t)c * p.T + t) : 0.f;
        }
        if (row < g.rowsA) {
          uint4 q;
          q.x = pack2(lrelu(v[0], p.pre_slope), lrelu(v[1], p.pre_slope), bf16);
          q.y = pack2(lrelu(v[2], p.pre_slope), lrelu(v[3], p.pre_slope), bf16);
          q.z = pack2(lrelu(v[4], p.pre_slope), lrelu(v[5], p.pre_slope), bf16);
          q.w = pack2(lrelu(v[6], p.pre_slope), lrelu(v[7], p.pre_slope), bf16);
          *reinterpret_cast<uint4*>(smem + (size_t)c8 * lboA + (size_t)row * 16) = q;
        }
      }
    }
    fence_proxy_async();
    mbar_arrive(bar_aready);

    const int q4 = warp & 3, hsel = warp >> 2;
    const int nch = g.Np >> 4;  // 16-column chunks
    if (p.nconv == 2) {
      // ---- epilogue 1: intermediate = lrelu(conv1 + b1) -> A (aliased), zero outside [0,T)
      mbar_wait(bar_accfull, 0, 10);
      tc_fence_after();
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
            *reinterpret_cast<uint4*>(smem + (size_t)(ch * 2 + h) * lboA + (size_t)row * 16) = q;
          }
        }
      }
      tc_fence_before();
      fence_proxy_async();
      mbar_arrive(bar_aready);
    }
    // ---- epilogue 2: y = ((acc + bias) + residual + acc_prev) / out_div
    mbar_wait(bar_accfull, (uint32_t)(p.nconv - 1), 11);
    tc_fence_after();
    const float* bias2 = bias_s + (p.nconv == 2 ? g.Np : 0);
    for (int i = 0; i < g.m; ++i) {
      const int row = i * 128 + q4 * 32 + lane;
      const int t = T0 + row;
      const bool ok = row < g.V && t < p.T;
      for (int ch = hsel; ch < nch; ch += 2) {
        uint32_t r[16];
        tc_ld16(tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(i * g.Np + ch * 16), r);
        tc_wait_ld();
        float res[16], acp[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int co = ch * 16 + e;
          const int64_t idx = ((int64_t)b * p.C + co) * p.T + t;
          const bool w = ok && co < p.C;
          res[e] = (w && p.residual) ? __ldg(p.residual + idx) : 0.f;
          acp[e] = (w && p.acc_prev) ? __ldg(p.acc_prev + idx) : 0.f;
        }
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int co = ch * 16 + e;
          if (ok && co < p.C) {
            float v = __uint_as_float(r[e]) + bias2[co];
            v += res[e];
            v += acp[e];
            if (p.out_div != 1.0f) v = v / p.out_div;
            p.y[((int64_t)b * p.C + co) * p.T + t] = v;
          }
        }
      }
    }
    tc_fence_before();
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
    const int nks_total = g.Np >> 4;  // 16-channel K steps per tap
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
          if (lane == 0) {
            const int nks = min(2, nks_total - kc * 2);
            const uint32_t wbase = sW + (uint32_t)s * g.stage_bytes;
            for (int i = 0; i < g.m; ++i) {
              for (int ks = 0; ks < nks; ++ks) {
                const uint32_t aaddr = sA + (uint32_t)(kc * 4 + ks * 2) * lboA + (uint32_t)(i * 128 + j * dil) * 16u;
                const uint32_t baddr = wbase + (uint32_t)(ks * 2) * lboB;
                const uint64_t ad = g.swap_lbo_sbo ? make_desc(aaddr, 128u, lboA) : make_desc(aaddr, lboA, 128u);
                const uint64_t bd = g.swap_lbo_sbo ? make_desc(baddr, 128u, lboB) : make_desc(baddr, lboB, 128u);
                tc_mma_f16(tmem + (uint32_t)(i * g.Np), ad, bd, g.idesc, (j | kc | ks) != 0 ? 1u : 0u);
              }
            }
            tc_commit(bar_empty(s));
          }
          __syncwarp();
        }
      }
      if (lane == 0) tc_commit(bar_accfull);
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
                                      int cout, int k, int Np, int nkc, int bf16) {
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
    img[idx] = bits;
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
    if (abytes + 2u * g.stage_bytes + misc + 256u > TC_SMEM_LIMIT) continue;
    g.m = m;
    int ns = (int)((TC_SMEM_LIMIT - abytes - misc - 256u) / g.stage_bytes);
    if (ns > TC_MAX_STAGES) ns = TC_MAX_STAGES;
    g.nstages = ns;
    g.off_w = (abytes + 127u) & ~127u;
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
                                               precision == AB_PREC_TC_BF16 ? 1 : 0);
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
  tc_conv_kernel<<<(unsigned)grid, TC_THREADS, smem, s>>>(p, g);
  AB_LAUNCH_CHECK("tc_conv_kernel");
  return AB_OK;
}

}  // namespace ab
