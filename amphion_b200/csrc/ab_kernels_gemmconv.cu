// tcgen05 implicit-GEMM kernel with N blocking: ConvTranspose1d (polyphase) and wide Conv1d.
//
//   D[row, n] = sum_{tap} sum_{ci} A[row + shift(tap), ci] * W[nb][tap][n, ci]
//   M = time rows (128 per MMA, m M-tiles per CTA), K = C_in per tap, N = one block of <= 256
//   "virtual output channels"; the CTA walks NB such blocks over ONE resident activation tile.
//
// ConvTranspose1d(stride u, padding (k-u)/2)  (hifigan.py:176-186, 206-207; bigvgan.py:254-276):
//   y[co, s*u - p + phi] = b[co] + sum_m sum_ci act(x)[ci, s - m] * W[ci, co, phi + u*m]
//   -> rows = input times s, virtual channel n = co_local*u + phi, taps m = 0..ceil(k/u)-1,
//   so every MAC is a useful one (no zero-stuffing) and a thread's TMEM row holds u consecutive
//   output samples of each channel.
// Conv1d ("same", dilation d): rows = output times, n = co, taps j = 0..k-1, shift j*d.
//
// TMEM is double buffered across N blocks (when NB > 1) so the epilogue of block nb overlaps the
// MMAs of block nb+1.  Operand layouts and the warp roles are those of ab_kernels_tc.cu.
#include <stdlib.h>

#include <algorithm>

#include "ab_tc.cuh"
#include "ab_tc_ptx.cuh"

namespace ab {

using namespace tcx;

namespace {

constexpr int GC_WORKER_WARPS = 8;
constexpr int GC_WORKERS = GC_WORKER_WARPS * 32;
constexpr int GC_THREADS = GC_WORKERS + 64;
constexpr int GC_MAX_STAGES = 8;
constexpr uint32_t GC_SMEM_LIMIT = 227 * 1024;

struct GcGeom {
  int mode;        // 0 conv, 1 conv-transpose
  int Kp, nkc;     // padded C_in, 32-channel chunks
  int Nb, NB;      // N block width (multiple of 16) and count
  int cc;          // convT: output channels per N block
  int ntaps;
  int m, nbuf;
  int rowsA;
  int tiles;       // per sequence
  int nstages;
  uint32_t stage_bytes;
  uint32_t off_w, off_bias, off_bar, smem_bytes;
  int row0_time;   // time of A row 0 is tile_origin - row0_time
  int Tout;        // output length
  int pad;         // convT: (k-u)/2
  int grouped;     // convT: 8-channel-group epilogue (u in {2,4,8}, cc % 8 == 0)
  int ctas;        // CTAs per SM this geometry was sized for (1 or 2): TMEM columns = 512 / ctas, shared memory likewise
  uint32_t idesc;
};

// ConvTranspose epilogue for one group of 8 output channels (8*U accumulator columns of one TMEM row):
// the row is input time s, column (c*U + phi) holds y[co0 + c, s*U - pad + phi].  Stores the fp32 samples
// (U consecutive floats per channel) and, optionally, the fp16 operand image row of each output time.
template <int U, class P, class G>
__device__ __forceinline__ void convT_group_store(const P& p, const G& g, uint32_t taddr, int b, int co0,
                                                  int s, const float* bias_s, int bf16) {
  uint32_t r[8 * U];
#pragma unroll
  for (int q = 0; q < (8 * U) / 16; ++q) {
    uint32_t tmp[16];
    tc_ld16(taddr + (uint32_t)(q * 16), tmp);
#pragma unroll
    for (int e = 0; e < 16; ++e) r[q * 16 + e] = tmp[e];
  }
  tc_wait_ld();
  const int tb = s * U - g.pad;
  float o[8][U];
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int f = 0; f < U; ++f) o[c][f] = __uint_as_float(r[c * U + f]) + bias_s[co0 + c];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    if (co0 + c < p.Cout) {
      float* yr = p.y + ((int64_t)b * p.Cout + co0 + c) * g.Tout;
      if (U >= 4 && tb >= 0 && tb + U <= g.Tout && ((tb & 3) == 0)) {
#pragma unroll
        for (int f = 0; f < U; f += 4)
          *reinterpret_cast<float4*>(yr + tb + f) = make_float4(o[c][f], o[c][f + 1], o[c][f + 2], o[c][f + 3]);
      } else {
#pragma unroll
        for (int f = 0; f < U; ++f)
          if (tb + f >= 0 && tb + f < g.Tout) yr[tb + f] = o[c][f];
      }
    }
  }
  if (p.yimg != nullptr) {
    const int c8n = (p.Cout + 15) >> 4 << 1;
    uint16_t* yi = p.yimg + ((size_t)b * c8n + (size_t)(co0 >> 3)) * (size_t)g.Tout * 8;
#pragma unroll
    for (int f = 0; f < U; ++f) {
      const int t = tb + f;
      if (t >= 0 && t < g.Tout) {
        uint4 q;
        q.x = pack2(lrelu(o[0][f], p.img_slope), lrelu(o[1][f], p.img_slope), bf16);
        q.y = pack2(lrelu(o[2][f], p.img_slope), lrelu(o[3][f], p.img_slope), bf16);
        q.z = pack2(lrelu(o[4][f], p.img_slope), lrelu(o[5][f], p.img_slope), bf16);
        q.w = pack2(lrelu(o[6][f], p.img_slope), lrelu(o[7][f], p.img_slope), bf16);
        *reinterpret_cast<uint4*>(yi + (size_t)t * 8) = q;
      }
    }
  }
}

// MINB = 2: two co-resident CTAs per SM (half the TMEM columns and shared memory each), so that one CTA's operand load
// and store phases run under the other's MMAs — the phases of ONE CTA are serial.
template <int MINB>
__global__ void __launch_bounds__(GC_THREADS, MINB) gemmconv_kernel(GcParams p, GcGeom g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / g.tiles, tile = blockIdx.x - b * g.tiles;
  const int R0 = tile * g.m * 128;   // first output row (time for conv, input time s for convT)
  const int bf16 = p.precision == AB_PREC_TC_BF16;

  const uint32_t sA = smem_u32(smem);
  const uint32_t sW = sA + g.off_w;
  float* bias_s = reinterpret_cast<float*>(smem + g.off_bias);   // [Cout padded]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + g.off_bar);
  const uint32_t bar0 = smem_u32(bars);
  auto bar_full = [&](int s) { return bar0 + 8u * s; };
  auto bar_empty = [&](int s) { return bar0 + 8u * (GC_MAX_STAGES + s); };
  const uint32_t bar_aready = bar0 + 8u * (2 * GC_MAX_STAGES);
  auto bar_accfull = [&](int q) { return bar0 + 8u * (2 * GC_MAX_STAGES + 1 + q); };
  auto bar_accempty = [&](int q) { return bar0 + 8u * (2 * GC_MAX_STAGES + 3 + q); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * GC_MAX_STAGES + 5);

  if (threadIdx.x == 0) {
    for (int s = 0; s < g.nstages; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1);
    }
    mbar_init(bar_aready, GC_WORKERS);
    for (int q = 0; q < 2; ++q) {
      mbar_init(bar_accfull(q), 1);
      mbar_init(bar_accempty(q), GC_WORKERS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
  }
  if (warp == GC_WORKER_WARPS + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u / MINB)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int ncols_buf = g.m * g.Nb;

  if (warp < GC_WORKER_WARPS) {
    // ===================== workers =====================
    const int nbias = g.mode ? g.NB * g.cc : g.NB * g.Nb;
    for (int i = threadIdx.x; i < nbias; i += GC_WORKERS)
      bias_s[i] = (p.bias != nullptr && i < p.Cout) ? __ldg(p.bias + i) : 0.f;
    if (p.ximg != nullptr) {
      // operand image in: register-free cp.async burst, zero fill outside [0, Tin)
      const int c8n = g.Kp >> 3;
      const uint16_t* xb = p.ximg + (size_t)b * c8n * p.Tin * 8;
      for (int c8 = warp; c8 < c8n; c8 += GC_WORKER_WARPS) {
        const uint16_t* xc = xb + (size_t)c8 * p.Tin * 8;
        for (int row = lane; row < g.rowsA; row += 32) {
          const int t = R0 - g.row0_time + row;
          const bool ok = t >= 0 && t < p.Tin;
          cp_async16(sA + unit_offset(g.rowsA, c8, row), ok ? (const void*)(xc + (size_t)t * 8) : (const void*)xb,
                     ok ? 16u : 0u);
        }
      }
      cp_async_wait_all();
    } else {
      const int ngrp = (g.rowsA + 127) >> 7;
      const int c8n = g.Kp >> 3;
      const float* xb = p.x + (int64_t)b * p.xsb;
      for (int item = warp; item < c8n * ngrp; item += GC_WORKER_WARPS) {
        const int c8 = item / ngrp, grp = item - c8 * ngrp;
        const int row0 = (grp << 7) + lane;
        float v[4][8];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int row = row0 + 32 * r;
          const int t = R0 - g.row0_time + row;
          const bool ok = row < g.rowsA && t >= 0 && t < p.Tin;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int c = c8 * 8 + e;
            v[r][e] = (ok && c < p.Cin) ? __ldg(xb + (int64_t)c * p.xsc + (int64_t)t * p.xst) : 0.f;
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
            *reinterpret_cast<uint4*>(smem + unit_offset(g.rowsA, c8, row)) = q;
          }
        }
      }
    }
    fence_proxy_async();
    mbar_arrive(bar_aready);

    const int q4 = warp & 3, hsel = warp >> 2;
    const int nch = g.Nb >> 4;
    for (int nb = 0; nb < g.NB; ++nb) {
      const int buf = nb % g.nbuf;
      mbar_wait(bar_accfull(buf), (uint32_t)(nb / g.nbuf) & 1u, 10);
      tc_fence_after();
      for (int i = 0; i < g.m; ++i) {
        const int row = i * 128 + q4 * 32 + lane;
        const uint32_t tbase = tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(buf * ncols_buf + i * g.Nb);
        if (g.mode == 1 && g.grouped) {
          // 8-channel groups (8*u columns each); this warp takes groups hsel, hsel+2, ...
          const int ngroups = g.cc >> 3;
          for (int grp = hsel; grp < ngroups; grp += 2) {
            const uint32_t ta = tbase + (uint32_t)(grp * 8 * p.u);
            const int co0 = nb * g.cc + grp * 8;
            if (p.u == 8) convT_group_store<8>(p, g, ta, b, co0, R0 + row, bias_s, bf16);
            else if (p.u == 4) convT_group_store<4>(p, g, ta, b, co0, R0 + row, bias_s, bf16);
            else convT_group_store<2>(p, g, ta, b, co0, R0 + row, bias_s, bf16);
          }
          continue;
        }
        for (int ch = hsel; ch < nch; ch += 2) {
          uint32_t r[16];
          tc_ld16(tbase + (uint32_t)(ch * 16), r);
          tc_wait_ld();
          if (g.mode == 0) {
            const int t = R0 + row;
            if (t < p.Tin) {
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                const int co = nb * g.Nb + ch * 16 + e;
                if (co < p.Cout) {
                  const int64_t idx = ((int64_t)b * p.Cout + co) * p.Tin + t;
                  float v = __uint_as_float(r[e]) + bias_s[co];
                  if (p.residual) v += __ldg(p.residual + idx);
                  if (p.post_tanh) v = tanhf(v);
                  p.y[idx] = v;
                }
              }
            }
          } else {
            // generic stride: row = input time s; column n = co_local*u + phi holds y[co, s*u - pad + phi]
            const int tb = (R0 + row) * p.u - g.pad;
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const int n = ch * 16 + e;
              const int col = n / p.u, phi = n - col * p.u;
              const int co = nb * g.cc + col;
              const int t = tb + phi;
              if (col < g.cc && co < p.Cout && t >= 0 && t < g.Tout)
                p.y[((int64_t)b * p.Cout + co) * g.Tout + t] = __uint_as_float(r[e]) + bias_s[co];
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(bar_accempty(buf));
    }
  } else if (warp == GC_WORKER_WARPS) {
    // ===================== TMA weight producer =====================
    if (lane == 0) {
      const int total = g.NB * g.ntaps * g.nkc;
      for (int it = 0; it < total; ++it) {
        const int s = it % g.nstages;
        const uint32_t ph = (uint32_t)(it / g.nstages) & 1u;
        mbar_wait(bar_empty(s), ph ^ 1u, 20);
        mbar_arrive_expect_tx(bar_full(s), g.stage_bytes);
        bulk_g2s(sW + (uint32_t)s * g.stage_bytes, static_cast<const uint8_t*>(p.w) + (size_t)it * g.stage_bytes,
                 g.stage_bytes, bar_full(s));
      }
    }
  } else {
    // ===================== MMA issuer =====================
    const uint32_t elected = elect_one_sync();
    const int nks_total = g.Kp >> 4;
    const uint64_t hi = desc_hi_sw32();
    const uint32_t kstepA = 2u * (uint32_t)g.rowsA, kstepB = 2u * (uint32_t)g.Nb;
    const uint32_t a16 = sA >> 4, w16 = sW >> 4, stage16 = g.stage_bytes >> 4;
    mbar_wait(bar_aready, 0, 30);
    tc_fence_after();
    int it = 0;
    for (int nb = 0; nb < g.NB; ++nb) {
      const int buf = nb % g.nbuf;
      mbar_wait(bar_accempty(buf), ((uint32_t)(nb / g.nbuf) & 1u) ^ 1u, 32);
      tc_fence_after();
      for (int tap = 0; tap < g.ntaps; ++tap) {
        const int shift = g.mode ? (g.ntaps - 1 - tap) : tap * p.d;
        for (int kc = 0; kc < g.nkc; ++kc, ++it) {
          const int s = it % g.nstages;
          const uint32_t ph = (uint32_t)(it / g.nstages) & 1u;
          mbar_wait(bar_full(s), ph, 31);
          tc_fence_after();
          const bool two = nks_total - kc * 2 >= 2;
          uint32_t alo = desc_lo_sw32(a16 + (uint32_t)(kc * 2) * kstepA + (uint32_t)shift * 2u);
          const uint32_t blo = desc_lo_sw32(w16 + (uint32_t)s * stage16);
          const uint32_t acc0 = (tap | kc) != 0 ? 1u : 0u;
          uint32_t td = tmem + (uint32_t)(buf * ncols_buf);
          for (int i = 0; i < g.m; ++i) {
            if (elected) {
              tc_mma_f16(td, hi | alo, hi | blo, g.idesc, acc0);
              if (two) tc_mma_f16(td, hi | (alo + kstepA), hi | (blo + kstepB), g.idesc, 1u);
            }
            alo += 256u;   // 128 rows x 32 B
            td += (uint32_t)g.Nb;
          }
          if (elected) tc_commit(bar_empty(s));
          __syncwarp();
        }
      }
      if (elected) tc_commit(bar_accfull(buf));
      __syncwarp();
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == GC_WORKER_WARPS + 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u / MINB) : "memory");
  }
}

// image: [nb][tap][kc] stages, each [Nb rows x 32 channels] in the operand layout of ab_tc_ptx.cuh
__global__ void gc_pack_weight_kernel(const float* __restrict__ w_t, uint16_t* __restrict__ img, GcGeom g,
                                      int cin, int cout, int k, int d_or_u, int bf16) {
  const int64_t per_stage = (int64_t)g.Nb * 32;
  const int64_t total = (int64_t)g.NB * g.ntaps * g.nkc * per_stage;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int e = (int)(idx & 7);
    int64_t r = idx >> 3;
    const int n = (int)(r % g.Nb);
    r /= g.Nb;
    const int c8l = (int)(r & 3);
    r >>= 2;
    const int kc = (int)(r % g.nkc);
    r /= g.nkc;
    const int tap = (int)(r % g.ntaps);
    const int nb = (int)(r / g.ntaps);
    const int ci = kc * 32 + c8l * 8 + e;
    float v = 0.f;
    if (g.mode == 0) {
      const int co = nb * g.Nb + n;
      if (ci < cin && co < cout) v = w_t[((int64_t)ci * k + tap) * cout + co];
    } else {
      const int u = d_or_u;
      const int col = n / u, phi = n - col * u;
      const int co = nb * g.cc + col;
      const int j = phi + u * tap;
      if (ci < cin && col < g.cc && co < cout && j < k) v = w_t[((int64_t)ci * k + j) * cout + co];
    }
    uint16_t bits;
    if (bf16) {
      __nv_bfloat16 h = __float2bfloat16_rn(v);
      bits = *reinterpret_cast<uint16_t*>(&h);
    } else {
      v = fminf(fmaxf(v, -65504.f), 65504.f);
      __half h = __float2half_rn(v);
      bits = *reinterpret_cast<uint16_t*>(&h);
    }
    const int64_t stage = ((int64_t)nb * g.ntaps + tap) * g.nkc + kc;
    const int unit = (c8l & 1) ^ ((n >> 2) & 1);
    img[stage * per_stage + (int64_t)(c8l >> 1) * g.Nb * 16 + (int64_t)n * 16 + unit * 8 + e] = bits;
  }
}

int rup(int x, int a) { return (x + a - 1) / a * a; }

// geometry that depends only on the layer (not on B / T): used for packing and image sizing
int gc_layer_geom(int mode, int cin, int cout, int k, int d_or_u, GcGeom& g) {
  if (cin <= 0 || cout <= 0 || k <= 0 || d_or_u <= 0) return fail(AB_ERR_ARG, "gemmconv: bad layer shape");
  g.mode = mode;
  g.Kp = rup(cin, 16);
  g.nkc = (g.Kp + 31) / 32;
  if (mode == 0) {
    if (!(k & 1)) return fail(AB_ERR_UNSUPPORTED, "gemmconv: conv kernel size must be odd");
    const int Np = rup(cout, 16);
    g.NB = (Np + 255) / 256;
    g.Nb = rup((Np + g.NB - 1) / g.NB, 16);
    g.cc = g.Nb;
    g.ntaps = k;
    g.pad = 0;
    g.grouped = 0;
  } else {
    const int u = d_or_u;
    if (k < u || ((k - u) & 1)) return fail(AB_ERR_UNSUPPORTED, "gemmconv: conv-transpose needs k >= stride, k-stride even");
    if (u > 64) return fail(AB_ERR_UNSUPPORTED, "gemmconv: stride %d too large", u);
    int cc = 256 / u;
    if (cc > cout) cc = cout;
    g.cc = cc;
    g.Nb = rup(cc * u, 16);
    g.NB = (cout + cc - 1) / cc;
    g.ntaps = (k + u - 1) / u;
    g.pad = (k - u) / 2;
    g.grouped = ((u == 2 || u == 4 || u == 8) && (cc % 8) == 0 && (cout % 8) == 0) ? 1 : 0;
  }
  g.stage_bytes = (uint32_t)g.Nb * 64u;
  // the activation tile is resident: even one M-tile (128 rows + tap reach) must fit next to two weight stages
  const int maxshift = mode ? (g.ntaps - 1) : (k - 1) * d_or_u;
  const uint32_t a1 = (uint32_t)rup(128 + maxshift, 8) * (uint32_t)g.Kp * 2u;
  if (a1 + 2u * g.stage_bytes + 8192u > GC_SMEM_LIMIT)
    return fail(AB_ERR_UNSUPPORTED, "gemmconv: C_in=%d needs the streaming variant (resident tile does not fit)", cin);
  return AB_OK;
}

int gc_full_geom(const GcParams& p, GcGeom& g, int ctas = 1) {
  int rc = gc_layer_geom(p.mode, p.Cin, p.Cout, p.k, p.mode ? p.u : p.d, g);
  if (rc != AB_OK) return rc;
  g.ctas = ctas;
  const int tmem_cols = 512 / ctas;
  const uint32_t smem_limit = ctas == 1 ? GC_SMEM_LIMIT : (GC_SMEM_LIMIT - 2048u) / 2u;
  g.nbuf = (g.NB > 1 && 2 * g.Nb <= tmem_cols) ? 2 : 1;
  const int maxshift = p.mode ? (g.ntaps - 1) : (p.k - 1) * p.d;
  g.row0_time = p.mode ? (g.ntaps - 1) : maxshift / 2;
  if (!p.mode && (maxshift & 1)) return fail(AB_ERR_UNSUPPORTED, "gemmconv: (k-1)*dilation must be even");
  const int rows_total = p.mode ? p.Tin + 1 : p.Tin;
  g.Tout = p.mode ? p.Tin * p.u : p.Tin;
  int m = tmem_cols / (g.nbuf * g.Nb);
  if (m < 1) return fail(AB_ERR_UNSUPPORTED, "gemmconv: N block %d too wide", g.Nb);
  if (m > 16) m = 16;
  while (m > 1 && (m - 1) * 128 >= rows_total) --m;
  const int nbias = rup(p.mode ? g.NB * g.cc : g.NB * g.Nb, 4);
  const uint32_t misc = (uint32_t)nbias * 4u + 8u * (2 * GC_MAX_STAGES + 5) + 16u;
  for (;; --m) {
    if (m < 1) return fail(AB_ERR_UNSUPPORTED, "gemmconv: Cin=%d does not fit shared memory", p.Cin);
    g.rowsA = rup(m * 128 + maxshift, 8);
    const uint32_t abytes = (uint32_t)g.rowsA * (uint32_t)g.Kp * 2u;
    if (abytes + 2u * g.stage_bytes + misc + 1280u > smem_limit) continue;
    g.m = m;
    int ns = (int)((smem_limit - abytes - misc - 1280u) / g.stage_bytes);
    g.nstages = std::min(ns, GC_MAX_STAGES);
    g.off_w = (abytes + 1023u) & ~1023u;
    g.off_bias = g.off_w + (uint32_t)g.nstages * g.stage_bytes;
    g.off_bar = (g.off_bias + (uint32_t)nbias * 4u + 15u) & ~15u;
    g.smem_bytes = g.off_bar + 8u * (2 * GC_MAX_STAGES + 5) + 16u;
    break;
  }
  g.tiles = (rows_total + g.m * 128 - 1) / (g.m * 128);
  const uint32_t fmt = p.precision == AB_PREC_TC_BF16 ? 1u : 0u;
  g.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(g.Nb >> 3) << 17) | ((128u >> 4) << 24);
  return AB_OK;
}


// ===========================================================================
// Streaming variant for wide layers (C_in too large for a resident activation tile, e.g. BigVGAN-large
// stage 0/1 with 768 / 384 channels): conv mode only, operand-image input only.
//   for each N block (<= 256 output channels, TMEM double buffered):
//     for each K chunk (<= 256 input channels): loader warps cp.async the chunk [rows x kchunk] of the
//       image into one of two smem buffers; the MMA warp sweeps the k taps x kchunk/16 K-steps over it.
// Warp roles: 0-3 epilogue, 4-7 activation loaders, 8 TMA weight producer, 9 MMA issuer.
// ===========================================================================
struct GsGeom {
  int Kp, kchunk, KA, nkc_l;   // padded C_in, channels per chunk, chunks, 32-channel stages per chunk
  int Nb, NB, nbuf;
  int ntaps, rowsA, tiles, nstages, hh;
  uint32_t chunk_bytes, stage_bytes, off_w, off_bias, off_bar, smem_bytes;
  uint32_t idesc;
  float out_scale;
  int mode, cc, pad, grouped, Tout;   // conv-transpose: channels per N block, (k-u)/2, 8-channel-group epilogue
};

__global__ void __launch_bounds__(320, 1) gemmconv_stream_kernel(GsParams p, GsGeom g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / g.tiles, tile = blockIdx.x - b * g.tiles;
  const int R0 = tile * 128;
  const uint32_t sA = smem_u32(smem);
  const uint32_t sW = sA + g.off_w;
  float* bias_s = reinterpret_cast<float*>(smem + g.off_bias);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + g.off_bar);
  const uint32_t bar0 = smem_u32(bars);
  auto bar_full = [&](int s) { return bar0 + 8u * s; };
  auto bar_empty = [&](int s) { return bar0 + 8u * (GC_MAX_STAGES + s); };
  auto bar_afull = [&](int q) { return bar0 + 8u * (2 * GC_MAX_STAGES + q); };
  auto bar_aempty = [&](int q) { return bar0 + 8u * (2 * GC_MAX_STAGES + 2 + q); };
  auto bar_accfull = [&](int q) { return bar0 + 8u * (2 * GC_MAX_STAGES + 4 + q); };
  auto bar_accempty = [&](int q) { return bar0 + 8u * (2 * GC_MAX_STAGES + 6 + q); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * GC_MAX_STAGES + 8);

  if (threadIdx.x == 0) {
    for (int s = 0; s < g.nstages; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1);
    }
    for (int q = 0; q < 2; ++q) {
      mbar_init(bar_afull(q), 128);
      mbar_init(bar_aempty(q), 1);
      mbar_init(bar_accfull(q), 1);
      mbar_init(bar_accempty(q), 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
  }
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < g.NB * g.Nb; i += 320) bias_s[i] = (p.bias != nullptr && i < p.Cout) ? __ldg(p.bias + i) : 0.f;
  const int bf16 = p.precision == AB_PREC_TC_BF16;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < 4) {
    // ===================== epilogue =====================
    const int q4 = warp;
    const int nch = g.Nb >> 4;
    const int row = q4 * 32 + lane;
    const int t = R0 + row;
    const bool ok = t < p.T;
    for (int nb = 0; nb < g.NB; ++nb) {
      const int buf = nb % g.nbuf;
      mbar_wait(bar_accfull(buf), (uint32_t)(nb / g.nbuf) & 1u, 10);
      tc_fence_after();
      const uint32_t tbase = tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(buf * g.Nb);
      if (g.mode == 1) {
        // polyphase conv-transpose: row = input time, 8-channel groups of 8*u columns
        const int ngroups = g.cc >> 3;
        for (int grp = 0; grp < ngroups; ++grp) {
          const uint32_t ta = tbase + (uint32_t)(grp * 8 * p.u);
          const int co0 = nb * g.cc + grp * 8;
          if (p.u == 8) convT_group_store<8>(p, g, ta, b, co0, R0 + row, bias_s, bf16);
          else if (p.u == 4) convT_group_store<4>(p, g, ta, b, co0, R0 + row, bias_s, bf16);
          else convT_group_store<2>(p, g, ta, b, co0, R0 + row, bias_s, bf16);
        }
      } else {
      for (int ch = 0; ch < nch; ++ch) {
        uint32_t r[16];
        tc_ld16(tbase + (uint32_t)(ch * 16), r);
        float res[16], acp[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int co = nb * g.Nb + ch * 16 + e;
          const bool w = ok && co < p.Cout;
          const int64_t off = ((int64_t)b * p.Cout + co) * p.T + t;
          res[e] = (w && p.residual) ? __ldg(p.residual + off) : 0.f;
          acp[e] = (w && p.acc_prev) ? __ldg(p.acc_prev + off) : 0.f;
        }
        tc_wait_ld();
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int co = nb * g.Nb + ch * 16 + e;
          if (ok && co < p.Cout) {
            float a = __uint_as_float(r[e]) + bias_s[co];
            a += res[e];
            a += acp[e];
            a *= g.out_scale;
            p.y[((int64_t)b * p.Cout + co) * p.T + t] = a;
          }
        }
      }
      }
      tc_fence_before();
      mbar_arrive(bar_accempty(buf));
    }
  } else if (warp < 8) {
    // ===================== activation loaders (cp.async from the operand image) =====================
    const int lw = warp - 4;
    const int c8_img = (p.Cin + 15) >> 4 << 1;           // 8-channel groups present in the image
    const int c8_chunk = g.kchunk >> 3;
    const uint16_t* xb = p.ximg ? p.ximg + (size_t)b * c8_img * p.T * 8 : nullptr;
    const int total = g.NB * g.KA;
    for (int it = 0; it < total; ++it) {
      const int ab = it & 1, kcA = it % g.KA;
      mbar_wait(bar_aempty(ab), (((uint32_t)(it >> 1)) & 1u) ^ 1u, 40);
      const uint32_t base = sA + (uint32_t)ab * g.chunk_bytes;
      if (p.ximg != nullptr) {
        for (int c8 = lw; c8 < c8_chunk; c8 += 4) {
          const int c8g = kcA * c8_chunk + c8;
          const uint16_t* xc = xb + (size_t)c8g * p.T * 8;
          for (int r = lane; r < g.rowsA; r += 32) {
            const int tt = R0 - g.hh + r;
            const bool okl = c8g < c8_img && tt >= 0 && tt < p.T;
            cp_async16(base + unit_offset(g.rowsA, c8, r), okl ? (const void*)(xc + (size_t)tt * 8) : (const void*)xb,
                       okl ? 16u : 0u);
          }
        }
        cp_async_wait_all();
      } else {
        // fp32 input: load, activate, convert (32 independent loads in flight per lane)
        const float* xf = p.x + (int64_t)b * p.Cin * p.T;
        const int ngrp = (g.rowsA + 127) >> 7;
        for (int item = lw; item < c8_chunk * ngrp; item += 4) {
          const int c8 = item / ngrp, grp = item - c8 * ngrp;
          const int row0 = (grp << 7) + lane;
          float v[4][8];
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int rr = row0 + 32 * r;
            const int tt = R0 - g.hh + rr;
            const bool okl = rr < g.rowsA && tt >= 0 && tt < p.T;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int c = kcA * g.kchunk + c8 * 8 + e;
              v[r][e] = (okl && c < p.Cin) ? __ldg(xf + (int64_t)c * p.T + tt) : 0.f;
            }
          }
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int rr = row0 + 32 * r;
            if (rr < g.rowsA) {
              uint4 q;
              q.x = pack2(lrelu(v[r][0], p.pre_slope), lrelu(v[r][1], p.pre_slope), bf16);
              q.y = pack2(lrelu(v[r][2], p.pre_slope), lrelu(v[r][3], p.pre_slope), bf16);
              q.z = pack2(lrelu(v[r][4], p.pre_slope), lrelu(v[r][5], p.pre_slope), bf16);
              q.w = pack2(lrelu(v[r][6], p.pre_slope), lrelu(v[r][7], p.pre_slope), bf16);
              *reinterpret_cast<uint4*>(smem + (size_t)ab * g.chunk_bytes + unit_offset(g.rowsA, c8, rr)) = q;
            }
          }
        }
      }
      fence_proxy_async();
      mbar_arrive(bar_afull(ab));
    }
  } else if (warp == 8) {
    // ===================== TMA weight producer =====================
    if (lane == 0) {
      const int total = g.NB * g.KA * g.ntaps * g.nkc_l;
      for (int it = 0; it < total; ++it) {
        const int s = it % g.nstages;
        const uint32_t ph = (uint32_t)(it / g.nstages) & 1u;
        mbar_wait(bar_empty(s), ph ^ 1u, 20);
        mbar_arrive_expect_tx(bar_full(s), g.stage_bytes);
        bulk_g2s(sW + (uint32_t)s * g.stage_bytes, static_cast<const uint8_t*>(p.w) + (size_t)it * g.stage_bytes,
                 g.stage_bytes, bar_full(s));
      }
    }
  } else {
    // ===================== MMA issuer =====================
    const uint32_t elected = elect_one_sync();
    const uint64_t hi = desc_hi_sw32();
    const uint32_t kstepA = 2u * (uint32_t)g.rowsA, kstepB = 2u * (uint32_t)g.Nb;
    const uint32_t w16 = sW >> 4, stage16 = g.stage_bytes >> 4;
    int it = 0, ita = 0;
    for (int nb = 0; nb < g.NB; ++nb) {
      const int buf = nb % g.nbuf;
      mbar_wait(bar_accempty(buf), ((uint32_t)(nb / g.nbuf) & 1u) ^ 1u, 32);
      tc_fence_after();
      const uint32_t td = tmem + (uint32_t)(buf * g.Nb);
      for (int kcA = 0; kcA < g.KA; ++kcA, ++ita) {
        const int ab = ita & 1;
        mbar_wait(bar_afull(ab), ((uint32_t)(ita >> 1)) & 1u, 33);
        tc_fence_after();
        const uint32_t a16 = (sA + (uint32_t)ab * g.chunk_bytes) >> 4;
        for (int tap = 0; tap < g.ntaps; ++tap) {
          for (int kc = 0; kc < g.nkc_l; ++kc, ++it) {
            const int s = it % g.nstages;
            const uint32_t ph = (uint32_t)(it / g.nstages) & 1u;
            mbar_wait(bar_full(s), ph, 31);
            tc_fence_after();
            const int shift = g.mode ? (g.ntaps - 1 - tap) : tap * p.d;
            const uint32_t alo = desc_lo_sw32(a16 + (uint32_t)(kc * 2) * kstepA + (uint32_t)shift * 2u);
            const uint32_t blo = desc_lo_sw32(w16 + (uint32_t)s * stage16);
            if (elected) {
              tc_mma_f16(td, hi | alo, hi | blo, g.idesc, (kcA | tap | kc) != 0 ? 1u : 0u);
              tc_mma_f16(td, hi | (alo + kstepA), hi | (blo + kstepB), g.idesc, 1u);
              tc_commit(bar_empty(s));
            }
            __syncwarp();
          }
        }
        if (elected) tc_commit(bar_aempty(ab));
        __syncwarp();
      }
      if (elected) tc_commit(bar_accfull(buf));
      __syncwarp();
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// image: [nb][kcA][tap][kc_l] stages, each [Nb rows x 32 channels] (SWIZZLE_32B rows)
__global__ void gs_pack_weight_kernel(const float* __restrict__ w_t, uint16_t* __restrict__ img, GsGeom g, int cin,
                                      int cout, int k, int u, int bf16) {
  const int64_t per_stage = (int64_t)g.Nb * 32;
  const int64_t total = (int64_t)g.NB * g.KA * g.ntaps * g.nkc_l * per_stage;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int e = (int)(idx & 7);
    int64_t r = idx >> 3;
    const int n = (int)(r % g.Nb);
    r /= g.Nb;
    const int c8l = (int)(r & 3);
    r >>= 2;
    const int kc = (int)(r % g.nkc_l);
    r /= g.nkc_l;
    const int tap = (int)(r % g.ntaps);
    r /= g.ntaps;
    const int kcA = (int)(r % g.KA);
    const int nb = (int)(r / g.KA);
    const int ci = kcA * g.kchunk + kc * 32 + c8l * 8 + e;
    float v = 0.f;
    if (g.mode == 0) {
      const int co = nb * g.Nb + n;
      if (ci < cin && co < cout) v = w_t[((int64_t)ci * k + tap) * cout + co];
    } else {
      const int col = n / u, phi = n - col * u;
      const int co = nb * g.cc + col;
      const int j = phi + u * tap;
      if (ci < cin && col < g.cc && co < cout && j < k) v = w_t[((int64_t)ci * k + j) * cout + co];
    }
    const int64_t stage = (((int64_t)nb * g.KA + kcA) * g.ntaps + tap) * g.nkc_l + kc;
    const int unit = (c8l & 1) ^ ((n >> 2) & 1);
    img[stage * per_stage + (int64_t)(c8l >> 1) * g.Nb * 16 + (int64_t)n * 16 + unit * 8 + e] =
        (uint16_t)(pack2(v, 0.f, bf16) & 0xffffu);
  }
}

int gs_layer_geom(int mode, int cin, int cout, int k, int d_or_u, GsGeom& g) {
  if (cin <= 0 || cout <= 0 || k <= 0 || d_or_u <= 0) return fail(AB_ERR_UNSUPPORTED, "gemmconv(stream): bad layer");
  g.mode = mode;
  g.Kp = rup(cin, 16);
  int maxshift;
  if (mode == 0) {
    if (!(k & 1)) return fail(AB_ERR_UNSUPPORTED, "gemmconv(stream): conv kernel size must be odd");
    const int Np = rup(cout, 16);
    g.NB = (Np + 255) / 256;
    g.Nb = rup((Np + g.NB - 1) / g.NB, 16);
    g.cc = g.Nb;
    g.ntaps = k;
    g.pad = 0;
    g.grouped = 0;
    maxshift = (k - 1) * d_or_u;
    if (maxshift & 1) return fail(AB_ERR_UNSUPPORTED, "gemmconv(stream): (k-1)*dilation must be even");
    g.hh = maxshift / 2;
  } else {
    const int u = d_or_u;
    if (k < u || ((k - u) & 1)) return fail(AB_ERR_UNSUPPORTED, "gemmconv(stream): conv-transpose needs k >= stride, k-stride even");
    if (!(u == 2 || u == 4 || u == 8)) return fail(AB_ERR_UNSUPPORTED, "gemmconv(stream): stride must be 2, 4 or 8");
    int cc = std::min(256 / u, cout);
    if ((cc % 8) || (cout % 8)) return fail(AB_ERR_UNSUPPORTED, "gemmconv(stream): channels must be a multiple of 8");
    g.cc = cc;
    g.Nb = rup(cc * u, 16);
    g.NB = (cout + cc - 1) / cc;
    g.ntaps = (k + u - 1) / u;
    g.pad = (k - u) / 2;
    g.grouped = 1;
    maxshift = g.ntaps - 1;
    g.hh = g.ntaps - 1;
  }
  g.nbuf = g.NB > 1 ? 2 : 1;
  g.rowsA = rup(128 + maxshift, 8);
  g.stage_bytes = (uint32_t)g.Nb * 64u;
  const uint32_t misc = (uint32_t)(g.NB * g.Nb) * 4u + 8u * (2 * GC_MAX_STAGES + 8) + 16u;
  for (int kchunk = 256; kchunk >= 32; kchunk -= 32) {
    const uint32_t cb = ((uint32_t)g.rowsA * (uint32_t)kchunk * 2u + 1023u) & ~1023u;
    if (2u * cb + 2u * g.stage_bytes + misc + 1280u > GC_SMEM_LIMIT) continue;
    g.kchunk = kchunk;
    g.chunk_bytes = cb;
    g.KA = (g.Kp + kchunk - 1) / kchunk;
    g.nkc_l = kchunk / 32;
    g.nstages = std::min((int)((GC_SMEM_LIMIT - 2u * cb - misc - 1280u) / g.stage_bytes), GC_MAX_STAGES);
    g.off_w = 2u * cb;
    g.off_bias = g.off_w + (uint32_t)g.nstages * g.stage_bytes;
    g.off_bar = (g.off_bias + (uint32_t)(g.NB * g.Nb) * 4u + 15u) & ~15u;
    g.smem_bytes = g.off_bar + 8u * (2 * GC_MAX_STAGES + 8) + 16u;
    return AB_OK;
  }
  return fail(AB_ERR_UNSUPPORTED, "gemmconv(stream): k=%d d|u=%d does not fit shared memory", k, d_or_u);
}

}  // namespace

bool gs_can_emit_image(int cout, int k, int u) {
  GsGeom g;
  return gs_layer_geom(1, 16, cout, k, u, g) == AB_OK && (cout % 16) == 0;
}

size_t gs_weight_image_bytes(int mode, int cin, int cout, int k, int d_or_u) {
  GsGeom g;
  if (gs_layer_geom(mode, cin, cout, k, d_or_u, g) != AB_OK) return 0;
  return (size_t)g.NB * g.KA * g.ntaps * g.nkc_l * g.stage_bytes;
}

int launch_gs_pack_weight(const float* w_t, void* image, int mode, int cin, int cout, int k, int d_or_u,
                          int precision, cudaStream_t s) {
  GsGeom g;
  int rc = gs_layer_geom(mode, cin, cout, k, d_or_u, g);
  if (rc != AB_OK) return rc;
  const int64_t total = (int64_t)g.NB * g.KA * g.ntaps * g.nkc_l * g.Nb * 32;
  const int blocks = (int)std::min<int64_t>((total + 255) / 256, 148 * 8);
  gs_pack_weight_kernel<<<blocks, 256, 0, s>>>(w_t, static_cast<uint16_t*>(image), g, cin, cout, k, d_or_u,
                                               precision == AB_PREC_TC_BF16 ? 1 : 0);
  AB_LAUNCH_CHECK("gs_pack_weight_kernel");
  return AB_OK;
}

int launch_gemmconv_stream(const GsParams& p, cudaStream_t s) {
  if ((!p.ximg && !p.x) || !p.y || !p.w) return fail(AB_ERR_ARG, "gemmconv(stream): null argument");
  if (p.B <= 0 || p.T <= 0) return fail(AB_ERR_ARG, "gemmconv(stream): bad shape");
  GsGeom g;
  int rc = gs_layer_geom(p.mode, p.Cin, p.Cout, p.k, p.mode ? p.u : p.d, g);
  if (rc != AB_OK) return rc;
  if (p.yimg != nullptr && !(p.mode == 1 && (p.Cout % 16) == 0))
    return fail(AB_ERR_UNSUPPORTED, "gemmconv(stream): cannot emit an operand image for this layer");
  g.Tout = p.mode ? p.T * p.u : p.T;
  g.tiles = ((p.mode ? p.T + 1 : p.T) + 127) / 128;
  const uint32_t fmt = p.precision == AB_PREC_TC_BF16 ? 1u : 0u;
  g.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(g.Nb >> 3) << 17) | ((128u >> 4) << 24);
  g.out_scale = 1.0f / p.out_div;
  static DeviceOnce configured;
  if (configured.need()) {
    AB_CUDA_TRY(cudaFuncSetAttribute(gemmconv_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GC_SMEM_LIMIT));
  }
  const int64_t grid = (int64_t)p.B * g.tiles;
  if (grid > 0x7fffffffll) return fail(AB_ERR_UNSUPPORTED, "gemmconv(stream): grid too large");
  const uint32_t smem = std::max<uint32_t>(g.smem_bytes, 120u * 1024u);
  gemmconv_stream_kernel<<<(unsigned)grid, 320, smem, s>>>(p, g);
  AB_LAUNCH_CHECK("gemmconv_stream_kernel");
  return AB_OK;
}

bool gc_can_emit_image(int cout, int k, int u) {
  GcGeom g;
  if (gc_layer_geom(1, 16, cout, k, u, g) != AB_OK) return false;
  return g.grouped && (cout % 16) == 0;
}

size_t gc_weight_image_bytes(int mode, int cin, int cout, int k, int d_or_u) {
  GcGeom g;
  if (gc_layer_geom(mode, cin, cout, k, d_or_u, g) != AB_OK) return 0;
  return (size_t)g.NB * g.ntaps * g.nkc * g.stage_bytes;
}

int launch_gc_pack_weight(const float* w_t, void* image, int mode, int cin, int cout, int k, int d_or_u,
                          int precision, cudaStream_t s) {
  GcGeom g;
  int rc = gc_layer_geom(mode, cin, cout, k, d_or_u, g);
  if (rc != AB_OK) return rc;
  const int64_t total = (int64_t)g.NB * g.ntaps * g.nkc * g.Nb * 32;
  const int blocks = (int)std::min<int64_t>((total + 255) / 256, 148 * 8);
  gc_pack_weight_kernel<<<blocks, 256, 0, s>>>(w_t, static_cast<uint16_t*>(image), g, cin, cout, k, d_or_u,
                                               precision == AB_PREC_TC_BF16 ? 1 : 0);
  AB_LAUNCH_CHECK("gc_pack_weight_kernel");
  return AB_OK;
}

int launch_gemmconv(const GcParams& p, cudaStream_t s) {
  if ((!p.x && !p.ximg) || !p.y || !p.w) return fail(AB_ERR_ARG, "gemmconv: null argument");
  if (p.ximg && (p.Cin % 16) != 0) return fail(AB_ERR_UNSUPPORTED, "gemmconv: operand-image input needs C_in %% 16 == 0");
  if (p.B <= 0 || p.Tin <= 0) return fail(AB_ERR_ARG, "gemmconv: bad shape");
  if (p.precision != AB_PREC_TC_F16 && p.precision != AB_PREC_TC_BF16) return fail(AB_ERR_ARG, "gemmconv: bad precision");
  // AB_GC_CTAS (debug / A-B knob): 1 = one CTA per SM with the whole TMEM (round-1 layout), 2 = two half-size CTAs
  // where the layer fits (default)
  static const int want_ctas = [] { const char* e = getenv("AB_GC_CTAS"); return e ? atoi(e) : 2; }();
  GcGeom g;
  int rc = AB_ERR_UNSUPPORTED;
  if (want_ctas >= 2) {
    rc = gc_full_geom(p, g, 2);
    // enough tiles to fill both slots of every SM, and rows per tile not cut below one M tile's worth of work
    if (rc == AB_OK && (int64_t)p.B * g.tiles < 2 * 148) rc = AB_ERR_UNSUPPORTED;
  }
  if (rc != AB_OK) rc = gc_full_geom(p, g, 1);
  if (rc != AB_OK) return rc;
  if (p.yimg != nullptr && !(p.mode == 1 && g.grouped && (p.Cout % 16) == 0))
    return fail(AB_ERR_UNSUPPORTED, "gemmconv: cannot emit an operand image for this layer");
  static DeviceOnce configured;
  if (configured.need()) {
    AB_CUDA_TRY(cudaFuncSetAttribute(gemmconv_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GC_SMEM_LIMIT));
    AB_CUDA_TRY(cudaFuncSetAttribute(gemmconv_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GC_SMEM_LIMIT / 2));
  }
  const int64_t grid = (int64_t)p.B * g.tiles;
  if (grid > 0x7fffffffll) return fail(AB_ERR_UNSUPPORTED, "gemmconv: grid too large");
  if (g.ctas == 2) {
    // at least 76 KB: never more than two CTAs per SM (each allocates 256 TMEM columns)
    const uint32_t smem = std::max<uint32_t>(g.smem_bytes, 76u * 1024u);
    gemmconv_kernel<2><<<(unsigned)grid, GC_THREADS, smem, s>>>(p, g);
  } else {
    const uint32_t smem = std::max<uint32_t>(g.smem_bytes, 120u * 1024u);
    gemmconv_kernel<1><<<(unsigned)grid, GC_THREADS, smem, s>>>(p, g);
  }
  AB_LAUNCH_CHECK("gemmconv_kernel");
  return AB_OK;
}

}  // namespace ab
