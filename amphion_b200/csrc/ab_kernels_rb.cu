// Persistent fused ResBlock kernel (sm_100a): a whole ResBlock1 / ResBlock2 chain per time tile.
//
//   for p in pairs:  x <- x + c2_p(lrelu(c1_p(lrelu(x, .1)) + b1_p, .1)) + b2_p     (hifigan.py:93-100)
//                or  x <- x + c_p(lrelu(x, .1)) + b_p                                (hifigan.py:139-144)
//   y = (x_L + branch_sum) / out_div                                                 (hifigan.py:208-214)
//
// One CTA per SM walks over pairs of time tiles (two "slots").  Per slot: one fp16/bf16 operand buffer in
// shared memory (SWIZZLE_32B K-major rows, time = row, a tap shift = a descriptor row offset, as in
// ab_kernels_tc.cu), one fp32 accumulator set in TMEM (256 columns) and one fp32 scratch tile in global memory
// that holds the residual stream x_p between pairs (thread-private, L2-resident: never read by another thread).
// The chain runs on R = 128*m rows with the block's halo on both sides and is recomputed in the halo, so the
// block reads its input once and writes its output once.  Roles:
//   warps 0..11  epilogue / loader warps (3 per TMEM lane quarter): while the tensor pipe runs conv s of one slot
//                they run the epilogue of conv s of the other slot (TMEM -> +bias, lrelu, cvt -> operand buffer,
//                residual -> accumulator init / scratch / output) and load the next tile
//   warp 12      weight producer: cp.async.bulk (TMA 1-D) of pre-swizzled weight stages into an mbarrier ring
//   warp 13      MMA issuer (the highest warp id: the scheduler prefers it over the epilogue warps of its
//                sub-partition): convs alternate between the two slots
// so the tensor pipe never waits for an epilogue that is shorter than a conv.
#include <stdlib.h>

#include <algorithm>
#include <type_traits>
#include <vector>

#include "ab_tc.cuh"
#include "ab_tc_issue.cuh"
#include "ab_tc_ptx.cuh"

namespace ab {

using namespace tcx;

namespace {

constexpr int RB_EPI_WARPS = 12;               // 3 per TMEM lane quarter; 14 warps -> 128 registers per thread (16 warps at 96 registers: slower)
constexpr int RB_PRODUCER_WARP = RB_EPI_WARPS, RB_MMA_WARP = RB_EPI_WARPS + 1;
constexpr int RB_EPI_GROUPS = RB_EPI_WARPS / 4;
constexpr int RB_EPI_THREADS = RB_EPI_WARPS * 32;
constexpr int RB_THREADS = RB_EPI_THREADS + 64;
constexpr int RB_MAX_STAGES = 8;
constexpr int RB_SLOT_COLS = 256;               // TMEM columns per slot
constexpr uint32_t RB_SMEM_LIMIT = 226 * 1024;      // + 1 KB of static shared memory

struct RbGeom {
  int Np, nkc, m, R, G, RB;     // padded channels, 32-channel K chunks, M tiles, rows, guard rows, buffer rows
  int E;                        // rows of the first conv's own halo that are loaded (c * dil[0])
  int Hlo, V;                   // row of the first valid output, valid outputs per tile
  int tiles, ntiles;            // per sequence, total
  int nsteps;                   // convs per tile
  int cps;                      // weight chunks (tap, 32 channels) per stage, in image order (tap-major)
  int split;                    // 1: conv1 -> D1 (cols 0..127 of the slot), every conv2 accumulates into D2 (cols 128..255),
                                //    which starts as x (+ branch sum): the residual stream never leaves TMEM
  int skip;                     // debug timing experiments (AB_RB_DEBUG_SKIP bitmask), results are wrong
  int nstages;
  uint32_t stage_bytes, chunk_bytes;
  uint32_t off_buf1, off_w, off_bias, off_bar, smem_bytes;
  uint32_t idesc;
  int korder;
  float out_scale;
  long long* dbg;
};

constexpr int DBG_SLOTS = 8;   // 0 total, 1 mma wait operand, 2 mma wait weights, 3 epi wait acc, 4 producer wait, 5 tiles

// ---- epilogue pieces (one item = 32 rows x 16 channels of one M tile) ---------------------------------------
__device__ __forceinline__ void load_bias16(float (&bv)[16], const float* b) {   // 64-byte aligned shared memory
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 f = *reinterpret_cast<const float4*>(b + 4 * q);
    bv[4 * q] = f.x; bv[4 * q + 1] = f.y; bv[4 * q + 2] = f.z; bv[4 * q + 3] = f.w;
  }
}

template <int BF16>
__device__ __forceinline__ void store_operand16(uint8_t* buf, int RBrows, int ch, int brow, const float (&v)[16]) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    uint4 q;
    q.x = pack2t<BF16>(v[8 * h + 0], v[8 * h + 1]);
    q.y = pack2t<BF16>(v[8 * h + 2], v[8 * h + 3]);
    q.z = pack2t<BF16>(v[8 * h + 4], v[8 * h + 5]);
    q.w = pack2t<BF16>(v[8 * h + 6], v[8 * h + 7]);
    *reinterpret_cast<uint4*>(buf + unit_offset(RBrows, ch * 2 + h, brow)) = q;
  }
}

// Epilogue arithmetic on packed fp32 pairs (add / mul.f32x2: IEEE round-to-nearest per lane, i.e. the same results as
// the scalar FADD / FMUL they replace, at half the issue slots).
__device__ __forceinline__ void add_bias16(const uint32_t (&r)[16], const float (&bv)[16], float (&a)[16]) {
#pragma unroll
  for (int e = 0; e < 16; e += 2)
    upk2(add2(pk2(__uint_as_float(r[e]), __uint_as_float(r[e + 1])), pk2(bv[e], bv[e + 1])), a[e], a[e + 1]);
}
// leaky_relu for 0 <= slope <= 1: max(v, slope * v)
__device__ __forceinline__ void lrelu16(const float (&a)[16], float slope, float (&v)[16]) {
  const f32x2 s2 = pk2(slope, slope);
#pragma unroll
  for (int e = 0; e < 16; e += 2) {
    float t0, t1;
    upk2(mul2(pk2(a[e], a[e + 1]), s2), t0, t1);
    v[e] = fmaxf(a[e], t0);
    v[e + 1] = fmaxf(a[e + 1], t1);
  }
}

template <int BF16>
__global__ void __launch_bounds__(RB_THREADS, 1) rb_kernel(RbParams p, RbGeom g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t s0 = smem_u32(smem);
  float* bias_s = reinterpret_cast<float*>(smem + g.off_bias);   // [nsteps][Np]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + g.off_bar);
  const uint32_t bar0 = smem_u32(bars);
  auto bar_full = [&](int s) { return bar0 + 8u * s; };
  auto bar_empty = [&](int s) { return bar0 + 8u * (RB_MAX_STAGES + s); };
  auto bar_opnd = [&](int s) { return bar0 + 8u * (2 * RB_MAX_STAGES + s); };
  auto bar_acc = [&](int s) { return bar0 + 8u * (2 * RB_MAX_STAGES + 2 + s); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * RB_MAX_STAGES + 4);
  // per-step parameters are indexed dynamically: keep them in shared memory, not in a local copy of the params
  __shared__ const void* w_s[2 * AB_RB_MAX_PAIRS];
  __shared__ const float* b_s[2 * AB_RB_MAX_PAIRS];
  __shared__ int dil_s[AB_RB_MAX_PAIRS];
  if (threadIdx.x < 2 * AB_RB_MAX_PAIRS) {
#pragma unroll
    for (int i = 0; i < 2 * AB_RB_MAX_PAIRS; ++i)
      if ((int)threadIdx.x == i) { w_s[i] = p.w[i]; b_s[i] = p.bias[i]; }
#pragma unroll
    for (int i = 0; i < AB_RB_MAX_PAIRS; ++i)
      if ((int)threadIdx.x == i) dil_s[i] = p.dil[i];
  }
  __syncthreads();

  if (threadIdx.x == 0) {
    for (int s = 0; s < g.nstages; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_opnd(s), RB_EPI_WARPS);
      mbar_init(bar_acc(s), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
  }
  if (warp == RB_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // biases of every conv of the chain; guard rows of both operand buffers are zeroed once (never written later)
  for (int i = threadIdx.x; i < g.nsteps * g.Np; i += RB_THREADS) {
    const int st = i / g.Np, c = i - st * g.Np;
    const float* src = b_s[st];
    bias_s[i] = (src != nullptr && c < p.C) ? __ldg(src + c) : 0.f;
  }
  // cumulative biases of the residual-closing convs: x_p = D2 + cum[p] when the stream lives in the accumulator
  for (int c = threadIdx.x; c < g.Np; c += RB_THREADS) {
    float acc = 0.f;
    for (int q = 0; q < p.npairs; ++q) {
      const float* src = b_s[q * p.nconv + p.nconv - 1];
      acc += (src != nullptr && c < p.C) ? __ldg(src + c) : 0.f;
      bias_s[(g.nsteps + q) * g.Np + c] = acc;
    }
  }
  {
    const int c8n = g.Np >> 3;
    const int ng = g.RB - g.R;   // guard rows below + above (+ padding)
    for (int u = threadIdx.x; u < 2 * c8n * ng; u += RB_THREADS) {
      const int s = u / (c8n * ng);
      const int r = u - s * (c8n * ng);
      const int c8 = r / ng;
      int row = r - c8 * ng;
      if (row >= g.G) row += g.R;
      *reinterpret_cast<uint4*>(smem + (uint32_t)s * g.off_buf1 + unit_offset(g.RB, c8, row)) = make_uint4(0, 0, 0, 0);
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const long long t_start = g.dbg ? clock64() : 0;
  const int npairs_total = (g.ntiles + 1) >> 1;

  if (warp < RB_EPI_WARPS) {
    // ===================== epilogue / loader warps =====================
    // The residual never goes through the epilogue's critical path: x_p (+ the branch sum before the last conv)
    // is written into the accumulator with tcgen05.st while the slot waits for its next conv, and that conv
    // accumulates on top of it.  The loads that feed the tcgen05.st are issued at the top of an item, before the
    // TMEM read and the arithmetic of that item, and the lines were prefetched into L2 when the tile was loaded.
    const int ew = warp;
    const int q4 = warp & 3;        // TMEM lane quarter this warp may access
    const int grp = ew >> 2;        // items are dealt round-robin to the RB_EPI_GROUPS warps of a quarter
    const int et = threadIdx.x;
    const int nch = g.Np >> 4;
    const int c8n = g.Np >> 3;
    const int nitems = g.m * nch;
    const int grp_i0 = grp / nch, grp_ch0 = grp - grp_i0 * nch;   // the only division of the item cursor
    const int last_pair = p.npairs - 1;
    long long wait_acc = 0;
    int ntile_done = 0;

    // tile load: operand buffer rows [-E, R + E) <- lrelu(x)[b, :, T0 - Hlo + row] as 16-bit K-major units (the
    // first conv's own halo E = c * dil[0] is loaded, not recomputed)
    auto load_tile = [&](int slot, int tile_id) {
      const int b = tile_id / g.tiles, tl = tile_id - b * g.tiles;
      const int tbase = tl * g.V - g.Hlo;
      const int rows = g.R + 2 * g.E;
      uint8_t* buf = smem + (uint32_t)slot * g.off_buf1;
      if (p.ximg != nullptr) {
        const uint32_t sb = s0 + (uint32_t)slot * g.off_buf1;
        const uint16_t* xb = p.ximg + (size_t)b * c8n * p.T * 8;
        for (int u = et; u < c8n * rows; u += RB_EPI_THREADS) {
          const int c8 = u / rows, row = u - c8 * rows - g.E;
          const int t = tbase + row;
          const bool ok = t >= 0 && t < p.T;
          cp_async16(sb + unit_offset(g.RB, c8, row + g.G),
                     ok ? (const void*)(xb + ((size_t)c8 * p.T + (size_t)t) * 8) : (const void*)xb, ok ? 16u : 0u);
        }
      } else {
        const float* xb = p.x + (int64_t)b * p.C * p.T;
        for (int u = et; u < c8n * rows; u += RB_EPI_THREADS) {
          const int c8 = u / rows, row = u - c8 * rows - g.E;
          const int t = tbase + row;
          const bool ok = t >= 0 && t < p.T;
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int c = c8 * 8 + e;
            v[e] = (ok && c < p.C) ? lrelu(__ldg(xb + (int64_t)c * p.T + t), p.slope) : 0.f;
          }
          uint4 q;
          q.x = pack2t<BF16>(v[0], v[1]);
          q.y = pack2t<BF16>(v[2], v[3]);
          q.z = pack2t<BF16>(v[4], v[5]);
          q.w = pack2t<BF16>(v[6], v[7]);
          *reinterpret_cast<uint4*>(buf + unit_offset(g.RB, c8, row + g.G)) = q;
        }
      }
      // L2 prefetch of the fp32 rows the accumulator inits will read (x, and the branch sum)
      const int segs = (g.R + 31) / 32 + 1;
      const int64_t bCT = (int64_t)b * p.C * p.T;
      for (int u = et; u < p.C * segs; u += RB_EPI_THREADS) {
        const int c = u / segs, sgi = u - c * segs;
        int t = tbase + sgi * 32;
        t = t < 0 ? 0 : (t >= p.T ? p.T - 1 : t);
        prefetch_l2(p.x + bCT + (int64_t)c * p.T + t);
        if (p.acc_prev != nullptr) prefetch_l2(p.acc_prev + bCT + (int64_t)c * p.T + t);
      }
    };
    // 16 fp32 x values of one item (rows of this lane), zero outside [0, T) / beyond C
    const bool sk_ld = g.skip & 1, sk_st = g.skip & 2, sk_tm = g.skip & 4, sk_sm = g.skip & 8;   // debug timing only
    // one predicate per item on the fast path (rows in range, all 16 channels exist), pointer increments
    const bool cfull = (p.C & 15) == 0;
    auto load_x16 = [&](float (&d)[16], const float* base, bool inr, int ch) {
      if (sk_ld) inr = false;
      if (inr && (cfull || ch * 16 + 16 <= p.C)) {
        const float* q = base;
#pragma unroll
        for (int e = 0; e < 16; ++e) { d[e] = __ldg(q); q += p.T; }
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) d[e] = (inr && ch * 16 + e < p.C) ? __ldg(base + (int64_t)e * p.T) : 0.f;
      }
    };
    auto load_acp16 = [&](float (&d)[16], const float* base, bool inr, int ch) {   // may alias y: plain loads
      if (sk_ld) inr = false;
      if (inr && (cfull || ch * 16 + 16 <= p.C)) {
        const float* q = base;
#pragma unroll
        for (int e = 0; e < 16; ++e) { d[e] = *q; q += p.T; }
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) d[e] = (inr && ch * 16 + e < p.C) ? base[(int64_t)e * p.T] : 0.f;
      }
    };
    auto tld16 = [&](uint32_t taddr, uint32_t (&r)[16]) {
      if (!sk_tm) tc_ld16(taddr, r);
      else {
#pragma unroll
        for (int e = 0; e < 16; ++e) r[e] = 0u;
      }
    };
    auto tst16 = [&](uint32_t taddr, const uint32_t (&r)[16]) { if (!sk_tm) tc_st16(taddr, r); };
    // accumulator init of a slot from x (nconv == 1: the first conv accumulates on the residual)
    auto init_acc_from_x = [&](int slot, int tile_id, uint32_t col0, bool add_acp) {
      const int b = tile_id / g.tiles, tl = tile_id - b * g.tiles;
      const int tbase = tl * g.V - g.Hlo;
      const uint32_t tslot = tmem + (uint32_t)(slot * RB_SLOT_COLS) + ((uint32_t)(q4 * 32) << 16);
      const int64_t bCT = (int64_t)b * p.C * p.T;
      for (int n = grp; n < nitems; n += RB_EPI_GROUPS) {
        const int i = n / nch, ch = n - i * nch;
        const int row = i * 128 + q4 * 32 + lane;
        const int t = tbase + row;
        const bool inr = t >= 0 && t < p.T;
        const int64_t off0 = bCT + (int64_t)(ch * 16) * p.T + t;
        float res[16];
        load_x16(res, p.x + off0, inr, ch);
        if (add_acp) {
          float acp[16];
          load_acp16(acp, p.acc_prev + off0, inr, ch);
#pragma unroll
          for (int e = 0; e < 16; ++e) res[e] += acp[e];
        }
        uint32_t r[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) r[e] = __float_as_uint(res[e]);
        tst16(tslot + col0 + (uint32_t)(i * g.Np + ch * 16), r);
      }
      tc_wait_st();
    };
    // which tiles get their accumulator pre-loaded at tile-load time: single-conv residual steps, and the split layout
    const bool init_at_load = p.nconv == 1 || g.split;
    const uint32_t init_col = g.split ? (uint32_t)(RB_SLOT_COLS / 2) : 0u;
    const bool init_acp = p.acc_prev != nullptr && !g.split && last_pair == 0;   // split: the branch sum joins in the last phase
    auto publish_operand = [&](int slot) {
      cp_async_wait_all();
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_opnd(slot));
    };

    uint32_t ph_acc = 0u;   // phase parity of bar_acc per slot (bit s)
    {
      const int q0 = blockIdx.x;
      if (q0 < npairs_total) {
        const bool two = 2 * q0 + 1 < g.ntiles;
        load_tile(0, 2 * q0);
        cp_async_commit();
        if (two) load_tile(1, 2 * q0 + 1);
        cp_async_commit();
        if (init_at_load) init_acc_from_x(0, 2 * q0, init_col, init_acp);
        cp_async_wait_group<1>();
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_opnd(0));
        if (two) {
          if (init_at_load) init_acc_from_x(1, 2 * q0 + 1, init_col, init_acp);
          publish_operand(1);
        }
      }
    }
    for (int q = blockIdx.x; q < npairs_total; q += gridDim.x) {
      const int nact = (2 * q + 1 < g.ntiles) ? 2 : 1;
      const int qn = q + gridDim.x;
      for (int step = 0; step < g.nsteps; ++step) {
        const int pair = step / p.nconv;
        const bool is_e1 = p.nconv == 2 && (step & 1) == 0;
        const bool final_step = step == g.nsteps - 1;
        const float* bias_c = bias_s + step * g.Np;
        for (int slot = 0; slot < nact; ++slot) {
          const int tile_id = 2 * q + slot;
          const int b = tile_id / g.tiles, tl = tile_id - b * g.tiles;
          const int tbase = tl * g.V - g.Hlo;
          uint8_t* buf = smem + (uint32_t)slot * g.off_buf1;
          const uint32_t tslot = tmem + (uint32_t)(slot * RB_SLOT_COLS) + ((uint32_t)(q4 * 32) << 16);
          float* scr = p.scratch ? p.scratch + ((size_t)blockIdx.x * 2 + slot) * ((size_t)g.R * g.Np) : nullptr;
          const int64_t bCT = (int64_t)b * p.C * p.T;
          {
            const long long t0 = g.dbg ? clock64() : 0;
            mbar_wait(bar_acc(slot), (ph_acc >> slot) & 1u, 40 + slot);
            if (g.dbg) wait_acc += clock64() - t0;
          }
          ph_acc ^= 1u << slot;
          tc_fence_after();
          // thread-private scratch of this slot: 16 contiguous floats per (item, lane) -> 4 x 16-byte accesses
          float* scr_lane = scr ? scr + ((size_t)q4 * 32 + lane) * 16 : nullptr;
          constexpr size_t scr_item = (size_t)4 * 32 * 16;
          // item cursor without divisions: item n = (M tile i, 16-channel chunk ch), n = grp, grp + G, ...
          struct Cur { int n, i, ch; };
          auto cur_first = [&]() { Cur c; c.n = grp; c.i = grp_i0; c.ch = grp_ch0; return c; };
          auto cur_next = [&](Cur& c) {
            c.n += RB_EPI_GROUPS;
            c.ch += RB_EPI_GROUPS;
            while (c.ch >= nch) { c.ch -= nch; ++c.i; }
          };
          const int rowl = q4 * 32 + lane;            // row of this lane inside an M tile
          const int tl0 = tbase + rowl;               // its time
          // operand-buffer byte offset of (row, 16-channel chunk): unit (2 ch) at off, unit (2 ch + 1) at off ^ 16
          const uint32_t chunk_stride = (uint32_t)g.RB * 32u;
          auto opnd_off = [&](int i, int ch) {
            const uint32_t brow = (uint32_t)(i * 128 + rowl + g.G);
            return (uint32_t)ch * chunk_stride + brow * 32u + (((brow >> 2) & 1u) << 4);
          };
          auto store_opnd = [&](uint32_t off, const float (&v)[16]) {
            if (sk_sm) return;
            uint4 q0, q1;
            q0.x = pack2t<BF16>(v[0], v[1]); q0.y = pack2t<BF16>(v[2], v[3]);
            q0.z = pack2t<BF16>(v[4], v[5]); q0.w = pack2t<BF16>(v[6], v[7]);
            q1.x = pack2t<BF16>(v[8], v[9]); q1.y = pack2t<BF16>(v[10], v[11]);
            q1.z = pack2t<BF16>(v[12], v[13]); q1.w = pack2t<BF16>(v[14], v[15]);
            *reinterpret_cast<uint4*>(buf + off) = q0;
            *reinterpret_cast<uint4*>(buf + (off ^ 16u)) = q1;
          };
          auto load_scr = [&](float (&res)[16], int n) {
            const float4* sp = reinterpret_cast<const float4*>(scr_lane + (size_t)n * scr_item);
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) {
              const float4 f = sk_ld ? make_float4(0.f, 0.f, 0.f, 0.f) : sp[qd];
              res[4 * qd] = f.x; res[4 * qd + 1] = f.y; res[4 * qd + 2] = f.z; res[4 * qd + 3] = f.w;
            }
          };
          if (is_e1) {
            // ---- intermediate = lrelu(conv1 + b1) -> operand buffer, zero outside [0, T) (conv2's zero padding);
            //      accumulator <- x_p (+ branch sum before the last conv): conv2 accumulates on the residual.
            const bool add_acp = pair == last_pair && p.acc_prev != nullptr;
            // TMEM -> bias, lrelu, mask, pack -> operand buffer; then the accumulator is overwritten with `res`
            auto e1_item = [&](const Cur& c, const float (&res)[16]) {
              const bool inr = (unsigned)(tl0 + c.i * 128) < (unsigned)p.T;
              uint32_t r[16];
              const uint32_t taddr = tslot + (uint32_t)(c.i * g.Np + c.ch * 16);
              tld16(taddr, r);
              float bv[16];
              load_bias16(bv, bias_c + c.ch * 16);
              tc_wait_ld();
              // rows outside [0, T) keep the zeros the tile load put there (conv2's zero padding): no store
              if (inr) {
                float a[16], v[16];
                add_bias16(r, bv, a);
                lrelu16(a, p.slope, v);
                store_opnd(opnd_off(c.i, c.ch), v);
              }
#pragma unroll
              for (int e = 0; e < 16; ++e) r[e] = __float_as_uint(res[e]);
              tst16(taddr, r);
            };
            if (g.split) {
              // split layout: D2 already holds the residual stream; only TMEM(D1) -> operand buffer
              for (Cur c = cur_first(); c.n < nitems; cur_next(c)) {
                const bool inr = (unsigned)(tl0 + c.i * 128) < (unsigned)p.T;
                uint32_t r[16];
                tld16(tslot + (uint32_t)(c.i * g.Np + c.ch * 16), r);
                float bv[16];
                load_bias16(bv, bias_c + c.ch * 16);
                tc_wait_ld();
                if (inr) {
                  float a[16], v[16];
                  add_bias16(r, bv, a);
                  lrelu16(a, p.slope, v);
                  store_opnd(opnd_off(c.i, c.ch), v);
                }
              }
            } else if (pair > 0 && !add_acp) {
              // residual from the scratch: the next item's 64 bytes are requested before this item is processed
              float res0[16], res1[16];
              Cur c0 = cur_first();
              if (c0.n < nitems) load_scr(res0, c0.n);
              while (c0.n < nitems) {
                Cur c1 = c0;
                cur_next(c1);
                if (c1.n < nitems) load_scr(res1, c1.n);
                e1_item(c0, res0);
                if (c1.n >= nitems) break;
                c0 = c1;
                cur_next(c0);
                if (c0.n < nitems) load_scr(res0, c0.n);
                e1_item(c1, res1);
              }
            } else {
              for (Cur c = cur_first(); c.n < nitems; cur_next(c)) {
                const int t = tl0 + c.i * 128;
                const bool inr = (unsigned)t < (unsigned)p.T;
                const int64_t off0 = bCT + (int64_t)(c.ch * 16) * p.T + t;
                float res[16];
                if (pair == 0) load_x16(res, p.x + off0, inr, c.ch);
                else load_scr(res, c.n);
                if (add_acp) {
                  float acp[16];
                  load_acp16(acp, p.acc_prev + off0, inr, c.ch);
#pragma unroll
                  for (int e = 0; e < 16; ++e) res[e] += acp[e];
                }
                e1_item(c, res);
              }
            }
            if (!g.split) tc_wait_st();
            publish_operand(slot);
          } else if (!final_step) {
            // ---- x_{p+1} = acc + b (the accumulator started from x_p).  nconv == 2: -> scratch for the next
            //      pair's accumulator init; nconv == 1: written straight back as the next conv's start value
            //      (+ the branch sum before the last conv).  lrelu(x_{p+1}) -> operand buffer.
            const bool add_acp = p.nconv == 1 && pair + 1 == last_pair && p.acc_prev != nullptr;
            if (g.split) {
              const float* cum = bias_s + (g.nsteps + pair) * g.Np;
              for (Cur c = cur_first(); c.n < nitems; cur_next(c)) {
                const bool inr = (unsigned)(tl0 + c.i * 128) < (unsigned)p.T;
                uint32_t r[16];
                tld16(tslot + (uint32_t)(RB_SLOT_COLS / 2) + (uint32_t)(c.i * g.Np + c.ch * 16), r);
                float bv[16];
                load_bias16(bv, cum + c.ch * 16);
                tc_wait_ld();
                if (inr) {
                  float a[16], v[16];
                  add_bias16(r, bv, a);
                  lrelu16(a, p.slope, v);
                  store_opnd(opnd_off(c.i, c.ch), v);
                }
              }
            } else
            for (Cur c = cur_first(); c.n < nitems; cur_next(c)) {
              const int t = tl0 + c.i * 128;
              const bool inr = (unsigned)t < (unsigned)p.T;
              float acp[16];
              if (add_acp) load_acp16(acp, p.acc_prev + bCT + (int64_t)(c.ch * 16) * p.T + t, inr, c.ch);
              uint32_t r[16];
              const uint32_t taddr = tslot + (uint32_t)(c.i * g.Np + c.ch * 16);
              tld16(taddr, r);
              float bv[16];
              load_bias16(bv, bias_c + c.ch * 16);
              tc_wait_ld();
              float a[16];
              add_bias16(r, bv, a);
              if (p.nconv == 2) {
                float4* sp = reinterpret_cast<float4*>(scr_lane + (size_t)c.n * scr_item);
#pragma unroll
                for (int qd = 0; qd < 4; ++qd)
                  if (!sk_st) sp[qd] = make_float4(a[4 * qd], a[4 * qd + 1], a[4 * qd + 2], a[4 * qd + 3]);
              } else {
#pragma unroll
                for (int e = 0; e < 16; ++e) r[e] = __float_as_uint(add_acp ? a[e] + acp[e] : a[e]);
                tst16(taddr, r);
              }
              if (inr) {
                float v[16];
                lrelu16(a, p.slope, v);
                store_opnd(opnd_off(c.i, c.ch), v);
              }
            }
            if (p.nconv == 1) tc_wait_st();
            publish_operand(slot);
          } else {
            // ---- last conv of the chain: y = (acc + b) / out_div (+ operand image of lrelu(y, img_slope)); the
            // operand buffer is free, so the next tile's load is issued first
            const int next_tile = 2 * qn + slot;
            const bool have_next = qn < npairs_total && next_tile < g.ntiles;
            if (have_next) load_tile(slot, next_tile);
            for (Cur c = cur_first(); c.n < nitems; cur_next(c)) {
              const int row = c.i * 128 + rowl;
              const int t = tbase + row;
              const bool ok = row >= g.Hlo && row < g.Hlo + g.V && t < p.T;
              const int64_t off0 = bCT + (int64_t)(c.ch * 16) * p.T + t;
              float acp[16];
              const bool fin_acp = g.split && p.acc_prev != nullptr;
              if (fin_acp) load_acp16(acp, p.acc_prev + off0, ok, c.ch);
              uint32_t r[16];
              tld16(tslot + (g.split ? (uint32_t)(RB_SLOT_COLS / 2) : 0u) + (uint32_t)(c.i * g.Np + c.ch * 16), r);
              float bv[16];
              load_bias16(bv, (g.split ? bias_s + (g.nsteps + last_pair) * g.Np : bias_c) + c.ch * 16);
              tc_wait_ld();
              float v[16];
              {
                float a[16];
                add_bias16(r, bv, a);
                const f32x2 sc2 = pk2(g.out_scale, g.out_scale);
#pragma unroll
                for (int e = 0; e < 16; e += 2) {
                  f32x2 a2 = pk2(a[e], a[e + 1]);
                  if (fin_acp) a2 = add2(a2, pk2(acp[e], acp[e + 1]));
                  upk2(mul2(a2, sc2), v[e], v[e + 1]);
                }
              }
              if (ok && !sk_st) {
                float* q = p.y + off0;
                if (cfull || c.ch * 16 + 16 <= p.C) {
#pragma unroll
                  for (int e = 0; e < 16; ++e) { *q = v[e]; q += p.T; }
                } else {
#pragma unroll
                  for (int e = 0; e < 16; ++e)
                    if (c.ch * 16 + e < p.C) q[(int64_t)e * p.T] = v[e];
                }
                if (p.yimg != nullptr) {
                  uint16_t* yi = p.yimg + (((size_t)b * c8n + (size_t)c.ch * 2) * p.T + (size_t)t) * 8;
#pragma unroll
                  float w[16];
                  lrelu16(v, p.img_slope, w);
#pragma unroll
                  for (int h = 0; h < 2; ++h) {
                    uint4 qv;
                    qv.x = pack2t<BF16>(w[8 * h + 0], w[8 * h + 1]);
                    qv.y = pack2t<BF16>(w[8 * h + 2], w[8 * h + 3]);
                    qv.z = pack2t<BF16>(w[8 * h + 4], w[8 * h + 5]);
                    qv.w = pack2t<BF16>(w[8 * h + 6], w[8 * h + 7]);
                    *reinterpret_cast<uint4*>(yi + (size_t)h * p.T * 8) = qv;
                  }
                }
              }
            }
            ++ntile_done;
            if (have_next) {
              if (init_at_load) init_acc_from_x(slot, next_tile, init_col, init_acp);
              publish_operand(slot);
            } else {
              tc_fence_before();
            }
          }
        }
      }
    }
    if (g.dbg != nullptr && threadIdx.x == 0) {
      g.dbg[blockIdx.x * DBG_SLOTS + 0] = clock64() - t_start;
      g.dbg[blockIdx.x * DBG_SLOTS + 3] = wait_acc;
      g.dbg[blockIdx.x * DBG_SLOTS + 5] = ntile_done;
    }
  } else if (warp == RB_PRODUCER_WARP) {
    // ===================== weight producer =====================
    if (lane == 0) {
      const int nchunks = p.k * g.nkc;                        // chunks per conv, image order
      const int per_conv = (nchunks + g.cps - 1) / g.cps;     // stages per conv (the last one may be short)
      int it = 0;
      long long wait_e = 0;
      for (int q = blockIdx.x; q < npairs_total; q += gridDim.x) {
        const int nact = (2 * q + 1 < g.ntiles) ? 2 : 1;
        for (int step = 0; step < g.nsteps; ++step) {
          const uint8_t* wsrc = static_cast<const uint8_t*>(w_s[step]);
          for (int slot = 0; slot < nact; ++slot) {
            for (int l = 0; l < per_conv; ++l, ++it) {
              const int s = it % g.nstages;
              const uint32_t ph = (uint32_t)(it / g.nstages) & 1u;
              const long long t0 = g.dbg ? clock64() : 0;
              mbar_wait(bar_empty(s), ph ^ 1u, 20);
              if (g.dbg) wait_e += clock64() - t0;
              const uint32_t bytes = (uint32_t)min(g.cps, nchunks - l * g.cps) * g.chunk_bytes;
              mbar_arrive_expect_tx(bar_full(s), bytes);
              bulk_g2s(s0 + g.off_w + (uint32_t)s * g.stage_bytes, wsrc + (size_t)l * g.stage_bytes, bytes, bar_full(s));
            }
          }
        }
      }
      if (g.dbg != nullptr) g.dbg[blockIdx.x * DBG_SLOTS + 4] = wait_e;
    }
  } else {
    // ===================== MMA issuer =====================
    const uint32_t elected = elect_one_sync();
    const int nks_total = g.Np >> 4;
    const uint64_t hi = desc_hi_sw32();
    const uint32_t kstepA = 2u * (uint32_t)g.RB;
    const uint32_t kstepB = 2u * (uint32_t)g.Np;
    const uint32_t w16 = (s0 + g.off_w) >> 4, stage16 = g.stage_bytes >> 4, chunk16 = g.chunk_bytes >> 4;
    const int nchunks = p.k * g.nkc;
    const int c = (p.k - 1) >> 1;
    uint32_t ph_opnd = 0u;
    int ring_s = 0;               // weight ring cursor (stage, phase parity)
    uint32_t ring_ph = 0u;
    long long wait_o = 0, wait_w = 0;
    for (int q = blockIdx.x; q < npairs_total; q += gridDim.x) {
      const int nact = (2 * q + 1 < g.ntiles) ? 2 : 1;
      for (int step = 0; step < g.nsteps; ++step) {
        const int pair = step / p.nconv;
        const int dil = (p.nconv == 2 && (step & 1)) ? 1 : dil_s[pair];
        // conv1 of a pair starts from zero; every conv that closes a residual step accumulates on x_p, which the
        // epilogue warps wrote into the accumulator
        const bool zero_init = p.nconv == 2 && (step & 1) == 0;
        for (int slot = 0; slot < nact; ++slot) {
          {
            const long long t0 = g.dbg ? clock64() : 0;
            mbar_wait(bar_opnd(slot), (ph_opnd >> slot) & 1u, 30 + slot);
            if (g.dbg) wait_o += clock64() - t0;
          }
          ph_opnd ^= 1u << slot;
          tc_fence_after();
          const uint32_t a16 = (s0 + (uint32_t)slot * g.off_buf1) >> 4;
          const uint32_t td = tmem + (uint32_t)(slot * RB_SLOT_COLS) + ((g.split && !zero_init) ? (uint32_t)(RB_SLOT_COLS / 2) : 0u);
          auto run_conv = [&](auto issue_first, auto issue_rest) {
            int kc = 0;                               // 32-channel K chunk of the current tap
            uint32_t arow = (uint32_t)(g.G - c * dil) * 2u;
            for (int q0 = 0; q0 < nchunks; q0 += g.cps) {
              {
                const long long t0 = g.dbg ? clock64() : 0;
                mbar_wait(bar_full(ring_s), ring_ph, 31);
                if (g.dbg) wait_w += clock64() - t0;
              }
              tc_fence_after();
              const int nq = min(g.cps, nchunks - q0);
              uint32_t blo = desc_lo_sw32(w16 + (uint32_t)ring_s * stage16);
              for (int kk = 0; kk < nq; ++kk) {
                const bool two = nks_total - kc * 2 >= 2;
                const uint32_t alo = desc_lo_sw32(a16 + (uint32_t)(kc * 2) * kstepA + arow);
                if (zero_init && (q0 | kk) == 0) issue_first(alo, blo, two);
                else issue_rest(alo, blo, two);
                blo += chunk16;
                if (++kc == g.nkc) { kc = 0; arow += (uint32_t)dil * 2u; }
              }
              if (elected) tc_commit(bar_empty(ring_s));
              __syncwarp();
              if (++ring_s == g.nstages) { ring_s = 0; ring_ph ^= 1u; }
            }
          };
          auto generic = [&](uint32_t alo, uint32_t blo, bool two, uint32_t acc0) {
            for (int h = 0; h < (two ? 2 : 1); ++h) {
              uint32_t ah = alo + (uint32_t)h * kstepA;
              const uint32_t bh = blo + (uint32_t)h * kstepB;
              uint32_t tdd = td;
              for (int i = 0; i < g.m; ++i) {
                if (elected) tc_mma_f16(tdd, hi | ah, hi | bh, g.idesc, h ? 1u : acc0);
                ah += 256u;
                tdd += (uint32_t)g.Np;
              }
            }
          };
#define AB_RUN(MM, KO)                                                                                            \
  run_conv([&](uint32_t alo, uint32_t blo, bool two) {                                                            \
             issue_stage<MM, KO, 1>(elected, td, (uint32_t)g.Np, hi, alo, blo, kstepA, kstepB, g.idesc, two);     \
           },                                                                                                     \
           [&](uint32_t alo, uint32_t blo, bool two) {                                                            \
             issue_stage<MM, KO, 0>(elected, td, (uint32_t)g.Np, hi, alo, blo, kstepA, kstepB, g.idesc, two);     \
           })
          if (g.m == 2 && !g.korder) AB_RUN(2, false);
          else if (g.m == 4 && g.korder) AB_RUN(4, true);
          else if (g.m == 8 && g.korder) AB_RUN(8, true);
          else if (g.m == 1 && !g.korder) AB_RUN(1, false);
          else if (g.m == 16 && g.korder) AB_RUN(16, true);
          else
            run_conv([&](uint32_t alo, uint32_t blo, bool two) { generic(alo, blo, two, 0u); },
                     [&](uint32_t alo, uint32_t blo, bool two) { generic(alo, blo, two, 1u); });
#undef AB_RUN
          if (elected) tc_commit(bar_acc(slot));
          __syncwarp();
        }
      }
    }
    if (g.dbg != nullptr && elected) {
      g.dbg[blockIdx.x * DBG_SLOTS + 1] = wait_o;
      g.dbg[blockIdx.x * DBG_SLOTS + 2] = wait_w;
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == RB_MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

int rb_round_up(int x, int a) { return (x + a - 1) / a * a; }

// halo (one side) of a chain of `npairs` pairs
int rb_halo(int k, const int* dil, int npairs, int nconv) {
  const int c = (k - 1) / 2;
  int h = 0;
  for (int i = 0; i < npairs; ++i) h += c * (dil[i] + (nconv == 2 ? 1 : 0));
  return h;
}

int rb_make_geom(const RbParams& p, RbGeom& g) {
  if (p.C <= 0 || rb_round_up(p.C, 16) > 128) return fail(AB_ERR_UNSUPPORTED, "rb: C=%d not in [1,128]", p.C);
  if (p.k <= 0 || !(p.k & 1) || p.k > 31) return fail(AB_ERR_UNSUPPORTED, "rb: need odd k <= 31");
  if (p.npairs < 1 || p.npairs > AB_RB_MAX_PAIRS || (p.nconv != 1 && p.nconv != 2)) return fail(AB_ERR_ARG, "rb: bad chain");
  g.Np = rb_round_up(p.C, 16);
  g.nkc = (g.Np + 31) / 32;
  g.split = (p.split && p.nconv == 2 && p.npairs > 1 && g.Np <= 64) ? 1 : 0;
  g.m = std::min((g.split ? RB_SLOT_COLS / 2 : RB_SLOT_COLS) / g.Np, 16);
  g.R = 128 * g.m;
  const int c = (p.k - 1) / 2;
  int dmax = 1;
  for (int i = 0; i < p.npairs; ++i) {
    if (p.dil[i] <= 0) return fail(AB_ERR_ARG, "rb: dilation must be positive");
    dmax = std::max(dmax, p.dil[i]);
  }
  g.G = c * dmax;
  g.RB = rb_round_up(g.R + 2 * g.G, 8);
  g.E = c * p.dil[0];
  const int H = rb_halo(p.k, p.dil, p.npairs, p.nconv) - g.E;   // rows lost on each side: every conv but the first
  g.Hlo = rb_round_up(H, 4);
  g.V = (g.R - g.Hlo - H) / 4 * 4;
  if (g.V < 32) return fail(AB_ERR_UNSUPPORTED, "rb: halo %d leaves no room in a %d-row tile", H, g.R);
  g.tiles = (p.T + g.V - 1) / g.V;
  const int64_t nt = (int64_t)p.B * g.tiles;
  if (nt > 0x3fffffffll) return fail(AB_ERR_UNSUPPORTED, "rb: too many tiles");
  g.ntiles = (int)nt;
  g.nsteps = p.npairs * p.nconv;
  g.chunk_bytes = (uint32_t)g.Np * 64u;
  // weight stage = cps consecutive chunks of the image (a chunk = one tap x 32 channels): the issuer pays one
  // barrier test per stage (~85 cycles even when the stage is already there), so a stage carries up to 32 MMAs
  static const int cps_pref = [] { const char* e = getenv("AB_RB_CPS"); return e ? atoi(e) : 0; }();
  const int mma_per_chunk = g.m * 2;
  g.cps = std::max(1, (32 + mma_per_chunk - 1) / mma_per_chunk);
  if (cps_pref > 0) g.cps = cps_pref;
  g.cps = std::min(g.cps, p.k * g.nkc);
  const uint32_t buf_bytes = (uint32_t)g.RB * (uint32_t)g.Np * 2u;
  g.off_buf1 = (buf_bytes + 1023u) & ~1023u;
  g.off_w = (g.off_buf1 + buf_bytes + 1023u) & ~1023u;
  const uint32_t tail = (uint32_t)(g.nsteps + p.npairs) * g.Np * 4u + 8u * (2 * RB_MAX_STAGES + 4) + 32u;
  g.stage_bytes = g.chunk_bytes * (uint32_t)g.cps;
  while (g.cps > 1 && g.off_w + 3u * g.stage_bytes + tail > RB_SMEM_LIMIT) {   // keep at least three stages in the ring
    --g.cps;
    g.stage_bytes = g.chunk_bytes * (uint32_t)g.cps;
  }
  if (g.off_w + 2u * g.stage_bytes + tail > RB_SMEM_LIMIT) return fail(AB_ERR_UNSUPPORTED, "rb: C=%d k=%d does not fit shared memory", p.C, p.k);
  g.nstages = std::min<int>((RB_SMEM_LIMIT - g.off_w - tail) / g.stage_bytes, RB_MAX_STAGES);
  g.off_bias = g.off_w + (uint32_t)g.nstages * g.stage_bytes;
  g.off_bar = (g.off_bias + (uint32_t)(g.nsteps + p.npairs) * g.Np * 4u + 15u) & ~15u;
  g.smem_bytes = g.off_bar + 8u * (2 * RB_MAX_STAGES + 4) + 16u;
  const uint32_t fmt = p.precision == AB_PREC_TC_BF16 ? 1u : 0u;
  g.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(g.Np >> 3) << 17) | ((128u >> 4) << 24);
  g.korder = g.m >= 4 ? 1 : 0;
  g.out_scale = 1.0f / p.out_div;
  g.dbg = nullptr;
  static const int skip = [] { const char* e = getenv("AB_RB_DEBUG_SKIP"); return e ? atoi(e) : 0; }();
  g.skip = skip;
  return AB_OK;
}

}  // namespace

bool rb_supported(int C, int k) { return C > 0 && rb_round_up(C, 16) <= 128 && (k & 1) && k <= 31; }

size_t rb_scratch_bytes() { return (size_t)148 * 2 * (size_t)(RB_SLOT_COLS * 128) * sizeof(float); }

// Modelled cycles per output row of a chain, calibrated on B200 (profiles/r2_rb_timing.txt): per tile the tensor
// pipe needs k * (Np/16) * m MMAs of ~(48 + 0.3 Np) cycles per conv, the epilogue warps ~8 k cycles per phase of a
// full 32 K-element tile (one pair per launch: ~26.5 k, bound by the HBM round trip of x / y / images), and the two
// overlap imperfectly: t = max + 0.35 min.  Used to choose between one fused launch per block and one launch per
// pair.  Returns 0 when the geometry is not served.
double rb_cost_per_row(int C, int k, const int* dil, int npairs, int nconv, int split) {
  if (!rb_supported(C, k)) return 0.0;
  const int Np = rb_round_up(C, 16);
  if (split && !(nconv == 2 && npairs > 1 && Np <= 64)) return 0.0;
  const int m = std::min((split ? RB_SLOT_COLS / 2 : RB_SLOT_COLS) / Np, 16), R = 128 * m;
  const int H = rb_halo(k, dil, npairs, nconv) - (k - 1) / 2 * dil[0];
  const int V = (R - rb_round_up(H, 4) - H) / 4 * 4;
  if (V < 32) return 0.0;
  const double fill = (double)(m * Np) / RB_SLOT_COLS;                       // tile elements / 32768
  const double mma = (double)nconv * npairs * k * (Np / 16.0) * m * (48.0 + 0.3 * Np);
  // an epilogue phase costs ~1.6 k cycles of hand-off latency plus ~0.2 cycles per element (measured: the split layout
  // removes the residual work but halves the tile, so its fixed share doubles — it rarely wins)
  const double phase = 1600.0 + fill * (split ? 6000.0 : 6400.0);
  const double epi = npairs == 1 ? fill * (nconv == 2 ? 26500.0 : 22000.0) : phase * (nconv * npairs + 1);
  return (std::max(mma, epi) + 0.35 * std::min(mma, epi)) / V;
}

int launch_rb(const RbParams& p, cudaStream_t s) {
  if (!p.x || !p.y) return fail(AB_ERR_ARG, "rb: null argument");
  if (p.B <= 0 || p.T <= 0) return fail(AB_ERR_ARG, "rb: bad shape");
  if (p.precision != AB_PREC_TC_F16 && p.precision != AB_PREC_TC_BF16) return fail(AB_ERR_ARG, "rb: bad precision");
  RbGeom g;
  int rc = rb_make_geom(p, g);
  if (rc != AB_OK) return rc;
  for (int i = 0; i < g.nsteps; ++i)
    if (!p.w[i]) return fail(AB_ERR_ARG, "rb: missing weight image %d", i);
  if (p.npairs > 1 && p.nconv == 2 && !g.split && !p.scratch) return fail(AB_ERR_ARG, "rb: a fused chain needs the scratch buffer");
  int dev = 0;
  AB_CUDA_TRY(cudaGetDevice(&dev));
  static DeviceOnce configured;
  if (configured.need()) {
    AB_CUDA_TRY(cudaFuncSetAttribute(rb_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RB_SMEM_LIMIT));
    AB_CUDA_TRY(cudaFuncSetAttribute(rb_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RB_SMEM_LIMIT));
  }
  static int nsm = 0;
  if (!nsm) AB_CUDA_TRY(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
  const int npairs_total = (g.ntiles + 1) / 2;
  const int grid = std::min(std::min(nsm, 148), npairs_total);
  // > half of the SM's shared memory: exactly one CTA (all 512 TMEM columns) per SM
  const uint32_t smem = std::max<uint32_t>(g.smem_bytes, 120u * 1024u);
  static const bool dbg_on = [] { const char* e = getenv("AB_RB_DEBUG_TIMING"); return e && e[0] == '1'; }();
  static long long* dbg_buf = nullptr;
  if (dbg_on) {   // debug only: the one place this file allocates
    if (!dbg_buf) AB_CUDA_TRY(cudaMalloc(&dbg_buf, sizeof(long long) * 148 * DBG_SLOTS));
    AB_CUDA_TRY(cudaMemsetAsync(dbg_buf, 0, sizeof(long long) * 148 * DBG_SLOTS, s));
    g.dbg = dbg_buf;
  }
  if (p.precision == AB_PREC_TC_BF16) rb_kernel<1><<<grid, RB_THREADS, smem, s>>>(p, g);
  else rb_kernel<0><<<grid, RB_THREADS, smem, s>>>(p, g);
  AB_LAUNCH_CHECK("rb_kernel");
  if (dbg_on) {
    AB_CUDA_TRY(cudaStreamSynchronize(s));
    std::vector<long long> h(148 * DBG_SLOTS);
    AB_CUDA_TRY(cudaMemcpy(h.data(), dbg_buf, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost));
    double a[DBG_SLOTS] = {};
    for (int i = 0; i < grid; ++i)
      for (int j = 0; j < DBG_SLOTS; ++j) a[j] += (double)h[(size_t)i * DBG_SLOTS + j] / grid;
    const double ideal = (double)g.m * (g.Np / 2.0) * (g.Np / 16.0) * p.k * g.nsteps * a[5];   // 8192 flop/clk/SM
    const double law = (double)g.m * (64.0 + g.Np / 2.0) * (g.Np / 16.0) * p.k * g.nsteps * a[5];
    fprintf(stderr,
            "[rb_timing] C=%d k=%d d=%d,%d,%d npairs=%d nconv=%d split=%d m=%d R=%d V=%d tiles=%d grid=%d stages=%dx%uB (%d chunks) | per CTA: "
            "total %.0f cycles, %.1f tiles | MMA ideal %.0f law %.0f | issuer waits: operand %.0f weights %.0f | "
            "epilogue waits acc %.0f | producer waits %.0f\n",
            p.C, p.k, p.dil[0], p.npairs > 1 ? p.dil[1] : 0, p.npairs > 2 ? p.dil[2] : 0, p.npairs, p.nconv, g.split, g.m, g.R, g.V,
            g.ntiles, grid, g.nstages, g.stage_bytes, g.cps, a[0], a[5], ideal, law, a[1], a[2], a[3], a[4]);
  }
  return AB_OK;
}

}  // namespace ab
