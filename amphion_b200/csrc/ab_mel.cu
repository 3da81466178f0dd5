// Mel front end: reflect-pad + frame + window  ->  cuFFT R2C  ->  fused
// magnitude / mel filterbank / log-clamp / energy.
//   utils/mel.py:20-170 (extract_linear_features, mel_spectrogram_torch,
//   extract_mel_features) and utils/stft.py:152-181,259-278 (TacotronSTFT).
// The framing kernel performs the same fp32 multiply torch.stft performs
// (frames * window) and the FFT is the libcufft.so.11 already mapped by torch,
// so the complex spectrum is bit-identical to torch.stft's on the same device.
#include <cufft.h>

#include <algorithm>
#include <map>
#include <vector>

#include "ab_common.cuh"

using namespace ab;

struct ab_mel {
  ab_mel_config cfg;
  int bins;
  std::map<int64_t, cufftHandle> plans;   // batch (B*F) -> plan; bounded LRU (variable-length corpora)
  std::map<int64_t, size_t> plan_ws;
  std::vector<int64_t> plan_lru;          // most recently used last
};

namespace {

constexpr int FT = 32;  // frames per CTA in the post kernel

// frames[(b*F + f)][n] = reflect(y)[b, f*hop + n - pad] * window_full[n]
// window_full = window centred in n_fft (zero outside), as torch.stft / pad_center do.
__global__ void __launch_bounds__(256) frame_window_kernel(const float* __restrict__ y,
                                                           const float* __restrict__ window,
                                                           float* __restrict__ frames, int T, int F,
                                                           int n_fft, int hop, int win, int pad) {
  const int64_t fr = blockIdx.x;  // b*F + f
  const int b = (int)(fr / F), f = (int)(fr - (int64_t)b * F);
  const float* yb = y + (int64_t)b * T;
  float* out = frames + fr * n_fft;
  const int lpad = (n_fft - win) / 2;
  for (int n = threadIdx.x; n < n_fft; n += 256) {
    int i = f * hop + n - pad;
    if (i < 0) i = -i;
    if (i >= T) i = 2 * (T - 1) - i;
    const int wi = n - lpad;
    const float w = (wi >= 0 && wi < win) ? __ldg(window + wi) : 0.f;
    out[n] = __fmul_rn(__ldg(yb + i), w);
  }
}

// per mel row: [lo, hi) = span of non-zero filter taps (pure work-skipping: adding
// exact zeros does not change a sum)
__global__ void mel_span_kernel(const float* __restrict__ basis, int n_mel, int bins, int2* __restrict__ span) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n_mel) return;
  int lo = bins, hi = 0;
  for (int k = 0; k < bins; ++k) {
    if (basis[(int64_t)m * bins + k] != 0.f) {
      if (k < lo) lo = k;
      hi = k + 1;
    }
  }
  if (hi == 0) lo = 0;
  span[m] = make_int2(lo, hi);
}

// One CTA = FT frames of one batch item.
__global__ void __launch_bounds__(256) mag_mel_kernel(const float2* __restrict__ spec,  // [B*F][bins]
                                                      const float* __restrict__ basis,  // [n_mel][bins]
                                                      const int2* __restrict__ span,
                                                      float* __restrict__ mag_out,      // [B][bins][F] or null
                                                      float* __restrict__ mel_out,      // [B][n_mel][F] or null
                                                      float* __restrict__ energy_out,   // [B][F] or null
                                                      int F, int bins, int n_mel, float eps, float clampv) {
  extern __shared__ float mag_s[];  // [bins][FT+1]
  __shared__ float epart[8][FT];
  const int b = blockIdx.y, f0 = blockIdx.x * FT;
  const int nf = min(FT, F - f0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // magnitude: warp w handles frames w, w+8, ...; lanes sweep bins (coalesced 8 B loads)
  for (int f = warp; f < nf; f += 8) {
    const float2* row = spec + ((int64_t)b * F + f0 + f) * bins;
    for (int k = lane; k < bins; k += 32) {
      const float2 c = __ldg(row + k);
      // torch: spec.pow(2).sum(-1) + eps, then sqrt  (utils/mel.py:165-166)
      const float s = __fadd_rn(__fadd_rn(__fmul_rn(c.x, c.x), __fmul_rn(c.y, c.y)), eps);
      mag_s[k * (FT + 1) + f] = __fsqrt_rn(s);
    }
  }
  __syncthreads();
  if (mag_out != nullptr) {
    for (int k = warp; k < bins; k += 8)
      if (lane < nf) mag_out[((int64_t)b * bins + k) * F + f0 + lane] = mag_s[k * (FT + 1) + lane];
  }
  if (energy_out != nullptr) {
    float e = 0.f;
    if (lane < nf)
      for (int k = warp; k < bins; k += 8) {
        const float v = mag_s[k * (FT + 1) + lane];
        e = fmaf(v, v, e);
      }
    epart[warp][lane] = e;
    __syncthreads();
    if (warp == 0 && lane < nf) {
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) tot += epart[w][lane];
      energy_out[(int64_t)b * F + f0 + lane] = sqrtf(tot);
    }
  }
  if (mel_out != nullptr) {
    for (int m = warp; m < n_mel; m += 8) {
      const int2 sp = span[m];
      const float* brow = basis + (int64_t)m * bins;
      float acc = 0.f;
      if (lane < nf)
        for (int k = sp.x; k < sp.y; ++k) acc = fmaf(__ldg(brow + k), mag_s[k * (FT + 1) + lane], acc);
      if (lane < nf) mel_out[((int64_t)b * n_mel + m) * F + f0 + lane] = logf(fmaxf(acc, clampv));
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Backward of the log-mel output with respect to the waveform (the mel loss of the vocoder trainers,
// gan_vocoder_trainer.py:368-396: L1 between extract_mel_features(y_gt) and extract_mel_features(y_pred)).
//   mel = log(max(acc, clamp)), acc = basis . mag, mag = sqrt(re^2 + im^2 + eps), (re, im) = rfft(frame * window)
// One CTA = FT frames of one batch item, as in mag_mel_kernel.  It turns the spectrum into the cotangent of the
// spectrum IN PLACE, already in the form cuFFT's C2R expects (interior bins halved: C2R doubles them).
__global__ void __launch_bounds__(256) mel_bwd_spec_kernel(float2* __restrict__ spec,          // [B*F][bins], in/out
                                                           const float* __restrict__ basis,   // [n_mel][bins]
                                                           const int2* __restrict__ span,
                                                           const float* __restrict__ gmel,    // [B][n_mel][F]
                                                           int F, int bins, int n_mel, float eps, float clampv) {
  extern __shared__ float bw_s[];           // mag [bins][FT+1], then gm [n_mel][FT+1]
  float* mag_s = bw_s;
  float* gm_s = bw_s + (size_t)bins * (FT + 1);
  const int b = blockIdx.y, f0 = blockIdx.x * FT;
  const int nf = min(FT, F - f0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int f = warp; f < nf; f += 8) {
    const float2* row = spec + ((int64_t)b * F + f0 + f) * bins;
    for (int k = lane; k < bins; k += 32) {
      const float2 c = row[k];
      const float sq = __fadd_rn(__fadd_rn(__fmul_rn(c.x, c.x), __fmul_rn(c.y, c.y)), eps);
      mag_s[k * (FT + 1) + f] = __fsqrt_rn(sq);
    }
  }
  __syncthreads();
  // d log(max(acc, clamp)) = g / acc where acc >= clamp (torch.clamp passes the gradient at the boundary), else 0
  for (int m = warp; m < n_mel; m += 8) {
    const int2 sp = span[m];
    const float* brow = basis + (int64_t)m * bins;
    float acc = 0.f;
    if (lane < nf)
      for (int k = sp.x; k < sp.y; ++k) acc = fmaf(__ldg(brow + k), mag_s[k * (FT + 1) + lane], acc);
    float g = 0.f;
    if (lane < nf && acc >= clampv) g = __ldg(gmel + ((int64_t)b * n_mel + m) * F + f0 + lane) / acc;
    gm_s[m * (FT + 1) + lane] = g;
  }
  __syncthreads();
  // d mag[k] = sum_m basis[m][k] gm[m]; stored as the ratio d mag / mag in place of mag
  for (int k = warp; k < bins; k += 8) {
    float acc = 0.f;
    for (int m = 0; m < n_mel; ++m) {
      const int2 sp = span[m];                                   // warp-uniform
      if (k >= sp.x && k < sp.y) acc = fmaf(__ldg(basis + (int64_t)m * bins + k), gm_s[m * (FT + 1) + lane], acc);
    }
    const float mg = mag_s[k * (FT + 1) + lane];
    mag_s[k * (FT + 1) + lane] = mg > 0.f ? acc / mg : 0.f;      // eps = 0 and a silent bin: torch yields NaN, we 0
  }
  __syncthreads();
  const int nyq = bins - 1;
  for (int f = warp; f < nf; f += 8) {
    float2* row = spec + ((int64_t)b * F + f0 + f) * bins;
    for (int k = lane; k < bins; k += 32) {
      const float2 c = row[k];
      const float r = mag_s[k * (FT + 1) + f] * ((k == 0 || k == nyq) ? 1.0f : 0.5f);
      row[k] = make_float2(c.x * r, c.y * r);
    }
  }
}

// grad_wav[b][i] = sum over the padded positions p that reflect onto i, over the frames f covering p:
//   window_full[p - f hop] * gframes[b F + f][p - f hop]      (adjoint of frame_window_kernel)
__global__ void __launch_bounds__(256) mel_bwd_ola_kernel(const float* __restrict__ gframes, const float* __restrict__ window,
                                                          float* __restrict__ gwav, int T, int F, int n_fft, int hop,
                                                          int win, int pad) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= T) return;
  const int lpad = (n_fft - win) / 2;
  const float* gb = gframes + (int64_t)b * F * n_fft;
  const int plen = T + 2 * pad;
  auto at = [&](int p) -> float {
    if (p < 0 || p >= plen) return 0.f;
    float acc = 0.f;
    int f_hi = p / hop;
    if (f_hi > F - 1) f_hi = F - 1;
    for (int f = f_hi; f >= 0; --f) {
      const int n = p - f * hop;
      if (n >= n_fft) break;
      const int wi = n - lpad;
      if (wi >= 0 && wi < win) acc = fmaf(__ldg(window + wi), __ldg(gb + (int64_t)f * n_fft + n), acc);
    }
    return acc;
  };
  float g = at(i + pad);
  // left reflection: padded p in [0, pad) reads y[pad - p]; right: p in [pad + T, plen) reads y[2 (T - 1) - (p - pad)]
  if (i >= 1 && i <= pad) g += at(pad - i);
  if (i <= T - 2 && i >= T - 1 - pad) g += at(pad + 2 * (T - 1) - i);
  gwav[(int64_t)b * T + i] = g;
}

// ---------------------------------------------------------------------------------------------------------
// iSTFT head (APNet.forward, models/vocoders/gan/generator/apnet.py:378-399, and ISTFT "same", :46-104):
//   pha = atan2(I, R); rea = exp(logamp) cos(pha); imag = exp(logamp) sin(pha)
//   audio = overlap-add(irfft(rea + i imag) * window) / overlap-add(window^2), trimmed by (win - hop) / 2 per side.
// Inputs are [B][bins][F] (frames fastest, as the convolutions write them); the spectrum for cuFFT is
// [B*F][bins] complex, so the kernel transposes 32 x 32 tiles through shared memory.
__global__ void __launch_bounds__(256) spectral_head_kernel(const float* __restrict__ logamp, const float* __restrict__ R,
                                                            const float* __restrict__ I, float* __restrict__ pha_out,
                                                            float* __restrict__ rea_out, float* __restrict__ imag_out,
                                                            float2* __restrict__ spec, int bins, int F, float amp_max) {
  __shared__ float2 tile[32][33];
  const int b = blockIdx.z, k0 = blockIdx.y * 32, f0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int k = k0 + r, f = f0 + tx;
    float2 c = make_float2(0.f, 0.f);
    if (k < bins && f < F) {
      const int64_t idx = ((int64_t)b * bins + k) * F + f;
      // I == null: R already holds the phase (Vocos' ISTFTHead, vocos.py:333-361: S = clip(exp(mag), max) e^{ip})
      const float ph = I != nullptr ? atan2f(__ldg(I + idx), __ldg(R + idx)) : __ldg(R + idx);
      float amp = expf(__ldg(logamp + idx));
      if (amp_max > 0.f) amp = fminf(amp, amp_max);
      float sn, cs;
      sincosf(ph, &sn, &cs);
      c = make_float2(amp * cs, amp * sn);
      if (pha_out) pha_out[idx] = ph;
      if (rea_out) rea_out[idx] = c.x;
      if (imag_out) imag_out[idx] = c.y;
    }
    tile[r][tx] = c;
  }
  __syncthreads();
  if (spec != nullptr)
    for (int r = ty; r < 32; r += 8) {
      const int f = f0 + r, k = k0 + tx;
      if (f < F && k < bins) spec[((int64_t)b * F + f) * bins + k] = tile[tx][r];
    }
}

// amplitude_phase_spectrum (utils/mel.py:244-280, APNet's training features): from the complex spectrum [B*F][bins]
// to log(|X| + 1e-5), atan2(im, re), re, im as [B][bins][F] — 32 x 32 tiles transposed through shared memory.
__global__ void __launch_bounds__(256) amp_phase_kernel(const float2* __restrict__ spec, float* __restrict__ logamp,
                                                        float* __restrict__ pha, float* __restrict__ rea,
                                                        float* __restrict__ imag, int bins, int F) {
  __shared__ float2 tile[32][33];
  const int b = blockIdx.z, k0 = blockIdx.y * 32, f0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int f = f0 + r, k = k0 + tx;
    tile[r][tx] = (f < F && k < bins) ? __ldg(spec + ((int64_t)b * F + f) * bins + k) : make_float2(0.f, 0.f);
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int k = k0 + r, f = f0 + tx;
    if (k < bins && f < F) {
      const float2 c = tile[tx][r];
      const int64_t idx = ((int64_t)b * bins + k) * F + f;
      // torch: log(abs(sqrt(re^2 + im^2)) + 1e-5)
      const float mag = __fsqrt_rn(__fadd_rn(__fmul_rn(c.x, c.x), __fmul_rn(c.y, c.y)));
      if (logamp) logamp[idx] = logf(mag + 1e-5f);
      if (pha) pha[idx] = atan2f(c.y, c.x);
      if (rea) rea[idx] = c.x;
      if (imag) imag[idx] = c.y;
    }
  }
}

// wav[b][i] = sum_f window[p - f hop] frames[b F + f][p - f hop] / n_fft  /  sum_f window[p - f hop]^2,  p = i + pad
__global__ void __launch_bounds__(256) istft_ola_kernel(const float* __restrict__ frames, const float* __restrict__ window,
                                                        float* __restrict__ wav, int F, int n_fft, int hop, int pad,
                                                        int L, float scale) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= L) return;
  const int p = i + pad;
  const float* fb = frames + (int64_t)b * F * n_fft;
  float acc = 0.f, env = 0.f;
  int f_hi = p / hop;
  if (f_hi > F - 1) f_hi = F - 1;
  for (int f = f_hi; f >= 0; --f) {
    const int n = p - f * hop;
    if (n >= n_fft) break;
    const float w = __ldg(window + n);
    acc = fmaf(w * scale, __ldg(fb + (int64_t)f * n_fft + n), acc);
    env = fmaf(w, w, env);
  }
  wav[(int64_t)b * L + i] = acc / env;
}

const char* cufft_err(cufftResult r) {
  switch (r) {
    case CUFFT_SUCCESS: return "CUFFT_SUCCESS";
    case CUFFT_INVALID_PLAN: return "CUFFT_INVALID_PLAN";
    case CUFFT_ALLOC_FAILED: return "CUFFT_ALLOC_FAILED";
    case CUFFT_INVALID_VALUE: return "CUFFT_INVALID_VALUE";
    case CUFFT_INTERNAL_ERROR: return "CUFFT_INTERNAL_ERROR";
    case CUFFT_EXEC_FAILED: return "CUFFT_EXEC_FAILED";
    case CUFFT_SETUP_FAILED: return "CUFFT_SETUP_FAILED";
    case CUFFT_INVALID_SIZE: return "CUFFT_INVALID_SIZE";
    default: return "CUFFT_<other>";
  }
}

#define AB_CUFFT_TRY(expr)                                                                   \
  do {                                                                                       \
    cufftResult _r = (expr);                                                                 \
    if (_r != CUFFT_SUCCESS) return fail(AB_ERR_CUDA, "%s failed: %s", #expr, cufft_err(_r)); \
  } while (0)

int get_plan(ab_mel* m, int64_t batch, cufftHandle* plan, size_t* ws, bool inverse = false) {
  constexpr size_t kMaxPlans = 8;
  const int64_t nbatch = batch;
  if (inverse) batch = -batch;             // cache key: C2R plans live under the negated batch
  auto touch = [&](int64_t b) {
    auto& l = m->plan_lru;
    l.erase(std::remove(l.begin(), l.end(), b), l.end());
    l.push_back(b);
  };
  auto it = m->plans.find(batch);
  if (it != m->plans.end()) {
    *plan = it->second;
    *ws = m->plan_ws[batch];
    touch(batch);
    return AB_OK;
  }
  while (m->plans.size() >= kMaxPlans && !m->plan_lru.empty()) {   // evict the least recently used plan
    const int64_t old = m->plan_lru.front();
    m->plan_lru.erase(m->plan_lru.begin());
    auto ev = m->plans.find(old);
    if (ev != m->plans.end()) { cufftDestroy(ev->second); m->plans.erase(ev); m->plan_ws.erase(old); }
  }
  if (nbatch > 0x7fffffffll) return fail(AB_ERR_UNSUPPORTED, "mel: too many frames");
  cufftHandle h;
  AB_CUFFT_TRY(cufftCreate(&h));
  AB_CUFFT_TRY(cufftSetAutoAllocation(h, 0));
  size_t sz = 0;
  int n[1] = {m->cfg.n_fft};
  AB_CUFFT_TRY(cufftMakePlanMany(h, 1, n, nullptr, 1, 0, nullptr, 1, 0, inverse ? CUFFT_C2R : CUFFT_R2C, (int)nbatch, &sz));
  m->plans[batch] = h;
  m->plan_ws[batch] = sz;
  touch(batch);
  *plan = h;
  *ws = sz;
  return AB_OK;
}


// ---------------------------------------------------------------------------------------------------------
// Fused front end for n_fft = 1024 (ab_mel_forward_fused): reflect-pad + frame + window -> 512-point complex FFT of
// the even/odd-packed frame (three radix-8 Stockham passes through shared memory, one warp per frame) -> real-FFT
// post-processing -> |.| -> mel filterbank (non-zero spans only) -> log-clamp (+ energy).  Nothing but the wav is
// read from and nothing but mel / energy is written to HBM: 74 MB for 64 x 10 s instead of ~0.9 GB.
// Persistent CTAs of 8 warps; a CTA takes 8 consecutive frames of one utterance so the output rows are written as
// full 32-byte sectors.  fp32 throughout; twiddles from sincospif (exact argument reduction).
// ---------------------------------------------------------------------------------------------------------
constexpr int FZ_N = 1024, FZ_H = 512, FZ_WARPS = 8, FZ_MAXVAL = 3072, FZ_MAXMEL = 128;

struct FusedSmem {
  float2 w512[FZ_H];            // exp(-2 pi i m / 512)
  float2 w1024[FZ_H + 1];       // exp(-2 pi i k / 1024), k = 0..512
  float window[FZ_N];           // window centred in n_fft
  float2 buf[FZ_WARPS][2][FZ_H];
  float mel_st[FZ_MAXMEL][FZ_WARPS];
  float en_st[FZ_WARPS];
  float vals[FZ_MAXVAL];        // non-zero filter taps, filter after filter
  int2 span[FZ_MAXMEL];         // [lo, hi) per filter
  int voff[FZ_MAXMEL];
  int fits;                     // all non-zero taps fit `vals` (else the taps are read from global memory)
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }   // a * (-i)
// XOR swizzle of the per-warp FFT buffers: every access pattern of the three passes (strides 1, 8 and 64 over 8-byte
// elements) then takes the minimum two wavefronts; unswizzled, pass 0's stores are 16-way bank conflicts
__device__ __forceinline__ int fsw(int i) { return i ^ ((i >> 4) & 15); }

// forward 8-point DFT, natural order in and out
__device__ __forceinline__ void dft8(float2 (&v)[8]) {
  const float h = 0.70710678118654752f;
  const float2 a0 = cadd(v[0], v[4]), a1 = csub(v[0], v[4]), a2 = cadd(v[2], v[6]), a3 = mul_mi(csub(v[2], v[6]));
  const float2 a4 = cadd(v[1], v[5]), a5 = csub(v[1], v[5]), a6 = cadd(v[3], v[7]), a7 = mul_mi(csub(v[3], v[7]));
  const float2 b0 = cadd(a0, a2), b2 = csub(a0, a2), b1 = cadd(a1, a3), b3 = csub(a1, a3);
  const float2 b4 = cadd(a4, a6), b6 = mul_mi(csub(a4, a6));
  const float2 s5 = cadd(a5, a7), d7 = csub(a5, a7);
  const float2 b5 = make_float2(h * (s5.x + s5.y), h * (s5.y - s5.x));       // * (1 - i)/sqrt2
  const float2 b7 = make_float2(h * (d7.y - d7.x), -h * (d7.x + d7.y));      // * (-1 - i)/sqrt2
  v[0] = cadd(b0, b4); v[4] = csub(b0, b4);
  v[1] = cadd(b1, b5); v[5] = csub(b1, b5);
  v[2] = cadd(b2, b6); v[6] = csub(b2, b6);
  v[3] = cadd(b3, b7); v[7] = csub(b3, b7);
}

__global__ void __launch_bounds__(FZ_WARPS * 32, 2)
mel_fused_kernel(const float* __restrict__ y, const float* __restrict__ window, const float* __restrict__ basis,
                 const int2* __restrict__ span_g, float* __restrict__ mel_out, float* __restrict__ energy_out, int B, int T,
                 int F, int hop, int win, int pad, int bins, int n_mel, float eps, float clampv) {
  extern __shared__ __align__(16) uint8_t fz_raw[];
  FusedSmem& sm = *reinterpret_cast<FusedSmem*>(fz_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int m = threadIdx.x; m < FZ_H; m += blockDim.x) {
    float sn, cs;
    sincospif(-(float)m / 256.0f, &sn, &cs);      // -2 pi m / 512
    sm.w512[m] = make_float2(cs, sn);
  }
  for (int k = threadIdx.x; k <= FZ_H; k += blockDim.x) {
    float sn, cs;
    sincospif(-(float)k / 512.0f, &sn, &cs);      // -2 pi k / 1024
    sm.w1024[k] = make_float2(cs, sn);
  }
  const int lpad = (FZ_N - win) / 2;
  for (int n = threadIdx.x; n < FZ_N; n += blockDim.x) {
    const int wi = n - lpad;
    sm.window[n] = (wi >= 0 && wi < win) ? __ldg(window + wi) : 0.f;
  }
  if (warp == 0) {
    // spans and the exclusive prefix sum of their lengths (tap offsets), 32 filters per round
    int carry = 0;
    for (int m0 = 0; m0 < n_mel; m0 += 32) {
      const int m = m0 + lane;
      const int2 sp = m < n_mel ? span_g[m] : make_int2(0, 0);
      const int len = sp.y - sp.x;
      int inc = len;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
      }
      if (m < n_mel) {
        sm.span[m] = sp;
        sm.voff[m] = carry + inc - len;
      }
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) sm.fits = carry <= FZ_MAXVAL;
  }
  __syncthreads();
  if (sm.fits) {
    for (int m = warp; m < n_mel; m += FZ_WARPS) {
      const int2 sp = sm.span[m];
      for (int k = sp.x + lane; k < sp.y; k += 32) sm.vals[sm.voff[m] + k - sp.x] = __ldg(basis + (int64_t)m * bins + k);
    }
  }
  __syncthreads();

  const int groups_per_seq = (F + FZ_WARPS - 1) / FZ_WARPS;
  const int64_t ngroups = (int64_t)B * groups_per_seq;
  float2* bufA = sm.buf[warp][0];
  float2* bufB = sm.buf[warp][1];
  for (int64_t gidx = blockIdx.x; gidx < ngroups; gidx += gridDim.x) {
    const int b = (int)(gidx / groups_per_seq), f0 = (int)(gidx - (int64_t)b * groups_per_seq) * FZ_WARPS;
    const int f = f0 + warp;
    const bool live = f < F;
    if (live) {
      const float* yb = y + (int64_t)b * T;
      const int base = f * hop - pad;
      // ---- pass 0 (Ns = 1): z[n] = (xw[2n], xw[2n+1]); v[r] = z[j + 64 r]; out[8 j + r]
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        const int j = lane + 32 * jj;
        float2 v[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const int n2 = 2 * (j + 64 * r);
          int i0 = base + n2, i1 = i0 + 1;
          if (i0 < 0) i0 = -i0;
          if (i0 >= T) i0 = 2 * (T - 1) - i0;
          if (i1 < 0) i1 = -i1;
          if (i1 >= T) i1 = 2 * (T - 1) - i1;
          v[r] = make_float2(__ldg(yb + i0) * sm.window[n2], __ldg(yb + i1) * sm.window[n2 + 1]);
        }
        dft8(v);
#pragma unroll
        for (int r = 0; r < 8; ++r) bufA[fsw(8 * j + r)] = v[r];
      }
      __syncwarp();
      // ---- pass 1 (Ns = 8): twiddle exp(-2 pi i r k / 64), k = j mod 8; out[(j / 8) * 64 + k + 8 r]
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        const int j = lane + 32 * jj, k = j & 7;
        float2 v[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const float2 x = bufA[fsw(j + 64 * r)];
          v[r] = r == 0 ? x : cmul(x, sm.w512[r * k * 8]);
        }
        dft8(v);
        const int o = (j >> 3) * 64 + k;
#pragma unroll
        for (int r = 0; r < 8; ++r) bufB[fsw(o + 8 * r)] = v[r];
      }
      __syncwarp();
      // ---- pass 2 (Ns = 64): twiddle exp(-2 pi i r j / 512); out[j + 64 r] = Z in natural order
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        const int j = lane + 32 * jj;
        float2 v[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const float2 x = bufB[fsw(j + 64 * r)];
          v[r] = r == 0 ? x : cmul(x, sm.w512[r * j]);
        }
        dft8(v);
#pragma unroll
        for (int r = 0; r < 8; ++r) bufA[fsw(j + 64 * r)] = v[r];
      }
      __syncwarp();
      // ---- real-FFT post-processing + magnitude: X[k] = Fe + W1024^k Fo, k = 0..512 -> mag in bufB (as floats)
      float* mag = reinterpret_cast<float*>(bufB);
      float e = 0.f;
      for (int k = lane; k <= FZ_H; k += 32) {
        const float2 a = bufA[fsw(k & (FZ_H - 1))];
        const float2 zb = bufA[fsw((FZ_H - k) & (FZ_H - 1))];
        const float2 bc = make_float2(zb.x, -zb.y);
        const float2 fe = make_float2(0.5f * (a.x + bc.x), 0.5f * (a.y + bc.y));
        const float2 d = csub(a, bc);
        const float2 fo = make_float2(0.5f * d.y, -0.5f * d.x);     // -i (A - conj B) / 2
        const float2 x = cadd(fe, cmul(sm.w1024[k], fo));
        const float mg = sqrtf(x.x * x.x + x.y * x.y + eps);
        mag[k] = mg;
        e = fmaf(mg, mg, e);
      }
      if (energy_out != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
        if (lane == 0) sm.en_st[warp] = sqrtf(e);
      }
      __syncwarp();
      for (int m = lane; m < n_mel; m += 32) {
        const int2 sp = sm.span[m];
        float acc = 0.f;
        if (sm.fits) {
          const float* vv = sm.vals + sm.voff[m] - sp.x;
          for (int k = sp.x; k < sp.y; ++k) acc = fmaf(vv[k], mag[k], acc);
        } else {
          const float* vv = basis + (int64_t)m * bins;
          for (int k = sp.x; k < sp.y; ++k) acc = fmaf(__ldg(vv + k), mag[k], acc);
        }
        sm.mel_st[m][warp] = logf(fmaxf(acc, clampv));
      }
    }
    __syncthreads();
    const int nf = min(FZ_WARPS, F - f0);
    for (int u = threadIdx.x; u < n_mel * FZ_WARPS; u += blockDim.x) {
      const int m = u / FZ_WARPS, w = u - m * FZ_WARPS;
      if (w < nf) mel_out[((int64_t)b * n_mel + m) * F + f0 + w] = sm.mel_st[m][w];
    }
    if (energy_out != nullptr && threadIdx.x < nf) energy_out[(int64_t)b * F + f0 + threadIdx.x] = sm.en_st[threadIdx.x];
    __syncthreads();
  }
}

struct MelLayout {
  int64_t F;
  size_t off_frames, off_spec, off_span, off_fft, total;
};

int mel_layout(ab_mel* m, int64_t B, int64_t T, MelLayout* L, cufftHandle* plan) {
  L->F = ab_mel_num_frames(m, T);
  if (L->F <= 0) return fail(AB_ERR_ARG, "mel: %lld samples are too few for n_fft=%d", (long long)T, m->cfg.n_fft);
  size_t fftws = 0;
  int rc = get_plan(m, B * L->F, plan, &fftws);
  if (rc != AB_OK) return rc;
  size_t off = 0;
  L->off_frames = off; off += align_up((size_t)B * L->F * m->cfg.n_fft * sizeof(float), 256);
  L->off_spec = off;   off += align_up((size_t)B * L->F * m->bins * sizeof(float2), 256);
  L->off_span = off;   off += align_up((size_t)std::max(m->cfg.n_mel, 1) * sizeof(int2), 256);
  L->off_fft = off;    off += align_up(fftws, 256);
  L->total = off;
  return AB_OK;
}

}  // namespace

extern "C" {

int ab_mel_create(const ab_mel_config* cfg, ab_mel** out) {
  if (!cfg || !out) return fail(AB_ERR_ARG, "ab_mel_create: null argument");
  if (cfg->n_fft <= 0 || (cfg->n_fft & 1) || cfg->hop <= 0 || cfg->win <= 0 || cfg->win > cfg->n_fft || cfg->n_mel < 0 || cfg->pad < 0)
    return fail(AB_ERR_ARG, "ab_mel_create: bad config (n_fft=%d hop=%d win=%d n_mel=%d pad=%d)", cfg->n_fft, cfg->hop, cfg->win, cfg->n_mel, cfg->pad);
  ab_mel* m = new ab_mel();
  m->cfg = *cfg;
  m->bins = cfg->n_fft / 2 + 1;
  *out = m;
  return AB_OK;
}

void ab_mel_destroy(ab_mel* m) {
  if (!m) return;
  for (auto& kv : m->plans) cufftDestroy(kv.second);
  delete m;
}

int64_t ab_mel_num_frames(const ab_mel* m, int64_t T) {
  if (!m) return 0;
  const int64_t padded = T + 2ll * m->cfg.pad;
  if (padded < m->cfg.n_fft || (m->cfg.pad > 0 && m->cfg.pad >= T)) return 0;
  return 1 + (padded - m->cfg.n_fft) / m->cfg.hop;
}

size_t ab_mel_workspace_bytes(const ab_mel* m, int64_t B, int64_t T) {
  if (!m || B <= 0 || T <= 0) return 0;
  MelLayout L;
  cufftHandle plan;
  if (mel_layout(const_cast<ab_mel*>(m), B, T, &L, &plan) != AB_OK) return 0;
  return L.total;
}

int ab_mel_forward(ab_mel* m, const float* dev_wav, int64_t B, int64_t T, const float* dev_window,
                   const float* dev_mel_basis, float* dev_mag, float* dev_mel, float* dev_energy,
                   void* ws, size_t ws_bytes, void* stream) {
  if (!m || !dev_wav || !dev_window || !ws) return fail(AB_ERR_ARG, "mel_forward: null argument");
  if (dev_mel && (!dev_mel_basis || m->cfg.n_mel <= 0)) return fail(AB_ERR_ARG, "mel_forward: mel output needs a mel basis and n_mel > 0");
  if (B <= 0 || B > 65535 || T <= 0 || T > (1ll << 30)) return fail(AB_ERR_ARG, "mel_forward: bad shape");
  MelLayout L;
  cufftHandle plan;
  int rc = mel_layout(m, B, T, &L, &plan);
  if (rc != AB_OK) return rc;
  if (ws_bytes < L.total) return fail(AB_ERR_WORKSPACE, "mel_forward: workspace %zu B < required %zu B", ws_bytes, L.total);
  if (reinterpret_cast<uintptr_t>(ws) & 255) return fail(AB_ERR_ARG, "mel_forward: workspace must be 256-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* base = static_cast<char*>(ws);
  float* frames = reinterpret_cast<float*>(base + L.off_frames);
  float2* spec = reinterpret_cast<float2*>(base + L.off_spec);
  int2* span = reinterpret_cast<int2*>(base + L.off_span);
  const int64_t nfr = B * L.F;
  frame_window_kernel<<<(unsigned)nfr, 256, 0, st>>>(dev_wav, dev_window, frames, (int)T, (int)L.F, m->cfg.n_fft,
                                                    m->cfg.hop, m->cfg.win, m->cfg.pad);
  AB_LAUNCH_CHECK("frame_window_kernel");
  AB_CUFFT_TRY(cufftSetStream(plan, st));
  AB_CUFFT_TRY(cufftSetWorkArea(plan, base + L.off_fft));
  AB_CUFFT_TRY(cufftExecR2C(plan, frames, reinterpret_cast<cufftComplex*>(spec)));
  if (dev_mel) {
    mel_span_kernel<<<(m->cfg.n_mel + 63) / 64, 64, 0, st>>>(dev_mel_basis, m->cfg.n_mel, m->bins, span);
    AB_LAUNCH_CHECK("mel_span_kernel");
  }
  const size_t smem = (size_t)m->bins * (FT + 1) * sizeof(float);
  if (smem > 200 * 1024) return fail(AB_ERR_UNSUPPORTED, "mel_forward: n_fft=%d too large", m->cfg.n_fft);
  static DeviceOnce configured;
  if (smem > 48 * 1024 && configured.need())
    AB_CUDA_TRY(cudaFuncSetAttribute(mag_mel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  dim3 grid((unsigned)ceil_div(L.F, FT), (unsigned)B);
  mag_mel_kernel<<<grid, 256, smem, st>>>(spec, dev_mel_basis, span, dev_mag, dev_mel, dev_energy, (int)L.F,
                                          m->bins, m->cfg.n_mel, m->cfg.eps, m->cfg.clamp);
  AB_LAUNCH_CHECK("mag_mel_kernel");
  return AB_OK;
}

size_t ab_mel_backward_workspace_bytes(const ab_mel* m, int64_t B, int64_t T) {
  if (!m || B <= 0 || T <= 0) return 0;
  MelLayout L;
  cufftHandle plan;
  ab_mel* mm = const_cast<ab_mel*>(m);
  if (mel_layout(mm, B, T, &L, &plan) != AB_OK) return 0;
  size_t inv_ws = 0;
  if (get_plan(mm, B * L.F, &plan, &inv_ws, true) != AB_OK) return 0;
  return L.total + align_up(inv_ws, 256);
}

int ab_mel_backward(ab_mel* m, const float* dev_wav, int64_t B, int64_t T, const float* dev_window,
                    const float* dev_mel_basis, const float* dev_grad_mel, float* dev_grad_wav,
                    void* ws, size_t ws_bytes, void* stream) {
  if (!m || !dev_wav || !dev_window || !dev_mel_basis || !dev_grad_mel || !dev_grad_wav || !ws)
    return fail(AB_ERR_ARG, "mel_backward: null argument");
  if (m->cfg.n_mel <= 0) return fail(AB_ERR_ARG, "mel_backward: the handle has no mel output");
  if (B <= 0 || B > 65535 || T <= 0 || T > (1ll << 30)) return fail(AB_ERR_ARG, "mel_backward: bad shape");
  MelLayout L;
  cufftHandle plan, iplan;
  int rc = mel_layout(m, B, T, &L, &plan);
  if (rc != AB_OK) return rc;
  size_t inv_ws = 0;
  rc = get_plan(m, B * L.F, &iplan, &inv_ws, true);
  if (rc != AB_OK) return rc;
  const size_t need = L.total + align_up(inv_ws, 256);
  if (ws_bytes < need) return fail(AB_ERR_WORKSPACE, "mel_backward: workspace %zu B < required %zu B", ws_bytes, need);
  if (reinterpret_cast<uintptr_t>(ws) & 255) return fail(AB_ERR_ARG, "mel_backward: workspace must be 256-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* base = static_cast<char*>(ws);
  float* frames = reinterpret_cast<float*>(base + L.off_frames);
  float2* spec = reinterpret_cast<float2*>(base + L.off_spec);
  int2* span = reinterpret_cast<int2*>(base + L.off_span);
  const int64_t nfr = B * L.F;
  // recompute the spectrum (cheaper than keeping B*F*bins complex numbers alive between forward and backward)
  frame_window_kernel<<<(unsigned)nfr, 256, 0, st>>>(dev_wav, dev_window, frames, (int)T, (int)L.F, m->cfg.n_fft,
                                                    m->cfg.hop, m->cfg.win, m->cfg.pad);
  AB_LAUNCH_CHECK("frame_window_kernel");
  AB_CUFFT_TRY(cufftSetStream(plan, st));
  AB_CUFFT_TRY(cufftSetWorkArea(plan, base + L.off_fft));
  AB_CUFFT_TRY(cufftExecR2C(plan, frames, reinterpret_cast<cufftComplex*>(spec)));
  mel_span_kernel<<<(m->cfg.n_mel + 63) / 64, 64, 0, st>>>(dev_mel_basis, m->cfg.n_mel, m->bins, span);
  AB_LAUNCH_CHECK("mel_span_kernel");
  const size_t smem = ((size_t)m->bins + m->cfg.n_mel) * (FT + 1) * sizeof(float);
  if (smem > 200 * 1024) return fail(AB_ERR_UNSUPPORTED, "mel_backward: n_fft=%d / n_mel=%d too large", m->cfg.n_fft, m->cfg.n_mel);
  static DeviceOnce configured;
  if (smem > 48 * 1024 && configured.need())
    AB_CUDA_TRY(cudaFuncSetAttribute(mel_bwd_spec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  dim3 grid((unsigned)ceil_div(L.F, FT), (unsigned)B);
  mel_bwd_spec_kernel<<<grid, 256, smem, st>>>(spec, dev_mel_basis, span, dev_grad_mel, (int)L.F, m->bins, m->cfg.n_mel,
                                               m->cfg.eps, m->cfg.clamp);
  AB_LAUNCH_CHECK("mel_bwd_spec_kernel");
  AB_CUFFT_TRY(cufftSetStream(iplan, st));
  AB_CUFFT_TRY(cufftSetWorkArea(iplan, base + L.total));
  AB_CUFFT_TRY(cufftExecC2R(iplan, reinterpret_cast<cufftComplex*>(spec), frames));
  dim3 ogrid((unsigned)ceil_div(T, (int64_t)256), (unsigned)B);
  mel_bwd_ola_kernel<<<ogrid, 256, 0, st>>>(frames, dev_window, dev_grad_wav, (int)T, (int)L.F, m->cfg.n_fft, m->cfg.hop,
                                            m->cfg.win, m->cfg.pad);
  AB_LAUNCH_CHECK("mel_bwd_ola_kernel");
  return AB_OK;
}

int ab_amplitude_phase_forward(ab_mel* m, const float* dev_wav, int64_t B, int64_t T, const float* dev_window,
                               float* dev_logamp, float* dev_pha, float* dev_rea, float* dev_imag,
                               void* ws, size_t ws_bytes, void* stream) {
  if (!m || !dev_wav || !dev_window || !ws) return fail(AB_ERR_ARG, "amplitude_phase: null argument");
  if (B <= 0 || B > 65535 || T <= 0 || T > (1ll << 30)) return fail(AB_ERR_ARG, "amplitude_phase: bad shape");
  MelLayout L;
  cufftHandle plan;
  int rc = mel_layout(m, B, T, &L, &plan);
  if (rc != AB_OK) return rc;
  if (ws_bytes < L.total) return fail(AB_ERR_WORKSPACE, "amplitude_phase: workspace %zu B < required %zu B", ws_bytes, L.total);
  if (reinterpret_cast<uintptr_t>(ws) & 255) return fail(AB_ERR_ARG, "amplitude_phase: workspace must be 256-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* base = static_cast<char*>(ws);
  float* frames = reinterpret_cast<float*>(base + L.off_frames);
  float2* spec = reinterpret_cast<float2*>(base + L.off_spec);
  frame_window_kernel<<<(unsigned)(B * L.F), 256, 0, st>>>(dev_wav, dev_window, frames, (int)T, (int)L.F, m->cfg.n_fft,
                                                          m->cfg.hop, m->cfg.win, m->cfg.pad);
  AB_LAUNCH_CHECK("frame_window_kernel");
  AB_CUFFT_TRY(cufftSetStream(plan, st));
  AB_CUFFT_TRY(cufftSetWorkArea(plan, base + L.off_fft));
  AB_CUFFT_TRY(cufftExecR2C(plan, frames, reinterpret_cast<cufftComplex*>(spec)));
  dim3 grid((unsigned)ceil_div(L.F, (int64_t)32), (unsigned)ceil_div((int64_t)m->bins, (int64_t)32), (unsigned)B);
  amp_phase_kernel<<<grid, 256, 0, st>>>(spec, dev_logamp, dev_pha, dev_rea, dev_imag, m->bins, (int)L.F);
  AB_LAUNCH_CHECK("amp_phase_kernel");
  return AB_OK;
}

int ab_spectral_head_forward(const float* dev_logamp, const float* dev_r, const float* dev_i, int64_t B, int64_t bins,
                             int64_t F, float amp_max, float* dev_pha, float* dev_rea, float* dev_imag, float* dev_spec,
                             void* stream) {
  if (!dev_logamp || !dev_r) return fail(AB_ERR_ARG, "spectral_head: null argument");
  if (B <= 0 || B > 65535 || bins <= 0 || F <= 0 || bins > (1 << 20) || F > (1ll << 30)) return fail(AB_ERR_ARG, "spectral_head: bad shape");
  dim3 grid((unsigned)ceil_div(F, (int64_t)32), (unsigned)ceil_div(bins, (int64_t)32), (unsigned)B);
  spectral_head_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(dev_logamp, dev_r, dev_i, dev_pha, dev_rea, dev_imag,
                                                                           reinterpret_cast<float2*>(dev_spec), (int)bins, (int)F, amp_max);
  AB_LAUNCH_CHECK("spectral_head_kernel");
  return AB_OK;
}

static int istft_layout(ab_mel* m, int64_t B, int64_t F, cufftHandle* plan, size_t* off_fft, size_t* total) {
  if (m->cfg.win != m->cfg.n_fft) return fail(AB_ERR_UNSUPPORTED, "istft: win_size %d != n_fft %d (apnet.py:83 multiplies the n_fft-long frames by the window)", m->cfg.win, m->cfg.n_fft);
  if (m->cfg.hop > m->cfg.win || ((m->cfg.win - m->cfg.hop) & 1)) return fail(AB_ERR_UNSUPPORTED, "istft: need hop <= win and win - hop even");
  size_t fftws = 0;
  int rc = get_plan(m, B * F, plan, &fftws, true);
  if (rc != AB_OK) return rc;
  *off_fft = align_up((size_t)B * F * m->cfg.n_fft * sizeof(float), 256);
  *total = *off_fft + align_up(fftws, 256);
  return AB_OK;
}

size_t ab_istft_workspace_bytes(const ab_mel* m, int64_t B, int64_t F) {
  if (!m || B <= 0 || F <= 0) return 0;
  cufftHandle plan;
  size_t off = 0, total = 0;
  if (istft_layout(const_cast<ab_mel*>(m), B, F, &plan, &off, &total) != AB_OK) return 0;
  return total;
}

int ab_istft_forward(ab_mel* m, float* dev_spec, int64_t B, int64_t F, const float* dev_window, float* dev_wav,
                     void* ws, size_t ws_bytes, void* stream) {
  if (!m || !dev_spec || !dev_window || !dev_wav || !ws) return fail(AB_ERR_ARG, "istft: null argument");
  if (B <= 0 || B > 65535 || F <= 0 || F * (int64_t)m->cfg.hop > (1ll << 30)) return fail(AB_ERR_ARG, "istft: bad shape");
  cufftHandle plan;
  size_t off_fft = 0, total = 0;
  int rc = istft_layout(m, B, F, &plan, &off_fft, &total);
  if (rc != AB_OK) return rc;
  if (ws_bytes < total) return fail(AB_ERR_WORKSPACE, "istft: workspace %zu B < required %zu B", ws_bytes, total);
  if (reinterpret_cast<uintptr_t>(ws) & 255) return fail(AB_ERR_ARG, "istft: workspace must be 256-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* base = static_cast<char*>(ws);
  float* frames = reinterpret_cast<float*>(base);
  AB_CUFFT_TRY(cufftSetStream(plan, st));
  AB_CUFFT_TRY(cufftSetWorkArea(plan, base + off_fft));
  AB_CUFFT_TRY(cufftExecC2R(plan, reinterpret_cast<cufftComplex*>(dev_spec), frames));
  const int pad = (m->cfg.win - m->cfg.hop) / 2;
  const int64_t L = (F - 1) * m->cfg.hop + m->cfg.win - 2 * pad;
  dim3 grid((unsigned)ceil_div(L, (int64_t)256), (unsigned)B);
  istft_ola_kernel<<<grid, 256, 0, st>>>(frames, dev_window, dev_wav, (int)F, m->cfg.n_fft, m->cfg.hop, pad, (int)L,
                                         1.0f / (float)m->cfg.n_fft);
  AB_LAUNCH_CHECK("istft_ola_kernel");
  return AB_OK;
}

int ab_mel_forward_fused(ab_mel* m, const float* dev_wav, int64_t B, int64_t T, const float* dev_window,
                         const float* dev_mel_basis, float* dev_mel, float* dev_energy, void* ws, size_t ws_bytes,
                         void* stream) {
  if (!m || !dev_wav || !dev_window || !dev_mel || !dev_mel_basis || !ws) return fail(AB_ERR_ARG, "mel_forward_fused: null argument");
  if (m->cfg.n_fft != FZ_N) return fail(AB_ERR_UNSUPPORTED, "mel_forward_fused: built for n_fft = 1024 (got %d)", m->cfg.n_fft);
  if (m->cfg.n_mel <= 0 || m->cfg.n_mel > FZ_MAXMEL) return fail(AB_ERR_UNSUPPORTED, "mel_forward_fused: n_mel %d not in [1,%d]", m->cfg.n_mel, FZ_MAXMEL);
  if (B <= 0 || T <= 0 || T > (1ll << 30)) return fail(AB_ERR_ARG, "mel_forward_fused: bad shape");
  const int64_t F = ab_mel_num_frames(m, T);
  if (F <= 0) return fail(AB_ERR_ARG, "mel: %lld samples are too few for n_fft=%d", (long long)T, m->cfg.n_fft);
  if (ws_bytes < (size_t)FZ_MAXMEL * sizeof(int2)) return fail(AB_ERR_WORKSPACE, "mel_forward_fused: workspace too small");
  if (reinterpret_cast<uintptr_t>(ws) & 255) return fail(AB_ERR_ARG, "mel_forward_fused: workspace must be 256-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int2* span = static_cast<int2*>(ws);
  mel_span_kernel<<<(m->cfg.n_mel + 63) / 64, 64, 0, st>>>(dev_mel_basis, m->cfg.n_mel, m->bins, span);
  AB_LAUNCH_CHECK("mel_span_kernel");
  static DeviceOnce configured;
  if (configured.need())
    AB_CUDA_TRY(cudaFuncSetAttribute(mel_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FusedSmem)));
  int dev = 0, nsm = 148;
  AB_CUDA_TRY(cudaGetDevice(&dev));
  AB_CUDA_TRY(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
  const int64_t ngroups = B * ((F + FZ_WARPS - 1) / FZ_WARPS);
  const int grid = (int)std::min<int64_t>(ngroups, 2ll * nsm);
  // the filterbank's non-zero taps must fit the shared-memory table: Slaney filterbanks hold ~2 taps per bin
  mel_fused_kernel<<<grid, FZ_WARPS * 32, sizeof(FusedSmem), st>>>(dev_wav, dev_window, dev_mel_basis, span, dev_mel, dev_energy,
                                                                  (int)B, (int)T, (int)F, m->cfg.hop, m->cfg.win, m->cfg.pad,
                                                                  m->bins, m->cfg.n_mel, m->cfg.eps, m->cfg.clamp);
  AB_LAUNCH_CHECK("mel_fused_kernel");
  return AB_OK;
}

}  // extern "C"
