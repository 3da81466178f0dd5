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

int get_plan(ab_mel* m, int64_t batch, cufftHandle* plan, size_t* ws) {
  constexpr size_t kMaxPlans = 8;
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
  if (batch > 0x7fffffffll) return fail(AB_ERR_UNSUPPORTED, "mel: too many frames");
  cufftHandle h;
  AB_CUFFT_TRY(cufftCreate(&h));
  AB_CUFFT_TRY(cufftSetAutoAllocation(h, 0));
  size_t sz = 0;
  int n[1] = {m->cfg.n_fft};
  AB_CUFFT_TRY(cufftMakePlanMany(h, 1, n, nullptr, 1, 0, nullptr, 1, 0, CUFFT_R2C, (int)batch, &sz));
  m->plans[batch] = h;
  m->plan_ws[batch] = sz;
  touch(batch);
  *plan = h;
  *ws = sz;
  return AB_OK;
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

}  // extern "C"
