// Waveform post-processing of utils/io.py:49-76 (save_audio) on the device, so that the D2H copy carries int16:
//   turn_up:      ratio = volume_peak / max(w.max(), |w.min()|);  w *= ratio            (:59-62)
//   add_silence:  fs // 20 zero samples before and after                                  (:64-68)
//   PCM_S 16:     what torchaudio.save(..., encoding="PCM_S", bits_per_sample=16) stores  (:76)
// Quantiser: q = clamp(floor(w * 32768 + 0.5), -32768, 32767) — the sox_io path of the reference's pinned
// torchaudio 2.0.2 (float -> int32 by * 2^31, then SOX_SAMPLE_TO_SIGNED_16BIT = (s + 0x8000) >> 16 with clip).
// torchaudio.save cannot run in the build container (no torchcodec / sox), so this last step is restated from
// that published algorithm, not pinned against an output of the reference ("parity unpinned" for the quantiser;
// the float arithmetic in front of it is pinned by tests/golden/save_audio.npz).
#include <algorithm>

#include "ab_common.cuh"

namespace ab {
namespace {

// |w| peak per utterance.  Non-negative floats order like their bit patterns, so atomicMax on the int view works.
__global__ void pcm_peak_kernel(const float* __restrict__ w, int64_t row_stride, const int64_t* __restrict__ lengths,
                                int64_t T, unsigned int* __restrict__ peak_bits) {
  const int b = blockIdx.y;
  const int64_t n = lengths ? min(lengths[b], T) : T;
  const float* row = w + (int64_t)b * row_stride;
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(__ldg(row + i)));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ float sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 8) {
    m = sm[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffu, m, o));
    if (threadIdx.x == 0) atomicMax(peak_bits + b, __float_as_uint(m));
  }
}

__device__ __forceinline__ int quant16(float v) {
  const float q = floorf(fmaf(v, 32768.0f, 0.5f));
  return (int)fminf(fmaxf(q, -32768.0f), 32767.0f);
}

// out[b, :] = [silence zeros | q(w[b, :len_b] * ratio_b) | silence zeros | zeros up to out_stride]
__global__ void pcm_quant_kernel(const float* __restrict__ w, int64_t row_stride, const int64_t* __restrict__ lengths,
                                 int64_t T, const unsigned int* __restrict__ peak_bits, float volume_peak,
                                 int64_t silence, int16_t* __restrict__ out, int64_t out_stride) {
  const int b = blockIdx.y;
  const int64_t n = lengths ? min(lengths[b], T) : T;
  float ratio = 1.0f;
  if (peak_bits != nullptr) {
    const float peak = __uint_as_float(peak_bits[b]);
    // all-zero utterance: the reference divides by zero there (0 * inf = NaN samples, utils/io.py:60-62);
    // a silent waveform stays silent instead
    ratio = peak > 0.f ? volume_peak / peak : 1.0f;   // IEEE division, as numpy's
  }
  const float* row = w + (int64_t)b * row_stride;
  int16_t* orow = out + (int64_t)b * out_stride;
  // pairs of samples -> one 32-bit store (out rows are 4-byte aligned: out_stride is even, checked on the host)
  const int64_t pairs = out_stride >> 1;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < pairs; p += (int64_t)gridDim.x * blockDim.x) {
    int q[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int64_t i = 2 * p + e - silence;
      q[e] = (i >= 0 && i < n) ? quant16(__ldg(row + i) * ratio) : 0;
    }
    reinterpret_cast<uint32_t*>(orow)[p] = ((uint32_t)(uint16_t)(int16_t)q[0]) | ((uint32_t)(uint16_t)(int16_t)q[1] << 16);
  }
}

}  // namespace
}  // namespace ab

extern "C" {

size_t ab_pcm16_workspace_bytes(int64_t batch) { return batch > 0 ? ab::align_up((size_t)batch * sizeof(unsigned int), 256) : 0; }

int ab_pcm16_forward(const float* dev_wav, int64_t batch, int64_t samples, int64_t row_stride,
                     const int64_t* dev_lengths, int32_t turn_up, float volume_peak, int64_t silence,
                     int16_t* dev_out, int64_t out_stride, void* dev_workspace, size_t workspace_bytes, void* stream) {
  using namespace ab;
  if (!dev_wav || !dev_out) return fail(AB_ERR_ARG, "pcm16: null argument");
  if (batch <= 0 || samples <= 0 || batch > 65535) return fail(AB_ERR_ARG, "pcm16: bad shape");
  if (row_stride < samples || silence < 0) return fail(AB_ERR_ARG, "pcm16: bad stride / silence");
  if (out_stride < samples + 2 * silence || (out_stride & 1) || (reinterpret_cast<uintptr_t>(dev_out) & 3))
    return fail(AB_ERR_ARG, "pcm16: out rows must hold samples + 2*silence, have an even stride and be 4-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  unsigned int* peaks = nullptr;
  const int nblk = (int)std::min<int64_t>(ceil_div(samples, 256 * 8), 148 * 4);
  if (turn_up) {
    if (!dev_workspace || workspace_bytes < ab_pcm16_workspace_bytes(batch)) return fail(AB_ERR_WORKSPACE, "pcm16: workspace too small");
    peaks = static_cast<unsigned int*>(dev_workspace);
    AB_CUDA_TRY(cudaMemsetAsync(peaks, 0, (size_t)batch * sizeof(unsigned int), st));
    pcm_peak_kernel<<<dim3((unsigned)nblk, (unsigned)batch), 256, 0, st>>>(dev_wav, row_stride, dev_lengths, samples, peaks);
    AB_LAUNCH_CHECK("pcm_peak_kernel");
  }
  const int qblk = (int)std::min<int64_t>(ceil_div(out_stride / 2, 256 * 4), 148 * 8);
  pcm_quant_kernel<<<dim3((unsigned)std::max(qblk, 1), (unsigned)batch), 256, 0, st>>>(dev_wav, row_stride, dev_lengths, samples, peaks,
                                                                          volume_peak, silence, dev_out, out_stride);
  AB_LAUNCH_CHECK("pcm_quant_kernel");
  return AB_OK;
}

}  // extern "C"
