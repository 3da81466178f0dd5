// Fully unrolled tcgen05.mma issue sequences shared by the tensor-core conv kernels (sm_100a).
#pragma once
#include "ab_tc_ptx.cuh"

namespace ab {
namespace tcx {

// MMAs of one weight stage (tap j, 32-channel K chunk), fully unrolled for M accumulator tiles.  The issuing
// thread is the bottleneck of the MMA phases (a few extra instructions per MMA cost tens of percent), so the
// per-MMA work is two uniform adds and the instruction itself: compile-time tile count, issue order and
// accumulate flag.  KOUTER: K-half outer / M-tile inner (m consecutive MMAs share the B descriptor).
template <int M, bool KOUTER, int FIRST, class Poll>
__device__ __forceinline__ void issue_stage(uint32_t elected, uint32_t tmem, uint32_t np, uint64_t hi, uint32_t alo,
                                            uint32_t blo, uint32_t kstepA, uint32_t kstepB, uint32_t idesc, bool two,
                                            Poll poll) {
  // `poll` runs between the MMAs (before the last group): the issuing thread is blocked behind the tensor pipe's short
  // queue anyway, so that is where the barrier test for the NEXT weight stage costs nothing
  if (KOUTER) {
#pragma unroll
    for (int i = 0; i < M; ++i)
      if (elected) tc_mma_f16_c<FIRST ? 0 : 1>(tmem + (uint32_t)i * np, hi | (alo + 256u * i), hi | blo, idesc);
    poll();
    if (two) {
#pragma unroll
      for (int i = 0; i < M; ++i)
        if (elected) tc_mma_f16_c<1>(tmem + (uint32_t)i * np, hi | (alo + kstepA + 256u * i), hi | (blo + kstepB), idesc);
    }
  } else {
#pragma unroll
    for (int i = 0; i < M; ++i) {
      if (i == M - 1) poll();
      if (elected) {
        tc_mma_f16_c<FIRST ? 0 : 1>(tmem + (uint32_t)i * np, hi | (alo + 256u * i), hi | blo, idesc);
        if (two) tc_mma_f16_c<1>(tmem + (uint32_t)i * np, hi | (alo + kstepA + 256u * i), hi | (blo + kstepB), idesc);
      }
    }
  }
}
template <int M, bool KOUTER, int FIRST>
__device__ __forceinline__ void issue_stage(uint32_t elected, uint32_t tmem, uint32_t np, uint64_t hi, uint32_t alo,
                                            uint32_t blo, uint32_t kstepA, uint32_t kstepB, uint32_t idesc, bool two) {
  issue_stage<M, KOUTER, FIRST>(elected, tmem, np, hi, alo, blo, kstepA, kstepB, idesc, two, [] {});
}

}  // namespace tcx
}  // namespace ab
