// Shared helpers for the LNST sm_100a kernels.
#pragma once
#ifdef LNST_CPU_EMU
#include "cpu_emu.h"   // test tooling: see tools/cpu_emu/cpu_emu.h
#else
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#define LNST_DYN_SMEM(type, name)                                   \
  extern __shared__ __align__(16) unsigned char lnst_dyn_smem_raw[]; \
  type* name = reinterpret_cast<type*>(lnst_dyn_smem_raw)
#define LNST_LAUNCH(kernel, grid, block, smem, stream, ...) \
  kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif

#include "../../include/lnst_b200.h"

#define LNST_OK 0
#define LNST_EARG (-1)

static inline int lnst_status() { return (int)cudaGetLastError(); }
static inline cudaStream_t lnst_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline unsigned lnst_blocks(int64_t n, int threads) {
  return (unsigned)((n + threads - 1) / threads);
}

struct SubVol { int oz, oy, ox, ez, ey, ex; };   // origin and extent of the region a launch covers
static inline SubVol make_subvol(const LnstBox* b, int D, int H, int W) {
  SubVol s = {0, 0, 0, D, H, W};
  if (b) {
    s.oz = b->lo[0]; s.oy = b->lo[1]; s.ox = b->lo[2];
    s.ez = b->hi[0] - b->lo[0] + 1; s.ey = b->hi[1] - b->lo[1] + 1; s.ex = b->hi[2] - b->lo[2] + 1;
  }
  return s;
}
static inline bool box_ok(const LnstBox* b, int D, int H, int W) {
  if (!b) return true;
  return b->lo[0] >= 0 && b->lo[1] >= 0 && b->lo[2] >= 0 && b->hi[0] < D && b->hi[1] < H && b->hi[2] < W &&
         b->lo[0] <= b->hi[0] && b->lo[1] <= b->hi[1] && b->lo[2] <= b->hi[2];
}

// np.nan_to_num for float32: NaN -> 0, +-inf -> +-FLT_MAX
__device__ __forceinline__ float lnst_nan_to_num(float x) {
  if (x != x) return 0.0f;
  if (x > 3.402823466e+38f) return 3.402823466e+38f;
  if (x < -3.402823466e+38f) return -3.402823466e+38f;
  return x;
}

__device__ __forceinline__ float lnst_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float lnst_warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
