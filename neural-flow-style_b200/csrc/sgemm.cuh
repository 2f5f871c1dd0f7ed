// Tiled SGEMM engine shared by the fp32 loss-network path (lossnet.cu) and the mixed-precision
// edge layers of the tensor-core path (conv_tc.cu): 64x64x16 tiles, 4x4 register micro-tiles,
// operand loaders / epilogues as functors.
#pragma once
#include "common.cuh"

#define GBM 64
#define GBN 64
#define GBK 16

// ---- operand loaders: element (row, k) of A [M,K] and (k, col) of B [K,N] ---------------
// An A loader is used in two stages: ``row(m)`` once per thread and tile row (the rows a thread stages do not change
// over the K loop), ``col(k)`` once per K step, ``load(row, col)`` per element.  For the implicit-im2col views this
// keeps every runtime integer division (pixel -> image/y/x, k -> tap/channel) out of the per-element path: the direct
// ``A(m, k)`` form spent ~6 divisions per staged element, several times the FMA work of the tile.
__device__ __forceinline__ float lnst_to_f32(float v) { return v; }
__device__ __forceinline__ void lnst_from_f32(float* p, float v) { *p = v; }
#ifndef LNST_CPU_EMU
#include <cuda_bf16.h>
__device__ __forceinline__ float lnst_to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ void lnst_from_f32(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
#endif

template <class TIn>
struct ConvAT {           // im2col view of x [n,H,W,Cin]: row = pixel, k = (ky*3+kx)*Cin + ci
  const TIn* x;
  int H, W, Cin;
  static constexpr bool kContigM = false;
  struct Row { int pix, py, px; };                 // linear pixel index (img*H + py)*W + px
  struct Col { int ci, dy, dx; };                  // channel, tap offset in [-1,1]
  __device__ __forceinline__ Row row(int m) const {
    const int px = m % W, t = m / W;
    return Row{m, t % H, px};
  }
  __device__ __forceinline__ Col col(int k) const {
    const int ci = k % Cin, tap = k / Cin;
    const int ky = tap / 3;
    return Col{ci, ky - 1, tap - 3 * ky - 1};
  }
  __device__ __forceinline__ float load(const Row& r, const Col& c) const {
    const int yy = r.py + c.dy, xx = r.px + c.dx;
    if (yy < 0 || yy >= H || xx < 0 || xx >= W) return 0.f;
    return lnst_to_f32(x[(int64_t)(r.pix + c.dy * W + c.dx) * Cin + c.ci]);
  }
};
typedef ConvAT<float> ConvA;
struct RowMajorA {        // A [M,K] row-major with leading dimension ld
  const float* a;
  int ld;
  static constexpr bool kContigM = false;
  typedef int Row;
  typedef int Col;
  __device__ __forceinline__ Row row(int m) const { return m; }
  __device__ __forceinline__ Col col(int k) const { return k; }
  __device__ __forceinline__ float load(Row m, Col k) const { return a[(int64_t)m * ld + k]; }
};
struct TransposedA {      // A = X^T where X [K,M] row-major (Gram: F^T)
  const float* a;
  int ld;
  static constexpr bool kContigM = true;
  typedef int Row;
  typedef int Col;
  __device__ __forceinline__ Row row(int m) const { return m; }
  __device__ __forceinline__ Col col(int k) const { return k; }
  __device__ __forceinline__ float load(Row m, Col k) const { return a[(int64_t)k * ld + m]; }
};
struct RowMajorB {        // B loaders: kContigK = false -> staged 64 columns wide (coalesced along n), B(k, n) per element;
  const float* b;         // kContigK = true -> 16 k wide (operands contiguous along k), kcol(k) once per K step + load(kcol, n)
  int ld;
  static constexpr bool kContigK = false;
  __device__ __forceinline__ float operator()(int k, int n) const { return b[(int64_t)k * ld + n]; }
};

// ---- epilogues ---------------------------------------------------------------------------
template <class TOut, class TMask>
struct ConvEpilogueT {    // y = [relu](acc + bias) [* (mask > 0)]
  TOut* y;
  const float* bias;
  const TMask* mask;
  int ld, relu;
  __device__ __forceinline__ void operator()(int m, int n, float acc) const {
    float v = acc + (bias ? bias[n] : 0.f);
    if (relu) v = fmaxf(v, 0.f);
    const int64_t o = (int64_t)m * ld + n;
    if (mask && !(lnst_to_f32(mask[o]) > 0.f)) v = 0.f;
    lnst_from_f32(y + o, v);
  }
};
typedef ConvEpilogueT<float, float> ConvEpilogue;
struct AtomicEpilogue {   // split-K accumulation
  float* c;
  int ld;
  __device__ __forceinline__ void operator()(int m, int n, float acc) const {
    atomicAdd(c + (int64_t)m * ld + n, acc);
  }
};
struct GramBwdEpilogue {  // g = (beta*g + coef*acc) * (F > 0)
  float* g;
  const float* F;
  int ld;
  float coef, beta;
  int relu_mask;
  __device__ __forceinline__ void operator()(int m, int n, float acc) const {
    const int64_t o = (int64_t)m * ld + n;
    float v = coef * acc;
    if (beta != 0.f) v += beta * g[o];
    g[o] = (!relu_mask || F[o] > 0.f) ? v : 0.f;
  }
};

template <class AL, class BL, class EP>
__global__ void __launch_bounds__(256) sgemm_k(AL A, BL B, EP ep, int M, int N, int K, int k_per_split) {
  __shared__ float As[GBK][GBM + 4];
  __shared__ float Bs[GBK][GBN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * GBM, n0 = blockIdx.x * GBN;
  const int kbeg = blockIdx.z * k_per_split;
  const int kend = min(K, kbeg + k_per_split);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // A elements this thread stages per 64x16 tile (same shared-memory layout as a flat e = tid + 256 j sweep):
  //   !kContigM: k column tid & 15, rows (tid >> 4) + 16 j;   kContigM: row tid & 63, k columns (tid >> 6) + 4 j
  constexpr int AR = AL::kContigM ? 1 : 4;
  typename AL::Row arow[AR];
  bool arow_ok[AR];
#pragma unroll
  for (int j = 0; j < AR; ++j) {
    const int m = m0 + (AL::kContigM ? (tid & 63) : (tid >> 4) + 16 * j);
    arow_ok[j] = m < M;
    arow[j] = A.row(arow_ok[j] ? m : 0);
  }

  for (int k0 = kbeg; k0 < kend; k0 += GBK) {
    if (AL::kContigM) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int kk = (tid >> 6) + 4 * j, k = k0 + kk;
        As[kk][tid & 63] = (arow_ok[0] && k < kend) ? A.load(arow[0], A.col(k)) : 0.f;
      }
    } else {
      const int kk = tid & 15, k = k0 + kk;
      const bool kok = k < kend;
      const typename AL::Col acol = A.col(kok ? k : kbeg);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        As[kk][(tid >> 4) + 16 * j] = (kok && arow_ok[j < AR ? j : 0]) ? A.load(arow[j < AR ? j : 0], acol) : 0.f;
    }
    if constexpr (BL::kContigK) {
      const int kk = tid & 15, k = k0 + kk;
      const bool kok = k < kend;
      const auto bcol = B.kcol(kok ? k : kbeg);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int nn = (tid >> 4) + 16 * j, n = n0 + nn;
        Bs[kk][nn] = (kok && n < N) ? B.load(bcol, n) : 0.f;
      }
    } else {
#pragma unroll
      for (int e = tid; e < GBN * GBK; e += 256) {
        const int nn = e % GBN, kk = e / GBN;
        const int n = n0 + nn, k = k0 + kk;
        Bs[kk][nn] = (n < N && k < kend) ? B(k, n) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GBK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < N) ep(m, n, acc[i][j]);
    }
  }
}

template <class AL, class BL, class EP>
static int run_sgemm(AL A, BL B, EP ep, int M, int N, int K, int splits, cudaStream_t s) {
  if (M <= 0 || N <= 0 || K <= 0) return LNST_OK;
  if (splits < 1) splits = 1;
  int kps = (K + splits - 1) / splits;
  kps = ((kps + GBK - 1) / GBK) * GBK;
  splits = (K + kps - 1) / kps;
  auto k = sgemm_k<AL, BL, EP>;
  LNST_LAUNCH(k, dim3((N + GBN - 1) / GBN, (M + GBM - 1) / GBM, splits), dim3(256), 0, s, A, B, ep, M, N, K, kps);
  return lnst_status();
}

