// fp32 NHWC operators for GraphDef loss networks (inception5h: reference styler_base.py:19-31,53-57,91-94 --
// `tf.import_graph_def` of tensorflow_inception_graph.pb; op set of that graph: Conv2D k in {1,3,5,7} stride
// {1,2} SAME, BiasAdd, Relu, MaxPool 3x3 stride {1,2} SAME, LRN, Concat).  Forward and DATA gradients only
// (the weights are frozen).  Every backward entry point ACCUMULATES into the input gradient when asked to:
// a tensor that feeds several branches (every inception module input feeds four) sums their cotangents.
//
// Convolutions run on the shared 64x64x16 SGEMM engine (sgemm.cuh) through implicit-im2col loaders; rows may
// be written with a leading dimension > Cout, so a branch can write straight into its slice of a concat buffer.
#include "sgemm.cuh"

struct Conv2dGeom { int H, W, Cin, Cout, kh, kw, stride, pt, pl, OH, OW; };

struct Conv2dA {          // im2col view of x [n,H,W,Cin]: row = output pixel, k = (ky*kw+kx)*Cin + ci
  const float* x;
  Conv2dGeom g;
  static constexpr bool kContigM = false;
  struct Row { int base, y0, x0; };                // img*H*W, top-left input coordinate of the window
  struct Col { int ci, ky, kx; };
  __device__ __forceinline__ Row row(int m) const {
    const int ox = m % g.OW, t = m / g.OW;
    const int oy = t % g.OH, img = t / g.OH;
    return Row{img * g.H * g.W, oy * g.stride - g.pt, ox * g.stride - g.pl};
  }
  __device__ __forceinline__ Col col(int k) const {
    const int ci = k % g.Cin, tap = k / g.Cin;
    const int ky = tap / g.kw;
    return Col{ci, ky, tap - g.kw * ky};
  }
  __device__ __forceinline__ float load(const Row& r, const Col& c) const {
    const int yy = r.y0 + c.ky, xx = r.x0 + c.kx;
    if (yy < 0 || yy >= g.H || xx < 0 || xx >= g.W) return 0.f;
    return x[(int64_t)(r.base + yy * g.W + xx) * g.Cin + c.ci];
  }
};

struct Conv2dGradA {      // rows = INPUT pixels, k = (ky*kw+kx)*Cout + co: the output pixel that tap (ky,kx) maps here
  const float* gy;
  const float* mask;      // post-ReLU output of the convolution (same indexing as gy) or NULL: gy * (mask > 0), the
  int ldg;                // tf.nn.relu gradient folded into the operand load
  Conv2dGeom g;
  static constexpr bool kContigM = false;
  struct Row { int base, ty0, tx0; };              // img*OH*OW, input coordinate + padding
  struct Col { int co, ky, kx; };
  __device__ __forceinline__ Row row(int m) const {
    const int ix = m % g.W, t = m / g.W;
    const int iy = t % g.H, img = t / g.H;
    return Row{img * g.OH * g.OW, iy + g.pt, ix + g.pl};
  }
  __device__ __forceinline__ Col col(int k) const {
    const int co = k % g.Cout, tap = k / g.Cout;
    const int ky = tap / g.kw;
    return Col{co, ky, tap - g.kw * ky};
  }
  __device__ __forceinline__ float load(const Row& r, const Col& c) const {
    const int ty = r.ty0 - c.ky, tx = r.tx0 - c.kx;
    if (ty < 0 || tx < 0) return 0.f;
    int oy = ty, ox = tx;
    if (g.stride != 1) {
      oy = ty / g.stride; ox = tx / g.stride;
      if (oy * g.stride != ty || ox * g.stride != tx) return 0.f;
    }
    if (oy >= g.OH || ox >= g.OW) return 0.f;
    const int64_t o = (int64_t)(r.base + oy * g.OW + ox) * ldg + c.co;
    if (mask && !(mask[o] > 0.f)) return 0.f;
    return gy[o];
  }
};

struct Conv2dGradB {      // (k = tap*Cout + co, n = ci) -> w[tap][ci][co]   (HWIO): contiguous along k
  const float* w;
  int Cin, Cout;
  static constexpr bool kContigK = true;
  __device__ __forceinline__ int64_t kcol(int k) const {
    const int co = k % Cout, tap = k / Cout;
    return (int64_t)tap * Cin * Cout + co;
  }
  __device__ __forceinline__ float load(int64_t kc, int n) const { return w[kc + (int64_t)n * Cout]; }
};

struct AccEpilogue {      // g = acc (+ g)
  float* g;
  int ld, accumulate;
  __device__ __forceinline__ void operator()(int m, int n, float acc) const {
    const int64_t o = (int64_t)m * ld + n;
    g[o] = accumulate ? g[o] + acc : acc;
  }
};

// Data gradient of a convolution with a THIN input (Cin <= 4: the network's first layer, conv2d0 7x7/2 on RGB).  As a GEMM
// it has N = Cin columns, and the 64-wide tiles of the shared SGEMM engine spend 95 % of their FMAs on padding (C5: 11.9 of
// the step's 21 ms).  Here a thread owns one input pixel: it walks the taps that map onto it (stride-aligned ones only),
// reads the Cout cotangents of that output pixel as float4s and keeps its Cin sums in registers; the weights sit in shared
// memory.  Same k order (tap-major, co inner) as the GEMM form.
template <int CIN>
__global__ void __launch_bounds__(128) conv2d_bwd_data_thin_k(const float* __restrict__ gy, const float* __restrict__ mask,
                                                              int ldg, const float* __restrict__ w, float* __restrict__ gx,
                                                              int n, Conv2dGeom g, int accumulate) {
  LNST_DYN_SMEM(float, ws);                                   // [kh*kw][CIN][Cout] (HWIO)
  const int nw = g.kh * g.kw * CIN * g.Cout;
  for (int i = threadIdx.x; i < nw; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= (int64_t)n * g.H * g.W) return;
  const int ix = (int)(m % g.W);
  const int64_t t = m / g.W;
  const int iy = (int)(t % g.H), img = (int)(t / g.H);
  float acc[CIN];
#pragma unroll
  for (int ci = 0; ci < CIN; ++ci) acc[ci] = 0.f;
  const bool vec = (g.Cout % 4 == 0) && (ldg % 4 == 0);
  for (int ky = 0; ky < g.kh; ++ky) {
    const int ty = iy + g.pt - ky;
    if (ty < 0) break;
    const int oy = ty / g.stride;
    if (oy * g.stride != ty || oy >= g.OH) continue;
    for (int kx = 0; kx < g.kw; ++kx) {
      const int tx = ix + g.pl - kx;
      if (tx < 0) break;
      const int ox = tx / g.stride;
      if (ox * g.stride != tx || ox >= g.OW) continue;
      const int64_t o = ((int64_t)(img * g.OH + oy) * g.OW + ox) * ldg;
      const float* wp = ws + (ky * g.kw + kx) * CIN * g.Cout;
      if (vec) {
        for (int co = 0; co < g.Cout; co += 4) {
          float4 gv = *reinterpret_cast<const float4*>(gy + o + co);
          if (mask) {
            const float4 mv = *reinterpret_cast<const float4*>(mask + o + co);
            if (!(mv.x > 0.f)) gv.x = 0.f;
            if (!(mv.y > 0.f)) gv.y = 0.f;
            if (!(mv.z > 0.f)) gv.z = 0.f;
            if (!(mv.w > 0.f)) gv.w = 0.f;
          }
#pragma unroll
          for (int ci = 0; ci < CIN; ++ci) {
            const float4 wv = *reinterpret_cast<const float4*>(wp + ci * g.Cout + co);
            acc[ci] = fmaf(gv.x, wv.x, acc[ci]); acc[ci] = fmaf(gv.y, wv.y, acc[ci]);
            acc[ci] = fmaf(gv.z, wv.z, acc[ci]); acc[ci] = fmaf(gv.w, wv.w, acc[ci]);
          }
        }
      } else {
        for (int co = 0; co < g.Cout; ++co) {
          float gv = gy[o + co];
          if (mask && !(mask[o + co] > 0.f)) gv = 0.f;
#pragma unroll
          for (int ci = 0; ci < CIN; ++ci) acc[ci] = fmaf(gv, wp[ci * g.Cout + co], acc[ci]);
        }
      }
    }
  }
#pragma unroll
  for (int ci = 0; ci < CIN; ++ci) {
    const int64_t q = m * CIN + ci;
    gx[q] = accumulate ? gx[q] + acc[ci] : acc[ci];
  }
}

__global__ void relu_fwd_k(const float* __restrict__ x, float* __restrict__ y, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = fmaxf(x[i], 0.f);
}

// tf.nn.relu gradient: g * (y > 0)
__global__ void relu_bwd_k(const float* __restrict__ gy, const float* __restrict__ y, float* __restrict__ gx, int64_t n,
                           int accumulate) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = y[i] > 0.f ? gy[i] : 0.f;
  gx[i] = accumulate ? gx[i] + v : v;
}

struct PoolGeom { int n, H, W, C, k, stride, pt, pl, OH, OW; };

// tf.nn.max_pool, padding cells never win (SAME pads with -inf)
__global__ void maxpool_fwd_k(const float* __restrict__ x, float* __restrict__ y, PoolGeom p) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)p.n * p.OH * p.OW * p.C;
  if (t >= total) return;
  const int c = (int)(t % p.C);
  int64_t r = t / p.C;
  const int ox = (int)(r % p.OW); r /= p.OW;
  const int oy = (int)(r % p.OH);
  const int img = (int)(r / p.OH);
  float best = -INFINITY;
  for (int ky = 0; ky < p.k; ++ky) {
    const int yy = oy * p.stride + ky - p.pt;
    if (yy < 0 || yy >= p.H) continue;
    for (int kx = 0; kx < p.k; ++kx) {
      const int xx = ox * p.stride + kx - p.pl;
      if (xx < 0 || xx >= p.W) continue;
      best = fmaxf(best, x[(((int64_t)img * p.H + yy) * p.W + xx) * p.C + c]);
    }
  }
  y[t] = best;
}

// MaxPoolGrad: the cotangent of an output goes to the FIRST maximum of its window in (row, column) scan order
// (strict > while scanning, TF's MaxPoolBackwardNoMask); windows overlap, so contributions are atomically added
// into g_x (zeroed or already holding other branches' gradient).
__global__ void maxpool_bwd_k(const float* __restrict__ gy, const float* __restrict__ x, float* __restrict__ gx,
                              PoolGeom p) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)p.n * p.OH * p.OW * p.C;
  if (t >= total) return;
  const int c = (int)(t % p.C);
  int64_t r = t / p.C;
  const int ox = (int)(r % p.OW); r /= p.OW;
  const int oy = (int)(r % p.OH);
  const int img = (int)(r / p.OH);
  float best = -INFINITY;
  int64_t arg = -1;
  for (int ky = 0; ky < p.k; ++ky) {
    const int yy = oy * p.stride + ky - p.pt;
    if (yy < 0 || yy >= p.H) continue;
    for (int kx = 0; kx < p.k; ++kx) {
      const int xx = ox * p.stride + kx - p.pl;
      if (xx < 0 || xx >= p.W) continue;
      const int64_t o = (((int64_t)img * p.H + yy) * p.W + xx) * p.C + c;
      const float v = x[o];
      if (v > best || arg < 0) { best = v; arg = o; }
    }
  }
  const float g = gy[t];
  if (arg >= 0 && g != 0.f) atomicAdd(gx + arg, g);
}

// tf.nn.avg_pool: mean over the window's cells that lie inside the image (SAME padding does not count)
__global__ void avgpool_fwd_k(const float* __restrict__ x, float* __restrict__ y, PoolGeom p) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)p.n * p.OH * p.OW * p.C;
  if (t >= total) return;
  const int c = (int)(t % p.C);
  int64_t r = t / p.C;
  const int ox = (int)(r % p.OW); r /= p.OW;
  const int oy = (int)(r % p.OH);
  const int img = (int)(r / p.OH);
  float acc = 0.f;
  int cnt = 0;
  for (int ky = 0; ky < p.k; ++ky) {
    const int yy = oy * p.stride + ky - p.pt;
    if (yy < 0 || yy >= p.H) continue;
    for (int kx = 0; kx < p.k; ++kx) {
      const int xx = ox * p.stride + kx - p.pl;
      if (xx < 0 || xx >= p.W) continue;
      acc += x[(((int64_t)img * p.H + yy) * p.W + xx) * p.C + c];
      ++cnt;
    }
  }
  y[t] = acc / (float)cnt;
}

// AvgPoolGrad in gather form: an input cell collects g_y / count from every window that contains it
__global__ void avgpool_bwd_k(const float* __restrict__ gy, float* __restrict__ gx, PoolGeom p, int accumulate) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)p.n * p.H * p.W * p.C;
  if (t >= total) return;
  const int c = (int)(t % p.C);
  int64_t r = t / p.C;
  const int ix = (int)(r % p.W); r /= p.W;
  const int iy = (int)(r % p.H);
  const int img = (int)(r / p.H);
  float acc = 0.f;
  for (int ky = 0; ky < p.k; ++ky) {
    const int ty = iy + p.pt - ky;
    if (ty < 0 || ty % p.stride) continue;
    const int oy = ty / p.stride;
    if (oy >= p.OH) continue;
    const int y0 = oy * p.stride - p.pt;
    const int ny = min(y0 + p.k, p.H) - max(y0, 0);
    for (int kx = 0; kx < p.k; ++kx) {
      const int tx = ix + p.pl - kx;
      if (tx < 0 || tx % p.stride) continue;
      const int ox = tx / p.stride;
      if (ox >= p.OW) continue;
      const int x0 = ox * p.stride - p.pl;
      const int nx = min(x0 + p.k, p.W) - max(x0, 0);
      acc += gy[(((int64_t)img * p.OH + oy) * p.OW + ox) * p.C + c] / (float)(ny * nx);
    }
  }
  gx[t] = accumulate ? gx[t] + acc : acc;
}

// tf.nn.local_response_normalization: y_c = x_c * (bias + alpha * sum_{|j-c|<=r} x_j^2)^-beta  (alpha is NOT divided
// by the window size, unlike Caffe / torch)
__device__ __forceinline__ float lrn_norm(const float* __restrict__ row, int C, int c, int radius, float bias,
                                          float alpha) {
  float s = 0.f;
  const int lo = max(c - radius, 0), hi = min(c + radius, C - 1);
  for (int j = lo; j <= hi; ++j) s = fmaf(row[j], row[j], s);
  return bias + alpha * s;
}

__global__ void lrn_fwd_k(const float* __restrict__ x, float* __restrict__ y, int64_t pixels, int C, int radius,
                          float bias, float alpha, float beta) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= pixels * C) return;
  const int c = (int)(t % C);
  const float* row = x + (t - c);
  y[t] = row[c] * powf(lrn_norm(row, C, c, radius, bias, alpha), -beta);
}

// LRNGrad: g_x[k] = sum_{j: |j-k|<=r} g_y[j] * (delta_jk N_j^-beta - 2 alpha beta x_k x_j N_j^(-beta-1))
__global__ void lrn_bwd_k(const float* __restrict__ gy, const float* __restrict__ x, float* __restrict__ gx,
                          int64_t pixels, int C, int radius, float bias, float alpha, float beta, int accumulate) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= pixels * C) return;
  const int k = (int)(t % C);
  const float* row = x + (t - k);
  const float* grow = gy + (t - k);
  const float xk = row[k];
  float acc = 0.f;
  const int lo = max(k - radius, 0), hi = min(k + radius, C - 1);
  for (int j = lo; j <= hi; ++j) {
    const float N = lrn_norm(row, C, j, radius, bias, alpha);
    const float nb = powf(N, -beta);
    float d = -2.f * alpha * beta * xk * row[j] * nb / N;
    if (j == k) d += nb;
    acc = fmaf(grow[j], d, acc);
  }
  gx[t] = accumulate ? gx[t] + acc : acc;
}

// dst[p, 0:C] (+)= src[p, 0:C] with independent row strides: concat (forward) and its slices (backward)
__global__ void copy_channels_k(const float* __restrict__ src, int ld_src, float* __restrict__ dst, int ld_dst, int C,
                                int64_t pixels, int accumulate) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= pixels * C) return;
  const int c = (int)(t % C);
  const int64_t p = t / C;
  const float v = src[p * ld_src + c];
  float* d = dst + p * ld_dst + c;
  *d = accumulate ? *d + v : v;
}

// ---------------------------------------------------------------------------------------------------
static bool conv_geom_ok(const Conv2dGeom& g, int n) {
  if (n < 1 || g.H < 1 || g.W < 1 || g.Cin < 1 || g.Cout < 1 || g.kh < 1 || g.kw < 1 || g.stride < 1 || g.pt < 0 ||
      g.pl < 0 || g.OH < 1 || g.OW < 1)
    return false;
  // every output pixel's window must start inside the padded input
  if ((int64_t)(g.OH - 1) * g.stride - g.pt >= g.H || (int64_t)(g.OW - 1) * g.stride - g.pl >= g.W) return false;
  return (int64_t)n * g.OH * g.OW <= 0x7fffffff && (int64_t)n * g.H * g.W <= 0x7fffffff &&
         (int64_t)g.kh * g.kw * g.Cin <= 0x7fffffff && (int64_t)g.kh * g.kw * g.Cout <= 0x7fffffff;
}

extern "C" int lnst_conv2d_f32(const float* x, const float* w, const float* bias, float* y, int32_t n, int32_t H,
                               int32_t W, int32_t Cin, int32_t Cout, int32_t kh, int32_t kw, int32_t stride,
                               int32_t pad_top, int32_t pad_left, int32_t OH, int32_t OW, int32_t ldy, int32_t relu,
                               void* stream) {
  Conv2dGeom g{H, W, Cin, Cout, kh, kw, stride, pad_top, pad_left, OH, OW};
  if (!x || !w || !y || !conv_geom_ok(g, n) || ldy < Cout) return LNST_EARG;
  Conv2dA A{x, g};
  RowMajorB B{w, (int)Cout};
  ConvEpilogue ep{y, bias, nullptr, (int)ldy, (int)relu};
  return run_sgemm(A, B, ep, n * OH * OW, Cout, kh * kw * Cin, 1, lnst_stream(stream));
}

extern "C" int lnst_conv2d_bwd_data_f32(const float* g_y, const float* relu_y, int32_t ldg, const float* w, float* g_x, int32_t n,
                                        int32_t H, int32_t W, int32_t Cin, int32_t Cout, int32_t kh, int32_t kw,
                                        int32_t stride, int32_t pad_top, int32_t pad_left, int32_t OH, int32_t OW,
                                        int32_t accumulate, void* stream) {
  Conv2dGeom g{H, W, Cin, Cout, kh, kw, stride, pad_top, pad_left, OH, OW};
  if (!g_y || !w || !g_x || !conv_geom_ok(g, n) || ldg < Cout) return LNST_EARG;
  const size_t wbytes = sizeof(float) * (size_t)kh * kw * Cin * Cout;
  if (Cin <= 4 && wbytes <= 48 * 1024 && (int64_t)n * OH * OW * ldg < 0x7fffffff) {   // thin input: one thread per input pixel
    const int64_t pixels = (int64_t)n * H * W;
    const dim3 grid_(lnst_blocks(pixels, 128)), blk(128);
    switch (Cin) {
      case 1: { auto k = conv2d_bwd_data_thin_k<1>; LNST_LAUNCH(k, grid_, blk, wbytes, lnst_stream(stream), g_y, relu_y, (int)ldg, w, g_x, (int)n, g, (int)accumulate); break; }
      case 2: { auto k = conv2d_bwd_data_thin_k<2>; LNST_LAUNCH(k, grid_, blk, wbytes, lnst_stream(stream), g_y, relu_y, (int)ldg, w, g_x, (int)n, g, (int)accumulate); break; }
      case 3: { auto k = conv2d_bwd_data_thin_k<3>; LNST_LAUNCH(k, grid_, blk, wbytes, lnst_stream(stream), g_y, relu_y, (int)ldg, w, g_x, (int)n, g, (int)accumulate); break; }
      default: { auto k = conv2d_bwd_data_thin_k<4>; LNST_LAUNCH(k, grid_, blk, wbytes, lnst_stream(stream), g_y, relu_y, (int)ldg, w, g_x, (int)n, g, (int)accumulate); break; }
    }
    return lnst_status();
  }
  Conv2dGradA A{g_y, relu_y, (int)ldg, g};
  Conv2dGradB B{w, (int)Cin, (int)Cout};
  AccEpilogue ep{g_x, (int)Cin, (int)accumulate};
  return run_sgemm(A, B, ep, n * H * W, Cin, kh * kw * Cout, 1, lnst_stream(stream));
}

extern "C" int lnst_relu_fwd(const float* x, float* y, int64_t n, void* stream) {
  if (n < 0) return LNST_EARG;
  if (n == 0) return LNST_OK;
  if (!x || !y) return LNST_EARG;
  LNST_LAUNCH(relu_fwd_k, dim3(lnst_blocks(n, 256)), dim3(256), 0, lnst_stream(stream), x, y, n);
  return lnst_status();
}

extern "C" int lnst_relu_bwd(const float* g_y, const float* y, float* g_x, int64_t n, int32_t accumulate, void* stream) {
  if (n < 0) return LNST_EARG;
  if (n == 0) return LNST_OK;
  if (!g_y || !y || !g_x) return LNST_EARG;
  LNST_LAUNCH(relu_bwd_k, dim3(lnst_blocks(n, 256)), dim3(256), 0, lnst_stream(stream), g_y, y, g_x, n, (int)accumulate);
  return lnst_status();
}

static bool pool_geom_ok(const PoolGeom& p) {
  return p.n >= 1 && p.H >= 1 && p.W >= 1 && p.C >= 1 && p.k >= 1 && p.stride >= 1 && p.pt >= 0 && p.pl >= 0 &&
         p.pt < p.k && p.pl < p.k && p.OH >= 1 && p.OW >= 1 && (int64_t)(p.OH - 1) * p.stride - p.pt < p.H &&
         (int64_t)(p.OW - 1) * p.stride - p.pl < p.W;
}

extern "C" int lnst_maxpool_fwd(const float* x, float* y, int32_t n, int32_t H, int32_t W, int32_t C, int32_t k,
                                int32_t stride, int32_t pad_top, int32_t pad_left, int32_t OH, int32_t OW,
                                void* stream) {
  PoolGeom p{n, H, W, C, k, stride, pad_top, pad_left, OH, OW};
  if (!x || !y || !pool_geom_ok(p)) return LNST_EARG;
  const int64_t total = (int64_t)n * OH * OW * C;
  LNST_LAUNCH(maxpool_fwd_k, dim3(lnst_blocks(total, 256)), dim3(256), 0, lnst_stream(stream), x, y, p);
  return lnst_status();
}

extern "C" int lnst_maxpool_bwd(const float* g_y, const float* x, float* g_x, int32_t n, int32_t H, int32_t W,
                                int32_t C, int32_t k, int32_t stride, int32_t pad_top, int32_t pad_left, int32_t OH,
                                int32_t OW, int32_t accumulate, void* stream) {
  PoolGeom p{n, H, W, C, k, stride, pad_top, pad_left, OH, OW};
  if (!g_y || !x || !g_x || !pool_geom_ok(p)) return LNST_EARG;
  if (!accumulate) cudaMemsetAsync(g_x, 0, sizeof(float) * (int64_t)n * H * W * C, lnst_stream(stream));
  const int64_t total = (int64_t)n * OH * OW * C;
  LNST_LAUNCH(maxpool_bwd_k, dim3(lnst_blocks(total, 256)), dim3(256), 0, lnst_stream(stream), g_y, x, g_x, p);
  return lnst_status();
}

extern "C" int lnst_avgpool_fwd(const float* x, float* y, int32_t n, int32_t H, int32_t W, int32_t C, int32_t k,
                                int32_t stride, int32_t pad_top, int32_t pad_left, int32_t OH, int32_t OW,
                                void* stream) {
  PoolGeom p{n, H, W, C, k, stride, pad_top, pad_left, OH, OW};
  if (!x || !y || !pool_geom_ok(p)) return LNST_EARG;
  const int64_t total = (int64_t)n * OH * OW * C;
  LNST_LAUNCH(avgpool_fwd_k, dim3(lnst_blocks(total, 256)), dim3(256), 0, lnst_stream(stream), x, y, p);
  return lnst_status();
}

extern "C" int lnst_avgpool_bwd(const float* g_y, float* g_x, int32_t n, int32_t H, int32_t W, int32_t C, int32_t k,
                                int32_t stride, int32_t pad_top, int32_t pad_left, int32_t OH, int32_t OW,
                                int32_t accumulate, void* stream) {
  PoolGeom p{n, H, W, C, k, stride, pad_top, pad_left, OH, OW};
  if (!g_y || !g_x || !pool_geom_ok(p)) return LNST_EARG;
  const int64_t total = (int64_t)n * H * W * C;
  LNST_LAUNCH(avgpool_bwd_k, dim3(lnst_blocks(total, 256)), dim3(256), 0, lnst_stream(stream), g_y, g_x, p,
              (int)accumulate);
  return lnst_status();
}

extern "C" int lnst_lrn_fwd(const float* x, float* y, int64_t pixels, int32_t C, int32_t depth_radius, float bias,
                            float alpha, float beta, void* stream) {
  if (!x || !y || pixels < 1 || C < 1 || depth_radius < 0) return LNST_EARG;
  LNST_LAUNCH(lrn_fwd_k, dim3(lnst_blocks(pixels * C, 256)), dim3(256), 0, lnst_stream(stream), x, y, pixels, (int)C,
              (int)depth_radius, bias, alpha, beta);
  return lnst_status();
}

extern "C" int lnst_lrn_bwd(const float* g_y, const float* x, float* g_x, int64_t pixels, int32_t C,
                            int32_t depth_radius, float bias, float alpha, float beta, int32_t accumulate,
                            void* stream) {
  if (!g_y || !x || !g_x || pixels < 1 || C < 1 || depth_radius < 0) return LNST_EARG;
  LNST_LAUNCH(lrn_bwd_k, dim3(lnst_blocks(pixels * C, 256)), dim3(256), 0, lnst_stream(stream), g_y, x, g_x, pixels,
              (int)C, (int)depth_radius, bias, alpha, beta, (int)accumulate);
  return lnst_status();
}

extern "C" int lnst_copy_channels(const float* src, int32_t ld_src, float* dst, int32_t ld_dst, int32_t C,
                                  int64_t pixels, int32_t accumulate, void* stream) {
  if (!src || !dst || C < 1 || ld_src < C || ld_dst < C || pixels < 1) return LNST_EARG;
  LNST_LAUNCH(copy_channels_k, dim3(lnst_blocks(pixels * C, 256)), dim3(256), 0, lnst_stream(stream), src, (int)ld_src,
              dst, (int)ld_dst, (int)C, pixels, (int)accumulate);
  return lnst_status();
}
