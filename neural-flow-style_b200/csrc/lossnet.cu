// Loss network (CUDA-core fp32 path) and the losses on its features.
//   reference: vgg.py:68-113 (slim VGG: 3x3 SAME conv + bias + ReLU, 2x2/2 avg-pool),
//              styler_base.py:96-102 (Gram), :152-185 (style loss), :135-148 (content),
//              :211-213 (TV).
// This file is the exact-arithmetic (fp32 FMA) implementation used for the tight parity
// tests and as the numerical reference of the tcgen05 path (conv_tc.cu).  One tiled SGEMM
// engine (64x64x16 tiles, 4x4 register micro-tiles) serves the implicit-GEMM convolution,
// the Gram matrix (split-K) and the Gram gradient; they differ only in operand loaders and
// epilogues.
#include "common.cuh"

#include "sgemm.cuh"

// ---- pooling -----------------------------------------------------------------------------
__global__ void avgpool2_fwd_k(const float* __restrict__ x, float* __restrict__ y, int n, int H, int W, int C) {
  const int OH = H / 2, OW = W / 2;
  const int64_t total = (int64_t)n * OH * OW * C;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int c = (int)(t % C);
  const int ox = (int)((t / C) % OW);
  const int oy = (int)((t / ((int64_t)C * OW)) % OH);
  const int img = (int)(t / ((int64_t)C * OW * OH));
  const float* b = x + (((int64_t)img * H + 2 * oy) * W + 2 * ox) * C + c;
  y[t] = (b[0] + b[C] + b[(int64_t)W * C] + b[(int64_t)W * C + C]) * 0.25f;
}
__global__ void avgpool2_bwd_k(const float* __restrict__ gy, const float* __restrict__ mask,
                               float* __restrict__ gx, int n, int H, int W, int C) {
  const int OH = H / 2, OW = W / 2;
  const int64_t total = (int64_t)n * H * W * C;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int c = (int)(t % C);
  const int xx = (int)((t / C) % W);
  const int yy = (int)((t / ((int64_t)C * W)) % H);
  const int img = (int)(t / ((int64_t)C * W * H));
  float g = 0.f;
  const int oy = yy >> 1, ox = xx >> 1;
  if (oy < OH && ox < OW) g = 0.25f * gy[(((int64_t)img * OH + oy) * OW + ox) * C + c];
  if (mask && !(mask[t] > 0.f)) g = 0.f;
  gx[t] = g;
}

// ---- losses ------------------------------------------------------------------------------
// G = G/denom - Gs ; loss += weight * sum(G^2)
// den_dev (may be NULL): the denominator lives on the device, denom = den_scale * den_dev[0] (style mask: 2 C * area of
// the mask, which depends on the render -- styler_base.py:165-169 -- and must not be read back inside a CUDA graph)
__global__ void gram_finish_k(float* __restrict__ G, const float* __restrict__ Gs, int n, float inv_denom,
                              float weight, float* __restrict__ loss, const float* __restrict__ den_dev,
                              float den_scale) {
  if (den_dev) inv_denom = 1.f / (den_scale * den_dev[0]);
  float s = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float d = G[i] * inv_denom;
    if (Gs) { d -= Gs[i]; s += d * d; }
    G[i] = d;
  }
  s = lnst_warp_sum(s);
  if (loss && Gs && (threadIdx.x & 31) == 0 && s != 0.f) atomicAdd(loss, weight * s);
}

// styler_base.py:143-148: -mean(f[...,c]) + mean|f[...,:c]| + mean|f[...,c+1:]|, or -mean(f)
__global__ void content_loss_k(const float* __restrict__ F, int64_t P, int C, int channel, float weight,
                               float* __restrict__ loss, float* __restrict__ gF, float beta, int relu_mask) {
  const int64_t total = P * C;
  float s = 0.f;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(t % C);
    const float f = F[t];
    float l, g;
    if (channel == 0) {
      l = -f / (float)total; g = -1.f / (float)total;
    } else if (c == channel) {
      l = -f / (float)P; g = -1.f / (float)P;
    } else {
      const float cnt = (c < channel) ? (float)P * (float)channel : (float)P * (float)(C - channel - 1);
      l = fabsf(f) / cnt;
      g = (f > 0.f ? 1.f : (f < 0.f ? -1.f : 0.f)) / cnt;
    }
    s += l;
    if (gF) gF[t] = (beta != 0.f ? beta * gF[t] : 0.f) + ((!relu_mask || f > 0.f) ? weight * g : 0.f);
  }
  s = lnst_warp_sum(s);
  // the LAST channel leaves `feature[..., c+1:]` empty: tf.reduce_mean of an empty tensor is NaN, so the reference logs
  // a NaN loss there (its gradient is unaffected: nothing flows into an empty slice) -- reproduced
  if (channel > 0 && channel == C - 1) s = __int_as_float(0x7fc00000);
  if (loss && (threadIdx.x & 31) == 0) atomicAdd(loss, weight * s);
}

// styler_base.py:137-141: mean((f - target*amp)^2) over the feature map (content target image)
__global__ void content_mse_k(const float* __restrict__ F, const float* __restrict__ T, int64_t total, float amp,
                              float weight, float* __restrict__ loss, float* __restrict__ gF, float beta,
                              int relu_mask) {
  float s = 0.f;
  const float inv = 1.f / (float)total;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const float f = F[t];
    const float d = f - T[t] * amp;
    s += d * d * inv;
    if (gF) gF[t] = (beta != 0.f ? beta * gF[t] : 0.f) + ((!relu_mask || f > 0.f) ? weight * 2.f * d * inv : 0.f);
  }
  s = lnst_warp_sum(s);
  if (loss && (threadIdx.x & 31) == 0) atomicAdd(loss, weight * s);
}

// tf.image.total_variation on one image [H,W,C]
__global__ void tv_loss_k(const float* __restrict__ d, int H, int W, int C, float weight,
                          float* __restrict__ loss, float* __restrict__ g) {
  const int64_t total = (int64_t)H * W * C;
  float s = 0.f;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)((t / C) % W), y = (int)(t / ((int64_t)C * W));
    const float v = d[t];
    float gr = 0.f;
    if (y + 1 < H) { const float e = d[t + (int64_t)W * C] - v; s += fabsf(e); gr -= (e > 0.f) - (e < 0.f); }
    if (x + 1 < W) { const float e = d[t + C] - v; s += fabsf(e); gr -= (e > 0.f) - (e < 0.f); }
    if (y > 0) { const float e = v - d[t - (int64_t)W * C]; gr += (e > 0.f) - (e < 0.f); }
    if (x > 0) { const float e = v - d[t - C]; gr += (e > 0.f) - (e < 0.f); }
    if (g) g[t] = weight * gr;
  }
  s = lnst_warp_sum(s);
  if (loss && (threadIdx.x & 31) == 0) atomicAdd(loss, weight * s);
}

// ---------------------------------------------------------------------------------------
// C-ABI
// ---------------------------------------------------------------------------------------
extern "C" int lnst_conv3x3_f32(const float* x, const float* w, const float* b, const float* mask, float* y,
                                int32_t n, int32_t H, int32_t W, int32_t Cin, int32_t Cout, int32_t relu,
                                void* stream) {
  if (!x || !w || !y || n < 1 || H < 1 || W < 1 || Cin < 1 || Cout < 1) return LNST_EARG;
  if ((int64_t)n * H * W > 0x7fffffff) return LNST_EARG;
  ConvA A{x, (int)H, (int)W, (int)Cin};
  RowMajorB B{w, (int)Cout};
  ConvEpilogue ep{y, b, mask, (int)Cout, (int)relu};
  return run_sgemm(A, B, ep, n * H * W, Cout, 9 * Cin, 1, lnst_stream(stream));
}

extern "C" int lnst_avgpool2_fwd(const float* x, float* y, int32_t n, int32_t H, int32_t W, int32_t C,
                                 void* stream) {
  if (!x || !y || n < 1 || H < 2 || W < 2 || C < 1) return LNST_EARG;
  const int64_t total = (int64_t)n * (H / 2) * (W / 2) * C;
  LNST_LAUNCH(avgpool2_fwd_k, dim3(lnst_blocks(total, 256)), dim3(256), 0, lnst_stream(stream), x, y, (int)n,
              (int)H, (int)W, (int)C);
  return lnst_status();
}

extern "C" int lnst_avgpool2_bwd(const float* g_y, const float* mask, float* g_x, int32_t n, int32_t H,
                                 int32_t W, int32_t C, void* stream) {
  if (!g_y || !g_x || n < 1 || H < 2 || W < 2 || C < 1) return LNST_EARG;
  const int64_t total = (int64_t)n * H * W * C;
  LNST_LAUNCH(avgpool2_bwd_k, dim3(lnst_blocks(total, 256)), dim3(256), 0, lnst_stream(stream), g_y, mask, g_x,
              (int)n, (int)H, (int)W, (int)C);
  return lnst_status();
}

// out[i] = x[i] * (num / (den_scale * den_dev[0])): a coefficient that depends on a device-resident denominator,
// folded into the small operand it multiplies (the Gram difference, a loss scalar)
__global__ void scale_by_dev_k(const float* __restrict__ x, int64_t n, float num, const float* __restrict__ den_dev,
                               float den_scale, float* __restrict__ out) {
  const float c = num / (den_scale * den_dev[0]);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = x[i] * c;
}
extern "C" int lnst_scale_by_dev(const float* x, int64_t n, float num, const float* den_dev, float den_scale, float* out,
                                 void* stream) {
  if (!x || !out || !den_dev || n < 1) return LNST_EARG;
  const unsigned nb = lnst_blocks(n, 256) > 256 ? 256 : lnst_blocks(n, 256);
  LNST_LAUNCH(scale_by_dev_k, dim3(nb), dim3(256), 0, lnst_stream(stream), x, n, num, den_dev, den_scale, out);
  return lnst_status();
}

extern "C" int lnst_gram_diff(const float* F, int64_t P, int32_t C, float denom, const float* Gs, float weight,
                              float* G, float* loss, void* stream) {
  return lnst_gram_diff_dev(F, P, C, denom, nullptr, 1.f, Gs, weight, G, loss, stream);
}

// lnst_gram_diff with the denominator on the device when den_dev != NULL: denom = den_scale * den_dev[0] (`denom` unused)
extern "C" int lnst_gram_diff_dev(const float* F, int64_t P, int32_t C, float denom, const float* den_dev, float den_scale,
                                  const float* Gs, float weight, float* G, float* loss, void* stream) {
  if (!F || !G || P < 1 || C < 1 || (!den_dev && !(denom > 0.f)) || P > 0x7fffffff) return LNST_EARG;
  cudaStream_t s = lnst_stream(stream);
  cudaMemsetAsync(G, 0, sizeof(float) * (int64_t)C * C, s);
  TransposedA A{F, (int)C};
  RowMajorB B{F, (int)C};
  AtomicEpilogue ep{G, (int)C};
  const int tiles = ((C + GBM - 1) / GBM) * ((C + GBN - 1) / GBN);
  int splits = (2 * 148 + tiles - 1) / tiles;             // ~2 waves of CTAs over the 148 SMs
  const int max_splits = (int)((P + 4 * GBK - 1) / (4 * GBK));
  if (splits > max_splits) splits = max_splits;
  int rc = run_sgemm(A, B, ep, C, C, (int)P, splits, s);
  if (rc) return rc;
  LNST_LAUNCH(gram_finish_k, dim3(lnst_blocks((int64_t)C * C, 256) > 64 ? 64 : lnst_blocks((int64_t)C * C, 256)),
              dim3(256), 0, s, G, Gs, (int)(C * C), den_dev ? 0.f : 1.f / denom, weight, loss, den_dev, den_scale);
  return lnst_status();
}

extern "C" int lnst_gram_bwd(const float* F, const float* G, int64_t P, int32_t C, float coef, float beta,
                             int32_t relu_mask, float* g_F, void* stream) {
  if (!F || !G || !g_F || P < 1 || C < 1 || P > 0x7fffffff) return LNST_EARG;
  RowMajorA A{F, (int)C};
  RowMajorB B{G, (int)C};
  GramBwdEpilogue ep{g_F, F, (int)C, coef, beta, (int)relu_mask};
  return run_sgemm(A, B, ep, (int)P, C, C, 1, lnst_stream(stream));
}

extern "C" int lnst_content_loss(const float* F, int64_t P, int32_t C, int32_t channel, float weight,
                                 float* loss, float* g_F, float beta, int32_t relu_mask, void* stream) {
  if (!F || P < 1 || C < 1 || channel < 0 || channel >= C) return LNST_EARG;
  const unsigned nb = lnst_blocks(P * C, 256) > 296 ? 296 : lnst_blocks(P * C, 256);
  LNST_LAUNCH(content_loss_k, dim3(nb), dim3(256), 0, lnst_stream(stream), F, P, (int)C, (int)channel, weight,
              loss, g_F, beta, (int)relu_mask);
  return lnst_status();
}

extern "C" int lnst_content_mse(const float* F, const float* target, int64_t n_el, float amp, float weight,
                                float* loss, float* g_F, float beta, int32_t relu_mask, void* stream) {
  if (!F || !target || n_el < 1) return LNST_EARG;
  const unsigned nb = lnst_blocks(n_el, 256) > 296 ? 296 : lnst_blocks(n_el, 256);
  LNST_LAUNCH(content_mse_k, dim3(nb), dim3(256), 0, lnst_stream(stream), F, target, n_el, amp, weight, loss, g_F,
              beta, (int)relu_mask);
  return lnst_status();
}

extern "C" int lnst_tv_loss(const float* d_img, int32_t H, int32_t W, int32_t C, float weight, float* loss,
                            float* g_img, void* stream) {
  if (!d_img || H < 1 || W < 1 || C < 1) return LNST_EARG;
  const int64_t total = (int64_t)H * W * C;
  const unsigned nb = lnst_blocks(total, 256) > 296 ? 296 : lnst_blocks(total, 256);
  LNST_LAUNCH(tv_loss_k, dim3(nb), dim3(256), 0, lnst_stream(stream), d_img, (int)H, (int)W, (int)C, weight, loss,
              g_img);
  return lnst_status();
}
