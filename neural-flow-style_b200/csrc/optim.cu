// Optimiser and per-iteration glue of the stylisation loop, plus semi-Lagrangian advection.
//   reference: TF-1.15 ApplyAdam as created at styler_3p.py:320-323 / styler_2p.py:251-254;
//              iterate bookkeeping styler_3p.py:336-363; temporal smoothing :382-383
//              (util.py:169-170 -> scipy.ndimage.gaussian_filter); advect transform.py:557-609.
#include "common.cuh"

// m += (g-m)(1-b1); v += (g^2-v)(1-b2); var -= lr_t m/(sqrt(v)+eps)   (eps outside the bias
// correction, lr_t = lr sqrt(1-b2^t)/(1-b1^t) computed by the caller).  A NaN gradient keeps
// m and v NaN for good and the variable NaN until it is re-assigned from the host-side iterate, as in
// TF; every consumer applies the reference's rule for it (clip -> +1, nan_to_num on the iterate).
__global__ void adam_step_k(float* __restrict__ var, const float* __restrict__ grad, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, float lr_t, float b1, float b2, float eps,
                            float gscale) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float g = grad[i] * gscale;
  const float mi = m[i] + (g - m[i]) * (1.f - b1);
  const float vi = v[i] + (g * g - v[i]) * (1.f - b2);
  m[i] = mi;
  v[i] = vi;
  var[i] = var[i] - lr_t * mi / (sqrtf(vi) + eps);
}

// Device-resident step counter of one AdamOptimizer (state = {beta1^t, beta2^t, lr_t}): computes
// lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t) in fp32 like TF's _prepare/_apply_dense and advances
// the powers, so that a whole optimisation step can be replayed from a CUDA graph with no host
// scalar changing between replays.
__global__ void adam_tick_k(float* __restrict__ state, float lr, float b1, float b2) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  const float b1p = state[0], b2p = state[1];
  state[2] = __fdiv_rn(__fmul_rn(lr, __fsqrt_rn(__fadd_rn(1.f, -b2p))), __fadd_rn(1.f, -b1p));
  state[0] = __fmul_rn(b1p, b1);
  state[1] = __fmul_rn(b2p, b2);
}

__global__ void adam_step_dev_k(float* __restrict__ var, const float* __restrict__ grad, float* __restrict__ m,
                                float* __restrict__ v, int64_t n, const float* __restrict__ state, float b1,
                                float b2, float eps, float gscale) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float lr_t = state[2];
  const float g = grad[i] * gscale;
  const float mi = m[i] + (g - m[i]) * (1.f - b1);
  const float vi = v[i] + (g * g - v[i]) * (1.f - b2);
  m[i] = mi;
  v[i] = vi;
  var[i] = var[i] - lr_t * mi / (sqrtf(vi) + eps);
}

// One iteration of a frame with a single Adam step (no views, or the mean-gradient view mode), fused:
// var <- g_opt (styler_3p.py:312), ApplyAdam, delta = (nan_to_num(var) - g_opt) [* mask] (:359-363) and, when
// no temporal filter runs between them, g_opt += delta (:385-386).  One pass over N*c elements instead of four.
__global__ void adam_iterate_dev_k(float* __restrict__ g_opt, const float* __restrict__ grad, float* __restrict__ m,
                                   float* __restrict__ v, int64_t n, const float* __restrict__ state, float b1,
                                   float b2, float eps, float gscale, const float* __restrict__ mask, int width,
                                   int mask_stride, float* __restrict__ var_out, float* __restrict__ delta,
                                   int apply) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float lr_t = state[2];
  const float x = g_opt[i];
  const float g = grad[i] * gscale;
  const float mi = m[i] + (g - m[i]) * (1.f - b1);
  const float vi = v[i] + (g * g - v[i]) * (1.f - b2);
  m[i] = mi;
  v[i] = vi;
  const float var = x - lr_t * mi / (sqrtf(vi) + eps);
  var_out[i] = var;
  float d = lnst_nan_to_num(var) - x;
  if (mask) d *= mask[(i / width) * mask_stride];
  delta[i] = d;
  if (apply) g_opt[i] = x + d;
}

// the same, four elements per thread (16-byte accesses; n % 4 == 0, 16-byte aligned arrays)
__global__ void adam_iterate_dev4_k(float* __restrict__ g_opt, const float* __restrict__ grad, float* __restrict__ m,
                                    float* __restrict__ v, int64_t n, const float* __restrict__ state, float b1,
                                    float b2, float eps, float gscale, const float* __restrict__ mask, int width,
                                    int mask_stride, float* __restrict__ var_out, float* __restrict__ delta,
                                    int apply) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  const float lr_t = state[2];
  const float4 x4 = *reinterpret_cast<const float4*>(g_opt + i), g4 = *reinterpret_cast<const float4*>(grad + i);
  const float4 m4 = *reinterpret_cast<const float4*>(m + i), v4 = *reinterpret_cast<const float4*>(v + i);
  const float x[4] = {x4.x, x4.y, x4.z, x4.w}, gg[4] = {g4.x, g4.y, g4.z, g4.w};
  const float mo[4] = {m4.x, m4.y, m4.z, m4.w}, vo[4] = {v4.x, v4.y, v4.z, v4.w};
  float mi[4], vi[4], var[4], d[4], xn[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float g = gg[e] * gscale;
    mi[e] = mo[e] + (g - mo[e]) * (1.f - b1);
    vi[e] = vo[e] + (g * g - vo[e]) * (1.f - b2);
    var[e] = x[e] - lr_t * mi[e] / (sqrtf(vi[e]) + eps);
    d[e] = lnst_nan_to_num(var[e]) - x[e];
    if (mask) d[e] *= mask[((i + e) / width) * mask_stride];
    xn[e] = x[e] + d[e];
  }
  *reinterpret_cast<float4*>(m + i) = make_float4(mi[0], mi[1], mi[2], mi[3]);
  *reinterpret_cast<float4*>(v + i) = make_float4(vi[0], vi[1], vi[2], vi[3]);
  *reinterpret_cast<float4*>(var_out + i) = make_float4(var[0], var[1], var[2], var[3]);
  *reinterpret_cast<float4*>(delta + i) = make_float4(d[0], d[1], d[2], d[3]);
  if (apply) *reinterpret_cast<float4*>(g_opt + i) = make_float4(xn[0], xn[1], xn[2], xn[3]);
}

__global__ void iterate_accumulate_k(float* __restrict__ acc, const float* __restrict__ var, int64_t n, int first) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = lnst_nan_to_num(var[i]);
  acc[i] = first ? x : acc[i] + x;
}

__global__ void iterate_delta_k(const float* __restrict__ g_new, float scale, const float* __restrict__ g_opt,
                                const float* __restrict__ mask, int width, int mask_stride, int64_t n,
                                float* __restrict__ delta) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float d = lnst_nan_to_num(g_new[i] * scale) - g_opt[i];
  if (mask) d *= mask[(i / width) * mask_stride];
  delta[i] = d;
}

__global__ void axpy_k(float* __restrict__ y, const float* __restrict__ x, float a, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] += a * x[i];
}

// tf.clip_by_value and its gradient (zero strictly outside [lo,hi]); styler_2p.py:68,88,94
__global__ void clip_fwd_k(const float* __restrict__ x, float lo, float hi, float* __restrict__ y, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = fmaxf(fminf(x[i], hi), lo);   // TF builds max(min(x,hi),lo): NaN reads as hi
}
__global__ void clip_bwd_k(const float* __restrict__ g, const float* __restrict__ x, float lo, float hi, float scale,
                           float* __restrict__ gx, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) gx[i] = (x[i] >= lo && x[i] <= hi) ? g[i] * scale : 0.f;
}
// out[i*C+c] = a[i*C+c] * b[i]   (colour field x density mask, styler_2p.py:71,100)
__global__ void mul_bcast_k(const float* __restrict__ a, const float* __restrict__ b, int C, float* __restrict__ out,
                            int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] * b[i / C];
}

// g[i*C+c] = beta*g + t[i*C+c] * m[i] [* (f > 0)]: the masked-Gram gradient back onto the unmasked feature
__global__ void masked_accumulate_k(const float* __restrict__ t, const float* __restrict__ m,
                                    const float* __restrict__ f, int relu, int C, float beta, float* __restrict__ g,
                                    int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = t[i] * m[i / C];
  if (relu && !(f[i] > 0.f)) v = 0.f;
  g[i] = (beta != 0.f ? beta * g[i] : 0.f) + v;
}

// 1-D Gaussian along axis 0 of x [T,M]; scipy 'reflect' boundary (d c b a | a b c d | d c b a)
#define LNST_MAX_GAUSS_RADIUS 64
struct GaussTaps { int radius; float w[LNST_MAX_GAUSS_RADIUS + 1]; };
__global__ void temporal_gauss_k(const float* __restrict__ x, float* __restrict__ y, int T, int64_t M, GaussTaps g) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  const int t = blockIdx.y;
  float s = 0.f;
  for (int o = -g.radius; o <= g.radius; ++o) {
    int q = t + o;
    // reflect (half-sample symmetric), repeated for radii larger than T
    while (q < 0 || q >= T) { if (q < 0) q = -q - 1; if (q >= T) q = 2 * T - 1 - q; }
    s += g.w[o < 0 ? -o : o] * x[(int64_t)q * M + j];
  }
  y[(int64_t)t * M + j] = s;
}

// ---- advection -----------------------------------------------------------------------------
struct AdvDims { int n[3]; float step[3]; int dim; };
__global__ void advect_k(const float* __restrict__ d, const float* __restrict__ vel, AdvDims a, int C,
                         float* __restrict__ out) {
  const int64_t cells = (int64_t)a.n[0] * a.n[1] * (a.dim == 3 ? a.n[2] : 1);
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= cells) return;
  int idx[3];
  if (a.dim == 3) { idx[2] = (int)(t % a.n[2]); idx[1] = (int)((t / a.n[2]) % a.n[1]); idx[0] = (int)(t / ((int64_t)a.n[2] * a.n[1])); }
  else { idx[1] = (int)(t % a.n[1]); idx[0] = (int)(t / a.n[1]); idx[2] = 0; }
  int lo[3], hi[3];
  float fr[3];
  for (int k = 0; k < a.dim; ++k) {
    const float g = __fadd_rn(-1.f, __fmul_rn(a.step[k], (float)idx[k])) - vel[t * a.dim + k];   // p' = p - v
    const float x = (g + 1.f) * ((float)a.n[k] - 1.f) * 0.5f;
    const int f = (int)floorf(x);
    lo[k] = min(max(f, 0), a.n[k] - 1);
    hi[k] = min(max(f + 1, 0), a.n[k] - 1);
    fr[k] = x - (float)lo[k];
  }
  for (int c = 0; c < C; ++c) {
    float o = 0.f;
    for (int corner = 0; corner < (1 << a.dim); ++corner) {
      int64_t lin = 0;
      float w = 1.f;
      for (int k = 0; k < a.dim; ++k) {
        const int bit = (corner >> (a.dim - 1 - k)) & 1;
        lin = lin * a.n[k] + (bit ? hi[k] : lo[k]);
        w *= bit ? fr[k] : (1.f - fr[k]);
      }
      o += w * d[lin * C + c];
    }
    out[t * C + c] = o;
  }
}

// ---------------------------------------------------------------------------------------
extern "C" int lnst_abi_version(void) { return LNST_ABI_VERSION; }

extern "C" int lnst_adam_step(float* var, const float* grad, float* m, float* v, int64_t n, float lr_t,
                              float beta1, float beta2, float eps, float gscale, void* stream) {
  if (n < 0) return LNST_EARG;
  if (n == 0) return LNST_OK;
  if (!var || !grad || !m || !v) return LNST_EARG;
  LNST_LAUNCH(adam_step_k, dim3(lnst_blocks(n, 256)), dim3(256), 0, lnst_stream(stream), var, grad, m, v, n, lr_t,
              beta1, beta2, eps, gscale);
  return lnst_status();
}

extern "C" int lnst_adam_step_dev(float* var, const float* grad, float* m, float* v, int64_t n, float* state,
                                  float lr, float beta1, float beta2, float eps, float gscale, void* stream) {
  if (!state || n < 0 || (n > 0 && (!var || !grad || !m || !v))) return LNST_EARG;   // an empty variable has no storage
  LNST_LAUNCH(adam_tick_k, dim3(1), dim3(32), 0, lnst_stream(stream), state, lr, beta1, beta2);
  if (n > 0)
    LNST_LAUNCH(adam_step_dev_k, dim3(lnst_blocks(n, 256)), dim3(256), 0, lnst_stream(stream), var, grad, m, v, n,
                (const float*)state, beta1, beta2, eps, gscale);
  return lnst_status();
}

extern "C" int lnst_adam_iterate_dev(float* g_opt, const float* grad, float* m, float* v, int64_t n, float* state,
                                     float lr, float beta1, float beta2, float eps, float gscale, const float* mask,
                                     int32_t width, int32_t mask_stride, float* var_out, float* delta, int32_t apply,
                                     void* stream) {
  if (!state || n < 0 || width < 1 || (n > 0 && (!g_opt || !grad || !m || !v || !var_out || !delta))) return LNST_EARG;
  LNST_LAUNCH(adam_tick_k, dim3(1), dim3(32), 0, lnst_stream(stream), state, lr, beta1, beta2);
  const bool vec4 = n % 4 == 0 && ((reinterpret_cast<uintptr_t>(g_opt) | reinterpret_cast<uintptr_t>(grad) |
                                    reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v) |
                                    reinterpret_cast<uintptr_t>(var_out) | reinterpret_cast<uintptr_t>(delta)) & 15u) == 0;
  if (n > 0 && vec4)
    LNST_LAUNCH(adam_iterate_dev4_k, dim3(lnst_blocks(n / 4, 256)), dim3(256), 0, lnst_stream(stream), g_opt, grad, m, v,
                n, (const float*)state, beta1, beta2, eps, gscale, mask, (int)width, (int)mask_stride, var_out, delta,
                (int)apply);
  else if (n > 0)
    LNST_LAUNCH(adam_iterate_dev_k, dim3(lnst_blocks(n, 256)), dim3(256), 0, lnst_stream(stream), g_opt, grad, m, v,
                n, (const float*)state, beta1, beta2, eps, gscale, mask, (int)width, (int)mask_stride, var_out, delta,
                (int)apply);
  return lnst_status();
}

extern "C" int lnst_iterate_accumulate(float* acc, const float* var, int64_t n, int32_t first, void* stream) {
  if (n < 0) return LNST_EARG;
  if (n == 0) return LNST_OK;
  if (!acc || !var) return LNST_EARG;
  LNST_LAUNCH(iterate_accumulate_k, dim3(lnst_blocks(n, 256)), dim3(256), 0, lnst_stream(stream), acc, var, n,
              (int)first);
  return lnst_status();
}

extern "C" int lnst_iterate_delta(const float* g_new, float scale, const float* g_opt, const float* mask,
                                  int32_t width, int32_t mask_stride, int64_t n, float* delta, void* stream) {
  if (n < 0 || width < 1) return LNST_EARG;
  if (n == 0) return LNST_OK;
  if (!g_new || !g_opt || !delta) return LNST_EARG;
  LNST_LAUNCH(iterate_delta_k, dim3(lnst_blocks(n, 256)), dim3(256), 0, lnst_stream(stream), g_new, scale, g_opt,
              mask, (int)width, (int)mask_stride, n, delta);
  return lnst_status();
}

// out[0] = scale * sum(x[0..n)) for a handful of per-view loss slots (one warp; n <= a few hundred): the per-iteration loss
// scalar of the loop (styler_3p.py:342 `np.mean(loss)`), so that the replayed step holds no framework reduction kernels
__global__ void sum_scale_k(const float* __restrict__ x, int n, float scale, float* __restrict__ out) {
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 32) s += x[i];
  s = lnst_warp_sum(s);
  if (threadIdx.x == 0) out[0] = s * scale;
}
extern "C" int lnst_sum_scale(const float* x, int32_t n, float scale, float* out, void* stream) {
  if (!x || !out || n < 1) return LNST_EARG;
  LNST_LAUNCH(sum_scale_k, dim3(1), dim3(32), 0, lnst_stream(stream), x, (int)n, scale, out);
  return lnst_status();
}
// x[0..n) = 0 on the stream (loss slots the loss kernels accumulate into)
extern "C" int lnst_zero(float* x, int64_t n, void* stream) {
  if (n < 0 || (n > 0 && !x)) return LNST_EARG;
  if (n > 0) cudaMemsetAsync(x, 0, sizeof(float) * n, lnst_stream(stream));
  return lnst_status();
}

extern "C" int lnst_axpy(float* y, const float* x, float a, int64_t n, void* stream) {
  if (n < 0) return LNST_EARG;
  if (n == 0) return LNST_OK;
  if (!y || !x) return LNST_EARG;
  LNST_LAUNCH(axpy_k, dim3(lnst_blocks(n, 256)), dim3(256), 0, lnst_stream(stream), y, x, a, n);
  return lnst_status();
}

extern "C" int lnst_clip_fwd(const float* x, float lo, float hi, float* y, int64_t n, void* stream) {
  if (n < 0) return LNST_EARG;
  if (n == 0) return LNST_OK;
  if (!x || !y) return LNST_EARG;
  LNST_LAUNCH(clip_fwd_k, dim3(lnst_blocks(n, 256)), dim3(256), 0, lnst_stream(stream), x, lo, hi, y, n);
  return lnst_status();
}

extern "C" int lnst_clip_bwd(const float* g, const float* x, float lo, float hi, float scale, float* gx, int64_t n,
                             void* stream) {
  if (n < 0) return LNST_EARG;
  if (n == 0) return LNST_OK;
  if (!g || !x || !gx) return LNST_EARG;
  LNST_LAUNCH(clip_bwd_k, dim3(lnst_blocks(n, 256)), dim3(256), 0, lnst_stream(stream), g, x, lo, hi, scale, gx, n);
  return lnst_status();
}

extern "C" int lnst_mul_bcast(const float* a, const float* b, int32_t C, float* out, int64_t n, void* stream) {
  if (n < 0 || C < 1) return LNST_EARG;
  if (n == 0) return LNST_OK;
  if (!a || !b || !out) return LNST_EARG;
  LNST_LAUNCH(mul_bcast_k, dim3(lnst_blocks(n, 256)), dim3(256), 0, lnst_stream(stream), a, b, (int)C, out, n);
  return lnst_status();
}

extern "C" int lnst_masked_accumulate(const float* t, const float* m, const float* f, int32_t relu, int32_t C,
                                      float beta, float* g, int64_t n, void* stream) {
  if (n < 0 || C < 1) return LNST_EARG;
  if (n == 0) return LNST_OK;
  if ((relu && !f) || !t || !m || !g) return LNST_EARG;
  LNST_LAUNCH(masked_accumulate_k, dim3(lnst_blocks(n, 256)), dim3(256), 0, lnst_stream(stream), t, m, f, (int)relu,
              (int)C, beta, g, n);
  return lnst_status();
}

extern "C" int lnst_temporal_gauss(const float* x, float* y, int32_t T, int64_t M, float sigma, void* stream) {
  if (T < 1 || M < 0 || !(sigma > 0.f)) return LNST_EARG;
  if (M == 0) return LNST_OK;                          // frames without particles: nothing to filter
  if (!x || !y) return LNST_EARG;
  GaussTaps g;
  g.radius = (int)(4.0 * (double)sigma + 0.5);           // scipy: int(truncate*sd + 0.5)
  if (g.radius > LNST_MAX_GAUSS_RADIUS) return LNST_EARG;
  double w[LNST_MAX_GAUSS_RADIUS + 1], sum = 0.0;
  for (int i = 0; i <= g.radius; ++i) {
    w[i] = exp(-0.5 * (double)i * (double)i / ((double)sigma * (double)sigma));
    sum += (i == 0 ? 1.0 : 2.0) * w[i];
  }
  for (int i = 0; i <= LNST_MAX_GAUSS_RADIUS; ++i) g.w[i] = i <= g.radius ? (float)(w[i] / sum) : 0.f;
  LNST_LAUNCH(temporal_gauss_k, dim3(lnst_blocks(M, 256), T), dim3(256), 0, lnst_stream(stream), x, y, (int)T, M, g);
  return lnst_status();
}

extern "C" int lnst_advect(const float* d, const float* vel, int32_t dim, const int32_t* dims, int32_t C,
                           float* out, void* stream) {
  if (!d || !vel || !out || !dims || (dim != 2 && dim != 3) || C < 1) return LNST_EARG;
  AdvDims a;
  a.dim = dim;
  int64_t cells = 1;
  for (int k = 0; k < 3; ++k) {
    a.n[k] = k < dim ? dims[k] : 1;
    if (a.n[k] < 1) return LNST_EARG;
    a.step[k] = a.n[k] > 1 ? 2.0f / (float)(a.n[k] - 1) : 0.f;
    cells *= a.n[k];
  }
  LNST_LAUNCH(advect_k, dim3(lnst_blocks(cells, 256)), dim3(256), 0, lnst_stream(stream), d, vel, a, (int)C, out);
  return lnst_status();
}
