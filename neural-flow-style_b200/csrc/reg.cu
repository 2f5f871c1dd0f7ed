// Field / variable regularisers of the stylisation loss and the library-level queries of the C-ABI.
//   reference: styler_3p.py:96-98 + styler_base.py:228-230 (pressure: mean(where(d > 0, d - 1, 0)^2) on the splatted
//              density), styler_base.py:217-223 (density: (sum d_i)^2 + 1e3 * sum -log(|d_i| + 1e-6) on the clipped
//              variable `self.d = clip(r_opt, -1, 1)`, styler_3p.py:74-76).
// Both add to the loss slots of the views they belong to and ACCUMULATE their gradient into the buffer the rest of the
// backward pass has filled, so the product step needs no eager tensor arithmetic for them.
#include "common.cuh"

// loss[0..n_loss) += w * mean(pr^2);  g_d += g_scale * pr   with pr = d > 0 ? d - rho0 : 0
__global__ void __launch_bounds__(256) pressure_reg_k(const float* __restrict__ d, int64_t cells, float rho0, float w_mean,
                                                      float g_scale, float* __restrict__ loss, int n_loss,
                                                      float* __restrict__ g_d) {
  __shared__ float part[8];
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += (int64_t)gridDim.x * blockDim.x) {
    const float dv = d[i];
    const float pr = dv > 0.f ? dv - rho0 : 0.f;
    acc += pr * pr;
    if (g_d && pr != 0.f) g_d[i] += g_scale * pr;
  }
  acc = lnst_warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float s = threadIdx.x < 8 ? part[threadIdx.x] : 0.f;
    s = lnst_warp_sum(s);
    if (threadIdx.x == 0 && loss)
      for (int k = 0; k < n_loss; ++k) atomicAdd(loss + k, s * w_mean);
  }
}

// pass 1: sums[0] += sum clip(var), sums[1] += sum -log(|clip(var)| + 1e-6)
__global__ void __launch_bounds__(256) density_reg_sums_k(const float* __restrict__ var, int64_t n, float* __restrict__ sums) {
  __shared__ float part[2][8];
  float a = 0.f, b = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float dv = fmaxf(fminf(var[i], 1.f), -1.f);
    a += dv;
    b -= logf(fabsf(dv) + 1e-6f);
  }
  a = lnst_warp_sum(a); b = lnst_warp_sum(b);
  if ((threadIdx.x & 31) == 0) { part[0][threadIdx.x >> 5] = a; part[1][threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x < 32) {
    float s = threadIdx.x < 8 ? part[0][threadIdx.x] : 0.f, t = threadIdx.x < 8 ? part[1][threadIdx.x] : 0.f;
    s = lnst_warp_sum(s); t = lnst_warp_sum(t);
    if (threadIdx.x == 0) { atomicAdd(sums, s); atomicAdd(sums + 1, t); }
  }
}
// pass 2: loss[0..n_loss) += w * (S^2 + 1e3 * L);  grad += g_w * [-1 <= var <= 1] * (2 S - 1e3 sign(dv) / (|dv| + 1e-6))
__global__ void __launch_bounds__(256) density_reg_apply_k(const float* __restrict__ var, int64_t n,
                                                           const float* __restrict__ sums, float w, float g_w,
                                                           float* __restrict__ loss, int n_loss, float* __restrict__ grad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const float S = sums[0];
  if (i == 0 && loss) {
    const float v = w * (S * S + 1e3f * sums[1]);
    for (int k = 0; k < n_loss; ++k) loss[k] += v;
  }
  if (i >= n || !grad) return;
  const float x = var[i];
  if (!(x >= -1.f && x <= 1.f)) return;                  // clip has zero gradient outside (and for NaN)
  const float sg = x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f);
  grad[i] += g_w * (2.f * S - 1e3f * sg / (fabsf(x) + 1e-6f));
}

extern "C" int lnst_pressure_reg(const float* d, int64_t cells, float rest_density, float w_mean, float g_scale,
                                 float* loss, int32_t n_loss, float* g_d, void* stream) {
  if (!d || cells < 1 || n_loss < 0 || (n_loss > 0 && !loss)) return LNST_EARG;
  const unsigned blocks = lnst_blocks(cells, 256) < 148u * 8u ? lnst_blocks(cells, 256) : 148u * 8u;
  LNST_LAUNCH(pressure_reg_k, dim3(blocks), dim3(256), 0, lnst_stream(stream), d, cells, rest_density,
              w_mean / (float)cells, g_scale, loss, (int)n_loss, g_d);
  return lnst_status();
}

extern "C" int lnst_density_reg(const float* var, int64_t n, float weight, float g_weight, float* sums, float* loss,
                                int32_t n_loss, float* grad, void* stream) {
  if (n < 0 || !sums || n_loss < 0 || (n_loss > 0 && !loss)) return LNST_EARG;
  cudaMemsetAsync(sums, 0, 2 * sizeof(float), lnst_stream(stream));
  if (n > 0) {
    if (!var) return LNST_EARG;
    const unsigned blocks = lnst_blocks(n, 256) < 148u * 8u ? lnst_blocks(n, 256) : 148u * 8u;
    LNST_LAUNCH(density_reg_sums_k, dim3(blocks), dim3(256), 0, lnst_stream(stream), var, n, sums);
  }
  LNST_LAUNCH(density_reg_apply_k, dim3(n > 0 ? lnst_blocks(n, 256) : 1u), dim3(256), 0, lnst_stream(stream), var, n,
              (const float*)sums, weight, g_weight, loss, (int)n_loss, grad);
  return lnst_status();
}

// ---- library queries -----------------------------------------------------------------------------------
extern "C" const char* lnst_version(void) { return "lnst-b200 2.0 (sm_100a)"; }

// Scratch a caller must provide to one call of entry point `op` for the given problem size, in bytes.  The library
// allocates nothing itself; every buffer an entry point takes beyond its inputs and outputs is named in its
// signature, and this returns their total so a host runtime can size an arena before the first call.
//   dims: op-specific, see include/lnst_b200.h (unknown op: -1)
extern "C" int64_t lnst_workspace_bytes(const char* op, const int64_t* dims, int32_t n_dims) {
  if (!op) return -1;
  auto is = [&](const char* s) { const char* a = op; while (*a && *a == *s) { ++a; ++s; } return *a == 0 && *s == 0; };
  auto dim = [&](int i) { return (dims && i < n_dims) ? dims[i] : 0; };
  if (is("splat_wavg_fwd")) return 4 * dim(0) * dim(1);            // num [nk, V]: dims = {V, nk}
  if (is("splat_wavg_fwd_tiled")) return 0;                        // tile accumulation lives in shared memory
  if (is("raymarch_fwd") || is("raymarch_bwd")) return 8 * dim(0) * dim(1);   // ray intervals int2 [nv, P]: {P, nv}
  if (is("image_max")) return 8 * dim(0);                          // stats [2 * n_img]
  if (is("normalize_bwd")) return 4 * dim(0);                      // dots [n_img]
  if (is("density_reg")) return 8;                                 // sums [2]
  if (is("adam_step_dev") || is("adam_iterate_dev")) return 12;    // state {beta1^t, beta2^t, lr_t}
  if (is("gram_diff_bf16_tc")) return 4 * dim(0) * dim(1) * dim(1);   // fp32 Gram accumulator [n, C, C]: {n, C}
  if (is("smooth3_relu_fwd") || is("smooth3_relu_bwd") || is("advect") || is("g2p") || is("rk4_advect") ||
      is("adam_step") || is("splat_sph_fwd") || is("splat_sph_bwd_pos") || is("splat_wavg_bwd") || is("rotate_fwd") ||
      is("rotate_bwd") || is("conv3x3_bf16_tc") || is("conv3x3_f32") || is("tv_loss") || is("content_loss") ||
      is("pressure_reg") || is("pressure_loss"))
    return 0;
  return -1;
}
