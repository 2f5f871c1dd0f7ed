// Grid -> particle gathers and the resimulation (data-prep) step that feeds the stylisation path:
// transform.py:771-1231 (g2p / g2p_cubic / g2p_linear) and test_smokegun_resim.py:36-77 (RK4 particle
// advection through the velocity grid, pressure loss of the re-splatted density).
//
// All of it is gather work: one thread per particle, the 64 (cubic) or 8 (linear) taps of all C
// channels of a particle sit next to each other in the [.., C] grid rows, neighbouring particles
// (cell-sorted by the host) hit the same L2 sectors.  The RK4 kernel keeps a particle in registers over
// its four velocity samples -- the reference builds four separate g2p sub-graphs (4 x 64 gathers of
// [N,3] rows each) -- and writes only x_adv.
#include "common.cuh"

#define LNST_G2P_MAXC 4

struct G2PDims { int n[3]; int dim; };

// transform.py:972-978 (_hermite): Catmull-Rom through B..C with tangents from A and D.  TensorFlow evaluates the
// expression one op at a time (every product and sum rounded to fp32), so the arithmetic is written with
// __fmul_rn/__fadd_rn: nvcc's default FMA contraction skips the rounding of the products and moved the 3-D cubic
// gather 3.4e-6 (relative) away from the reference vectors on the B200 while the CPU interpreter matched them.
__device__ __forceinline__ float g2p_hermite(float A, float B, float C, float D, float t) {
  const float a = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(A, -0.5f), __fmul_rn(B, 1.5f)), __fmul_rn(C, -1.5f)),
                            __fmul_rn(D, 0.5f));
  const float b = __fadd_rn(__fadd_rn(__fadd_rn(A, __fmul_rn(B, -2.5f)), __fmul_rn(C, 2.0f)), __fmul_rn(D, -0.5f));
  const float c = __fadd_rn(__fmul_rn(A, -0.5f), __fmul_rn(C, 0.5f));
  const float at3 = __fmul_rn(__fmul_rn(__fmul_rn(a, t), t), t);
  const float bt2 = __fmul_rn(__fmul_rn(b, t), t);
  return __fadd_rn(__fadd_rn(__fadd_rn(at3, bt2), __fmul_rn(c, t)), B);
}

// Clamped tap indices and the fractional offset of one axis (transform.py:812-838 cubic, :1144-1162 linear).
// NB the reference measures the offset from the CLAMPED index (x1 is reassigned by clip_by_value before
// `dx = x - (x1 + 0.5)`, :999 / :1200), so particles outside the grid extrapolate; restated as is.
template <int TAPS>
__device__ __forceinline__ float g2p_axis(float pos01, int len, int* idx) {
  const float x = __fmul_rn(pos01, (float)len);   // rounded product: floorf(x - 0.5f) must not see an FMA
  const int f = (int)floorf(x - 0.5f);
  const int first = TAPS == 4 ? f - 1 : f;
#pragma unroll
  for (int i = 0; i < TAPS; ++i) idx[i] = min(max(first + i, 0), len - 1);
  const int anchor = TAPS == 4 ? idx[1] : idx[0];
  return x - ((float)anchor + 0.5f);
}

// Sample all C channels of grid g ([n0,n1,(n2),C]) at normalised position pos ((z,)y,x order = grid axes).
template <int DIM, bool LINEAR>
__device__ __forceinline__ void g2p_sample(const float* __restrict__ g, const G2PDims& d, int C, int nc,
                                           const float* pos, float* out) {
  constexpr int TAPS = LINEAR ? 2 : 4;
  int ix[TAPS], iy[TAPS], iz[TAPS];
  const float tx = g2p_axis<TAPS>(pos[0], d.n[0], ix);
  const float ty = g2p_axis<TAPS>(pos[1], d.n[1], iy);
  float tz = 0.f;
  if (DIM == 3) tz = g2p_axis<TAPS>(pos[2], d.n[2], iz);
  else {
#pragma unroll
    for (int i = 0; i < TAPS; ++i) iz[i] = 0;
  }
  const int64_t s0 = (int64_t)d.n[1] * (DIM == 3 ? d.n[2] : 1), s1 = DIM == 3 ? d.n[2] : 1;
#pragma unroll 1
  for (int c = 0; c < nc; ++c) {                      // C = row stride in floats, nc = channels sampled
    if (LINEAR) {
      // transform.py:1196-1227: weights (1-dx)(1-dy)[(1-dz)] ..., add_n in corner order 000,001,010,...
      float o = 0.f;
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const float wab = __fmul_rn(a ? tx : 1.f - tx, b ? ty : 1.f - ty);
          if (DIM == 3) {
#pragma unroll
            for (int e = 0; e < 2; ++e)
              o = __fadd_rn(o, __fmul_rn(__fmul_rn(wab, e ? tz : 1.f - tz), g[(ix[a] * s0 + iy[b] * s1 + iz[e]) * C + c]));
          } else {
            o = __fadd_rn(o, __fmul_rn(wab, g[(ix[a] * s0 + iy[b] * s1) * C + c]));
          }
        }
      out[c] = o;
    } else if (DIM == 3) {
      // :1078-1102: Hermite along the first axis for each (y,z) pair, then along y, then along z.  The z
      // planes are walked one at a time (16 loads in flight, a rotating 4-entry window of plane results):
      // fully unrolled, the 64 taps x 64-bit addresses need > 255 registers.
      float Iz0 = 0.f, Iz1 = 0.f, Iz2 = 0.f, Iz3 = 0.f;
      const int zfirst = (int)floorf(__fmul_rn(pos[2], (float)d.n[2]) - 0.5f) - 1;
#pragma unroll 1
      for (int e = 0; e < 4; ++e) {
        const int ze = min(max(zfirst + e, 0), d.n[2] - 1);
        float Iy[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int64_t o = iy[b] * s1 + ze;
          Iy[b] = g2p_hermite(g[(ix[0] * s0 + o) * C + c], g[(ix[1] * s0 + o) * C + c], g[(ix[2] * s0 + o) * C + c],
                              g[(ix[3] * s0 + o) * C + c], tx);
        }
        Iz0 = Iz1; Iz1 = Iz2; Iz2 = Iz3;
        Iz3 = g2p_hermite(Iy[0], Iy[1], Iy[2], Iy[3], ty);
      }
      out[c] = g2p_hermite(Iz0, Iz1, Iz2, Iz3, tz);
    } else {
      float Iy[4];
#pragma unroll
      for (int b = 0; b < 4; ++b)
        Iy[b] = g2p_hermite(g[(ix[0] * s0 + iy[b]) * C + c], g[(ix[1] * s0 + iy[b]) * C + c],
                            g[(ix[2] * s0 + iy[b]) * C + c], g[(ix[3] * s0 + iy[b]) * C + c], tx);
      out[c] = g2p_hermite(Iy[0], Iy[1], Iy[2], Iy[3], ty);
    }
  }
}

template <int DIM, bool LINEAR>
__global__ void __launch_bounds__(128) g2p_k(const float* __restrict__ g, G2PDims d, int C, const float* __restrict__ p,
                                             const float* __restrict__ disp, int64_t n, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float pos[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < DIM; ++k) pos[k] = p[i * DIM + k] + (disp ? disp[i * DIM + k] : 0.f);
  for (int c0 = 0; c0 < C; c0 += LNST_G2P_MAXC) {          // C is 1 (density) or dim (velocity) in the drivers
    float o[LNST_G2P_MAXC];
    const int cc = min(C - c0, LNST_G2P_MAXC);
    g2p_sample<DIM, LINEAR>(g + c0, d, C, cc, pos, o);
    for (int c = 0; c < cc; ++c) out[i * C + c0 + c] = o[c];
  }
}

// test_smokegun_resim.py:36-55: v = g2p(u,x); v1 = g2p(u, x + v/2); v2 = g2p(u, x + v1/2); v3 = g2p(u, x + v2);
// x_adv = x + time_step * (v + 2 v1 + 2 v2 + v3)/6.  u has DIM channels in the order of the position axes.
template <int DIM, bool LINEAR>
__global__ void __launch_bounds__(128) rk4_advect_k(const float* __restrict__ u, G2PDims d, const float* __restrict__ x,
                                                    int64_t n, float time_step, float* __restrict__ x_adv,
                                                    float* __restrict__ v_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float x0[3] = {0.f, 0.f, 0.f}, xs[3] = {0.f, 0.f, 0.f};
  float v[LNST_G2P_MAXC], v1[LNST_G2P_MAXC], v2[LNST_G2P_MAXC], v3[LNST_G2P_MAXC];
#pragma unroll
  for (int k = 0; k < DIM; ++k) x0[k] = x[i * DIM + k];
  g2p_sample<DIM, LINEAR>(u, d, DIM, DIM, x0, v);
#pragma unroll
  for (int k = 0; k < DIM; ++k) xs[k] = __fadd_rn(x0[k], __fmul_rn(v[k], 0.5f));
  g2p_sample<DIM, LINEAR>(u, d, DIM, DIM, xs, v1);
#pragma unroll
  for (int k = 0; k < DIM; ++k) xs[k] = __fadd_rn(x0[k], __fmul_rn(v1[k], 0.5f));
  g2p_sample<DIM, LINEAR>(u, d, DIM, DIM, xs, v2);
#pragma unroll
  for (int k = 0; k < DIM; ++k) xs[k] = x0[k] + v2[k];
  g2p_sample<DIM, LINEAR>(u, d, DIM, DIM, xs, v3);
#pragma unroll
  for (int k = 0; k < DIM; ++k) {
    const float vm = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(v[k], __fmul_rn(v1[k], 2.f)), __fmul_rn(v2[k], 2.f)), v3[k]), 6.f);
    x_adv[i * DIM + k] = __fadd_rn(x0[k], __fmul_rn(vm, time_step));
    if (v_out) v_out[i * DIM + k] = vm;
  }
}

// test_smokegun_resim.py:69-74: pressure = where(d_rec > 0, d_rec - rho0, 0); loss = mean(pressure^2).
// One pass: block-reduced loss (one atomic per block) and the gradient 2*w*pressure/cells in place of d_rec's
// cotangent.  `loss` accumulates (caller zeroes it).
__global__ void __launch_bounds__(256) pressure_loss_k(const float* __restrict__ d_rec, int64_t cells, float rho0,
                                                       float weight, float* __restrict__ loss,
                                                       float* __restrict__ g_d) {
  __shared__ float part[8];
  const float inv = 1.0f / (float)cells;
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += (int64_t)gridDim.x * blockDim.x) {
    const float dv = d_rec[i];
    const float pr = dv > 0.f ? dv - rho0 : 0.f;
    acc += pr * pr;
    if (g_d) g_d[i] = 2.f * weight * inv * pr;
  }
  acc = lnst_warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float s = threadIdx.x < 8 ? part[threadIdx.x] : 0.f;
    s = lnst_warp_sum(s);
    if (threadIdx.x == 0) atomicAdd(loss, s * inv * weight);
  }
}

// out[z,y,x] = a[z,y,x] - b[z,H-1-y,x]: the residual between a density grid and a splatted field, whose H
// axis is stored flipped (test_smokegun_resim.py:92 `d - d_hi[:,:,::-1]`, :106 d_diff).  b == NULL: out = a.
__global__ void sub_fliph_k(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int D,
                            int H, int W) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t cells = (int64_t)D * H * W;
  if (t >= cells) return;
  const int x = (int)(t % W), y = (int)((t / W) % H);
  const int64_t z = t / ((int64_t)W * H);
  out[t] = a[t] - b[(z * H + (H - 1 - y)) * W + x];
}

// ---------------------------------------------------------------------------------------------------
static bool g2p_dims(G2PDims& d, int32_t dim, const int32_t* dims) {
  if (!dims || (dim != 2 && dim != 3)) return false;
  d.dim = dim;
  for (int k = 0; k < 3; ++k) {
    d.n[k] = k < dim ? dims[k] : 1;
    if (d.n[k] < 1) return false;
  }
  return true;
}

extern "C" int lnst_g2p(const float* g, int32_t dim, const int32_t* dims, int32_t C, const float* p,
                        const float* disp, int64_t n, int32_t linear, float* out, void* stream) {
  G2PDims d;
  if (n < 0 || C < 1 || !g2p_dims(d, dim, dims)) return LNST_EARG;
  if (n == 0) return LNST_OK;                       // an empty particle set has no storage to point at
  if (!g || !p || !out) return LNST_EARG;
  const dim3 grid_(lnst_blocks(n, 128)), blk(128);
  if (dim == 3) {
    if (linear) { auto k = g2p_k<3, true>; LNST_LAUNCH(k, grid_, blk, 0, lnst_stream(stream), g, d, (int)C, p, disp, n, out); }
    else { auto k = g2p_k<3, false>; LNST_LAUNCH(k, grid_, blk, 0, lnst_stream(stream), g, d, (int)C, p, disp, n, out); }
  } else {
    if (linear) { auto k = g2p_k<2, true>; LNST_LAUNCH(k, grid_, blk, 0, lnst_stream(stream), g, d, (int)C, p, disp, n, out); }
    else { auto k = g2p_k<2, false>; LNST_LAUNCH(k, grid_, blk, 0, lnst_stream(stream), g, d, (int)C, p, disp, n, out); }
  }
  return lnst_status();
}

extern "C" int lnst_rk4_advect(const float* u, int32_t dim, const int32_t* dims, const float* x, int64_t n,
                               float time_step, int32_t linear, float* x_adv, float* v_out, void* stream) {
  G2PDims d;
  if (n < 0 || !g2p_dims(d, dim, dims)) return LNST_EARG;
  if (n == 0) return LNST_OK;
  if (!u || !x || !x_adv) return LNST_EARG;
  const dim3 grid_(lnst_blocks(n, 128)), blk(128);
  if (dim == 3) {
    if (linear) { auto k = rk4_advect_k<3, true>; LNST_LAUNCH(k, grid_, blk, 0, lnst_stream(stream), u, d, x, n, time_step, x_adv, v_out); }
    else { auto k = rk4_advect_k<3, false>; LNST_LAUNCH(k, grid_, blk, 0, lnst_stream(stream), u, d, x, n, time_step, x_adv, v_out); }
  } else {
    if (linear) { auto k = rk4_advect_k<2, true>; LNST_LAUNCH(k, grid_, blk, 0, lnst_stream(stream), u, d, x, n, time_step, x_adv, v_out); }
    else { auto k = rk4_advect_k<2, false>; LNST_LAUNCH(k, grid_, blk, 0, lnst_stream(stream), u, d, x, n, time_step, x_adv, v_out); }
  }
  return lnst_status();
}

extern "C" int lnst_pressure_loss(const float* d_rec, int64_t cells, float rest_density, float weight, float* loss,
                                  float* g_d, void* stream) {
  if (!d_rec || !loss || cells < 1) return LNST_EARG;
  const unsigned blocks = lnst_blocks(cells, 256) < 148u * 8u ? lnst_blocks(cells, 256) : 148u * 8u;   // grid-stride, 8 CTAs/SM
  LNST_LAUNCH(pressure_loss_k, dim3(blocks), dim3(256), 0, lnst_stream(stream), d_rec, cells, rest_density, weight,
              loss, g_d);
  return lnst_status();
}

extern "C" int lnst_sub_fliph(const float* a, const float* b, float* out, int32_t D, int32_t H, int32_t W,
                              void* stream) {
  if (!a || !b || !out || D < 1 || H < 1 || W < 1) return LNST_EARG;
  const int64_t cells = (int64_t)D * H * W;
  LNST_LAUNCH(sub_fliph_k, dim3(lnst_blocks(cells, 256)), dim3(256), 0, lnst_stream(stream), a, b, out, (int)D, (int)H,
              (int)W);
  return lnst_status();
}
