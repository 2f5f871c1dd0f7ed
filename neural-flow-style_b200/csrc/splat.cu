// Particle -> grid splatting kernels (reference: transform.py:1233-1245 W, :1310-1453 p2g,
// :1577-1704 p2g_wavg).  One thread per particle; the (2 nsize+1)^dim target cells are
// visited in registers, the H axis is written flipped (transform.py:1404,1452).
//
// Arithmetic notes (parity with the fp32 oracle): the position -> (cell, offset) part uses
// explicitly un-fused multiplies/adds (__fmul_rn/__fadd_rn) because a contracted FMA changes
// the offset by up to one ulp of the (large) domain coordinate.
#include "common.cuh"
#include "splat_common.cuh"

// Visit every in-range target cell of a particle: f(cell_linear_index_with_H_flip, d[DIM], |d|)
template <int DIM, class F>
__device__ __forceinline__ void for_each_target(const Particle<DIM>& pt, const LnstGrid& g, F f) {
  const int ns = g.nsize;
  const int H = g.res[1], W = g.res[2];
  if (DIM == 3) {
    const int D = g.res[0];
    for (int sz = -ns; sz <= ns; ++sz) {
      const int z = pt.idx[0] + sz;
      const float dz = __fadd_rn(pt.r[0], -__fmul_rn((float)sz, g.cell));
      for (int sy = -ns; sy <= ns; ++sy) {
        const int y = pt.idx[1] + sy;
        const float dy = __fadd_rn(pt.r[1], -__fmul_rn((float)sy, g.cell));
        for (int sx = -ns; sx <= ns; ++sx) {
          const int x = pt.idx[2] + sx;
          if (z < 0 || z >= D || y < 0 || y >= H || x < 0 || x >= W) continue;
          const float dx = __fadd_rn(pt.r[2], -__fmul_rn((float)sx, g.cell));
          float d[3] = {dz, dy, dx};
          const float len = sqrtf(dx * dx + dy * dy + dz * dz);
          f(((int64_t)z * H + (H - 1 - y)) * W + x, d, len);
        }
      }
    }
  } else {
    for (int sy = -ns; sy <= ns; ++sy) {
      const int y = pt.idx[0] + sy;
      const float dy = __fadd_rn(pt.r[0], -__fmul_rn((float)sy, g.cell));
      for (int sx = -ns; sx <= ns; ++sx) {
        const int x = pt.idx[1] + sx;
        if (y < 0 || y >= H || x < 0 || x >= W) continue;
        const float dx = __fadd_rn(pt.r[1], -__fmul_rn((float)sx, g.cell));
        float d[3] = {dy, dx, 0.f};
        const float len = sqrtf(dx * dx + dy * dy);
        f((int64_t)(H - 1 - y) * W + x, d, len);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// SPH splat (p2g)
// ---------------------------------------------------------------------------------------
template <int DIM>
__global__ void splat_sph_fwd_k(const float* __restrict__ p, const float* __restrict__ disp, int64_t n,
                                LnstGrid g, float h, float sigma, float scale,
                                const float* __restrict__ pc, const float* __restrict__ pd, int C,
                                float rho, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Particle<DIM> pt = load_particle<DIM>(p, disp, i, g);
  if (!pt.valid) return;
  const float inv_h = 1.f / h;
  if (pc == nullptr) {
    for_each_target<DIM>(pt, g, [&](int64_t cell, const float*, float len) {
      const float w = cubic_w(len * inv_h, sigma);
      if (w != 0.f) atomicAdd(out + cell, scale * w);
    });
  } else {
    const float den = pd ? pd[i] : rho;
    for_each_target<DIM>(pt, g, [&](int64_t cell, const float*, float len) {
      const float w = cubic_w(len * inv_h, sigma);
      if (w != 0.f)
        for (int c = 0; c < C; ++c) atomicAdd(out + cell * C + c, scale * w * pc[i * C + c] / den);
    });
  }
}

template <int DIM>
__global__ void splat_sph_bwd_pos_k(const float* __restrict__ p, const float* __restrict__ disp,
                                    int64_t n, LnstGrid g, float h, float sigma, float scale,
                                    const float* __restrict__ g_out, float* __restrict__ g_p) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Particle<DIM> pt = load_particle<DIM>(p, disp, i, g);
  float acc[3] = {0.f, 0.f, 0.f};
  if (pt.valid) {
    const float inv_h = 1.f / h;
    for_each_target<DIM>(pt, g, [&](int64_t cell, const float* d, float len) {
      if (len > 0.f) {   // sqrt'(0): TF gives NaN, defined as 0 here
        const float dw = cubic_dw(len * inv_h, sigma);
        const float c = g_out[cell] * scale * dw * inv_h / len;
        acc[0] += c * d[0];
        acc[1] += c * d[1];
        acc[2] += c * d[2];
      }
    });
  }
#pragma unroll
  for (int a = 0; a < DIM; ++a) g_p[i * DIM + a] = acc[a] * pt.dpd[a];
}

template <int DIM>
__global__ void splat_sph_bwd_color_k(const float* __restrict__ p, int64_t n, LnstGrid g, float h,
                                      float sigma, float scale, const float* __restrict__ pd, int C,
                                      float rho, const float* __restrict__ g_out,
                                      float* __restrict__ g_pc) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Particle<DIM> pt = load_particle<DIM>(p, nullptr, i, g);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (pt.valid) {
    const float inv_h = 1.f / h;
    const float den = pd ? pd[i] : rho;
    for_each_target<DIM>(pt, g, [&](int64_t cell, const float*, float len) {
      const float w = cubic_w(len * inv_h, sigma);
      if (w != 0.f)
        for (int c = 0; c < C; ++c) acc[c] += g_out[cell * C + c] * (scale * w / den);
    });
  }
  for (int c = 0; c < C; ++c) g_pc[i * C + c] = acc[c];
}

// ---------------------------------------------------------------------------------------
// weighted-average splat (p2g_wavg)
// ---------------------------------------------------------------------------------------
template <int DIM>
__global__ void splat_wavg_wmap_k(const float* __restrict__ p, int64_t n, LnstGrid g, SplatKernels ks,
                                  int nk, int64_t cells, float* __restrict__ wmap) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Particle<DIM> pt = load_particle<DIM>(p, nullptr, i, g);
  if (!pt.valid) return;
  for_each_target<DIM>(pt, g, [&](int64_t cell, const float*, float len) {
    for (int k = 0; k < nk; ++k) {
      const float w = cubic_w(len * ks.inv_h[k], ks.sigma[k]);
      if (w != 0.f) atomicAdd(wmap + k * cells + cell, w);
    }
  });
}

template <int DIM>
__global__ void splat_wavg_num_k(const float* __restrict__ p, const float* __restrict__ r,
                                 const float* __restrict__ var, int64_t n, LnstGrid g, SplatKernels ks,
                                 int nk, int64_t cells, float* __restrict__ num) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Particle<DIM> pt = load_particle<DIM>(p, nullptr, i, g);
  if (!pt.valid) return;
  float x[LNST_MAX_NK];
  for (int k = 0; k < nk; ++k) {
    float v = var ? var[i * nk + k] : 0.f;
    v = fmaxf(fminf(v, 1.f), -1.f);               // styler_3p.py:74; TF order max(min(x,1),-1): NaN reads as +1
    x[k] = r[i * nk + k] + v;                     // :76
  }
  for_each_target<DIM>(pt, g, [&](int64_t cell, const float*, float len) {
    for (int k = 0; k < nk; ++k) {
      const float w = cubic_w(len * ks.inv_h[k], ks.sigma[k]);
      if (w != 0.f) atomicAdd(num + k * cells + cell, w * x[k]);
    }
  });
}

__global__ void splat_wavg_combine_k(const float* __restrict__ wmap, const float* __restrict__ num,
                                     int nk, int64_t cells, float* __restrict__ out) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cells) return;
  float s = 0.f;
  for (int k = 0; k < nk; ++k) {
    const float w = wmap[k * cells + c];
    const float v = num[k * cells + c];
    s += (w > 1e-6f) ? v / w : v;                 // transform.py:1703
  }
  out[c] = s;
}

// ---- 3-D, nsize = 1, NK kernels: the density-mode hot path (styler_3p.py:79-87) -------------------
// Fully unrolled 27-cell stencil with one 32-bit anchor; particles whose stencil leaves the grid take
// the generic path.  Same arithmetic as for_each_target (offsets, sum of squares, sqrtf), so a cell
// gets the same weight from either path and from the wmap / num / gradient kernels alike.
template <class F>
__device__ __forceinline__ void stencil27(const Particle<3>& pt, const LnstGrid& g, F f) {
  const int H = g.res[1], W = g.res[2];
  const int base = (pt.idx[0] * H + (H - 1 - pt.idx[1])) * W + pt.idx[2];
  float dz[3], dy[3], dx[3];
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    const float o = __fmul_rn((float)(s - 1), g.cell);
    dz[s] = __fadd_rn(pt.r[0], -o);
    dy[s] = __fadd_rn(pt.r[1], -o);
    dx[s] = __fadd_rn(pt.r[2], -o);
  }
#pragma unroll
  for (int sz = 0; sz < 3; ++sz)
#pragma unroll
    for (int sy = 0; sy < 3; ++sy)
#pragma unroll
      for (int sx = 0; sx < 3; ++sx) {
        const float len = sqrtf(dx[sx] * dx[sx] + dy[sy] * dy[sy] + dz[sz] * dz[sz]);
        f(base + (sz - 1) * H * W - (sy - 1) * W + (sx - 1), len);
      }
}
__device__ __forceinline__ bool stencil_inside(const Particle<3>& pt, const LnstGrid& g) {
  return pt.idx[0] >= 1 && pt.idx[0] < g.res[0] - 1 && pt.idx[1] >= 1 && pt.idx[1] < g.res[1] - 1 &&
         pt.idx[2] >= 1 && pt.idx[2] < g.res[2] - 1;
}

template <int NK>
__global__ void __launch_bounds__(256) splat_wavg_num3_k(const float* __restrict__ p, const float* __restrict__ r,
                                                         const float* __restrict__ var, int64_t n, LnstGrid g,
                                                         SplatKernels ks, int64_t cells, float* __restrict__ num) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Particle<3> pt = load_particle<3>(p, nullptr, i, g);
  if (!pt.valid) return;
  float x[NK];
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    float v = var ? var[i * NK + k] : 0.f;
    v = fmaxf(fminf(v, 1.f), -1.f);               // styler_3p.py:74; TF order max(min(x,1),-1): NaN reads as +1
    x[k] = r[i * NK + k] + v;                     // :76
  }
  if (stencil_inside(pt, g)) {
    stencil27(pt, g, [&](int cell, float len) {
#pragma unroll
      for (int k = 0; k < NK; ++k) {
        const float w = cubic_w(len * ks.inv_h[k], ks.sigma[k]);
        if (w != 0.f) atomicAdd(num + k * cells + cell, w * x[k]);
      }
    });
  } else {
    for_each_target<3>(pt, g, [&](int64_t cell, const float*, float len) {
#pragma unroll
      for (int k = 0; k < NK; ++k) {
        const float w = cubic_w(len * ks.inv_h[k], ks.sigma[k]);
        if (w != 0.f) atomicAdd(num + k * cells + cell, w * x[k]);
      }
    });
  }
}

// coef = d where(wm > eps, num/wm, num) / d num as TF computes it (transform.py:1703): 1/wm, 1, or NaN
// where wm == 0 (the untaken division branch contributes 0/0).  Positions are constants in density
// mode, so this is computed once per (frame, octave) next to wmap.
__global__ void splat_wavg_coef_k(const float* __restrict__ wmap, int64_t total, float* __restrict__ coef) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const float wm = wmap[i];
  coef[i] = (wm > 1e-6f) ? 1.f / wm : (wm == 0.f ? __int_as_float(0x7fc00000) : 1.f);
}

template <int NK>
__global__ void __launch_bounds__(256) splat_wavg_bwd3_k(const float* __restrict__ p, const float* __restrict__ var,
                                                         int64_t n, LnstGrid g, SplatKernels ks, int64_t cells,
                                                         const float* __restrict__ coef,
                                                         const float* __restrict__ g_out, float* __restrict__ g_var) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Particle<3> pt = load_particle<3>(p, nullptr, i, g);
  float acc[NK];
#pragma unroll
  for (int k = 0; k < NK; ++k) acc[k] = 0.f;
  if (pt.valid) {
    if (stencil_inside(pt, g)) {
      stencil27(pt, g, [&](int cell, float len) {
        const float go = g_out[cell];
#pragma unroll
        for (int k = 0; k < NK; ++k)
          acc[k] += cubic_w(len * ks.inv_h[k], ks.sigma[k]) * (coef[k * cells + cell] * go);
      });
    } else {
      for_each_target<3>(pt, g, [&](int64_t cell, const float*, float len) {
        const float go = g_out[cell];
#pragma unroll
        for (int k = 0; k < NK; ++k)
          acc[k] += cubic_w(len * ks.inv_h[k], ks.sigma[k]) * (coef[k * cells + cell] * go);
      });
    }
  }
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    const float v = var ? var[i * NK + k] : 0.f;
    g_var[i * NK + k] = (v >= -1.f && v <= 1.f) ? acc[k] : 0.f;   // clip_by_value gradient
  }
}

// box variant: combines only the cells of the sub-volume and clears the num it consumed
__global__ void splat_wavg_combine_box_k(const float* __restrict__ wmap, float* __restrict__ num, int nk,
                                         int64_t cells, int H, int W, SubVol sv, float* __restrict__ out) {
  // 32-bit index arithmetic (a volume has fewer than 2^31 cells): the 64-bit divisions cost more than the
  // five memory accesses of a cell
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (unsigned)(sv.ez * sv.ey * sv.ex)) return;
  const unsigned q = t / (unsigned)sv.ex;
  const int x = sv.ox + (int)(t - q * (unsigned)sv.ex);
  const unsigned zq = q / (unsigned)sv.ey;
  const int y = sv.oy + (int)(q - zq * (unsigned)sv.ey), z = sv.oz + (int)zq;
  const int64_t c = ((int64_t)z * H + y) * W + x;
  float s = 0.f;
  for (int k = 0; k < nk; ++k) {
    const float w = wmap[k * cells + c];
    const float v = num[k * cells + c];
    if (v != 0.f) num[k * cells + c] = 0.f;
    s += (w > 1e-6f) ? v / w : v;                 // transform.py:1703
  }
  out[c] = s;
}

template <int DIM>
__global__ void splat_wavg_bwd_k(const float* __restrict__ p, const float* __restrict__ var, int64_t n,
                                 LnstGrid g, SplatKernels ks, int nk, int64_t cells,
                                 const float* __restrict__ wmap, const float* __restrict__ g_out,
                                 float* __restrict__ g_var) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Particle<DIM> pt = load_particle<DIM>(p, nullptr, i, g);
  float acc[LNST_MAX_NK] = {0.f, 0.f, 0.f, 0.f};
  if (pt.valid) {
    for_each_target<DIM>(pt, g, [&](int64_t cell, const float*, float len) {
      const float go = g_out[cell];
      for (int k = 0; k < nk; ++k) {
        const float w = cubic_w(len * ks.inv_h[k], ks.sigma[k]);
        const float wm = wmap[k * cells + cell];
        // gradient of where(wm>eps, num/wm, num) w.r.t. num as TF computes it: the untaken
        // division branch contributes 0/wm, which is NaN when wm == 0 (transform.py:1703).
        const float coef = (wm > 1e-6f) ? 1.f / wm : (wm == 0.f ? __int_as_float(0x7fc00000) : 1.f);
        acc[k] += w * (coef * go);
      }
    });
  }
  for (int k = 0; k < nk; ++k) {
    const float v = var ? var[i * nk + k] : 0.f;
    const float pass = (v >= -1.f && v <= 1.f) ? 1.f : 0.f;   // clip_by_value gradient
    g_var[i * nk + k] = (pass != 0.f) ? acc[k] : 0.f;
  }
}

// ---------------------------------------------------------------------------------------
// C-ABI
// ---------------------------------------------------------------------------------------
// home cell (linear index over [D,H,W], y NOT flipped; -1: outside the domain / padding row, -2: valid but its cell index
// rounded onto the grid's far face) and offset from that cell's centre, for the per-cell particle lists of the gather splat
__global__ void splat_cells_k(const float* __restrict__ p, int64_t n, LnstGrid g, int* __restrict__ cell,
                              float* __restrict__ rel) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Particle<3> pt = load_particle<3>(p, nullptr, i, g);
  int c = -1;
  if (pt.valid) {
    const bool in = pt.idx[0] >= 0 && pt.idx[0] < g.res[0] && pt.idx[1] >= 0 && pt.idx[1] < g.res[1] && pt.idx[2] >= 0 &&
                    pt.idx[2] < g.res[2];
    c = in ? (pt.idx[0] * g.res[1] + pt.idx[1]) * g.res[2] + pt.idx[2] : -2;
  }
  cell[i] = c;
  rel[i * 3] = pt.r[0]; rel[i * 3 + 1] = pt.r[1]; rel[i * 3 + 2] = pt.r[2];
}

extern "C" int lnst_splat_cells(const float* p, int64_t n, const LnstGrid* g, int32_t* cell, float* rel, void* stream) {
  if (!grid_ok(g) || g->dim != 3 || n < 0 || (n > 0 && (!p || !cell || !rel))) return LNST_EARG;
  if (n == 0) return LNST_OK;
  if (grid_cells(g) >= 0x7fffffff) return LNST_EARG;
  LNST_LAUNCH(splat_cells_k, dim3(lnst_blocks(n, 256)), dim3(256), 0, lnst_stream(stream), p, n, *g, (int*)cell, rel);
  return lnst_status();
}

extern "C" int lnst_splat_sph_fwd(const float* p, const float* disp, int64_t n, const LnstGrid* g,
                                  float h, float scale, const float* pc, const float* pd, int32_t C,
                                  float rest_density, float* out, void* stream) {
  if (!grid_ok(g) || (n > 0 && !p) || !out || n < 0 || !(h > 0.f) || (pc && (C < 1 || C > 4))) return LNST_EARG;   // an empty set has no storage
  if (n == 0) return LNST_OK;
  const float sigma = sigma_for(g->dim, h);
  const int T = 256;
  if (g->dim == 3) {
    auto k = splat_sph_fwd_k<3>;
    LNST_LAUNCH(k, dim3(lnst_blocks(n, T)), dim3(T), 0, lnst_stream(stream), p, disp, n, *g, h, sigma,
                scale, pc, pd, (int)C, rest_density, out);
  } else {
    auto k = splat_sph_fwd_k<2>;
    LNST_LAUNCH(k, dim3(lnst_blocks(n, T)), dim3(T), 0, lnst_stream(stream), p, disp, n, *g, h, sigma,
                scale, pc, pd, (int)C, rest_density, out);
  }
  return lnst_status();
}

extern "C" int lnst_splat_sph_bwd_pos(const float* p, const float* disp, int64_t n, const LnstGrid* g,
                                      float h, float scale, const float* g_out, float* g_p,
                                      void* stream) {
  if (!grid_ok(g) || (n > 0 && (!p || !g_p)) || !g_out || n < 0 || !(h > 0.f)) return LNST_EARG;
  if (n == 0) return LNST_OK;
  const float sigma = sigma_for(g->dim, h);
  const int T = 256;
  if (g->dim == 3) {
    auto k = splat_sph_bwd_pos_k<3>;
    LNST_LAUNCH(k, dim3(lnst_blocks(n, T)), dim3(T), 0, lnst_stream(stream), p, disp, n, *g, h, sigma,
                scale, g_out, g_p);
  } else {
    auto k = splat_sph_bwd_pos_k<2>;
    LNST_LAUNCH(k, dim3(lnst_blocks(n, T)), dim3(T), 0, lnst_stream(stream), p, disp, n, *g, h, sigma,
                scale, g_out, g_p);
  }
  return lnst_status();
}

extern "C" int lnst_splat_sph_bwd_color(const float* p, int64_t n, const LnstGrid* g, float h,
                                        float scale, const float* pd, int32_t C, float rest_density,
                                        const float* g_out, float* g_pc, void* stream) {
  if (!grid_ok(g) || (n > 0 && (!p || !g_pc)) || !g_out || n < 0 || !(h > 0.f) || C < 1 || C > 4) return LNST_EARG;
  if (n == 0) return LNST_OK;
  const float sigma = sigma_for(g->dim, h);
  const int T = 256;
  if (g->dim == 3) {
    auto k = splat_sph_bwd_color_k<3>;
    LNST_LAUNCH(k, dim3(lnst_blocks(n, T)), dim3(T), 0, lnst_stream(stream), p, n, *g, h, sigma, scale,
                pd, (int)C, rest_density, g_out, g_pc);
  } else {
    auto k = splat_sph_bwd_color_k<2>;
    LNST_LAUNCH(k, dim3(lnst_blocks(n, T)), dim3(T), 0, lnst_stream(stream), p, n, *g, h, sigma, scale,
                pd, (int)C, rest_density, g_out, g_pc);
  }
  return lnst_status();
}

extern "C" int lnst_splat_wavg_wmap(const float* p, int64_t n, const LnstGrid* g, const float* h,
                                    int32_t nk, float* wmap, void* stream) {
  SplatKernels ks;
  if (!grid_ok(g) || (n > 0 && !p) || !wmap || n < 0 || !fill_kernels(ks, g ? g->dim : 3, h, nk)) return LNST_EARG;
  const int64_t cells = grid_cells(g);
  cudaMemsetAsync(wmap, 0, sizeof(float) * cells * nk, lnst_stream(stream));
  if (n == 0) return lnst_status();
  const int T = 256;
  if (g->dim == 3) {
    auto k = splat_wavg_wmap_k<3>;
    LNST_LAUNCH(k, dim3(lnst_blocks(n, T)), dim3(T), 0, lnst_stream(stream), p, n, *g, ks, (int)nk, cells,
                wmap);
  } else {
    auto k = splat_wavg_wmap_k<2>;
    LNST_LAUNCH(k, dim3(lnst_blocks(n, T)), dim3(T), 0, lnst_stream(stream), p, n, *g, ks, (int)nk, cells,
                wmap);
  }
  return lnst_status();
}

extern "C" int lnst_splat_wavg_fwd(const float* p, const float* r, const float* var, int64_t n,
                                   const LnstGrid* g, const float* h, int32_t nk, const float* wmap,
                                   float* num, float* out, void* stream) {
  return lnst_splat_wavg_fwd_box(p, r, var, n, g, h, nk, wmap, num, out, nullptr, stream);
}

extern "C" int lnst_splat_wavg_fwd_box(const float* p, const float* r, const float* var, int64_t n,
                                       const LnstGrid* g, const float* h, int32_t nk, const float* wmap,
                                       float* num, float* out, const LnstBox* box, void* stream) {
  SplatKernels ks;
  if (!grid_ok(g) || (n > 0 && (!p || !r)) || !wmap || !num || !out || n < 0 ||
      !fill_kernels(ks, g ? g->dim : 3, h, nk))
    return LNST_EARG;
  const int Dz = g->dim == 3 ? g->res[0] : 1;
  if (!box_ok(box, Dz, g->res[1], g->res[2])) return LNST_EARG;
  const int64_t cells = grid_cells(g);
  if (!box) cudaMemsetAsync(num, 0, sizeof(float) * cells * nk, lnst_stream(stream));
  const int T = 256;
  if (n > 0 && g->dim == 3 && g->nsize == 1) {
    const dim3 grid_(lnst_blocks(n, T)), blk(T);
    switch (nk) {
      case 1: { auto k = splat_wavg_num3_k<1>; LNST_LAUNCH(k, grid_, blk, 0, lnst_stream(stream), p, r, var, n, *g, ks, cells, num); break; }
      case 2: { auto k = splat_wavg_num3_k<2>; LNST_LAUNCH(k, grid_, blk, 0, lnst_stream(stream), p, r, var, n, *g, ks, cells, num); break; }
      case 3: { auto k = splat_wavg_num3_k<3>; LNST_LAUNCH(k, grid_, blk, 0, lnst_stream(stream), p, r, var, n, *g, ks, cells, num); break; }
      default: { auto k = splat_wavg_num3_k<4>; LNST_LAUNCH(k, grid_, blk, 0, lnst_stream(stream), p, r, var, n, *g, ks, cells, num); break; }
    }
  } else if (n > 0) {
    if (g->dim == 3) {
      auto k = splat_wavg_num_k<3>;
      LNST_LAUNCH(k, dim3(lnst_blocks(n, T)), dim3(T), 0, lnst_stream(stream), p, r, var, n, *g, ks,
                  (int)nk, cells, num);
    } else {
      auto k = splat_wavg_num_k<2>;
      LNST_LAUNCH(k, dim3(lnst_blocks(n, T)), dim3(T), 0, lnst_stream(stream), p, r, var, n, *g, ks,
                  (int)nk, cells, num);
    }
  }
  if (box) {
    const SubVol sv = make_subvol(box, Dz, g->res[1], g->res[2]);
    LNST_LAUNCH(splat_wavg_combine_box_k, dim3(lnst_blocks((int64_t)sv.ez * sv.ey * sv.ex, T)), dim3(T), 0,
                lnst_stream(stream), wmap, num, (int)nk, cells, (int)g->res[1], (int)g->res[2], sv, out);
    return lnst_status();
  }
  LNST_LAUNCH(splat_wavg_combine_k, dim3(lnst_blocks(cells, T)), dim3(T), 0, lnst_stream(stream), wmap,
              (const float*)num, (int)nk, cells, out);
  return lnst_status();
}

extern "C" int lnst_splat_wavg_coef(const float* wmap, int32_t nk, int64_t cells, float* coef, void* stream) {
  if (!wmap || !coef || nk < 1 || nk > LNST_MAX_NK || cells < 1) return LNST_EARG;
  LNST_LAUNCH(splat_wavg_coef_k, dim3(lnst_blocks(cells * nk, 256)), dim3(256), 0, lnst_stream(stream), wmap,
              cells * nk, coef);
  return lnst_status();
}

extern "C" int lnst_splat_wavg_bwd_coef(const float* p, const float* var, int64_t n, const LnstGrid* g,
                                        const float* h, int32_t nk, const float* coef, const float* g_out,
                                        float* g_var, void* stream) {
  SplatKernels ks;
  if (!grid_ok(g) || (n > 0 && (!p || !g_var)) || !coef || !g_out || n < 0 || !fill_kernels(ks, g ? g->dim : 3, h, nk))
    return LNST_EARG;
  if (g->dim != 3 || g->nsize != 1) return LNST_EARG;
  if (n == 0) return LNST_OK;
  const int64_t cells = grid_cells(g);
  const dim3 grid_(lnst_blocks(n, 256)), blk(256);
  switch (nk) {
    case 1: { auto k = splat_wavg_bwd3_k<1>; LNST_LAUNCH(k, grid_, blk, 0, lnst_stream(stream), p, var, n, *g, ks, cells, coef, g_out, g_var); break; }
    case 2: { auto k = splat_wavg_bwd3_k<2>; LNST_LAUNCH(k, grid_, blk, 0, lnst_stream(stream), p, var, n, *g, ks, cells, coef, g_out, g_var); break; }
    case 3: { auto k = splat_wavg_bwd3_k<3>; LNST_LAUNCH(k, grid_, blk, 0, lnst_stream(stream), p, var, n, *g, ks, cells, coef, g_out, g_var); break; }
    default: { auto k = splat_wavg_bwd3_k<4>; LNST_LAUNCH(k, grid_, blk, 0, lnst_stream(stream), p, var, n, *g, ks, cells, coef, g_out, g_var); break; }
  }
  return lnst_status();
}

extern "C" int lnst_splat_wavg_bwd(const float* p, const float* var, int64_t n, const LnstGrid* g,
                                   const float* h, int32_t nk, const float* wmap, const float* g_out,
                                   float* g_var, void* stream) {
  SplatKernels ks;
  if (!grid_ok(g) || (n > 0 && (!p || !g_var)) || !wmap || !g_out || n < 0 || !fill_kernels(ks, g ? g->dim : 3, h, nk))
    return LNST_EARG;
  if (n == 0) return LNST_OK;
  const int64_t cells = grid_cells(g);
  const int T = 256;
  if (g->dim == 3) {
    auto k = splat_wavg_bwd_k<3>;
    LNST_LAUNCH(k, dim3(lnst_blocks(n, T)), dim3(T), 0, lnst_stream(stream), p, var, n, *g, ks, (int)nk,
                cells, wmap, g_out, g_var);
  } else {
    auto k = splat_wavg_bwd_k<2>;
    LNST_LAUNCH(k, dim3(lnst_blocks(n, T)), dim3(T), 0, lnst_stream(stream), p, var, n, *g, ks, (int)nk,
                cells, wmap, g_out, g_var);
  }
  return lnst_status();
}
