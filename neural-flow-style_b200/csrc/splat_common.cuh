// Particle -> cell arithmetic shared by the splat kernels (csrc/splat.cu: one thread per particle, atomics; csrc/tiles_tma.cu:
// one thread per cell, gather).  Both evaluate a (particle, cell) weight with exactly these functions.
//   reference: transform.py:1233-1245 (W), :1316-1347 (cell index and offset).
#pragma once
#include "common.cuh"

#define LNST_MAX_NK 4
struct SplatKernels {
  float h[LNST_MAX_NK];
  float inv_h[LNST_MAX_NK];
  float sigma[LNST_MAX_NK];
};

template <int DIM>
struct Particle {
  bool valid;
  int idx[DIM];
  float r[DIM];    // offset from the centre of the particle's own cell (domain units)
  float dpd[DIM];  // d(domain coordinate)/d(normalised coordinate): domain, or 0 where clamped
};

template <int DIM>
__device__ __forceinline__ Particle<DIM> load_particle(const float* __restrict__ p,
                                                       const float* __restrict__ disp, int64_t i,
                                                       const LnstGrid& g) {
  Particle<DIM> o;
  o.valid = true;
  const int off = 3 - DIM;
#pragma unroll
  for (int a = 0; a < DIM; ++a) {
    float pn = p[i * DIM + a];
    if (disp != nullptr) pn = __fadd_rn(pn, disp[i * DIM + a]);
    const float dom = g.domain[off + a];
    float pd = __fmul_rn(pn, dom);
    float gs = dom;
    if (g.clip) {
      const float hi = __fadd_rn(dom, -1e-6f);
      if (pd < 0.f) { pd = 0.f; gs = 0.f; }
      if (pd > hi) { pd = hi; gs = 0.f; }
      if (pd != pd) o.valid = false;
    } else if (!(pd >= 0.f && pd < dom)) {
      o.valid = false;
    }
    const float f = floorf(pd / g.cell);
    o.idx[a] = (int)f;
    o.r[a] = __fadd_rn(pd, -__fmul_rn(f + 0.5f, g.cell));
    o.dpd[a] = gs;
  }
  return o;
}

__device__ __forceinline__ float cubic_w(float q, float sigma) {
  if (q > 1.f) return 0.f;
  const float a = 6.f * (q * q * q - q * q) + 1.f;
  const float omq = 1.f - q;
  const float b = 2.f * omq * omq * omq;
  return sigma * (q <= 0.5f ? a : b);
}
__device__ __forceinline__ float cubic_dw(float q, float sigma) {
  if (q > 1.f) return 0.f;
  const float omq = 1.f - q;
  return sigma * (q <= 0.5f ? (18.f * q * q - 12.f * q) : (-6.f * omq * omq));
}


static inline float sigma_for(int dim, float h) {
  const double pi = 3.14159265358979323846;
  return dim == 3 ? (float)(8.0 / pi / ((double)h * h * h)) : (float)(40.0 / 7.0 / pi / ((double)h * h));
}
static inline bool grid_ok(const LnstGrid* g) {
  return g && (g->dim == 2 || g->dim == 3) && g->res[1] > 0 && g->res[2] > 0 &&
         (g->dim == 2 || g->res[0] > 0) && g->cell > 0.f && g->nsize >= 0 && g->nsize <= 8;
}
static inline int64_t grid_cells(const LnstGrid* g) {
  return (int64_t)(g->dim == 3 ? g->res[0] : 1) * g->res[1] * g->res[2];
}
static inline bool fill_kernels(SplatKernels& ks, int dim, const float* h, int nk) {
  if (!h || nk < 1 || nk > LNST_MAX_NK) return false;
  for (int k = 0; k < LNST_MAX_NK; ++k) {
    ks.h[k] = k < nk ? h[k] : 1.f;
    ks.inv_h[k] = 1.f / ks.h[k];
    ks.sigma[k] = k < nk ? sigma_for(dim, h[k]) : 0.f;
    if (!(ks.h[k] > 0.f)) return false;
  }
  return true;
}

