// Ray geometry shared by the rotated ray-march kernels (csrc/render.cu: global-memory gathers; csrc/tiles_tma.cu:
// TMA-staged shared-memory slabs).  Both evaluate a sample with exactly these functions, so their images agree bit for bit.
//   reference: transform.py:611-628 (rotate), :343-433 (_interpolate3d), :152-177 (mgrid); styler_3p.py:148-158.
#pragma once
#include "common.cuh"

__device__ __forceinline__ float lin_coord(int i, float step) {
  return __fadd_rn(-1.f, __fmul_rn(step, (float)i));   // tf.linspace: start + step*i
}


struct RayGeo {
  int D, H, W, HW;
  int D2, H2, W2;        // L - 2
  float mD, mH, mW;      // L - 1
  float sD, sH, sW;      // lattice steps
  float hD, hH, hW;      // (L - 1) / 2
};
static inline RayGeo make_geo(int D, int H, int W) {
  RayGeo g;
  g.D = D; g.H = H; g.W = W; g.HW = H * W;
  g.D2 = D - 2; g.H2 = H - 2; g.W2 = W - 2;
  g.mD = (float)(D - 1); g.mH = (float)(H - 1); g.mW = (float)(W - 1);
  g.sD = 2.0f / (float)(D - 1); g.sH = 2.0f / (float)(H - 1); g.sW = 2.0f / (float)(W - 1);
  g.hD = 0.5f * g.mD; g.hH = 0.5f * g.mH; g.hW = 0.5f * g.mW;
  return g;
}

struct RayLine { float cz, cy, cx, kz, ky, kx; };   // voxel position of depth index i: c + k * i

__device__ __forceinline__ RayLine ray_line(const float* __restrict__ R, float gh, float gw, const RayGeo& g) {
  RayLine l;
  const float az = fmaf(R[1], gh, R[2] * gw), ay = fmaf(R[4], gh, R[5] * gw), ax = fmaf(R[7], gh, R[8] * gw);
  l.cz = fmaf(az - R[0], g.hD, g.hD); l.kz = R[0] * (g.sD * g.hD);
  l.cy = fmaf(ay - R[3], g.hH, g.hH); l.ky = R[3] * (g.sD * g.hH);
  l.cx = fmaf(ax - R[6], g.hW, g.hW); l.kx = R[6] * (g.sD * g.hW);
  return l;
}

// Box in float form for the interval test: a sample at position z touches voxels floor(z), floor(z)+1,
// so it matters iff lo-1 < z < hi+1; a bound on a face of the volume is open-ended because positions
// beyond the face are clamped onto it (edge replication).
struct BoxF { float lo[3], hi[3]; };
static inline BoxF make_boxf(const LnstBox* b, int D, int H, int W) {
  const float inf = __builtin_huge_valf();
  BoxF f;
  const int L[3] = {D, H, W};
  for (int a = 0; a < 3; ++a) {
    f.lo[a] = (!b || b->lo[a] <= 0) ? -inf : (float)(b->lo[a] - 1);
    f.hi[a] = (!b || b->hi[a] >= L[a] - 1) ? inf : (float)(b->hi[a] + 1);
  }
  return f;
}
__device__ __forceinline__ void axis_interval(float c, float k, float lo, float hi, float& t0, float& t1) {
  if (fabsf(k) < 1e-12f) {
    if (!(c > lo && c < hi)) { t0 = 1e30f; t1 = -1e30f; }
    return;
  }
  float a = (lo - c) / k, b = (hi - c) / k;
  if (k < 0.f) { const float t = a; a = b; b = t; }
  t0 = fmaxf(t0, a);
  t1 = fminf(t1, b);
}
// inclusive depth-index range of the samples that can touch the box (empty: lo > hi)
__device__ __forceinline__ void ray_interval(const RayLine& l, const RayGeo& g, const BoxF& bf, int& ilo, int& ihi) {
  float t0 = 0.f, t1 = g.mD;
  axis_interval(l.cz, l.kz, bf.lo[0], bf.hi[0], t0, t1);
  axis_interval(l.cy, l.ky, bf.lo[1], bf.hi[1], t0, t1);
  axis_interval(l.cx, l.kx, bf.lo[2], bf.hi[2], t0, t1);
  if (!(t0 <= t1 + 2.f)) { ilo = 1; ihi = 0; return; }
  ilo = max(0, (int)ceilf(t0) - 1);
  ihi = min(g.D - 1, (int)floorf(t1) + 1);
}

// Occupancy bricks (4^3 voxels, one byte each; set where any voxel within reach of the brick is active --
// the caller dilates by two voxels and one brick, see Styler._workspace): shrink [ilo, ihi] from both
// ends while the ray is in empty bricks, testing every `stride`-th sample (stride * max|k| <= 4 voxels, so
// no sample between two tested ones can touch an active voxel the tests missed).
#define LNST_BRICK 4
struct Bricks { const unsigned char* occ; int by, bx; };
__device__ __forceinline__ bool brick_hit(float cz, float cy, float cx, float kz, float ky, float kx, float fi,
                                          float mD, float mH, float mW, const unsigned char* __restrict__ occ,
                                          int by, int bx) {
  const int z = (int)fminf(fmaxf(fmaf(kz, fi, cz), 0.f), mD);
  const int y = (int)fminf(fmaxf(fmaf(ky, fi, cy), 0.f), mH);
  const int x = (int)fminf(fmaxf(fmaf(kx, fi, cx), 0.f), mW);
  return occ[((z / LNST_BRICK) * by + (y / LNST_BRICK)) * bx + (x / LNST_BRICK)] != 0;
}
__device__ __forceinline__ void refine_interval(const RayLine l, const RayGeo g, const unsigned char* occ, int by, int bx,
                                                int& ilo, int& ihi) {
  if (occ == nullptr || ilo > ihi) return;
  const float km = fmaxf(fmaxf(fabsf(l.kz), fabsf(l.ky)), fmaxf(fabsf(l.kx), 1e-6f));
  const int stride = max(1, (int)(4.f / km));
  const int lo0 = ilo, hi0 = ihi;
  int a = ilo, b = ihi;
  while (a <= b && !brick_hit(l.cz, l.cy, l.cx, l.kz, l.ky, l.kx, (float)a, g.mD, g.mH, g.mW, occ, by, bx)) a += stride;
  if (a > b) { ilo = 1; ihi = 0; return; }
  while (b > a && !brick_hit(l.cz, l.cy, l.cx, l.kz, l.ky, l.kx, (float)b, g.mD, g.mH, g.mW, occ, by, bx)) b -= stride;
  ilo = max(lo0, a - stride + 1);
  ihi = min(hi0, b + stride - 1);
}

struct Cell { int idx; float fz, fy, fx; };

__device__ __forceinline__ Cell locate(const RayLine& l, float fi, const RayGeo& g) {
  const float z = fminf(fmaxf(fmaf(l.kz, fi, l.cz), 0.f), g.mD);
  const float y = fminf(fmaxf(fmaf(l.ky, fi, l.cy), 0.f), g.mH);
  const float x = fminf(fmaxf(fmaf(l.kx, fi, l.cx), 0.f), g.mW);
  const int z0 = min((int)z, g.D2), y0 = min((int)y, g.H2), x0 = min((int)x, g.W2);   // z,y,x >= 0: trunc = floor
  Cell c;
  c.fz = z - (float)z0; c.fy = y - (float)y0; c.fx = x - (float)x0;
  c.idx = (z0 * g.H + y0) * g.W + x0;
  return c;
}

__device__ __forceinline__ float sample_cell(const float* __restrict__ vol, const Cell& c, const RayGeo& g) {
  const float* p = vol + c.idx;
  const float v000 = p[0], v001 = p[1], v010 = p[g.W], v011 = p[g.W + 1];
  const float* q = p + g.HW;
  const float v100 = q[0], v101 = q[1], v110 = q[g.W], v111 = q[g.W + 1];
  const float a00 = fmaf(c.fx, v001 - v000, v000), a01 = fmaf(c.fx, v011 - v010, v010);
  const float a10 = fmaf(c.fx, v101 - v100, v100), a11 = fmaf(c.fx, v111 - v110, v110);
  const float b0 = fmaf(c.fy, a01 - a00, a00), b1 = fmaf(c.fy, a11 - a10, a10);
  return fmaf(c.fz, b1 - b0, b0);
}

// The four voxels of one z-plane of a cell: [y0x0, y0x1, y1x0, y1x1].  A ray whose anchor advances by exactly
// one voxel in depth between two consecutive samples (the usual case for the reference's small view angles,
// |R[0]| ~ 1) sees the far plane of one sample as the near plane of the next: the values (forward) and the
// gradient contributions (backward) of that plane are carried in registers instead of being re-read /
// scattered twice.  Same arithmetic per sample as sample_cell, so images stay bit-identical.
struct Plane4 { float a, b, c, d; };
__device__ __forceinline__ Plane4 load_plane(const float* __restrict__ p, int W) {
  Plane4 v;
  v.a = p[0]; v.b = p[1]; v.c = p[W]; v.d = p[W + 1];
  return v;
}
__device__ __forceinline__ float lerp_planes(const Plane4& lo, const Plane4& hi, const Cell& c) {
  const float a00 = fmaf(c.fx, lo.b - lo.a, lo.a), a01 = fmaf(c.fx, lo.d - lo.c, lo.c);
  const float a10 = fmaf(c.fx, hi.b - hi.a, hi.a), a11 = fmaf(c.fx, hi.d - hi.c, hi.c);
  const float b0 = fmaf(c.fy, a01 - a00, a00), b1 = fmaf(c.fy, a11 - a10, a10);
  return fmaf(c.fz, b1 - b0, b0);
}
#define LNST_NO_CELL (-0x40000000)

__device__ __forceinline__ float fast_exp2(float x) {
#ifdef LNST_CPU_EMU
  return exp2f(x);
#else
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#endif
}

