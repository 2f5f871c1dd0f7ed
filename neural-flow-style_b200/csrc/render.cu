// Differentiable rotation + emission/absorption volume rendering, and the image glue between
// the render and the loss network.
//   reference: transform.py:611-628 (rotate), :343-433 (_interpolate3d), :152-177 (mgrid);
//              styler_3p.py:148-158 (render, /max); styler_base.py:33-45 (resize, x255, RGB);
//              vgg.py:50-53 (mean subtraction).
// The rotated volume of the reference ([n_views,D,H,W], 32 MB per view at 200^3) is never
// materialised: each ray samples the source volume while it marches.
#include "common.cuh"
#include "render_common.cuh"

#define RM_UNROLL 4
#ifndef RMB_UNROLL
#define RMB_UNROLL 1   // backward: samples in flight per thread; 1 + __launch_bounds__(128, 8) = 64 registers, 8 CTAs per SM (2 needed 80: 6 CTAs)
#endif
#ifndef RMB_MINB
#define RMB_MINB 8
#endif

struct VolDims {
  int D, H, W;
  float sD, sH, sW;   // linspace steps 2/(L-1) (0 when L == 1), transform.py:175
};
static inline VolDims make_dims(int D, int H, int W) {
  VolDims v;
  v.D = D; v.H = H; v.W = W;
  v.sD = D > 1 ? 2.0f / (float)(D - 1) : 0.f;
  v.sH = H > 1 ? 2.0f / (float)(H - 1) : 0.f;
  v.sW = W > 1 ? 2.0f / (float)(W - 1) : 0.f;
  return v;
}

struct Corner8 {
  int z0, z1, y0, y1, x0, x1;
  float fz, fy, fx;
};

// rotated sample position of lattice point (gd,gh,gw) -> clamped corner indices + fractions
__device__ __forceinline__ Corner8 rotate_sample(const float* __restrict__ R, float gd, float gh, float gw,
                                                 const VolDims& v) {
  const float u0 = fmaf(R[2], gw, fmaf(R[1], gh, R[0] * gd));
  const float u1 = fmaf(R[5], gw, fmaf(R[4], gh, R[3] * gd));
  const float u2 = fmaf(R[8], gw, fmaf(R[7], gh, R[6] * gd));
  const float z = (u0 + 1.f) * ((float)v.D - 1.f) * 0.5f;
  const float y = (u1 + 1.f) * ((float)v.H - 1.f) * 0.5f;
  const float x = (u2 + 1.f) * ((float)v.W - 1.f) * 0.5f;
  Corner8 c;
  const int zf = (int)floorf(z), yf = (int)floorf(y), xf = (int)floorf(x);
  c.z0 = min(max(zf, 0), v.D - 1); c.z1 = min(max(zf + 1, 0), v.D - 1);
  c.y0 = min(max(yf, 0), v.H - 1); c.y1 = min(max(yf + 1, 0), v.H - 1);
  c.x0 = min(max(xf, 0), v.W - 1); c.x1 = min(max(xf + 1, 0), v.W - 1);
  c.fz = z - (float)c.z0; c.fy = y - (float)c.y0; c.fx = x - (float)c.x0;
  return c;
}

__device__ __forceinline__ float sample8(const float* __restrict__ vol, const Corner8& c, const VolDims& v) {
  const int64_t HW = (int64_t)v.H * v.W;
  const int64_t b00 = c.z0 * HW + (int64_t)c.y0 * v.W, b01 = c.z0 * HW + (int64_t)c.y1 * v.W;
  const int64_t b10 = c.z1 * HW + (int64_t)c.y0 * v.W, b11 = c.z1 * HW + (int64_t)c.y1 * v.W;
  const float gz = 1.f - c.fz, gy = 1.f - c.fy, gx = 1.f - c.fx;
  float o = gz * gy * gx * vol[b00 + c.x0];
  o += gz * gy * c.fx * vol[b00 + c.x1];
  o += gz * c.fy * gx * vol[b01 + c.x0];
  o += gz * c.fy * c.fx * vol[b01 + c.x1];
  o += c.fz * gy * gx * vol[b10 + c.x0];
  o += c.fz * gy * c.fx * vol[b10 + c.x1];
  o += c.fz * c.fy * gx * vol[b11 + c.x0];
  o += c.fz * c.fy * c.fx * vol[b11 + c.x1];
  return o;
}

__device__ __forceinline__ void scatter8(float* __restrict__ gv, const Corner8& c, const VolDims& v, float g) {
  const int64_t HW = (int64_t)v.H * v.W;
  const int64_t b00 = c.z0 * HW + (int64_t)c.y0 * v.W, b01 = c.z0 * HW + (int64_t)c.y1 * v.W;
  const int64_t b10 = c.z1 * HW + (int64_t)c.y0 * v.W, b11 = c.z1 * HW + (int64_t)c.y1 * v.W;
  const float gz = 1.f - c.fz, gy = 1.f - c.fy, gx = 1.f - c.fx;
  atomicAdd(gv + b00 + c.x0, g * (gz * gy * gx));
  atomicAdd(gv + b00 + c.x1, g * (gz * gy * c.fx));
  atomicAdd(gv + b01 + c.x0, g * (gz * c.fy * gx));
  atomicAdd(gv + b01 + c.x1, g * (gz * c.fy * c.fx));
  atomicAdd(gv + b10 + c.x0, g * (c.fz * gy * gx));
  atomicAdd(gv + b10 + c.x1, g * (c.fz * gy * c.fx));
  atomicAdd(gv + b11 + c.x0, g * (c.fz * c.fy * gx));
  atomicAdd(gv + b11 + c.x1, g * (c.fz * c.fy * c.fx));
}

__global__ void rotate_fwd_k(const float* __restrict__ vol, const float* __restrict__ rot, VolDims v,
                             float* __restrict__ out) {
  const int64_t V = (int64_t)v.D * v.H * v.W;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= V) return;
  const int view = blockIdx.y;
  const int w = (int)(t % v.W), h = (int)((t / v.W) % v.H), i = (int)(t / ((int64_t)v.W * v.H));
  const Corner8 c = rotate_sample(rot + 9 * view, lin_coord(i, v.sD), lin_coord(h, v.sH), lin_coord(w, v.sW), v);
  out[view * V + t] = sample8(vol, c, v);
}

// gradient of rotate() w.r.t. the volume: the 8-corner scatter-add of every rotated sample's cotangent
// (TF: gather gradient = unsorted_segment_sum, transform.py:385-428); g_vol accumulates over all views
__global__ void rotate_bwd_k(const float* __restrict__ g_out, const float* __restrict__ rot, VolDims v,
                             float* __restrict__ g_vol) {
  const int64_t V = (int64_t)v.D * v.H * v.W;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= V) return;
  const int view = blockIdx.y;
  const float g = g_out[view * V + t];
  if (g == 0.f) return;
  const int w = (int)(t % v.W), h = (int)((t / v.W) % v.H), i = (int)(t / ((int64_t)v.W * v.H));
  const Corner8 c = rotate_sample(rot + 9 * view, lin_coord(i, v.sD), lin_coord(h, v.sH), lin_coord(w, v.sW), v);
  scatter8(g_vol, c, v, g);
}

// one thread per pixel column (view, h, w); marches from the camera side (high D) down
__global__ void raymarch_fwd_k(const float* __restrict__ vol, const float* __restrict__ rot, VolDims v,
                               SubVol sv, float tau, int liquid, float* __restrict__ img,
                               float* __restrict__ stot, float* __restrict__ stats) {
  const int P = v.H * v.W;
  const int pix0 = blockIdx.x * blockDim.x + threadIdx.x;
  const bool act = pix0 < P;
  if (!act && stats == nullptr) return;
  const int pix = act ? pix0 : P - 1;            // threads past the image only take part in the maximum
  const int view = blockIdx.y;
  const int h = pix / v.W, w = pix % v.W;
  const float* R = rot ? rot + 9 * view : nullptr;
  const float gh = lin_coord(h, v.sH), gw = lin_coord(w, v.sW);
  float S = 0.f, I = 0.f;
  // active sub-volume (unrotated rays only): the density is zero outside it
  int i_lo = 0, i_hi = v.D - 1;
  if (!R) {
    i_lo = sv.oz; i_hi = sv.oz + sv.ez - 1;
    if (h < sv.oy || h >= sv.oy + sv.ey || w < sv.ox || w >= sv.ox + sv.ex) i_hi = i_lo - 1;
  }
  if (!act) i_hi = i_lo - 1;
  // RM_UNROLL depth steps are sampled before any of them is consumed, so their (independent)
  // loads are in flight together; only the running transmittance is sequential.
  for (int i0 = i_hi; i0 >= i_lo; i0 -= RM_UNROLL) {
    float d[RM_UNROLL];
#pragma unroll
    for (int u = 0; u < RM_UNROLL; ++u) {
      const int i = i0 - u;
      d[u] = 0.f;
      if (i >= i_lo) {
        if (R) {
          const Corner8 c = rotate_sample(R, lin_coord(i, v.sD), gh, gw, v);
          d[u] = sample8(vol, c, v);
        } else {
          d[u] = vol[(int64_t)i * P + pix];
        }
      }
    }
#pragma unroll
    for (int u = 0; u < RM_UNROLL; ++u) {
      if (i0 - u >= i_lo) {
        S += d[u];                              // inclusive reverse cumsum, styler_3p.py:155
        if (!liquid) I += d[u] * expf(-S * tau);
      }
    }
  }
  if (liquid) I = 1.f - expf(-S * tau);         // styler_3p.py:150-152
  if (act) {
    img[(int64_t)view * P + pix] = I;
    stot[(int64_t)view * P + pix] = S;
  }
  if (stats != nullptr) {                       // the view's maximum (image_max_k), one atomic per warp
    const float m = lnst_warp_max(act ? I : 0.f);
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(stats + 2 * view), __float_as_int(m));
  }
}

// d I / d d_k = T_k - tau * sum_{i<=k} d_i T_i  (smoke);  tau * exp(-tau * S_total) (liquid)
__global__ void raymarch_bwd_k(const float* __restrict__ vol, const float* __restrict__ rot, VolDims v,
                               SubVol sv, float tau, int liquid, const float* __restrict__ stot,
                               const float* __restrict__ g_img, float* __restrict__ g_vol, int use_atomic,
                               const float* __restrict__ nimg, const float* __restrict__ nstats,
                               const float* __restrict__ ndots) {
  const int P = v.H * v.W;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= P) return;
  const int view = blockIdx.y;
  const int h = pix / v.W, w = pix % v.W;
  const float* R = rot ? rot + 9 * view : nullptr;
  float gI = g_img[(int64_t)view * P + pix];
  if (nstats != nullptr) {                       // g_img is d loss / d (img / max): normalize_bwd_k, inline
    const float m = nstats[2 * view], ties = nstats[2 * view + 1];
    gI = gI / m;
    if (nimg[(int64_t)view * P + pix] == m) gI -= ndots[view] / (m * m) / ties;
  }
  const float St = stot[(int64_t)view * P + pix];
  if (gI == 0.f) return;
  const float gh = lin_coord(h, v.sH), gw = lin_coord(w, v.sW);
  const float gl = liquid ? gI * tau * expf(-St * tau) : 0.f;
  float below = 0.f, Pk = 0.f;
  int i_lo = 0, i_end = v.D;
  if (!R) {
    i_lo = sv.oz; i_end = sv.oz + sv.ez;
    if (h < sv.oy || h >= sv.oy + sv.ey || w < sv.ox || w >= sv.ox + sv.ex) return;
  }
  for (int i0 = i_lo; i0 < i_end; i0 += RM_UNROLL) {
    Corner8 c[RM_UNROLL];
    float d[RM_UNROLL];
#pragma unroll
    for (int u = 0; u < RM_UNROLL; ++u) {
      const int i = i0 + u;
      d[u] = 0.f;
      if (i < i_end) {
        if (R) {
          c[u] = rotate_sample(R, lin_coord(i, v.sD), gh, gw, v);
          if (!liquid) d[u] = sample8(vol, c[u], v);
        } else if (!liquid) {
          d[u] = vol[(int64_t)i * P + pix];
        }
      }
    }
#pragma unroll
    for (int u = 0; u < RM_UNROLL; ++u) {
      const int i = i0 + u;
      if (i < i_end) {
        float g;
        if (liquid) {
          g = gl;
        } else {
          const float T = expf(-(St - below) * tau);
          Pk += d[u] * T;
          below += d[u];
          g = gI * (T - tau * Pk);
        }
        if (R) scatter8(g_vol, c[u], v, g);
        else if (use_atomic) atomicAdd(g_vol + (int64_t)i * P + pix, g);
        else g_vol[(int64_t)i * P + pix] += g;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// Fast path of the rotated ray-march (every axis >= 2 voxels).
//  * the sample position is affine in the depth index: pos(i) = c + k*i per axis (one FFMA each);
//  * _interpolate3d's "clamp floor and floor+1 separately, weights from the unclamped fraction"
//    (transform.py:385-428) equals clamping the coordinate to [0, L-1] and interpolating in the
//    cell [min(floor, L-2), +1]: both corners coincide outside the volume, so the weights sum to
//    one on the edge voxel either way.  The second corner is then always idx + 1 / + W / + H*W:
//    one 32-bit anchor per sample and immediate-offset loads;
//  * backward: lane l's x1 corners are lane l+1's x0 corners whenever their anchors differ by
//    one voxel (the common case for the small view angles of the reference, config.py:63-70), so
//    they travel by warp shuffle and one red.global covers both: ~4.7 instead of 8 per sample.
// ---------------------------------------------------------------------------------------
// [i_lo, i_hi] of every ray of every view: the box slab test, then the brick refinement.  Depends only on
// the view matrices, the box and the bricks -- not on the density -- so it runs once per view set.
__global__ void __launch_bounds__(128) ray_intervals_k(const float* __restrict__ rot, RayGeo g, BoxF bf, Bricks br,
                                                        int2* __restrict__ iv) {
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= g.HW) return;
  const int view = blockIdx.y;
  const int h = pix / g.W, w = pix - h * g.W;
  const RayLine l = ray_line(rot + 9 * view, lin_coord(h, g.sH), lin_coord(w, g.sW), g);
  int lo, hi;
  ray_interval(l, g, bf, lo, hi);
  refine_interval(l, g, br.occ, br.by, br.bx, lo, hi);
  iv[(int64_t)view * g.HW + pix] = make_int2(lo, hi);
}

// Exact live interval of every ray: first and last sample whose trilinear footprint touches a voxel where the smoothed
// density can be non-zero.  `touch` [D,H,W] (one byte per voxel) is set for an ANCHOR voxel v when any of the 8 voxels
// {v, v+1}^3 is active, so one byte per sample decides.  Costs about one forward march, so it is meant for view sets
// that stay fixed over the iterations (sample_type 'uniform'); the brick test above is the cheap conservative version.
__global__ void __launch_bounds__(128) ray_intervals_exact_k(const float* __restrict__ rot, RayGeo g, BoxF bf,
                                                              const unsigned char* __restrict__ touch,
                                                              int2* __restrict__ iv) {
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= g.HW) return;
  const int view = blockIdx.y;
  const int h = pix / g.W, w = pix - h * g.W;
  const RayLine l = ray_line(rot + 9 * view, lin_coord(h, g.sH), lin_coord(w, g.sW), g);
  int lo, hi;
  ray_interval(l, g, bf, lo, hi);
  while (lo <= hi && touch[locate(l, (float)lo, g).idx] == 0) ++lo;
  while (hi > lo && touch[locate(l, (float)hi, g).idx] == 0) --hi;
  if (lo > hi) { lo = 1; hi = 0; }
  iv[(int64_t)view * g.HW + pix] = make_int2(lo, hi);
}

__global__ void __launch_bounds__(128) raymarch_rot_fwd_k(const float* __restrict__ vol, const float* __restrict__ rot,
                                                           RayGeo g, BoxF bf, const int2* __restrict__ iv,
                                                           float ntl2, int liquid, float* __restrict__ img,
                                                           float* __restrict__ stot, float* __restrict__ stats) {
  const int pix0 = blockIdx.x * blockDim.x + threadIdx.x;
  const bool act = pix0 < g.HW;
  if (!act && stats == nullptr) return;
  const int pix = act ? pix0 : g.HW - 1;               // threads past the image only take part in the maximum
  const int view = blockIdx.y;
  const int h = pix / g.W, w = pix - h * g.W;
  const RayLine l = ray_line(rot + 9 * view, lin_coord(h, g.sH), lin_coord(w, g.sW), g);
  float S = 0.f, I = 0.f;
  int i_lo, i0;
  if (iv) {                                            // precomputed by ray_intervals_k (box + bricks)
    const int2 r = iv[(int64_t)view * g.HW + pix];
    i_lo = r.x; i0 = r.y;
  } else {
    ray_interval(l, g, bf, i_lo, i0);                  // the density is zero outside [i_lo, i0]
  }
  if (!act) { i_lo = 1; i0 = 0; }
  // marching towards the eye (descending i): sample i's near plane is sample i-1's far plane when the
  // anchor steps back by exactly one voxel in depth
  Plane4 near_prev;
  near_prev.a = near_prev.b = near_prev.c = near_prev.d = 0.f;
  int idx_prev = LNST_NO_CELL;
  for (; i0 >= i_lo + RM_UNROLL - 1; i0 -= RM_UNROLL) {   // full groups: no per-sample bounds checks
    Cell c[RM_UNROLL];
    Plane4 lo[RM_UNROLL], hi[RM_UNROLL];
    bool cont[RM_UNROLL];
#pragma unroll
    for (int u = 0; u < RM_UNROLL; ++u) c[u] = locate(l, (float)(i0 - u), g);
#pragma unroll
    for (int u = 0; u < RM_UNROLL; ++u) {
      const float* p = vol + c[u].idx;
      lo[u] = load_plane(p, g.W);
      cont[u] = c[u].idx + g.HW == (u == 0 ? idx_prev : c[u - 1].idx);
      if (!cont[u]) hi[u] = load_plane(p + g.HW, g.W);
    }
#pragma unroll
    for (int u = 0; u < RM_UNROLL; ++u) {
      if (cont[u]) hi[u] = (u == 0) ? near_prev : lo[u - 1];
      const float d = lerp_planes(lo[u], hi[u], c[u]);
      S += d;                                           // inclusive reverse cumsum, styler_3p.py:155
      I = fmaf(d, fast_exp2(S * ntl2), I);
    }
    near_prev = lo[RM_UNROLL - 1];
    idx_prev = c[RM_UNROLL - 1].idx;
  }
  for (; i0 >= i_lo; --i0) {
    const Cell c = locate(l, (float)i0, g);
    const float* p = vol + c.idx;
    const Plane4 lo = load_plane(p, g.W);
    const Plane4 hi = (c.idx + g.HW == idx_prev) ? near_prev : load_plane(p + g.HW, g.W);
    const float d = lerp_planes(lo, hi, c);
    S += d;
    I = fmaf(d, fast_exp2(S * ntl2), I);
    near_prev = lo;
    idx_prev = c.idx;
  }
  if (liquid) I = 1.f - fast_exp2(S * ntl2);            // styler_3p.py:150-152
  if (act) {
    img[(int64_t)view * g.HW + pix] = I;
    stot[(int64_t)view * g.HW + pix] = S;
  }
  if (stats != nullptr) {                               // the view's maximum (image_max_k), one atomic per warp
    const float m = lnst_warp_max(act ? I : 0.f);
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(stats + 2 * view), __float_as_int(m));
  }
}

// d I / d d_k = T_k - tau * sum_{i<=k} d_i T_i  (smoke);  tau * exp(-tau * S_total) (liquid)
// PATCH: a warp is a 4 x 8 pixel patch instead of 32 pixels of one row (a block = 4 rows x 32 columns).  The warp walks
// the union of its rays' depth intervals in lockstep, and the rays of a compact patch enter and leave the density
// together: 89 % of the lane-slots are live at C3 against 75 % for rows (tools/dev/iv_probe.py).  It also lets the y1 row
// of a sample merge into the lane one row down (+8), like the x1 column merges into lane + 1.
template <bool MERGE, bool PATCH>
__global__ void __launch_bounds__(128, RMB_MINB) raymarch_rot_bwd_k(const float* __restrict__ vol, const float* __restrict__ rot,
                                                           RayGeo g, BoxF bf, const int2* __restrict__ iv,
                                                           float tau, float ntl2, int liquid,
                                                           const float* __restrict__ stot,
                                                           const float* __restrict__ g_img, float* __restrict__ g_vol,
                                                           int tiles_w, const float* __restrict__ nimg,
                                                           const float* __restrict__ nstats,
                                                           const float* __restrict__ ndots) {
  const int view = blockIdx.y;
  const int lane = threadIdx.x & 31;
  int pix;
  bool active;
  if (PATCH) {
    const int by = blockIdx.x / tiles_w, bx = blockIdx.x - by * tiles_w;
    const int ph = by * 4 + (lane >> 3), pw = (bx * 4 + (threadIdx.x >> 5)) * 8 + (lane & 7);
    active = ph < g.H && pw < g.W;
    pix = active ? ph * g.W + pw : 0;
  } else {
    pix = blockIdx.x * blockDim.x + threadIdx.x;
    active = pix < g.HW;
  }
  float gI = 0.f, St = 0.f;
  if (active) {
    gI = g_img[(int64_t)view * g.HW + pix];
    if (nstats != nullptr) {                             // g_img is d loss / d (img / max): normalize_bwd_k, inline
      const float m = nstats[2 * view], ties = nstats[2 * view + 1];
      const float x = nimg[(int64_t)view * g.HW + pix];
      gI = gI / m;
      if (x == m) gI -= ndots[view] / (m * m) / ties;
    }
    St = stot[(int64_t)view * g.HW + pix];
    active = gI != 0.f;
  }
  if (!MERGE && !active) return;
  if (MERGE && !__any_sync(0xffffffffu, active)) return;
  const int pc = active ? pix : 0;
  const int h = pc / g.W, w = pc - h * g.W;
  const RayLine l = ray_line(rot + 9 * view, lin_coord(h, g.sH), lin_coord(w, g.sW), g);
  const float gl = liquid ? gI * tau * fast_exp2(St * ntl2) : 0.f;
  float below = 0.f, Pk = 0.f;
  // samples below the interval have zero density (below = Pk = 0 there) and, like those above it,
  // no footprint inside the box
  int i_lo, i_hi;
  if (iv) {
    const int2 r = iv[(int64_t)view * g.HW + pc];
    i_lo = r.x; i_hi = r.y;
  } else {
    ray_interval(l, g, bf, i_lo, i_hi);
  }
  if (!active || i_lo > i_hi) { i_lo = 0x7fffffff; i_hi = -1; }
  int w_lo = i_lo, w_hi = i_hi;                        // warp-uniform loop bounds (the merge shuffles)
  if (MERGE) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      w_lo = min(w_lo, __shfl_xor_sync(0xffffffffu, w_lo, o));
      w_hi = max(w_hi, __shfl_xor_sync(0xffffffffu, w_hi, o));
    }
  }
  // marching away from the eye (ascending i).  carry: the ray advances in +z by about one voxel per sample,
  // so a sample's far plane is usually the next sample's near plane -- its 4 voxel values are reused and its
  // 4 gradient contributions wait in registers (pend) to be merged into the next sample's near plane: one
  // red.global per voxel and sample pair instead of two.  Rays running the other way scatter directly.
  const bool carry = rot[9 * view] > 0.5f;             // warp-uniform (one view per blockIdx.y)
  Plane4 far_prev, pend;
  far_prev.a = far_prev.b = far_prev.c = far_prev.d = 0.f;
  pend = far_prev;
  int idx_prev = LNST_NO_CELL;                         // anchor of the previous live, sampled cell
  int pend_idx = -1;                                   // voxel index of the plane waiting in pend
  for (int i0 = w_lo; i0 <= w_hi; i0 += RMB_UNROLL) {
    Cell c[RMB_UNROLL];
    float d[RMB_UNROLL];
    Plane4 lo[RMB_UNROLL], hi[RMB_UNROLL];
    bool cont[RMB_UNROLL], lv[RMB_UNROLL];
#pragma unroll
    for (int u = 0; u < RMB_UNROLL; ++u) {
      const int i = min(i0 + u, g.D - 1);              // tail: a repeated sample, masked below
      c[u] = locate(l, (float)i, g);
      lv[u] = i0 + u >= i_lo && i0 + u <= i_hi;
    }
#pragma unroll
    for (int u = 0; u < RMB_UNROLL; ++u) {
      cont[u] = false;
      if (!liquid && lv[u]) {
        const float* p = vol + c[u].idx;
        const int ip = (u == 0) ? idx_prev : (lv[u - 1] ? c[u - 1].idx : LNST_NO_CELL);
        cont[u] = c[u].idx == ip + g.HW;
        hi[u] = load_plane(p + g.HW, g.W);
        if (!cont[u]) lo[u] = load_plane(p, g.W);
      }
    }
#pragma unroll
    for (int u = 0; u < RMB_UNROLL; ++u) {
      d[u] = 0.f;
      if (!liquid && lv[u]) {
        if (cont[u]) lo[u] = (u == 0) ? far_prev : hi[u - 1];
        d[u] = lerp_planes(lo[u], hi[u], c[u]);
      }
    }
    if (!liquid) {
      far_prev = hi[RMB_UNROLL - 1];
      idx_prev = lv[RMB_UNROLL - 1] ? c[RMB_UNROLL - 1].idx : LNST_NO_CELL;
    }
#pragma unroll
    for (int u = 0; u < RMB_UNROLL; ++u) {
      const bool live = lv[u];
      float gk;
      if (liquid) {
        gk = gl;
      } else {
        const float T = fast_exp2((St - below) * ntl2);
        Pk = fmaf(d[u], T, Pk);
        below += d[u];
        gk = gI * fmaf(-tau, Pk, T);
      }
      if (!live) gk = 0.f;
      const float fz = c[u].fz, fy = c[u].fy, fx = c[u].fx;
      const float g1 = gk * fz, g0 = gk - g1;
      const float g01 = g0 * fy, g00 = g0 - g01, g11 = g1 * fy, g10 = g1 - g11;
      float c001 = g00 * fx, c011 = g01 * fx, c101 = g10 * fx, c111 = g11 * fx;
      float c000 = g00 - c001, c010 = g01 - c011, c100 = g10 - c101, c110 = g11 - c111;
      float* p = g_vol + c[u].idx;
      float* q = p + g.HW;
      if (carry) {
        // the plane waiting in pend: merge it into this sample's near plane, or write it out
        if (pend_idx >= 0) {
          if (live && pend_idx == c[u].idx) {
            c000 += pend.a; c001 += pend.b; c010 += pend.c; c011 += pend.d;
          } else {
            float* f = g_vol + pend_idx;
            atomicAdd(f, pend.a); atomicAdd(f + 1, pend.b); atomicAdd(f + g.W, pend.c); atomicAdd(f + g.W + 1, pend.d);
          }
          pend_idx = -1;
        }
        bool give = false;
        if (MERGE) {                                   // near plane only: the far plane is merged when it is written
          const int my = live ? c[u].idx : -2;
          const int nb = __shfl_down_sync(0xffffffffu, my, 1);
          give = lane < 31 && my >= 0 && nb == my + 1;
          const float r00 = __shfl_up_sync(0xffffffffu, give ? c001 : 0.f, 1);
          const float r01 = __shfl_up_sync(0xffffffffu, give ? c011 : 0.f, 1);
          if (lane > 0) { c000 += r00; c010 += r01; }
        }
        bool give_y = false;
        if (MERGE && PATCH) {                          // (y1, x0) goes to the lane one patch row down when its anchor is one voxel down
          const int my = live ? c[u].idx : -0x40000000;
          const int nb8 = __shfl_down_sync(0xffffffffu, my, 8);
          give_y = lane < 24 && live && nb8 == my + g.W;
          const float ry = __shfl_up_sync(0xffffffffu, give_y ? c010 : 0.f, 8);
          if (lane >= 8) c000 += ry;
        }
        if (live) {
          atomicAdd(p, c000);
          if (!give_y) atomicAdd(p + g.W, c010);
          if (!give) { atomicAdd(p + 1, c001); atomicAdd(p + g.W + 1, c011); }
          pend.a = c100; pend.b = c101; pend.c = c110; pend.d = c111;
          pend_idx = c[u].idx + g.HW;
        }
      } else if (MERGE) {
        const int my = live ? c[u].idx : -2;
        const int nb = __shfl_down_sync(0xffffffffu, my, 1);
        const bool give = lane < 31 && my >= 0 && nb == my + 1;
        const float r00 = __shfl_up_sync(0xffffffffu, give ? c001 : 0.f, 1);
        const float r01 = __shfl_up_sync(0xffffffffu, give ? c011 : 0.f, 1);
        const float r10 = __shfl_up_sync(0xffffffffu, give ? c101 : 0.f, 1);
        const float r11 = __shfl_up_sync(0xffffffffu, give ? c111 : 0.f, 1);
        if (lane > 0) { c000 += r00; c010 += r01; c100 += r10; c110 += r11; }
        if (live) {
          atomicAdd(p, c000); atomicAdd(p + g.W, c010); atomicAdd(q, c100); atomicAdd(q + g.W, c110);
          if (!give) {
            atomicAdd(p + 1, c001); atomicAdd(p + g.W + 1, c011); atomicAdd(q + 1, c101); atomicAdd(q + g.W + 1, c111);
          }
        }
      } else if (live) {
        atomicAdd(p, c000); atomicAdd(p + 1, c001); atomicAdd(p + g.W, c010); atomicAdd(p + g.W + 1, c011);
        atomicAdd(q, c100); atomicAdd(q + 1, c101); atomicAdd(q + g.W, c110); atomicAdd(q + g.W + 1, c111);
      }
    }
  }
  if (pend_idx >= 0) {
    float* f = g_vol + pend_idx;
    atomicAdd(f, pend.a); atomicAdd(f + 1, pend.b); atomicAdd(f + g.W, pend.c); atomicAdd(f + g.W + 1, pend.d);
  }
}

// ---------------------------------------------------------------------------------------
// image glue
// ---------------------------------------------------------------------------------------
// stats[2v] = max (images are >= 0: int compare on the bit pattern is order preserving)
__global__ void image_max_k(const float* __restrict__ img, int64_t n_pix, float* __restrict__ stats) {
  const int view = blockIdx.y;
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pix; i += (int64_t)gridDim.x * blockDim.x)
    m = fmaxf(m, img[view * n_pix + i]);
  m = lnst_warp_max(m);
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(stats + 2 * view), __float_as_int(m));
}
__global__ void image_ties_k(const float* __restrict__ img, int64_t n_pix, float* __restrict__ stats) {
  const int view = blockIdx.y;
  const float m = stats[2 * view];
  float c = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pix; i += (int64_t)gridDim.x * blockDim.x)
    c += (img[view * n_pix + i] == m) ? 1.f : 0.f;
  c = lnst_warp_sum(c);
  if ((threadIdx.x & 31) == 0 && c != 0.f) atomicAdd(stats + 2 * view + 1, c);
}
__global__ void normalize_fwd_k(const float* __restrict__ img, const float* __restrict__ stats, int64_t n_pix,
                                float* __restrict__ gray) {
  const int view = blockIdx.y;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pix) return;
  gray[view * n_pix + i] = img[view * n_pix + i] / stats[2 * view];
}
// normalize_fwd_k + image_ties_k in one pass over the image (stats[2v] = max already final, stats[2v+1] zero on entry)
__global__ void normalize_ties_fwd_k(const float* __restrict__ img, float* __restrict__ stats, int64_t n_pix,
                                     float* __restrict__ gray) {
  const int view = blockIdx.y;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const float m = stats[2 * view];
  float c = 0.f;
  if (i < n_pix) {
    const float x = img[view * n_pix + i];
    gray[view * n_pix + i] = x / m;
    c = (x == m) ? 1.f : 0.f;
  }
  c = lnst_warp_sum(c);
  if ((threadIdx.x & 31) == 0 && c != 0.f) atomicAdd(stats + 2 * view + 1, c);
}
__global__ void dot_k(const float* __restrict__ a, const float* __restrict__ b, int64_t n_pix,
                      float* __restrict__ dots) {
  const int view = blockIdx.y;
  float s = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pix; i += (int64_t)gridDim.x * blockDim.x)
    s += a[view * n_pix + i] * b[view * n_pix + i];
  s = lnst_warp_sum(s);
  if ((threadIdx.x & 31) == 0) atomicAdd(dots + view, s);
}
// y = x/m: dL/dx_p = g_p/m - [x_p == m] * (sum_q g_q x_q) / m^2 / ties
__global__ void normalize_bwd_k(const float* __restrict__ img, const float* __restrict__ stats,
                                const float* __restrict__ g_gray, const float* __restrict__ dots,
                                int64_t n_pix, float* __restrict__ g_img) {
  const int view = blockIdx.y;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pix) return;
  const float m = stats[2 * view], ties = stats[2 * view + 1];
  const float x = img[view * n_pix + i];
  float g = g_gray[view * n_pix + i] / m;
  if (x == m) g -= dots[view] / (m * m) / ties;
  g_img[view * n_pix + i] = g;
}

// legacy bilinear: src = dst * in/out, lower = floor, upper = min(lower+1, in-1)
__global__ void resize_fwd_k(const float* __restrict__ in, int H, int W, int C, int OH, int OW, float sy,
                             float sx, float* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)OH * OW * C;
  if (t >= total) return;
  const int img = blockIdx.y;
  const int c = (int)(t % C), ox = (int)((t / C) % OW), oy = (int)(t / ((int64_t)C * OW));
  const float fy = (float)oy * sy, fx = (float)ox * sx;
  const int y0 = min((int)floorf(fy), H - 1), x0 = min((int)floorf(fx), W - 1);
  const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
  const float ly = fy - (float)y0, lx = fx - (float)x0;
  const float* b = in + (int64_t)img * H * W * C;
  const float tl = b[((int64_t)y0 * W + x0) * C + c], tr = b[((int64_t)y0 * W + x1) * C + c];
  const float bl = b[((int64_t)y1 * W + x0) * C + c], br = b[((int64_t)y1 * W + x1) * C + c];
  const float top = tl + (tr - tl) * lx, bot = bl + (br - bl) * lx;
  out[(int64_t)img * total + t] = top + (bot - top) * ly;
}
__global__ void resize_bwd_k(const float* __restrict__ g_out, int H, int W, int C, int OH, int OW, float sy,
                             float sx, float* __restrict__ g_in) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)OH * OW * C;
  if (t >= total) return;
  const int img = blockIdx.y;
  const int c = (int)(t % C), ox = (int)((t / C) % OW), oy = (int)(t / ((int64_t)C * OW));
  const float fy = (float)oy * sy, fx = (float)ox * sx;
  const int y0 = min((int)floorf(fy), H - 1), x0 = min((int)floorf(fx), W - 1);
  const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
  const float ly = fy - (float)y0, lx = fx - (float)x0;
  const float g = g_out[(int64_t)img * total + t];
  float* b = g_in + (int64_t)img * H * W * C;
  atomicAdd(b + ((int64_t)y0 * W + x0) * C + c, g * (1.f - ly) * (1.f - lx));
  atomicAdd(b + ((int64_t)y0 * W + x1) * C + c, g * (1.f - ly) * lx);
  atomicAdd(b + ((int64_t)y1 * W + x0) * C + c, g * ly * (1.f - lx));
  atomicAdd(b + ((int64_t)y1 * W + x1) * C + c, g * ly * lx);
}

// tf.compat.v1.image.resize(BICUBIC), align_corners=False, legacy coordinates src = dst*in/out, Keys kernel
// a = -0.75, taps clamped to the image (styler_base.py:166: the style mask).  Closed-form weights (TF reads
// them from a 1024-entry table; the difference is below fp32 noise of the loss -- DESIGN.md section 5).
__device__ __forceinline__ float keys_w(float t) {
  const float a = -0.75f;
  t = fabsf(t);
  if (t <= 1.f) return ((a + 2.f) * t - (a + 3.f)) * t * t + 1.f;
  if (t < 2.f) return ((a * t - 5.f * a) * t + 8.f * a) * t - 4.f * a;
  return 0.f;
}
__global__ void resize_bicubic_fwd_k(const float* __restrict__ in, int H, int W, int C, int OH, int OW, float sy,
                                     float sx, float* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)OH * OW * C;
  if (t >= total) return;
  const int img = blockIdx.y;
  const int c = (int)(t % C), ox = (int)((t / C) % OW), oy = (int)(t / ((int64_t)C * OW));
  const float fy = (float)oy * sy, fx = (float)ox * sx;
  const float by = floorf(fy), bx = floorf(fx);
  const float* b = in + (int64_t)img * H * W * C;
  // rows first, then columns: the summation order of the separable restatement
  float colv[4];
#pragma unroll
  for (int kx = 0; kx < 4; ++kx) colv[kx] = 0.f;
  float acc = 0.f;
#pragma unroll
  for (int kx = 0; kx < 4; ++kx) {
    const int xx = min(max((int)bx + kx - 1, 0), W - 1);
    float r = 0.f;
#pragma unroll
    for (int ky = 0; ky < 4; ++ky) {
      const int yy = min(max((int)by + ky - 1, 0), H - 1);
      r += b[((int64_t)yy * W + xx) * C + c] * keys_w(fy - (by + (float)(ky - 1)));
    }
    acc += r * keys_w(fx - (bx + (float)(kx - 1)));
  }
  out[(int64_t)img * total + t] = acc;
}

// Transpose of resize_bicubic_fwd_k: every output pixel's cotangent goes to its 16 (clamped) taps with the same
// weights; taps clamped onto the border accumulate there.  g_in is zeroed by the entry point.
__global__ void resize_bicubic_bwd_k(const float* __restrict__ g_out, int H, int W, int C, int OH, int OW, float sy,
                                     float sx, float* __restrict__ g_in) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)OH * OW * C;
  if (t >= total) return;
  const int img = blockIdx.y;
  const int c = (int)(t % C), ox = (int)((t / C) % OW), oy = (int)(t / ((int64_t)C * OW));
  const float fy = (float)oy * sy, fx = (float)ox * sx;
  const float by = floorf(fy), bx = floorf(fx);
  const float g = g_out[(int64_t)img * total + t];
  float* b = g_in + (int64_t)img * H * W * C;
#pragma unroll
  for (int kx = 0; kx < 4; ++kx) {
    const int xx = min(max((int)bx + kx - 1, 0), W - 1);
    const float wx = keys_w(fx - (bx + (float)(kx - 1)));
#pragma unroll
    for (int ky = 0; ky < 4; ++ky) {
      const int yy = min(max((int)by + ky - 1, 0), H - 1);
      atomicAdd(b + ((int64_t)yy * W + xx) * C + c, g * wx * keys_w(fy - (by + (float)(ky - 1))));
    }
  }
}

// out[p] (+)= sum_c a[p,c] * b[p,c] + scale * (*scalar): the style-mask cotangent (styler_base.py:165-169):
// d loss / d mask = <d loss / d (F*m), F> per pixel plus the pixel-independent term through the masked area.
__global__ void rowdot_k(const float* __restrict__ a, const float* __restrict__ b, int C, int64_t P,
                         const float* __restrict__ scalar, float scale, int accumulate, float* __restrict__ out) {
  const int64_t p = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);     // one warp per pixel row
  if (p >= P) return;
  const int lane = threadIdx.x & 31;
  float acc = 0.f;
  for (int c = lane; c < C; c += 32) acc = fmaf(a[p * C + c], b[p * C + c], acc);
  acc = lnst_warp_sum(acc);
  if (lane == 0) {
    const float v = acc + (scalar ? scale * scalar[0] : 0.f);
    out[p] = accumulate ? out[p] + v : v;
  }
}

__constant__ float kMeanRGB[3] = {(float)(0.485 * 255), (float)(0.456 * 255), (float)(0.406 * 255)};   // vgg.py:16-18

__global__ void to_net_input_fwd_k(const float* __restrict__ gray, int64_t total_pix, int Cg, float s,
                                   float* __restrict__ d_img, float* __restrict__ x) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total_pix) return;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v = gray[i * Cg + (Cg == 1 ? 0 : c)] * s;
    d_img[i * 3 + c] = v;
    if (x) x[i * 3 + c] = v - kMeanRGB[c];
  }
}
__global__ void to_net_input_bwd_k(const float* __restrict__ g_x, int64_t total_pix, int Cg, float s,
                                   float* __restrict__ g_gray) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total_pix) return;
  const float a = g_x[i * 3], b = g_x[i * 3 + 1], c = g_x[i * 3 + 2];
  if (Cg == 1) {
    g_gray[i] = (a + b + c) * s;
  } else {
    g_gray[i * 3] = a * s; g_gray[i * 3 + 1] = b * s; g_gray[i * 3 + 2] = c * s;
  }
}

// ---------------------------------------------------------------------------------------
// C-ABI
// ---------------------------------------------------------------------------------------
extern "C" int lnst_rotate_fwd(const float* vol, const float* rot, int32_t n_views, int32_t D, int32_t H,
                               int32_t W, float* out, void* stream) {
  if (!vol || !rot || !out || n_views < 1 || D < 1 || H < 1 || W < 1) return LNST_EARG;
  const VolDims v = make_dims(D, H, W);
  const int64_t V = (int64_t)D * H * W;
  LNST_LAUNCH(rotate_fwd_k, dim3(lnst_blocks(V, 256), n_views), dim3(256), 0, lnst_stream(stream), vol, rot,
              v, out);
  return lnst_status();
}

extern "C" int lnst_rotate_bwd(const float* g_out, const float* rot, int32_t n_views, int32_t D, int32_t H,
                               int32_t W, float* g_vol, void* stream) {
  if (!g_out || !rot || !g_vol || n_views < 1 || D < 1 || H < 1 || W < 1) return LNST_EARG;
  const VolDims v = make_dims(D, H, W);
  const int64_t V = (int64_t)D * H * W;
  LNST_LAUNCH(rotate_bwd_k, dim3(lnst_blocks(V, 256), n_views), dim3(256), 0, lnst_stream(stream), g_out, rot,
              v, g_vol);
  return lnst_status();
}

// tuning switch (tests / microbenchmarks): 0 = plain atomics, 1 = shuffle-merge the x-neighbour atomics of the backward
// (warps = pixel rows), 2 = warps = 4 x 8 pixel patches with x and y merges (default)
static int lnst_raymarch_merge = 2;
extern "C" int lnst_set_raymarch_merge(int32_t mode) { lnst_raymarch_merge = mode < 0 ? 0 : (mode > 2 ? 2 : mode); return LNST_OK; }

static inline Bricks make_bricks(const unsigned char* occ, int H, int W) {
  Bricks b;
  b.occ = occ;
  b.by = (H + LNST_BRICK - 1) / LNST_BRICK;
  b.bx = (W + LNST_BRICK - 1) / LNST_BRICK;
  return b;
}

extern "C" int lnst_ray_intervals(const float* rot, int32_t n_views, int32_t D, int32_t H, int32_t W,
                                  const LnstBox* box, const unsigned char* bricks, int32_t* intervals, void* stream) {
  if (!rot || !intervals || n_views < 1 || D < 2 || H < 2 || W < 2 || !box_ok(box, D, H, W)) return LNST_EARG;
  if ((int64_t)D * H * W >= 0x7fffffff) return LNST_EARG;
  const RayGeo g = make_geo(D, H, W);
  LNST_LAUNCH(ray_intervals_k, dim3(lnst_blocks((int64_t)H * W, 128), n_views), dim3(128), 0, lnst_stream(stream), rot,
              g, make_boxf(box, D, H, W), make_bricks(bricks, H, W), reinterpret_cast<int2*>(intervals));
  return lnst_status();
}

extern "C" int lnst_ray_intervals_exact(const float* rot, int32_t n_views, int32_t D, int32_t H, int32_t W,
                                        const LnstBox* box, const unsigned char* touch, int32_t* intervals, void* stream) {
  if (!rot || !intervals || !touch || n_views < 1 || D < 2 || H < 2 || W < 2 || !box_ok(box, D, H, W)) return LNST_EARG;
  if ((int64_t)D * H * W >= 0x7fffffff) return LNST_EARG;
  const RayGeo g = make_geo(D, H, W);
  LNST_LAUNCH(ray_intervals_exact_k, dim3(lnst_blocks((int64_t)H * W, 128), n_views), dim3(128), 0, lnst_stream(stream),
              rot, g, make_boxf(box, D, H, W), touch, reinterpret_cast<int2*>(intervals));
  return lnst_status();
}

extern "C" int lnst_raymarch_fwd_box(const float* vol, const float* rot, int32_t n_views, int32_t D, int32_t H,
                                     int32_t W, float tau, int32_t liquid, const LnstBox* box,
                                     const int32_t* intervals, float* img, float* stot, void* stream) {
  return lnst_raymarch_fwd_max_box(vol, rot, n_views, D, H, W, tau, liquid, box, intervals, img, stot, nullptr, stream);
}

// The same march; stats[2 v] = max over view v's pixels is reduced by the kernel itself (stats zero on entry, may be NULL).
extern "C" int lnst_raymarch_fwd_max_box(const float* vol, const float* rot, int32_t n_views, int32_t D, int32_t H,
                                         int32_t W, float tau, int32_t liquid, const LnstBox* box,
                                         const int32_t* intervals, float* img, float* stot, float* stats, void* stream) {
  if (!vol || !img || !stot || n_views < 1 || D < 1 || H < 1 || W < 1 || !box_ok(box, D, H, W)) return LNST_EARG;
  if (!rot && n_views != 1) return LNST_EARG;
  if (rot && D >= 2 && H >= 2 && W >= 2 && (int64_t)D * H * W < 0x7fffffff) {
    const RayGeo g = make_geo(D, H, W);
    LNST_LAUNCH(raymarch_rot_fwd_k, dim3(lnst_blocks((int64_t)H * W, 128), n_views), dim3(128), 0,
                lnst_stream(stream), vol, rot, g, make_boxf(box, D, H, W), reinterpret_cast<const int2*>(intervals),
                -tau * 1.4426950408889634f, (int)liquid, img, stot, stats);
    return lnst_status();
  }
  const VolDims v = make_dims(D, H, W);
  LNST_LAUNCH(raymarch_fwd_k, dim3(lnst_blocks((int64_t)H * W, 128), n_views), dim3(128), 0,
              lnst_stream(stream), vol, rot, v, make_subvol(box, D, H, W), tau, (int)liquid, img, stot, stats);
  return lnst_status();
}

extern "C" int lnst_raymarch_fwd(const float* vol, const float* rot, int32_t n_views, int32_t D, int32_t H,
                                 int32_t W, float tau, int32_t liquid, float* img, float* stot,
                                 void* stream) {
  return lnst_raymarch_fwd_box(vol, rot, n_views, D, H, W, tau, liquid, nullptr, nullptr, img, stot, stream);
}

extern "C" int lnst_raymarch_bwd_box(const float* vol, const float* rot, int32_t n_views, int32_t D, int32_t H,
                                     int32_t W, float tau, int32_t liquid, const LnstBox* box,
                                     const int32_t* intervals, const float* stot, const float* g_img,
                                     float* g_vol, void* stream) {
  return lnst_raymarch_bwd_norm_box(vol, rot, n_views, D, H, W, tau, liquid, box, intervals, stot, g_img, nullptr, nullptr,
                                    nullptr, g_vol, stream);
}

// The same backward march fed with the cotangent of the NORMALISED image gray = img / max(img): lnst_normalize_bwd's
// second pass runs inside the kernel's prologue (img, stats = {max, ties} per view, dots[v] = sum g_gray * img).  With
// stats == NULL this is lnst_raymarch_bwd_box.  Needs the rotated-march kernel.
extern "C" int lnst_raymarch_bwd_norm_box(const float* vol, const float* rot, int32_t n_views, int32_t D, int32_t H,
                                          int32_t W, float tau, int32_t liquid, const LnstBox* box,
                                          const int32_t* intervals, const float* stot, const float* g_img,
                                          const float* img, const float* stats, const float* dots, float* g_vol,
                                          void* stream) {
  if (!vol || !stot || !g_img || !g_vol || n_views < 1 || D < 1 || H < 1 || W < 1 || !box_ok(box, D, H, W))
    return LNST_EARG;
  if (stats && (!img || !dots)) return LNST_EARG;
  if (!rot && n_views != 1) return LNST_EARG;
  if (rot && D >= 2 && H >= 2 && W >= 2 && (int64_t)D * H * W < 0x7fffffff) {
    const RayGeo g = make_geo(D, H, W);
    const BoxF bf = make_boxf(box, D, H, W);
    const int2* br = reinterpret_cast<const int2*>(intervals);
    const float ntl2 = -tau * 1.4426950408889634f;
    if (lnst_raymarch_merge == 2) {                   // warps = 4 x 8 pixel patches, blocks = 4 rows x 32 columns
      const int tiles_w = (W + 31) / 32, tiles_h = (H + 3) / 4;
      auto k = raymarch_rot_bwd_k<true, true>;
      LNST_LAUNCH(k, dim3((unsigned)(tiles_w * tiles_h), n_views), dim3(128), 0, lnst_stream(stream), vol, rot, g, bf,
                  br, tau, ntl2, (int)liquid, stot, g_img, g_vol, tiles_w, img, stats, dots);
    } else if (lnst_raymarch_merge) {
      auto k = raymarch_rot_bwd_k<true, false>;
      LNST_LAUNCH(k, dim3(lnst_blocks((int64_t)H * W, 128), n_views), dim3(128), 0, lnst_stream(stream), vol, rot, g,
                  bf, br, tau, ntl2, (int)liquid, stot, g_img, g_vol, 0, img, stats, dots);
    } else {
      auto k = raymarch_rot_bwd_k<false, false>;
      LNST_LAUNCH(k, dim3(lnst_blocks((int64_t)H * W, 128), n_views), dim3(128), 0, lnst_stream(stream), vol, rot, g,
                  bf, br, tau, ntl2, (int)liquid, stot, g_img, g_vol, 0, img, stats, dots);
    }
    return lnst_status();
  }
  const VolDims v = make_dims(D, H, W);
  LNST_LAUNCH(raymarch_bwd_k, dim3(lnst_blocks((int64_t)H * W, 128), n_views), dim3(128), 0,
              lnst_stream(stream), vol, rot, v, make_subvol(box, D, H, W), tau, (int)liquid, stot, g_img, g_vol, 0, img,
              stats, dots);
  return lnst_status();
}

extern "C" int lnst_raymarch_bwd(const float* vol, const float* rot, int32_t n_views, int32_t D, int32_t H,
                                 int32_t W, float tau, int32_t liquid, const float* stot,
                                 const float* g_img, float* g_vol, void* stream) {
  return lnst_raymarch_bwd_box(vol, rot, n_views, D, H, W, tau, liquid, nullptr, nullptr, stot, g_img, g_vol, stream);
}

extern "C" int lnst_image_max(const float* img, int32_t n_img, int64_t n_pix, float* stats, void* stream) {
  if (!img || !stats || n_img < 1 || n_pix < 1) return LNST_EARG;
  cudaMemsetAsync(stats, 0, sizeof(float) * 2 * n_img, lnst_stream(stream));
  const unsigned nb = (unsigned)((n_pix + 1023) / 1024 > 64 ? 64 : (n_pix + 1023) / 1024);
  LNST_LAUNCH(image_max_k, dim3(nb, n_img), dim3(256), 0, lnst_stream(stream), img, n_pix, stats);
  LNST_LAUNCH(image_ties_k, dim3(nb, n_img), dim3(256), 0, lnst_stream(stream), img, n_pix, stats);
  return lnst_status();
}

extern "C" int lnst_normalize_fwd(const float* img, const float* stats, int32_t n_img, int64_t n_pix,
                                  float* gray, void* stream) {
  if (!img || !stats || !gray || n_img < 1 || n_pix < 1) return LNST_EARG;
  LNST_LAUNCH(normalize_fwd_k, dim3(lnst_blocks(n_pix, 256), n_img), dim3(256), 0, lnst_stream(stream), img,
              stats, n_pix, gray);
  return lnst_status();
}

// gray = img / max and stats[2v+1] = number of pixels at the maximum, in one pass; stats[2v] = max must be final
// (lnst_raymarch_fwd_max_*) and stats[2v+1] zero on entry.
extern "C" int lnst_normalize_ties_fwd(const float* img, float* stats, int32_t n_img, int64_t n_pix, float* gray,
                                       void* stream) {
  if (!img || !stats || !gray || n_img < 1 || n_pix < 1) return LNST_EARG;
  LNST_LAUNCH(normalize_ties_fwd_k, dim3(lnst_blocks(n_pix, 256), n_img), dim3(256), 0, lnst_stream(stream), img,
              stats, n_pix, gray);
  return lnst_status();
}

extern "C" int lnst_normalize_bwd(const float* img, const float* stats, const float* g_gray, int32_t n_img,
                                  int64_t n_pix, float* dots, float* g_img, void* stream) {
  if (!img || !stats || !g_gray || !dots || !g_img || n_img < 1 || n_pix < 1) return LNST_EARG;
  cudaMemsetAsync(dots, 0, sizeof(float) * n_img, lnst_stream(stream));
  const unsigned nb = (unsigned)((n_pix + 1023) / 1024 > 64 ? 64 : (n_pix + 1023) / 1024);
  LNST_LAUNCH(dot_k, dim3(nb, n_img), dim3(256), 0, lnst_stream(stream), g_gray, img, n_pix, dots);
  LNST_LAUNCH(normalize_bwd_k, dim3(lnst_blocks(n_pix, 256), n_img), dim3(256), 0, lnst_stream(stream), img,
              stats, g_gray, (const float*)dots, n_pix, g_img);
  return lnst_status();
}

extern "C" int lnst_resize_bilinear_fwd(const float* in, int32_t n_img, int32_t H, int32_t W, int32_t C,
                                        int32_t OH, int32_t OW, float* out, void* stream) {
  if (!in || !out || n_img < 1 || H < 1 || W < 1 || C < 1 || OH < 1 || OW < 1) return LNST_EARG;
  const float sy = (float)H / (float)OH, sx = (float)W / (float)OW;
  LNST_LAUNCH(resize_fwd_k, dim3(lnst_blocks((int64_t)OH * OW * C, 256), n_img), dim3(256), 0,
              lnst_stream(stream), in, (int)H, (int)W, (int)C, (int)OH, (int)OW, sy, sx, out);
  return lnst_status();
}

extern "C" int lnst_resize_bilinear_bwd(const float* g_out, int32_t n_img, int32_t H, int32_t W, int32_t C,
                                        int32_t OH, int32_t OW, float* g_in, void* stream) {
  if (!g_out || !g_in || n_img < 1 || H < 1 || W < 1 || C < 1 || OH < 1 || OW < 1) return LNST_EARG;
  const float sy = (float)H / (float)OH, sx = (float)W / (float)OW;
  cudaMemsetAsync(g_in, 0, sizeof(float) * (int64_t)n_img * H * W * C, lnst_stream(stream));
  LNST_LAUNCH(resize_bwd_k, dim3(lnst_blocks((int64_t)OH * OW * C, 256), n_img), dim3(256), 0,
              lnst_stream(stream), g_out, (int)H, (int)W, (int)C, (int)OH, (int)OW, sy, sx, g_in);
  return lnst_status();
}

extern "C" int lnst_resize_bicubic_fwd(const float* in, int32_t n_img, int32_t H, int32_t W, int32_t C, int32_t OH,
                                       int32_t OW, float* out, void* stream) {
  if (!in || !out || n_img < 1 || H < 1 || W < 1 || C < 1 || OH < 1 || OW < 1) return LNST_EARG;
  const float sy = (float)H / (float)OH, sx = (float)W / (float)OW;
  LNST_LAUNCH(resize_bicubic_fwd_k, dim3(lnst_blocks((int64_t)OH * OW * C, 256), n_img), dim3(256), 0,
              lnst_stream(stream), in, (int)H, (int)W, (int)C, (int)OH, (int)OW, sy, sx, out);
  return lnst_status();
}

extern "C" int lnst_resize_bicubic_bwd(const float* g_out, int32_t n_img, int32_t H, int32_t W, int32_t C, int32_t OH,
                                       int32_t OW, float* g_in, void* stream) {
  if (!g_out || !g_in || n_img < 1 || H < 1 || W < 1 || C < 1 || OH < 1 || OW < 1) return LNST_EARG;
  const float sy = (float)H / (float)OH, sx = (float)W / (float)OW;
  cudaMemsetAsync(g_in, 0, sizeof(float) * (int64_t)n_img * H * W * C, lnst_stream(stream));
  LNST_LAUNCH(resize_bicubic_bwd_k, dim3(lnst_blocks((int64_t)OH * OW * C, 256), n_img), dim3(256), 0,
              lnst_stream(stream), g_out, (int)H, (int)W, (int)C, (int)OH, (int)OW, sy, sx, g_in);
  return lnst_status();
}

extern "C" int lnst_rowdot(const float* a, const float* b, int32_t C, int64_t P, const float* scalar, float scale,
                           int32_t accumulate, float* out, void* stream) {
  if (!a || !b || !out || C < 1 || P < 1) return LNST_EARG;
  LNST_LAUNCH(rowdot_k, dim3(lnst_blocks(P, 8)), dim3(256), 0, lnst_stream(stream), a, b, (int)C, P, scalar, scale,
              (int)accumulate, out);
  return lnst_status();
}

extern "C" int lnst_to_net_input_fwd(const float* gray, int32_t n_img, int64_t n_pix, int32_t Cg, float s,
                                     float* d_img, float* x, void* stream) {
  if (!gray || !d_img || n_img < 1 || n_pix < 1 || (Cg != 1 && Cg != 3)) return LNST_EARG;
  const int64_t total = (int64_t)n_img * n_pix;
  LNST_LAUNCH(to_net_input_fwd_k, dim3(lnst_blocks(total, 256)), dim3(256), 0, lnst_stream(stream), gray,
              total, (int)Cg, s, d_img, x);
  return lnst_status();
}

extern "C" int lnst_to_net_input_bwd(const float* g_x, int32_t n_img, int64_t n_pix, int32_t Cg, float s,
                                     float* g_gray, void* stream) {
  if (!g_x || !g_gray || n_img < 1 || n_pix < 1 || (Cg != 1 && Cg != 3)) return LNST_EARG;
  const int64_t total = (int64_t)n_img * n_pix;
  LNST_LAUNCH(to_net_input_bwd_k, dim3(lnst_blocks(total, 256)), dim3(256), 0, lnst_stream(stream), g_x, total,
              (int)Cg, s, g_gray);
  return lnst_status();
}
