// Field post-processing: 3x3x3 separable smoothing (SAME zero padding) fused with max(d,0)
// (reference: styler_3p.py:112-125, executed there as tf.nn.conv3d + tf.maximum).
//
// One thread produces FOUR consecutive outputs along W from a 3 x 3 x 6 register window, so
// each input row is read as part of a contiguous, coalesced segment and reused by the four
// outputs; rows above/below come from L1/L2.  The sign bit of a zero output records
// "pre-activation < 0" (stored as -0.0f) so the backward pass can apply TF's maximum()
// gradient (which passes at equality) without a separate mask volume.
#include "common.cuh"

#define SM_VEC 4

__device__ __forceinline__ float ld_or_zero(const float* __restrict__ v, int z, int y, int x, int D, int H,
                                            int W) {
  if (z < 0 || z >= D || y < 0 || y >= H || x < 0 || x >= W) return 0.f;
  return v[((int64_t)z * H + y) * W + x];
}

// mode 0: out = relu(conv(in)) with signed zero; mode 1: g_in = conv(g_out * pass(out))
template <int MODE>
__global__ void smooth3_k(const float* __restrict__ in, const float* __restrict__ aux,
                          float* __restrict__ out, int D, int H, int W, SubVol sv, float w_side, float w_mid,
                          int do_conv) {
  const int xw = (sv.ex + SM_VEC - 1) / SM_VEC;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)sv.ez * sv.ey * xw;
  if (t >= total) return;
  const int x0 = sv.ox + (int)(t % xw) * SM_VEC;
  const int y = sv.oy + (int)((t / xw) % sv.ey);
  const int z = sv.oz + (int)(t / ((int64_t)xw * sv.ey));
  const int x_end = sv.ox + sv.ex;
  float acc[SM_VEC] = {0.f, 0.f, 0.f, 0.f};
  const int r = do_conv ? 1 : 0;
  for (int dz = -r; dz <= r; ++dz) {
    const float wz = do_conv ? (dz == 0 ? w_mid : w_side) : 1.f;
    for (int dy = -r; dy <= r; ++dy) {
      const float wy = do_conv ? (dy == 0 ? w_mid : w_side) : 1.f;
      float row[SM_VEC + 2];
#pragma unroll
      for (int j = 0; j < SM_VEC + 2; ++j) {
        const int x = x0 - 1 + j;
        float v = 0.f;
        if (do_conv || (j >= 1 && j <= SM_VEC)) {
          v = ld_or_zero(in, z + dz, y + dy, x, D, H, W);
          if (MODE == 1 && v != 0.f) {
            // backward: mask the incoming gradient by the forward pre-activation sign
            const float o = ld_or_zero(aux, z + dz, y + dy, x, D, H, W);
            if (__float_as_uint(o) >> 31) v = 0.f;   // pre-activation < 0 (stored as -0.0f)
          }
        }
        row[j] = v;
      }
      const float wzy = wz * wy;
#pragma unroll
      for (int j = 0; j < SM_VEC; ++j) {
        if (do_conv)
          acc[j] += wzy * (w_side * row[j] + w_mid * row[j + 1] + w_side * row[j + 2]);
        else
          acc[j] += row[j + 1];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < SM_VEC; ++j) {
    const int x = x0 + j;
    if (x < x_end) {
      float v = acc[j];
      if (MODE == 0) v = (v < 0.f) ? -0.0f : v;
      out[((int64_t)z * H + y) * W + x] = v;
    }
  }
}

// Column form of the same filter (do_conv only): a thread owns SM_VEC outputs along W and walks ZT slices along
// D.  Per slice it loads the 3 x (SM_VEC+2) window once and reduces it over y and x (the kernel is the outer
// product k1 x k1 x k1); the z reduction runs over a rolling window of three such slice values held in
// registers.  18 loads per 4 outputs and slice instead of 54 -- the direct form was L1-bound (ncu: l1tex 83-89 %).
// Every output is computed by the same expression whatever the chunking, so box and full-volume runs agree
// bit for bit.
template <int MODE>
__device__ __forceinline__ void smooth_slice(const float* __restrict__ in, const float* __restrict__ aux, int z, int y,
                                             int x0, int D, int H, int W, float w_side, float w_mid, float (&p)[SM_VEC]) {
#pragma unroll
  for (int j = 0; j < SM_VEC; ++j) p[j] = 0.f;
  if (z < 0 || z >= D) return;
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy) {
    const int yy = y + dy;
    if (yy < 0 || yy >= H) continue;
    const float wy = dy == 0 ? w_mid : w_side;
    const int64_t base = ((int64_t)z * H + yy) * W;
    float row[SM_VEC + 2];
#pragma unroll
    for (int j = 0; j < SM_VEC + 2; ++j) {
      const int x = x0 - 1 + j;
      float v = 0.f;
      if (x >= 0 && x < W) {
        v = in[base + x];
        if (MODE == 1 && v != 0.f && (__float_as_uint(aux[base + x]) >> 31)) v = 0.f;   // pre-activation < 0
      }
      row[j] = v;
    }
#pragma unroll
    for (int j = 0; j < SM_VEC; ++j) p[j] += wy * (w_side * row[j] + w_mid * row[j + 1] + w_side * row[j + 2]);
  }
}

template <int MODE>
__global__ void smooth3_col_k(const float* __restrict__ in, const float* __restrict__ aux, float* __restrict__ out,
                              int D, int H, int W, SubVol sv, float w_side, float w_mid, int ZT) {
  const int xw = (sv.ex + SM_VEC - 1) / SM_VEC;
  const int zc = (sv.ez + ZT - 1) / ZT;
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;       // 32-bit index arithmetic: < 2^31 cells
  if (t >= (unsigned)(zc * sv.ey * xw)) return;
  const unsigned q = t / (unsigned)xw, zq = q / (unsigned)sv.ey;
  const int x0 = sv.ox + (int)(t - q * (unsigned)xw) * SM_VEC;
  const int y = sv.oy + (int)(q - zq * (unsigned)sv.ey);
  const int z0 = sv.oz + (int)zq * ZT;
  const int z1 = min(z0 + ZT, sv.oz + sv.ez), x_end = sv.ox + sv.ex;
  float a[SM_VEC], b[SM_VEC], c[SM_VEC];
  smooth_slice<MODE>(in, aux, z0 - 1, y, x0, D, H, W, w_side, w_mid, a);
  smooth_slice<MODE>(in, aux, z0, y, x0, D, H, W, w_side, w_mid, b);
  for (int z = z0; z < z1; ++z) {
    smooth_slice<MODE>(in, aux, z + 1, y, x0, D, H, W, w_side, w_mid, c);
#pragma unroll
    for (int j = 0; j < SM_VEC; ++j) {
      float v = w_side * a[j] + w_mid * b[j] + w_side * c[j];
      if (MODE == 0) v = (v < 0.f) ? -0.0f : v;
      if (x0 + j < x_end) out[((int64_t)z * H + y) * W + x0 + j] = v;
      a[j] = b[j];
      b[j] = c[j];
    }
  }
}

// slices per thread: long columns while the launch still fills the machine
static inline int smooth_zt(const SubVol& sv) {
  const int64_t cols = (int64_t)sv.ey * ((sv.ex + SM_VEC - 1) / SM_VEC);
  return (cols * ((sv.ez + 7) / 8) >= 148 * 1536) ? 8 : 4;
}

static inline void smooth_weights(int k, float& side, float& mid) {
  // k1 = [1,k,1], K = k1 x k1 x k1 / sum(K) with sum(K) = (k+2)^3 (styler_3p.py:115-120)
  const float s = (float)(k + 2);
  side = 1.f / s;
  mid = (float)k / s;
}

__global__ void fill_box_k(float* __restrict__ vol, int H, int W, SubVol sv, float value) {
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;       // 32-bit index arithmetic: < 2^31 cells
  if (t >= (unsigned)(sv.ez * sv.ey * sv.ex)) return;
  const unsigned q = t / (unsigned)sv.ex, zq = q / (unsigned)sv.ey;
  const int x = sv.ox + (int)(t - q * (unsigned)sv.ex), y = sv.oy + (int)(q - zq * (unsigned)sv.ey);
  const int z = sv.oz + (int)zq;
  vol[((int64_t)z * H + y) * W + x] = value;
}

// one 16-byte piece of a row per thread: a float4 store where the piece lies inside the box, scalar stores at its x edges
__global__ void fill_box4_k(float* __restrict__ vol, int H, int W, SubVol sv, int x4lo, int nx4, float value) {
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (unsigned)(sv.ez * sv.ey * nx4)) return;
  const unsigned q = t / (unsigned)nx4, zq = q / (unsigned)sv.ey;
  const int x = (x4lo + (int)(t - q * (unsigned)nx4)) * 4, y = sv.oy + (int)(q - zq * (unsigned)sv.ey);
  const int z = sv.oz + (int)zq;
  float* p = vol + ((int64_t)z * H + y) * W + x;
  if (x >= sv.ox && x + 3 < sv.ox + sv.ex) {
    *reinterpret_cast<float4*>(p) = make_float4(value, value, value, value);
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (x + e >= sv.ox && x + e < sv.ox + sv.ex) p[e] = value;
  }
}

extern "C" int lnst_fill_box(float* vol, int32_t D, int32_t H, int32_t W, const LnstBox* box, float value,
                             void* stream) {
  if (!vol || D < 1 || H < 1 || W < 1 || !box_ok(box, D, H, W)) return LNST_EARG;
  const SubVol sv = make_subvol(box, D, H, W);
  if (W % 4 == 0 && (reinterpret_cast<uintptr_t>(vol) & 15u) == 0) {
    const int x4lo = sv.ox / 4, nx4 = (sv.ox + sv.ex + 3) / 4 - x4lo;
    const int64_t total4 = (int64_t)sv.ez * sv.ey * nx4;
    LNST_LAUNCH(fill_box4_k, dim3(lnst_blocks(total4, 256)), dim3(256), 0, lnst_stream(stream), vol, (int)H, (int)W, sv,
                x4lo, nx4, value);
    return lnst_status();
  }
  const int64_t total = (int64_t)sv.ez * sv.ey * sv.ex;
  LNST_LAUNCH(fill_box_k, dim3(lnst_blocks(total, 256)), dim3(256), 0, lnst_stream(stream), vol, (int)H, (int)W, sv,
              value);
  return lnst_status();
}

extern "C" int lnst_smooth3_relu_fwd_box(const float* in, float* out, int32_t D, int32_t H, int32_t W,
                                         int32_t k, const LnstBox* box, void* stream) {
  if (!in || !out || D < 1 || H < 1 || W < 1 || !box_ok(box, D, H, W)) return LNST_EARG;
  float side = 0.f, mid = 1.f;
  if (k > 0) smooth_weights(k, side, mid);
  const SubVol sv = make_subvol(box, D, H, W);
  if (k > 0) {
    const int zt = smooth_zt(sv);
    const int64_t total = (int64_t)((sv.ez + zt - 1) / zt) * sv.ey * ((sv.ex + SM_VEC - 1) / SM_VEC);
    auto kern = smooth3_col_k<0>;
    LNST_LAUNCH(kern, dim3(lnst_blocks(total, 128)), dim3(128), 0, lnst_stream(stream), in, (const float*)nullptr, out,
                (int)D, (int)H, (int)W, sv, side, mid, zt);
    return lnst_status();
  }
  const int64_t total = (int64_t)sv.ez * sv.ey * ((sv.ex + SM_VEC - 1) / SM_VEC);
  auto kern = smooth3_k<0>;
  LNST_LAUNCH(kern, dim3(lnst_blocks(total, 256)), dim3(256), 0, lnst_stream(stream), in,
              (const float*)nullptr, out, (int)D, (int)H, (int)W, sv, side, mid, (int)(k > 0));
  return lnst_status();
}

extern "C" int lnst_smooth3_relu_bwd_box(const float* g_out, const float* out, float* g_in, int32_t D,
                                         int32_t H, int32_t W, int32_t k, const LnstBox* box, void* stream) {
  if (!g_out || !out || !g_in || D < 1 || H < 1 || W < 1 || !box_ok(box, D, H, W)) return LNST_EARG;
  float side = 0.f, mid = 1.f;
  if (k > 0) smooth_weights(k, side, mid);
  const SubVol sv = make_subvol(box, D, H, W);
  if (k > 0) {
    const int zt = smooth_zt(sv);
    const int64_t total = (int64_t)((sv.ez + zt - 1) / zt) * sv.ey * ((sv.ex + SM_VEC - 1) / SM_VEC);
    auto kern = smooth3_col_k<1>;
    LNST_LAUNCH(kern, dim3(lnst_blocks(total, 128)), dim3(128), 0, lnst_stream(stream), g_out, out, g_in, (int)D,
                (int)H, (int)W, sv, side, mid, zt);
    return lnst_status();
  }
  const int64_t total = (int64_t)sv.ez * sv.ey * ((sv.ex + SM_VEC - 1) / SM_VEC);
  auto kern = smooth3_k<1>;
  LNST_LAUNCH(kern, dim3(lnst_blocks(total, 256)), dim3(256), 0, lnst_stream(stream), g_out, out, g_in,
              (int)D, (int)H, (int)W, sv, side, mid, (int)(k > 0));
  return lnst_status();
}

extern "C" int lnst_smooth3_relu_fwd(const float* in, float* out, int32_t D, int32_t H, int32_t W,
                                     int32_t k, void* stream) {
  return lnst_smooth3_relu_fwd_box(in, out, D, H, W, k, nullptr, stream);
}

extern "C" int lnst_smooth3_relu_bwd(const float* g_out, const float* out, float* g_in, int32_t D,
                                     int32_t H, int32_t W, int32_t k, void* stream) {
  return lnst_smooth3_relu_bwd_box(g_out, out, g_in, D, H, W, k, nullptr, stream);
}
