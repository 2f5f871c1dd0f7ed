// Volume kernels of the stylisation step with their 3-D tiles staged in shared memory by TMA (BASELINE north_star:
// "the splat, ray-march and advect kernels stage 3D tiles in shared memory via TMA").  CUDA only -- the SIMT versions in
// field.cu / render.cu / splat.cu / optim.cu stay as the path of the CPU interpreter, of volumes TMA cannot address
// (W % 4 != 0) and of views too oblique for a fixed slab box.
//   reference: styler_3p.py:112-125 (3x3x3 smoothing + ReLU), :148-158 (render), transform.py:611-628,343-433 (rotate),
//              transform.py:1577-1704 (p2g_wavg), :557-609 (advect).
#include <cuda_bf16.h>
#include "tma_tiles.cuh"
#include "render_common.cuh"
#include "splat_common.cuh"

// =====================================================================================================================
// 3x3x3 smoothing (+ ReLU / ReLU-mask): one TMA box {SX+4, SY+2, SZ+2} per tile of SZ x SY x SX outputs.  The zero
// fill of out-of-volume coordinates IS the SAME padding of tf.nn.conv3d (styler_3p.py:112-121).  A thread owns one (y,x)
// column of the tile and walks its SZ outputs with a rolling window of three plane values; a plane value is the 3x3
// in-plane filter read from shared memory (9 LDS, conflict-free: a warp is 32 consecutive x of one row).
// =====================================================================================================================
namespace sm3 {
constexpr int SZ = 8, SY = 16, SX = 32;
constexpr int BZ = SZ + 2, BY = SY + 2, BX = SX + 8;        // box: the innermost start coordinate must be a multiple of 4
                                                            // floats (TMA: 16-byte aligned global address of the box start;
                                                            // measured: x = -1 or 1 raises an illegal-instruction fault),
                                                            // so the halo column x0-1 sits 0..3 floats into the box
constexpr int BOX = BZ * BY * BX;                           // 7200 floats
constexpr int BOX_PAD = (BOX * 4 + 127) / 128 * 32;         // floats to the next 128-byte boundary (second TMA destination)
constexpr int THREADS = SY * SX;                            // 512
}

// MODE 0: out = relu(conv(in)), negative pre-activations stored as -0.0f (field.cu);  MODE 1: g_in = conv(g_out * pass(out))
template <int MODE>
__global__ void __launch_bounds__(sm3::THREADS) smooth3_tma_k(const __grid_constant__ CUtensorMap map_in,
                                                              const __grid_constant__ CUtensorMap map_aux,
                                                              float* __restrict__ out, int H, int W, SubVol sv,
                                                              float w_side, float w_mid, int tiles_y, int tiles_x) {
  using namespace sm3;
  extern __shared__ unsigned char smem_raw[];                 // TMA destinations need 128-byte alignment: aligned by hand
  float* tin = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  float* taux = tin + BOX_PAD;                                  // MODE 1 only
  __shared__ __align__(8) uint64_t bar_store;
  const uint32_t bar = tma::smem_u32(&bar_store);
  const int t = blockIdx.x;
  const int tx_ = t % tiles_x, ty_ = (t / tiles_x) % tiles_y, tz_ = t / (tiles_x * tiles_y);
  const int z0 = sv.oz + tz_ * SZ, y0 = sv.oy + ty_ * SY, x0 = sv.ox + tx_ * SX;
  const int xs = (x0 - 1) & ~3;                             // box start: x0-1 rounded down to a multiple of 4 (also below 0)
  const int xo = x0 - 1 - xs;                               // 0..3
  if (threadIdx.x == 0) {
    tma::mbar_init(bar, 1);
    tma::fence_mbar_init();
    tma::mbar_expect_tx(bar, (MODE == 1 ? 2 : 1) * BOX * 4);
    tma::load_3d(tma::smem_u32(tin), &map_in, bar, xs, y0 - 1, z0 - 1);
    if (MODE == 1) tma::load_3d(tma::smem_u32(taux), &map_aux, bar, xs, y0 - 1, z0 - 1);
  }
  __syncthreads();
  tma::mbar_wait(bar, 0);
  if (MODE == 1) {                                          // mask the incoming gradient by the forward pre-activation sign
    for (int i = threadIdx.x; i < BOX; i += THREADS)
      if (__float_as_uint(taux[i]) >> 31) tin[i] = 0.f;
    __syncthreads();
  }
  const int ly = threadIdx.x >> 5, lx = threadIdx.x & 31;
  const int y = y0 + ly, x = x0 + lx;
  const bool col_ok = y < sv.oy + sv.ey && x < sv.ox + sv.ex;
  const int z_end = min(z0 + SZ, sv.oz + sv.ez);
  auto plane = [&](int zp) -> float {                       // same expression as field.cu smooth_slice
    const float* r = tin + (zp * BY + ly) * BX + lx + xo;
    float p = 0.f;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const float wy = dy == 1 ? w_mid : w_side;
      p += wy * (w_side * r[dy * BX] + w_mid * r[dy * BX + 1] + w_side * r[dy * BX + 2]);
    }
    return p;
  };
  float a = plane(0), b = plane(1);
  float* o = out + ((int64_t)z0 * H + y) * W + x;
  for (int z = z0; z < z_end; ++z) {
    const float c = plane(z - z0 + 2);
    float v = w_side * a + w_mid * b + w_side * c;
    if (MODE == 0) v = (v < 0.f) ? -0.0f : v;
    if (col_ok) *o = v;
    o += (int64_t)H * W;
    a = b; b = c;
  }
}

static inline void smooth_weights_tma(int k, float& side, float& mid) {
  const float s = (float)(k + 2);                           // k1 = [1,k,1] / (k+2) per axis (styler_3p.py:115-120)
  side = 1.f / s;
  mid = (float)k / s;
}

template <int MODE>
static int launch_smooth_tma(const float* in, const float* aux, float* out, int D, int H, int W, int k,
                             const LnstBox* box, cudaStream_t st) {
  using namespace sm3;
  CUtensorMap m_in, m_aux;
  if (!tma::make_volume_map(&m_in, in, D, H, W, BZ, BY, BX)) return LNST_EARG;
  if (MODE == 1) { if (!tma::make_volume_map(&m_aux, aux, D, H, W, BZ, BY, BX)) return LNST_EARG; }
  else m_aux = m_in;
  float side, mid;
  smooth_weights_tma(k, side, mid);
  const SubVol sv = make_subvol(box, D, H, W);
  const int tz = (sv.ez + SZ - 1) / SZ, ty = (sv.ey + SY - 1) / SY, tx = (sv.ex + SX - 1) / SX;
  const int smem = (MODE == 1 ? 2 : 1) * BOX_PAD * 4 + 128;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(smooth3_tma_k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * BOX_PAD * 4 + 128);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  smooth3_tma_k<MODE><<<(unsigned)(tz * ty * tx), THREADS, smem, st>>>(m_in, m_aux, out, H, W, sv, side, mid, ty, tx);
  return (int)cudaGetLastError();
}

extern "C" int lnst_tma_supported(void) { return tma::encode_fn() != nullptr ? 1 : 0; }

extern "C" int lnst_smooth3_relu_fwd_tma(const float* in, float* out, int32_t D, int32_t H, int32_t W, int32_t k,
                                         const LnstBox* box, void* stream) {
  if (!in || !out || D < 1 || H < 1 || W < 1 || k < 1 || !box_ok(box, D, H, W)) return LNST_EARG;
  return launch_smooth_tma<0>(in, nullptr, out, D, H, W, k, box, lnst_stream(stream));
}
extern "C" int lnst_smooth3_relu_bwd_tma(const float* g_out, const float* out, float* g_in, int32_t D, int32_t H,
                                         int32_t W, int32_t k, const LnstBox* box, void* stream) {
  if (!g_out || !out || !g_in || D < 1 || H < 1 || W < 1 || k < 1 || !box_ok(box, D, H, W)) return LNST_EARG;
  return launch_smooth_tma<1>(g_out, out, g_in, D, H, W, k, box, lnst_stream(stream));
}

// =====================================================================================================================
// Weighted-average splat (p2g_wavg, density mode), forward, as a GATHER: one thread per output cell walks the particles
// of its 27 neighbour cells (per-cell lists built once per (frame, octave) -- positions are constants in density mode)
// and accumulates num_k = sum w x and wmap_k = sum w in registers, so there are no atomics, no `num` volumes and no
// separate combine pass (the scatter version: 27 nk atomics per particle + 16 V bytes of combine traffic).  A CTA's
// 4 x 8 x 32 tile of results is assembled in shared memory -- rows already in the reference's flipped H order
// (transform.py:1703) -- and leaves with ONE TMA store; the store clips the tile at the volume's faces.
//   cstart [V+1]: first list entry of every cell (cells in unflipped (z,y,x) order); order [Nv]: particle index of every
//   list entry; rel [Nv,3]: that particle's offset from its cell centre (lnst_splat_cells), in list order.
// =====================================================================================================================
namespace sg {
constexpr int TZ = 4, TY = 8, TX = 32;
constexpr int THREADS = TY * TX;
constexpr int ROWS = (TZ + 2) * (TY + 2);                    // (z,y) rows of cells a tile's outputs can receive from
constexpr int RC = TX + 3;                                   // cell boundaries per row: cells x0-1 .. x0+TX, + the end
constexpr int PMAX = 3072;                                   // particles staged per tile (C3: ~2000); more -> global path
}

// The tile's neighbourhood particles are staged in shared memory first (their offsets and x = r + clip(var), in list
// order -- every (z,y) row of cells is one contiguous list range), with the per-cell list boundaries beside them; each
// thread then walks the 27 neighbour cells of its output cells out of shared memory.  Tiles whose neighbourhood holds
// more than PMAX particles walk the lists in global memory instead.
template <int NK>
__global__ void __launch_bounds__(sg::THREADS) splat_wavg_gather_k(const __grid_constant__ CUtensorMap map_out,
                                                                   const int* __restrict__ cstart,
                                                                   const int* __restrict__ order,
                                                                   const float* __restrict__ rel,
                                                                   const float* __restrict__ r,
                                                                   const float* __restrict__ var, LnstGrid g,
                                                                   SplatKernels ks, int z_base, int y_base, int x_base,
                                                                   int tiles_y, int tiles_x) {
  using namespace sg;
  extern __shared__ unsigned char smem_raw[];
  float* tile = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));   // TMA source
  float* s_rel = tile + TZ * TY * TX;                        // [PMAX][3]
  float* s_x = s_rel + 3 * PMAX;                             // [NK][PMAX]
  int* s_cell = reinterpret_cast<int*>(s_x + NK * PMAX);     // [ROWS][RC] list boundaries, relative to the staged arrays
  __shared__ int s_row_j0[ROWS], s_row_off[ROWS + 1];
  const int D = g.res[0], H = g.res[1], W = g.res[2];
  const int t = blockIdx.x;
  const int tx_ = t % tiles_x, ty_ = (t / tiles_x) % tiles_y, tz_ = t / (tiles_x * tiles_y);
  // tiles are laid out over OUTPUT rows (R = H - 1 - y, transform.py:1703) so that the store's coordinates are never
  // negative (a negative row start faults on the B200 like a misaligned column start; overhang past the far faces is fine)
  const int z0 = z_base + tz_ * TZ, r0 = y_base + ty_ * TY, x0 = x_base + tx_ * TX;
  const int ly = threadIdx.x >> 5, lx = threadIdx.x & 31;
  const int y = H - 1 - (r0 + ly), x = x0 + lx;
  // ---- stage: row q = (rz, ry) holds the cells (z0 - 1 + rz, y_hi + 1 - ry, x0 - 1 .. x0 + TX) ------------------------------
  const int y_hi = H - 1 - r0;                               // unflipped y of output row r0 (the tile's highest y)
  auto row_base = [&](int q, bool& ok) -> int {
    const int zz = z0 - 1 + q / (TY + 2), yy = y_hi + 1 - q % (TY + 2);
    ok = zz >= 0 && zz < D && yy >= 0 && yy < H;
    return (zz * H + yy) * W;
  };
  auto clampx = [&](int xx) -> int { return min(max(xx, 0), W); };
  if (threadIdx.x < ROWS) {
    bool ok;
    const int rb = row_base(threadIdx.x, ok);
    const int j0 = ok ? cstart[rb + clampx(x0 - 1)] : 0, j1 = ok ? cstart[rb + clampx(x0 + TX + 1)] : 0;
    s_row_j0[threadIdx.x] = j0;
    s_row_off[threadIdx.x + 1] = j1 - j0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    s_row_off[0] = 0;
    for (int q = 0; q < ROWS; ++q) { acc += s_row_off[q + 1]; s_row_off[q + 1] = acc; }
  }
  __syncthreads();
  const int total = s_row_off[ROWS];
  const bool staged = total <= PMAX;
  if (staged) {
    for (int e = threadIdx.x; e < ROWS * RC; e += THREADS) {
      const int q = e / RC, c = e - q * RC;
      bool ok;
      const int rb = row_base(q, ok);
      s_cell[e] = ok ? cstart[rb + clampx(x0 - 1 + c)] - s_row_j0[q] + s_row_off[q] : s_row_off[q];
    }
    for (int e = threadIdx.x; e < total; e += THREADS) {      // staged entry e: which row's list does it come from?
      int lo = 0, hi = ROWS;                                 // largest q with s_row_off[q] <= e
      while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s_row_off[mid] <= e) lo = mid; else hi = mid; }
      const int j = s_row_j0[lo] + (e - s_row_off[lo]);
      s_rel[3 * e] = rel[3 * j]; s_rel[3 * e + 1] = rel[3 * j + 1]; s_rel[3 * e + 2] = rel[3 * j + 2];
      const int i = order[j];
#pragma unroll
      for (int k = 0; k < NK; ++k) {
        float v = var ? var[(int64_t)i * NK + k] : 0.f;
        v = fmaxf(fminf(v, 1.f), -1.f);                    // styler_3p.py:74; TF order max(min(x,1),-1): NaN reads as +1
        s_x[k * PMAX + e] = r[(int64_t)i * NK + k] + v;
      }
    }
    __syncthreads();
  }
  float off[3];                                              // (target - home) * cell for -1, 0, +1, as stencil27 forms it
#pragma unroll
  for (int q = 0; q < 3; ++q) off[q] = __fmul_rn((float)(q - 1), g.cell);
#pragma unroll 1
  for (int tz = 0; tz < TZ; ++tz) {
    const int z = z0 + tz;
    float num[NK], wm[NK];
#pragma unroll
    for (int k = 0; k < NK; ++k) { num[k] = 0.f; wm[k] = 0.f; }
    if (x < W && y >= 0 && z < D) {
#pragma unroll 1
      for (int dz = -1; dz <= 1; ++dz) {
        const float oz = dz < 0 ? off[2] : (dz == 0 ? off[1] : off[0]);   // off[1 - dz]
#pragma unroll 1
        for (int dy = -1; dy <= 1; ++dy) {
          const float oy = dy < 0 ? off[2] : (dy == 0 ? off[1] : off[0]);
          if (staged) {
            // row of (z + dz, y + dy): rz = tz + 1 + dz, ry = (y_hi + 1) - (y + dy) = ly + 1 - dy
            const int* cb = s_cell + ((tz + 1 + dz) * (TY + 2) + (ly + 1 - dy)) * RC + lx;   // cell x - 1 is local lx
            const int j0 = cb[0], b1 = cb[1], b2 = cb[2], j1 = cb[3];
            for (int j = j0; j < j1; ++j) {
              const float ox = j < b1 ? off[2] : (j < b2 ? off[1] : off[0]);  // home x - 1, x, x + 1 -> target - home = +1, 0, -1
              const float ddz = __fadd_rn(s_rel[3 * j], -oz);
              const float ddy = __fadd_rn(s_rel[3 * j + 1], -oy);
              const float ddx = __fadd_rn(s_rel[3 * j + 2], -ox);
              const float len = sqrtf(ddx * ddx + ddy * ddy + ddz * ddz);
#pragma unroll
              for (int k = 0; k < NK; ++k) {
                const float w = cubic_w(len * ks.inv_h[k], ks.sigma[k]);
                if (w != 0.f) { num[k] = fmaf(w, s_x[k * PMAX + j], num[k]); wm[k] += w; }
              }
            }
          } else {
            const int zz = z + dz, yy = y + dy;
            if (zz < 0 || zz >= D || yy < 0 || yy >= H) continue;
            const int row = (zz * H + yy) * W;
            const int j0 = cstart[row + max(x - 1, 0)], j1 = cstart[row + min(x + 2, W)];
            const int b1 = cstart[row + x], b2 = cstart[row + x + 1];
            for (int j = j0; j < j1; ++j) {
              const float ox = j < b1 ? off[2] : (j < b2 ? off[1] : off[0]);
              const float ddz = __fadd_rn(rel[3 * j], -oz);
              const float ddy = __fadd_rn(rel[3 * j + 1], -oy);
              const float ddx = __fadd_rn(rel[3 * j + 2], -ox);
              const float len = sqrtf(ddx * ddx + ddy * ddy + ddz * ddz);
              const int i = order[j];
#pragma unroll
              for (int k = 0; k < NK; ++k) {
                const float w = cubic_w(len * ks.inv_h[k], ks.sigma[k]);
                if (w != 0.f) {
                  float v = var ? var[(int64_t)i * NK + k] : 0.f;
                  v = fmaxf(fminf(v, 1.f), -1.f);
                  num[k] = fmaf(w, r[(int64_t)i * NK + k] + v, num[k]);
                  wm[k] += w;
                }
              }
            }
          }
        }
      }
    }
    float sres = 0.f;
#pragma unroll
    for (int k = 0; k < NK; ++k) sres += (wm[k] > 1e-6f) ? num[k] / wm[k] : num[k];   // transform.py:1703
    tile[(tz * TY + ly) * TX + lx] = sres;
  }
  tma::fence_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    tma::store_3d(&map_out, tma::smem_u32(tile), x0, r0, z0);
    tma::store_commit();
    tma::store_wait_all();
  }
}

template <int NK>
static int launch_gather(const CUtensorMap& mo, const int* cstart, const int* order, const float* rel, const float* r,
                         const float* var, const LnstGrid& g, const SplatKernels& ks, int oz, int y_lo, int x_base, int tz,
                         int ty, int tx, cudaStream_t st) {
  using namespace sg;
  const int smem = 128 + 4 * (TZ * TY * TX + 3 * PMAX + NK * PMAX + ROWS * RC);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(splat_wavg_gather_k<NK>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  splat_wavg_gather_k<NK><<<(unsigned)(tz * ty * tx), THREADS, smem, st>>>(mo, cstart, order, rel, r, var, g, ks, oz, y_lo,
                                                                          x_base, ty, tx);
  return (int)cudaGetLastError();
}

extern "C" int lnst_splat_wavg_fwd_gather(const int32_t* cstart, const int32_t* order, const float* rel, const float* r,
                                          const float* var, const LnstGrid* g, const float* h, int32_t nk, float* out,
                                          const LnstBox* box, void* stream) {
  using namespace sg;
  SplatKernels ks;
  if (!grid_ok(g) || g->dim != 3 || g->nsize != 1 || g->clip || !cstart || !order || !rel || !r || !out ||
      !fill_kernels(ks, 3, h, nk))
    return LNST_EARG;
  const int D = g->res[0], H = g->res[1], W = g->res[2];
  if (!box_ok(box, D, H, W) || grid_cells(g) >= 0x7fffffff) return LNST_EARG;
  CUtensorMap mo;
  if (!tma::make_volume_map(&mo, out, D, H, W, TZ, TY, TX)) return LNST_EARG;
  const SubVol sv = make_subvol(box, D, H, W);               // box rows are OUTPUT rows (H already flipped)
  const int y_lo = sv.oy;
  const int x_base = sv.ox & ~3;                             // TMA: innermost start coordinate a multiple of 4 floats
  const int ex = sv.ox + sv.ex - x_base;
  const int tz = (sv.ez + TZ - 1) / TZ, ty = (sv.ey + TY - 1) / TY, tx = (ex + TX - 1) / TX;
  // tiles hang over the box's far faces: the cells there receive their true value (zero outside the particles' reach,
  // which the box contains), clipped at the volume's faces by the store
  cudaStream_t st = lnst_stream(stream);
  switch (nk) {
    case 1: return launch_gather<1>(mo, cstart, order, rel, r, var, *g, ks, sv.oz, y_lo, x_base, tz, ty, tx, st);
    case 2: return launch_gather<2>(mo, cstart, order, rel, r, var, *g, ks, sv.oz, y_lo, x_base, tz, ty, tx, st);
    case 3: return launch_gather<3>(mo, cstart, order, rel, r, var, *g, ks, sv.oz, y_lo, x_base, tz, ty, tx, st);
    default: return launch_gather<4>(mo, cstart, order, rel, r, var, *g, ks, sv.oz, y_lo, x_base, tz, ty, tx, st);
  }
}

// =====================================================================================================================
// Rotated ray-march, forward.  A CTA owns a TH x TW tile of pixels of one view and walks the volume in slabs of BZ
// planes, front (high z) to back: every slab is ONE TMA box {BX, BY, BZ} -- the bounding box of the tile's sample
// footprints inside those planes -- double buffered, so the samples of a slab come out of shared memory while the next
// slab is in flight.  A thread is one ray, as in render.cu, and evaluates each sample with the same functions
// (render_common.cuh) in the same order: the image is bit-identical to raymarch_rot_fwd_k's.
// Views whose footprint does not fit the fixed box (|angle| beyond ~15 degrees) or that do not run towards +z take the
// gather path inside the same kernel -- the decision is per CTA and on the device, because the view matrices live in a
// device buffer that is refreshed between CUDA-graph replays (Poisson view sampling).
// =====================================================================================================================
namespace rm {
constexpr int TH = 8, TW = 32;
constexpr int BY = 16, BX = 44;                             // x origin rounded down to a multiple of 4 (TMA alignment): +3 columns
constexpr int THREADS = TH * TW;                            // planes per slab (BZ): template parameter, 8 / 12 / 16
}

struct Anchor { int z0, y0, x0; float fz, fy, fx; };
// same arithmetic as locate() (render_common.cuh), with the anchor kept as three indices
__device__ __forceinline__ Anchor locate3(const RayLine& l, float fi, const RayGeo& g) {
  const float z = fminf(fmaxf(fmaf(l.kz, fi, l.cz), 0.f), g.mD);
  const float y = fminf(fmaxf(fmaf(l.ky, fi, l.cy), 0.f), g.mH);
  const float x = fminf(fmaxf(fmaf(l.kx, fi, l.cx), 0.f), g.mW);
  Anchor a;
  a.z0 = min((int)z, g.D2); a.y0 = min((int)y, g.H2); a.x0 = min((int)x, g.W2);
  a.fz = z - (float)a.z0; a.fy = y - (float)a.y0; a.fx = x - (float)a.x0;
  return a;
}

// Bounding rows / columns of the tile's sample positions over the planes [za, zb + 1]: along a ray y = Y0 + r z with
// Y0 = cy - r cz affine in the pixel, so the extremes sit at the tile's corner pixels and the slab's end planes.
struct TileSpan { float lo, hi, r; };                       // Y0 range over the tile's 4 corners, slope dy/dz
__device__ __forceinline__ TileSpan tile_span(const float* __restrict__ R, const RayGeo& g, int h0, int h1, int w0, int w1,
                                              int axis) {
  float lo = 3.0e38f, hi = -3.0e38f, r = 0.f;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const RayLine l = ray_line(R, lin_coord((c & 1) ? h1 : h0, g.sH), lin_coord((c & 2) ? w1 : w0, g.sW), g);
    const float k = axis == 1 ? l.ky : l.kx, cc = axis == 1 ? l.cy : l.cx;
    r = k / l.kz;
    const float v = cc - r * l.cz;
    lo = fminf(lo, v); hi = fmaxf(hi, v);
  }
  TileSpan s = {lo, hi, r};
  return s;
}
__device__ __forceinline__ int span_origin(const TileSpan& s, int za, int zb, float m, bool align4) {
  const float a = s.r * (float)za, b = s.r * (float)(zb + 1);
  const float lo = fminf(fmaxf(s.lo + fminf(a, b), 0.f), m);  // positions are clamped to the volume, so is their range
  const int o = (int)floorf(lo) - 1;                          // one voxel of slack against round-off
  return align4 ? (o & ~3) : o;                               // innermost TMA coordinate: a multiple of 4 floats
}

template <int BZ>
__global__ void __launch_bounds__(rm::THREADS) raymarch_fwd_tma_k(const __grid_constant__ CUtensorMap map_vol,
                                                                  const float* __restrict__ vol,
                                                                  const float* __restrict__ rot, RayGeo g, BoxF bf,
                                                                  const int2* __restrict__ iv, float ntl2, int liquid,
                                                                  float* __restrict__ img, float* __restrict__ stot,
                                                                  int tiles_w, float* __restrict__ stats) {
  using namespace rm;
  constexpr int SLAB = BZ * BY * BX;                         // floats per slab (a multiple of 32: 128-byte aligned buffers)
  extern __shared__ unsigned char smem_raw[];
  float* slab = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));   // 2 x SLAB
  __shared__ __align__(8) uint64_t bar_store[2];
  __shared__ int s_ztop, s_zbot;
  const int view = blockIdx.y;
  const float* R = rot + 9 * view;
  const int th = blockIdx.x / tiles_w, tw = blockIdx.x - th * tiles_w;
  const int h0 = th * TH, w0 = tw * TW;
  const int ly = threadIdx.x >> 5, lx = threadIdx.x & 31;
  const int h = h0 + ly, w = w0 + lx;
  const bool valid = h < g.H && w < g.W;
  const int hc = min(h, g.H - 1), wc = min(w, g.W - 1);
  const RayLine l = ray_line(R, lin_coord(hc, g.sH), lin_coord(wc, g.sW), g);
  int i_lo = 1, i = 0;                                       // [i_lo, i]: the ray's live samples, marched downwards
  if (valid) {
    if (iv) { const int2 r = iv[(int64_t)view * g.HW + h * g.W + w]; i_lo = r.x; i = r.y; }
    else ray_interval(l, g, bf, i_lo, i);
  }
  const bool live = valid && i_lo <= i;
  if (threadIdx.x == 0) {
    s_ztop = -1; s_zbot = 0x7fffffff;
    tma::mbar_init(tma::smem_u32(&bar_store[0]), 1);
    tma::mbar_init(tma::smem_u32(&bar_store[1]), 1);
    tma::fence_mbar_init();
  }
  __syncthreads();
  // tile-uniform: does the slab box hold this view's footprint?  (kz, ky, kx do not depend on the pixel)
  const int h1 = min(h0 + TH, g.H) - 1, w1 = min(w0 + TW, g.W) - 1;
  const TileSpan sy = tile_span(R, g, h0, h1, w0, w1, 1), sx = tile_span(R, g, h0, h1, w0, w1, 2);
  const bool fits = l.kz > 0.5f && (sy.hi - sy.lo) + fabsf(sy.r) * (float)BZ + 4.f <= (float)BY &&
                    (sx.hi - sx.lo) + fabsf(sx.r) * (float)BZ + 7.f <= (float)BX;
  float S = 0.f, I = 0.f;
  if (!fits) {                                               // gather path (render.cu arithmetic, no plane carry)
    for (; i >= i_lo; --i) {
      const Cell c = locate(l, (float)i, g);
      const float* p = vol + c.idx;
      const Plane4 lo = load_plane(p, g.W), hi = load_plane(p + g.HW, g.W);
      const float d = lerp_planes(lo, hi, c);
      S += d;
      I = fmaf(d, fast_exp2(S * ntl2), I);
    }
  } else {
    if (live) {
      atomicMax(&s_ztop, locate3(l, (float)i, g).z0);
      atomicMin(&s_zbot, locate3(l, (float)i_lo, g).z0);
    }
    __syncthreads();
    const int ztop = s_ztop, zbot = s_zbot;
    const int nslab = ztop >= zbot ? (ztop - zbot) / (BZ - 1) + 1 : 0;
    // slab s holds the anchors [za, za + BZ - 2] with za = ztop - (BZ - 2) - s (BZ - 1)
    if (threadIdx.x == 0) {
      for (int s = 0; s < 2 && s < nslab; ++s) {
        const int za = ztop - (BZ - 2) - s * (BZ - 1);
        const uint32_t bar = tma::smem_u32(&bar_store[s]);
        tma::mbar_expect_tx(bar, SLAB * 4);
        tma::load_3d(tma::smem_u32(slab + s * SLAB), &map_vol, bar, span_origin(sx, za, za + BZ - 2, g.mW, true),
                     span_origin(sy, za, za + BZ - 2, g.mH, false), za);
      }
    }
    for (int s = 0; s < nslab; ++s) {
      const int buf = s & 1;
      const int za = ztop - (BZ - 2) - s * (BZ - 1);
      const int oy = span_origin(sy, za, za + BZ - 2, g.mH, false), ox = span_origin(sx, za, za + BZ - 2, g.mW, true);
      tma::mbar_wait(tma::smem_u32(&bar_store[buf]), (uint32_t)((s >> 1) & 1));
      const float* sb = slab + buf * SLAB;
      // interior slab: the staged box lies inside the volume and no sample of it is clamped -- positions need no
      // clamping and every footprint is inside the box by construction (tile-uniform test)
      const bool interior = za >= 1 && za + BZ <= g.D - 1 && oy >= 0 && oy + BY <= g.H && ox >= 0 && ox + BX <= g.W;
      if (interior) {
        const int base = -((za * BY + oy) * BX + ox);
        auto sample_in = [&](float fi, int& z0) -> float {   // same values as locate3 + lerp_planes, clamps are no-ops here
          const float z = fmaf(l.kz, fi, l.cz), y = fmaf(l.ky, fi, l.cy), x = fmaf(l.kx, fi, l.cx);
          z0 = (int)z;
          const int y0 = (int)y, x0 = (int)x;
          const float* p = sb + ((z0 * BY + y0) * BX + x0 + base);
          Plane4 lo, hi;
          lo.a = p[0]; lo.b = p[1]; lo.c = p[BX]; lo.d = p[BX + 1];
          hi.a = p[BY * BX]; hi.b = p[BY * BX + 1]; hi.c = p[BY * BX + BX]; hi.d = p[BY * BX + BX + 1];
          Cell c; c.idx = 0; c.fz = z - (float)z0; c.fy = y - (float)y0; c.fx = x - (float)x0;
          return lerp_planes(lo, hi, c);
        };
        while (i - 3 >= i_lo) {
          if ((int)fmaf(l.kz, (float)(i - 3), l.cz) < za) break;     // the group's lowest sample is in a later slab
          int zz;
          const float d0 = sample_in((float)i, zz), d1 = sample_in((float)(i - 1), zz), d2 = sample_in((float)(i - 2), zz),
                      d3 = sample_in((float)(i - 3), zz);
          S += d0; I = fmaf(d0, fast_exp2(S * ntl2), I);     // inclusive reverse cumsum, styler_3p.py:155
          S += d1; I = fmaf(d1, fast_exp2(S * ntl2), I);
          S += d2; I = fmaf(d2, fast_exp2(S * ntl2), I);
          S += d3; I = fmaf(d3, fast_exp2(S * ntl2), I);
          i -= 4;
        }
        while (i >= i_lo) {
          if ((int)fmaf(l.kz, (float)i, l.cz) < za) break;
          int zz;
          const float d = sample_in((float)i, zz);
          S += d;
          I = fmaf(d, fast_exp2(S * ntl2), I);
          --i;
        }
      } else {
        auto sample = [&](const Anchor& a) -> float {        // one trilinear sample out of the staged slab
          const int ry = a.y0 - oy, rx = a.x0 - ox;
          Plane4 lo, hi;
          if ((unsigned)ry < (unsigned)(BY - 1) && (unsigned)rx < (unsigned)(BX - 1)) {
            const float* p = sb + ((a.z0 - za) * BY + ry) * BX + rx;
            lo.a = p[0]; lo.b = p[1]; lo.c = p[BX]; lo.d = p[BX + 1];
            const float* q = p + BY * BX;
            hi.a = q[0]; hi.b = q[1]; hi.c = q[BX]; hi.d = q[BX + 1];
          } else {                                           // clamped onto a face, outside the staged box: gather
            const float* p = vol + ((int64_t)a.z0 * g.H + a.y0) * g.W + a.x0;
            lo = load_plane(p, g.W); hi = load_plane(p + g.HW, g.W);
          }
          Cell c; c.idx = 0; c.fz = a.fz; c.fy = a.fy; c.fx = a.fx;
          return lerp_planes(lo, hi, c);
        };
        while (i >= i_lo) {
          const Anchor a = locate3(l, (float)i, g);
          if (a.z0 < za) break;                              // belongs to a slab further back
          const float d = sample(a);
          S += d;
          I = fmaf(d, fast_exp2(S * ntl2), I);
          --i;
        }
      }
      __syncthreads();                                       // every ray is done with this buffer
      if (threadIdx.x == 0 && s + 2 < nslab) {
        const int zn = ztop - (BZ - 2) - (s + 2) * (BZ - 1);
        const uint32_t bar = tma::smem_u32(&bar_store[buf]);
        tma::mbar_expect_tx(bar, SLAB * 4);
        tma::load_3d(tma::smem_u32(slab + buf * SLAB), &map_vol, bar, span_origin(sx, zn, zn + BZ - 2, g.mW, true),
                     span_origin(sy, zn, zn + BZ - 2, g.mH, false), zn);
      }
    }
  }
  if (liquid) I = 1.f - fast_exp2(S * ntl2);                 // styler_3p.py:150-152
  if (valid) {
    img[(int64_t)view * g.HW + h * g.W + w] = I;
    stot[(int64_t)view * g.HW + h * g.W + w] = S;
  }
  if (stats != nullptr) {                                    // the view's maximum (image_max_k), one atomic per warp
    const float m = lnst_warp_max(valid ? I : 0.f);
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(stats + 2 * view), __float_as_int(m));
  }
}


// =====================================================================================================================
// Rotated ray-march, backward.  Same CTA tile and TMA slabs as the forward kernel (the density is re-sampled from shared
// memory to rebuild T_k and the running sums), ascending in z.  What bounds this kernel is the scatter of the 8 corner
// contributions of every sample (red.global.add.f32: LSU-bound, ~1.3 cycles per lane and instruction), so the thread
// layout is chosen for MERGING them before they leave the SM:
//   * a warp is a 4 x 8 pixel patch; its lanes walk the depth index in lockstep (warp-uniform i range per slab);
//   * depth: a sample's far plane waits in registers and is merged into the next sample's near plane (anchor + 1 in z);
//   * x: lane (r,c) hands its x1 column to lane (r,c+1) when that lane's anchor is one voxel to the right;
//   * y: lane (r,c) hands its (y1,x0) value to lane (r+1,c) when that lane's anchor is one voxel down.
// With the reference's small view angles nearly every voxel then receives ONE atomic per view instead of eight.
// Slabs overlap by two planes so that a warp-uniform i range exists whose samples all sit inside one slab (the z shear
// across a 4 x 8 patch is below two voxels for every view the slab box admits).
// =====================================================================================================================
template <int BZ>
__global__ void __launch_bounds__(rm::THREADS) raymarch_bwd_tma_k(const __grid_constant__ CUtensorMap map_vol,
                                                                  const float* __restrict__ vol,
                                                                  const float* __restrict__ rot, RayGeo g, BoxF bf,
                                                                  const int2* __restrict__ iv, float tau, float ntl2,
                                                                  const float* __restrict__ stot,
                                                                  const float* __restrict__ g_img,
                                                                  float* __restrict__ g_vol, int tiles_w) {
  using namespace rm;
  constexpr int SLAB = BZ * BY * BX;
  constexpr int ADV = BZ - 3;                                 // slab advance: BZ - 1 anchors per slab, two of them overlap
  extern __shared__ unsigned char smem_raw[];
  float* slab = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  __shared__ __align__(8) uint64_t bar_store[2];
  __shared__ int s_ztop, s_zbot;
  const int view = blockIdx.y;
  const float* R = rot + 9 * view;
  const int th = blockIdx.x / tiles_w, tw = blockIdx.x - th * tiles_w;
  const int h0 = th * TH, w0 = tw * TW;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ly = (warp >> 2) * 4 + (lane >> 3), lx = (warp & 3) * 8 + (lane & 7);   // 4 x 8 pixel patch per warp
  const int h = h0 + ly, w = w0 + lx;
  const bool valid = h < g.H && w < g.W;
  const int hc = min(h, g.H - 1), wc = min(w, g.W - 1);
  const RayLine l = ray_line(R, lin_coord(hc, g.sH), lin_coord(wc, g.sW), g);
  float gI = 0.f, St = 0.f;
  int i_lo = 0x7fffffff, i_hi = -1;
  if (valid) {
    gI = g_img[(int64_t)view * g.HW + h * g.W + w];
    St = stot[(int64_t)view * g.HW + h * g.W + w];
    if (gI != 0.f) {
      if (iv) { const int2 r = iv[(int64_t)view * g.HW + h * g.W + w]; i_lo = r.x; i_hi = r.y; }
      else ray_interval(l, g, bf, i_lo, i_hi);
      if (i_lo > i_hi) { i_lo = 0x7fffffff; i_hi = -1; }
    }
  }
  const bool has = i_lo <= i_hi;
  if (threadIdx.x == 0) {
    s_ztop = -1; s_zbot = 0x7fffffff;
    tma::mbar_init(tma::smem_u32(&bar_store[0]), 1);
    tma::mbar_init(tma::smem_u32(&bar_store[1]), 1);
    tma::fence_mbar_init();
  }
  __syncthreads();
  const int h1 = min(h0 + TH, g.H) - 1, w1 = min(w0 + TW, g.W) - 1;
  const TileSpan sy = tile_span(R, g, h0, h1, w0, w1, 1), sx = tile_span(R, g, h0, h1, w0, w1, 2);
  // z shear across a warp's 4 x 8 patch must stay below the two overlapping planes
  const float shear = 3.f * fabsf(R[1] * g.sH * g.hD) + 7.f * fabsf(R[2] * g.sW * g.hD);
  const bool fits = l.kz > 0.5f && shear < 1.9f && (sy.hi - sy.lo) + fabsf(sy.r) * (float)BZ + 4.f <= (float)BY &&
                    (sx.hi - sx.lo) + fabsf(sx.r) * (float)BZ + 7.f <= (float)BX;
  if (!fits) {                                               // oblique view: plain 8-corner scatter, no staging
    float below = 0.f, Pk = 0.f;
    for (int i = i_lo; i <= i_hi; ++i) {
      const Cell c = locate(l, (float)i, g);
      const float* p = vol + c.idx;
      const Plane4 lo = load_plane(p, g.W), hi = load_plane(p + g.HW, g.W);
      const float d = lerp_planes(lo, hi, c);
      const float T = fast_exp2((St - below) * ntl2);
      Pk = fmaf(d, T, Pk);
      below += d;
      const float gk = gI * fmaf(-tau, Pk, T);
      const float g1 = gk * c.fz, g0 = gk - g1;
      const float g01 = g0 * c.fy, g00 = g0 - g01, g11 = g1 * c.fy, g10 = g1 - g11;
      const float c001 = g00 * c.fx, c011 = g01 * c.fx, c101 = g10 * c.fx, c111 = g11 * c.fx;
      float* q = g_vol + c.idx;
      atomicAdd(q, g00 - c001); atomicAdd(q + 1, c001); atomicAdd(q + g.W, g01 - c011); atomicAdd(q + g.W + 1, c011);
      q += g.HW;
      atomicAdd(q, g10 - c101); atomicAdd(q + 1, c101); atomicAdd(q + g.W, g11 - c111); atomicAdd(q + g.W + 1, c111);
    }
    return;
  }
  if (has) {
    atomicMax(&s_ztop, locate3(l, (float)i_hi, g).z0);
    atomicMin(&s_zbot, locate3(l, (float)i_lo, g).z0);
  }
  __syncthreads();
  const int ztop = s_ztop, zbot = s_zbot;
  if (ztop < zbot) return;                                   // nothing live in this tile
  const int nslab = (ztop - zbot) / ADV + 1;                 // slab s: anchors [za, za + BZ - 2], za = zbot + s ADV
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2 && s < nslab; ++s) {
      const int za = zbot + s * ADV;
      const uint32_t bar = tma::smem_u32(&bar_store[s]);
      tma::mbar_expect_tx(bar, SLAB * 4);
      tma::load_3d(tma::smem_u32(slab + s * SLAB), &map_vol, bar, span_origin(sx, za, za + BZ - 2, g.mW, true),
                   span_origin(sy, za, za + BZ - 2, g.mH, false), za);
    }
  }
  // warp-uniform depth range
  int w_lo = i_lo, w_hi = i_hi;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    w_lo = min(w_lo, __shfl_xor_sync(0xffffffffu, w_lo, o));
    w_hi = max(w_hi, __shfl_xor_sync(0xffffffffu, w_hi, o));
  }
  // first depth index of this lane whose anchor plane is >= z (anchors do not decrease with i: kz > 0.5)
  auto first_i = [&](int z) -> int {
    int i = (int)ceilf(((float)z - l.cz) / l.kz);
    i = min(max(i, 0), g.D - 1);
    while (i > 0 && locate3(l, (float)(i - 1), g).z0 >= z) --i;
    while (i < g.D - 1 && locate3(l, (float)i, g).z0 < z) ++i;
    if (locate3(l, (float)i, g).z0 < z) i = g.D;             // never reaches plane z
    return i;
  };
  float below = 0.f, Pk = 0.f;
  Plane4 pend;                                               // far-plane contributions waiting for the next sample
  pend.a = pend.b = pend.c = pend.d = 0.f;
  int pend_idx = -1;
  int i_beg = w_lo;
  const bool xr = (lane & 7) != 7, yr = lane < 24;           // has a right / lower neighbour inside the warp's patch
  for (int s = 0; s < nslab; ++s) {
    const int buf = s & 1;
    const int za = zbot + s * ADV;
    const int oy = span_origin(sy, za, za + BZ - 2, g.mH, false), ox = span_origin(sx, za, za + BZ - 2, g.mW, true);
    int i_end = w_hi + 1;
    if (s + 1 < nslab) {
      int f = has ? first_i(za + ADV) : 0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) f = max(f, __shfl_xor_sync(0xffffffffu, f, o));
      i_end = min(i_end, f);
    }
    tma::mbar_wait(tma::smem_u32(&bar_store[buf]), (uint32_t)((s >> 1) & 1));
    const float* sb = slab + buf * SLAB;
    const int base = -((za * BY + oy) * BX + ox);
    for (int i = i_beg; i < i_end; ++i) {
      const bool live = i >= i_lo && i <= i_hi;
      const Anchor a = locate3(l, (float)min(i, g.D - 1), g);
      const int gidx = (a.z0 * g.H + a.y0) * g.W + a.x0;
      float gk = 0.f;
      if (live) {
        const int ry = a.y0 - oy, rx = a.x0 - ox, rz = a.z0 - za;
        Plane4 lo, hi;
        if ((unsigned)ry < (unsigned)(BY - 1) && (unsigned)rx < (unsigned)(BX - 1) && (unsigned)rz < (unsigned)(BZ - 1)) {
          const float* p = sb + ((a.z0 * BY + a.y0) * BX + a.x0 + base);
          lo.a = p[0]; lo.b = p[1]; lo.c = p[BX]; lo.d = p[BX + 1];
          hi.a = p[BY * BX]; hi.b = p[BY * BX + 1]; hi.c = p[BY * BX + BX]; hi.d = p[BY * BX + BX + 1];
        } else {                                             // clamped onto a face / rim of the box: gather
          const float* p = vol + gidx;
          lo = load_plane(p, g.W); hi = load_plane(p + g.HW, g.W);
        }
        Cell c; c.idx = 0; c.fz = a.fz; c.fy = a.fy; c.fx = a.fx;
        const float d = lerp_planes(lo, hi, c);
        const float T = fast_exp2((St - below) * ntl2);
        Pk = fmaf(d, T, Pk);
        below += d;
        gk = gI * fmaf(-tau, Pk, T);
      }
      const float g1 = gk * a.fz, g0 = gk - g1;
      const float g01 = g0 * a.fy, g00 = g0 - g01, g11 = g1 * a.fy, g10 = g1 - g11;
      float n01 = g00 * a.fx, n11 = g01 * a.fx;              // near plane: (y0,x1), (y1,x1)
      float n00 = g00 - n01, n10 = g01 - n11;                //             (y0,x0), (y1,x0)
      const float f01 = g10 * a.fx, f11 = g11 * a.fx;        // far plane
      const float f00 = g10 - f01, f10 = g11 - f11;
      // depth: merge the waiting plane into this sample's near plane, or write it out
      if (pend_idx >= 0) {
        if (live && pend_idx == gidx) {
          n00 += pend.a; n01 += pend.b; n10 += pend.c; n11 += pend.d;
        } else {
          float* f = g_vol + pend_idx;
          atomicAdd(f, pend.a); atomicAdd(f + 1, pend.b); atomicAdd(f + g.W, pend.c); atomicAdd(f + g.W + 1, pend.d);
        }
        pend_idx = -1;
      }
      // x: hand the x1 column to the right neighbour when its anchor is one voxel to the right
      const int my = live ? gidx : -0x40000000;
      const int nbx = __shfl_down_sync(0xffffffffu, my, 1);
      const bool give_x = xr && live && nbx == my + 1;
      const float rx0 = __shfl_up_sync(0xffffffffu, give_x ? n01 : 0.f, 1);
      const float rx1 = __shfl_up_sync(0xffffffffu, give_x ? n11 : 0.f, 1);
      if ((lane & 7) != 0) { n00 += rx0; n10 += rx1; }
      // y: hand the (y1,x0) value to the lane one row down when its anchor is one voxel down
      const int nby = __shfl_down_sync(0xffffffffu, my, 8);
      const bool give_y = yr && live && nby == my + g.W;
      const float ry0 = __shfl_up_sync(0xffffffffu, give_y ? n10 : 0.f, 8);
      if (lane >= 8) n00 += ry0;
      if (live) {
        float* p = g_vol + gidx;
        atomicAdd(p, n00);
        if (!give_y) atomicAdd(p + g.W, n10);
        if (!give_x) { atomicAdd(p + 1, n01); atomicAdd(p + g.W + 1, n11); }
        pend.a = f00; pend.b = f01; pend.c = f10; pend.d = f11;
        pend_idx = gidx + g.HW;
      }
    }
    i_beg = max(i_beg, i_end);
    __syncthreads();                                         // every warp is done with this buffer
    if (threadIdx.x == 0 && s + 2 < nslab) {
      const int zn = zbot + (s + 2) * ADV;
      const uint32_t bar = tma::smem_u32(&bar_store[buf]);
      tma::mbar_expect_tx(bar, SLAB * 4);
      tma::load_3d(tma::smem_u32(slab + buf * SLAB), &map_vol, bar, span_origin(sx, zn, zn + BZ - 2, g.mW, true),
                   span_origin(sy, zn, zn + BZ - 2, g.mH, false), zn);
    }
  }
  if (pend_idx >= 0) {
    float* f = g_vol + pend_idx;
    atomicAdd(f, pend.a); atomicAdd(f + 1, pend.b); atomicAdd(f + g.W, pend.c); atomicAdd(f + g.W + 1, pend.d);
  }
}

static int rm_slab_planes = 12;   // tuning switch (microbenchmarks): planes per TMA slab of the ray-march kernels, 8 / 12 / 16
                                  // (C3 on the B200, exact intervals: forward 78 / 68 / 70 us; the gather kernel 81 us)
extern "C" int lnst_set_raymarch_slab(int32_t planes) {
  if (planes != 8 && planes != 12 && planes != 16) return LNST_EARG;
  rm_slab_planes = planes;
  return LNST_OK;
}

template <int BZ>
static int launch_rm_fwd(const CUtensorMap& mv, const float* vol, const float* rot, int n_views, const RayGeo& g,
                         const BoxF& bf, const int2* iv, float ntl2, int liquid, float* img, float* stot, cudaStream_t st,
                         float* stats) {
  using namespace rm;
  const int tiles_h = (g.H + TH - 1) / TH, tiles_w = (g.W + TW - 1) / TW;
  const int smem = 2 * BZ * BY * BX * 4 + 128;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(raymarch_fwd_tma_k<BZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  raymarch_fwd_tma_k<BZ><<<dim3((unsigned)(tiles_h * tiles_w), (unsigned)n_views), THREADS, smem, st>>>(
      mv, vol, rot, g, bf, iv, ntl2, liquid, img, stot, tiles_w, stats);
  return (int)cudaGetLastError();
}

extern "C" int lnst_raymarch_fwd_tma(const float* vol, const float* rot, int32_t n_views, int32_t D, int32_t H,
                                     int32_t W, float tau, int32_t liquid, const LnstBox* box, const int32_t* intervals,
                                     float* img, float* stot, void* stream) {
  return lnst_raymarch_fwd_max_tma(vol, rot, n_views, D, H, W, tau, liquid, box, intervals, img, stot, nullptr, stream);
}

// The same march; stats[2 v] = max over view v's pixels comes out of the kernel's last instructions (stats must be zero on
// entry; lnst_image_max's first pass and its memset disappear).  stats may be NULL.
extern "C" int lnst_raymarch_fwd_max_tma(const float* vol, const float* rot, int32_t n_views, int32_t D, int32_t H,
                                         int32_t W, float tau, int32_t liquid, const LnstBox* box,
                                         const int32_t* intervals, float* img, float* stot, float* stats, void* stream) {
  using namespace rm;
  if (!vol || !rot || !img || !stot || n_views < 1 || D < 2 || H < 2 || W < 2 || !box_ok(box, D, H, W)) return LNST_EARG;
  if ((int64_t)D * H * W >= 0x7fffffff) return LNST_EARG;
  const int bz = rm_slab_planes;
  CUtensorMap mv;
  if (!tma::make_volume_map(&mv, vol, D, H, W, bz, BY, BX)) return LNST_EARG;
  const RayGeo g = make_geo(D, H, W);
  const BoxF bf = make_boxf(box, D, H, W);
  const int2* iv = reinterpret_cast<const int2*>(intervals);
  const float ntl2 = -tau * 1.4426950408889634f;
  cudaStream_t st = lnst_stream(stream);
  if (bz == 8) return launch_rm_fwd<8>(mv, vol, rot, n_views, g, bf, iv, ntl2, (int)liquid, img, stot, st, stats);
  if (bz == 12) return launch_rm_fwd<12>(mv, vol, rot, n_views, g, bf, iv, ntl2, (int)liquid, img, stot, st, stats);
  return launch_rm_fwd<16>(mv, vol, rot, n_views, g, bf, iv, ntl2, (int)liquid, img, stot, st, stats);
}

template <int BZ>
static int launch_rm_bwd(const CUtensorMap& mv, const float* vol, const float* rot, int n_views, const RayGeo& g,
                         const BoxF& bf, const int2* iv, float tau, const float* stot, const float* g_img, float* g_vol,
                         cudaStream_t st) {
  using namespace rm;
  const int tiles_h = (g.H + TH - 1) / TH, tiles_w = (g.W + TW - 1) / TW;
  const int smem = 2 * BZ * BY * BX * 4 + 128;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(raymarch_bwd_tma_k<BZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  raymarch_bwd_tma_k<BZ><<<dim3((unsigned)(tiles_h * tiles_w), (unsigned)n_views), THREADS, smem, st>>>(
      mv, vol, rot, g, bf, iv, tau, -tau * 1.4426950408889634f, stot, g_img, g_vol, tiles_w);
  return (int)cudaGetLastError();
}

// smoke render only (the liquid render's gradient does not depend on the samples: lnst_raymarch_bwd_box handles it)
extern "C" int lnst_raymarch_bwd_tma(const float* vol, const float* rot, int32_t n_views, int32_t D, int32_t H,
                                     int32_t W, float tau, const LnstBox* box, const int32_t* intervals,
                                     const float* stot, const float* g_img, float* g_vol, void* stream) {
  using namespace rm;
  if (!vol || !rot || !stot || !g_img || !g_vol || n_views < 1 || D < 2 || H < 2 || W < 2 || !box_ok(box, D, H, W))
    return LNST_EARG;
  if ((int64_t)D * H * W >= 0x3fffffff) return LNST_EARG;
  const int bz = rm_slab_planes;
  CUtensorMap mv;
  if (!tma::make_volume_map(&mv, vol, D, H, W, bz, BY, BX)) return LNST_EARG;
  const RayGeo g = make_geo(D, H, W);
  const BoxF bf = make_boxf(box, D, H, W);
  const int2* iv = reinterpret_cast<const int2*>(intervals);
  cudaStream_t st = lnst_stream(stream);
  if (bz == 8) return launch_rm_bwd<8>(mv, vol, rot, n_views, g, bf, iv, tau, stot, g_img, g_vol, st);
  if (bz == 12) return launch_rm_bwd<12>(mv, vol, rot, n_views, g, bf, iv, tau, stot, g_img, g_vol, st);
  return launch_rm_bwd<16>(mv, vol, rot, n_views, g, bf, iv, tau, stot, g_img, g_vol, st);
}

// =====================================================================================================================
// Semi-Lagrangian advection, order 1 (transform.py:557-609), 3-D scalar field: out(p) = d(p - v(p)), trilinear, edge
// clamped.  A CTA owns an 8 x 8 x 32 tile of outputs; the source box of the tile -- the tile grown by `reach` cells,
// the caller's bound on the back-trace length -- is ONE TMA box, and the 8 corners of every back-traced point come out
// of shared memory.  Points whose back-trace leaves the staged box (|v| above the bound) gather from global memory, so
// the result never depends on the bound.  Same arithmetic per output as advect_k (optim.cu): bit-identical fields.
// =====================================================================================================================
namespace adv {
constexpr int TZ = 8, TY = 8, TX = 32;
constexpr int THREADS = TY * TX;
}

template <int R>
__global__ void __launch_bounds__(adv::THREADS) advect3_tma_k(const __grid_constant__ CUtensorMap map_d,
                                                              const float* __restrict__ d, const float* __restrict__ vel,
                                                              int D, int H, int W, float sD, float sH, float sW,
                                                              float* __restrict__ out, int tiles_y, int tiles_x) {
  using namespace adv;
  constexpr int BZ = TZ + 2 * R + 1, BY = TY + 2 * R + 1, BX = ((TX + 2 * R + 1 + 3) + 3) & ~3;
  extern __shared__ unsigned char smem_raw[];
  float* box = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  __shared__ __align__(8) uint64_t bar_store;
  const uint32_t bar = tma::smem_u32(&bar_store);
  const int t = blockIdx.x;
  const int tx_ = t % tiles_x, ty_ = (t / tiles_x) % tiles_y, tz_ = t / (tiles_x * tiles_y);
  const int z0 = tz_ * TZ, y0 = ty_ * TY, x0 = tx_ * TX;
  const int oz = z0 - R, oy = y0 - R, ox = (x0 - R) & ~3;    // x start: a multiple of 4 floats (tma_tiles.cuh)
  if (threadIdx.x == 0) {
    tma::mbar_init(bar, 1);
    tma::fence_mbar_init();
    tma::mbar_expect_tx(bar, BZ * BY * BX * 4);
    tma::load_3d(tma::smem_u32(box), &map_d, bar, ox, oy, oz);
  }
  __syncthreads();
  const int ly = threadIdx.x >> 5, lx = threadIdx.x & 31;
  const int y = y0 + ly, x = x0 + lx;
  const bool col = y < H && x < W;
  // velocities of the column first (independent of the staged box): TZ x 3 loads in flight under the TMA copy
  float vz[TZ], vy[TZ], vx[TZ];
#pragma unroll
  for (int k = 0; k < TZ; ++k) {
    const int z = z0 + k;
    vz[k] = vy[k] = vx[k] = 0.f;
    if (col && z < D) {
      const int64_t c = ((int64_t)z * H + y) * W + x;
      vz[k] = vel[c * 3]; vy[k] = vel[c * 3 + 1]; vx[k] = vel[c * 3 + 2];
    }
  }
  tma::mbar_wait(bar, 0);
  const float mD = (float)D - 1.f, mH = (float)H - 1.f, mW = (float)W - 1.f;
#pragma unroll
  for (int k = 0; k < TZ; ++k) {
    const int z = z0 + k;
    if (!(col && z < D)) continue;
    // advect_k's arithmetic: g = linspace(-1,1)[idx] - v;  x = (g + 1) (n - 1) / 2;  floor / floor+1 clamped separately
    const float gz = __fadd_rn(-1.f, __fmul_rn(sD, (float)z)) - vz[k];
    const float gy = __fadd_rn(-1.f, __fmul_rn(sH, (float)y)) - vy[k];
    const float gx = __fadd_rn(-1.f, __fmul_rn(sW, (float)x)) - vx[k];
    const float pz = (gz + 1.f) * mD * 0.5f, py = (gy + 1.f) * mH * 0.5f, px = (gx + 1.f) * mW * 0.5f;
    const int fz = (int)floorf(pz), fy = (int)floorf(py), fx = (int)floorf(px);
    const int lz = min(max(fz, 0), D - 1), hz = min(max(fz + 1, 0), D - 1);
    const int lyy = min(max(fy, 0), H - 1), hy = min(max(fy + 1, 0), H - 1);
    const int lxx = min(max(fx, 0), W - 1), hx = min(max(fx + 1, 0), W - 1);
    const float wz = pz - (float)lz, wy = py - (float)lyy, wx = px - (float)lxx;
    float v[8];
    const bool inside = lz >= oz && hz < oz + BZ && lyy >= oy && hy < oy + BY && lxx >= ox && hx < ox + BX;
    if (inside) {
      const float* b = box - ((oz * BY + oy) * BX + ox);
      v[0] = b[(lz * BY + lyy) * BX + lxx]; v[1] = b[(lz * BY + lyy) * BX + hx];
      v[2] = b[(lz * BY + hy) * BX + lxx];  v[3] = b[(lz * BY + hy) * BX + hx];
      v[4] = b[(hz * BY + lyy) * BX + lxx]; v[5] = b[(hz * BY + lyy) * BX + hx];
      v[6] = b[(hz * BY + hy) * BX + lxx];  v[7] = b[(hz * BY + hy) * BX + hx];
    } else {
      v[0] = d[((int64_t)lz * H + lyy) * W + lxx]; v[1] = d[((int64_t)lz * H + lyy) * W + hx];
      v[2] = d[((int64_t)lz * H + hy) * W + lxx];  v[3] = d[((int64_t)lz * H + hy) * W + hx];
      v[4] = d[((int64_t)hz * H + lyy) * W + lxx]; v[5] = d[((int64_t)hz * H + lyy) * W + hx];
      v[6] = d[((int64_t)hz * H + hy) * W + lxx];  v[7] = d[((int64_t)hz * H + hy) * W + hx];
    }
    // corner order and weight products of advect_k: corner bits (z, y, x), w = wz' * wy' * wx', o += w * value
    float o = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float w = 1.f;
      w *= (c & 4) ? wz : (1.f - wz);
      w *= (c & 2) ? wy : (1.f - wy);
      w *= (c & 1) ? wx : (1.f - wx);
      o += w * v[c];
    }
    out[((int64_t)z * H + y) * W + x] = o;
  }
}

template <int R>
static int launch_advect(const float* d, const float* vel, int D, int H, int W, float* out, cudaStream_t st) {
  using namespace adv;
  constexpr int BZ = TZ + 2 * R + 1, BY = TY + 2 * R + 1, BX = ((TX + 2 * R + 1 + 3) + 3) & ~3;
  CUtensorMap md;
  if (!tma::make_volume_map(&md, d, D, H, W, BZ, BY, BX)) return LNST_EARG;
  const int smem = BZ * BY * BX * 4 + 128;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(advect3_tma_k<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const int tz = (D + TZ - 1) / TZ, ty = (H + TY - 1) / TY, tx = (W + TX - 1) / TX;
  advect3_tma_k<R><<<(unsigned)(tz * ty * tx), THREADS, smem, st>>>(
      md, d, vel, D, H, W, D > 1 ? 2.0f / (float)(D - 1) : 0.f, H > 1 ? 2.0f / (float)(H - 1) : 0.f,
      W > 1 ? 2.0f / (float)(W - 1) : 0.f, out, ty, tx);
  return (int)cudaGetLastError();
}

// d [D,H,W] scalar field, vel [D,H,W,3] in normalised units per step (channel i <-> axis i), out [D,H,W].
// reach: bound on the back-trace length in cells the staged box is sized for (1..4); longer back-traces still give the
// exact result (global gathers).
extern "C" int lnst_advect3_tma(const float* d, const float* vel, int32_t D, int32_t H, int32_t W, int32_t reach,
                                float* out, void* stream) {
  if (!d || !vel || !out || D < 1 || H < 1 || W < 1 || (int64_t)D * H * W >= 0x7fffffff) return LNST_EARG;
  cudaStream_t st = lnst_stream(stream);
  if (reach <= 1) return launch_advect<1>(d, vel, D, H, W, out, st);
  if (reach == 2) return launch_advect<2>(d, vel, D, H, W, out, st);
  if (reach == 3) return launch_advect<3>(d, vel, D, H, W, out, st);
  return launch_advect<4>(d, vel, D, H, W, out, st);
}

// =====================================================================================================================
// Data gradient of conv1_1 w.r.t. a GRAY render (styler_base.py:41-43 folds the RGB replication into the weights):
// g_gray[p] = sum_tap sum_ch g[p + tap][ch] * wg[tap][ch], zero padded -- 64 (or 2 x 64 split) bf16 channels in, ONE fp32
// channel out.  On the tensor cores this is an N = 16 MMA whose time is the A operand's shared-memory reads (0.064 ms at
// C3 in bf16x3); here a CTA stages the 18 x 18 pixel patch of its 16 x 16 tile with ONE TMA box (out-of-image pixels are
// zero filled = the SAME padding) and a thread sums its pixel's 9 x 64 products on the CUDA cores, fp32 weights.
// Bank conflicts: a pixel row is 128 / 256 B, so the lanes of a quarter-warp read the 16-byte chunk (j + lane) % NCH of
// their pixel in step j -- 8 distinct bank groups -- and fetch the matching weights.
// =====================================================================================================================
namespace cfb {
constexpr int T = 16;                                        // tile of T x T pixels, patch (T+2)^2
constexpr int THREADS = T * T;
}

template <int SPLIT>
__global__ void __launch_bounds__(cfb::THREADS) conv_first_bwd_gray_direct_k(const __grid_constant__ CUtensorMap map_g,
                                                                            const float* __restrict__ wg,
                                                                            float* __restrict__ g_gray, int H, int W,
                                                                            int tiles_y, int tiles_x) {
  using namespace cfb;
  constexpr int C2 = SPLIT ? 128 : 64;                       // bf16 elements per pixel row
  constexpr int NCH = C2 / 8;                                // 16-byte chunks per pixel row
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  const uint4* patch = reinterpret_cast<const uint4*>(base);                // [(T+2)*(T+2)][NCH]
  float* sw = reinterpret_cast<float*>(base + (T + 2) * (T + 2) * C2 * 2);   // [9][64]
  __shared__ __align__(8) uint64_t bar_store;
  const uint32_t bar = tma::smem_u32(&bar_store);
  const int t = blockIdx.x;
  const int tx_ = t % tiles_x, ty_ = (t / tiles_x) % tiles_y, img = t / (tiles_x * tiles_y);
  const int y0 = ty_ * T, x0 = tx_ * T;
  if (threadIdx.x == 0) {
    tma::mbar_init(bar, 1);
    tma::fence_mbar_init();
    tma::mbar_expect_tx(bar, (T + 2) * (T + 2) * C2 * 2);
    tma::load_4d(tma::smem_u32(base), &map_g, bar, 0, x0 - 1, y0 - 1, img);
  }
  for (int i = threadIdx.x; i < 9 * 64; i += THREADS) sw[i] = wg[i];
  __syncthreads();
  tma::mbar_wait(bar, 0);
  const int ly = threadIdx.x / T, lx = threadIdx.x % T;
  const int rot = threadIdx.x & 7;                           // chunk rotation of this lane within its quarter-warp
  float acc = 0.f;
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int dy = tap / 3, dx = tap - 3 * dy;
    const uint4* row = patch + ((ly + dy) * (T + 2) + lx + dx) * NCH;
    const float* w = sw + tap * 64;
#pragma unroll
    for (int j = 0; j < 8; ++j) {                            // 8 chunks of 8 channels
      const int c = (j + rot) & 7;
      const float4 w0 = *reinterpret_cast<const float4*>(w + c * 8), w1 = *reinterpret_cast<const float4*>(w + c * 8 + 4);
      const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
      const uint4 hi = row[c];
      const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&hi);
      float v[8];
#pragma unroll
      for (int e = 0; e < 4; ++e) { const float2 f = __bfloat1622float2(hh[e]); v[2 * e] = f.x; v[2 * e + 1] = f.y; }
      if (SPLIT) {
        const uint4 lo = row[8 + c];
        const __nv_bfloat162* ll = reinterpret_cast<const __nv_bfloat162*>(&lo);
#pragma unroll
        for (int e = 0; e < 4; ++e) { const float2 f = __bfloat1622float2(ll[e]); v[2 * e] += f.x; v[2 * e + 1] += f.y; }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) acc = fmaf(v[e], wv[e], acc);
    }
  }
  const int y = y0 + ly, x = x0 + lx;
  if (y < H && x < W) g_gray[((int64_t)img * H + y) * W + x] = acc;
}

// g bf16 [n,H,W,64] (split = 0) or [n,H,W,128] = [hi | lo] (split = 1); wg fp32 [9,64] = the data-gradient weights of
// conv1_1 summed over the three input channels, times the net-input scale; g_gray fp32 [n,H,W]
extern "C" int lnst_conv_first_bwd_gray_direct(const void* g, int32_t split, const float* wg, float* g_gray, int32_t n,
                                               int32_t H, int32_t W, void* stream) {
  using namespace cfb;
  if (!g || !wg || !g_gray || n < 1 || H < 1 || W < 1) return LNST_EARG;
  const int C2 = split ? 128 : 64;
  CUtensorMap mg;
  if (!tma::make_nhwc_bf16_map(&mg, g, n, H, W, C2, T + 2, T + 2)) return LNST_EARG;
  const int ty = (H + T - 1) / T, tx = (W + T - 1) / T;
  const int smem = (T + 2) * (T + 2) * C2 * 2 + 9 * 64 * 4 + 128;
  cudaStream_t st = lnst_stream(stream);
  if (split) {
    static bool configured = false;
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(conv_first_bwd_gray_direct_k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e != cudaSuccess) return (int)e;
      configured = true;
    }
    conv_first_bwd_gray_direct_k<1><<<(unsigned)(n * ty * tx), THREADS, smem, st>>>(mg, wg, g_gray, H, W, ty, tx);
  } else {
    static bool configured = false;
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(conv_first_bwd_gray_direct_k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e != cudaSuccess) return (int)e;
      configured = true;
    }
    conv_first_bwd_gray_direct_k<0><<<(unsigned)(n * ty * tx), THREADS, smem, st>>>(mg, wg, g_gray, H, W, ty, tx);
  }
  return (int)cudaGetLastError();
}
