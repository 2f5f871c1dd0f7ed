/* lnst_b200.h -- C-ABI of the B200-native LNST stylisation hot path.
 *
 * The reference (byungsook/neural-flow-style) has NO native code and no FFI: its hot path is
 * a TensorFlow-1.15 graph executed by `sess.run([train_op, total_loss])`
 * (styler_3p.py:331,354; styler_2p.py:257).  Each entry point below replaces the stock TF
 * op(s) cited beside it.  The library is what a maintainer would bind in place of those ops
 * (ctypes stub shown in INTEGRATION.md); `neural-flow-style_b200/lnst/_lib.py` is that binding.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer (fp32 unless stated), owned by the caller; the
 *     library allocates nothing, keeps no global state, never synchronises the stream;
 *   - `stream` is a cudaStream_t passed as void*;
 *   - return 0 = ok; <0 = argument error (nothing launched); >0 = cudaError_t of the launch;
 *   - volumes are contiguous [D,H,W] (one frame, one channel), images NHWC, particles AoS
 *     [N,dim] in the reference's normalised ((z,)y,x) order, weights HWIO (slim layout).
 */
#ifndef LNST_B200_H
#define LNST_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define LNST_ABI_VERSION 1

/* Grid description shared by the particle<->grid ops (transform.py:1316-1343). */
typedef struct LnstGrid {
  int32_t dim;        /* 2 or 3 */
  int32_t res[3];     /* D,H,W (dim==2: res[0] = 1, positions are (y,x)) */
  float domain[3];    /* domain size in the same order (dim==2: domain[0] unused) */
  float cell;         /* float32(domain[first]) / float32(res[first]) (transform.py:1328-1330) */
  int32_t nsize;      /* neighbour half-width: (2 nsize+1)^dim target cells */
  int32_t clip;       /* 1: clamp positions to [0,domain-1e-6]; 0: drop outliers (:1320-1325) */
} LnstGrid;

/* Active region of a volume: inclusive voxel bounds in the stored [D,H,W] layout.  The `_box` entry
 * points below take one (NULL = whole volume) when the caller knows that the density is exactly zero
 * outside it and that no gradient is needed there -- in density mode the particle positions are
 * constants (styler_3p.py:60-76), so the bounding box of sum-of-weights > 0, grown by one voxel for
 * the 3x3x3 blur, is fixed for a whole run.  Results inside the box are identical to the full-volume
 * call; nothing outside the box is read or written (the caller keeps those voxels zero). */
typedef struct LnstBox {
  int32_t lo[3];
  int32_t hi[3];
} LnstBox;

int lnst_abi_version(void);
/* Human-readable library version (static storage). */
const char* lnst_version(void);
/* Bytes of caller-provided scratch one call of entry point `op` (its name without the lnst_ prefix) needs beyond its
 * inputs and outputs, for the problem size in `dims` (the library allocates nothing).  dims per op:
 *   splat_wavg_fwd {V, nk} (the num volume); raymarch_fwd / raymarch_bwd {P, n_views} (ray intervals); image_max {n_img};
 *   normalize_bwd {n_img}; density_reg {}; adam_step_dev / adam_iterate_dev {} (the 3-float state);
 *   gram_diff_bf16_tc {n_img, C}.  Ops that need none return 0; an unknown op returns -1. */
int64_t lnst_workspace_bytes(const char* op, const int64_t* dims, int32_t n_dims);

/* ---- particle -> grid (transform.py:1310-1453 p2g, :1577-1704 p2g_wavg) ------------------ */
/* SPH splat: out[cell] += scale * W(|x_p - x_cell|/h).  `disp` (may be NULL) is added to p
 * (styler_3p.py:58).  With colours (pc != NULL, C channels, out is [cells,C]):
 * out += scale*W*pc/pd (pd NULL => divide by rest_density).  `out` must be zeroed by the
 * caller.  H axis is written flipped (:1404,:1452). */
int lnst_splat_sph_fwd(const float* p, const float* disp, int64_t n, const LnstGrid* g, float h,
                       float scale, const float* pc, const float* pd, int32_t C, float rest_density,
                       float* out, void* stream);
/* d loss / d p (normalised units) for the scalar SPH splat; g_p [n,dim] is overwritten. */
int lnst_splat_sph_bwd_pos(const float* p, const float* disp, int64_t n, const LnstGrid* g, float h,
                           float scale, const float* g_out, float* g_p, void* stream);
/* d loss / d pc for the colour splat (styler_2p.py:74-75); g_pc [n,C] is overwritten. */
int lnst_splat_sph_bwd_color(const float* p, int64_t n, const LnstGrid* g, float h, float scale,
                             const float* pd, int32_t C, float rest_density, const float* g_out,
                             float* g_pc, void* stream);
/* Weighted-average splat, nk kernels with support radii h[k] (styler_3p.py:79-87).
 * wmap [nk,cells]: sum of weights (positions are constant in density mode => computed once). */
int lnst_splat_wavg_wmap(const float* p, int64_t n, const LnstGrid* g, const float* h, int32_t nk,
                         float* wmap, void* stream);
/* out[cell] = sum_k where(wmap_k>1e-6, num_k/wmap_k, num_k), num_k = sum W_k (r_k + clip(var_k,-1,1)).
 * `num` [nk,cells] is workspace (zeroed inside). */
int lnst_splat_wavg_fwd(const float* p, const float* r, const float* var, int64_t n, const LnstGrid* g,
                        const float* h, int32_t nk, const float* wmap, float* num, float* out,
                        void* stream);
/* Box variant: only cells inside `box` are combined; `num` must be zero on entry and is zero again on
 * exit (the combine pass clears what it reads), so no per-step memset of the workspace is needed. */
int lnst_splat_wavg_fwd_box(const float* p, const float* r, const float* var, int64_t n, const LnstGrid* g,
                            const float* h, int32_t nk, const float* wmap, float* num, float* out,
                            const LnstBox* box, void* stream);
/* g_var [n,nk] overwritten; reproduces TF's where/div NaN rule (transform.py:1703). */
int lnst_splat_wavg_bwd(const float* p, const float* var, int64_t n, const LnstGrid* g, const float* h,
                        int32_t nk, const float* wmap, const float* g_out, float* g_var, void* stream);

/* Density-mode fast path (3-D, nsize = 1): coef [nk,cells] = d out/d num per cell (1/wmap, 1, or NaN where
 * wmap == 0 -- TF's where/div rule), computed once per (frame, octave) like wmap; the gradient kernel
 * then needs no division.  Same result as lnst_splat_wavg_bwd. */
int lnst_splat_wavg_coef(const float* wmap, int32_t nk, int64_t cells, float* coef, void* stream);
/* Home cell of every particle (linear index over the unflipped [D,H,W] grid; -1: outside the domain / padding row; -2:
 * valid but rounded onto the far face) and its offset from that cell's centre [n,3]: input of the per-cell particle
 * lists the gather splat (lnst_splat_wavg_fwd_gather) walks.  3-D grids. */
int lnst_splat_cells(const float* p, int64_t n, const LnstGrid* g, int32_t* cell, float* rel, void* stream);
int lnst_splat_wavg_bwd_coef(const float* p, const float* var, int64_t n, const LnstGrid* g, const float* h,
                             int32_t nk, const float* coef, const float* g_out, float* g_var, void* stream);

/* ---- field post-processing (styler_3p.py:112-125: conv3d [1,k,1]^3/sum SAME + max(d,0)) -- */
/* out = relu(smooth(in)); cells whose pre-activation is < 0 are stored as -0.0f so that the
 * backward pass can apply TF's maximum() gradient rule (passes at equality) without a mask. */
int lnst_smooth3_relu_fwd(const float* in, float* out, int32_t D, int32_t H, int32_t W, int32_t k,
                          void* stream);
int lnst_smooth3_relu_bwd(const float* g_out, const float* out, float* g_in, int32_t D, int32_t H,
                          int32_t W, int32_t k, void* stream);

int lnst_smooth3_relu_fwd_box(const float* in, float* out, int32_t D, int32_t H, int32_t W, int32_t k,
                              const LnstBox* box, void* stream);
int lnst_smooth3_relu_bwd_box(const float* g_out, const float* out, float* g_in, int32_t D, int32_t H,
                              int32_t W, int32_t k, const LnstBox* box, void* stream);
/* vol[box] = value */
int lnst_fill_box(float* vol, int32_t D, int32_t H, int32_t W, const LnstBox* box, float value, void* stream);

/* ---- rotate + render (transform.py:611-628,343-433; styler_3p.py:148-158) ---------------- */
/* rot: [n_views,9] row-major rotation matrices, or NULL for the unrotated render. */
int lnst_rotate_fwd(const float* vol, const float* rot, int32_t n_views, int32_t D, int32_t H,
                    int32_t W, float* out, void* stream);
/* Gradient of lnst_rotate_fwd w.r.t. the volume: g_vol [D,H,W] += 8-corner scatter of g_out [n_views,D,H,W]
 * (accumulates over the views: zero it first). */
int lnst_rotate_bwd(const float* g_out, const float* rot, int32_t n_views, int32_t D, int32_t H,
                    int32_t W, float* g_vol, void* stream);
/* Fused rotate + emission/absorption ray-march.  img [n_views,H,W] = sum_i d_i T_i (smoke) or
 * 1-exp(-tau sum d) (liquid); stot [n_views,H,W] = sum_i d_i (saved for the backward). */
int lnst_raymarch_fwd(const float* vol, const float* rot, int32_t n_views, int32_t D, int32_t H,
                      int32_t W, float tau, int32_t liquid, float* img, float* stot, void* stream);
/* g_vol [D,H,W] += d loss/d vol from g_img [n_views,H,W] (accumulates: zero it first). */
int lnst_raymarch_bwd(const float* vol, const float* rot, int32_t n_views, int32_t D, int32_t H,
                      int32_t W, float tau, int32_t liquid, const float* stot, const float* g_img,
                      float* g_vol, void* stream);

/* Box variants: rays are marched only through the depth interval whose trilinear footprints can touch
 * the box (density is zero elsewhere, so img/stot are bit-identical to the full march; g_vol receives
 * the exact gradient for every ACTIVE voxel inside the box and nothing outside it).
 * `intervals` (may be NULL): int32 [n_views,H,W,2] = the inclusive depth-index range of every ray, from
 * lnst_ray_intervals; NULL = derive it from the box inside the kernel. */
int lnst_raymarch_fwd_box(const float* vol, const float* rot, int32_t n_views, int32_t D, int32_t H,
                          int32_t W, float tau, int32_t liquid, const LnstBox* box, const int32_t* intervals,
                          float* img, float* stot, void* stream);
int lnst_raymarch_bwd_box(const float* vol, const float* rot, int32_t n_views, int32_t D, int32_t H,
                          int32_t W, float tau, int32_t liquid, const LnstBox* box, const int32_t* intervals,
                          const float* stot, const float* g_img, float* g_vol, void* stream);
/* Fused image glue for the smoke render (styler_3p.py:158 `d /= tf.reduce_max(d)`): the forward march also reduces
 * stats[2*v] = max of view v (stats zero on entry; lnst_image_max's first pass), and the backward march takes the cotangent
 * of the NORMALISED image and applies lnst_normalize_bwd's second pass while loading it (img = the un-normalised render,
 * stats = {max, ties} per view, dots[v] = sum_p g_gray[v,p] * img[v,p]).  With stats == NULL they are the calls above.
 * Same results as the separate calls. */
int lnst_raymarch_fwd_max_box(const float* vol, const float* rot, int32_t n_views, int32_t D, int32_t H, int32_t W,
                              float tau, int32_t liquid, const LnstBox* box, const int32_t* intervals, float* img,
                              float* stot, float* stats, void* stream);
int lnst_raymarch_bwd_norm_box(const float* vol, const float* rot, int32_t n_views, int32_t D, int32_t H, int32_t W,
                               float tau, int32_t liquid, const LnstBox* box, const int32_t* intervals, const float* stot,
                               const float* g_gray, const float* img, const float* stats, const float* dots, float* g_vol,
                               void* stream);
/* Ray intervals for a set of views: the slab test against `box`, then shrunk from both ends while the ray
 * runs through empty occupancy bricks.  `bricks` (may be NULL): one byte per 4x4x4-voxel brick,
 * [ceil(D/4), ceil(H/4), ceil(W/4)], non-zero where the brick or anything within two voxels plus one brick
 * of it is active.  Depends on the views, not on the density: run it once per view set.  Only zeros are
 * skipped, so images stay bit-identical and every ACTIVE voxel still receives its exact gradient. */
int lnst_ray_intervals(const float* rot, int32_t n_views, int32_t D, int32_t H, int32_t W, const LnstBox* box,
                       const unsigned char* bricks, int32_t* intervals, void* stream);
/* Exact variant for view sets that stay fixed: touch [D,H,W] bytes, non-zero for an anchor voxel v when any voxel of
 * {v, v+1}^3 can hold smoothed density.  Same output layout; about the cost of one forward march. */
int lnst_ray_intervals_exact(const float* rot, int32_t n_views, int32_t D, int32_t H, int32_t W, const LnstBox* box,
                             const unsigned char* touch, int32_t* intervals, void* stream);
/* Tuning switch for lnst_raymarch_bwd (process-wide, default 1): 1 = neighbouring lanes merge their
 * shared x-corner contributions by warp shuffle before the atomics; 0 = eight atomics per sample. */
int lnst_set_raymarch_merge(int32_t on);

/* ---- image glue (styler_3p.py:158; styler_base.py:33-45; vgg.py:50-53) -------------------- */
/* stats[2*v+0] = max over image v, stats[2*v+1] = number of pixels attaining it. */
int lnst_image_max(const float* img, int32_t n_img, int64_t n_pix, float* stats, void* stream);
/* gray[v,p] = img[v,p] / max_v  (smoke normalisation, :158). */
int lnst_normalize_fwd(const float* img, const float* stats, int32_t n_img, int64_t n_pix, float* gray,
                       void* stream);
/* lnst_normalize_fwd plus the ties count of lnst_image_max in one pass: stats[2*v] = max must be final
 * (lnst_raymarch_fwd_max_*), stats[2*v+1] zero on entry. */
int lnst_normalize_ties_fwd(const float* img, float* stats, int32_t n_img, int64_t n_pix, float* gray, void* stream);
/* g_img from g_gray incl. the reduce_max gradient (split evenly among ties); `dots`
 * [n_img] is workspace. */
int lnst_normalize_bwd(const float* img, const float* stats, const float* g_gray, int32_t n_img,
                       int64_t n_pix, float* dots, float* g_img, void* stream);
/* tf.compat.v1.image.resize(BILINEAR), legacy coordinates, C channels NHWC. */
int lnst_resize_bilinear_fwd(const float* in, int32_t n_img, int32_t H, int32_t W, int32_t C,
                             int32_t OH, int32_t OW, float* out, void* stream);
int lnst_resize_bilinear_bwd(const float* g_out, int32_t n_img, int32_t H, int32_t W, int32_t C,
                             int32_t OH, int32_t OW, float* g_in, void* stream);
/* tf.compat.v1.image.resize(BICUBIC) with legacy coordinates (Keys a = -0.75, clamped taps): the style
 * mask of styler_base.py:165-169.  in [n,H,W,C] -> out [n,OH,OW,C]. */
int lnst_resize_bicubic_fwd(const float* in, int32_t n_img, int32_t H, int32_t W, int32_t C, int32_t OH,
                            int32_t OW, float* out, void* stream);
/* Gradient of lnst_resize_bicubic_fwd w.r.t. its input (the 3-D style mask is the render itself, so the mask
 * carries a gradient: styler_base.py:165-169); g_in [n,H,W,C] is overwritten. */
int lnst_resize_bicubic_bwd(const float* g_out, int32_t n_img, int32_t H, int32_t W, int32_t C, int32_t OH,
                            int32_t OW, float* g_in, void* stream);
/* out[p] (+)= sum_c a[p,c] b[p,c] + scale * scalar[0]  (scalar: device float, may be NULL): cotangent of the style
 * mask, <d loss/d (F m), F> per pixel plus the term through the masked area in the Gram denominator. */
int lnst_rowdot(const float* a, const float* b, int32_t C, int64_t P, const float* scalar, float scale,
                int32_t accumulate, float* out, void* stream);
/* d_img[v,p,c] = s*gray[v,p,(c)] (gray has Cg = 1 or 3 channels), x = d_img - mean_rgb. */
int lnst_to_net_input_fwd(const float* gray, int32_t n_img, int64_t n_pix, int32_t Cg, float s,
                          float* d_img, float* x, void* stream);
/* g_gray[v,p,(c)] = s * (sum over the replicated channels of g_x). */
int lnst_to_net_input_bwd(const float* g_x, int32_t n_img, int64_t n_pix, int32_t Cg, float s,
                          float* g_gray, void* stream);

/* ---- loss network, CUDA-core fp32 path (vgg.py:68-113) ------------------------------------ */
/* y = [relu](conv3x3_SAME(x, w) + b); x [n,H,W,Cin], w [3,3,Cin,Cout] (HWIO), y [n,H,W,Cout].
 * mask (may be NULL): y *= (mask > 0) elementwise (fused ReLU backward for dgrad chains). */
int lnst_conv3x3_f32(const float* x, const float* w, const float* b, const float* mask, float* y,
                     int32_t n, int32_t H, int32_t W, int32_t Cin, int32_t Cout, int32_t relu,
                     void* stream);
int lnst_avgpool2_fwd(const float* x, float* y, int32_t n, int32_t H, int32_t W, int32_t C, void* stream);
/* g_x = pool_bwd(g_y) * (mask > 0) (mask may be NULL). */
int lnst_avgpool2_bwd(const float* g_y, const float* mask, float* g_x, int32_t n, int32_t H, int32_t W,
                      int32_t C, void* stream);

/* ---- loss network, tensor-core path (same reference lines; bf16 NHWC activations) ---------- */
/* 1 when the driver exposes cuTensorMapEncodeTiled (TMA descriptors can be built). */
int lnst_tc_supported(void);
/* tcgen05 + TMA implicit-GEMM 3x3 SAME convolution.  x [n,H,W,Cin] bf16, w_packed
 * [9,Cout,Cin] bf16 (tap-major, K-major rows), bias fp32 [Cout] or NULL, mask bf16 like y or
 * NULL, y [n,H,W,Cout] bf16.  Cin and Cout must be multiples of 64.  fp32 accumulation. */
int lnst_conv3x3_bf16_tc(const void* x, const void* w_packed, const float* bias, const void* mask,
                         void* y, int32_t n, int32_t H, int32_t W, int32_t Cin, int32_t Cout,
                         int32_t relu, void* stream);
/* CUDA-core convolution with bf16 <-> fp32 I/O for the thin edge layers (conv1_1, Cin = 3, and
 * its data gradient, Cout = 3).  w fp32 HWIO; x_bf16 / y_bf16 select the element types; mask bf16. */
int lnst_conv3x3_mixed(const void* x, int32_t x_bf16, const float* w, const float* b, const void* mask,
                       void* y, int32_t y_bf16, int32_t n, int32_t H, int32_t W, int32_t Cin,
                       int32_t Cout, int32_t relu, void* stream);
/* Dedicated kernels for VGG's conv1_1 (3 -> 64, K = 27): forward x fp32 [n,H,W,3] -> y bf16
 * [n,H,W,64] with bias + ReLU (w fp32 HWIO [3,3,3,64]); data gradient g bf16 [n,H,W,64] -> gx fp32
 * [n,H,W,3] (wd fp32 [3,3,64,3] = flipped/transposed w). */
int lnst_conv_first_fwd(const float* x, const float* w, const float* b, void* y, int32_t n, int32_t H,
                        int32_t W, void* stream);
int lnst_conv_first_bwd(const void* g, const float* wd, float* gx, int32_t n, int32_t H, int32_t W,
                        void* stream);
/* Tuning switch (process-wide, default 1): 1 = persistent convolution kernel (one CTA per SM walks the
 * tile list, TMEM double-buffered), 0 = one CTA per tile. */
int lnst_set_conv_persistent(int32_t on);
/* Developer probe (tools/umma_probe.py): one M=128,N=16,K=64 MMA whose A operand is tap (ky,kx) of an 18x10
 * pixel patch addressed only through the shared-memory descriptor; x bf16 [180,64], b bf16 [16,64], out
 * fp32 [128,16]. */
int lnst_umma_probe(const void* x, const void* b, float* out, int32_t pitched, int32_t ky, int32_t kx,
                    int32_t base_mode, void* stream);
/* Tuning switch (default 2): 1 = layers whose weights fit in shared memory load one halo'd patch per tile
 * and read the 9 taps through descriptor offsets; 2 = every layer does (weights streamed when they do not
 * fit); 0 = one TMA tile per tap everywhere. */
int lnst_set_conv_halo(int32_t on);
/* Tuning switch (tests / microbenchmarks): 1 (default) = the gray data gradient of conv1_1 (lnst_conv_first_bwd_gray[_x3]_tc)
 * runs as one GEMM per halo'd patch with the 9 taps as the N dimension plus a 9-term gather; 0 = the halo kernel with
 * one MMA chain per tap. */
int lnst_set_conv_first_col(int32_t on);
/* Tuning switch (tests / microbenchmarks): 1 = lnst_conv_first_fwd_gray_x3 runs as one K = 64 GEMM per 128-pixel tile
 * (gray values and weights split into three bf16 pieces, fp32-accurate); 0 (default) = the CUDA-core kernel, which is
 * faster in the step. */
int lnst_set_conv_first_mma(int32_t on);
/* lnst_conv_first_bwd_gray[_x3]_tc plus dots[i] += sum_p g_gray[i,p] * img[i,p] out of the same kernel (the reduction
 * lnst_normalize_bwd starts with); dots [n] zero on entry, img fp32 [n,H,W] the un-normalised render. */
int lnst_conv_first_bwd_gray_dot_tc(const void* g, const void* wd16, float* g_gray, const float* img, float* dots,
                                    int32_t split, int32_t n, int32_t H, int32_t W, void* stream);
/* Tuning switch (tests / microbenchmarks): 1 (default) = lnst_gram_diff_bf16x3_tc sums hi^T hi + hi^T lo + lo^T hi in one
 * TMEM tile per 128 x 128 block (C % 128 == 0); 0 = the 2C x 2C Gram of the split rows plus a finishing pass. */
int lnst_set_gram_split3(int32_t on);
/* Data gradient of conv1_1 on tensor cores: g bf16 [n,H,W,64], wd16 bf16 [9,16,64] (rows 0..2 = the
 * flipped/transposed 64->3 weights, rows 3..15 zero) -> gx fp32 [n,H,W,3]. */
int lnst_conv_first_bwd_tc(const void* g, const void* wd16, float* gx, int32_t n, int32_t H, int32_t W,
                           void* stream);
/* conv1_1 forward and data gradient for a GRAY render replicated to RGB (styler_base.py:41-43, vgg.py:50-53):
 * x_c = s*gray - mean_c is never materialised.  ws[9][64] = s*sum_c w[tap,c,:], wm[9][64] = sum_c w[tap,c,:]*mean_c,
 * bsum[64] = b - sum_tap wm; y bf16 [n,H,W,64] = relu(conv).  Backward: wd16 bf16 [9,16,64] with row 0 =
 * s*sum_c of the flipped/transposed rows, rows 1..15 zero -> g_gray fp32 [n,H,W]. */
int lnst_conv_first_fwd_gray(const float* gray, const float* ws, const float* wm, const float* bsum, void* y, int32_t n,
                             int32_t H, int32_t W, void* stream);
int lnst_conv_first_bwd_gray_tc(const void* g, const void* wd16, float* g_gray, int32_t n, int32_t H, int32_t W,
                                void* stream);
/* Gram matrices on tensor cores, batched over images (styler_base.py:96-102,152-185):
 * G[i] = F[i]^T F[i] / denom - Gs (fp32 [n,C,C]; Gs NULL => no subtraction), Gd = bf16 copy of G,
 * loss[i] += weight * sum(G[i]^2).  F bf16 [n,P,C], C a multiple of 64.  tcgen05 with MN-major
 * operands, split-K over the pixels. */
int lnst_gram_diff_bf16_tc(const void* F, int32_t n, int64_t P, int32_t C, float denom, const float* Gs,
                           float weight, float* G, void* Gd, float* loss, void* stream);
/* Gram-loss gradient on tensor cores: g = (addend + coef * F x Gd) [* (F > 0) if relu_mask].
 * F, addend (may be NULL or == g), g: bf16 [n,H,W,C]; Gd bf16 [n,C,C] (symmetric). */
int lnst_gram_bwd_bf16_tc(const void* F, const void* Gd, float coef, const void* addend, int32_t relu_mask,
                          void* g, int32_t n, int32_t H, int32_t W, int32_t C, void* stream);
int lnst_avgpool2_bf16_fwd(const void* x, void* y, int32_t n, int32_t H, int32_t W, int32_t C,
                           void* stream);
int lnst_avgpool2_bf16_bwd(const void* g_y, const void* mask, void* g_x, int32_t n, int32_t H,
                           int32_t W, int32_t C, void* stream);
int lnst_f32_to_bf16(const float* x, void* y, int64_t n, void* stream);
int lnst_bf16_to_f32(const void* x, float* y, int64_t n, void* stream);

/* ---- TMA-tiled volume kernels (csrc/tiles_tma.cu) ---------------------------------------------------------------
 * Same results as the entry points they shadow, with the 3-D tiles of the volume staged in shared memory by TMA bulk
 * tensor copies.  They need rows of a multiple of 4 floats and 16-byte aligned volume pointers (LNST_EARG otherwise: the
 * caller then uses the SIMT entry point).  lnst_tma_supported(): 1 when the driver exposes cuTensorMapEncodeTiled. */
int lnst_tma_supported(void);
/* tuning switch for microbenchmarks: planes per TMA slab of the ray-march kernels (8, 12 or 16; default 12) */
int lnst_set_raymarch_slab(int32_t planes);
int lnst_smooth3_relu_fwd_tma(const float* in, float* out, int32_t D, int32_t H, int32_t W, int32_t k, const LnstBox* box,
                              void* stream);
int lnst_smooth3_relu_bwd_tma(const float* g_out, const float* out, float* g_in, int32_t D, int32_t H, int32_t W,
                              int32_t k, const LnstBox* box, void* stream);
/* p2g_wavg forward as a gather over per-cell particle lists: cstart [V+1] (first entry of every cell, cells in unflipped
 * (z,y,x) order), order [Nv] (particle index per entry), rel [Nv,3] (lnst_splat_cells offsets in list order).  Writes
 * out = sum_k where(wmap_k > 1e-6, num_k / wmap_k, num_k) for the cells of `box` (output coordinates, NULL = all) with
 * one TMA store per 4 x 8 x 32 tile; wmap is accumulated on the fly.  nsize = 1, clip = 0. */
int lnst_splat_wavg_fwd_gather(const int32_t* cstart, const int32_t* order, const float* rel, const float* r,
                               const float* var, const LnstGrid* g, const float* h, int32_t nk, float* out,
                               const LnstBox* box, void* stream);
/* images bit-identical to lnst_raymarch_fwd_box (same per-sample arithmetic, same order along the ray) */
int lnst_raymarch_fwd_tma(const float* vol, const float* rot, int32_t n_views, int32_t D, int32_t H, int32_t W, float tau,
                          int32_t liquid, const LnstBox* box, const int32_t* intervals, float* img, float* stot,
                          void* stream);
/* lnst_raymarch_fwd_max_box on the TMA-staged kernel */
int lnst_raymarch_fwd_max_tma(const float* vol, const float* rot, int32_t n_views, int32_t D, int32_t H, int32_t W,
                              float tau, int32_t liquid, const LnstBox* box, const int32_t* intervals, float* img,
                              float* stot, float* stats, void* stream);

/* lnst_advect for a 3-D scalar field (dim = 3, C = 1): d [D,H,W], vel [D,H,W,3], out [D,H,W].  The source box of each
 * 8 x 8 x 32 output tile, grown by `reach` cells (1..4: the caller's bound on the back-trace length), is one TMA box;
 * longer back-traces gather from global memory, so the result does not depend on the bound. */
int lnst_advect3_tma(const float* d, const float* vel, int32_t D, int32_t H, int32_t W, int32_t reach, float* out,
                     void* stream);
/* Data gradient of conv1_1 w.r.t. a gray render on the CUDA cores, the 18 x 18 patch of every 16 x 16 pixel tile staged by
 * one TMA box: g bf16 [n,H,W,64] (split = 0) or [n,H,W,128] = [hi | lo] (split = 1), wg fp32 [9,64] (data-gradient
 * weights summed over the three input channels times the input scale) -> g_gray fp32 [n,H,W].  Same result as
 * lnst_conv_first_bwd_gray[_x3]_tc with un-rounded weights. */
int lnst_conv_first_bwd_gray_direct(const void* g, int32_t split, const float* wg, float* g_gray, int32_t n, int32_t H,
                                    int32_t W, void* stream);
/* smoke render (liquid = 0) only; g_vol accumulates like lnst_raymarch_bwd_box */
int lnst_raymarch_bwd_tma(const float* vol, const float* rot, int32_t n_views, int32_t D, int32_t H, int32_t W, float tau,
                          const LnstBox* box, const int32_t* intervals, const float* stot, const float* g_img,
                          float* g_vol, void* stream);

/* ---- bf16x3 ("split") tensor-core loss network: fp32-tolerance results at tensor-core speed (vgg.py:89-113 is fp32) ----
 * Every fp32 value v travels as two bf16 halves hi = bf16(v), lo = bf16(v - hi): an NHWC row of C logical channels is
 * 2C bf16 = [hi(0..C-1) | lo(0..C-1)], weights are packed [9, Cout, 2*Cin] = [Whi | Wlo].  A convolution runs three K
 * passes -- x_hi*W_hi, x_lo*W_hi, x_hi*W_lo -- into one fp32 TMEM accumulator (products carry 16 mantissa bits; the
 * lo*lo term, 2^-16 relative, is dropped) and its epilogue splits the fp32 result again.  Same kernels and argument
 * meaning as the bf16 entry points above; `C`, `Cin`, `Cout` are LOGICAL channel counts. */
int lnst_conv3x3_bf16x3_tc(const void* x, const void* w_packed2, const float* bias, const void* mask, void* y,
                           int32_t n, int32_t H, int32_t W, int32_t Cin, int32_t Cout, int32_t relu, void* stream);
/* The same convolution with the 2x2 average pool that follows it in the network (vgg.py:96,102 `pool1`, `pool2`) written by
 * the same epilogue: y_pool bf16 [n,H/2,W/2,2*Cout], bit-identical to lnst_avgpool2_bf16x3_fwd(y). */
int lnst_conv3x3_pool_bf16x3_tc(const void* x, const void* w_packed2, const float* bias, const void* mask, void* y,
                                void* y_pool, int32_t n, int32_t H, int32_t W, int32_t Cin, int32_t Cout, int32_t relu,
                                void* stream);
/* Data gradient through a 3x3 convolution plus the Gram-loss gradient of the layer it lands on, masked together:
 * y = relu_mask(F) * (x (*) w_packed2 + F x Gd2s), F split [n,H,W,2*Cout] (the layer's features), Gd2s split [n,Cout,2*Cout]
 * from lnst_gram_diff_scaled_bf16x3_tc with gd_scale = the loss coefficient.  One kernel when Cout % 128 == 0 (the tile
 * accumulates both products); otherwise the convolution followed by lnst_gram_bwd_bf16x3_tc.  styler_base.py:98-109. */
int lnst_conv3x3_gram_bf16x3_tc(const void* x, const void* w_packed2, const void* F, const void* Gd2s, void* y, int32_t n,
                                int32_t H, int32_t W, int32_t Cin, int32_t Cout, void* stream);
/* Data gradient through a 3x3 convolution whose input is a 2x2 average pool's output, written at the pool's INPUT level:
 * g_fine [n,2H,2W,2*Cout] = lnst_avgpool2_bf16x3_bwd(lnst_conv3x3_bf16x3_tc(x, ...), mask = fine_act), bit-identical, from
 * the convolution's epilogue (vgg.py:96,102).  H, W = the coarse level.  LNST_EARG when the layer's weights are resident in
 * shared memory (9 * 2*Cin * Cout * 2 B fit): then call the two separately. */
int lnst_conv3x3_unpool_bf16x3_tc(const void* x, const void* w_packed2, const void* fine_act, void* g_fine, int32_t n,
                                  int32_t H, int32_t W, int32_t Cin, int32_t Cout, void* stream);
/* lnst_gram_diff_bf16x3_tc with Gd2 = split(gd_scale * G); gd_scale != 1 needs C % 128 == 0. */
int lnst_gram_diff_scaled_bf16x3_tc(const void* F, int32_t n, int64_t P, int32_t C, float denom, const float* Gs,
                                    float weight, float gd_scale, float* G2, float* G, void* Gd2, float* loss, void* stream);
/* G2: fp32 scratch [n,2C,2C]; G fp32 [n,C,C] = F^T F/denom - Gs; Gd2 bf16 [n,C,2C] = split copy of G (may be NULL). */
int lnst_gram_diff_bf16x3_tc(const void* F, int32_t n, int64_t P, int32_t C, float denom, const float* Gs, float weight,
                             float* G2, float* G, void* Gd2, float* loss, void* stream);
int lnst_gram_bwd_bf16x3_tc(const void* F, const void* Gd2, float coef, const void* addend, int32_t relu_mask, void* g,
                            int32_t n, int32_t H, int32_t W, int32_t C, void* stream);
int lnst_conv_first_fwd_x3(const float* x, const float* w, const float* b, void* y, int32_t n, int32_t H, int32_t W,
                           void* stream);
int lnst_conv_first_fwd_gray_x3(const float* gray, const float* ws, const float* wm, const float* bsum, void* y,
                                int32_t n, int32_t H, int32_t W, void* stream);
/* wd16 bf16 [9,16,128] = [hi | lo] rows */
int lnst_conv_first_bwd_x3_tc(const void* g, const void* wd16, float* gx, int32_t n, int32_t H, int32_t W, void* stream);
int lnst_conv_first_bwd_gray_x3_tc(const void* g, const void* wd16, float* g_gray, int32_t n, int32_t H, int32_t W,
                                   void* stream);
int lnst_avgpool2_bf16x3_fwd(const void* x, void* y, int32_t n, int32_t H, int32_t W, int32_t C, void* stream);
int lnst_avgpool2_bf16x3_bwd(const void* g_y, const void* mask, void* g_x, int32_t n, int32_t H, int32_t W, int32_t C,
                             void* stream);
/* fp32 [rows, C] <-> split bf16 [rows, 2C] */
int lnst_f32_to_bf16x3(const float* x, void* y, int64_t rows, int32_t C, void* stream);
int lnst_bf16x3_to_f32(const void* x, float* y, int64_t rows, int32_t C, void* stream);

/* ---- losses (styler_base.py:96-102,135-185,211-213) --------------------------------------- */
/* G [C,C] = F^T F / denom - Gs (the difference is what both the loss and its gradient need);
 * loss[0] += weight * sum(G^2).  F [P,C].  Gs NULL => G = F^T F / denom (style-target pass). */
int lnst_gram_diff(const float* F, int64_t P, int32_t C, float denom, const float* Gs, float weight,
                   float* G, float* loss, void* stream);
/* lnst_gram_diff with the denominator on the device: denom = den_scale * den_dev[0] when den_dev != NULL (`denom` is then
 * unused).  The 3-D style mask's Gram denominators are 2 C * area(mask) with the mask = the current render
 * (styler_base.py:165-169): reading the area back would break the step's CUDA graph. */
int lnst_gram_diff_dev(const float* F, int64_t P, int32_t C, float denom, const float* den_dev, float den_scale,
                       const float* Gs, float weight, float* G, float* loss, void* stream);
/* out[i] = x[i] * num / (den_scale * den_dev[0]): folds a coefficient with a device-resident denominator into the small
 * operand it multiplies (the Gram difference ahead of lnst_gram_bwd, a loss scalar ahead of lnst_rowdot). */
int lnst_scale_by_dev(const float* x, int64_t n, float num, const float* den_dev, float den_scale, float* out, void* stream);
/* g_F = (beta*g_F + coef * F G) [* (F > 0) when relu_mask: F is a post-ReLU conv output and
 * g_F becomes the gradient w.r.t. the pre-activation] */
int lnst_gram_bwd(const float* F, const float* G, int64_t P, int32_t C, float coef, float beta,
                  int32_t relu_mask, float* g_F, void* stream);
/* content loss without target image (styler_base.py:143-148): loss[0] += weight*L and
 * g_F = beta*g_F + weight * dL/dF [* (F > 0) when relu_mask]; g_F may be NULL. */
int lnst_content_loss(const float* F, int64_t P, int32_t C, int32_t channel, float weight, float* loss,
                      float* g_F, float beta, int32_t relu_mask, void* stream);
/* Content loss against a target image's feature (styler_base.py:137-141): loss += weight *
 * mean((F - target*amp)^2) over the n_el elements; g_F = beta*g_F + weight * d/dF [* (F > 0)]. */
int lnst_content_mse(const float* F, const float* target, int64_t n_el, float amp, float weight, float* loss,
                     float* g_F, float beta, int32_t relu_mask, void* stream);
/* anisotropic L1 total variation of d_img [H,W,C]; g_img overwritten with weight * d tv. */
int lnst_tv_loss(const float* d_img, int32_t H, int32_t W, int32_t C, float weight, float* loss,
                 float* g_img, void* stream);

/* ---- optimiser + loop glue (styler_3p.py:320-363) ----------------------------------------- */
/* TF-1.15 ApplyAdam with g = grad*gscale; var is passed through nan_to_num afterwards. */
int lnst_adam_step(float* var, const float* grad, float* m, float* v, int64_t n, float lr_t, float beta1,
                   float beta2, float eps, float gscale, void* stream);
/* Same update with the step counter resident on the device (CUDA-graph replayable): state[3] =
 * {beta1^t, beta2^t, lr_t}, initialised by the caller to {beta1, beta2, 0}.  Computes lr_t =
 * lr*sqrt(1-beta2^t)/(1-beta1^t) in fp32, applies the update, then advances the powers. */
int lnst_adam_step_dev(float* var, const float* grad, float* m, float* v, int64_t n, float* state, float lr,
                       float beta1, float beta2, float eps, float gscale, void* stream);
/* A whole single-Adam-step iteration of one frame (styler_3p.py:312,331,359-363,385-386) in one pass:
 * var = g_opt; ApplyAdam(var, grad*gscale) with the device-resident step counter `state`; var_out = var;
 * delta = (nan_to_num(var) - g_opt) * (mask ? mask[(i/width)*mask_stride] : 1); apply != 0: g_opt += delta. */
int lnst_adam_iterate_dev(float* g_opt, const float* grad, float* m, float* v, int64_t n, float* state, float lr,
                          float beta1, float beta2, float eps, float gscale, const float* mask, int32_t width,
                          int32_t mask_stride, float* var_out, float* delta, int32_t apply, void* stream);
/* acc = (first ? 0 : acc) + nan_to_num(var) */
int lnst_iterate_accumulate(float* acc, const float* var, int64_t n, int32_t first, void* stream);
/* delta[i] = (nan_to_num(g_new[i]*scale) - g_opt[i]) * (mask ? mask[(i/width)*mask_stride] : 1) */
int lnst_iterate_delta(const float* g_new, float scale, const float* g_opt, const float* mask,
                       int32_t width, int32_t mask_stride, int64_t n, float* delta, void* stream);
/* g[i,c] = beta*g[i,c] + t[i,c]*m[i] [* (f[i,c] > 0)]: gradient of a masked Gram loss back onto the
 * unmasked feature (styler_base.py:167). */
int lnst_masked_accumulate(const float* t, const float* m, const float* f, int32_t relu, int32_t C, float beta,
                           float* g, int64_t n, void* stream);
/* scipy.ndimage.gaussian_filter along the first axis of x [T,M] (mode reflect, truncate 4). */
int lnst_temporal_gauss(const float* x, float* y, int32_t T, int64_t M, float sigma, void* stream);
/* y += a*x */
int lnst_axpy(float* y, const float* x, float a, int64_t n, void* stream);
/* out[0] = scale * sum(x[0..n)) (the iteration's mean loss over its views, styler_3p.py:342); x[0..n) = 0 on the stream. */
int lnst_sum_scale(const float* x, int32_t n, float scale, float* out, void* stream);
int lnst_zero(float* x, int64_t n, void* stream);
/* tf.clip_by_value (styler_2p.py:68,88,94) and its gradient gx = scale*g inside [lo,hi], 0 outside. */
int lnst_clip_fwd(const float* x, float lo, float hi, float* y, int64_t n, void* stream);
int lnst_clip_bwd(const float* g, const float* x, float lo, float hi, float scale, float* gx, int64_t n,
                  void* stream);
/* out[i] = a[i] * b[i / C]  (colour field x density mask, styler_2p.py:71,100); n = elements of a. */
int lnst_mul_bcast(const float* a, const float* b, int32_t C, float* out, int64_t n, void* stream);

/* ---- semi-Lagrangian advection, order 1 (transform.py:557-609) ---------------------------- */
/* d [X,Y,(Z),C], vel [X,Y,(Z),dim] in normalised units; dims[0..dim-1]. */
int lnst_advect(const float* d, const float* vel, int32_t dim, const int32_t* dims, int32_t C, float* out,
                void* stream);

/* ---- grid -> particle gathers and the resimulation step (transform.py:771-1231, -------------
 *      test_smokegun_resim.py:17-112: the data-prep stage that produces the stylisation inputs) --- */
/* g2p: sample all C channels of a grid g [n0,n1,(n2),C] at normalised particle positions p (+ disp, may be
 * NULL) [n,dim] in the grid's axis order.  linear = 0: Catmull-Rom (g2p_cubic, :778-1108, 4^dim clamped
 * taps, offsets measured from the clamped anchor as the reference does); 1: g2p_linear (:1110-1231).
 * out [n,C] is overwritten. */
int lnst_g2p(const float* g, int32_t dim, const int32_t* dims, int32_t C, const float* p, const float* disp,
             int64_t n, int32_t linear, float* out, void* stream);
/* RK4 particle advection through a velocity grid u [n0,n1,(n2),dim] (normalised units, channel k <-> axis k):
 * v, v1 = u(x + v/2), v2 = u(x + v1/2), v3 = u(x + v2); x_adv = x + time_step*(v + 2 v1 + 2 v2 + v3)/6
 * (test_smokegun_resim.py:36-55).  One kernel; v_out (may be NULL) receives the blended velocity. */
int lnst_rk4_advect(const float* u, int32_t dim, const int32_t* dims, const float* x, int64_t n, float time_step,
                    int32_t linear, float* x_adv, float* v_out, void* stream);
/* loss += weight * mean(where(d_rec > 0, d_rec - rest_density, 0)^2) and, if g_d != NULL, its gradient w.r.t.
 * d_rec (test_smokegun_resim.py:69-74; styler_3p.py:96-98).  The caller zeroes `loss` (one device float). */
int lnst_pressure_loss(const float* d_rec, int64_t cells, float rest_density, float weight, float* loss,
                       float* g_d, void* stream);
/* Loss regularisers of the stylisation step, accumulating into the step's buffers:
 * pressure (styler_3p.py:96-98, styler_base.py:228-230): loss[0..n_loss) += w_mean * mean(pr^2) and
 * g_d[i] += g_scale * pr_i with pr = d > 0 ? d - rest_density : 0 (g_d may be NULL). */
int lnst_pressure_reg(const float* d, int64_t cells, float rest_density, float w_mean, float g_scale, float* loss,
                      int32_t n_loss, float* g_d, void* stream);
/* density (styler_base.py:217-223 on dv = clip(var, -1, 1)): loss[0..n_loss) += weight * ((sum dv)^2 + 1e3 * sum
 * -log(|dv| + 1e-6)); grad[i] += g_weight * [-1 <= var_i <= 1] * (2 sum dv - 1e3 sign(dv_i) / (|dv_i| + 1e-6)).
 * sums: 2 floats of scratch. */
int lnst_density_reg(const float* var, int64_t n, float weight, float g_weight, float* sums, float* loss,
                     int32_t n_loss, float* grad, void* stream);
/* out[z,y,x] = a[z,y,x] - b[z,H-1-y,x]: residual between a density grid and a splatted field (whose H axis is
 * stored flipped): `d - d_hi[:,:,::-1]` and d_diff of test_smokegun_resim.py:92,106. */
int lnst_sub_fliph(const float* a, const float* b, float* out, int32_t D, int32_t H, int32_t W, void* stream);

/* ---- GraphDef loss networks (inception5h; styler_base.py:19-31,53-57,91-94) -- fp32 NHWC ------------
 * The reference imports tensorflow_inception_graph.pb with tf.import_graph_def and reads layers by tensor
 * name; these replace the TF ops of that graph (Conv2D + BiasAdd, Relu, MaxPool, LRN, Concat) and their
 * gradients w.r.t. the data.  `accumulate` != 0: the result is ADDED to g_x (a tensor feeding several
 * branches sums their cotangents). */
/* y[pix, 0:Cout] = [relu](conv(x, w HWIO) + bias), stride/padding explicit (TF SAME: pad_top = pad_total/2);
 * ldy >= Cout = row stride of y in floats (write into a channel slice of a concat buffer). bias may be NULL. */
int lnst_conv2d_f32(const float* x, const float* w, const float* bias, float* y, int32_t n, int32_t H, int32_t W,
                    int32_t Cin, int32_t Cout, int32_t kh, int32_t kw, int32_t stride, int32_t pad_top,
                    int32_t pad_left, int32_t OH, int32_t OW, int32_t ldy, int32_t relu, void* stream);
/* g_x [n,H,W,Cin] (+)= conv2d data gradient of g_y (row stride ldg >= Cout).  relu_y (may be NULL): the post-ReLU
 * output of a convolution that ran with relu = 1, indexed like g_y -- the cotangent is masked by (relu_y > 0) while
 * it is loaded (tf.nn.relu's gradient without a pass of its own). */
int lnst_conv2d_bwd_data_f32(const float* g_y, const float* relu_y, int32_t ldg, const float* w, float* g_x, int32_t n, int32_t H,
                             int32_t W, int32_t Cin, int32_t Cout, int32_t kh, int32_t kw, int32_t stride,
                             int32_t pad_top, int32_t pad_left, int32_t OH, int32_t OW, int32_t accumulate,
                             void* stream);
int lnst_relu_fwd(const float* x, float* y, int64_t n, void* stream);
/* g_x (+)= g_y * (y > 0) */
int lnst_relu_bwd(const float* g_y, const float* y, float* g_x, int64_t n, int32_t accumulate, void* stream);
/* tf.nn.max_pool k x k (padding cells never win) and MaxPoolGrad (first maximum in scan order takes the cotangent). */
int lnst_maxpool_fwd(const float* x, float* y, int32_t n, int32_t H, int32_t W, int32_t C, int32_t k, int32_t stride,
                     int32_t pad_top, int32_t pad_left, int32_t OH, int32_t OW, void* stream);
int lnst_maxpool_bwd(const float* g_y, const float* x, float* g_x, int32_t n, int32_t H, int32_t W, int32_t C,
                     int32_t k, int32_t stride, int32_t pad_top, int32_t pad_left, int32_t OH, int32_t OW,
                     int32_t accumulate, void* stream);
/* tf.nn.avg_pool k x k (SAME padding cells are not counted) and AvgPoolGrad -- the inception head's avgpool0. */
int lnst_avgpool_fwd(const float* x, float* y, int32_t n, int32_t H, int32_t W, int32_t C, int32_t k, int32_t stride,
                     int32_t pad_top, int32_t pad_left, int32_t OH, int32_t OW, void* stream);
int lnst_avgpool_bwd(const float* g_y, float* g_x, int32_t n, int32_t H, int32_t W, int32_t C, int32_t k, int32_t stride,
                     int32_t pad_top, int32_t pad_left, int32_t OH, int32_t OW, int32_t accumulate, void* stream);
/* tf.nn.lrn: y_c = x_c (bias + alpha sum_{|j-c|<=depth_radius} x_j^2)^-beta, and LRNGrad. */
int lnst_lrn_fwd(const float* x, float* y, int64_t pixels, int32_t C, int32_t depth_radius, float bias, float alpha,
                 float beta, void* stream);
int lnst_lrn_bwd(const float* g_y, const float* x, float* g_x, int64_t pixels, int32_t C, int32_t depth_radius,
                 float bias, float alpha, float beta, int32_t accumulate, void* stream);
/* dst[p, 0:C] (+)= src[p, 0:C] with row strides ld_src / ld_dst: tf.concat along channels and its gradient slices. */
int lnst_copy_channels(const float* src, int32_t ld_src, float* dst, int32_t ld_dst, int32_t C, int64_t pixels,
                       int32_t accumulate, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LNST_B200_H */
