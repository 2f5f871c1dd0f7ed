// Loss-network convolutions on the 5th-generation tensor cores (tcgen05) fed by TMA.
//   reference: vgg.py:68-113 -- slim.conv2d 3x3 SAME + bias + ReLU, executed there by cuDNN.
//
// 3x3 convolution as an implicit GEMM without im2col in memory:
//   D[pixel, co] = sum_{tap=(ky,kx)} sum_{ci} X[pixel + (ky-1,kx-1), ci] * Wt[tap][co][ci]
// * M tile = 128 output pixels = a TH x TW spatial patch of one image; N tile = 64/128 output
//   channels; K is walked as 9 taps x (Cin/64) channel chunks.
// * A operand: for every (tap, chunk) ONE 4-D tiled TMA load of the box
//   {64 ch, TW, TH, 1} at coordinates (c0, w0+kx-1, h0+ky-1, img) of the NHWC bf16 activation.
//   TMA zero-fills out-of-bounds coordinates (incl. negative ones), which IS the SAME
//   padding; the box lands in shared memory as 128 rows x 128 B = the K-major SWIZZLE_128B
//   UMMA operand layout, so no thread ever touches operand data.
// * B operand: weights pre-packed [tap][Cout][Cin] bf16 (K-major), 3-D TMA box {64, BN, 1}.
// * tcgen05.mma (cta_group::1, kind::f16, M=128, N=BN, K=16) issued by one thread, fp32
//   accumulators in TMEM; smem ring of 3 stages with mbarrier full/empty pairs
//   (tcgen05.commit releases a stage); 2 CTAs per SM overlap one tile's epilogue with the
//   other's main loop.
// * Epilogue: 4 warps tcgen05.ld their TMEM sub-partition, add bias, ReLU (forward) or apply
//   the ReLU mask of the layer below (data-gradient chain), convert to bf16, 16-byte stores.
// The data gradient is the same kernel on flipped/transposed weights (wd[ky,kx,co,ci] =
// w[2-ky,2-kx,ci,co]).  Layers with fewer than 64 input or output channels (conv1_1 and its
// gradient) run on the CUDA-core engine of sgemm.cuh with bf16 <-> fp32 I/O.
#include <cuda.h>
#include <cuda_bf16.h>
#include "sgemm.cuh"

namespace tc {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;          // bf16 elements = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int STAGES = 3;
constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;   // 16 KiB
constexpr int NUM_THREADS = 192;     // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2-5: epilogue
constexpr int PERSIST_THREADS = 320; // per-tap persistent kernel: 8 epilogue warps (two per TMEM sub-partition)

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// One lane of the (fully active) warp; the same lane every time for the same mask.  The producer and MMA
// warps run their loops with ALL lanes so that stage indices, phases, coordinates and descriptors stay
// warp-uniform (uniform registers), and only the TMA / MMA / commit instructions are guarded by the
// elected lane.  Running the loops inside `if (lane == 0)` made the compiler wrap every UTCHMMA / UTMALDG
// in an R2UR + ELECT + BRA.U.ANY "uniformise" loop: ~100 / ~70 issue slots per k-step in a single warp,
// i.e. 400-600 cycles against the 256 cycles its four MMAs need (ncu source view, profiles/).
__device__ __forceinline__ bool elect_one() {
#ifdef LNST_CPU_EMU
  return (threadIdx.x & 31) == 0;
#else
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}" : "=r"(pred));
  return pred != 0;
#endif
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start>>4 [0,14) | LBO>>4 [16,30) (unused for swizzled K-major, 1) | SBO>>4 [32,46) = 1024 B between
// 8-row groups | version=1 [46,48) | base_offset=0 (tiles are 1024 B aligned) | layout=2 [61,64)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// cute::UMMA::InstrDescriptor for kind::f16: c_format=F32 [4,6), a/b_format=BF16 [7,10)/[10,13),
// a/b K-major (bits 15,16 = 0), N>>3 [17,23), M>>4 [24,29)
__device__ __forceinline__ uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// Descriptors split into a constant high word and a low word that only carries the start address
// (>>4): advancing an operand is ONE 32-bit add in the single issuing thread.  (Rebuilding the 64-bit
// descriptor with shifts/ors for every MMA cost ~150 issue cycles per instruction -- more than the
// 32-64 cycles the tensor pipe needs for it -- and capped every kernel here at 30-45 % tensor-active.)
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr & 0x3FFFFu) >> 4) | ((lbo_bytes >> 4) << 16);
}
__device__ __forceinline__ uint32_t desc_hi_sw128(uint32_t sbo_bytes) {
  return (sbo_bytes >> 4) | (1u << 14) | (2u << 29);      // SBO | version=1 (bit 46) | SWIZZLE_128B (bits 61-63)
}
__device__ __forceinline__ void umma_bf16_lh(uint32_t tmem_d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
      "}" ::"r"(tmem_d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(accumulate) : "memory");
}
// The same MMA with the A operand kept in / taken from the collector buffer: two consecutive MMAs that share their A
// slice (x_hi * W_hi, then x_hi * W_lo in the bf16x3 K loop) read it from shared memory once (SASS: A_KEEP / A_REUSE).
__device__ __forceinline__ void umma_bf16_lh_keep(uint32_t tmem_d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi,
                                                  uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], da, db, %5, p;\n\t"
      "}" ::"r"(tmem_d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_bf16_lh_reuse(uint32_t tmem_d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi,
                                                   uint32_t idesc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, 1, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%0], da, db, %5, p;\n\t"
      "}" ::"r"(tmem_d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// the same load without the wait: several can be in flight before one tcgen05.wait::ld
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

struct ConvShape {
  int H, W, Cin, Cout;
  int TH, TW, tiles_w, tiles_h;
  int relu;
  int taps;      // 9: 3x3 convolution; 1: per-pixel GEMM (Gram gradient F x G)
  int w_img;     // 1: third coordinate of the weight map is the image index (per-image B matrix)
  float scale;   // multiplies the accumulator before addend / bias
  int out_ch;    // OUT3 epilogue: fp32 channels written per pixel (3, or 1 for the gray render)
  // bf16x3 ("split") operands: a value v travels as hi = bf16(v), lo = bf16(v - hi) in channels [0,C) and [C,2C) of one
  // NHWC row, weights as [Whi | Wlo] along Cin, and the K loop runs 3 passes -- x_hi*W_hi, x_lo*W_hi, x_hi*W_lo -- into
  // the same fp32 accumulator: products carry 16 mantissa bits instead of 8 at 3x the MMA count (the lo*lo term, 2^-16
  // relative, is dropped).  nchunks = K-loop length in 64-channel chunks (Cin/64, or 3*Cin/64), wchunks = weight
  // chunks per tap (Cin/64 or 2*Cin/64), kc = Cin/64; chunk c reads activation chunk xchunk(c) and weight chunk wchunk(c).
  int split, kc, nchunks, wchunks;
  int ldy, ldm;  // row strides (elements) of the output / mask / addend tensors: Cout, or 2*Cout when split
  int gram_kc;   // halo kernel, GRAM: 64-channel chunks of the feature tensor whose Gram gradient joins the tile (else 0)
};
__device__ __forceinline__ int xchunk(const ConvShape& s, int c) { return (s.split && c >= 2 * s.kc) ? c - 2 * s.kc : c; }
__device__ __forceinline__ int wchunk(const ConvShape& s, int c) { return (s.split && c >= s.kc) ? c - s.kc : c; }
static inline void shape_plain(ConvShape& s) {
  s.gram_kc = 0;
  s.split = 0; s.kc = s.Cin / 64; s.nchunks = s.kc; s.wchunks = s.kc; s.ldy = s.Cout; s.ldm = s.Cout;
}
static inline void shape_split(ConvShape& s) {
  s.gram_kc = 0;
  s.split = 1; s.kc = s.Cin / 64; s.nchunks = 3 * s.kc; s.wchunks = 2 * s.kc; s.ldy = 2 * s.Cout; s.ldm = 2 * s.Cout;
}
// hi/lo split of an fp32 pair
__device__ __forceinline__ void split2(float f0, float f1, __nv_bfloat162& hi, __nv_bfloat162& lo) {
  hi = __floats2bfloat162_rn(f0, f1);
  const float2 h = __bfloat1622float2(hi);
  lo = __floats2bfloat162_rn(f0 - h.x, f1 - h.y);
}

// y = mask( relu?( scale * (X (*) W) + addend + bias ) )
template <int BLOCK_N>
__global__ void __launch_bounds__(NUM_THREADS, 2)
conv3x3_tc_k(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
             const float* __restrict__ bias, const __nv_bfloat16* __restrict__ mask,
             const __nv_bfloat16* __restrict__ addend, __nv_bfloat16* __restrict__ y, ConvShape s) {
  constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operand tiles need 1024-byte alignment
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a = base;
  const uint32_t smem_b = base + STAGES * A_BYTES;
  const uint32_t bars = smem_b + STAGES * B_BYTES;      // full[STAGES], empty[STAGES], tmem_full
  const uint32_t tmem_slot = bars + 8 * (2 * STAGES + 1);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x;
  const int tw = tile % s.tiles_w, th = (tile / s.tiles_w) % s.tiles_h, img = tile / (s.tiles_w * s.tiles_h);
  const int h0 = th * s.TH, w0 = tw * s.TW, n0 = blockIdx.y * BLOCK_N;
  const int kchunks = s.nchunks;
  const int num_kb = s.taps * kchunks;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_x);
    prefetch_tmap(&map_w);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(bars + 8 * i, 1);
      mbar_init(bars + 8 * (STAGES + i), 1);
    }
    mbar_init(bars + 8 * (2 * STAGES), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // whole warp: allocate BLOCK_N TMEM columns (power of two >= 32)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(BLOCK_N));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      for (int kb = 0; kb < num_kb; ++kb) {
        const int st = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(bars + 8 * (STAGES + st), ph ^ 1);                 // slot free
        const uint32_t full = bars + 8 * st;
        mbar_expect_tx(full, A_BYTES + B_BYTES);
        const int tap = kb / kchunks, cc = kb - tap * kchunks, c0 = xchunk(s, cc) * BLOCK_K;
        int ky = 1, kx = 1;
        if (s.taps == 9) { ky = tap / 3; kx = tap - 3 * ky; }
        tma_load_4d(smem_a + st * A_BYTES, &map_x, full, c0, w0 + kx - 1, h0 + ky - 1, img);
        tma_load_3d(smem_b + st * B_BYTES, &map_w, full, wchunk(s, cc) * BLOCK_K, n0, s.w_img ? img : tap);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      const uint32_t idesc = umma_idesc_bf16(BLOCK_M, BLOCK_N);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int st = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(bars + 8 * st, ph);                                // operands landed
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t alo = desc_lo(smem_a + st * A_BYTES, 16), blo = desc_lo(smem_b + st * B_BYTES, 16);
        const uint32_t hi = desc_hi_sw128(1024);
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; ++k)   // advancing 16 bf16 = 32 B inside the 128 B swizzle row
          umma_bf16_lh(tmem_d, alo + k * 2, hi, blo + k * 2, hi, idesc, (kb | k) != 0 ? 1u : 0u);
        umma_commit(bars + 8 * (STAGES + st));                       // frees the slot when the MMAs retire
      }
      umma_commit(bars + 8 * (2 * STAGES));                          // accumulator complete
    }
  } else {
    // ===== epilogue: warps 2..5, TMEM sub-partition = warp % 4 =====
    const int q = warp & 3;
    const int r = q * 32 + lane;                                      // row of the tile = pixel
    const int ph_ = h0 + r / s.TW, pw_ = w0 + r % s.TW;
    const bool valid = ph_ < s.H && pw_ < s.W;
    const int64_t pix = ((int64_t)img * s.H + ph_) * s.W + pw_;
    // The ReLU mask of this thread's row does not depend on the accumulator: fetch it while the
    // main loop runs (all loads of a 32-channel chunk in flight together) and keep one bit per channel.
    uint32_t mbits[BLOCK_N / 32];
#pragma unroll
    for (int c = 0; c < BLOCK_N / 32; ++c) mbits[c] = 0xffffffffu;
    if (mask && valid) {
      const uint4* msk = reinterpret_cast<const uint4*>(mask + pix * s.Cout + n0);
#pragma unroll
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        uint4 mv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) mv[j] = msk[c * 4 + j];
        uint32_t bits = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const __nv_bfloat16* mh = reinterpret_cast<const __nv_bfloat16*>(&mv[j]);
#pragma unroll
          for (int e = 0; e < 8; ++e) bits |= (__bfloat162float(mh[e]) > 0.f ? 1u : 0u) << (j * 8 + e);
        }
        mbits[c] = bits;
      }
    }
    // first addend chunk: also independent of the accumulator
    uint4 av[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) av[j] = make_uint4(0, 0, 0, 0);
    const uint4* add = (addend && valid) ? reinterpret_cast<const uint4*>(addend + pix * s.Cout + n0) : nullptr;
    if (add) {
#pragma unroll
      for (int j = 0; j < 4; ++j) av[j] = add[j];
    }
    mbar_wait(bars + 8 * (2 * STAGES), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
    for (int c = 0; c < BLOCK_N / 32; ++c) {
      uint32_t v[32];
      tmem_ld32(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
      uint4 an[4];                                                     // next chunk's addend, in flight during the math
#pragma unroll
      for (int j = 0; j < 4; ++j) an[j] = make_uint4(0, 0, 0, 0);
      if (add && c + 1 < BLOCK_N / 32) {
#pragma unroll
        for (int j = 0; j < 4; ++j) an[j] = add[(c + 1) * 4 + j];
      }
      if (valid) {
        const int co = n0 + c * 32;
        __nv_bfloat16* dst = y + pix * s.Cout + co;
        uint4 ov[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const __nv_bfloat16* ah = reinterpret_cast<const __nv_bfloat16*>(&av[j]);
          __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&ov[j]);
#pragma unroll
          for (int e = 0; e < 8; e += 2) {
            float f0 = __uint_as_float(v[j * 8 + e]) * s.scale, f1 = __uint_as_float(v[j * 8 + e + 1]) * s.scale;
            if (addend) { f0 += __bfloat162float(ah[e]); f1 += __bfloat162float(ah[e + 1]); }
            if (bias) { f0 += bias[co + j * 8 + e]; f1 += bias[co + j * 8 + e + 1]; }
            if (s.relu) { f0 = fmaxf(f0, 0.f); f1 = fmaxf(f1, 0.f); }
            if (!((mbits[c] >> (j * 8 + e)) & 1u)) f0 = 0.f;
            if (!((mbits[c] >> (j * 8 + e + 1)) & 1u)) f1 = 0.f;
            oh[e >> 1] = __floats2bfloat162_rn(f0, f1);
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(dst + j * 8) = ov[j];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) av[j] = an[j];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(BLOCK_N));
  }
}

// ---- persistent variant --------------------------------------------------------------------------
// One CTA per SM walks the tile list (tile = 128 pixels x BLOCK_N channels).  The TMA producer and the
// MMA issuer keep one smem ring (PSTAGES deep, ~192 KB) running ACROSS tile boundaries, and the
// accumulator is double-buffered in TMEM (2 x BLOCK_N columns): while the four epilogue warps drain
// tile i (tcgen05.ld, bias/ReLU/mask/addend, bf16 stores), the tensor pipe already works on tile i+1.
// Per-CTA set-up (barrier init, TMEM allocation, descriptor prefetch) is paid once per SM instead of
// once per tile, and the mask / addend rows of the next tile are fetched before its accumulator is
// complete.
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// Coalesced write-out of a warp's 32 accumulator rows (NV x 16 B each).  A thread owns one pixel row, so a plain
// per-thread store makes every warp-wide instruction touch 32 different lines with 16 B each: the LSU
// serialises them and the N = 128 epilogues took longer than their tile's MMAs (knock-out experiments, DESIGN.md section 4: conv2_1
// 41 us without the stores, 80 us with them).  Rows go through a swizzled per-warp staging buffer instead and
// come back out so that NV consecutive lanes write one row: every instruction covers whole 64/128-byte runs.
__device__ __forceinline__ void st_shared_v4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
// position of 16-byte piece j inside row `row`: every 8 lanes of a warp-wide access hit 8 different bank groups
template <int NV>
__device__ __forceinline__ uint32_t rows_swz(int j, int row) {
  return NV == 8 ? (uint32_t)((j ^ row) & 7) : (uint32_t)((j + (row >> 1)) & 3);
}
template <int NV, class RowPtr>
__device__ __forceinline__ void warp_rows_store(uint32_t stage, int lane, const uint4 (&ov)[NV], RowPtr rowptr) {
  constexpr int RPI = 32 / NV;                         // rows per store instruction
#pragma unroll
  for (int j = 0; j < NV; ++j) st_shared_v4(stage + lane * (NV * 16) + (rows_swz<NV>(j, lane) << 4), ov[j]);
  __syncwarp();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int row = i * RPI + lane / NV, cj = lane % NV;
    const uint4 v = ld_shared_v4(stage + row * (NV * 16) + (rows_swz<NV>(cj, row) << 4));
    __nv_bfloat16* dst = rowptr(row);
    if (dst) *reinterpret_cast<uint4*>(dst + cj * 8) = v;
  }
  __syncwarp();
}

template <int BLOCK_N> struct PersistStages { static constexpr int value = (BLOCK_N == 128) ? 5 : 6; };

template <int BLOCK_N>
__global__ void __launch_bounds__(PERSIST_THREADS, 1)
conv3x3_tc_persist_k(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                     const float* __restrict__ bias, const __nv_bfloat16* __restrict__ mask,
                     const __nv_bfloat16* __restrict__ addend, __nv_bfloat16* __restrict__ y, ConvShape s,
                     int n_blocks_n, int n_tiles) {
  constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
  constexpr int PSTAGES = PersistStages<BLOCK_N>::value;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage_base = base;                     // 8 epilogue warps x 4 KiB (fp32 accumulator chunks, transposed)
  const uint32_t smem_a = base + 8 * 4096;
  const uint32_t smem_b = smem_a + PSTAGES * A_BYTES;
  const uint32_t bars = smem_b + PSTAGES * B_BYTES;     // full[P], empty[P], tfull[2], tempty[2]
  const uint32_t bar_tfull = bars + 8 * (2 * PSTAGES);
  const uint32_t bar_tempty = bar_tfull + 16;
  const uint32_t tmem_slot = bar_tempty + 16;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kchunks = s.nchunks;
  const int num_kb = s.taps * kchunks;
  const int tiles_sp = s.tiles_w * s.tiles_h;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_x);
    prefetch_tmap(&map_w);
    for (int i = 0; i < PSTAGES; ++i) {
      mbar_init(bars + 8 * i, 1);
      mbar_init(bars + 8 * (PSTAGES + i), 1);
    }
    mbar_init(bar_tfull, 1); mbar_init(bar_tfull + 8, 1);
    mbar_init(bar_tempty, 8); mbar_init(bar_tempty + 8, 8);          // one arrival per epilogue warp
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(2 * BLOCK_N));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot_ptr;

  if (warp == 0) {
    // ===== TMA producer (whole warp in the loop, one elected lane issues) =====
    const bool leader = elect_one();
    uint32_t st = 0, ph = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int nb = t % n_blocks_n, sp = t / n_blocks_n;
      const int img = sp / tiles_sp, rem = sp - img * tiles_sp;
      const int th = rem / s.tiles_w, tw = rem - th * s.tiles_w;
      const int h0 = th * s.TH, w0 = tw * s.TW, n0 = nb * BLOCK_N;
      for (int tap = 0; tap < s.taps; ++tap) {
        int ky = 1, kx = 1;
        if (s.taps == 9) { ky = tap / 3; kx = tap - 3 * ky; }
        const int wsel = s.w_img ? img : tap;
        for (int c = 0; c < kchunks; ++c) {
          mbar_wait(bars + 8 * (PSTAGES + st), ph ^ 1);
          if (leader) {
            const uint32_t full = bars + 8 * st;
            mbar_expect_tx(full, A_BYTES + B_BYTES);
            tma_load_4d(smem_a + st * A_BYTES, &map_x, full, xchunk(s, c) * BLOCK_K, w0 + kx - 1, h0 + ky - 1, img);
            tma_load_3d(smem_b + st * B_BYTES, &map_w, full, wchunk(s, c) * BLOCK_K, n0, wsel);
          }
          if (++st == PSTAGES) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (whole warp in the loop, one elected lane issues) =====
    const bool leader = elect_one();
    const uint32_t idesc = umma_idesc_bf16(BLOCK_M, BLOCK_N);
    const uint32_t alo_base = desc_lo(smem_a, 16), blo_base = desc_lo(smem_b, 16);
    const uint32_t hi = desc_hi_sw128(1024);
    uint32_t st = 0, ph = 0, lt = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++lt) {
      const uint32_t buf = lt & 1, bph = (lt >> 1) & 1;
      mbar_wait(bar_tempty + 8 * buf, bph ^ 1);                      // epilogue has drained this buffer
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t acc = tmem_d + buf * BLOCK_N;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(bars + 8 * st, ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t alo = alo_base + st * (A_BYTES >> 4), blo = blo_base + st * (B_BYTES >> 4);
        if (leader) {
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
            umma_bf16_lh(acc, alo + k * 2, hi, blo + k * 2, hi, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(bars + 8 * (PSTAGES + st));
        }
        if (++st == PSTAGES) { st = 0; ph ^= 1; }
      }
      if (leader) umma_commit(bar_tfull + 8 * buf);
    }
  } else {
    // ===== epilogue: warps 2..9.  TMEM sub-partition = warp % 4; warps 2-5 take the lower half of the tile's
    // columns, warps 6-9 the upper half.  An accumulator row lives in ONE thread, but a row-per-thread access to the
    // mask / addend / output rows touches 32 different 128-byte lines per instruction with 16 bytes each: the LSU
    // serialised them (ncu: l1tex 74 %, tensor pipe 12 % on the Gram-gradient GEMM, ~10 000 wavefronts per tile).  So
    // the fp32 accumulator chunk (32 rows x 32 columns per warp) goes through a swizzled shared-memory tile and ALL
    // the arithmetic happens in the transposed layout, where 4 lanes cover 64 contiguous bytes of a row =====
    constexpr int EN = BLOCK_N / 2;
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const uint32_t stage = stage_base + (uint32_t)(warp - 2) * 4096u;   // 32 rows x 128 B, 16-byte piece j of row r at j ^ (r & 7)
    const int sub = lane >> 2, pj = lane & 3;                            // coalesced layout: row 8 i + sub, columns 8 pj .. 8 pj + 7
    uint32_t lt = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++lt) {
      const int nb = t % n_blocks_n, sp = t / n_blocks_n;
      const int img = sp / tiles_sp, rem = sp - img * tiles_sp;
      const int th = rem / s.tiles_w, tw = rem - th * s.tiles_w;
      const int n0 = nb * BLOCK_N + half * EN;
      int64_t pix[4];
      bool valid[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int rr = q * 32 + i * 8 + sub;
        const int ph_ = th * s.TH + rr / s.TW, pw_ = tw * s.TW + rr % s.TW;
        valid[i] = ph_ < s.H && pw_ < s.W;
        pix[i] = ((int64_t)img * s.H + ph_) * s.W + pw_;
      }
      const uint32_t buf = lt & 1, bph = (lt >> 1) & 1;
      const uint32_t acc = tmem_d + buf * BLOCK_N + half * EN + ((uint32_t)(q * 32) << 16);
#pragma unroll
      for (int c = 0; c < EN / 32; ++c) {
        const int co = n0 + c * 32 + pj * 8;                             // this lane's 8 output channels
        // the chunk's mask / addend pieces, requested before the accumulator is waited for
        uint4 mv[4], av[4], al[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          mv[i] = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);   // bf16 1.0: keep
          av[i] = make_uint4(0, 0, 0, 0);
          al[i] = make_uint4(0, 0, 0, 0);
          if (valid[i]) {
            if (mask) mv[i] = *reinterpret_cast<const uint4*>(mask + pix[i] * s.ldm + co);
            if (addend) {
              av[i] = *reinterpret_cast<const uint4*>(addend + pix[i] * s.ldy + co);
              if (s.split) al[i] = *reinterpret_cast<const uint4*>(addend + pix[i] * s.ldy + s.Cout + co);
            }
          }
        }
        if (c == 0) {
          mbar_wait(bar_tfull + 8 * buf, bph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        uint32_t v[32];
        tmem_ld32(acc + (uint32_t)(c * 32), v);
        if (c == EN / 32 - 1) {
          // every TMEM read of this warp has completed: hand the buffer back to the MMA issuer
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          st_shared_v4(stage + lane * 128 + ((uint32_t)((j ^ lane) & 7) << 4),
                       make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
        __syncwarp();
        float bb[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) bb[e] = bias ? bias[co + e] : 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int row = i * 8 + sub;
          const uint4 x0 = ld_shared_v4(stage + row * 128 + ((uint32_t)(((2 * pj) ^ row) & 7) << 4));
          const uint4 x1 = ld_shared_v4(stage + row * 128 + ((uint32_t)(((2 * pj + 1) ^ row) & 7) << 4));
          const uint32_t xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
          const __nv_bfloat16* mh = reinterpret_cast<const __nv_bfloat16*>(&mv[i]);
          const __nv_bfloat16* ah = reinterpret_cast<const __nv_bfloat16*>(&av[i]);
          const __nv_bfloat16* alh = reinterpret_cast<const __nv_bfloat16*>(&al[i]);
          uint4 ov, ol;
          __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&ov);
          __nv_bfloat162* ohl = reinterpret_cast<__nv_bfloat162*>(&ol);
#pragma unroll
          for (int e = 0; e < 8; e += 2) {
            float f0 = __uint_as_float(xv[e]) * s.scale, f1 = __uint_as_float(xv[e + 1]) * s.scale;
            if (addend) { f0 += __bfloat162float(ah[e]); f1 += __bfloat162float(ah[e + 1]); }
            if (addend && s.split) { f0 += __bfloat162float(alh[e]); f1 += __bfloat162float(alh[e + 1]); }
            if (bias) { f0 += bb[e]; f1 += bb[e + 1]; }
            if (s.relu) { f0 = fmaxf(f0, 0.f); f1 = fmaxf(f1, 0.f); }
            if (!(__bfloat162float(mh[e]) > 0.f)) f0 = 0.f;
            if (!(__bfloat162float(mh[e + 1]) > 0.f)) f1 = 0.f;
            if (s.split) split2(f0, f1, oh[e >> 1], ohl[e >> 1]);
            else oh[e >> 1] = __floats2bfloat162_rn(f0, f1);
          }
          if (valid[i]) {
            __nv_bfloat16* dst = y + pix[i] * s.ldy + co;
            *reinterpret_cast<uint4*>(dst) = ov;
            if (s.split) *reinterpret_cast<uint4*>(dst + s.Cout) = ol;
          }
        }
        __syncwarp();
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(2 * BLOCK_N));
  }
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint64_t umma_desc_sw128_ex(uint32_t saddr, uint32_t sbo_bytes, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7u) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- halo'd-patch variant: every input byte crosses L2 -> smem ONCE per tile ------------------------
// The per-tap kernel above re-reads its 128 x 64 A tile for each of the 9 taps (ncu: lts sectors ~16x
// the input, tensor pipe 24-43 % active at Cin = 64).  Here the tile is TH x TW = 16 x 8 pixels and ONE
// TMA box {64 ch, TW+2, TH+2} per 64-channel chunk brings the halo'd patch (180 rows of 128 B,
// SWIZZLE_128B, OOB = zero = SAME padding).  The A operand of tap (ky,kx) is the same patch seen through
// another descriptor: start + (ky*(TW+2) + kx)*128 B, SBO = (TW+2)*128 B (one 8-row core group = one
// tile row) -- the swizzle is a function of the absolute smem address, so no base offset is needed
// (verified on the device by tools/umma_probe.py).  Weights stay RESIDENT in smem for the whole kernel
// when the layer's 9*Cin*Cout*2 B fit beside the patch ring, else they stream through their own ring.
// Persistent, TMEM double-buffered, epilogue as above.  OUT3: BLOCK_N = 16, fp32 [pixel,3] output of
// columns 0..2 (the data gradient of conv1_1).
constexpr int HTH = 16, HTW = 8;
constexpr int PATCH_ROWS = (HTH + 2) * (HTW + 2);          // 180
constexpr int PATCH_BYTES = PATCH_ROWS * 128;              // 23040
constexpr int PATCH_STRIDE = 23 * 1024;                    // stage stride, keeps every stage 1024 B aligned

struct HaloCfg {
  int sa, sb;          // patch stages, weight stages (sb = 0: weights resident)
  int n_blocks_n, n_tiles;
};

template <int BLOCK_N> struct HaloAcc { static constexpr int value = (BLOCK_N >= 128) ? 4 : 8; };
// store staging: 128-byte row chunks (8 x 16 B) for the wide tiles, 64-byte ones where shared memory is tight
template <int BLOCK_N> struct HaloStage { static constexpr int nv = 4; static constexpr int bytes = (BLOCK_N >= 128) ? 4 * nv * 512 : 0; };

// GRAM (streamed weights only): after the nine taps the tile also accumulates F x Gd -- the Gram-loss gradient of the layer
// this data gradient lands on (vgg.py style layers; styler_base.py:98-109) -- from the layer's own features F (same
// pixels: the centre tap of an F patch) and the per-image matrix Gd (pre-scaled by the loss coefficient), so the separate
// per-pixel GEMM with its read-modify-write of the whole gradient tensor disappears.  s.gram_kc = C_F / 64.
// UPS: the output is the gradient of a 2x2 average pool's OUTPUT; the epilogue writes the pool's INPUT gradient instead
// (lnst_avgpool2_bf16x3_bwd: 0.25 * value to the four pixels of the quad, each under the ReLU mask of the layer below),
// through a transposed shared-memory tile so that 4 lanes cover 64 contiguous bytes of a row.  Even H and W at the fine level.
template <int BLOCK_N, bool OUT3, bool RES, bool POOL = false, bool GRAM = false, bool UPS = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv3x3_halo_k(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
               const float* __restrict__ bias, const __nv_bfloat16* __restrict__ mask,
               __nv_bfloat16* __restrict__ y, float* __restrict__ y3, ConvShape s, HaloCfg cfg,
               __nv_bfloat16* __restrict__ ypool, const __grid_constant__ CUtensorMap map_f,
               const __grid_constant__ CUtensorMap map_g, const __nv_bfloat16* __restrict__ ups_mask) {
  constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
  // accumulator ring in TMEM: the round trip MMA-complete -> tfull -> epilogue -> tempty -> next MMA costs about
  // 4000 cycles (measured: with two buffers every tile took >= 2000 cycles even with 1/9 of the MMAs and no
  // epilogue work), longer than a whole tile's MMAs at BLOCK_N <= 128, so all 512 columns are used as a ring
  constexpr int NACC = HaloAcc<BLOCK_N>::value;
  constexpr int TCOLS = NACC * BLOCK_N;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(16) float sbias[512];              // the layer's bias, read as broadcast float4s
  if (bias)
    for (int i = threadIdx.x; i < s.Cout && i < 512; i += blockDim.x) sbias[i] = bias[i];
  constexpr bool resident = RES;                        // weights resident in smem (cfg.sb == 0) or streamed
  // resident + split: 2 kc patch loads per tile (hi and lo of every channel group), the MMAs of a hi chunk in pairs that
  // share their A slice through the collector buffer; streamed weights keep the three-pass order (their ring runs 8 taps
  // ahead, and pairing two weight tiles per tap halves that distance: +25 us per step when it was tried there)
  const bool pairs = resident && s.split;
  const int kchunks = pairs ? 2 * s.kc : s.nchunks;
  const int n_wslots = resident ? 9 * s.wchunks : cfg.sb;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a = base;
  const uint32_t smem_b = base + cfg.sa * PATCH_STRIDE;
  const uint32_t bars = smem_b + n_wslots * B_BYTES;
  const uint32_t bar_afull = bars, bar_aempty = bars + 8 * cfg.sa;
  const uint32_t bar_bfull = bar_aempty + 8 * cfg.sa, bar_bempty = bar_bfull + 8 * (resident ? 1 : cfg.sb);
  const uint32_t bar_tfull = bar_bempty + 8 * (resident ? 1 : cfg.sb), bar_tempty = bar_tfull + 8 * NACC;
  const uint32_t tmem_slot = bar_tempty + 8 * NACC;
  const uint32_t stage_base = (tmem_slot + 16 + 127u) & ~127u;   // 4 epilogue warps x 4 KiB (warp_rows_store)
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_sp = s.tiles_w * s.tiles_h;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_x);
    prefetch_tmap(&map_w);
    for (int i = 0; i < cfg.sa; ++i) { mbar_init(bar_afull + 8 * i, 1); mbar_init(bar_aempty + 8 * i, 1); }
    for (int i = 0; i < (resident ? 1 : cfg.sb); ++i) { mbar_init(bar_bfull + 8 * i, 1); mbar_init(bar_bempty + 8 * i, 1); }
    for (int i = 0; i < NACC; ++i) { mbar_init(bar_tfull + 8 * i, 1); mbar_init(bar_tempty + 8 * i, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TCOLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot_ptr;

  if (warp == 0) {
    // ===== TMA producer (whole warp in the loop, one elected lane issues) =====
    const bool leader = elect_one();
    if (resident && leader) {                           // all 9 x kchunks weight tiles, once (n0 = 0)
      mbar_expect_tx(bar_bfull, 9 * s.wchunks * B_BYTES);
      for (int tap = 0; tap < 9; ++tap)
        for (int c = 0; c < s.wchunks; ++c)
          tma_load_3d(smem_b + (tap * s.wchunks + c) * B_BYTES, &map_w, bar_bfull, c * BLOCK_K, 0, tap);
    }
    uint32_t sta = 0, pha = 0, stb = 0, phb = 0;
    for (int t = blockIdx.x; t < cfg.n_tiles; t += gridDim.x) {
      const int nb = t % cfg.n_blocks_n, sp = t / cfg.n_blocks_n;
      const int img = sp / tiles_sp, rem = sp - img * tiles_sp;
      const int th = rem / s.tiles_w, tw = rem - th * s.tiles_w;
      const int h0 = th * HTH, w0 = tw * HTW, n0 = nb * BLOCK_N;
      for (int c = 0; c < kchunks; ++c) {
        // resident weights, split operands: hi_0, lo_0, hi_1, lo_1, ... (a hi chunk meets W_hi and W_lo back to back)
        const int xc = pairs ? ((c & 1) ? s.kc + (c >> 1) : (c >> 1)) : xchunk(s, c);
        mbar_wait(bar_aempty + 8 * sta, pha ^ 1);
        if (leader) {
          mbar_expect_tx(bar_afull + 8 * sta, PATCH_BYTES);
          tma_load_4d(smem_a + sta * PATCH_STRIDE, &map_x, bar_afull + 8 * sta, xc * BLOCK_K, w0 - 1, h0 - 1, img);
        }
        if (++sta == (uint32_t)cfg.sa) { sta = 0; pha ^= 1; }
        if (!resident) {
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(bar_bempty + 8 * stb, phb ^ 1);
            if (leader) {
              mbar_expect_tx(bar_bfull + 8 * stb, B_BYTES);
              tma_load_3d(smem_b + stb * B_BYTES, &map_w, bar_bfull + 8 * stb, wchunk(s, c) * BLOCK_K, n0, tap);
            }
            if (++stb == (uint32_t)cfg.sb) { stb = 0; phb ^= 1; }
          }
        }
      }
      if (GRAM && !resident) {
        // F_hi x Gd_hi, F_lo x Gd_hi, F_hi x Gd_lo: chunk c reads feature chunk fx(c) and matrix chunk gw(c)
        const int gk = s.gram_kc;
        for (int c = 0; c < 3 * gk; ++c) {
          const int fx = c >= 2 * gk ? c - 2 * gk : c, gw = c >= gk ? c - gk : c;
          mbar_wait(bar_aempty + 8 * sta, pha ^ 1);
          if (leader) {
            mbar_expect_tx(bar_afull + 8 * sta, PATCH_BYTES);
            tma_load_4d(smem_a + sta * PATCH_STRIDE, &map_f, bar_afull + 8 * sta, fx * BLOCK_K, w0 - 1, h0 - 1, img);
          }
          if (++sta == (uint32_t)cfg.sa) { sta = 0; pha ^= 1; }
          mbar_wait(bar_bempty + 8 * stb, phb ^ 1);
          if (leader) {
            mbar_expect_tx(bar_bfull + 8 * stb, B_BYTES);
            tma_load_3d(smem_b + stb * B_BYTES, &map_g, bar_bfull + 8 * stb, gw * BLOCK_K, n0, img);
          }
          if (++stb == (uint32_t)cfg.sb) { stb = 0; phb ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (whole warp in the loop, one elected lane issues) =====
    const bool leader = elect_one();
    const uint32_t idesc = umma_idesc_bf16(BLOCK_M, BLOCK_N);
    const uint32_t ahi = desc_hi_sw128((HTW + 2) * 128), bhi = desc_hi_sw128(1024);
    const uint32_t alo_base = desc_lo(smem_a, 16), blo_base = desc_lo(smem_b, 16);
    uint32_t sta = 0, pha = 0, stb = 0, phb = 0, buf = 0, bph = 0;
    if (resident) {
      mbar_wait(bar_bfull, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    for (int t = blockIdx.x; t < cfg.n_tiles; t += gridDim.x) {
      mbar_wait(bar_tempty + 8 * buf, bph ^ 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t acc = tmem_d + buf * BLOCK_N;
      for (int c = 0; c < kchunks; ++c) {
        mbar_wait(bar_afull + 8 * sta, pha);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t alo0 = alo_base + sta * (PATCH_STRIDE >> 4);
        if (resident) {
          // one straight-line block of 36 MMAs per chunk: the issuing warp is alone on its scheduler, every
          // instruction costs it ~4-5 cycles, and a tile's MMAs only take 32-64 cycles each -- the per-tap
          // branches and barrier bookkeeping of the streamed path (~50 instructions per tap) capped the
          // kernel at ~2000 cycles per chunk whatever the MMA count (knock-out experiments)
          if (leader) {
            const uint32_t bstep = (uint32_t)s.wchunks * (B_BYTES >> 4);
            const int wc = pairs ? (c >> 1) : wchunk(s, c);            // W_hi chunk of this channel group
            uint32_t blo = blo_base + wc * (B_BYTES >> 4);             // slot tap * wchunks + wc
            if (pairs && !(c & 1)) {                                   // hi chunk: x_hi * W_hi, x_hi * W_lo
              const uint32_t lo_off = (uint32_t)s.kc * (B_BYTES >> 4);
#pragma unroll
              for (int tap = 0; tap < 9; ++tap) {
                const int ky = tap / 3, kx = tap - 3 * ky;             // compile-time after unrolling
                const uint32_t alo = alo0 + ((ky * (HTW + 2) + kx) * 128 >> 4);
#pragma unroll
                for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                  umma_bf16_lh_keep(acc, alo + k * 2, ahi, blo + k * 2, bhi, idesc, (tap | k) != 0 ? 1u : (c != 0 ? 1u : 0u));
                  umma_bf16_lh_reuse(acc, alo + k * 2, ahi, blo + lo_off + k * 2, bhi, idesc);
                }
                blo += bstep;
              }
            } else {
#pragma unroll
              for (int tap = 0; tap < 9; ++tap) {
                const int ky = tap / 3, kx = tap - 3 * ky;             // compile-time after unrolling
                const uint32_t alo = alo0 + ((ky * (HTW + 2) + kx) * 128 >> 4);
#pragma unroll
                for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                  umma_bf16_lh(acc, alo + k * 2, ahi, blo + k * 2, bhi, idesc, (tap | k) != 0 ? 1u : (c != 0 ? 1u : 0u));
                blo += bstep;
              }
            }
          }
        } else {
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const int ky = tap / 3, kx = tap - 3 * ky;
            mbar_wait(bar_bfull + 8 * stb, phb);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t blo = blo_base + stb * (B_BYTES >> 4);
            const uint32_t alo = alo0 + ((ky * (HTW + 2) + kx) * 128 >> 4);
            if (leader) {
#pragma unroll
              for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                umma_bf16_lh(acc, alo + k * 2, ahi, blo + k * 2, bhi, idesc, (c | tap | k) != 0 ? 1u : 0u);
              umma_commit(bar_bempty + 8 * stb);
            }
            if (++stb == (uint32_t)cfg.sb) { stb = 0; phb ^= 1; }
          }
        }
        if (leader) umma_commit(bar_aempty + 8 * sta);
        if (++sta == (uint32_t)cfg.sa) { sta = 0; pha ^= 1; }
      }
      if (GRAM && !resident) {
        for (int c = 0; c < 3 * s.gram_kc; ++c) {
          mbar_wait(bar_afull + 8 * sta, pha);
          mbar_wait(bar_bfull + 8 * stb, phb);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t alo = alo_base + sta * (PATCH_STRIDE >> 4) + (((HTW + 2) + 1) * 128 >> 4);   // centre tap
          const uint32_t blo = blo_base + stb * (B_BYTES >> 4);
          if (leader) {
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
              umma_bf16_lh(acc, alo + k * 2, ahi, blo + k * 2, bhi, idesc, 1u);
            umma_commit(bar_bempty + 8 * stb);
            umma_commit(bar_aempty + 8 * sta);
          }
          if (++stb == (uint32_t)cfg.sb) { stb = 0; phb ^= 1; }
          if (++sta == (uint32_t)cfg.sa) { sta = 0; pha ^= 1; }
        }
      }
      if (leader) umma_commit(bar_tfull + 8 * buf);
      if (++buf == NACC) { buf = 0; bph ^= 1; }
    }
  } else {
    // ===== epilogue: warps 2..5, TMEM sub-partition = warp % 4 =====
    const int q = warp & 3;
    const int r = q * 32 + lane;
    uint32_t buf = 0, bph = 0;
    for (int t = blockIdx.x; t < cfg.n_tiles; t += gridDim.x) {
      const int nb = t % cfg.n_blocks_n, sp = t / cfg.n_blocks_n;
      const int img = sp / tiles_sp, rem = sp - img * tiles_sp;
      const int th = rem / s.tiles_w, tw = rem - th * s.tiles_w;
      const int n0 = nb * BLOCK_N;
      const int ph_ = th * HTH + (r >> 3), pw_ = tw * HTW + (r & 7);
      const bool valid = ph_ < s.H && pw_ < s.W;
      const int64_t pix = ((int64_t)img * s.H + ph_) * s.W + pw_;
      const uint32_t acc = tmem_d + buf * BLOCK_N + ((uint32_t)(q * 32) << 16);
      if (OUT3) {
        mbar_wait(bar_tfull + 8 * buf, bph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t v[16];
        tmem_ld16(acc, v);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
        if (valid) {
          if (s.out_ch == 1) {
            y3[pix] = __uint_as_float(v[0]) * s.scale;
          } else {
            float* o = y3 + pix * 3;
            o[0] = __uint_as_float(v[0]) * s.scale; o[1] = __uint_as_float(v[1]) * s.scale; o[2] = __uint_as_float(v[2]) * s.scale;
          }
        }
      } else {
        constexpr int NCH = (BLOCK_N >= 32) ? BLOCK_N / 32 : 1;
        uint32_t mbits[NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) mbits[c] = 0xffffffffu;
        if (mask && valid) {
          const uint4* msk = reinterpret_cast<const uint4*>(mask + pix * s.ldm + n0);
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            uint4 mv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) mv[j] = msk[c * 4 + j];
            uint32_t bits = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const __nv_bfloat16* mh = reinterpret_cast<const __nv_bfloat16*>(&mv[j]);
#pragma unroll
              for (int e = 0; e < 8; ++e) bits |= (__bfloat162float(mh[e]) > 0.f ? 1u : 0u) << (j * 8 + e);
            }
            mbits[c] = bits;
          }
        }
        mbar_wait(bar_tfull + 8 * buf, bph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // whole accumulator row into registers (1 CTA/SM: registers are plentiful), then hand the TMEM
        // buffer straight back: the MMAs of the tile after next never wait for this tile's stores
        uint32_t v[NCH * 32];
#pragma unroll
        for (int c = 0; c < NCH; ++c) tmem_ld32_nowait(acc + (uint32_t)(c * 32), v + c * 32);
        tmem_ld_wait();
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
        {
          constexpr int NV = HaloStage<BLOCK_N>::nv;                 // 16-byte pieces per staged row chunk
          constexpr int NCK = (BLOCK_N >= NV * 8) ? BLOCK_N / (NV * 8) : 1;
#pragma unroll
          for (int ck = 0; ck < NCK; ++ck) {
            uint4 ov[NV], ol[NV];
#pragma unroll
            for (int jj = 0; jj < NV; ++jj) {
              const int col = ck * NV * 8 + jj * 8;                  // first of 8 accumulator columns
              const int c = col >> 5, j = (col >> 3) & 3;            // TMEM chunk of 32 columns, 8-column group
              __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&ov[jj]);
              __nv_bfloat162* ohl = reinterpret_cast<__nv_bfloat162*>(&ol[jj]);
              const float4 b0 = bias ? *reinterpret_cast<const float4*>(sbias + n0 + col) : make_float4(0, 0, 0, 0);
              const float4 b1 = bias ? *reinterpret_cast<const float4*>(sbias + n0 + col + 4) : make_float4(0, 0, 0, 0);
              const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
              for (int e = 0; e < 8; e += 2) {
                float f0 = fmaf(__uint_as_float(v[col + e]), s.scale, bb[e]);
                float f1 = fmaf(__uint_as_float(v[col + e + 1]), s.scale, bb[e + 1]);
                if (s.relu) { f0 = fmaxf(f0, 0.f); f1 = fmaxf(f1, 0.f); }
                if (!((mbits[c] >> (j * 8 + e)) & 1u)) f0 = 0.f;
                if (!((mbits[c] >> (j * 8 + e + 1)) & 1u)) f1 = 0.f;
                if (s.split) split2(f0, f1, oh[e >> 1], ohl[e >> 1]);
                else oh[e >> 1] = __floats2bfloat162_rn(f0, f1);
              }
            }
            if (POOL && NV == 4) {
              // fused 2x2 average pool of the split output (avgpool2_split_fwd_k): a warp is 4 tile rows x 8 columns, so
              // a quad is lanes {l, l^1, l^8, l^9}.  Reduce-scatter over the quad: after the x step a lane keeps 16 of
              // the chunk's 32 channels, after the y step 8, fully summed -- 24 shuffles per chunk instead of 64 -- and
              // writes them.  Values are the ROUNDED ones (hi + lo), summed (a + b) + (c + d) like the stand-alone kernel.
              float fr[32];
#pragma unroll
              for (int jj = 0; jj < NV; ++jj) {
                const __nv_bfloat162* oh = reinterpret_cast<const __nv_bfloat162*>(&ov[jj]);
                const __nv_bfloat162* ohl = reinterpret_cast<const __nv_bfloat162*>(&ol[jj]);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 h = __bfloat1622float2(oh[e]), l = __bfloat1622float2(ohl[e]);
                  fr[jj * 8 + 2 * e] = h.x + l.x;
                  fr[jj * 8 + 2 * e + 1] = h.y + l.y;
                }
              }
              const bool dx = (lane & 1) != 0, dy = (lane & 8) != 0;
              float k1[16], k2[8];
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                const float recv = __shfl_xor_sync(0xffffffffu, dx ? fr[e] : fr[16 + e], 1);
                k1[e] = (dx ? fr[16 + e] : fr[e]) + recv;
              }
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float recv = __shfl_xor_sync(0xffffffffu, dy ? k1[e] : k1[8 + e], 8);
                k2[e] = ((dy ? k1[8 + e] : k1[e]) + recv) * 0.25f;
              }
              const int OHp = s.H >> 1, OWp = s.W >> 1;
              const int py = ph_ >> 1, px = pw_ >> 1;
              if (py < OHp && px < OWp) {
                uint4 ph4, pl4;
                __nv_bfloat162* a = reinterpret_cast<__nv_bfloat162*>(&ph4);
                __nv_bfloat162* b = reinterpret_cast<__nv_bfloat162*>(&pl4);
#pragma unroll
                for (int e = 0; e < 8; e += 2) split2(k2[e], k2[e + 1], a[e >> 1], b[e >> 1]);
                __nv_bfloat16* dstp = ypool + (((int64_t)img * OHp + py) * OWp + px) * s.ldy + n0 + ck * 32 +
                                      (dx ? 16 : 0) + (dy ? 8 : 0);
                *reinterpret_cast<uint4*>(dstp) = ph4;
                *reinterpret_cast<uint4*>(dstp + s.Cout) = pl4;
              }
            }
            if (UPS && NV == 4) {
              // (hi + lo) * 0.25 per value, like the stand-alone kernel; y is the FINE-level gradient [n, 2H, 2W, 2 Cout]
              const uint32_t stg = stage_base + (uint32_t)q * 4096u;
#pragma unroll
              for (int jj = 0; jj < NV; ++jj) {
                const __nv_bfloat162* oh = reinterpret_cast<const __nv_bfloat162*>(&ov[jj]);
                const __nv_bfloat162* ohl = reinterpret_cast<const __nv_bfloat162*>(&ol[jj]);
                float fr[8];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 h = __bfloat1622float2(oh[e]), l = __bfloat1622float2(ohl[e]);
                  fr[2 * e] = (h.x + l.x) * 0.25f;
                  fr[2 * e + 1] = (h.y + l.y) * 0.25f;
                }
                st_shared_v4(stg + lane * 128 + ((uint32_t)(((2 * jj) ^ lane) & 7) << 4),
                             make_uint4(__float_as_uint(fr[0]), __float_as_uint(fr[1]), __float_as_uint(fr[2]), __float_as_uint(fr[3])));
                st_shared_v4(stg + lane * 128 + ((uint32_t)(((2 * jj + 1) ^ lane) & 7) << 4),
                             make_uint4(__float_as_uint(fr[4]), __float_as_uint(fr[5]), __float_as_uint(fr[6]), __float_as_uint(fr[7])));
              }
              __syncwarp();
              const int sub = lane >> 2, pj = lane & 3;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int row = i * 8 + sub, rr = q * 32 + row;
                const int plh = th * HTH + (rr >> 3), plw = tw * HTW + (rr & 7);
                const uint4 x0 = ld_shared_v4(stg + row * 128 + ((uint32_t)(((2 * pj) ^ row) & 7) << 4));
                const uint4 x1 = ld_shared_v4(stg + row * 128 + ((uint32_t)(((2 * pj + 1) ^ row) & 7) << 4));
                const float xv[8] = {__uint_as_float(x0.x), __uint_as_float(x0.y), __uint_as_float(x0.z), __uint_as_float(x0.w),
                                     __uint_as_float(x1.x), __uint_as_float(x1.y), __uint_as_float(x1.z), __uint_as_float(x1.w)};
                if (plh < s.H && plw < s.W) {
                  const int co = n0 + ck * 32 + pj * 8;
                  uint4 mv[4];
                  int64_t off[4];
#pragma unroll
                  for (int qd = 0; qd < 4; ++qd) {
                    off[qd] = ((((int64_t)img * 2 * s.H + 2 * plh + (qd >> 1)) * (2 * s.W)) + 2 * plw + (qd & 1)) * s.ldy + co;
                    mv[qd] = *reinterpret_cast<const uint4*>(ups_mask + off[qd]);      // hi half carries the sign
                  }
#pragma unroll
                  for (int qd = 0; qd < 4; ++qd) {
                    const __nv_bfloat16* mh = reinterpret_cast<const __nv_bfloat16*>(&mv[qd]);
                    uint4 oh4, ol4;
                    __nv_bfloat162* a = reinterpret_cast<__nv_bfloat162*>(&oh4);
                    __nv_bfloat162* b = reinterpret_cast<__nv_bfloat162*>(&ol4);
#pragma unroll
                    for (int e = 0; e < 8; e += 2)
                      split2(__bfloat162float(mh[e]) > 0.f ? xv[e] : 0.f, __bfloat162float(mh[e + 1]) > 0.f ? xv[e + 1] : 0.f,
                             a[e >> 1], b[e >> 1]);
                    *reinterpret_cast<uint4*>(y + off[qd]) = oh4;
                    *reinterpret_cast<uint4*>(y + off[qd] + s.Cout) = ol4;
                  }
                }
              }
              __syncwarp();
            } else if (BLOCK_N < 128) {                              // narrow tiles: the direct stores keep up with the MMAs
              if (valid) {
                __nv_bfloat16* dst = y + pix * s.ldy + n0 + ck * NV * 8;
#pragma unroll
                for (int jj = 0; jj < NV; ++jj) *reinterpret_cast<uint4*>(dst + jj * 8) = ov[jj];
                if (s.split) {
#pragma unroll
                  for (int jj = 0; jj < NV; ++jj) *reinterpret_cast<uint4*>(dst + s.Cout + jj * 8) = ol[jj];
                }
              }
            } else {
              warp_rows_store<NV>(stage_base + (uint32_t)q * (NV * 512u), lane, ov, [&](int row) -> __nv_bfloat16* {
                const int rr = q * 32 + row;
                const int ph2 = th * HTH + (rr >> 3), pw2 = tw * HTW + (rr & 7);
                if (ph2 >= s.H || pw2 >= s.W) return nullptr;
                return y + (((int64_t)img * s.H + ph2) * s.W + pw2) * s.ldy + n0 + ck * NV * 8;
              });
              if (s.split)
                warp_rows_store<NV>(stage_base + (uint32_t)q * (NV * 512u), lane, ol, [&](int row) -> __nv_bfloat16* {
                  const int rr = q * 32 + row;
                  const int ph2 = th * HTH + (rr >> 3), pw2 = tw * HTW + (rr & 7);
                  if (ph2 >= s.H || pw2 >= s.W) return nullptr;
                  return y + (((int64_t)img * s.H + ph2) * s.W + pw2) * s.ldy + s.Cout + n0 + ck * NV * 8;
                });
            }
          }
        }
      }
      if (++buf == NACC) { buf = 0; bph ^= 1; }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(TCOLS));
  }
}

// ---- descriptor probe (developer tool, tools/umma_probe.py) -------------------------------------------
// One CTA: TMA-load a (TH+2) x (TW+2) = 18 x 10 pixel patch of 64 channels (rows of 128 B, SWIZZLE_128B)
// either densely (row pitch 10) or one image row per 2048 B slot (row pitch 16), then run ONE
// M=128,N=16,K=64 MMA whose A operand is tap (ky,kx) of the patch, addressed only through the
// descriptor (start offset, SBO = pitch, base_offset).  Tells which addressing the hardware accepts
// for reading the 9 taps of a 3x3 convolution out of one halo'd patch.
__global__ void __launch_bounds__(128, 1)
umma_probe_k(const __grid_constant__ CUtensorMap map_dense, const __grid_constant__ CUtensorMap map_row,
             const __grid_constant__ CUtensorMap map_b, float* __restrict__ out, int pitched, int ky, int kx,
             int base_mode) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a = base;                       // up to 18 * 2048 = 36 KiB
  const uint32_t smem_b = base + 40 * 1024;           // 16 rows x 128 B
  const uint32_t bars = smem_b + 2048;
  const uint32_t tmem_slot = bars + 16;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bars, 1);
    mbar_init(bars + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(32));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot_ptr;
  const int pitch_rows = pitched ? 16 : 10;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bars, 180 * 128 + 16 * 128);
    if (!pitched) {
      tma_load_2d(smem_a, &map_dense, bars, 0, 0);
    } else {
      for (int r = 0; r < 18; ++r) tma_load_2d(smem_a + r * 2048, &map_row, bars, 0, r * 10);
    }
    tma_load_2d(smem_b, &map_b, bars, 0, 0);
    mbar_wait(bars, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = umma_idesc_bf16(128, 16);
    for (int k = 0; k < 4; ++k) {
      const uint32_t a = smem_a + (ky * pitch_rows + kx) * 128 + k * 32;
      const uint32_t bo = base_mode == 1 ? ((a >> 7) & 7u) : (base_mode == 2 ? (uint32_t)kx : 0u);
      umma_bf16(tmem_d, umma_desc_sw128_ex(a, pitch_rows * 128, bo), umma_desc_sw128(smem_b + k * 32), idesc,
                k != 0 ? 1u : 0u);
    }
    umma_commit(bars + 8);
  }
  __syncwarp();
  mbar_wait(bars + 8, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t v[16];
  tmem_ld16(tmem_d + ((uint32_t)(warp * 32) << 16), v);
  const int row = warp * 32 + lane;
  for (int j = 0; j < 16; ++j) out[row * 16 + j] = __uint_as_float(v[j]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(32));
  }
}

// ---- Gram matrix G = F^T F on tensor cores ---------------------------------------------------
// F bf16 [n, P, C] is the NHWC activation seen as P = h*w rows of C channels.  Both operands of
// F^T F are "MN-major" for the MMA (the contraction index p is the slow one in memory), so the
// TMA box {64 channels, 64 pixels} lands as 64 K-rows x 128 B and is consumed through an
// MN-major SWIZZLE_128B descriptor: 64-channel groups are LBO = 8 KiB apart, 8-pixel groups
// SBO = 1 KiB apart, one K=16 MMA step advances 2 KiB.  The pixel range is split over CTAs
// (split-K); partial 128x128 tiles are reduced with coalesced fp32 atomics, written transposed
// (G is symmetric) so that consecutive TMEM lanes hit consecutive addresses.
constexpr int G_STAGES = 4;
constexpr int G_BOX_BYTES = 64 * 128;             // one {64 ch, 64 px} box
constexpr int G_OP_BYTES = 2 * G_BOX_BYTES;       // 128 channels x 64 pixels

__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(G_BOX_BYTES >> 4) << 16;        // LBO: next 64-element group along M/N
  d |= (uint64_t)(1024 >> 4) << 32;               // SBO: next 8-row group along K
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
gram_tc_k(const __grid_constant__ CUtensorMap map_f, float* __restrict__ G, int P, int C, int tiles_n,
          int k_per_split, int sym) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a = base;
  const uint32_t smem_b = base + G_STAGES * G_OP_BYTES;
  const uint32_t bars = smem_b + G_STAGES * G_OP_BYTES;
  const uint32_t tmem_slot = bars + 8 * (2 * G_STAGES + 1);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int m0 = (blockIdx.x / tiles_n) * 128, n0 = (blockIdx.x % tiles_n) * 128;
  if (sym) {                                             // F^T F is symmetric: only the tiles with m0 <= n0 are launched
    int i = 0, rem = blockIdx.x;                         // (the epilogue stores them transposed: row block >= column block)
    while (rem >= tiles_n - i) { rem -= tiles_n - i; ++i; }
    m0 = i * 128; n0 = (i + rem) * 128;
  }
  const int img = blockIdx.z;
  const int p_beg = blockIdx.y * k_per_split;
  const int p_end = min(P, p_beg + k_per_split);
  const int num_kb = (p_end - p_beg + 63) / 64;
  const bool diag = (m0 == n0);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_f);
    for (int i = 0; i < G_STAGES; ++i) {
      mbar_init(bars + 8 * i, 1);
      mbar_init(bars + 8 * (G_STAGES + i), 1);
    }
    mbar_init(bars + 8 * (2 * G_STAGES), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int st = kb % G_STAGES;
        const uint32_t ph = (kb / G_STAGES) & 1;
        mbar_wait(bars + 8 * (G_STAGES + st), ph ^ 1);
        const uint32_t full = bars + 8 * st;
        mbar_expect_tx(full, diag ? G_OP_BYTES : 2 * G_OP_BYTES);
        const int p0 = p_beg + kb * 64;
        // rows beyond P (or beyond this split's share of a partial block) must not be counted twice:
        // k_per_split is a multiple of 64, so only the global tail is partial and TMA zero-fills it.
        tma_load_3d(smem_a + st * G_OP_BYTES, &map_f, full, m0, p0, img);
        tma_load_3d(smem_a + st * G_OP_BYTES + G_BOX_BYTES, &map_f, full, m0 + 64, p0, img);
        if (!diag) {
          tma_load_3d(smem_b + st * G_OP_BYTES, &map_f, full, n0, p0, img);
          tma_load_3d(smem_b + st * G_OP_BYTES + G_BOX_BYTES, &map_f, full, n0 + 64, p0, img);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // kind::f16, BF16 x BF16 -> F32, both operands MN-major (bits 15 and 16), M = N = 128
      const uint32_t idesc = umma_idesc_bf16(128, 128) | (1u << 15) | (1u << 16);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int st = kb % G_STAGES;
        const uint32_t ph = (kb / G_STAGES) & 1;
        mbar_wait(bars + 8 * st, ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a0 = smem_a + st * G_OP_BYTES;
        const uint32_t b0 = diag ? a0 : smem_b + st * G_OP_BYTES;
        const uint32_t alo = desc_lo(a0, G_BOX_BYTES), blo = desc_lo(b0, G_BOX_BYTES);   // LBO: next 64-ch group
        const uint32_t hi = desc_hi_sw128(1024);                                        // SBO: next 8 pixels
#pragma unroll
        for (int k = 0; k < 64 / UMMA_K; ++k)
          umma_bf16_lh(tmem_d, alo + k * (2048 >> 4), hi, blo + k * (2048 >> 4), hi, idesc, (kb | k) != 0 ? 1u : 0u);
        umma_commit(bars + 8 * (G_STAGES + st));
      }
      umma_commit(bars + 8 * (2 * G_STAGES));
    }
  } else if (num_kb > 0) {
    const int q = warp & 3;
    const int m = m0 + q * 32 + lane;
    mbar_wait(bars + 8 * (2 * G_STAGES), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float* Gi = G + (int64_t)img * C * C;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t v[32];
      tmem_ld32(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
      if (m < C) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int nn = n0 + c * 32 + j;
          if (nn < C) atomicAdd(Gi + (int64_t)nn * C + m, __uint_as_float(v[j]));   // transposed: G symmetric
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(128));
  }
}

// Split features F' = [hi | lo] (bf16 [n, P, 2C]): G = hi^T hi + hi^T lo + lo^T hi accumulated in ONE TMEM tile (the lo^T lo
// term, 2^-16 relative, is dropped like in the convolutions).  The 2C x 2C Gram of the split rows (gram_tc_k on F') costs
// four 128 x 128 products per output tile, re-loads every slab once per product and leaves the sum to a finishing kernel
// that reads mirrored blocks; here a k-block brings the hi and lo slabs of the row block (and of the column block off
// the diagonal) ONCE -- 32 / 64 KiB per 64 pixels, three stages in flight -- and feeds 12 MMAs.  C % 128 == 0.
// Output tile (i <= j) is stored transposed at rows of block j, columns of block i, like gram_tc_k (sym).
constexpr int G3_STAGES = 3;
constexpr int G3_STAGE_BYTES = 4 * G_OP_BYTES;    // hi_i, lo_i, hi_j, lo_j: 128 channels x 64 pixels each
__global__ void __launch_bounds__(NUM_THREADS, 1)
gram_split_tc_k(const __grid_constant__ CUtensorMap map_f, float* __restrict__ G, int P, int C, int tiles_n,
                int k_per_split) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + G3_STAGES * G3_STAGE_BYTES;
  const uint32_t tmem_slot = bars + 8 * (2 * G3_STAGES + 1);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int bi = 0, rem = blockIdx.x;
  while (rem >= tiles_n - bi) { rem -= tiles_n - bi; ++bi; }
  const int m0 = bi * 128, n0 = (bi + rem) * 128;
  const int img = blockIdx.z;
  const int p_beg = blockIdx.y * k_per_split;
  const int p_end = min(P, p_beg + k_per_split);
  const int num_kb = (p_end - p_beg + 63) / 64;
  const bool diag = (m0 == n0);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_f);
    for (int i = 0; i < G3_STAGES; ++i) {
      mbar_init(bars + 8 * i, 1);
      mbar_init(bars + 8 * (G3_STAGES + i), 1);
    }
    mbar_init(bars + 8 * (2 * G3_STAGES), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot_ptr;

  if (warp == 0) {
    const bool leader = elect_one();
    uint32_t st = 0, ph = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(bars + 8 * (G3_STAGES + st), ph ^ 1);
      if (leader) {
        const uint32_t full = bars + 8 * st;
        const uint32_t dst = base + st * G3_STAGE_BYTES;
        mbar_expect_tx(full, (diag ? 2 : 4) * G_OP_BYTES);
        const int p0 = p_beg + kb * 64;
#pragma unroll
        for (int h = 0; h < 2; ++h) {                      // h = 0: hi (channels [0,C)), 1: lo ([C,2C))
          tma_load_3d(dst + h * G_OP_BYTES, &map_f, full, h * C + m0, p0, img);
          tma_load_3d(dst + h * G_OP_BYTES + G_BOX_BYTES, &map_f, full, h * C + m0 + 64, p0, img);
          if (!diag) {
            tma_load_3d(dst + (2 + h) * G_OP_BYTES, &map_f, full, h * C + n0, p0, img);
            tma_load_3d(dst + (2 + h) * G_OP_BYTES + G_BOX_BYTES, &map_f, full, h * C + n0 + 64, p0, img);
          }
        }
      }
      if (++st == G3_STAGES) { st = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    const bool leader = elect_one();
    const uint32_t idesc = umma_idesc_bf16(128, 128) | (1u << 15) | (1u << 16);      // both operands MN-major
    const uint32_t hi = desc_hi_sw128(1024);
    uint32_t st = 0, ph = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(bars + 8 * st, ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (leader) {
        const uint32_t a_hi = desc_lo(base + st * G3_STAGE_BYTES, G_BOX_BYTES);
        const uint32_t a_lo = a_hi + (G_OP_BYTES >> 4);
        const uint32_t b_hi = diag ? a_hi : a_hi + (2 * G_OP_BYTES >> 4);
        const uint32_t b_lo = b_hi + (G_OP_BYTES >> 4);
#pragma unroll
        for (int k = 0; k < 64 / UMMA_K; ++k) {
          const uint32_t o = k * (2048 >> 4);
          umma_bf16_lh(tmem_d, a_hi + o, hi, b_hi + o, hi, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_bf16_lh(tmem_d, a_hi + o, hi, b_lo + o, hi, idesc, 1u);
          umma_bf16_lh(tmem_d, a_lo + o, hi, b_hi + o, hi, idesc, 1u);
        }
        umma_commit(bars + 8 * (G3_STAGES + st));
      }
      if (++st == G3_STAGES) { st = 0; ph ^= 1; }
    }
    if (leader) umma_commit(bars + 8 * (2 * G3_STAGES));
  } else if (num_kb > 0) {
    const int q = warp & 3;
    const int m = m0 + q * 32 + lane;
    mbar_wait(bars + 8 * (2 * G3_STAGES), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float* Gi = G + (int64_t)img * C * C;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t v[32];
      tmem_ld32(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
#pragma unroll
      for (int j = 0; j < 32; ++j)
        atomicAdd(Gi + (int64_t)(n0 + c * 32 + j) * C + m, __uint_as_float(v[j]));     // transposed: G symmetric
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(128));
  }
}

// finishing pass of gram_split_tc_k: Graw [n, C, C] holds the blocks with row block >= column block.  One block per 1024
// elements (four independent loads per thread in flight), one loss atomic per block (per-warp atomics onto the nine
// loss slots serialised: ~10 us for a 0.6-2.4 MB pass).
__global__ void __launch_bounds__(256) gram_finish_split3_k(const float* __restrict__ Graw, float* __restrict__ G,
                                                            const float* __restrict__ Gs, __nv_bfloat16* __restrict__ Gd2,
                                                            int C, float inv_denom, float weight, float* __restrict__ loss,
                                                            float gd_scale) {
  __shared__ float part[8];
  const int img = blockIdx.y;
  const int n_el = C * C;
  const float* g = Graw + (int64_t)img * n_el;
  float d[4];
  int idx[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int i = blockIdx.x * 1024 + j * 256 + threadIdx.x;
    idx[j] = i;
    d[j] = 0.f;
    if (i < n_el) {
      const int r = i / C, c = i - r * C;
      d[j] = (r >> 7) >= (c >> 7) ? g[i] : g[(int64_t)c * C + r];
    }
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int i = idx[j];
    if (i >= n_el) continue;
    float v = d[j] * inv_denom;
    if (Gs) { v -= Gs[i]; s += v * v; }
    G[(int64_t)img * n_el + i] = v;
    if (Gd2) {
      const int r = i / C, c = i - r * C;
      const int64_t a = (int64_t)r * 2 * C + c;
      const float vs = v * gd_scale;                    // the operand of the gradient GEMM may carry the loss coefficient
      const __nv_bfloat16 h = __float2bfloat16_rn(vs);
      Gd2[(int64_t)img * 2 * n_el + a] = h;
      Gd2[(int64_t)img * 2 * n_el + a + C] = __float2bfloat16_rn(vs - __bfloat162float(h));
    }
  }
  if (loss && Gs) {                                     // uniform branch: every thread of the block takes it
    s = lnst_warp_sum(s);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) tot += part[w];
      if (tot != 0.f) atomicAdd(loss + img, weight * tot);
    }
  }
}

// G <- G/denom - Gs (fp32, in place), bf16 copy for the gradient GEMM, loss[img] += weight * sum(G^2)
__global__ void gram_finish_bf16_k(float* __restrict__ G, const float* __restrict__ Gs, __nv_bfloat16* __restrict__ Gd,
                                   int n_el, float inv_denom, float weight, float* __restrict__ loss) {
  const int img = blockIdx.y;
  float s = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_el; i += gridDim.x * blockDim.x) {
    float d = G[(int64_t)img * n_el + i] * inv_denom;
    if (Gs) { d -= Gs[i]; s += d * d; }
    G[(int64_t)img * n_el + i] = d;
    if (Gd) Gd[(int64_t)img * n_el + i] = __float2bfloat16_rn(d);
  }
  s = lnst_warp_sum(s);
  if (loss && Gs && (threadIdx.x & 31) == 0 && s != 0.f) atomicAdd(loss + img, weight * s);
}

// Split features: G2 = F'^T F' with F' = [hi | lo] is [2C, 2C] = [[hh, hl], [lh, ll]]; F^T F = hh + hl + lh + ll.
// G <- that / denom - Gs (fp32), Gd2 [C, 2C] = [hi | lo] split copy for the gradient GEMM, loss += weight * sum(G^2).
__global__ void gram_finish_split_k(const float* __restrict__ G2, float* __restrict__ G, const float* __restrict__ Gs,
                                    __nv_bfloat16* __restrict__ Gd2, int C, float inv_denom, float weight,
                                    float* __restrict__ loss) {
  const int img = blockIdx.y;
  const int n_el = C * C;
  const float* g2 = G2 + (int64_t)img * 4 * n_el;
  float s = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_el; i += gridDim.x * blockDim.x) {
    const int r = i / C, c = i - r * C;
    const int64_t a = (int64_t)r * 2 * C + c;
    // G2 holds the 128 x 128 blocks with row block >= column block (gram_tc_k, sym): the others are read mirrored
    auto at = [&](int R, int Cc) -> float {
      return (R >> 7) >= (Cc >> 7) ? g2[(int64_t)R * 2 * C + Cc] : g2[(int64_t)Cc * 2 * C + R];
    };
    float d = ((at(r, c) + at(r, C + c)) + (at(C + r, c) + at(C + r, C + c))) * inv_denom;
    if (Gs) { d -= Gs[i]; s += d * d; }
    G[(int64_t)img * n_el + i] = d;
    if (Gd2) {
      const __nv_bfloat16 h = __float2bfloat16_rn(d);
      Gd2[(int64_t)img * 2 * n_el + a] = h;
      Gd2[(int64_t)img * 2 * n_el + a + C] = __float2bfloat16_rn(d - __bfloat162float(h));
    }
  }
  s = lnst_warp_sum(s);
  if (loss && Gs && (threadIdx.x & 31) == 0 && s != 0.f) atomicAdd(loss + img, weight * s);
}

// ---- host side: tensor maps ------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

static bool make_map(CUtensorMap* m, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                     const cuuint32_t* box) {
  cuuint32_t ones[5] = {1, 1, 1, 1, 1};
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides, box, ones,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int conv_persistent = 1;   // tuning switch: 1 = persistent kernel, 0 = one CTA per tile

template <int BLOCK_N>
static int launch_conv(const CUtensorMap& mx, const CUtensorMap& mw, const float* bias, const __nv_bfloat16* mask,
                       const __nv_bfloat16* addend, __nv_bfloat16* y, const ConvShape& s, int n_img,
                       cudaStream_t stream) {
  const int smem = STAGES * (A_BYTES + BLOCK_N * BLOCK_K * 2) + 8 * (2 * STAGES + 1) + 16 + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_tc_k<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  if (conv_persistent) {
    constexpr int PSTAGES = PersistStages<BLOCK_N>::value;
    const int psmem = 8 * 4096 + PSTAGES * (A_BYTES + BLOCK_N * BLOCK_K * 2) + 8 * (2 * PSTAGES + 4) + 16 + 1024;
    static bool pconfigured = false;
    static int sms = 148;
    if (!pconfigured) {
      cudaError_t e = cudaFuncSetAttribute(conv3x3_tc_persist_k<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           psmem);
      if (e != cudaSuccess) return (int)e;
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      pconfigured = true;
    }
    const int n_blocks_n = s.Cout / BLOCK_N;
    const int n_tiles = s.tiles_w * s.tiles_h * n_img * n_blocks_n;
    const int grid = n_tiles < sms ? n_tiles : sms;
    conv3x3_tc_persist_k<BLOCK_N><<<grid, PERSIST_THREADS, psmem, stream>>>(mx, mw, bias, mask, addend, y, s, n_blocks_n,
                                                                        n_tiles);
    return (int)cudaGetLastError();
  }
  dim3 grid(s.tiles_w * s.tiles_h * n_img, s.Cout / BLOCK_N);
  conv3x3_tc_k<BLOCK_N><<<grid, NUM_THREADS, smem, stream>>>(mx, mw, bias, mask, addend, y, s);
  return (int)cudaGetLastError();
}

// choose the TH x TW = 128 patch that wastes the fewest pixels
static void pick_tile(int H, int W, int& TH, int& TW) {
  long best = -1;
  for (int tw = 8; tw <= 128; tw *= 2) {
    const int th = 128 / tw;
    const long cover = (long)((H + th - 1) / th) * th * ((W + tw - 1) / tw) * tw;
    if (best < 0 || cover < best || (cover == best && tw == 16)) { best = cover; TH = th; TW = tw; }
  }
}

static int conv_halo = 2;         // tuning switch: 0 = per-tap kernel, 1 = halo'd-patch kernel where the weights stay
                                  // resident, 2 = halo'd-patch kernel for every 3x3 convolution (weights streamed otherwise)

// smem plan of the halo kernel: weights resident when they fit beside >= 3 patch stages
template <int BLOCK_N, bool OUT3>
static int launch_halo(const void* x, const void* wmat, const float* bias, const __nv_bfloat16* mask,
                       __nv_bfloat16* y, float* y3, int n, int H, int W, int Cin, int Cout, int relu, float scale,
                       cudaStream_t stream, int out_ch = 3, int split = 0, __nv_bfloat16* ypool = nullptr,
                       const void* gram_f = nullptr, const void* gram_g = nullptr, int* gram_fused = nullptr,
                       const __nv_bfloat16* ups_mask = nullptr) {
  constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
  constexpr int MAX_SMEM = 232448;                        // 227 KiB opt-in limit per CTA
  const int stage_bytes = ups_mask ? 4 * 4096 : HaloStage<BLOCK_N>::bytes;             // UPS: fp32 chunks, 4 KiB per epilogue warp
  const int budget = MAX_SMEM - 2048 - 1024 - 512 - stage_bytes - 128;   // static bias table, alignment slack, barriers, store staging
  ConvShape s;
  s.H = H; s.W = W; s.Cin = Cin; s.Cout = Cout; s.relu = relu; s.taps = 9; s.w_img = 0; s.scale = scale; s.out_ch = out_ch;
  s.TH = HTH; s.TW = HTW;
  s.tiles_w = (W + HTW - 1) / HTW;
  s.tiles_h = (H + HTH - 1) / HTH;
  if (split) shape_split(s); else shape_plain(s);
  const int kchunks = s.wchunks;                          // weight chunks per tap (what residency has to hold)
  const int xC = split ? 2 * Cin : Cin;                   // channels of a row of x / of the weight rows
  HaloCfg cfg;
  cfg.n_blocks_n = OUT3 ? 1 : Cout / BLOCK_N;
  cfg.n_tiles = s.tiles_w * s.tiles_h * n * cfg.n_blocks_n;
  const int wres = 9 * kchunks * B_BYTES;
  if (cfg.n_blocks_n == 1 && wres + 3 * PATCH_STRIDE <= budget) {
    cfg.sb = 0;
    cfg.sa = (budget - wres) / PATCH_STRIDE;
  } else {
    cfg.sb = BLOCK_N == 128 ? 8 : 9;
    cfg.sa = (budget - cfg.sb * B_BYTES) / PATCH_STRIDE;
  }
  if (cfg.sa > 6) cfg.sa = 6;
  if (cfg.sa < 2) return LNST_EARG;
  const int n_wslots = cfg.sb == 0 ? 9 * kchunks : cfg.sb;
  const int smem = cfg.sa * PATCH_STRIDE + n_wslots * B_BYTES + 8 * (2 * cfg.sa + 2 * (cfg.sb ? cfg.sb : 1) + 16) + 16 + 1024 + stage_bytes + 128;
  CUtensorMap mx, mw;
  {
    const cuuint64_t dims[4] = {(cuuint64_t)xC, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n};
    const cuuint64_t strides[3] = {(cuuint64_t)xC * 2, (cuuint64_t)W * xC * 2, (cuuint64_t)H * W * xC * 2};
    const cuuint32_t box[4] = {(cuuint32_t)BLOCK_K, (cuuint32_t)(HTW + 2), (cuuint32_t)(HTH + 2), 1};
    if (!make_map(&mx, x, 4, dims, strides, box)) return LNST_EARG;
  }
  {
    const int rows = OUT3 ? BLOCK_N : Cout;
    const cuuint64_t dims[3] = {(cuuint64_t)xC, (cuuint64_t)rows, 9};
    const cuuint64_t strides[2] = {(cuuint64_t)xC * 2, (cuuint64_t)rows * xC * 2};
    const cuuint32_t box[3] = {(cuuint32_t)BLOCK_K, (cuuint32_t)BLOCK_N, 1};
    if (!make_map(&mw, wmat, 3, dims, strides, box)) return LNST_EARG;
  }
  static bool configured = false;
  static int sms = 148;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_halo_k<BLOCK_N, OUT3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         MAX_SMEM - 2048);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(conv3x3_halo_k<BLOCK_N, OUT3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             MAX_SMEM - 2048);
    if (e != cudaSuccess) return (int)e;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    configured = true;
  }
  const int grid = cfg.n_tiles < sms ? cfg.n_tiles : sms;
  if (gram_fused) *gram_fused = 0;
  if constexpr (!OUT3) if (ups_mask != nullptr) {
    if (!split || cfg.sb == 0 || ypool != nullptr || gram_f != nullptr) return LNST_EARG;   // streamed-weight layers only
    static bool uconfigured = false;
    if (!uconfigured) {
      cudaError_t e = cudaFuncSetAttribute(conv3x3_halo_k<BLOCK_N, false, false, false, false, true>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_SMEM - 2048);
      if (e != cudaSuccess) return (int)e;
      uconfigured = true;
    }
    conv3x3_halo_k<BLOCK_N, false, false, false, false, true><<<grid, NUM_THREADS, smem, stream>>>(
        mx, mw, bias, mask, y, y3, s, cfg, nullptr, mx, mw, ups_mask);
    return (int)cudaGetLastError();
  }
  if constexpr (!OUT3 && BLOCK_N == 128) if (gram_f != nullptr && gram_g != nullptr && split && cfg.sb != 0 && ypool == nullptr) {
    // Gram-loss gradient of the layer this data gradient lands on, accumulated by the same tiles (features F and the
    // pre-scaled per-image matrix Gd: Cout logical channels each)
    CUtensorMap mf, mg;
    const int fC = 2 * Cout;
    {
      const cuuint64_t dims[4] = {(cuuint64_t)fC, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n};
      const cuuint64_t strides[3] = {(cuuint64_t)fC * 2, (cuuint64_t)W * fC * 2, (cuuint64_t)H * W * fC * 2};
      const cuuint32_t box[4] = {(cuuint32_t)BLOCK_K, (cuuint32_t)(HTW + 2), (cuuint32_t)(HTH + 2), 1};
      if (!make_map(&mf, gram_f, 4, dims, strides, box)) return LNST_EARG;
    }
    {
      const cuuint64_t dims[3] = {(cuuint64_t)fC, (cuuint64_t)Cout, (cuuint64_t)n};
      const cuuint64_t strides[2] = {(cuuint64_t)fC * 2, (cuuint64_t)Cout * fC * 2};
      const cuuint32_t box[3] = {(cuuint32_t)BLOCK_K, (cuuint32_t)BLOCK_N, 1};
      if (!make_map(&mg, gram_g, 3, dims, strides, box)) return LNST_EARG;
    }
    static bool gconfigured = false;
    if (!gconfigured) {
      cudaError_t e = cudaFuncSetAttribute(conv3x3_halo_k<BLOCK_N, false, false, false, true>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_SMEM - 2048);
      if (e != cudaSuccess) return (int)e;
      gconfigured = true;
    }
    s.gram_kc = Cout / 64;
    conv3x3_halo_k<BLOCK_N, false, false, false, true><<<grid, NUM_THREADS, smem, stream>>>(mx, mw, bias, mask, y, y3, s, cfg,
                                                                                          nullptr, mf, mg, nullptr);
    if (gram_fused) *gram_fused = 1;
    return (int)cudaGetLastError();
  }
  if (ypool != nullptr) {
    // the pooled output is a separate instantiation: with the pool code behind a run-time test every halo kernel
    // grew by 14-35 registers and ran 12-19 % slower (ncu launch lists, profiles/)
    if (OUT3 || !split) return LNST_EARG;
  }
  if constexpr (!OUT3) if (ypool != nullptr) {
    static bool pconfigured = false;
    if (!pconfigured) {
      cudaError_t e = cudaFuncSetAttribute(conv3x3_halo_k<BLOCK_N, false, true, true>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_SMEM - 2048);
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(conv3x3_halo_k<BLOCK_N, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 MAX_SMEM - 2048);
      if (e != cudaSuccess) return (int)e;
      pconfigured = true;
    }
    if (cfg.sb == 0)
      conv3x3_halo_k<BLOCK_N, false, true, true><<<grid, NUM_THREADS, smem, stream>>>(mx, mw, bias, mask, y, y3, s, cfg, ypool, mx, mw, nullptr);
    else
      conv3x3_halo_k<BLOCK_N, false, false, true><<<grid, NUM_THREADS, smem, stream>>>(mx, mw, bias, mask, y, y3, s, cfg, ypool, mx, mw, nullptr);
    return (int)cudaGetLastError();
  }
  if (cfg.sb == 0)
    conv3x3_halo_k<BLOCK_N, OUT3, true><<<grid, NUM_THREADS, smem, stream>>>(mx, mw, bias, mask, y, y3, s, cfg, ypool, mx, mw, nullptr);
  else
    conv3x3_halo_k<BLOCK_N, OUT3, false><<<grid, NUM_THREADS, smem, stream>>>(mx, mw, bias, mask, y, y3, s, cfg, ypool, mx, mw, nullptr);
  return (int)cudaGetLastError();
}

// ---- elementwise helpers on bf16 ---------------------------------------------------------------
__global__ void f32_to_bf16_k(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int64_t n) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(x + i);
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&a);
    o.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(y + i) = o;
  } else {
    for (int64_t j = i; j < n; ++j) y[j] = __float2bfloat16_rn(x[j]);
  }
}
__global__ void bf16_to_f32_k(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, int64_t n) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const uint2 v = *reinterpret_cast<const uint2*>(x + i);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&v.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&v.y);
    const float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    *reinterpret_cast<float4*>(y + i) = make_float4(fa.x, fa.y, fb.x, fb.y);
  } else {
    for (int64_t j = i; j < n; ++j) y[j] = __bfloat162float(x[j]);
  }
}
// fp32 [rows, C] <-> split bf16 [rows, 2C] = [hi | lo] (bf16x3 operand layout, see ConvShape)
__global__ void f32_to_split_k(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int64_t rows, int C) {
  const int C2 = C / 2;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows * C2) return;
  const int64_t r = t / C2;
  const int c = (int)(t - r * C2) * 2;
  const float2 v = *reinterpret_cast<const float2*>(x + r * C + c);
  __nv_bfloat162 hi, lo;
  split2(v.x, v.y, hi, lo);
  *reinterpret_cast<__nv_bfloat162*>(y + r * 2 * C + c) = hi;
  *reinterpret_cast<__nv_bfloat162*>(y + r * 2 * C + C + c) = lo;
}
__global__ void split_to_f32_k(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, int64_t rows, int C) {
  const int C2 = C / 2;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows * C2) return;
  const int64_t r = t / C2;
  const int c = (int)(t - r * C2) * 2;
  const float2 h = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(x + r * 2 * C + c));
  const float2 l = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(x + r * 2 * C + C + c));
  *reinterpret_cast<float2*>(y + r * C + c) = make_float2(h.x + l.x, h.y + l.y);
}
__device__ __forceinline__ void load8_split(const __nv_bfloat16* __restrict__ p, int C, float* f) {
  const uint4 h = *reinterpret_cast<const uint4*>(p), l = *reinterpret_cast<const uint4*>(p + C);
  const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&h);
  const __nv_bfloat162* ll = reinterpret_cast<const __nv_bfloat162*>(&l);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 a = __bfloat1622float2(hh[e]), b = __bfloat1622float2(ll[e]);
    f[2 * e] = a.x + b.x; f[2 * e + 1] = a.y + b.y;
  }
}
__device__ __forceinline__ void store8_split(__nv_bfloat16* __restrict__ p, int C, const float* f) {
  uint4 h, l;
  __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&h);
  __nv_bfloat162* ll = reinterpret_cast<__nv_bfloat162*>(&l);
#pragma unroll
  for (int e = 0; e < 4; ++e) split2(f[2 * e], f[2 * e + 1], hh[e], ll[e]);
  *reinterpret_cast<uint4*>(p) = h;
  *reinterpret_cast<uint4*>(p + C) = l;
}
// 2x2/2 average pool on split rows: the four (hi + lo) values are summed in fp32 and split again
__global__ void avgpool2_split_fwd_k(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int n, int H,
                                     int W, int C) {
  const int OH = H / 2, OW = W / 2, C8 = C / 8;
  const unsigned total = (unsigned)n * OH * OW * C8;
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const unsigned t1 = t / (unsigned)C8, t2 = t1 / (unsigned)OW, t3 = t2 / (unsigned)OH;
  const int c = (int)(t - t1 * C8) * 8;
  const int ox = (int)(t1 - t2 * OW), oy = (int)(t2 - t3 * OH), img = (int)t3;
  const int64_t R = 2 * (int64_t)C;                                 // row stride
  const __nv_bfloat16* b = x + (((int64_t)img * H + 2 * oy) * W + 2 * ox) * R + c;
  float f0[8], f1[8], f2[8], f3[8], o[8];
  load8_split(b, C, f0); load8_split(b + R, C, f1); load8_split(b + (int64_t)W * R, C, f2); load8_split(b + (int64_t)W * R + R, C, f3);
#pragma unroll
  for (int e = 0; e < 8; ++e) o[e] = ((f0[e] + f1[e]) + (f2[e] + f3[e])) * 0.25f;   // the order of the fused epilogue (conv3x3_halo_k)
  store8_split(y + (((int64_t)img * OH + oy) * OW + ox) * R + c, C, o);
}
__global__ void avgpool2_split_bwd_k(const __nv_bfloat16* __restrict__ gy, const __nv_bfloat16* __restrict__ mask,
                                     __nv_bfloat16* __restrict__ gx, int n, int H, int W, int C) {
  const int OH = H / 2, OW = W / 2, C8 = C / 8;
  const int QH = (H + 1) / 2, QW = (W + 1) / 2;
  const unsigned total = (unsigned)n * QH * QW * C8;
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const unsigned t1 = t / (unsigned)C8, t2 = t1 / (unsigned)QW, t3 = t2 / (unsigned)QH;
  const int c = (int)(t - t1 * C8) * 8;
  const int ox = (int)(t1 - t2 * QW), oy = (int)(t2 - t3 * QH), img = (int)t3;
  const int64_t R = 2 * (int64_t)C;
  float f[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) f[e] = 0.f;
  if (oy < OH && ox < OW) {
    load8_split(gy + (((int64_t)img * OH + oy) * OW + ox) * R + c, C, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] *= 0.25f;
  }
  const int y0 = 2 * oy, x0 = 2 * ox;
  const int64_t o00 = (((int64_t)img * H + y0) * W + x0) * R + c;
  const bool hx = x0 + 1 < W, hy = y0 + 1 < H;
  const int64_t off[4] = {o00, o00 + R, o00 + (int64_t)W * R, o00 + (int64_t)W * R + R};
  const bool ok[4] = {true, hx, hy, hx && hy};
  uint4 mv[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    mv[q] = make_uint4(0, 0, 0, 0);
    if (mask && ok[q]) mv[q] = *reinterpret_cast<const uint4*>(mask + off[q]);     // hi half carries the sign
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (!ok[q]) continue;
    const __nv_bfloat16* mh = reinterpret_cast<const __nv_bfloat16*>(&mv[q]);
    float o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = (mask && !(__bfloat162float(mh[e]) > 0.f)) ? 0.f : f[e];
    store8_split(gx + off[q], C, o);
  }
}
// 2x2/2 average pool on NHWC bf16, 8 channels (16 bytes) per thread
__global__ void avgpool2_bf16_fwd_k(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int n, int H,
                                    int W, int C) {
  const int OH = H / 2, OW = W / 2, C8 = C / 8;
  const unsigned total = (unsigned)n * OH * OW * C8;               // 32-bit index arithmetic (host checks the range)
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const unsigned t1 = t / (unsigned)C8, t2 = t1 / (unsigned)OW, t3 = t2 / (unsigned)OH;
  const int c = (int)(t - t1 * C8) * 8;
  const int ox = (int)(t1 - t2 * OW);
  const int oy = (int)(t2 - t3 * OH);
  const int img = (int)t3;
  const __nv_bfloat16* b = x + (((int64_t)img * H + 2 * oy) * W + 2 * ox) * C + c;
  const uint4 v00 = *reinterpret_cast<const uint4*>(b), v01 = *reinterpret_cast<const uint4*>(b + C);
  const uint4 v10 = *reinterpret_cast<const uint4*>(b + (int64_t)W * C);
  const uint4 v11 = *reinterpret_cast<const uint4*>(b + (int64_t)W * C + C);
  const __nv_bfloat162* a0 = reinterpret_cast<const __nv_bfloat162*>(&v00);
  const __nv_bfloat162* a1 = reinterpret_cast<const __nv_bfloat162*>(&v01);
  const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&v10);
  const __nv_bfloat162* a3 = reinterpret_cast<const __nv_bfloat162*>(&v11);
  uint4 o;
  __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 f0 = __bfloat1622float2(a0[e]), f1 = __bfloat1622float2(a1[e]);
    const float2 f2 = __bfloat1622float2(a2[e]), f3 = __bfloat1622float2(a3[e]);
    oh[e] = __floats2bfloat162_rn((f0.x + f1.x + f2.x + f3.x) * 0.25f, (f0.y + f1.y + f2.y + f3.y) * 0.25f);
  }
  *reinterpret_cast<uint4*>(y + (((int64_t)img * OH + oy) * OW + ox) * C + c) = o;
}
// One thread per 2x2 input quad and 8 channels: the pooled gradient is read once, the four mask vectors are
// independent loads in flight together, four 16-byte stores.  Quads hanging over an odd edge (VALID pooling
// never read those pixels) get zero gradient.
__global__ void avgpool2_bf16_bwd_k(const __nv_bfloat16* __restrict__ gy, const __nv_bfloat16* __restrict__ mask,
                                    __nv_bfloat16* __restrict__ gx, int n, int H, int W, int C) {
  const int OH = H / 2, OW = W / 2, C8 = C / 8;
  const int QH = (H + 1) / 2, QW = (W + 1) / 2;
  const unsigned total = (unsigned)n * QH * QW * C8;               // 32-bit index arithmetic (host checks the range)
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const unsigned t1 = t / (unsigned)C8, t2 = t1 / (unsigned)QW, t3 = t2 / (unsigned)QH;
  const int c = (int)(t - t1 * C8) * 8;
  const int ox = (int)(t1 - t2 * QW);
  const int oy = (int)(t2 - t3 * QH);
  const int img = (int)t3;
  uint4 g = make_uint4(0, 0, 0, 0);
  if (oy < OH && ox < OW) g = *reinterpret_cast<const uint4*>(gy + (((int64_t)img * OH + oy) * OW + ox) * C + c);
  const int y0 = 2 * oy, x0 = 2 * ox;
  const int64_t o00 = (((int64_t)img * H + y0) * W + x0) * C + c;
  const bool hx = x0 + 1 < W, hy = y0 + 1 < H;
  const int64_t off[4] = {o00, o00 + C, o00 + (int64_t)W * C, o00 + (int64_t)W * C + C};
  const bool ok[4] = {true, hx, hy, hx && hy};
  uint4 mv[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    mv[q] = make_uint4(0, 0, 0, 0);
    if (mask && ok[q]) mv[q] = *reinterpret_cast<const uint4*>(mask + off[q]);
  }
  const __nv_bfloat16* gh = reinterpret_cast<const __nv_bfloat16*>(&g);
  float f[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) f[e] = 0.25f * __bfloat162float(gh[e]);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (!ok[q]) continue;
    const __nv_bfloat16* mh = reinterpret_cast<const __nv_bfloat16*>(&mv[q]);
    uint4 ov;
    __nv_bfloat16* oh = reinterpret_cast<__nv_bfloat16*>(&ov);
#pragma unroll
    for (int e = 0; e < 8; ++e) oh[e] = __float2bfloat16_rn((mask && !(__bfloat162float(mh[e]) > 0.f)) ? 0.f : f[e]);
    *reinterpret_cast<uint4*>(gx + off[q]) = ov;
  }
}

// ---- conv1_1 (3 -> 64) and its data gradient (64 -> 3): K = 27, far too thin for an MMA tile ----
// forward: a thread computes FP x 8 outputs (FP consecutive pixels of a row, 8 channels); every weight
// float4 fetched from shared memory feeds FP FMAs per lane, the (FP+2) x 3 x 3 input window is read
// once per tap row (the 8 lanes of a pixel group read the same addresses: one broadcast transaction).
constexpr int FP = 4;
// relu'd accumulators of one pixel's 8 channels -> bf16 (or hi/lo halves of a 128-channel split row)
__device__ __forceinline__ void store_first8(__nv_bfloat16* __restrict__ y, int64_t pixel, int cg, const float* acc, int split) {
  uint4 o, ol;
  __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
  __nv_bfloat162* ohl = reinterpret_cast<__nv_bfloat162*>(&ol);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float f0 = fmaxf(acc[2 * j], 0.f), f1 = fmaxf(acc[2 * j + 1], 0.f);
    if (split) split2(f0, f1, oh[j], ohl[j]);
    else oh[j] = __floats2bfloat162_rn(f0, f1);
  }
  if (split) {
    *reinterpret_cast<uint4*>(y + pixel * 128 + cg) = o;
    *reinterpret_cast<uint4*>(y + pixel * 128 + 64 + cg) = ol;
  } else {
    *reinterpret_cast<uint4*>(y + pixel * 64 + cg) = o;
  }
}

__global__ void __launch_bounds__(256) conv_first_fwd_k(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ b, __nv_bfloat16* __restrict__ y,
                                                        int n, int H, int W, int split) {
  __shared__ __align__(16) float ws[27 * 64];
  __shared__ float bs[64];
  for (int i = threadIdx.x; i < 27 * 64; i += blockDim.x) ws[i] = w[i];
  if (threadIdx.x < 64) bs[threadIdx.x] = b ? b[threadIdx.x] : 0.f;
  __syncthreads();
  const int WG = (W + FP - 1) / FP;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t grp = t >> 3;
  const int cg = (int)(t & 7) * 8;
  if (grp >= (int64_t)n * H * WG) return;
  const int px = (int)(grp % WG) * FP, py = (int)((grp / WG) % H);
  const int64_t img = grp / ((int64_t)WG * H);
  float acc[FP][8];
#pragma unroll
  for (int p = 0; p < FP; ++p)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[p][j] = bs[cg + j];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int yy = py + ky - 1;
    if (yy < 0 || yy >= H) continue;
    const float* row = x + (img * H + yy) * (int64_t)W * 3;
    float in[(FP + 2) * 3];
#pragma unroll
    for (int c = 0; c < FP + 2; ++c) {
      const int xx = px + c - 1;
      const bool ok = xx >= 0 && xx < W;
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) in[c * 3 + ci] = ok ? row[xx * 3 + ci] : 0.f;
    }
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) {
        const float4* wr = reinterpret_cast<const float4*>(ws + ((ky * 3 + kx) * 3 + ci) * 64 + cg);
        const float4 w0 = wr[0], w1 = wr[1];
#pragma unroll
        for (int p = 0; p < FP; ++p) {
          const float xv = in[(p + kx) * 3 + ci];
          acc[p][0] = fmaf(xv, w0.x, acc[p][0]); acc[p][1] = fmaf(xv, w0.y, acc[p][1]);
          acc[p][2] = fmaf(xv, w0.z, acc[p][2]); acc[p][3] = fmaf(xv, w0.w, acc[p][3]);
          acc[p][4] = fmaf(xv, w1.x, acc[p][4]); acc[p][5] = fmaf(xv, w1.y, acc[p][5]);
          acc[p][6] = fmaf(xv, w1.z, acc[p][6]); acc[p][7] = fmaf(xv, w1.w, acc[p][7]);
        }
      }
    }
  }
#pragma unroll
  for (int p = 0; p < FP; ++p) {
    if (px + p >= W) break;
    store_first8(y, (img * H + py) * (int64_t)W + px + p, cg, acc[p], split);
  }
}

// conv1_1 on a GRAY render (every 3-D driver: the image is one channel replicated to RGB, styler_base.py:41-43):
// x_c = s*g - mean_c, so sum_c w[tap,c,co]*x_c = g*(s*sum_c w) - sum_c w*mean_c.  ws[tap][co] = s*sum_c w[tap,c,co],
// wm[tap][co] = sum_c w[tap,c,co]*mean_c, bsum[co] = b[co] - sum_tap wm[tap][co]: 9 FMAs per output instead of 27
// and no [n,H,W,3] network input in memory.  SAME padding pads x (not g) with zeros: a tap outside the image
// contributes neither term, so border pixels add wm[tap] back for their missing taps.
__global__ void __launch_bounds__(256) conv_first_fwd_gray_k(const float* __restrict__ gimg, const float* __restrict__ ws,
                                                             const float* __restrict__ wm, const float* __restrict__ bsum,
                                                             __nv_bfloat16* __restrict__ y, int n, int H, int W, int split) {
  __shared__ __align__(16) float s_ws[9 * 64];
  __shared__ __align__(16) float s_wm[9 * 64];
  __shared__ float s_b[64];
  for (int i = threadIdx.x; i < 9 * 64; i += blockDim.x) { s_ws[i] = ws[i]; s_wm[i] = wm[i]; }
  if (threadIdx.x < 64) s_b[threadIdx.x] = bsum[threadIdx.x];
  __syncthreads();
  const int WG = (W + FP - 1) / FP;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t grp = t >> 3;
  const int cg = (int)(t & 7) * 8;
  if (grp >= (int64_t)n * H * WG) return;
  const int px = (int)(grp % WG) * FP, py = (int)((grp / WG) % H);
  const int64_t img = grp / ((int64_t)WG * H);
  float acc[FP][8];
#pragma unroll
  for (int p = 0; p < FP; ++p)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[p][j] = s_b[cg + j];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int yy = py + ky - 1;
    const bool rowok = yy >= 0 && yy < H;
    const float* row = gimg + (img * H + (rowok ? yy : 0)) * (int64_t)W;
    float in[FP + 2];
#pragma unroll
    for (int c = 0; c < FP + 2; ++c) {
      const int xx = px + c - 1;
      in[c] = (rowok && xx >= 0 && xx < W) ? row[xx] : 0.f;
    }
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const float4* wr = reinterpret_cast<const float4*>(s_ws + (ky * 3 + kx) * 64 + cg);
      const float4 w0 = wr[0], w1 = wr[1];
#pragma unroll
      for (int p = 0; p < FP; ++p) {
        const float xv = in[p + kx];
        acc[p][0] = fmaf(xv, w0.x, acc[p][0]); acc[p][1] = fmaf(xv, w0.y, acc[p][1]);
        acc[p][2] = fmaf(xv, w0.z, acc[p][2]); acc[p][3] = fmaf(xv, w0.w, acc[p][3]);
        acc[p][4] = fmaf(xv, w1.x, acc[p][4]); acc[p][5] = fmaf(xv, w1.y, acc[p][5]);
        acc[p][6] = fmaf(xv, w1.z, acc[p][6]); acc[p][7] = fmaf(xv, w1.w, acc[p][7]);
      }
    }
  }
  if (py == 0 || py == H - 1 || px == 0 || px + FP >= W) {          // border: give back the mean terms of missing taps
#pragma unroll
    for (int p = 0; p < FP; ++p) {
      const int xc = px + p;
      for (int ky = 0; ky < 3; ++ky)
        for (int kx = 0; kx < 3; ++kx) {
          const int yy = py + ky - 1, xx = xc + kx - 1;
          if (yy < 0 || yy >= H || xx < 0 || xx >= W) {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[p][j] += s_wm[(ky * 3 + kx) * 64 + cg + j];
          }
        }
    }
  }
#pragma unroll
  for (int p = 0; p < FP; ++p) {
    if (px + p >= W) break;
    store_first8(y, (img * H + py) * (int64_t)W + px + p, cg, acc[p], split);
  }
}

// data gradient: a thread computes BP consecutive pixels x 3 outputs, walking the 64 gradient channels
// in chunks of 8 (one 16-byte load per window column); each chunk's 3 x 8 x 3 weights come from
// shared memory as broadcast float4s and feed BP x 72 FMAs.  wd fp32 [3,3,64,3].
constexpr int BP = 2;
__global__ void __launch_bounds__(128) conv_first_bwd_k(const __nv_bfloat16* __restrict__ g,
                                                        const float* __restrict__ wd, float* __restrict__ gx,
                                                        int n, int H, int W) {
  __shared__ __align__(16) float ws[9 * 64 * 3];
  for (int i = threadIdx.x; i < 9 * 64 * 3; i += blockDim.x) ws[i] = wd[i];
  __syncthreads();
  const int WG = (W + BP - 1) / BP;
  const int64_t grp = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (grp >= (int64_t)n * H * WG) return;
  const int px = (int)(grp % WG) * BP, py = (int)((grp / WG) % H);
  const int64_t img = grp / ((int64_t)WG * H);
  float acc[BP][3];
#pragma unroll
  for (int p = 0; p < BP; ++p) acc[p][0] = acc[p][1] = acc[p][2] = 0.f;
  for (int ky = 0; ky < 3; ++ky) {
    const int yy = py + ky - 1;
    if (yy < 0 || yy >= H) continue;
    const __nv_bfloat16* row = g + (img * H + yy) * (int64_t)W * 64;
#pragma unroll 2
    for (int c8 = 0; c8 < 8; ++c8) {
      float gv[BP + 2][8];
#pragma unroll
      for (int c = 0; c < BP + 2; ++c) {
        const int xx = px + c - 1;
        uint4 q = make_uint4(0, 0, 0, 0);
        if (xx >= 0 && xx < W) q = *reinterpret_cast<const uint4*>(row + (int64_t)xx * 64 + c8 * 8);
        const __nv_bfloat162* hq = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __bfloat1622float2(hq[e]);
          gv[c][2 * e] = f.x; gv[c][2 * e + 1] = f.y;
        }
      }
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        // 8 channels x 3 outputs = 24 consecutive floats = 6 float4
        const float4* wt = reinterpret_cast<const float4*>(ws + ((ky * 3 + kx) * 64 + c8 * 8) * 3);
        float wv[24];
#pragma unroll
        for (int q4 = 0; q4 < 6; ++q4) {
          const float4 t4 = wt[q4];
          wv[4 * q4] = t4.x; wv[4 * q4 + 1] = t4.y; wv[4 * q4 + 2] = t4.z; wv[4 * q4 + 3] = t4.w;
        }
#pragma unroll
        for (int p = 0; p < BP; ++p)
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float f = gv[p + kx][e];
            acc[p][0] = fmaf(f, wv[e * 3], acc[p][0]);
            acc[p][1] = fmaf(f, wv[e * 3 + 1], acc[p][1]);
            acc[p][2] = fmaf(f, wv[e * 3 + 2], acc[p][2]);
          }
      }
    }
  }
#pragma unroll
  for (int p = 0; p < BP; ++p) {
    if (px + p >= W) break;
    float* o = gx + ((img * H + py) * (int64_t)W + px + p) * 3;
    o[0] = acc[p][0]; o[1] = acc[p][1]; o[2] = acc[p][2];
  }
}

// ---- data gradient of conv1_1 for a gray render as ONE GEMM per patch + a 9-term gather -------------------
// g_gray[p] = sum_tap sum_c g[p + tap, c] * w[tap, c].  The halo kernel above spends one M=128, N=16 MMA chain per
// TAP on it (9 x 3 passes x 4 k-steps per tile, every one of them bound by the 4 KiB A operand it pulls out of shared
// memory: 0.064 ms at C3 for 20 GFLOP).  Here the taps are the N dimension instead: Z[q, tap] = sum_c g[q, c] * w[tap, c]
// for every pixel q of the halo'd 18 x 10 patch (two M=128 blocks over the patch rows as they lie in shared memory,
// N = 16 of which 9 are used), i.e. each activation byte goes through the tensor pipe once per pass instead of nine
// times, and the output pixel sums its nine Z[q + tap, tap] from a shared-memory copy of Z.
// One tile per CTA, three CTAs per SM: loads, MMAs and the gather of neighbouring CTAs overlap.
constexpr int CF_BTILE = 16 * 128;                          // one 64-channel chunk of the weights: 16 rows x 128 B
template <bool SPLIT>
__global__ void __launch_bounds__(128) conv_first_bwd_gray_col_k(const __grid_constant__ CUtensorMap map_g,
                                                                  const __nv_bfloat16* __restrict__ wd16,
                                                                  float* __restrict__ g_gray, int H, int W, int tiles_w,
                                                                  int tiles_h, float scale, const float* __restrict__ dimg,
                                                                  float* __restrict__ dots) {
  constexpr int NCH = SPLIT ? 2 : 1;                         // activation / weight chunks: [hi | lo]
  __shared__ float s_part[4];
  constexpr int XC = 64 * NCH;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a = base;                              // NCH patches at PATCH_STRIDE; the second M block of the
  const uint32_t a_end = smem_a + (NCH - 1) * PATCH_STRIDE + 256 * 128;   // last one reads (unused) rows up to here
  const uint32_t smem_b = a_end;                             // NCH x 2 KiB
  const uint32_t zbuf = smem_b + NCH * CF_BTILE;             // 180 x 9 floats
  const uint32_t bars = zbuf + 180 * 9 * 4 + 8;
  const uint32_t tmem_slot = bars + 16;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));     // generic pointer to `base`
  float* Z = reinterpret_cast<float*>(gen + (zbuf - base));
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(gen + (tmem_slot - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_sp = tiles_w * tiles_h;
  const int img = blockIdx.x / tiles_sp, rem = blockIdx.x - img * tiles_sp;
  const int th = rem / tiles_w, tw = rem - th * tiles_w;

  // weights -> the K-major SWIZZLE_128B operand layout, by hand: row = tap (9 of 16, the rest zero), 16-byte piece j of
  // row r sits at piece j ^ (r & 7) of its 128-byte line
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    const int r = threadIdx.x >> 3, j = threadIdx.x & 7;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < 9) v = *reinterpret_cast<const uint4*>(wd16 + ((size_t)r * 16) * XC + ch * 64 + j * 8);   // row 0 of tap r
    st_shared_v4(smem_b + ch * CF_BTILE + r * 128 + ((j ^ (r & 7)) << 4), v);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_g);
    mbar_init(bars, 1);
    mbar_init(bars + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(32));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot_ptr;

  if (warp == 0) {
    const bool leader = elect_one();
    if (leader) {
      mbar_expect_tx(bars, NCH * PATCH_BYTES);
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch)
        tma_load_4d(smem_a + ch * PATCH_STRIDE, &map_g, bars, ch * 64, tw * HTW - 1, th * HTH - 1, img);
    }
    mbar_wait(bars, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (leader) {
      const uint32_t idesc = umma_idesc_bf16(BLOCK_M, 16);
      const uint32_t hi = desc_hi_sw128(1024);
      const uint32_t alo_base = desc_lo(smem_a, 16), blo_base = desc_lo(smem_b, 16);
      constexpr int NP = SPLIT ? 3 : 1;                      // passes: hi*Whi, lo*Whi, hi*Wlo
#pragma unroll
      for (int m = 0; m < 2; ++m) {
#pragma unroll
        for (int ps = 0; ps < NP; ++ps) {
          const int ac = (ps == 1) ? 1 : 0, bc = (ps == 2) ? 1 : 0;
          const uint32_t alo = alo_base + ((ac * PATCH_STRIDE + m * 128 * 128) >> 4);
          const uint32_t blo = blo_base + ((bc * CF_BTILE) >> 4);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
            umma_bf16_lh(tmem_d + m * 16, alo + k * 2, hi, blo + k * 2, hi, idesc, (ps | k) != 0 ? 1u : 0u);
        }
      }
      umma_commit(bars + 8);
    }
  }
  mbar_wait(bars + 8, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  {
    const int r = warp * 32 + lane;
    uint32_t v[16];
    tmem_ld16(tmem_d + ((uint32_t)(warp * 32) << 16), v);
#pragma unroll
    for (int t = 0; t < 9; ++t) Z[r * 9 + t] = __uint_as_float(v[t]);
    tmem_ld16(tmem_d + ((uint32_t)(warp * 32) << 16) + 16, v);
    if (128 + r < PATCH_ROWS) {
#pragma unroll
      for (int t = 0; t < 9; ++t) Z[(128 + r) * 9 + t] = __uint_as_float(v[t]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  {
    const int r = threadIdx.x, py = r >> 3, px = r & 7;
    const int hh = th * HTH + py, ww = tw * HTW + px;
    float acc = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) acc += Z[((py + ky) * (HTW + 2) + px + kx) * 9 + ky * 3 + kx];
    const bool ok = hh < H && ww < W;
    const int64_t pix = ((int64_t)img * H + hh) * W + ww;
    if (ok) g_gray[pix] = acc * scale;
    if (dots != nullptr) {                                   // dots[img] += sum g_gray * image (lnst_normalize_bwd's first pass)
      const float pr = lnst_warp_sum(ok ? acc * scale * dimg[pix] : 0.f);
      if (lane == 0) s_part[warp] = pr;
      __syncthreads();
      if (threadIdx.x == 0) atomicAdd(dots + img, (s_part[0] + s_part[1]) + (s_part[2] + s_part[3]));
    }
  }
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(32));
  }
}

// ---- conv1_1 of a gray render (bf16x3 output) as ONE K = 64 GEMM per 128-pixel tile ----------------------------------
// y[p, co] = relu( sum_tap g[p + tap] * ws[tap, co] + bsum[co] (+ the mean terms of taps outside the image) ), the
// arithmetic of conv_first_fwd_gray_k.  The CUDA-core kernel spends ~1200 instructions per pixel (576 FMAs + the split
// epilogue) and runs at 35 us against a 14 us write floor.  Here the 9 taps are the K dimension: every gray value and every
// weight is split into THREE bf16 pieces (24 mantissa bits) and the six products of total order <= 2 are laid side by side,
//   A row (pixel):   [ g1 | g1 | g2 | g1 | g2 | g3 ] x 9 taps = 54 of 64 columns
//   B row (channel): [ w1 | w2 | w1 | w3 | w2 | w1 ]
// so four K = 16 MMAs give the fp32-accurate sum (fp32 accumulation in TMEM) and the threads only build the operand rows
// and run the epilogue (bias, border terms, ReLU, hi/lo split, coalesced stores).  One 16 x 8-pixel tile per CTA.
__device__ __forceinline__ void split3(float v, __nv_bfloat16& a, __nv_bfloat16& b, __nv_bfloat16& c) {
  a = __float2bfloat16_rn(v);
  const float r1 = v - __bfloat162float(a);
  b = __float2bfloat16_rn(r1);
  c = __float2bfloat16_rn(r1 - __bfloat162float(b));
}
__global__ void __launch_bounds__(128) conv_first_fwd_gray_mma_k(const float* __restrict__ gimg, const float* __restrict__ ws,
                                                                 const float* __restrict__ wm, const float* __restrict__ bsum,
                                                                 __nv_bfloat16* __restrict__ y, int H, int W, int tiles_w,
                                                                 int tiles_h) {
  // dynamic shared memory, aligned by hand to 1024 B (the operand swizzle is a function of the absolute address):
  // A = 128 pixels x 64 bf16 (K-major SWIZZLE_128B), B = 64 channels x 64 bf16, then 4 x 2 KiB of store staging
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t s_a = sbase, s_b = sbase + 128 * 128, s_stage = s_b + 64 * 128;
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_tmem;
  __shared__ float s_bias[64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_sp = tiles_w * tiles_h;
  const int img = blockIdx.x / tiles_sp, rem = blockIdx.x - img * tiles_sp;
  const int th = rem / tiles_w, tw = rem - th * tiles_w;
  const int r = threadIdx.x;                                // accumulator row = tile pixel (r >> 3, r & 7)
  const int ph = th * HTH + (r >> 3), pw = tw * HTW + (r & 7);
  const bool valid = ph < H && pw < W;

  // ---- operand rows: the pixel's 9 neighbours (zero outside the image: x is padded, not g -- see the border terms) ----
  {
    __nv_bfloat16 p1[9], p2[9], p3[9];
    const float* gi = gimg + (int64_t)img * H * W;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int yy = ph + t / 3 - 1, xx = pw + t % 3 - 1;
      const float g = (valid && yy >= 0 && yy < H && xx >= 0 && xx < W) ? gi[(int64_t)yy * W + xx] : 0.f;
      split3(g, p1[t], p2[t], p3[t]);
    }
    __nv_bfloat16 row[64];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      row[t] = p1[t]; row[9 + t] = p1[t]; row[18 + t] = p2[t]; row[27 + t] = p1[t]; row[36 + t] = p2[t]; row[45 + t] = p3[t];
    }
#pragma unroll
    for (int t = 54; t < 64; ++t) row[t] = __float2bfloat16_rn(0.f);
    const uint32_t a0 = s_a + r * 128;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      st_shared_v4(a0 + ((uint32_t)((j ^ r) & 7) << 4), *reinterpret_cast<const uint4*>(&row[8 * j]));
  }
  if (r < 64) {
    __nv_bfloat16 q1[9], q2[9], q3[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) split3(ws[t * 64 + r], q1[t], q2[t], q3[t]);
    __nv_bfloat16 row[64];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      row[t] = q1[t]; row[9 + t] = q2[t]; row[18 + t] = q1[t]; row[27 + t] = q3[t]; row[36 + t] = q2[t]; row[45 + t] = q1[t];
    }
#pragma unroll
    for (int t = 54; t < 64; ++t) row[t] = __float2bfloat16_rn(0.f);
    const uint32_t b0 = s_b + r * 128;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      st_shared_v4(b0 + ((uint32_t)((j ^ r) & 7) << 4), *reinterpret_cast<const uint4*>(&row[8 * j]));
    s_bias[r] = bsum[r];
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  const uint32_t bar = smem_u32(&s_bar);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = s_tmem;
  if (warp == 0) {
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(BLOCK_M, 64);
      const uint32_t hi = desc_hi_sw128(1024);
      const uint32_t alo = desc_lo(s_a, 16), blo = desc_lo(s_b, 16);
#pragma unroll
      for (int k = 0; k < BLOCK_K / UMMA_K; ++k) umma_bf16_lh(tmem_d, alo + k * 2, hi, blo + k * 2, hi, idesc, k != 0 ? 1u : 0u);
      umma_commit(bar);
    }
  }
  mbar_wait(bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t v[64];
  tmem_ld32_nowait(tmem_d + ((uint32_t)(warp * 32) << 16), v);
  tmem_ld32_nowait(tmem_d + ((uint32_t)(warp * 32) << 16) + 32, v + 32);
  tmem_ld_wait();
  float f[64];
#pragma unroll
  for (int c = 0; c < 64; ++c) f[c] = __uint_as_float(v[c]) + s_bias[c];
  if (valid && (ph == 0 || ph == H - 1 || pw == 0 || pw == W - 1)) {   // border: give back the mean terms of missing taps
    for (int t = 0; t < 9; ++t) {
      const int yy = ph + t / 3 - 1, xx = pw + t % 3 - 1;
      if (yy < 0 || yy >= H || xx < 0 || xx >= W) {
#pragma unroll
        for (int c = 0; c < 64; ++c) f[c] += wm[t * 64 + c];
      }
    }
  }
#pragma unroll
  for (int ck = 0; ck < 2; ++ck) {
    uint4 ov[4], ol[4];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&ov[jj]);
      __nv_bfloat162* ohl = reinterpret_cast<__nv_bfloat162*>(&ol[jj]);
#pragma unroll
      for (int e = 0; e < 8; e += 2)
        split2(fmaxf(f[ck * 32 + jj * 8 + e], 0.f), fmaxf(f[ck * 32 + jj * 8 + e + 1], 0.f), oh[e >> 1], ohl[e >> 1]);
    }
    auto rowptr = [&](int row, int half) -> __nv_bfloat16* {
      const int rr = warp * 32 + row;
      const int ph2 = th * HTH + (rr >> 3), pw2 = tw * HTW + (rr & 7);
      if (ph2 >= H || pw2 >= W) return nullptr;
      return y + (((int64_t)img * H + ph2) * W + pw2) * 128 + half * 64 + ck * 32;
    };
    warp_rows_store<4>(s_stage + (uint32_t)warp * 2048u, lane, ov, [&](int row) { return rowptr(row, 0); });
    warp_rows_store<4>(s_stage + (uint32_t)warp * 2048u, lane, ol, [&](int row) { return rowptr(row, 1); });
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(64));
  }
}
// tuning switch: 1 = the kernel above for lnst_conv_first_fwd_gray_x3, 0 (default) = the CUDA-core kernel.  Measured in the
// C3 step: +13 us with the MMA form -- one tile per CTA serialises operand build, MMA and epilogue behind two block-wide
// syncs and a TMEM allocation, and 112 registers leave 4 CTAs per SM; it would have to be persistent to win.
static int conv_first_mma = 0;

static int conv_first_col = 1;    // tuning switch: 1 = the one-GEMM-per-patch kernel above for the gray data gradient, 0 = halo kernel

static int launch_first_bwd_col(const void* g, const void* wd16, float* g_gray, int n, int H, int W, int split,
                                cudaStream_t stream, const float* dimg = nullptr, float* dots = nullptr) {
  const int nch = split ? 2 : 1, xC = 64 * nch;
  CUtensorMap mg;
  const cuuint64_t dims[4] = {(cuuint64_t)xC, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n};
  const cuuint64_t strides[3] = {(cuuint64_t)xC * 2, (cuuint64_t)W * xC * 2, (cuuint64_t)H * W * xC * 2};
  const cuuint32_t box[4] = {(cuuint32_t)BLOCK_K, (cuuint32_t)(HTW + 2), (cuuint32_t)(HTH + 2), 1};
  if (!make_map(&mg, g, 4, dims, strides, box)) return LNST_EARG;
  const int smem = (nch - 1) * PATCH_STRIDE + 256 * 128 + nch * CF_BTILE + 180 * 9 * 4 + 8 + 16 + 16 + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_first_bwd_gray_col_k<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 68 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv_first_bwd_gray_col_k<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 68 * 1024);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const int tiles_w = (W + HTW - 1) / HTW, tiles_h = (H + HTH - 1) / HTH;
  const unsigned grid = (unsigned)(tiles_w * tiles_h * n);
  if (split)
    conv_first_bwd_gray_col_k<true><<<grid, 128, smem, stream>>>(mg, (const __nv_bfloat16*)wd16, g_gray, H, W, tiles_w,
                                                                 tiles_h, 1.0f, dimg, dots);
  else
    conv_first_bwd_gray_col_k<false><<<grid, 128, smem, stream>>>(mg, (const __nv_bfloat16*)wd16, g_gray, H, W, tiles_w,
                                                                  tiles_h, 1.0f, dimg, dots);
  return (int)cudaGetLastError();
}

}  // namespace tc

// ---------------------------------------------------------------------------------------
// C-ABI
// ---------------------------------------------------------------------------------------
extern "C" int lnst_conv_first_fwd(const float* x, const float* w, const float* b, void* y, int32_t n, int32_t H,
                                   int32_t W, void* stream) {
  if (!x || !w || !y || n < 1 || H < 1 || W < 1) return LNST_EARG;
  const int64_t threads = (int64_t)n * H * ((W + tc::FP - 1) / tc::FP) * 8;
  tc::conv_first_fwd_k<<<lnst_blocks(threads, 256), 256, 0, lnst_stream(stream)>>>(x, w, b, (__nv_bfloat16*)y, n, H, W, 0);
  return lnst_status();
}
extern "C" int lnst_conv_first_fwd_x3(const float* x, const float* w, const float* b, void* y, int32_t n, int32_t H,
                                      int32_t W, void* stream) {
  if (!x || !w || !y || n < 1 || H < 1 || W < 1) return LNST_EARG;
  const int64_t threads = (int64_t)n * H * ((W + tc::FP - 1) / tc::FP) * 8;
  tc::conv_first_fwd_k<<<lnst_blocks(threads, 256), 256, 0, lnst_stream(stream)>>>(x, w, b, (__nv_bfloat16*)y, n, H, W, 1);
  return lnst_status();
}

extern "C" int lnst_conv_first_fwd_gray(const float* gray, const float* ws, const float* wm, const float* bsum, void* y,
                                        int32_t n, int32_t H, int32_t W, void* stream) {
  if (!gray || !ws || !wm || !bsum || !y || n < 1 || H < 1 || W < 1) return LNST_EARG;
  const int64_t threads = (int64_t)n * H * ((W + tc::FP - 1) / tc::FP) * 8;
  tc::conv_first_fwd_gray_k<<<lnst_blocks(threads, 256), 256, 0, lnst_stream(stream)>>>(gray, ws, wm, bsum,
                                                                                       (__nv_bfloat16*)y, n, H, W, 0);
  return lnst_status();
}
extern "C" int lnst_conv_first_fwd_gray_x3(const float* gray, const float* ws, const float* wm, const float* bsum,
                                           void* y, int32_t n, int32_t H, int32_t W, void* stream) {
  if (!gray || !ws || !wm || !bsum || !y || n < 1 || H < 1 || W < 1) return LNST_EARG;
  if (tc::conv_first_mma) {
    const int tiles_w = (W + tc::HTW - 1) / tc::HTW, tiles_h = (H + tc::HTH - 1) / tc::HTH;
    tc::conv_first_fwd_gray_mma_k<<<(unsigned)(tiles_w * tiles_h * n), 128, 128 * 128 + 64 * 128 + 4 * 2048 + 1024,
                                    lnst_stream(stream)>>>(
        gray, ws, wm, bsum, (__nv_bfloat16*)y, H, W, tiles_w, tiles_h);
    return lnst_status();
  }
  const int64_t threads = (int64_t)n * H * ((W + tc::FP - 1) / tc::FP) * 8;
  tc::conv_first_fwd_gray_k<<<lnst_blocks(threads, 256), 256, 0, lnst_stream(stream)>>>(gray, ws, wm, bsum,
                                                                                       (__nv_bfloat16*)y, n, H, W, 1);
  return lnst_status();
}
extern "C" int lnst_set_conv_first_mma(int32_t on) { tc::conv_first_mma = on ? 1 : 0; return LNST_OK; }

extern "C" int lnst_conv_first_bwd(const void* g, const float* wd, float* gx, int32_t n, int32_t H, int32_t W,
                                   void* stream) {
  if (!g || !wd || !gx || n < 1 || H < 1 || W < 1) return LNST_EARG;
  const int64_t threads = (int64_t)n * H * ((W + tc::BP - 1) / tc::BP);
  tc::conv_first_bwd_k<<<lnst_blocks(threads, 128), 128, 0, lnst_stream(stream)>>>((const __nv_bfloat16*)g, wd, gx, n,
                                                                                     H, W);
  return lnst_status();
}

// developer probe: x bf16 [180,64] (18 x 10 patch rows), b bf16 [16,64], out fp32 [128,16]
extern "C" int lnst_umma_probe(const void* x, const void* b, float* out, int32_t pitched, int32_t ky, int32_t kx,
                               int32_t base_mode, void* stream) {
  using namespace tc;
  CUtensorMap md, mr, mb;
  const cuuint64_t dx[2] = {64, 180}, sx[1] = {128};
  const cuuint32_t bd[2] = {64, 180}, br[2] = {64, 10}, bb[2] = {64, 16};
  const cuuint64_t db[2] = {64, 16};
  if (!make_map(&md, x, 2, dx, sx, bd) || !make_map(&mr, x, 2, dx, sx, br) || !make_map(&mb, b, 2, db, sx, bb))
    return LNST_EARG;
  const int smem = 44 * 1024 + 1024;
  cudaFuncSetAttribute(umma_probe_k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  umma_probe_k<<<1, 128, smem, lnst_stream(stream)>>>(md, mr, mb, out, pitched, ky, kx, base_mode);
  return lnst_status();
}

extern "C" int lnst_set_conv_halo(int32_t on) { tc::conv_halo = on; return LNST_OK; }   // 2 = also with streamed weights

// Data gradient of conv1_1 on tensor cores: g bf16 [n,H,W,64] (x) wd16 bf16 [9,16,64] (rows 0..2 = the three
// input channels, rows 3..15 zero) -> gx fp32 [n,H,W,3].
extern "C" int lnst_conv_first_bwd_tc(const void* g, const void* wd16, float* gx, int32_t n, int32_t H, int32_t W,
                                      void* stream) {
  if (!g || !wd16 || !gx || n < 1 || H < 1 || W < 1) return LNST_EARG;
  return tc::launch_halo<16, true>(g, wd16, nullptr, nullptr, nullptr, gx, n, H, W, 64, 16, 0, 1.0f,
                                   lnst_stream(stream));
}

// The same for a gray render: wd16 row 0 = s * sum_c of the three data-gradient rows -> g_gray fp32 [n,H,W].
extern "C" int lnst_conv_first_bwd_gray_tc(const void* g, const void* wd16, float* g_gray, int32_t n, int32_t H,
                                           int32_t W, void* stream) {
  if (!g || !wd16 || !g_gray || n < 1 || H < 1 || W < 1) return LNST_EARG;
  if (tc::conv_first_col) return tc::launch_first_bwd_col(g, wd16, g_gray, n, H, W, 0, lnst_stream(stream));
  return tc::launch_halo<16, true>(g, wd16, nullptr, nullptr, nullptr, g_gray, n, H, W, 64, 16, 0, 1.0f,
                                   lnst_stream(stream), 1);
}

// split (bf16x3) variants: g bf16 [n,H,W,128] = [hi | lo], wd16 bf16 [9,16,128] = [hi | lo] rows
extern "C" int lnst_conv_first_bwd_x3_tc(const void* g, const void* wd16, float* gx, int32_t n, int32_t H, int32_t W,
                                         void* stream) {
  if (!g || !wd16 || !gx || n < 1 || H < 1 || W < 1) return LNST_EARG;
  return tc::launch_halo<16, true>(g, wd16, nullptr, nullptr, nullptr, gx, n, H, W, 64, 16, 0, 1.0f,
                                   lnst_stream(stream), 3, 1);
}
extern "C" int lnst_conv_first_bwd_gray_x3_tc(const void* g, const void* wd16, float* g_gray, int32_t n, int32_t H,
                                              int32_t W, void* stream) {
  if (!g || !wd16 || !g_gray || n < 1 || H < 1 || W < 1) return LNST_EARG;
  if (tc::conv_first_col) return tc::launch_first_bwd_col(g, wd16, g_gray, n, H, W, 1, lnst_stream(stream));
  return tc::launch_halo<16, true>(g, wd16, nullptr, nullptr, nullptr, g_gray, n, H, W, 64, 16, 0, 1.0f,
                                   lnst_stream(stream), 1, 1);
}
extern "C" int lnst_set_conv_first_col(int32_t on) { tc::conv_first_col = on ? 1 : 0; return LNST_OK; }
// The gray data gradient of conv1_1 plus dots[i] += sum_p g_gray[i,p] * img[i,p] from the same kernel (the reduction
// lnst_normalize_bwd starts with; dots zero on entry).  split: g and wd16 carry [hi | lo] halves (bf16x3).
extern "C" int lnst_conv_first_bwd_gray_dot_tc(const void* g, const void* wd16, float* g_gray, const float* img, float* dots,
                                               int32_t split, int32_t n, int32_t H, int32_t W, void* stream) {
  if (!g || !wd16 || !g_gray || !img || !dots || n < 1 || H < 1 || W < 1) return LNST_EARG;
  return tc::launch_first_bwd_col(g, wd16, g_gray, n, H, W, split ? 1 : 0, lnst_stream(stream), img, dots);
}

extern "C" int lnst_set_conv_persistent(int32_t on) { tc::conv_persistent = on ? 1 : 0; return LNST_OK; }

extern "C" int lnst_tc_supported(void) { return tc::encode_fn() != nullptr ? 1 : 0; }

// shared host path of the 3x3 convolution (taps = 9, one weight set) and the per-pixel GEMM
// (taps = 1, one B matrix per image)
static int run_tc_gemm(const void* x, const void* wmat, const float* bias, const void* mask, const void* addend,
                       void* y, int n, int H, int W, int Cin, int Cout, int relu, int taps, int w_img, float scale,
                       void* stream, int split = 0, void* ypool = nullptr, const void* gram_f = nullptr,
                       const void* gram_g = nullptr, int* gram_fused = nullptr, const void* ups_mask = nullptr) {
  using namespace tc;
  if (!x || !wmat || !y || n < 1 || H < 1 || W < 1 || Cin < 64 || Cout < 64 || Cin % 64 || Cout % 64)
    return LNST_EARG;
  // The halo'd-patch kernel reads every input byte once per tile instead of once per tap; the weights stay
  // resident in shared memory when 9*Cin*Cout*2 B fit beside >= 3 patch stages (conv1_2, conv2_1 and their data
  // gradients) and stream through their own ring otherwise.  Since the issue loops run warp-uniform it beats the
  // per-tap kernel on every layer (tools/convbench.py, 18 images: conv2_2 45.5 vs 59.0 us, conv3_1 28.0 vs 35.5),
  // whose 32 KiB of TMA writes per k-step compete with the MMAs' own operand reads for shared-memory bandwidth.
  const int bn_ = (Cout % 128 == 0) ? 128 : 64;
  const int wch = (split ? 2 : 1) * (Cin / 64);
  const bool resident = (Cout == bn_) && (9 * wch * bn_ * 128 + 3 * PATCH_STRIDE <= 232448 - 2048 - 1024 - 512 - (bn_ >= 128 ? 8192 : 0) - 128);
  if (taps == 9 && !w_img && !addend && (split || (conv_halo && (resident || conv_halo == 2)))) {
    if (ypool && !split) return LNST_EARG;
    if (Cout % 128 == 0)
      return launch_halo<128, false>(x, wmat, bias, (const __nv_bfloat16*)mask, (__nv_bfloat16*)y, nullptr, n, H, W,
                                     Cin, Cout, relu, scale, lnst_stream(stream), 3, split, (__nv_bfloat16*)ypool,
                                     gram_f, gram_g, gram_fused, (const __nv_bfloat16*)ups_mask);
    return launch_halo<64, false>(x, wmat, bias, (const __nv_bfloat16*)mask, (__nv_bfloat16*)y, nullptr, n, H, W, Cin,
                                  Cout, relu, scale, lnst_stream(stream), 3, split, (__nv_bfloat16*)ypool, nullptr, nullptr,
                                  nullptr, (const __nv_bfloat16*)ups_mask);
  }
  if (ypool || ups_mask) return LNST_EARG;
  if (split && (!conv_persistent || taps != 1)) return LNST_EARG;    // split operands: halo kernel, or the persistent per-pixel GEMM
  ConvShape s;
  s.H = H; s.W = W; s.Cin = Cin; s.Cout = Cout; s.relu = relu;
  s.taps = taps; s.w_img = w_img; s.scale = scale; s.out_ch = 3;
  if (split) shape_split(s); else shape_plain(s);
  const int xC = split ? 2 * Cin : Cin;
  pick_tile(H, W, s.TH, s.TW);
  s.tiles_w = (W + s.TW - 1) / s.TW;
  s.tiles_h = (H + s.TH - 1) / s.TH;
  const int BN = (Cout % 128 == 0) ? 128 : 64;
  CUtensorMap mx, mw;
  {
    const cuuint64_t dims[4] = {(cuuint64_t)xC, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n};
    const cuuint64_t strides[3] = {(cuuint64_t)xC * 2, (cuuint64_t)W * xC * 2, (cuuint64_t)H * W * xC * 2};
    const cuuint32_t box[4] = {(cuuint32_t)BLOCK_K, (cuuint32_t)s.TW, (cuuint32_t)s.TH, 1};
    if (!make_map(&mx, x, 4, dims, strides, box)) return LNST_EARG;
  }
  {
    const cuuint64_t dims[3] = {(cuuint64_t)xC, (cuuint64_t)Cout, (cuuint64_t)(w_img ? n : taps)};
    const cuuint64_t strides[2] = {(cuuint64_t)xC * 2, (cuuint64_t)Cout * xC * 2};
    const cuuint32_t box[3] = {(cuuint32_t)BLOCK_K, (cuuint32_t)BN, 1};
    if (!make_map(&mw, wmat, 3, dims, strides, box)) return LNST_EARG;
  }
  if (BN == 128)
    return launch_conv<128>(mx, mw, bias, (const __nv_bfloat16*)mask, (const __nv_bfloat16*)addend,
                            (__nv_bfloat16*)y, s, n, lnst_stream(stream));
  return launch_conv<64>(mx, mw, bias, (const __nv_bfloat16*)mask, (const __nv_bfloat16*)addend, (__nv_bfloat16*)y,
                         s, n, lnst_stream(stream));
}

extern "C" int lnst_conv3x3_bf16_tc(const void* x, const void* w_packed, const float* bias, const void* mask,
                                    void* y, int32_t n, int32_t H, int32_t W, int32_t Cin, int32_t Cout,
                                    int32_t relu, void* stream) {
  return run_tc_gemm(x, w_packed, bias, mask, nullptr, y, n, H, W, Cin, Cout, relu, 9, 0, 1.0f, stream);
}

// bf16x3: x [n,H,W,2*Cin] = [hi | lo], w_packed [9, Cout, 2*Cin] = [Whi | Wlo], mask (ReLU mask of the layer below) and
// y [n,H,W,2*Cout] split rows.  Three K passes (hi*hi, lo*hi, hi*lo) into one fp32 accumulator (ConvShape).
extern "C" int lnst_conv3x3_bf16x3_tc(const void* x, const void* w_packed, const float* bias, const void* mask,
                                      void* y, int32_t n, int32_t H, int32_t W, int32_t Cin, int32_t Cout,
                                      int32_t relu, void* stream) {
  return run_tc_gemm(x, w_packed, bias, mask, nullptr, y, n, H, W, Cin, Cout, relu, 9, 0, 1.0f, stream, 1);
}
// Data gradient through a 3x3 convolution PLUS the Gram-loss gradient of the layer it lands on, masked together:
//   y = relu_mask(F) * ( x (*) w  +  F x Gd )      F split [n,H,W,2*Cout], Gd2s split [n,Cout,2*Cout] = coef * (G - Gs)
// One kernel when the layer's weights are streamed (Cout % 128 == 0); otherwise the convolution followed by the per-pixel GEMM.
extern "C" int lnst_conv3x3_gram_bf16x3_tc(const void* x, const void* w_packed, const void* F, const void* Gd2s, void* y,
                                           int32_t n, int32_t H, int32_t W, int32_t Cin, int32_t Cout, void* stream) {
  if (!F || !Gd2s) return LNST_EARG;
  int fused = 0;
  int rc = run_tc_gemm(x, w_packed, nullptr, F, nullptr, y, n, H, W, Cin, Cout, 0, 9, 0, 1.0f, stream, 1, nullptr, F, Gd2s,
                       &fused);
  if (rc != 0 || fused) return rc;
  return run_tc_gemm(F, Gd2s, nullptr, F, y, y, n, H, W, Cout, Cout, 0, 1, 1, 1.0f, stream, 1);
}

// Data gradient through a 3x3 convolution whose INPUT is the output of a 2x2 average pool: the epilogue writes the gradient
// of the pool's input, g_fine [n, 2H, 2W, 2*Cout] = 0.25 * g under the ReLU mask of the layer below the pool (fine_act, same
// shape) -- lnst_conv3x3_bf16x3_tc followed by lnst_avgpool2_bf16x3_bwd, bit-identical.  H, W: the coarse level.  Returns
// LNST_EARG when the layer's weights are resident in shared memory (no room for the staging tile): call the two separately.
extern "C" int lnst_conv3x3_unpool_bf16x3_tc(const void* x, const void* w_packed, const void* fine_act, void* g_fine,
                                             int32_t n, int32_t H, int32_t W, int32_t Cin, int32_t Cout, void* stream) {
  if (!fine_act || !g_fine) return LNST_EARG;
  return run_tc_gemm(x, w_packed, nullptr, nullptr, nullptr, g_fine, n, H, W, Cin, Cout, 0, 9, 0, 1.0f, stream, 1, nullptr,
                     nullptr, nullptr, nullptr, fine_act);
}

// The same convolution, with the 2x2 average pool of its output (lnst_avgpool2_bf16x3_fwd) written by the same epilogue:
// y_pool bf16 [n, H/2, W/2, 2*Cout].  Bit-identical to the two separate calls.
extern "C" int lnst_conv3x3_pool_bf16x3_tc(const void* x, const void* w_packed, const float* bias, const void* mask,
                                           void* y, void* y_pool, int32_t n, int32_t H, int32_t W, int32_t Cin,
                                           int32_t Cout, int32_t relu, void* stream) {
  if (!y_pool || H < 2 || W < 2) return LNST_EARG;
  return run_tc_gemm(x, w_packed, bias, mask, nullptr, y, n, H, W, Cin, Cout, relu, 9, 0, 1.0f, stream, 1, y_pool);
}
// F split [n,H,W,2C], Gd2 bf16 [n,C,2C] = [hi | lo] of the (symmetric) Gram difference, addend / g split rows
extern "C" int lnst_gram_bwd_bf16x3_tc(const void* F, const void* Gd2, float coef, const void* addend,
                                       int32_t relu_mask, void* g, int32_t n, int32_t H, int32_t W, int32_t C,
                                       void* stream) {
  return run_tc_gemm(F, Gd2, nullptr, relu_mask ? F : nullptr, addend, g, n, H, W, C, C, 0, 1, 1, coef, stream, 1);
}

// Gram-loss gradient on tensor cores: g[img] = (addend + coef * F[img] x Gd[img]) * (F > 0 if relu_mask).
// F bf16 [n,H,W,C] (post-ReLU features), Gd bf16 [n,C,C] (symmetric: rows are K-major), g bf16.
extern "C" int lnst_gram_bwd_bf16_tc(const void* F, const void* Gd, float coef, const void* addend,
                                     int32_t relu_mask, void* g, int32_t n, int32_t H, int32_t W, int32_t C,
                                     void* stream) {
  return run_tc_gemm(F, Gd, nullptr, relu_mask ? F : nullptr, addend, g, n, H, W, C, C, 0, 1, 1, coef, stream);
}

// Gram difference on tensor cores, batched over images: G[i] = F[i]^T F[i] / denom - Gs (fp32, [n,C,C]),
// Gd = bf16 copy of G (operand of lnst_gram_bwd_bf16_tc), loss[i] += weight * sum(G[i]^2).
// F bf16 [n,P,C]; Gs fp32 [C,C] or NULL (then G = F^T F / denom: the style-target pass).
// raw F^T F of a bf16 feature map into G (fp32 [n,C,C], zeroed here); sym: only the blocks with row block >= column block
static int gram_raw(const void* F, int n, int64_t P, int C, float* G, cudaStream_t st, int sym) {
  using namespace tc;
  CUtensorMap mf;
  {
    const cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)P, (cuuint64_t)n};
    const cuuint64_t strides[2] = {(cuuint64_t)C * 2, (cuuint64_t)P * C * 2};
    const cuuint32_t box[3] = {64, 64, 1};
    if (!make_map(&mf, F, 3, dims, strides, box)) return LNST_EARG;
  }
  const int tiles_n = (C + 127) / 128;
  const int tiles = sym ? tiles_n * (tiles_n + 1) / 2 : tiles_n * tiles_n;
  int splits = (148 + tiles * n - 1) / (tiles * n);            // one wave of CTAs: fewer split-K atomics per element
  const int max_splits = (int)((P + 255) / 256);              // at least 4 K-blocks per CTA
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int kps = (int)((P + splits - 1) / splits);
  kps = ((kps + 63) / 64) * 64;
  splits = (int)((P + kps - 1) / kps);
  const int smem = 2 * G_STAGES * G_OP_BYTES + 8 * (2 * G_STAGES + 1) + 16 + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gram_tc_k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  cudaMemsetAsync(G, 0, sizeof(float) * (size_t)n * C * C, st);
  gram_tc_k<<<dim3(tiles, splits, n), NUM_THREADS, smem, st>>>(mf, G, (int)P, (int)C, tiles_n, kps, sym);
  return (int)cudaGetLastError();
}

extern "C" int lnst_gram_diff_bf16_tc(const void* F, int32_t n, int64_t P, int32_t C, float denom, const float* Gs,
                                      float weight, float* G, void* Gd, float* loss, void* stream) {
  using namespace tc;
  if (!F || !G || n < 1 || P < 1 || C < 64 || C % 64 || !(denom > 0.f) || P > 0x7fffffff) return LNST_EARG;
  cudaStream_t st = lnst_stream(stream);
  const int rc = gram_raw(F, n, P, C, G, st, 0);
  if (rc != 0) return rc;
  const int n_el = C * C;
  const unsigned nb = lnst_blocks(n_el, 256) > 32 ? 32 : lnst_blocks(n_el, 256);
  gram_finish_bf16_k<<<dim3(nb, n), 256, 0, st>>>(G, Gs, (__nv_bfloat16*)Gd, n_el, 1.f / denom, weight, loss);
  return lnst_status();
}

// Split features F bf16 [n,P,2C]: the 2C x 2C Gram of the split rows (same tensor-core kernel) lands in the scratch G2
// [n,2C,2C] fp32 and its four blocks are folded into F^T F (hi*hi + hi*lo + lo*hi + lo*lo); G [n,C,C] fp32 = that / denom
// - Gs, Gd2 [n,C,2C] its split copy, loss[i] += weight * sum(G[i]^2).
static int gram_split3 = 1;       // tuning switch: 1 = gram_split_tc_k (hi/lo products summed in TMEM), 0 = 2C x 2C Gram of the split rows
extern "C" int lnst_set_gram_split3(int32_t on) { gram_split3 = on ? 1 : 0; return LNST_OK; }
extern "C" int lnst_gram_diff_bf16x3_tc(const void* F, int32_t n, int64_t P, int32_t C, float denom, const float* Gs,
                                        float weight, float* G2, float* G, void* Gd2, float* loss, void* stream) {
  return lnst_gram_diff_scaled_bf16x3_tc(F, n, P, C, denom, Gs, weight, 1.0f, G2, G, Gd2, loss, stream);
}
// The same with Gd2 = split(gd_scale * G): the gradient GEMM's operand carries the loss coefficient, so that F x Gd2 can be
// accumulated next to other products (lnst_conv3x3_gram_bf16x3_tc).  gd_scale != 1 needs C % 128 == 0.
extern "C" int lnst_gram_diff_scaled_bf16x3_tc(const void* F, int32_t n, int64_t P, int32_t C, float denom, const float* Gs,
                                               float weight, float gd_scale, float* G2, float* G, void* Gd2, float* loss,
                                               void* stream) {
  const float gram_gd_scale = gd_scale;
  if (!F || !G2 || !G || n < 1 || P < 1 || C < 64 || C % 64 || !(denom > 0.f) || P > 0x7fffffff) return LNST_EARG;
  if (gd_scale != 1.0f && !(C % 128 == 0 && gram_split3)) return LNST_EARG;
  if (C % 128 == 0 && gram_split3) {
    using namespace tc;
    cudaStream_t st = lnst_stream(stream);
    CUtensorMap mf;
    const cuuint64_t dims[3] = {(cuuint64_t)(2 * C), (cuuint64_t)P, (cuuint64_t)n};
    const cuuint64_t strides[2] = {(cuuint64_t)(2 * C) * 2, (cuuint64_t)P * (2 * C) * 2};
    const cuuint32_t box[3] = {64, 64, 1};
    if (!make_map(&mf, F, 3, dims, strides, box)) return LNST_EARG;
    const int tiles_n = C / 128, tiles = tiles_n * (tiles_n + 1) / 2;
    int splits = (148 + tiles * n - 1) / (tiles * n);           // about one wave of CTAs
    const int max_splits = (int)((P + 255) / 256);             // at least 4 k-blocks per CTA
    if (splits > max_splits) splits = max_splits;
    if (splits > 16) splits = 16;                              // every split adds 16 K atomics onto the same tile (1-2 images per rank)
    if (splits < 1) splits = 1;
    int kps = (int)((P + splits - 1) / splits);
    kps = ((kps + 63) / 64) * 64;
    splits = (int)((P + kps - 1) / kps);
    const int smem = G3_STAGES * G3_STAGE_BYTES + 8 * (2 * G3_STAGES + 1) + 16 + 1024;
    static bool configured = false;
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(gram_split_tc_k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e != cudaSuccess) return (int)e;
      configured = true;
    }
    cudaMemsetAsync(G2, 0, sizeof(float) * (size_t)n * C * C, st);
    gram_split_tc_k<<<dim3(tiles, splits, n), NUM_THREADS, smem, st>>>(mf, G2, (int)P, (int)C, tiles_n, kps);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    const int n_el3 = C * C;
    const unsigned nb3 = (unsigned)((n_el3 + 1023) / 1024);
    gram_finish_split3_k<<<dim3(nb3, n), 256, 0, st>>>(G2, G, Gs, (__nv_bfloat16*)Gd2, (int)C, 1.f / denom, weight, loss,
                                                       gram_gd_scale);
    return lnst_status();
  }
  const int rc = gram_raw(F, n, P, 2 * C, G2, lnst_stream(stream), 1);     // symmetric: 10 of the 16 tiles at C = 256
  if (rc != 0) return rc;
  const int n_el = C * C;
  const unsigned nb = lnst_blocks(n_el, 256) > 32 ? 32 : lnst_blocks(n_el, 256);
  tc::gram_finish_split_k<<<dim3(nb, n), 256, 0, lnst_stream(stream)>>>(G2, G, Gs, (__nv_bfloat16*)Gd2, (int)C,
                                                                        1.f / denom, weight, loss);
  return lnst_status();
}

extern "C" int lnst_avgpool2_bf16x3_fwd(const void* x, void* y, int32_t n, int32_t H, int32_t W, int32_t C,
                                        void* stream) {
  if (!x || !y || n < 1 || H < 2 || W < 2 || C < 8 || C % 8) return LNST_EARG;
  const int64_t total = (int64_t)n * (H / 2) * (W / 2) * (C / 8);
  if (total >= 0x7fffffff) return LNST_EARG;
  tc::avgpool2_split_fwd_k<<<lnst_blocks(total, 256), 256, 0, lnst_stream(stream)>>>(
      (const __nv_bfloat16*)x, (__nv_bfloat16*)y, n, H, W, C);
  return lnst_status();
}
extern "C" int lnst_avgpool2_bf16x3_bwd(const void* g_y, const void* mask, void* g_x, int32_t n, int32_t H, int32_t W,
                                        int32_t C, void* stream) {
  if (!g_y || !g_x || n < 1 || H < 2 || W < 2 || C < 8 || C % 8) return LNST_EARG;
  const int64_t total = (int64_t)n * ((H + 1) / 2) * ((W + 1) / 2) * (C / 8);
  if (total >= 0x7fffffff) return LNST_EARG;
  tc::avgpool2_split_bwd_k<<<lnst_blocks(total, 256), 256, 0, lnst_stream(stream)>>>(
      (const __nv_bfloat16*)g_y, (const __nv_bfloat16*)mask, (__nv_bfloat16*)g_x, n, H, W, C);
  return lnst_status();
}
extern "C" int lnst_f32_to_bf16x3(const float* x, void* y, int64_t rows, int32_t C, void* stream) {
  if (!x || !y || rows < 0 || C < 2 || C % 2) return LNST_EARG;
  if (rows == 0) return LNST_OK;
  tc::f32_to_split_k<<<lnst_blocks(rows * (C / 2), 256), 256, 0, lnst_stream(stream)>>>(x, (__nv_bfloat16*)y, rows, C);
  return lnst_status();
}
extern "C" int lnst_bf16x3_to_f32(const void* x, float* y, int64_t rows, int32_t C, void* stream) {
  if (!x || !y || rows < 0 || C < 2 || C % 2) return LNST_EARG;
  if (rows == 0) return LNST_OK;
  tc::split_to_f32_k<<<lnst_blocks(rows * (C / 2), 256), 256, 0, lnst_stream(stream)>>>((const __nv_bfloat16*)x, y, rows, C);
  return lnst_status();
}

// CUDA-core convolution with mixed I/O types for the thin edge layers (conv1_1: Cin = 3; its
// data gradient: Cout = 3).  w is fp32 HWIO like lnst_conv3x3_f32.
extern "C" int lnst_conv3x3_mixed(const void* x, int32_t x_bf16, const float* w, const float* b, const void* mask,
                                  void* y, int32_t y_bf16, int32_t n, int32_t H, int32_t W, int32_t Cin,
                                  int32_t Cout, int32_t relu, void* stream) {
  if (!x || !w || !y || n < 1 || H < 1 || W < 1 || Cin < 1 || Cout < 1) return LNST_EARG;
  if ((int64_t)n * H * W > 0x7fffffff) return LNST_EARG;
  const int M = n * H * W, K = 9 * Cin;
  RowMajorB B{w, (int)Cout};
  typedef __nv_bfloat16 bf;
  cudaStream_t s = lnst_stream(stream);
  if (!x_bf16 && y_bf16) {
    ConvAT<float> A{(const float*)x, (int)H, (int)W, (int)Cin};
    ConvEpilogueT<bf, bf> ep{(bf*)y, b, (const bf*)mask, (int)Cout, (int)relu};
    return run_sgemm(A, B, ep, M, Cout, K, 1, s);
  }
  if (x_bf16 && !y_bf16) {
    ConvAT<bf> A{(const bf*)x, (int)H, (int)W, (int)Cin};
    ConvEpilogueT<float, bf> ep{(float*)y, b, (const bf*)mask, (int)Cout, (int)relu};
    return run_sgemm(A, B, ep, M, Cout, K, 1, s);
  }
  if (x_bf16 && y_bf16) {
    ConvAT<bf> A{(const bf*)x, (int)H, (int)W, (int)Cin};
    ConvEpilogueT<bf, bf> ep{(bf*)y, b, (const bf*)mask, (int)Cout, (int)relu};
    return run_sgemm(A, B, ep, M, Cout, K, 1, s);
  }
  return LNST_EARG;
}

extern "C" int lnst_avgpool2_bf16_fwd(const void* x, void* y, int32_t n, int32_t H, int32_t W, int32_t C,
                                      void* stream) {
  if (!x || !y || n < 1 || H < 2 || W < 2 || C < 8 || C % 8) return LNST_EARG;
  const int64_t total = (int64_t)n * (H / 2) * (W / 2) * (C / 8);
  if (total >= 0x7fffffff) return LNST_EARG;
  tc::avgpool2_bf16_fwd_k<<<lnst_blocks(total, 256), 256, 0, lnst_stream(stream)>>>(
      (const __nv_bfloat16*)x, (__nv_bfloat16*)y, n, H, W, C);
  return lnst_status();
}

extern "C" int lnst_avgpool2_bf16_bwd(const void* g_y, const void* mask, void* g_x, int32_t n, int32_t H, int32_t W,
                                      int32_t C, void* stream) {
  if (!g_y || !g_x || n < 1 || H < 2 || W < 2 || C < 8 || C % 8) return LNST_EARG;
  const int64_t total = (int64_t)n * ((H + 1) / 2) * ((W + 1) / 2) * (C / 8);   // one thread per 2x2 quad, 8 channels
  if (total >= 0x7fffffff) return LNST_EARG;
  tc::avgpool2_bf16_bwd_k<<<lnst_blocks(total, 256), 256, 0, lnst_stream(stream)>>>(
      (const __nv_bfloat16*)g_y, (const __nv_bfloat16*)mask, (__nv_bfloat16*)g_x, n, H, W, C);
  return lnst_status();
}

extern "C" int lnst_f32_to_bf16(const float* x, void* y, int64_t n, void* stream) {
  if (!x || !y || n < 0) return LNST_EARG;
  if (n == 0) return LNST_OK;
  tc::f32_to_bf16_k<<<lnst_blocks((n + 3) / 4, 256), 256, 0, lnst_stream(stream)>>>(x, (__nv_bfloat16*)y, n);
  return lnst_status();
}

extern "C" int lnst_bf16_to_f32(const void* x, float* y, int64_t n, void* stream) {
  if (!x || !y || n < 0) return LNST_EARG;
  if (n == 0) return LNST_OK;
  tc::bf16_to_f32_k<<<lnst_blocks((n + 3) / 4, 256), 256, 0, lnst_stream(stream)>>>((const __nv_bfloat16*)x, y, n);
  return lnst_status();
}
