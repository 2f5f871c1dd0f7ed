// CPU interpreter for the SIMT kernels in neural-flow-style_b200/csrc -- TEST TOOLING ONLY.
//
// The build container has no GPU.  To debug indexing and arithmetic of the hand-written
// kernels before spending GPU minutes, the very same .cu sources are compiled with g++
// against this header (-DLNST_CPU_EMU): every CUDA thread of a block becomes a cooperative
// fiber (ucontext), blocks run one after another, __syncthreads()/__shfl_*_sync are fiber
// barriers, atomics are plain read-modify-writes.  The resulting liblnst_emu.so is loaded
// ONLY by tests (tests/test_emu_*.py) through an explicit handle; the product loader
// (lnst/_lib.py) knows nothing about it and fails loudly when the CUDA library is absent.
// tcgen05/TMA kernels are not emulated.
#pragma once
#include <ucontext.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>
#include <algorithm>

struct uint3_ { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct int3 { int x, y, z; };
struct int4 { int x, y, z, w; };
static inline float2 make_float2(float a, float b) { return {a, b}; }
static inline float3 make_float3(float a, float b, float c) { return {a, b, c}; }
static inline float4 make_float4(float a, float b, float c, float d) { return {a, b, c, d}; }
static inline int3 make_int3(int a, int b, int c) { return {a, b, c}; }
static inline int2 make_int2(int a, int b) { return {a, b}; }

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaPeekAtLastError() { return 0; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { memcpy(d, s, n); return 0; }
enum { cudaMemcpyDeviceToDevice = 3 };

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __constant__ static

namespace emu {
struct Fiber {
  ucontext_t ctx;
  std::vector<char> stack;
  uint3_ tid;
  bool done = false;
};
struct Block {
  std::vector<Fiber> fibers;
  ucontext_t sched;
  int cur = -1;
  uint3_ bid;
  dim3 bdim, gdim;
  // barrier state: phase counters per barrier id (0 = block, 1+warp = warps)
  std::vector<long> arrived;
  std::vector<long> phase;
  std::vector<char> dyn;
  std::vector<float> shfl_f;   // per-thread exchange slot
  std::vector<int> shfl_i;
  const std::function<void()>* body = nullptr;
};
inline Block*& B() { static Block* b = nullptr; return b; }

inline int linear_tid() {
  Block* b = B();
  const uint3_& t = b->fibers[b->cur].tid;
  return (int)(t.x + b->bdim.x * (t.y + b->bdim.y * t.z));
}
inline void yield() {
  Block* b = B();
  swapcontext(&b->fibers[b->cur].ctx, &b->sched);
}
// generic barrier among `count` fibers sharing barrier id `id`
inline void barrier(int id, int count) {
  Block* b = B();
  long my = b->phase[id];
  if (++b->arrived[id] == count) { b->arrived[id] = 0; b->phase[id]++; }
  while (b->phase[id] == my) yield();
}
inline void trampoline() {
  Block* b = B();
  (*b->body)();
  b->fibers[b->cur].done = true;
  swapcontext(&b->fibers[b->cur].ctx, &b->sched);
}
inline void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  const int nt = (int)(block.x * block.y * block.z);
  const int nwarp = (nt + 31) / 32;
  static Block blk;   // fiber stacks are reused across launches
  blk.bdim = block; blk.gdim = grid;
  if ((int)blk.fibers.size() < nt) blk.fibers.resize(nt);
  blk.dyn.assign(smem + 16, 0);
  blk.shfl_f.assign(nt, 0.f); blk.shfl_i.assign(nt, 0);
  blk.body = &body;
  for (auto& f : blk.fibers) if (f.stack.empty()) f.stack.resize(128 * 1024);
  B() = &blk;
  // LNST_EMU_ORDER=reverse runs blocks and the fibers of a block in descending order: a kernel whose result depends
  // on the (unspecified) execution order of threads between barriers, or of blocks, then fails its parity test --
  // a poor man's racecheck for missing __syncthreads / inter-block assumptions.
  static const bool rev = [] { const char* e = getenv("LNST_EMU_ORDER"); return e && e[0] == 'r'; }();
  const unsigned long nblk = (unsigned long)grid.x * grid.y * grid.z;
  for (unsigned long bl = 0; bl < nblk; ++bl) {
    const unsigned long bi = rev ? nblk - 1 - bl : bl;
    const unsigned bx = (unsigned)(bi % grid.x), by = (unsigned)((bi / grid.x) % grid.y), bz = (unsigned)(bi / ((unsigned long)grid.x * grid.y));
    blk.bid = {bx, by, bz};
    blk.arrived.assign(1 + nwarp, 0);
    blk.phase.assign(1 + nwarp, 0);
    int i = 0;
    for (unsigned tz = 0; tz < block.z; ++tz)
    for (unsigned ty = 0; ty < block.y; ++ty)
    for (unsigned tx = 0; tx < block.x; ++tx, ++i) {
      Fiber& f = blk.fibers[i];
      f.tid = {tx, ty, tz};
      f.done = false;
      getcontext(&f.ctx);
      f.ctx.uc_stack.ss_sp = f.stack.data();
      f.ctx.uc_stack.ss_size = f.stack.size();
      f.ctx.uc_link = &blk.sched;
      makecontext(&f.ctx, (void (*)())trampoline, 0);
    }
    int remaining = nt;
    while (remaining > 0) {
      int progressed = 0;
      for (int kk = 0; kk < nt; ++kk) {
        const int k = rev ? nt - 1 - kk : kk;
        if (blk.fibers[k].done) continue;
        blk.cur = k;
        swapcontext(&blk.sched, &blk.fibers[k].ctx);
        if (blk.fibers[k].done) { --remaining; }
        ++progressed;
      }
      if (!progressed) break;
    }
  }
  B() = nullptr;
}
}  // namespace emu

#define threadIdx (emu::B()->fibers[emu::B()->cur].tid)
#define blockIdx (emu::B()->bid)
#define blockDim (emu::B()->bdim)
#define gridDim (emu::B()->gdim)

static inline void __syncthreads() {
  emu::Block* b = emu::B();
  emu::barrier(0, (int)(b->bdim.x * b->bdim.y * b->bdim.z));
}
static inline void __syncwarp(unsigned = 0xffffffffu) {}
static inline void __threadfence() {}

// warp exchange: all 32 lanes of the warp must call (full mask), as in the real kernels
template <class T> static inline T emu_exchange(T v, int src_lane_in_warp, std::vector<T>& slots) {
  emu::Block* b = emu::B();
  const int t = emu::linear_tid();
  const int nt = (int)(b->bdim.x * b->bdim.y * b->bdim.z);
  const int w = t / 32;
  const int lanes = std::min(32, nt - w * 32);
  slots[t] = v;
  emu::barrier(1 + w, lanes);
  int src = w * 32 + src_lane_in_warp;
  T r = (src_lane_in_warp >= 0 && src_lane_in_warp < lanes) ? slots[src] : v;
  emu::barrier(1 + w, lanes);
  return r;
}
static inline float __shfl_xor_sync(unsigned, float v, int m) { int l = emu::linear_tid() % 32; return emu_exchange<float>(v, l ^ m, emu::B()->shfl_f); }
static inline int __shfl_xor_sync(unsigned, int v, int m) { int l = emu::linear_tid() % 32; return emu_exchange<int>(v, l ^ m, emu::B()->shfl_i); }
static inline float __shfl_down_sync(unsigned, float v, int d) { int l = emu::linear_tid() % 32; return emu_exchange<float>(v, l + d, emu::B()->shfl_f); }
static inline int __shfl_down_sync(unsigned, int v, int d) { int l = emu::linear_tid() % 32; return emu_exchange<int>(v, l + d, emu::B()->shfl_i); }
static inline float __shfl_up_sync(unsigned, float v, int d) { int l = emu::linear_tid() % 32; return emu_exchange<float>(v, l - d, emu::B()->shfl_f); }
static inline int __shfl_up_sync(unsigned, int v, int d) { int l = emu::linear_tid() % 32; return emu_exchange<int>(v, l - d, emu::B()->shfl_i); }
static inline bool __any_sync(unsigned, bool p) {
  int r = p ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) r |= emu_exchange<int>(r, (emu::linear_tid() % 32) ^ o, emu::B()->shfl_i);
  return r != 0;
}
static inline float __shfl_sync(unsigned, float v, int s) { return emu_exchange<float>(v, s, emu::B()->shfl_f); }
static inline int __shfl_sync(unsigned, int v, int s) { return emu_exchange<int>(v, s, emu::B()->shfl_i); }

static inline float atomicAdd(float* a, float v) { float o = *a; *a = o + v; return o; }
static inline int atomicAdd(int* a, int v) { int o = *a; *a = o + v; return o; }
static inline unsigned atomicAdd(unsigned* a, unsigned v) { unsigned o = *a; *a = o + v; return o; }
static inline int atomicMax(int* a, int v) { int o = *a; *a = std::max(o, v); return o; }
static inline unsigned atomicMax(unsigned* a, unsigned v) { unsigned o = *a; *a = std::max(o, v); return o; }
static inline int atomicMin(int* a, int v) { int o = *a; *a = std::min(o, v); return o; }

static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline float __uint_as_float(unsigned i) { float f; memcpy(&f, &i, 4); return f; }
#define __expf(x) expf(x)
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __saturatef(float x) { return x < 0 ? 0 : (x > 1 ? 1 : x); }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline float __fsqrt_rn(float a) { volatile float r = sqrtf(a); return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }

#define LNST_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>( \
    (reinterpret_cast<uintptr_t>(emu::B()->dyn.data()) + 15) & ~uintptr_t(15))
#define LNST_LAUNCH(kernel, grid, block, smem, stream, ...) \
  emu::launch((grid), (block), (smem), [&]() { kernel(__VA_ARGS__); })
