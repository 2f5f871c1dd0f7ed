// developer probe: which fp32 3-D TMA box configurations does the hardware accept?   ./tma_selftest bx by bz align
#include <cstdio>
#include <cstdlib>
#include "../../neural-flow-style_b200/csrc/tma_tiles.cuh"
__global__ void k(const __grid_constant__ CUtensorMap m, float* out, int n, int off, int x, int y, int z, int variant) {
  extern __shared__ unsigned char raw[];
  const uint32_t base = ((tma::smem_u32(raw) + 1023u) & ~1023u) + off;
  float* t = reinterpret_cast<float*>(raw + (base - tma::smem_u32(raw)));
  const uint32_t b = base + 65536;                       // barrier inside the dynamic allocation, like conv_tc.cu
  if (threadIdx.x == 0) {
    tma::mbar_init(b, 1);
    tma::fence_mbar_init();
  }
  if (variant & 1) __syncthreads();
  if (threadIdx.x == 0) {
    if (variant & 2) tma::prefetch_map(&m);
    tma::mbar_expect_tx(b, n * 4);
    tma::load_3d(base, &m, b, x, y, z);
  }
  __syncthreads();
  tma::mbar_wait(b, 0);
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = t[i];
}
int main(int argc, char** argv) {
  const int bx = atoi(argv[1]), by = atoi(argv[2]), bz = atoi(argv[3]), off = atoi(argv[4]);
  const int D = 64, H = 64, W = 64;
  float* v; float* o;
  cudaMalloc(&v, D * H * W * 4); cudaMalloc(&o, 1 << 20);
  float* h = (float*)malloc(D * H * W * 4);
  for (int i = 0; i < D * H * W; ++i) h[i] = (float)i;
  cudaMemcpy(v, h, D * H * W * 4, cudaMemcpyHostToDevice);
  CUtensorMap m;
  const int swz = argc > 5 ? atoi(argv[5]) : 0, l2 = argc > 6 ? atoi(argv[6]) : 1, dt = argc > 7 ? atoi(argv[7]) : 7;
  {
    const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D};
    const cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4};
    const cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bz};
    const cuuint32_t ones[3] = {1, 1, 1};
    CUresult r = tma::encode_fn()(&m, (CUtensorMapDataType)dt, 3, v, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  (CUtensorMapSwizzle)swz, (CUtensorMapL2promotion)l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 2; }
  }
  { const unsigned long long* q = (const unsigned long long*)&m; for (int i = 0; i < 16; ++i) printf("%016llx ", q[i]); printf("\n"); }
  const int n = bx * by * bz;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000);
  k<<<1, 128, 70000, 0>>>(m, o, n, off, argc > 9 ? atoi(argv[9]) : -1, 2, 3, argc > 8 ? atoi(argv[8]) : 0);
  cudaError_t e = cudaDeviceSynchronize();
  float r[4];
  if (e == cudaSuccess) cudaMemcpy(r, o, 16, cudaMemcpyDeviceToHost);
  printf("swz %d l2 %d dt %d box %d %d %d off %d: %s  first %g %g %g\n", swz, l2, dt, bx, by, bz, off, cudaGetErrorString(e), r[0], r[1], r[2]);
  return 0;
}
