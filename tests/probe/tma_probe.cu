// Bring-up probe (test infrastructure): which fp32 TMA box configurations load correctly on this GPU?
//   tma_probe <swizzle 0|32|64|128> <box_inner> <box_rows> <rank 2|3> <t0> <f0>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

__device__ __forceinline__ uint32_t s32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__global__ void probe(const __grid_constant__ CUtensorMap tm, float* out, int rank, int c0, int c1, int c2, int bytes) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(bytes));
    if (rank == 3)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                   ::"r"(s32(smem)), "l"(&tm), "r"(s32(&bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(s32(smem)), "l"(&tm), "r"(s32(&bar)), "r"(c0), "r"(c1) : "memory");
  }
  uint32_t ok = 0;
  for (int i = 0; i < 1000000 && !ok; ++i)
    asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(s32(&bar)));
  __syncthreads();
  for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) out[i] = reinterpret_cast<float*>(smem)[i];
  if (threadIdx.x == 0) out[bytes / 4] = ok ? 1.f : 0.f;
}

int main(int argc, char** argv) {
  int swz = atoi(argv[1]), bi = atoi(argv[2]), br = atoi(argv[3]), rank = atoi(argv[4]), t0 = atoi(argv[5]), f0 = atoi(argv[6]);
  const int B = 3, F = 128, T = 1000;
  std::vector<float> h(size_t(B) * F * T);
  for (size_t i = 0; i < h.size(); ++i) h[i] = float(i % 100003);
  float *d, *o;
  cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  int bytes = bi * br * 4;
  cudaMalloc(&o, bytes + 4);
  CUtensorMap tm;
  cuuint64_t gdim[3] = {(cuuint64_t)T, (cuuint64_t)F, (cuuint64_t)B};
  cuuint64_t gstr[2] = {(cuuint64_t)T * 4, (cuuint64_t)F * T * 4};
  cuuint32_t box[3] = {(cuuint32_t)bi, (cuuint32_t)br, 1};
  cuuint32_t es[3] = {1, 1, 1};
  if (rank == 2) { gdim[1] = (cuuint64_t)F * B; }
  CUtensorMapSwizzle sw = swz == 0 ? CU_TENSOR_MAP_SWIZZLE_NONE : swz == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : swz == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
  CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("swz=%d box=%dx%d rank=%d: ENCODE FAILED %d\n", swz, bi, br, rank, (int)r); return 0; }
  probe<<<1, 128, 16384>>>(tm, o, rank, t0, f0, 1, bytes);
  cudaError_t e = cudaGetLastError(); if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("swz=%d box=%dx%d rank=%d t0=%d: KERNEL ERROR %s\n", swz, bi, br, rank, t0, cudaGetErrorString(e)); return 0; }
  std::vector<float> g(bytes / 4 + 1);
  cudaMemcpy(g.data(), o, bytes + 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  if (swz == 0) {
    for (int r2 = 0; r2 < br; ++r2)
      for (int c = 0; c < bi; ++c) {
        size_t row = (rank == 3) ? size_t(1) * F + f0 + r2 : size_t(f0 + r2);
        float want = h[row * T + t0 + c];
        if (g[r2 * bi + c] != want) ++bad;
      }
  }
  printf("swz=%d box=%dx%d rank=%d t0=%d f0=%d: done=%g mismatches=%d first=%g\n", swz, bi, br, rank, t0, f0, g[bytes / 4], bad, g[0]);
  return 0;
}
