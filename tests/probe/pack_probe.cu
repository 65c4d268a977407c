// Throughput probe (sm_100a): how many fp32->bf16x2 packs (F2FP.BF16.F32.PACK_AB), MUFU.EX2 and integer-pipe packs
// (two IADD + PRMT) one SM retires per clock.  Bring-up tool, not part of the library:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pack_probe tests/probe/pack_probe.cu && /tmp/pack_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void probe(uint32_t* out, float a, float b, int iters, long long* cyc) {
  float x[8];
  uint32_t acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { x[i] = a + float(threadIdx.x) * 1e-3f + float(i); acc[i] = 0; }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) {            // F2FP pack
        uint32_t r;
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x[i]), "f"(x[(i + 1) & 7]));
        acc[i] += r;
      } else if (MODE == 1) {     // MUFU.EX2
        float r;
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x[i]));
        acc[i] += __float_as_uint(r);
      } else if (MODE == 2) {     // integer-pipe pack: round half up, take the upper halves
        const uint32_t u0 = __float_as_uint(x[i]) + 0x8000u, u1 = __float_as_uint(x[(i + 1) & 7]) + 0x8000u;
        uint32_t r;
        asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(r) : "r"(u0), "r"(u1));
        acc[i] += r;
      } else if (MODE == 3) {     // F2FP pack + EX2 together (do they share a pipe?)
        uint32_t r;
        float e;
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x[i]), "f"(x[(i + 1) & 7]));
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x[i]));
        acc[i] += r + __float_as_uint(e);
      } else {                    // baseline: only the IADD
        acc[i] += __float_as_uint(x[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] += b;
  }
  const long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, uint32_t* out, long long* cyc, int threads) {
  const int iters = 4096, blocks = 148;
  probe<MODE><<<blocks, threads>>>(out, 1.0f, 1e-4f, 16, cyc);
  probe<MODE><<<blocks, threads>>>(out, 1.0f, 1e-4f, iters, cyc);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < blocks; ++i) avg += double(h[i]);
  avg /= blocks;
  printf("%-28s %4d threads/SM: %8.0f clk, %6.1f thread-ops/clk/SM\n", name, threads, avg,
         double(threads) * iters * 8 / avg);
}

int main() {
  uint32_t* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 148 * 8);
  for (int threads : {256, 1024}) {
    run<4>("IADD only (baseline)", out, cyc, threads);
    run<0>("F2FP.BF16.PACK_AB + IADD", out, cyc, threads);
    run<1>("MUFU.EX2 + IADD", out, cyc, threads);
    run<2>("2 IADD + PRMT + IADD", out, cyc, threads);
    run<3>("F2FP + MUFU.EX2 + 2 IADD", out, cyc, threads);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
