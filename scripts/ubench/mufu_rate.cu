// Micro-benchmark: SFU throughput of the activation candidates on sm_100a (ops per clock per SM).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_rate mufu_rate.cu
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

template <int OP>
__global__ void k(float* out, int iters, float seed) {
  float x[8];
  unsigned u[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { x[i] = seed + 0.01f * (threadIdx.x + i); u[i] = __float_as_uint(x[i]) | 0x3c003c00u; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x[i]));
      if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      if (OP == 3) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(u[i]));
      if (OP == 4) asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(u[i]));
      if (OP == 5) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u[i]));
      if (OP == 6) x[i] = fmaf(x[i], 0.999f, 0.001f);
      if (OP == 7) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i])); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[i])); }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i] + __uint_as_float(u[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run(const char* name, int elems_per_op) {
  int dev = 0, sms = 0, khz = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  float* out;
  const int threads = 512, blocks = sms * 2, iters = 20000;
  cudaMalloc(&out, sizeof(float) * threads * blocks);
  k<OP><<<blocks, threads>>>(out, 100, 0.5f);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<OP><<<blocks, threads>>>(out, iters, 0.5f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double ops = double(blocks) * threads * iters * 8.0 * (OP == 7 ? 1 : 1);
  const double per_clk_sm = ops / (ms * 1e-3) / sms / (khz * 1e3);
  printf("%-28s %8.3f ms  %6.2f lane-ops/clk/SM (at the nominal %d MHz)  -> %6.2f elements/clk/SM\n", name, ms, per_clk_sm,
         khz / 1000, per_clk_sm * elems_per_op);
  cudaFree(out);
}

int main() {
  run<0>("tanh.approx.f32", 1);
  run<1>("ex2.approx.ftz.f32", 1);
  run<2>("rcp.approx.ftz.f32", 1);
  run<7>("ex2 + rcp (one SiLU)", 1);
  run<3>("tanh.approx.f16x2", 2);
  run<4>("tanh.approx.bf16x2", 2);
  run<5>("ex2.approx.f16x2", 2);
  run<6>("fma.f32 (reference)", 1);
  return 0;
}
