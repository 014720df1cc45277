// tools/ubench_fma.cu — issue-rate micro-benchmark: scalar FFMA vs packed fma.rn.f32x2 (FFMA2) vs MUFU.EX2 on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/ubench_fma.cu -o tools/bin/ubench_fma
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096, CHAINS = 8;

__global__ void k_ffma(float* out, float a, float b)
{
  float v[CHAINS];
#pragma unroll
  for(int i = 0; i < CHAINS; i++) v[i] = threadIdx.x + i;
  for(int it = 0; it < ITERS; it++)
#pragma unroll
    for(int i = 0; i < CHAINS; i++) v[i] = __fmaf_rn(v[i], a, b);
  float s = 0;
#pragma unroll
  for(int i = 0; i < CHAINS; i++) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ unsigned long long fma2(unsigned long long x, unsigned long long a, unsigned long long b)
{
  unsigned long long d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(x), "l"(a), "l"(b));
  return d;
}

__global__ void k_ffma2(float* out, float a, float b)
{
  unsigned long long v[CHAINS], a2, b2;
  asm("mov.b64 %0, {%1,%1};" : "=l"(a2) : "f"(a));
  asm("mov.b64 %0, {%1,%1};" : "=l"(b2) : "f"(b));
#pragma unroll
  for(int i = 0; i < CHAINS; i++) { float f = threadIdx.x + i; asm("mov.b64 %0, {%1,%1};" : "=l"(v[i]) : "f"(f)); }
  for(int it = 0; it < ITERS; it++)
#pragma unroll
    for(int i = 0; i < CHAINS; i++) v[i] = fma2(v[i], a2, b2);
  float s = 0;
#pragma unroll
  for(int i = 0; i < CHAINS; i++) { float lo, hi; asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v[i])); s += lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_ex2(float* out, float a)
{
  float v[CHAINS];
#pragma unroll
  for(int i = 0; i < CHAINS; i++) v[i] = (threadIdx.x + i) * 1e-3f;
  for(int it = 0; it < ITERS; it++)
#pragma unroll
    for(int i = 0; i < CHAINS; i++) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
  float s = 0;
#pragma unroll
  for(int i = 0; i < CHAINS; i++) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mixed: 6 FFMA + 1 EX2 + 2 ALU-ish (FADD is fma pipe? FMNMX is alu) per step, like a blend inner loop
__global__ void k_mix(float* out, float a, float b)
{
  float v[CHAINS];
#pragma unroll
  for(int i = 0; i < CHAINS; i++) v[i] = (threadIdx.x + i) * 1e-3f;
  for(int it = 0; it < ITERS / 8; it++)
#pragma unroll
    for(int i = 0; i < CHAINS; i++)
    {
      float x = v[i];
#pragma unroll
      for(int k = 0; k < 6; k++) x = __fmaf_rn(x, a, b);
      float e; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x));
      v[i] = fminf(e, x) + b;
    }
  float s = 0;
#pragma unroll
  for(int i = 0; i < CHAINS; i++) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float timeit(F f)
{
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); for(int i = 0; i < 5; i++) f(); cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / 5;
}

int main()
{
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount, blocks = sms * 2, threads = 1024;
  int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  float* out; cudaMalloc(&out, sizeof(float) * blocks * threads);
  const double warpsPerSmsp = 2.0 * threads / 32 / 4;
  auto report = [&](const char* name, float ms, double instrPerThread, double flopsPerInstrLane) {
    const double cyc = ms * 1e-3 * khz * 1e3;                       // at max clock
    const double perSmsp = instrPerThread * warpsPerSmsp;           // warp-instr issued per SMSP
    printf("%-8s %.3f ms  %.2f cycles/warp-instr/SMSP  %.1f TFLOP/s\n", name, ms, cyc / perSmsp,
           flopsPerInstrLane * instrPerThread * blocks * threads / (ms * 1e-3) / 1e12);
  };
  report("FFMA", timeit([&] { k_ffma<<<blocks, threads>>>(out, 1.0001f, 0.5f); }), double(ITERS) * CHAINS, 2);
  report("FFMA2", timeit([&] { k_ffma2<<<blocks, threads>>>(out, 1.0001f, 0.5f); }), double(ITERS) * CHAINS, 4);
  report("EX2", timeit([&] { k_ex2<<<blocks, threads>>>(out, 1.0f); }), double(ITERS) * CHAINS, 1);
  report("MIX9", timeit([&] { k_mix<<<blocks, threads>>>(out, 1.0001f, 0.5f); }), double(ITERS / 8) * CHAINS * 9, 1);
  printf("sms=%d clock=%d kHz %s\n", sms, khz, cudaGetLastError() == cudaSuccess ? "ok" : "ERR");
  return 0;
}
