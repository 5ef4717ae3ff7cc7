// icache_probe.cu — how many instructions per cycle an SM sub-partition sustains on a straight-line loop body of N
// instructions (N x 16 bytes of code), for 1..4 warps per sub-partition; and the dependent-issue latency / throughput of
// DFMA.  Diagnostic for the rollout kernel (its step loop is ~25 KB of code).
#include <cstdio>
#include <cuda_runtime.h>

template <int N> __global__ void body(float* out, int iters, float a, float b) {
  float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < N / 8; i++) {
      x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
      x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
__global__ void dfma_dep(double* out, int iters, double a, double b) {
  double x = threadIdx.x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 64; i++) x = fma(x, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
__global__ void dfma_ind(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
__global__ void ddiv_dep(double* out, int iters, double a) {
  double x = threadIdx.x + 2.0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) x = a / x + 1.5;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

template <typename F> float time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}
template <int N> void run(float* d, int sms, double mhz) {
  for (int warps = 1; warps <= 4; warps *= 2) {
    const int threads = 128 * warps;
    const int iters = (1 << 22) / N;
    float ms = time_ms([&] { body<N><<<sms, threads>>>(d, iters, 1.0001f, 0.5f); });
    const double inst = (double)N * iters * (threads / 32);  // warp instructions per SM
    const double cycles = ms * 1e-3 * mhz * 1e6;
    printf("body %6d instr (%4d KB) warps/SMSP %d: IPC/SM %.2f\n", N, N * 16 / 1024, warps, inst / cycles);
  }
}
int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  const double mhz = p.clockRate / 1000.0;
  printf("%s, %d SMs, %.0f MHz (nominal)\n", p.name, sms, mhz);
  float* d;
  cudaMalloc(&d, 1 << 24);
  run<256>(d, sms, mhz); run<512>(d, sms, mhz); run<1024>(d, sms, mhz); run<1536>(d, sms, mhz); run<2048>(d, sms, mhz);
  run<3072>(d, sms, mhz); run<4096>(d, sms, mhz); run<6144>(d, sms, mhz); run<8192>(d, sms, mhz); run<16384>(d, sms, mhz);
  double* dd = (double*)d;
  for (int warps = 1; warps <= 4; warps *= 2) {
    const int threads = 128 * warps, iters = 20000;
    float ms = time_ms([&] { dfma_dep<<<sms, threads>>>(dd, iters, 1.0000001, 0.5); });
    printf("DFMA dependent chain, %d warps/SMSP: %.1f cycles per DFMA per warp\n", warps, ms * 1e-3 * mhz * 1e6 / (64.0 * iters));
    ms = time_ms([&] { dfma_ind<<<sms, threads>>>(dd, iters, 1.0000001, 0.5); });
    printf("DFMA independent x8, %d warps/SMSP: %.2f DFMA warp-instr per cycle per SM\n", warps, 64.0 * iters * (threads / 32) / (ms * 1e-3 * mhz * 1e6));
    ms = time_ms([&] { ddiv_dep<<<sms, threads>>>(dd, iters / 8, 3.7); });
    printf("double division dependent chain, %d warps/SMSP: %.0f cycles per division(+add) per warp\n", warps, ms * 1e-3 * mhz * 1e6 / (16.0 * (iters / 8)));
  }
  return 0;
}
