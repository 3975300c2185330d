// tools/pipe_peaks.cu -- measured issue peaks of the pipes the compute-bound
// samplers live on ("FP64/FP32/INT peaks are not in MEASURED_PEAKS.json:
// measure with a DFMA/FFMA/IADD microbenchmark on the box and record").
// Each kernel runs 8 independent dependent chains per thread, 1024 threads per
// CTA, 2 CTAs per SM (64 warps per SM), so the pipe, not latency, is the limit.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_peaks tools/pipe_peaks.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kIters = 4096;

__global__ void __launch_bounds__(1024, 2) k_dfma(double* out, double a, double b) {
  double x[8];
  for (int j = 0; j < 8; ++j) x[j] = threadIdx.x + j;
  for (int i = 0; i < kIters; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = __fma_rn(x[j], a, b);
  double s = 0; for (int j = 0; j < 8; ++j) s += x[j];
  if (s == 12345.678) out[0] = s;
}
__global__ void __launch_bounds__(1024, 2) k_ffma(float* out, float a, float b) {
  float x[8];
  for (int j = 0; j < 8; ++j) x[j] = threadIdx.x + j;
  for (int i = 0; i < kIters; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = __fmaf_rn(x[j], a, b);
  float s = 0; for (int j = 0; j < 8; ++j) s += x[j];
  if (s == 12345.678f) out[0] = s;
}
__global__ void __launch_bounds__(1024, 2) k_lop3(unsigned* out, unsigned a, unsigned b) {
  unsigned x[8];
  for (int j = 0; j < 8; ++j) x[j] = threadIdx.x + j;
  for (int i = 0; i < kIters; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = (x[j] ^ a ^ (x[(j + 1) & 7] & b));   // one LOP3 each
  unsigned s = 0; for (int j = 0; j < 8; ++j) s += x[j];
  if (s == 12345u) out[0] = s;
}
__global__ void __launch_bounds__(1024, 2) k_mufu(float* out, float a) {
  float x[8];
  for (int j = 0; j < 8; ++j) x[j] = 1.0f + threadIdx.x * 1e-3f + j;
  for (int i = 0; i < kIters; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x[j])); x[j] = y + a; }
  float s = 0; for (int j = 0; j < 8; ++j) s += x[j];
  if (s == 12345.678f) out[0] = s;
}

template <typename F> double time_ms(F launch) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount, grid = sms * 2;
  void* buf; cudaMalloc(&buf, 64);
  const double ops = double(grid) * 1024 * 8 * kIters;       // thread-level instructions
  double t;
  t = time_ms([&] { k_dfma<<<grid, 1024>>>((double*)buf, 1.0000001, 1e-9); });
  printf("DFMA : %.3f ms  %.2f T thread-inst/s  = %.1f lanes/clk/SM at %.0f MHz\n", t, ops / t / 1e9, ops / (t * 1e-3) / sms / (p.clockRate * 1e3), p.clockRate / 1e3);
  t = time_ms([&] { k_ffma<<<grid, 1024>>>((float*)buf, 1.0000001f, 1e-9f); });
  printf("FFMA : %.3f ms  %.2f T thread-inst/s  = %.1f lanes/clk/SM\n", t, ops / t / 1e9, ops / (t * 1e-3) / sms / (p.clockRate * 1e3));
  t = time_ms([&] { k_lop3<<<grid, 1024>>>((unsigned*)buf, 0x9e3779b9u, 0xdeadbeefu); });
  printf("LOP3 : %.3f ms  %.2f T thread-inst/s  = %.1f lanes/clk/SM\n", t, ops / t / 1e9, ops / (t * 1e-3) / sms / (p.clockRate * 1e3));
  t = time_ms([&] { k_mufu<<<grid, 1024>>>((float*)buf, 0.5f); });
  printf("MUFU+FADD pair : %.3f ms  %.2f T pairs/s  = %.1f lanes/clk/SM\n", t, ops / t / 1e9, ops / (t * 1e-3) / sms / (p.clockRate * 1e3));
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
