// fp64 tensor-core (mma.sync f64) throughput probe for sm_100a: independent accumulator chains per warp.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) dmma884(double* out, int iters) {
  double c[8][2]; for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0; for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) dmma1688(double* out, int iters) {
  double c[4][4]; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) c[i][j] = threadIdx.x + i + j;
  double a0 = 1.0 + threadIdx.x * 1e-9, a1 = a0 * 0.5, a2 = a0 * 0.25, a3 = a0 * 0.125, b0 = 1.0 - threadIdx.x * 1e-9, b1 = b0 * 0.5;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
  }
  double s = 0; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) dfma(double* out, int iters) {
  double x[8]; for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i; const double a = 1.0000001, b = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0; for (int i = 0; i < 8; ++i) s += x[i]; out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class F> double timeit(F f) { cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); f(); cudaDeviceSynchronize(); cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); return ms; }
int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0); const int grid = sms * 8, iters = 20000; double* out; cudaMalloc(&out, (size_t)grid * 256 * 8);
  double ms = timeit([&] { dmma884<<<grid, 256>>>(out, iters); });
  printf("dmma m8n8k4 : %.2f TFLOP/s\n", (double)grid * 8 * iters * 8 * 512 / (ms * 1e-3) / 1e12);
  ms = timeit([&] { dmma1688<<<grid, 256>>>(out, iters); });
  printf("dmma m16n8k8: %.2f TFLOP/s\n", (double)grid * 8 * iters * 4 * 2048 / (ms * 1e-3) / 1e12);
  ms = timeit([&] { dfma<<<grid, 256>>>(out, iters); });
  printf("dfma        : %.2f TFLOP/s\n", (double)grid * 256 * iters * 64 * 2 / (ms * 1e-3) / 1e12);
  return 0;
}
