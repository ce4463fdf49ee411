// Dependent-chain latency and warp-throughput of the fp64 / conversion / MUFU instructions the
// disaggregation loop is made of (B200).  nvcc -arch=sm_100a -O3 lat.cu -o lat && ./lat
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
__global__ void chain_dfma(double *o, long long *t, double a, double b) {
  double x = o[threadIdx.x];
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = fma(x, a, b);
  long long t1 = clock64();
  o[threadIdx.x] = x; if (threadIdx.x == 0) t[blockIdx.x] = t1 - t0;
}
__global__ void chain_rcp(double *o, long long *t) {
  double x = o[threadIdx.x];
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) { double y; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); x = y; }
  long long t1 = clock64();
  o[threadIdx.x] = x; if (threadIdx.x == 0) t[blockIdx.x] = t1 - t0;
}
__global__ void chain_f2f(double *o, long long *t) {
  double x = o[threadIdx.x];
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) { float f; asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f) : "d"(x)); double y; asm volatile("cvt.f64.f32 %0, %1;" : "=d"(y) : "f"(f)); x = y; }
  long long t1 = clock64();
  o[threadIdx.x] = x; if (threadIdx.x == 0) t[blockIdx.x] = t1 - t0;
}
// throughput: 8 independent chains per thread
template <int OP>
__global__ void tput(double *o, long long *t, double a, double b) {
  double x[8];
  for (int k = 0; k < 8; ++k) x[k] = o[threadIdx.x] + k;
  long long t0 = clock64();
  for (int i = 0; i < N / 8; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (OP == 0) x[k] = fma(x[k], a, b);
      if (OP == 1) { double y; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x[k])); x[k] = y; }
      if (OP == 2) { float f; asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f) : "d"(x[k])); x[k] += (double)__float_as_int(f) * 0; asm volatile("" : "+d"(x[k])); }
    }
  }
  long long t1 = clock64();
  double s = 0; for (int k = 0; k < 8; ++k) s += x[k];
  o[threadIdx.x] = s; if (threadIdx.x == 0) t[blockIdx.x] = t1 - t0;
}
int main() {
  double *o; long long *t; cudaMalloc(&o, 1024 * 8); cudaMalloc(&t, 1024 * 8); cudaMemset(o, 0x3f, 1024 * 8);
  long long h;
  for (int rep = 0; rep < 2; ++rep) {
    chain_dfma<<<1, 32>>>(o, t, 1.0000001, 1e-9); cudaMemcpy(&h, t, 8, cudaMemcpyDeviceToHost); printf("DFMA dependent latency %.2f cyc\n", (double)h / N);
    chain_rcp<<<1, 32>>>(o, t); cudaMemcpy(&h, t, 8, cudaMemcpyDeviceToHost); printf("MUFU.RCP64H dependent latency %.2f cyc\n", (double)h / N);
    chain_f2f<<<1, 32>>>(o, t); cudaMemcpy(&h, t, 8, cudaMemcpyDeviceToHost); printf("F2F f64->f32->f64 round trip latency %.2f cyc\n", (double)h / N);
    for (int w = 1; w <= 16; w *= 2) {
      tput<0><<<1, 128 * w>>>(o, t, 1.0000001, 1e-9); cudaMemcpy(&h, t, 8, cudaMemcpyDeviceToHost);
      printf("DFMA   %2d warps/SMSP: %.2f cyc per warp-instr per SMSP\n", w, (double)h / N / w);
      tput<1><<<1, 128 * w>>>(o, t, 0, 0); cudaMemcpy(&h, t, 8, cudaMemcpyDeviceToHost);
      printf("RCP64H %2d warps/SMSP: %.2f cyc per warp-instr per SMSP\n", w, (double)h / N / w);
      tput<2><<<1, 128 * w>>>(o, t, 0, 0); cudaMemcpy(&h, t, 8, cudaMemcpyDeviceToHost);
      printf("F2F    %2d warps/SMSP: %.2f cyc per warp-instr per SMSP (plus a DFMA)\n", w, (double)h / N / w);
    }
  }
  return 0;
}
