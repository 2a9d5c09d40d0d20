// Third microbenchmark (B200): conversions and shared-memory loads measured so that the compiler cannot
// hoist or simplify them -- every result feeds a loop-carried value.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o microbench3 microbench3.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 2048
#define ILP 8

template <int OP>
__global__ void __launch_bounds__(256) bench(float* out, float seed, int stride) {
  __shared__ double sm[2048];
  float f[ILP];
  double d[ILP];
  int idx[ILP];
  for (int k = 0; k < ILP; ++k) { f[k] = seed + k + threadIdx.x * 1e-3f; d[k] = f[k]; idx[k] = (threadIdx.x + 32 * k) & 1023; }
  for (int k = threadIdx.x; k < 2048; k += 256) sm[k] = (double)((k * stride) & 1023);
  __syncthreads();
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int k = 0; k < ILP; ++k) {
      if (OP == 0) { d[k] = d[k] + (double)f[k]; f[k] = f[k] + 1.0f; }                 // F2F.F64.F32 + DADD + FADD
      if (OP == 1) { d[k] = d[k] + 1.5; f[k] = f[k] + 1.0f; }                          // DADD + FADD (reference)
      if (OP == 2) { f[k] = f[k] + (float)d[k]; d[k] = d[k] + 1.5; }                   // F2F.F32.F64 + FADD + DADD
      if (OP == 3) { double v = sm[idx[k]]; d[k] += v; idx[k] = (idx[k] + stride) & 1023; }   // LDS.64 (dependent address) + DADD
      if (OP == 4) { float v = ((float*)sm)[idx[k]]; f[k] += v; idx[k] = (idx[k] + stride) & 1023; }   // LDS.32 + FADD
      if (OP == 5) { double2 v = ((double2*)sm)[idx[k] & 511]; d[k] += v.x + v.y; idx[k] = (idx[k] + stride) & 1023; }  // LDS.128 + 2 DADD
      if (OP == 6) {  // int widening f32 -> f64 + DADD
        unsigned b = __float_as_uint(f[k]);
        unsigned hi = (((b & 0x7fffffffu) >> 3) + 0x38000000u) | (b & 0x80000000u);
        d[k] = d[k] + __hiloint2double((int)hi, (int)(b << 29));
        f[k] = f[k] + 1.0f;
      }
    }
  }
  float acc = 0.f;
  for (int k = 0; k < ILP; ++k) acc += f[k] + (float)d[k] + (float)idx[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int OP>
void run(const char* name, float* out, int stride = 1) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  int blocks = 148 * 8, threads = 256;
  bench<OP><<<blocks, threads>>>(out, 1.f, stride);
  cudaEventRecord(a);
  bench<OP><<<blocks, threads>>>(out, 1.f, stride);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  double ops = (double)blocks * threads * ITERS * ILP;
  printf("%-52s %8.3f ms  %7.1f lanes/clk/SM at 1.9 GHz\n", name, ms, ops / ms / 1e6 / 148 / 1.9);
}

int main() {
  float* out;
  cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
  run<1>("DADD + FADD (reference)", out);
  run<0>("F2F.F64.F32 + DADD + FADD", out);
  run<2>("F2F.F32.F64 + FADD + DADD", out);
  run<6>("int widen f32->f64 + DADD + FADD", out);
  run<3>("LDS.64 + DADD + index update (conflict free)", out, 1);
  run<4>("LDS.32 + FADD + index update (conflict free)", out, 1);
  run<5>("LDS.128 + 2 DADD + index update", out, 1);
  return 0;
}
