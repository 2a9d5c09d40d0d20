// Instruction-throughput microbenchmark used to size the fused kernel's arithmetic budget on B200:
// FP64 add/fma, f32<->f64 conversions (F2F) and their integer-pipe replacements.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o microbench microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define ILP 8

template <int OP>
__global__ void bench(float* out, float seed) {
  float f[ILP];
  double d[ILP];
  unsigned u[ILP];
  for (int k = 0; k < ILP; ++k) { f[k] = seed + k + threadIdx.x * 1e-3f; d[k] = f[k]; u[k] = __float_as_uint(f[k]); }
  const double c1 = 1.000001, c2 = 0.999999;
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int k = 0; k < ILP; ++k) {
      if (OP == 0) f[k] = f[k] + 1.0001f;                                   // FADD
      if (OP == 1) d[k] = d[k] + c1;                                        // DADD
      if (OP == 2) d[k] = fma(d[k], c2, c1);                                // DFMA
      if (OP == 3) { d[k] = (double)f[k]; f[k] = f[k] + (float)(int)(d[k] > 3.0); }  // F2F.F64.F32 (+FADD)
      if (OP == 4) { f[k] = (float)d[k]; d[k] = d[k] + (double)0.0 + c1 * (f[k] > 1e30f); }  // F2F.F32.F64 (+pred)
      if (OP == 5) {  // integer widening f32 -> f64 bits (normal numbers)
        unsigned b = u[k];
        unsigned hi = ((b & 0x7fffffffu) >> 3) + 0x38000000u;
        hi |= b & 0x80000000u;
        unsigned lo = b << 29;
        d[k] = __hiloint2double((int)hi, (int)lo);
        u[k] = b + (unsigned)(d[k] > 1e300);
      }
      if (OP == 6) {  // magic-number round of f64 to the f32 grid (2 DADD + 2 ALU)
        int hi = __double2hiint(d[k]);
        double m = __hiloint2double((hi & 0xfff00000) + (29 << 20), 0);
        d[k] = (d[k] + m) - m + c1;
      }
      if (OP == 7) f[k] = __shfl_xor_sync(0xffffffffu, f[k], 1) + 1.f;       // SHFL
    }
  }
  float acc = 0.f;
  for (int k = 0; k < ILP; ++k) acc += f[k] + (float)d[k] + (float)u[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int OP>
void run(const char* name, float* out) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  int blocks = 148 * 8, threads = 256;
  bench<OP><<<blocks, threads>>>(out, 1.f);
  cudaEventRecord(a);
  bench<OP><<<blocks, threads>>>(out, 1.f);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  double ops = (double)blocks * threads * ITERS * ILP;
  printf("%-28s %8.3f ms  %8.2f Gop/s/SM  (%.1f lanes/clk/SM at 1.9 GHz)\n", name, ms, ops / ms / 1e6 / 148,
         ops / ms / 1e6 / 148 / 1.9);
}

int main() {
  float* out;
  cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
  run<0>("FADD", out);
  run<1>("DADD", out);
  run<2>("DFMA", out);
  run<3>("F2F f32->f64 (+FADD)", out);
  run<4>("F2F f64->f32 (+DADD,FSETP)", out);
  run<5>("int widen f32->f64 (+DSETP)", out);
  run<6>("magic round (2 DADD+2 ALU+DADD)", out);
  run<7>("SHFL (+FADD)", out);
  return 0;
}
