// Second instruction-throughput microbenchmark (B200): the candidate building blocks of the exact
// window-sum arithmetic -- 64-bit integer adds, int64<->f64 magic conversions, FP64 ops, conversions,
// shared-memory 32/64/128-bit loads -- alone and mixed, to see which pipes overlap.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o microbench2 microbench2.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 2048
#define ILP 8

template <int OP>
__global__ void __launch_bounds__(256) bench(float* out, float seed, long long lseed) {
  __shared__ double sm[256 * 4 + 64];
  float f[ILP];
  double d[ILP];
  long long q[ILP];
  unsigned u[ILP];
  for (int k = 0; k < ILP; ++k) {
    f[k] = seed + k + threadIdx.x * 1e-3f; d[k] = f[k]; u[k] = __float_as_uint(f[k]); q[k] = lseed + k * 77 + threadIdx.x;
  }
  for (int k = threadIdx.x; k < 256 * 4 + 64; k += 256) sm[k] = k * seed;
  __syncthreads();
  const double c1 = 1.000001, c2 = 0.999999;
  const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int k = 0; k < ILP; ++k) {
      if (OP == 0) q[k] = q[k] + (q[(k + 1) % ILP] ^ lseed);                 // 64-bit add (+xor)
      if (OP == 1) { q[k] += lseed; }                                        // 64-bit add only
      if (OP == 2) {  // f64 -> biased int64 via magic DFMA, accumulate in int64
        double t = fma(d[k], 1048576.0, MAGIC);
        q[k] += __double_as_longlong(t);
        d[k] = d[k] + c1;
      }
      if (OP == 3) {  // int64 -> f64 via magic OR + DADD
        long long s = q[k] + lseed;
        double t = __longlong_as_double(s | 0x4330000000000000ll) - 4503599627370496.0;
        d[k] = t * c2;
        q[k] = s ^ __double2hiint(d[k]);
      }
      if (OP == 4) d[k] = d[k] * c2;                                         // DMUL
      if (OP == 5) { f[k] = (float)d[k]; d[k] = d[k] + c1; }                 // F2F.F32.F64 + DADD
      if (OP == 6) { d[k] = (double)f[k] + c1; f[k] += 1.f; }                // F2F.F64.F32 + DADD + FADD
      if (OP == 7) { d[k] += sm[(threadIdx.x + k * 32 + (i & 3) * 256) & 1023]; }  // LDS.64 + DADD
      if (OP == 8) { f[k] += ((float*)sm)[(threadIdx.x + k * 32 + (i & 3) * 256) & 2047]; }  // LDS.32 + FADD
      if (OP == 9) {  // mixed: LDS.64 x2, 64-bit add x2 (sliding int64 window)
        const long long* s = (const long long*)sm;
        q[k] += s[(threadIdx.x + k * 32 + (i & 3) * 256) & 1023] - s[(threadIdx.x + k * 32 + 17 + (i & 3) * 256) & 1023];
      }
      if (OP == 10) {  // full "mean" chain: int64 diff -> f64 -> mul -> f32 -> FADD/FMUL/FADD
        long long s = q[k] + lseed;
        double t = (__longlong_as_double(s | 0x4330000000000000ll) - 4503599627370496.0) * c2;
        float m = (float)t;
        f[k] = f[k] + 0.5f * (seed - m);
        q[k] = s;
      }
      if (OP == 11) { u[k] = (u[k] << 3) + 0x38000000u; }                    // SHF/IADD
      if (OP == 12) { f[k] = f[k] * 1.0001f; }                               // FMUL
      if (OP == 13) { q[k] = (long long)__float2ll_rn(f[k]) + q[k]; f[k] += 1.f; }  // F2I.S64
      if (OP == 14) { d[k] = (double)q[k] + d[k]; q[k] += lseed; }            // I2F.F64.S64
    }
  }
  float acc = 0.f;
  for (int k = 0; k < ILP; ++k) acc += f[k] + (float)d[k] + (float)u[k] + (float)q[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int OP>
void run(const char* name, float* out, int blocks_per_sm = 8) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  int blocks = 148 * blocks_per_sm, threads = 256;
  bench<OP><<<blocks, threads>>>(out, 1.f, 12345);
  cudaEventRecord(a);
  bench<OP><<<blocks, threads>>>(out, 1.f, 12345);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  double ops = (double)blocks * threads * ITERS * ILP;
  printf("%-44s %8.3f ms  %8.2f Gop/s/SM  (%.1f lanes/clk/SM at 1.9 GHz)\n", name, ms, ops / ms / 1e6 / 148,
         ops / ms / 1e6 / 148 / 1.9);
}

int main() {
  float* out;
  cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
  run<0>("int64 add + xor", out);
  run<1>("int64 add", out);
  run<2>("f64->int64 magic DFMA + int64 add + DADD", out);
  run<3>("int64->f64 magic (add,or,DADD,DMUL,xor)", out);
  run<4>("DMUL", out);
  run<5>("F2F.F32.F64 + DADD", out);
  run<6>("F2F.F64.F32 + DADD + FADD", out);
  run<7>("LDS.64 + DADD", out);
  run<8>("LDS.32 + FADD", out);
  run<9>("2x LDS.64 + 2x int64 add", out);
  run<10>("mean chain (int64->f64, DMUL, F2F, 3 f32)", out);
  run<11>("SHF+IADD", out);
  run<12>("FMUL", out);
  run<13>("F2I.S64.F32 + int64 add + FADD", out);
  run<14>("I2F.F64.S64 + DADD + int64 add", out);
  return 0;
}
