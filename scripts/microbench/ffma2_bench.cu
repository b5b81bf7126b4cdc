// Microbenchmark (B200, sm_100a): issue rate and dependent latency of packed fp32x2 (FFMA2 / FADD2 / FMUL2) against
// scalar FFMA, per SM, for 1..16 warps per scheduler and 1..8 independent chains per thread.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu ; prints one line per case.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, int MODE>
__global__ void k(float* out, float a, float b, int iters, long long* cycles) {
  float2 acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f - i);
  const float2 av = make_float2(a, a * 1.0001f), bv = make_float2(b, b * 0.9999f);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (MODE == 0) {            // scalar FFMA x2 (same flops as one FFMA2)
        acc[i].x = fmaf(acc[i].x, a, b);
        acc[i].y = fmaf(acc[i].y, a, b);
      } else if (MODE == 1) {     // FFMA2, packed operands
        acc[i] = __ffma2_rn(acc[i], av, bv);
      } else if (MODE == 2) {     // FFMA2, broadcast scalar operands
        acc[i] = __ffma2_rn(acc[i], make_float2(a, a), make_float2(b, b));
      } else if (MODE == 3) {     // FADD2
        acc[i] = __fadd2_rn(acc[i], bv);
      } else if (MODE == 4) {     // FMUL2
        acc[i] = __fmul2_rn(acc[i], av);
      } else if (MODE == 5) {     // scalar FFMA x1 (half the flops)
        acc[i].x = fmaf(acc[i].x, a, b);
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int ILP, int MODE>
void run(const char* name, int threads, float* out, long long* cyc) {
  const int iters = 4096;
  k<ILP, MODE><<<148, threads>>>(out, 1.0001f, 1e-6f, iters, cyc);
  cudaDeviceSynchronize();
  k<ILP, MODE><<<148, threads>>>(out, 1.0001f, 1e-6f, iters, cyc);
  cudaDeviceSynchronize();
  long long c;
  cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
  const double warps = threads / 32.0;
  const double inst_per_warp = (double)iters * ILP * (MODE == 0 ? 2 : 1);
  // warp-instructions per cycle per SM, and cycles per instruction per warp (latency when ILP = 1 and one warp per scheduler)
  printf("%-22s ILP=%d warps/SM=%2d  cycles=%9lld  warp-inst/clk/SM=%6.3f  pair-FMA/clk/SM=%7.1f  clk/inst/warp=%6.2f\n", name, ILP, (int)warps, c,
         inst_per_warp * warps / c, (double)iters * ILP * warps * 32 / c, c / inst_per_warp);
}

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaMalloc(&cyc, sizeof(long long));
  const int thr[] = {128, 256, 512, 1024};
  for (int t : thr) {
    run<1, 0>("FFMA x2 (scalar)", t, out, cyc);
    run<1, 1>("FFMA2 packed", t, out, cyc);
    run<1, 2>("FFMA2 broadcast", t, out, cyc);
    run<4, 0>("FFMA x2 (scalar)", t, out, cyc);
    run<4, 1>("FFMA2 packed", t, out, cyc);
    run<4, 2>("FFMA2 broadcast", t, out, cyc);
    run<4, 3>("FADD2", t, out, cyc);
    run<4, 4>("FMUL2", t, out, cyc);
    run<4, 5>("FFMA x1 (scalar)", t, out, cyc);
    run<8, 0>("FFMA x2 (scalar)", t, out, cyc);
    run<8, 1>("FFMA2 packed", t, out, cyc);
  }
  return 0;
}
