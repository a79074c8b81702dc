// Microbenchmark: issue / pipe throughput of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) and of a
// blend-like mix (packed FMAs + MUFU.EX2) on sm_100a.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
// -O3 -o ffma2_bench ffma2_bench.cu ; run on the B200 box.  Prints FMA lane-ops / clk / SM.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint64_t pk(float a, float b){ uint64_t r; asm("mov.b64 %0, {%1,%2};":"=l"(r):"f"(a),"f"(b)); return r;}
__device__ __forceinline__ void upk(uint64_t v, float&a, float&b){ asm("mov.b64 {%0,%1}, %2;":"=f"(a),"=f"(b):"l"(v));}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c){ uint64_t r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;":"=l"(r):"l"(a),"l"(b),"l"(c)); return r;}
__device__ __forceinline__ float ex2(float x){ float y; asm volatile("ex2.approx.ftz.f32 %0, %1;":"=f"(y):"f"(x)); return y;}

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b) {
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-3f + i;
  if (MODE == 0) {            // 16 independent scalar FFMA chains
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(acc[i]) : "f"(a), "f"(b));
    }
  } else if (MODE == 1) {     // 8 independent packed chains (same 16 lanes of state)
    uint64_t p[8], pa = pk(a, a), pb = pk(b, b);
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] = pk(acc[2 * i], acc[2 * i + 1]);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], pa, pb);
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], pa, pb);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) upk(p[i], acc[2 * i], acc[2 * i + 1]);
  } else if (MODE == 2) {     // blend-fwd like: per pair of pixels 9 packed FMA + 2 MUFU + 4 scalar ALU-ish
    uint64_t p[8], pa = pk(a, a), pb = pk(b, b);
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] = pk(acc[2 * i], acc[2 * i + 1]);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        uint64_t q0 = fma2(p[4 * i], pa, pb), q1 = fma2(q0, pa, pb), q2 = fma2(q1, q0, pb);
        float x, y; upk(q2, x, y);
        x = ex2(x); y = ex2(y);
        x = (x > 1e-4f) ? x : 0.f; y = (y > 1e-4f) ? y : 0.f;
        uint64_t w = pk(x, y);
        p[4 * i] = fma2(w, pa, p[4 * i]);
        p[4 * i + 1] = fma2(w, pb, p[4 * i + 1]);
        p[4 * i + 2] = fma2(w, q0, p[4 * i + 2]);
        p[4 * i + 3] = fma2(w, q1, p[4 * i + 3]);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) upk(p[i], acc[2 * i], acc[2 * i + 1]);
  } else if (MODE == 3) {     // same mix, scalar
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float q0, q1, q2;
        asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(q0) : "f"(acc[4 * i]), "f"(a), "f"(b));
        asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(q1) : "f"(q0), "f"(a), "f"(b));
        asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(q2) : "f"(q1), "f"(q0), "f"(b));
        float x = ex2(q2);
        x = (x > 1e-4f) ? x : 0.f;
        asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[4 * i]) : "f"(x), "f"(a));
        asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[4 * i + 1]) : "f"(x), "f"(b));
        asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[4 * i + 2]) : "f"(x), "f"(q0));
        asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[4 * i + 3]) : "f"(x), "f"(q1));
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
double run(const char* name, double lane_fma_per_iter, int blocks_per_sm) {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  float* out; cudaMalloc(&out, sizeof(float) * sms * blocks_per_sm * 256);
  const int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<sms * blocks_per_sm, 256>>>(out, 1000, 0.999f, 1e-3f);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0);
    k<MODE><<<sms * blocks_per_sm, 256>>>(out, iters, 0.999f, 1e-3f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  double threads = (double)sms * blocks_per_sm * 256;
  double lane_fma = threads * iters * lane_fma_per_iter;
  double per_clk_sm = lane_fma / (best * 1e-3) / ((double)clk_khz * 1e3) / sms;
  printf("%-28s blocks/SM %d  %.3f ms  %.1f FMA lane-ops/clk/SM (at max clock %d MHz)\n", name, blocks_per_sm, best,
         per_clk_sm, clk_khz / 1000);
  cudaFree(out);
  return per_clk_sm;
}

int main() {
  for (int b : {2, 4, 8}) {
    run<0>("scalar FFMA x16", 16, b);
    run<1>("packed FFMA2 x8 (x2)", 32, b);
    run<3>("blend-like mix, scalar", 4 * 7, b);      // 7 FMA + 1 MUFU + 2 ALU per pixel, 4 pixels
    run<2>("blend-like mix, packed", 2 * 2 * 7, b);  // per pixel pair: 7 FFMA2 + 2 MUFU + 4 ALU + pack
  }
  return 0;
}
