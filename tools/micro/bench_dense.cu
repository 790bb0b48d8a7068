// Microbenchmark (dev tool, not part of the library): cycles of the 128x64x64 smem GEMM
// building blocks of the update kernel under different thread tilings.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -I pantheonrl_b200/csrc \
//        tools/micro/bench_dense.cu -o gpurun_out/bench_dense && gpurun_out/bench_dense
#include <cstdio>
#include <vector>
#include "pth_mlp.cuh"
void pth_set_error(const char*, ...) {}
using namespace pthmlp;

// variant: dense layer with SPT samples x JT outputs per thread, explicit prefetch depth
template <bool TANH, int NTH, int SPT, int PF>
__device__ __forceinline__ void dense_v(const float* A, const float* W, const float* bias, float* Out, int tid) {
  constexpr int BTS = 128, LDA_ = BTS + 4;
  constexpr int TXN = BTS / SPT, NY = NTH / TXN, JT = HID / NY;
  const int tx = tid % TXN, ty = tid / TXN;
  float acc[JT][SPT];
#pragma unroll
  for (int jj = 0; jj < JT; ++jj) {
    const float bj = bias[jj * NY + ty];
#pragma unroll
    for (int ss = 0; ss < SPT; ++ss) acc[jj][ss] = bj;
  }
  auto lda = [&](int k, float (&a)[SPT]) {
    if constexpr (SPT == 8) {
      const float4 a0 = *reinterpret_cast<const float4*>(A + k * LDA_ + tx * 4);
      const float4 a1 = *reinterpret_cast<const float4*>(A + k * LDA_ + 64 + tx * 4);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
    } else {
      const float4 a0 = *reinterpret_cast<const float4*>(A + k * LDA_ + tx * 4);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
    }
  };
  if constexpr (PF == 0) {
#pragma unroll 2
    for (int k0 = 0; k0 < HID; k0 += 4) {
      float4 w[JT];
#pragma unroll
      for (int jj = 0; jj < JT; ++jj) w[jj] = *reinterpret_cast<const float4*>(W + (jj * NY + ty) * LDW + k0);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        float a[SPT];
        lda(k0 + kk, a);
#pragma unroll
        for (int jj = 0; jj < JT; ++jj) {
          const float wk = kk == 0 ? w[jj].x : (kk == 1 ? w[jj].y : (kk == 2 ? w[jj].z : w[jj].w));
#pragma unroll
          for (int ss = 0; ss < SPT; ++ss) acc[jj][ss] = fmaf(a[ss], wk, acc[jj][ss]);
        }
      }
    }
  } else {
    // explicit register double buffering: loads of step k0+4 are issued before the FMAs of k0
    float4 w[2][JT];
    float a[2][4][SPT];
#pragma unroll
    for (int jj = 0; jj < JT; ++jj) w[0][jj] = *reinterpret_cast<const float4*>(W + (jj * NY + ty) * LDW);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) lda(kk, a[0][kk]);
#pragma unroll
    for (int it = 0; it < HID / 4; ++it) {
      const int cur = it & 1, nxt = cur ^ 1;
      if (it + 1 < HID / 4) {
#pragma unroll
        for (int jj = 0; jj < JT; ++jj)
          w[nxt][jj] = *reinterpret_cast<const float4*>(W + (jj * NY + ty) * LDW + (it + 1) * 4);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) lda((it + 1) * 4 + kk, a[nxt][kk]);
      }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
#pragma unroll
        for (int jj = 0; jj < JT; ++jj) {
          const float wk = kk == 0 ? w[cur][jj].x : (kk == 1 ? w[cur][jj].y : (kk == 2 ? w[cur][jj].z : w[cur][jj].w));
#pragma unroll
          for (int ss = 0; ss < SPT; ++ss) acc[jj][ss] = fmaf(a[cur][kk][ss], wk, acc[jj][ss]);
        }
    }
  }
#pragma unroll
  for (int jj = 0; jj < JT; ++jj) {
    float* o = Out + (jj * NY + ty) * LDA_;
    float v[SPT];
#pragma unroll
    for (int ss = 0; ss < SPT; ++ss) v[ss] = TANH ? pth_tanhf(acc[jj][ss]) : acc[jj][ss];
    *reinterpret_cast<float4*>(o + tx * 4) = make_float4(v[0], v[1], v[2], v[3]);
    if constexpr (SPT == 8) *reinterpret_cast<float4*>(o + 64 + tx * 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
}

struct Sm {
  float A[HID * LDA], B[HID * LDA], W[HID * LDW], bias[HID];
};

template <bool TANH, int NTH, int SPT, int PF>
__global__ void __launch_bounds__(NTH) k_dense(long long* out, int reps) {
  extern __shared__ __align__(16) unsigned char raw[];
  Sm& sm = *reinterpret_cast<Sm*>(raw);
  const int tid = threadIdx.x;
  for (int i = tid; i < HID * LDA; i += NTH) { sm.A[i] = 0.001f * (i % 97); sm.B[i] = 0.f; }
  for (int i = tid; i < HID * LDW; i += NTH) sm.W[i] = 0.01f * ((i % 13) - 6);
  if (tid < HID) sm.bias[tid] = 0.1f;
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    dense_v<TANH, NTH, SPT, PF>(sm.A, sm.W, sm.bias, sm.B, tid);
    __syncthreads();
    dense_v<TANH, NTH, SPT, PF>(sm.B, sm.W, sm.bias, sm.A, tid);
    __syncthreads();
  }
  const long long t1 = clock64();
  if (tid == 0) out[blockIdx.x] = (t1 - t0) / (2 * reps);
}

template <bool TANH, int NTH, int SPT, int PF>
void run(const char* name) {
  long long* d;
  cudaMalloc(&d, 148 * 8);
  auto kern = k_dense<TANH, NTH, SPT, PF>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Sm));
  kern<<<148, NTH, sizeof(Sm)>>>(d, 50);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<long long> h(148);
  cudaMemcpy(h.data(), d, 148 * 8, cudaMemcpyDeviceToHost);
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, kern);
  printf("%-44s %6lld cyc/layer (ideal 4096 FFMA-issue)  regs %d  %s\n", name, h[0], fa.numRegs, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<false, 256, 8, 0>("256 thr acc[4][8] no tanh (current)");
  run<true, 256, 8, 0>("256 thr acc[4][8] tanh (current)");
  run<false, 256, 8, 1>("256 thr acc[4][8] no tanh, reg prefetch");
  run<true, 256, 8, 1>("256 thr acc[4][8] tanh, reg prefetch");
  run<false, 512, 8, 0>("512 thr acc[2][8] no tanh");
  run<true, 512, 8, 0>("512 thr acc[2][8] tanh");
  run<false, 512, 4, 0>("512 thr acc[4][4] no tanh");
  run<true, 512, 4, 0>("512 thr acc[4][4] tanh");
  run<false, 512, 8, 1>("512 thr acc[2][8] no tanh, reg prefetch");
  run<true, 512, 4, 1>("512 thr acc[4][4] tanh, reg prefetch");
  run<false, 1024, 4, 0>("1024 thr acc[2][4] no tanh");
  run<true, 1024, 4, 0>("1024 thr acc[2][4] tanh");
  return 0;
}
