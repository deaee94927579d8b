// Micro-benchmark: legacy mma.sync.m16n8k16 (bf16 -> fp32) issue rate per SM on B200, to size the
// per-(sequence, head) attention products that are too small for tcgen05 tiles.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__global__ void probe(float* out, int iters) {
  float c[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
  uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, 0x3f803f80u, 0x3f803f80u}, b0 = 0x3f803f80u, b1 = threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out;
  cudaMalloc(&out, 148 * 1024 * sizeof(float) * 4);
  int sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  for (int warps : {4, 8, 16, 32}) {
    const int iters = 4096;
    probe<<<sms, warps * 32>>>(out, 16);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<<<sms, warps * 32>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double mmas = (double)sms * warps * iters * 8;
    const double flops = mmas * 2.0 * 16 * 8 * 16;
    printf("warps/SM %2d: %.3f ms  %.1f TFLOP/s  %.1f MAC/clk/SM (at %d MHz)\n", warps, ms, flops / ms / 1e9,
           mmas * 2048 / sms / (ms * 1e-3 * khz * 1e3), khz / 1000);
  }
  return cudaGetLastError() != cudaSuccess;
}
