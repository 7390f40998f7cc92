// micro-benchmarks that calibrate the latency model used for the PCG kernel design (run under gpurun)
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
__global__ void k_lat(double* out, long long* cyc, int n) {
  __shared__ double sh[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sh[i] = 1.0 + i * 1e-9;
  __syncthreads();
  double a = out[0], b = 1.0000001, c = 1e-9;
  long long t0, t1;
  // DFMA chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < n; ++i) a = fma(a, b, c);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  // DADD chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < n; ++i) a = a + c;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[1] = t1 - t0;
  // double shuffle chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < n; ++i) a = __shfl_xor_sync(0xffffffffu, a, 1) + c;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[2] = t1 - t0;
  // LDS dependent chain (index from loaded value)
  int idx = threadIdx.x & 1023;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < n; ++i) {
    double v = sh[idx];
    idx = (idx + (int)v) & 1023;
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[3] = t1 - t0;
  a += idx;
  // syncthreads
  t0 = clock64();
  for (int i = 0; i < n; ++i) __syncthreads();
  t1 = clock64();
  if (threadIdx.x == 0) cyc[4] = t1 - t0;
  // double division chain
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < n; ++i) a = b / (a + 2.0);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[5] = t1 - t0;
  // clock64 overhead
  t0 = clock64();
  long long acc = 0;
#pragma unroll 16
  for (int i = 0; i < n; ++i) acc += clock64();
  t1 = clock64();
  if (threadIdx.x == 0) cyc[6] = t1 - t0;
  // nanosleep(20)
  t0 = clock64();
  for (int i = 0; i < 64; ++i) __nanosleep(20);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[7] = (t1 - t0) * n / 64;
  // FFMA chain for reference
  float fa = (float)a, fb = 1.0001f, fc = 1e-6f;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < n; ++i) fa = fmaf(fa, fb, fc);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[8] = t1 - t0;
  out[threadIdx.x + blockIdx.x * blockDim.x] = a + (double)acc + fa;
}
// DFMA throughput: many independent chains, all warps
__global__ void k_thr(double* out, long long* cyc, int n) {
  double a[8];
  for (int k = 0; k < 8; ++k) a[k] = out[k] + threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = fma(a[k], 1.0000001, 1e-9);
  }
  long long t1 = clock64();
  double s = 0;
  for (int k = 0; k < 8; ++k) s += a[k];
  out[threadIdx.x + blockIdx.x * blockDim.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
// global memory: relaxed.gpu load round trip (pointer chase), and cross-SM ping-pong latency
__global__ void k_gl(unsigned* buf, long long* cyc, int n) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned idx = 0;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
      unsigned v;
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(buf + idx) : "memory");
      idx = v;
    }
    long long t1 = clock64();
    cyc[0] = t1 - t0;
    buf[4096] = idx;
  }
}
__global__ void k_pingpong(unsigned* flags, long long* cyc, int n, int other) {
  // block 0 and block `other` bounce a counter
  unsigned* f = flags;
  if (threadIdx.x != 0) return;
  if (blockIdx.x == 0) {
    long long t0 = clock64();
    for (int i = 1; i <= n; ++i) {
      asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(f), "r"(2 * i - 1) : "memory");
      unsigned v;
      do {
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(f + 32) : "memory");
      } while (v != (unsigned)(2 * i));
    }
    long long t1 = clock64();
    cyc[0] = t1 - t0;
  } else if ((int)blockIdx.x == other) {
    for (int i = 1; i <= n; ++i) {
      unsigned v;
      do {
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
      } while (v != (unsigned)(2 * i - 1));
      asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(f + 32), "r"(2 * i) : "memory");
    }
  }
}
int main() {
  double* out;
  long long* cyc;
  unsigned* buf;
  cudaMalloc(&out, 1 << 22);
  cudaMemset(out, 0, 1 << 22);
  cudaMallocManaged(&cyc, 64 * 8);
  cudaMalloc(&buf, 1 << 20);
  const int n = 4096;
  const char* names[] = {"DFMA chain", "DADD chain", "shfl(double)+DADD chain", "LDS dependent chain", "__syncthreads (16 warps)",
                         "double div chain", "clock64 read", "nanosleep(20)", "FFMA chain"};
  for (int threads : {32, 512}) {
    k_lat<<<1, threads>>>(out, cyc, n);
    cudaDeviceSynchronize();
    printf("-- %d threads, 1 block: cycles per op\n", threads);
    for (int k = 0; k < 9; ++k) printf("  %-28s %.1f\n", names[k], (double)cyc[k] / n);
  }
  for (int threads : {128, 512, 1024}) {
    k_thr<<<148, threads>>>(out, cyc, n);
    cudaDeviceSynchronize();
    printf("DFMA throughput %d thr/SM: %.2f warp-DFMA/cycle/SM\n", threads, (double)n * 8 * (threads / 32) / cyc[0]);
  }
  {
    unsigned h[8192];
    for (int i = 0; i < 4096; ++i) h[i] = (i * 97 + 64) % 4096;
    cudaMemcpy(buf, h, 4096 * 4, cudaMemcpyHostToDevice);
    k_gl<<<1, 32>>>(buf, cyc, n);
    cudaDeviceSynchronize();
    printf("ld.relaxed.gpu dependent chain: %.1f cycles\n", (double)cyc[0] / n);
  }
  for (int other : {1, 2, 74, 147}) {
    cudaMemset(buf, 0, 4096);
    k_pingpong<<<148, 32>>>(buf, cyc, 2000, other);
    cudaDeviceSynchronize();
    printf("st->ld ping-pong block 0 <-> %d: %.1f cycles round trip (2 hops)\n", other, (double)cyc[0] / 2000);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
