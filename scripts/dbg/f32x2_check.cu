#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
__device__ __forceinline__ uint64_t pack2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) { uint64_t r; asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float dist(float a, float b, float c, float d, float x, float y, float z) {
  return __fadd_rn(__fadd_rn(__fmul_rn(a, x), __fmul_rn(b, y)), __fadd_rn(__fmul_rn(c, z), d));
}
__global__ void k(const float4* h, const float4* p, int n, unsigned* bad, float* ex) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 h0 = h[2 * i], h1 = h[2 * i + 1], P = p[i];
  uint64_t ha = pack2(h0.x, h1.x), hb = pack2(h0.y, h1.y), hc = pack2(h0.z, h1.z), hd = pack2(h0.w, h1.w);
  uint64_t x2 = pack2(P.x, P.x), y2 = pack2(P.y, P.y), z2 = pack2(P.z, P.z);
  uint64_t s = add2(add2(mul2(ha, x2), mul2(hb, y2)), add2(mul2(hc, z2), hd));
  float s0, s1; unpack2(s, s0, s1);
  float r0 = dist(h0.x, h0.y, h0.z, h0.w, P.x, P.y, P.z), r1 = dist(h1.x, h1.y, h1.z, h1.w, P.x, P.y, P.z);
  bool b0 = __float_as_uint(s0) != __float_as_uint(r0) && !(isnan(s0) && isnan(r0));
  bool b1 = __float_as_uint(s1) != __float_as_uint(r1) && !(isnan(s1) && isnan(r1));
  if (b0 || b1) { if (atomicAdd(bad, 1) == 0) { ex[0]=b0? s0:s1; ex[1]=b0? r0:r1; ex[2]=b0?0:1; ex[3]=P.x; ex[4]=P.y; ex[5]=P.z; } }
}
int main() {
  const int n = 1 << 23;
  float4 *hh = (float4*)malloc(2*n*16), *hp = (float4*)malloc(n*16);
  srand(7);
  auto rf = [](float s){ return (rand() / (float)RAND_MAX - 0.5f) * 2.f * s; };
  for (int i = 0; i < 2*n; ++i) { float a=rf(1),b=rf(1),c=rf(1); float nn=sqrtf(a*a+b*b+c*c); hh[i] = make_float4(a/nn,b/nn,c/nn,rf(4)); }
  for (int i = 0; i < n; ++i) hp[i] = make_float4(rf(3), rf(3), rf(6), 0.f);
  float4 *dh, *dp; unsigned* db; float* dex;
  cudaMalloc(&dh, 2*n*16); cudaMalloc(&dp, n*16); cudaMalloc(&db, 4); cudaMalloc(&dex, 32);
  cudaMemcpy(dh, hh, 2*n*16, cudaMemcpyHostToDevice); cudaMemcpy(dp, hp, n*16, cudaMemcpyHostToDevice);
  cudaMemset(db, 0, 4); cudaMemset(dex, 0, 32);
  k<<<(n+255)/256, 256>>>(dh, dp, n, db, dex);
  unsigned b; float ex[8];
  cudaMemcpy(&b, db, 4, cudaMemcpyDeviceToHost); cudaMemcpy(ex, dex, 32, cudaMemcpyDeviceToHost);
  printf("expression mismatches %u of %d\n", b, n);
  printf("ex: packed %a scalar %a lane %g  P=(%a,%a,%a)\n", ex[0], ex[1], ex[2], ex[3], ex[4], ex[5]);
  return 0;
}
