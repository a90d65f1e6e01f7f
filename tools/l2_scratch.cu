// l2_scratch.cu -- does a write-then-read scratch that persistent warps REUSE stay in the 126 MB L2 while the
// step's own 32 B/node stream through it?  (Question behind a K7F-style 1-D table kernel: K2's (delta, lambda) round trip
// costs 32 of its 70 B/node.)  Each warp walks blocks of 32 lines x n nodes: forward reads grid / predict (evict_first)
// and writes (d, l) to ITS scratch block; backward reads the scratch in reverse and writes grid' / predict' (evict_first).
// Run under `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` for the DRAM bytes per node.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o l2_scratch l2_scratch.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long pol_first() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ unsigned long long pol_last() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ double ld_ef(const double *p, unsigned long long pol) {
  double v;
  asm volatile("ld.global.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ void st_ef(double *p, double v, unsigned long long pol) {
  asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol));
}
__device__ __forceinline__ void st_sc(double2 *p, double2 v, int hint, unsigned long long pol) {
  if (hint) asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(pol));
  else *p = v;
}
__device__ __forceinline__ double2 ld_sc(const double2 *p, int hint, unsigned long long pol) {
  double2 v;
  if (hint) asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
  else asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}

__global__ void __launch_bounds__(128) walk(double *G, double *P, double2 *S, int n, long long nblocks, int reuse, int hint) {
  const int lane = threadIdx.x & 31;
  const unsigned long long pf = pol_first(), pl = pol_last();
  const long long W = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  for (long long b = w; b < nblocks; b += W) {
    double *g = G + (size_t)b * n * 32 + lane, *p = P + (size_t)b * n * 32 + lane;
    double2 *s = S + (size_t)(reuse ? w : b) * n * 32 + lane;
    double d = 0.5, l = 0.25;
#pragma unroll 8
    for (int i = 0; i < n; ++i) {
      const double gv = ld_ef(g + i * 32, pf), pv = ld_ef(p + i * 32, pf);
      d = d * 0.5 + gv;
      l = l * 0.5 + pv;
      st_sc(s + i * 32, make_double2(d, l), hint, pl);
    }
    double x = d;
#pragma unroll 8
    for (int i = n - 1; i >= 0; --i) {
      const double2 v = ld_sc(s + i * 32, hint, pl);
      x = v.x * 1e-3 * x + v.y;
      st_ef(g + i * 32, x, pf);
      st_ef(p + i * 32, 2.0 * x - v.y, pf);
    }
  }
}

int main(int argc, char **argv) {
  const int n = 256;
  const long long nblocks = 32768;   // 1 M lines
  const size_t elems = (size_t)nblocks * n * 32;
  double *G, *P;
  double2 *S;
  cudaMalloc(&G, elems * 8);
  cudaMalloc(&P, elems * 8);
  cudaMalloc(&S, elems * 16);        // large enough for the no-reuse variant
  cudaMemset(G, 0, elems * 8);
  cudaMemset(P, 0, elems * 8);
  cudaMemset(S, 0, elems * 16);
  const int cfg[][3] = {{4, 1, 0}, {4, 1, 1}, {6, 1, 0}, {6, 1, 1}, {8, 1, 0}, {8, 1, 1}, {12, 1, 0}, {16, 1, 0}, {24, 1, 0}, {24, 0, 0}};
  for (auto &c : cfg) {
    const int wps = c[0], reuse = c[1], hint = c[2];
    const int ctas = 148 * wps / 4;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    walk<<<ctas, 128>>>(G, P, S, n, nblocks, reuse, hint);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    walk<<<ctas, 128>>>(G, P, S, n, nblocks, reuse, hint);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("warps/SM %2d reuse %d hint %d scratch %6.1f MB  %.3f ms  (%.0f GB/s of the 32 B/node stream)  %s\n", wps, reuse, hint,
           reuse ? 148.0 * wps * n * 32 * 16 / 1e6 : elems * 16 / 1e6, ms, elems * 32.0 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
