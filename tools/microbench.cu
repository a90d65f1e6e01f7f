// tools/microbench.cu -- (1) validates gs_div (shared-reciprocal division) against div.rn.f64 on random and
// adversarial operands; (2) measures FP64 pipe throughput / dependent latency and the cost of the division variants.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -fmad=false -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include "../gridsplines_b200/csrc/gs_device.cuh"

__device__ __forceinline__ uint64_t splitmix(uint64_t &s) {
  uint64_t z = (s += 0x9e3779b97f4a7c15ull);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}
// mode 0: mantissas uniform, exponents in [-40, 40]; 1: a slightly below b (quotient just under 1);
// 2: small integers / small integers scaled; 3: b with low mantissa bits only
__device__ double make_operand(uint64_t r, int mode, bool second, double first) {
  uint64_t mant = r & 0x000fffffffffffffull;
  int e = (int)((r >> 52) % 81) - 40;
  uint64_t sign = (r >> 63) << 63;
  double x = __longlong_as_double((long long)(sign | ((uint64_t)(1023 + e) << 52) | mant));
  if (mode == 1 && second) {
    x = first * (1.0 + (double)((r >> 20) & 0xffff) * 1.1102230246251565e-16);
  } else if (mode == 2) {
    x = (double)((r & 0xfffff) + 1) * ((r >> 40) & 1 ? 0.125 : 3.0);
  } else if (mode == 3) {
    x = __longlong_as_double((long long)(sign | ((uint64_t)(1023 + e) << 52) | (mant & 0xfffull) | (mant & 0x000ff00000000000ull)));
  }
  return x;
}
__global__ void k_check_div(uint64_t seed, int iters, int mode, unsigned long long *mism, unsigned long long *count) {
  uint64_t s = seed + (uint64_t)(blockIdx.x * blockDim.x + threadIdx.x) * 0x632be59bd9b4e019ull;
  unsigned long long bad = 0;
  for (int i = 0; i < iters; ++i) {
    double b = make_operand(splitmix(s), mode, false, 0.0);
    double a1 = make_operand(splitmix(s), mode, true, b);
    double a2 = make_operand(splitmix(s), mode, true, b);
    GsRcp r = gs_rcp(b);
    double q1 = gs_div(a1, r), q2 = gs_div(a2, r);
    double n1 = a1 / b, n2 = a2 / b;
    bad += (__double_as_longlong(q1) != __double_as_longlong(n1));
    bad += (__double_as_longlong(q2) != __double_as_longlong(n2));
    // the reciprocal itself
    bad += (__double_as_longlong(r.y) != __double_as_longlong(1.0 / b));
  }
  atomicAdd(mism, bad);
  atomicAdd(count, 3ull * iters);
}

// ---- throughput / latency -------------------------------------------------
template <int ILP>
__global__ void k_dfma(double *out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int j = 0; j < ILP; ++j) x[j] = threadIdx.x * 1e-3 + j;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < ILP; ++j) x[j] = fma(x[j], a, b);
  }
  double s = 0;
#pragma unroll
  for (int j = 0; j < ILP; ++j) s += x[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>  // 0 native a/b, 1 rcp + 1 quotient, 2 rcp + 2 quotients (Thomas pair), 3 invariant divisor
__global__ void k_divs(double *out, int iters, double a0, double b0) {
  double a = a0 + threadIdx.x * 1e-3, b = b0 + threadIdx.x * 1e-4, acc = 0.0, acc2 = 0.0;
  GsRcp inv = gs_rcp(b0);
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) { acc += a / b; acc2 += (a + 1.0) / b; }
    if (MODE == 1) { acc += gs_div(a, gs_rcp(b)); acc2 += gs_div(a + 1.0, gs_rcp(b + 1.0)); }
    if (MODE == 2) { GsRcp r = gs_rcp(b); acc += gs_div(a, r); acc2 += gs_div(a + 1.0, r); }
    if (MODE == 3) { acc += gs_div(a, inv); acc2 += gs_div(a + 1.0, inv); }
    a += 1e-3; b += 1e-3;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + acc2;
}
template <int MODE>  // dependent chain: x = c / (d + e * x)  (the Thomas delta recurrence)
__global__ void k_chain(double *out, int iters, double c, double d, double e) {
  double x = 0.1 + threadIdx.x * 1e-6;
  for (int i = 0; i < iters; ++i) {
    double den = d + e * x;
    x = MODE == 0 ? (c / den) : gs_div(c, gs_rcp(den));
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

template <typename F> float time_ms(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main(int argc, char **argv) {
  long long check_iters = argc > 1 ? atoll(argv[1]) : 2000;
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  int sms = prop.multiProcessorCount; double ghz = prop.clockRate * 1e-6;
  printf("device %s, %d SMs, %.3f GHz nominal\n", prop.name, sms, ghz);
  unsigned long long *d; cudaMalloc(&d, 16); 
  for (int mode = 0; mode < 4; ++mode) {
    cudaMemset(d, 0, 16);
    k_check_div<<<sms * 8, 256>>>(0x1234567ull + mode, (int)check_iters, mode, d, d + 1);
    unsigned long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("div check mode %d: %llu mismatches in %llu quotients/reciprocals  (%s)\n", mode, h[0], h[1], cudaGetErrorString(cudaGetLastError()));
  }
  double *out; cudaMalloc(&out, (size_t)sms * 32 * 1024 * 8);
  int iters = 4096;
  auto report = [&](const char *name, float ms, double ops_per_thread, int threads) {
    double ops = ops_per_thread * threads;
    printf("%-44s %8.3f ms  %8.2f Gop/s  %6.2f op/clk/SM (nominal clk)\n", name, ms, ops / ms * 1e-6, ops / (ms * 1e-3) / (sms * ghz * 1e9));
  };
  { int T = sms * 1024; float ms = time_ms([&] { k_dfma<8><<<sms * 4, 256>>>(out, iters, 1.0000001, 1e-9); }); report("DFMA ILP8, 1024 thr/SM (throughput)", ms, 8.0 * iters, T); }
  { int T = sms * 128; float ms = time_ms([&] { k_dfma<1><<<sms, 128>>>(out, iters, 1.0000001, 1e-9); });
    printf("%-44s %8.3f ms  -> dependent DFMA latency ~ %.1f clk (nominal clk)\n", "DFMA ILP1, 128 thr/SM (latency)", ms, ms * 1e-3 * ghz * 1e9 / iters); (void)T; }
  { int T = sms * 1024; 
    float m0 = time_ms([&] { k_divs<0><<<sms * 4, 256>>>(out, iters, 1.5, 2.5); }); report("native a/b (x2 per iter)", m0, 2.0 * iters, T);
    float m1 = time_ms([&] { k_divs<1><<<sms * 4, 256>>>(out, iters, 1.5, 2.5); }); report("gs_div(a, gs_rcp(b)) (x2 per iter)", m1, 2.0 * iters, T);
    float m2 = time_ms([&] { k_divs<2><<<sms * 4, 256>>>(out, iters, 1.5, 2.5); }); report("shared rcp, 2 quotients", m2, 2.0 * iters, T);
    float m3 = time_ms([&] { k_divs<3><<<sms * 4, 256>>>(out, iters, 1.5, 2.5); }); report("invariant divisor, 2 quotients", m3, 2.0 * iters, T); }
  { float m0 = time_ms([&] { k_chain<0><<<sms, 128>>>(out, iters, -0.3, 2.1, 0.4); });
    float m1 = time_ms([&] { k_chain<1><<<sms, 128>>>(out, iters, -0.3, 2.1, 0.4); });
    printf("Thomas delta chain, native div : %.1f clk/node (nominal clk)\n", m0 * 1e-3 * ghz * 1e9 / iters);
    printf("Thomas delta chain, gs_div     : %.1f clk/node (nominal clk)\n", m1 * 1e-3 * ghz * 1e9 / iters); }
  return 0;
}
