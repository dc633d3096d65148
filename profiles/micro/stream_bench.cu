// Micro-benchmark behind the design of the solver loops' vector kernels (solvers.cu): the CG update's first half,
//   x += alpha p;  r -= alpha Ap;  sum r.r        (4 arrays read, 2 written, 48 n bytes)
// on n = 2^24 doubles (C4's vectors), written several ways.  Not part of the product; built and run by
// profiles/gpu_session_d.sh:  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o stream_bench stream_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

// (a) round-1 style: one wave of 2 CTAs per SM, 8 scalar items per thread and array, 256 apart
template <int ITEMS, bool RR>
__global__ void __launch_bounds__(256) k_scalar(int64_t n, double alpha, const double* __restrict__ p, const double* __restrict__ Ap,
                                                double* __restrict__ x, double* __restrict__ r, double* __restrict__ part) {
  constexpr int TILE = 256 * ITEMS;
  const int64_t chunk = ((n + gridDim.x - 1) / gridDim.x + 255) / 256 * 256;
  const int64_t lo = RR ? (int64_t)blockIdx.x * TILE : (int64_t)blockIdx.x * chunk;
  const int64_t hi = RR ? n : (lo + chunk < n ? lo + chunk : n);
  const int64_t step = RR ? (int64_t)gridDim.x * TILE : TILE;
  double acc = 0.0;
  for (int64_t tile = lo; tile < hi; tile += step) {
    double pv[ITEMS], av[ITEMS], xv[ITEMS], rv[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
      const int64_t k = tile + threadIdx.x + (int64_t)i * 256;
      const bool ok = k < hi;
      pv[i] = ok ? p[k] : 0.0; av[i] = ok ? Ap[k] : 0.0; xv[i] = ok ? x[k] : 0.0; rv[i] = ok ? r[k] : 0.0;
    }
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
      const int64_t k = tile + threadIdx.x + (int64_t)i * 256;
      if (k < hi) {
        x[k] = xv[i] + alpha * pv[i];
        const double v = rv[i] - alpha * av[i];
        r[k] = v;
        acc += v * v;
      }
    }
  }
  if (acc == 123.456) part[blockIdx.x] = acc;
}

// (b) 128-bit accesses: a thread owns PAIRS pairs per array and tile, 256 pairs apart
template <int PAIRS>
__global__ void __launch_bounds__(256) k_vec2(int64_t n, double alpha, const double2* __restrict__ p, const double2* __restrict__ Ap,
                                              double2* __restrict__ x, double2* __restrict__ r, double* __restrict__ part) {
  constexpr int TILE = 256 * PAIRS;  // in pairs
  const int64_t n2 = n / 2;
  double acc = 0.0;
  for (int64_t tile = (int64_t)blockIdx.x * TILE; tile < n2; tile += (int64_t)gridDim.x * TILE) {
    double2 pv[PAIRS], av[PAIRS], xv[PAIRS], rv[PAIRS];
#pragma unroll
    for (int i = 0; i < PAIRS; i++) {
      const int64_t k = tile + threadIdx.x + (int64_t)i * 256;
      const bool ok = k < n2;
      const double2 z = make_double2(0.0, 0.0);
      pv[i] = ok ? p[k] : z; av[i] = ok ? Ap[k] : z; xv[i] = ok ? x[k] : z; rv[i] = ok ? r[k] : z;
    }
#pragma unroll
    for (int i = 0; i < PAIRS; i++) {
      const int64_t k = tile + threadIdx.x + (int64_t)i * 256;
      if (k < n2) {
        x[k] = make_double2(xv[i].x + alpha * pv[i].x, xv[i].y + alpha * pv[i].y);
        const double2 v = make_double2(rv[i].x - alpha * av[i].x, rv[i].y - alpha * av[i].y);
        r[k] = v;
        acc += v.x * v.x + v.y * v.y;
      }
    }
  }
  if (acc == 123.456) part[blockIdx.x] = acc;
}

// (d) the fused CG update as the product runs it: phase 1 (x, r, r.r), grid barrier, phase 2 (p = r + beta p), one wave
template <int ITEMS, bool FENCE_ALL>
__global__ void __launch_bounds__(256, 2) k_fused(int64_t n, double alpha, double beta, double* p, const double* __restrict__ Ap,
                                                  double* __restrict__ x, double* r, double* __restrict__ part,
                                                  unsigned int* ticket, unsigned int* gen) {
  constexpr int TILE = 256 * ITEMS;
  __shared__ double red[8];
  __shared__ int s_last;
  const unsigned int gen0 = *reinterpret_cast<volatile unsigned int*>(gen);
  double acc = 0.0;
  for (int64_t tile = (int64_t)blockIdx.x * TILE; tile < n; tile += (int64_t)gridDim.x * TILE) {
    double pv[ITEMS], av[ITEMS], xv[ITEMS], rv[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
      const int64_t k = tile + threadIdx.x + (int64_t)i * 256;
      const bool ok = k < n;
      pv[i] = ok ? p[k] : 0.0; av[i] = ok ? Ap[k] : 0.0; xv[i] = ok ? x[k] : 0.0; rv[i] = ok ? r[k] : 0.0;
    }
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
      const int64_t k = tile + threadIdx.x + (int64_t)i * 256;
      if (k < n) {
        x[k] = xv[i] + alpha * pv[i];
        const double v = rv[i] - alpha * av[i];
        r[k] = v;
        acc += v * v;
      }
    }
  }
  for (int d = 16; d; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; w++) t += red[w];
    part[blockIdx.x] = t;
    __threadfence();
    s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) {
    if (threadIdx.x == 0) while (*reinterpret_cast<volatile unsigned int*>(gen) == gen0) {}
    __syncthreads();
    if (FENCE_ALL) __threadfence();
  } else {
    if (threadIdx.x == 0) { *ticket = 0; __threadfence(); atomicExch(gen, gen0 + 1u); }
    __syncthreads();
  }
  for (int64_t tile = (int64_t)blockIdx.x * TILE; tile < n; tile += (int64_t)gridDim.x * TILE) {
    double rv[ITEMS], pv[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
      const int64_t k = tile + threadIdx.x + (int64_t)i * 256;
      rv[i] = k < n ? r[k] : 0.0; pv[i] = k < n ? p[k] : 0.0;
    }
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
      const int64_t k = tile + threadIdx.x + (int64_t)i * 256;
      if (k < n) p[k] = rv[i] + beta * pv[i];
    }
  }
}

// (e) phase 2 alone
template <int ITEMS>
__global__ void __launch_bounds__(256, 2) k_phase2(int64_t n, double beta, double* __restrict__ p, const double* __restrict__ r) {
  constexpr int TILE = 256 * ITEMS;
  for (int64_t tile = (int64_t)blockIdx.x * TILE; tile < n; tile += (int64_t)gridDim.x * TILE) {
    double rv[ITEMS], pv[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
      const int64_t k = tile + threadIdx.x + (int64_t)i * 256;
      rv[i] = k < n ? r[k] : 0.0; pv[i] = k < n ? p[k] : 0.0;
    }
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
      const int64_t k = tile + threadIdx.x + (int64_t)i * 256;
      if (k < n) p[k] = rv[i] + beta * pv[i];
    }
  }
}

// (c) plain copy for reference: y = x with 128-bit accesses
__global__ void __launch_bounds__(256) k_copy(int64_t n2, const double2* __restrict__ a, double2* __restrict__ b) {
  for (int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x; k < n2; k += (int64_t)gridDim.x * 256) b[k] = a[k];
}

template <typename F>
static void run(const char* name, double bytes, F launch) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 5; i++) launch();
  CK(cudaEventRecord(e0));
  const int reps = 50;
  for (int i = 0; i < reps; i++) launch();
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  CK(cudaGetLastError());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  printf("%-44s %8.1f us  %7.1f GB/s\n", name, 1e3 * ms / reps, bytes * reps / (ms * 1e-3) / 1e9);
}

int main() {
  const int64_t n = 1 << 24;
  double *p, *Ap, *x, *r, *part;
  CK(cudaMalloc(&p, 8 * n)); CK(cudaMalloc(&Ap, 8 * n)); CK(cudaMalloc(&x, 8 * n)); CK(cudaMalloc(&r, 8 * n)); CK(cudaMalloc(&part, 8 * 65536));
  CK(cudaMemset(p, 0, 8 * n)); CK(cudaMemset(Ap, 0, 8 * n)); CK(cudaMemset(x, 0, 8 * n)); CK(cudaMemset(r, 0, 8 * n));
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const double bytes = 48.0 * n;
  printf("n = %lld doubles, %d SMs, 48 n = %.0f MB per launch\n", (long long)n, sms, bytes / 1e6);
  run("copy 128-bit, grid = n/512 (2 x 134 MB)", 16.0 * n, [&] { k_copy<<<(unsigned)(n / 2 / 256), 256>>>(n / 2, (const double2*)p, (double2*)x); });
  run("copy 128-bit, grid = 8 x SMs", 16.0 * n, [&] { k_copy<<<8 * sms, 256>>>(n / 2, (const double2*)p, (double2*)x); });
  for (int c : {2, 4, 8}) {
    char nm[128];
    snprintf(nm, sizeof nm, "scalar 8 items, chunk per CTA, %d CTAs/SM", c);
    run(nm, bytes, [&] { k_scalar<8, false><<<c * sms, 256>>>(n, 0.5, p, Ap, x, r, part); });
    snprintf(nm, sizeof nm, "scalar 8 items, round robin, %d CTAs/SM", c);
    run(nm, bytes, [&] { k_scalar<8, true><<<c * sms, 256>>>(n, 0.5, p, Ap, x, r, part); });
    snprintf(nm, sizeof nm, "scalar 4 items, round robin, %d CTAs/SM", c);
    run(nm, bytes, [&] { k_scalar<4, true><<<c * sms, 256>>>(n, 0.5, p, Ap, x, r, part); });
    snprintf(nm, sizeof nm, "scalar 2 items, round robin, %d CTAs/SM", c);
    run(nm, bytes, [&] { k_scalar<2, true><<<c * sms, 256>>>(n, 0.5, p, Ap, x, r, part); });
    snprintf(nm, sizeof nm, "128-bit 4 pairs, round robin, %d CTAs/SM", c);
    run(nm, bytes, [&] { k_vec2<4><<<c * sms, 256>>>(n, 0.5, (const double2*)p, (const double2*)Ap, (double2*)x, (double2*)r, part); });
    snprintf(nm, sizeof nm, "128-bit 2 pairs, round robin, %d CTAs/SM", c);
    run(nm, bytes, [&] { k_vec2<2><<<c * sms, 256>>>(n, 0.5, (const double2*)p, (const double2*)Ap, (double2*)x, (double2*)r, part); });
    snprintf(nm, sizeof nm, "128-bit 1 pair, round robin, %d CTAs/SM", c);
    run(nm, bytes, [&] { k_vec2<1><<<c * sms, 256>>>(n, 0.5, (const double2*)p, (const double2*)Ap, (double2*)x, (double2*)r, part); });
  }
  {
    unsigned int* cnt;
    CK(cudaMalloc(&cnt, 8)); CK(cudaMemset(cnt, 0, 8));
    const double b72 = 72.0 * n;
    run("FUSED update (72 n B), barrier, fence by all, 2 CTAs/SM", b72, [&] { k_fused<8, true><<<2 * sms, 256>>>(n, 0.5, 0.25, p, Ap, x, r, part, cnt, cnt + 1); });
    run("FUSED update (72 n B), barrier, no fence, 2 CTAs/SM", b72, [&] { k_fused<8, false><<<2 * sms, 256>>>(n, 0.5, 0.25, p, Ap, x, r, part, cnt, cnt + 1); });
    run("two kernels: phase 1 + phase 2 (72 n B), 2 CTAs/SM", b72, [&] {
      k_scalar<8, true><<<2 * sms, 256>>>(n, 0.5, p, Ap, x, r, part);
      k_phase2<8><<<2 * sms, 256>>>(n, 0.25, p, r); });
    run("phase 2 alone (24 n B), 2 CTAs/SM", 24.0 * n, [&] { k_phase2<8><<<2 * sms, 256>>>(n, 0.25, p, r); });
    for (size_t lim : {(size_t)0, (size_t)32 << 20, (size_t)96 << 20}) {
      // the same with a persisting-L2 set-aside claimed (what cask_b200_cg_device did once per context in round 1)
      int maxp = 0;
      CK(cudaDeviceGetAttribute(&maxp, cudaDevAttrMaxPersistingL2CacheSize, 0));
      const size_t want = lim < (size_t)maxp ? lim : (size_t)maxp;
      CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want));
      char nm[128];
      snprintf(nm, sizeof nm, "FUSED update, persisting L2 set-aside %zu MB (max %d MB)", want >> 20, maxp >> 20);
      run(nm, b72, [&] { k_fused<8, true><<<2 * sms, 256>>>(n, 0.5, 0.25, p, Ap, x, r, part, cnt, cnt + 1); });
      snprintf(nm, sizeof nm, "scalar 8 items RR 2 CTAs/SM, set-aside %zu MB", want >> 20);
      run(nm, bytes, [&] { k_scalar<8, true><<<2 * sms, 256>>>(n, 0.5, p, Ap, x, r, part); });
    }
    CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0));
  }
  run("128-bit 1 pair, grid = n/512 CTAs", bytes, [&] { k_vec2<1><<<(unsigned)(n / 512), 256>>>(n, 0.5, (const double2*)p, (const double2*)Ap, (double2*)x, (double2*)r, part); });
  run("128-bit 2 pairs, grid = n/1024 CTAs", bytes, [&] { k_vec2<2><<<(unsigned)(n / 1024), 256>>>(n, 0.5, (const double2*)p, (const double2*)Ap, (double2*)x, (double2*)r, part); });
  run("scalar 2 items, grid = n/512 CTAs", bytes, [&] { k_scalar<2, true><<<(unsigned)(n / 512), 256>>>(n, 0.5, p, Ap, x, r, part); });
  return 0;
}
