// For the record (north_star: "batched per-block SVD as one-sided Jacobi with one warp per block, using
// shuffles"): a COOPERATIVE one-sided Jacobi - 8 lanes per 8x8 block, one column per lane, columns
// exchanged with shuffles, four blocks per warp - against the library's mapping, one THREAD per block
// with the whole matrix in its registers (tmf::jacobi_svd8<false>, csrc/tmf_math.cuh).  Values only
// (what the fused extract / faithful embed need), same rotation formulas, same stop rule.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I thatsmyface_b200/csrc \
//        profiles/microbench/coop_jacobi.cu -o profiles/microbench/coop_jacobi && profiles/microbench/coop_jacobi
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "tmf_math.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

// ---- one thread per block (the library's mapping; staged through shared memory like k_svd8x8) ----
__global__ void __launch_bounds__(128) k_thread_per_block(const float* __restrict__ blocks, long long n, float* __restrict__ S) {
  __shared__ float sm[128 * 65];
  const long long b0 = (long long)blockIdx.x * 128;
  const int nb = (int)min((long long)128, n - b0), t = threadIdx.x;
  for (int idx = t; idx < nb * 64; idx += 128) sm[(idx >> 6) * 65 + (idx & 63)] = __ldg(blocks + b0 * 64 + idx);
  __syncthreads();
  float s[8];
  if (t < nb) {
    float a[64], unscale;
#pragma unroll
    for (int k = 0; k < 64; ++k) a[k] = sm[t * 65 + k];
    tmf::jacobi_svd8<false>(a, nullptr, unscale);
    float n2[8];
    tmf::column_norms2(a, n2);
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = tmf::f_sqrt(n2[j]) * unscale;
  }
  __syncthreads();
  if (t < nb) {
#pragma unroll
    for (int j = 0; j < 8; ++j) sm[t * 9 + j] = s[j];
  }
  __syncthreads();
  for (int idx = t; idx < nb * 8; idx += 128) S[b0 * 8 + idx] = sm[(idx >> 3) * 9 + (idx & 7)];
}

// ---- cooperative: lane j of an 8-lane group holds column j; partner in round m is lane j ^ m ----
__global__ void __launch_bounds__(128) k_cooperative(const float* __restrict__ blocks, long long n, float* __restrict__ S) {
  const long long blk = ((long long)blockIdx.x * 128 + threadIdx.x) >> 3;
  const int j = threadIdx.x & 7;
  const bool live = blk < n;
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = live ? __ldg(blocks + blk * 64 + 8 * i + j) : 0.0f;   // column j (coalesced over the group's 8 lanes per row)
  // power-of-two pre-scaling as tmf::jacobi_svd8
  float own = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) own = fmaf(x[i], x[i], own);
  float frob2 = own;
#pragma unroll
  for (int m = 1; m < 8; m <<= 1) frob2 += __shfl_xor_sync(0xffffffffu, frob2, m);
  float unscale = 1.0f;
  const bool ok = (frob2 > 0.0f) && (frob2 < INFINITY);
  if (ok) {
    const float sc = tmf::pow2_scale_for(frob2, unscale);
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] *= sc;
    own *= sc * sc;
  }
  bool more = ok;
  for (int sweep = 0; sweep < TMF_JACOBI_MAX_SWEEPS; ++sweep) {
    if (!__any_sync(0xffffffffu, more)) break;      // the four blocks of a warp move together (shuffles need every lane)
    float worst = 0.0f;
    own = 0.f;                                       // refresh the carried norm once per sweep
#pragma unroll
    for (int i = 0; i < 8; ++i) own = fmaf(x[i], x[i], own);
#pragma unroll 1
    for (int m = 1; m < 8; ++m) {                    // 7 rounds: the XOR tournament meets all 28 pairs once
      float y[8], ga = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) { y[i] = __shfl_xor_sync(0xffffffffu, x[i], m); ga = fmaf(x[i], y[i], ga); }
      const float other = __shfl_xor_sync(0xffffffffu, own, m);
      const bool lower = (j & m) == 0;               // the lower lane of the pair plays "column p"
      const float al = lower ? own : other, be = lower ? other : own;
      float c, s, t;
      worst = fmaxf(worst, tmf::jacobi_cs(al, be, ga, c, s, &t));
      const float sg = lower ? -s : s;               // p' = c p - s q ;  q' = s p + c q
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = fmaf(c, x[i], sg * y[i]);
      own = lower ? fmaxf(own - t * ga, 0.0f) : own + t * ga;
    }
    // the pair's two lanes saw the same `worst` contributions only for their own pairs: reduce over the group
#pragma unroll
    for (int m = 1; m < 8; m <<= 1) worst = fmaxf(worst, __shfl_xor_sync(0xffffffffu, worst, m));
    more = more && (worst > TMF_JACOBI_MORE);
  }
  own = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) own = fmaf(x[i], x[i], own);
  if (live) S[blk * 8 + j] = tmf::f_sqrt(own) * unscale;   // unsorted, like the thread-per-block kernel here
}

static double time_ms(void (*launch)(const float*, long long, float*), const float* d_in, long long n, float* d_out) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int k = 0; k < 3; ++k) launch(d_in, n, d_out);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  const int reps = 20;
  for (int k = 0; k < reps; ++k) launch(d_in, n, d_out);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / reps;
}
static void launch_tpb(const float* in, long long n, float* out) { k_thread_per_block<<<(unsigned)((n + 127) / 128), 128>>>(in, n, out); }
static void launch_coop(const float* in, long long n, float* out) { k_cooperative<<<(unsigned)((n * 8 + 127) / 128), 128>>>(in, n, out); }

int main() {
  const long long nmax = 1000000;
  std::vector<float> h((size_t)nmax * 64);
  unsigned s = 12345u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (float)(s >> 8) / 16777216.0f; };
  for (long long b = 0; b < nmax; ++b) {
    // "DCT of a natural block"-like: energy decaying away from (0,0); every third block plain noise
    const bool noise = (b % 3) == 2;
    for (int i = 0; i < 8; ++i)
      for (int j = 0; j < 8; ++j) h[(size_t)b * 64 + 8 * i + j] = (rnd() - 0.5f) * (noise ? 1.0f : 4.0f / (1.0f + 3.0f * (i + j) * (i + j)));
    if (!noise) h[(size_t)b * 64] += 3.5f;
  }
  float *d_in, *d_a, *d_b;
  CK(cudaMalloc(&d_in, (size_t)nmax * 64 * 4)); CK(cudaMalloc(&d_a, (size_t)nmax * 8 * 4)); CK(cudaMalloc(&d_b, (size_t)nmax * 8 * 4));
  CK(cudaMemcpy(d_in, h.data(), (size_t)nmax * 64 * 4, cudaMemcpyHostToDevice));
  printf("one-sided Jacobi, values only, 8x8 fp32 blocks: one thread per block (library mapping) vs 8 lanes per block (cooperative, shuffles)\n");
  for (long long n : {1024LL, 4096LL, 32400LL, 129600LL, 518400LL, 1000000LL}) {
    const double ta = time_ms(launch_tpb, d_in, n, d_a), tb = time_ms(launch_coop, d_in, n, d_b);
    std::vector<float> a((size_t)n * 8), b((size_t)n * 8);
    CK(cudaMemcpy(a.data(), d_a, (size_t)n * 8 * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(b.data(), d_b, (size_t)n * 8 * 4, cudaMemcpyDeviceToHost));
    double worst = 0;
    for (long long k = 0; k < n; ++k) {       // compare as sorted sets, relative to sigma0
      float u[8], v[8];
      for (int q = 0; q < 8; ++q) { u[q] = a[(size_t)k * 8 + q]; v[q] = b[(size_t)k * 8 + q]; }
      for (int p = 0; p < 8; ++p) for (int q = p + 1; q < 8; ++q) { if (u[q] > u[p]) { float t = u[p]; u[p] = u[q]; u[q] = t; } if (v[q] > v[p]) { float t = v[p]; v[p] = v[q]; v[q] = t; } }
      for (int q = 0; q < 8; ++q) { const double d = fabs((double)u[q] - v[q]) / fmax(u[0], 1e-30f); if (d > worst) worst = d; }
    }
    printf("blocks %8lld: thread-per-block %8.4f ms = %7.3f G blocks/s | cooperative %8.4f ms = %7.3f G blocks/s | ratio %.2f | max |dsigma|/sigma0 %.1e\n",
           n, ta, n / ta / 1e6, tb, n / tb / 1e6, ta / tb, worst);
  }
  return 0;
}
