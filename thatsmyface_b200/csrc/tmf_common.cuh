// Shared by every translation unit of libtmfwm: error reporting, the block geometry of a
// batch, block-row loads/stores, and the launcher entry points each TU exports to api.cu.
//
// Mapping, every fused kernel: ONE THREAD OWNS ONE BLOCK; adjacent threads own adjacent
// blocks of one block-row, so a warp's accesses to image row r cover 32*24 = 768
// contiguous bytes.  All block arithmetic is in the owning thread's registers with
// compile-time indices: no shuffles, no exchange between threads, no redundant work.
//
//   fast_kernels.cu      FAST mode, block size 8: k_embed_tile (TMA-tiled, persistent),
//                        k_embed_fast / k_extract_fast / k_sigma0_fast (per-thread global
//                        accesses; any alignment)
//   fast_n_kernels.cu    FAST mode, the UI's other block sizes (4..16)
//   faithful_kernels.cu  FAITHFUL mode (DCT -> one-sided Jacobi -> IDCT), every block size
//   taps.cu              svd / dct / colour / pixel-format taps, strip round trip
//   wm_map.cu            watermark map on the device (PIL LANCZOS restated)
//   ctx.cu               host-buffer pipeline (H2D / kernel / D2H)
//   api.cu               C ABI of embed / extract / sigma0: argument checks + dispatch
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/tmf_wm.h"
#include "tmf_tunables.h"

namespace tmfi {

// ---- errors (api.cu) ------------------------------------------------------------------
int fail(int code, const char* fmt, ...);          // formats into the thread-local message, returns code
int check_launch(const char* what);                // cudaGetLastError -> TMF_ERR_CUDA
int cuda_fail(const char* what, cudaError_t e);    // clears the sticky error, formats, returns TMF_ERR_CUDA
const char* last_error_message();

#define TMF_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return ::tmfi::cuda_fail(#call, e_); } while (0)

constexpr int kThreads = 128;

// n / d for n < 2^31 as one 32x32->64 multiply and a shift: mul = ceil(2^shift / d),
// shift = 31 + ceil(log2 d) (exact: the error term n*e/(d*2^shift) stays below 1/d).
struct FastDiv {
  uint32_t mul, shift;
};
inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  uint32_t s = 0;
  while ((1ull << s) < d) ++s;
  f.shift = 31 + s;
  f.mul = (uint32_t)(((1ull << f.shift) + d - 1) / d);
  return f;
}

struct BlockGeom {
  int h, w, nbh, nbw;
  int bs;                     // block size
  long long blocks_per_img;   // nbh * nbw
  long long total_blocks;     // n * blocks_per_img  (< 2^31, make_geom)
  size_t img_stride;          // bytes between images
  size_t row_pitch;           // 3 * w
  uint32_t pitch32;           // the same in 32 bits (make_geom: 3 * w < 2^32)
  FastDiv div_bpi, div_nbw;   // block index -> (image, block row, block column) without divisions
};

// ---- host helpers (api.cu) -------------------------------------------------------------
int make_geom(int n, int h, int w, size_t img_stride, int block, BlockGeom& g);
inline unsigned grid_for(long long items, int per_cta) { return (unsigned)((items + per_cta - 1) / per_cta); }
// widest access (8, 4 or 1 bytes) the 24-byte block rows of a block-size-8 batch allow
int pick_vec(const BlockGeom& g, const void* p0, const void* p1, const void* p2 = nullptr);
// 4, 2 or 1: alignment shared by every block row of a batch (other block sizes)
int row_alignment(const BlockGeom& g, const void* p0, const void* p1, const void* p2 = nullptr);
int sm_count();   // multiprocessors of the current device (cached per device)

// ---- launchers, one per fused kernel family (each returns TMF_OK or an error) -------------
// fast_kernels.cu (block size 8)
int launch_embed_fast8(const uint8_t* rgb, uint8_t* out, const BlockGeom& g, const uint8_t* wm, int wm_shared,
                       double alpha, cudaStream_t st);
int launch_extract_fast8(const uint8_t* wmk, const uint8_t* orig, uint8_t* out_wm, const BlockGeom& g, double alpha,
                         cudaStream_t st);
int launch_sigma0_fast8(const uint8_t* rgb, float* sigma0, const BlockGeom& g, cudaStream_t st);
// which kernel the last FAST block-8 embed / extract on this thread used: 1 = TMA-tiled, 0 = per-thread
int last_fast_path();
// fast_n_kernels.cu (block sizes 4, 6, 10, 12, 14, 16)
int launch_embed_fast_n(const uint8_t* rgb, uint8_t* out, const BlockGeom& g, const uint8_t* wm, int wm_shared,
                        double alpha, cudaStream_t st);
int launch_extract_fast_n(const uint8_t* wmk, const uint8_t* orig, uint8_t* out_wm, const BlockGeom& g, double alpha,
                          cudaStream_t st);
int launch_sigma0_fast_n(const uint8_t* rgb, float* sigma0, const BlockGeom& g, cudaStream_t st);
// faithful_kernels.cu (every supported block size)
int launch_embed_faithful(const uint8_t* rgb, uint8_t* out, const BlockGeom& g, const uint8_t* wm, int wm_shared,
                          double alpha, int literal, cudaStream_t st);
int launch_extract_faithful(const uint8_t* wmk, const uint8_t* orig, uint8_t* out_wm, const BlockGeom& g,
                            double alpha, cudaStream_t st);
int launch_sigma0_faithful(const uint8_t* rgb, float* sigma0, const BlockGeom& g, cudaStream_t st);
// taps.cu: pixels outside whole blocks take the colour round trip only
int launch_strip_roundtrip(const uint8_t* rgb, uint8_t* out, const BlockGeom& g, int n, cudaStream_t st);

#if defined(__CUDACC__)
// ---------------------------------------------------------------------------
// device side
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t fastdiv(uint32_t n, FastDiv f) {
  return (uint32_t)(((unsigned long long)n * f.mul) >> f.shift);
}

// block index -> byte offset of the block's first pixel; N = block size
template <int N>
__device__ __forceinline__ size_t block_origin(const BlockGeom& g, long long gb, long long& img, int& by, int& bx,
                                               uint32_t* in_img = nullptr) {
  const uint32_t n = (uint32_t)gb;                       // total_blocks < 2^31 (make_geom)
  const uint32_t im = fastdiv(n, g.div_bpi);
  const uint32_t r = n - im * (uint32_t)g.blocks_per_img;
  const uint32_t y = fastdiv(r, g.div_nbw);
  img = im;
  by = (int)y;
  bx = (int)(r - y * (uint32_t)g.nbw);
  if (in_img) *in_img = r;                               // block index inside its image (the shared map's index)
  // one 32x32->64 multiply per term (IMAD.WIDE.U32): no 64x64 products
  return (size_t)im * g.img_stride + (unsigned long long)(y * (uint32_t)N) * g.pitch32 + (uint32_t)bx * (3u * N);
}

// 24-byte block-row load/store with the widest access the alignment allows
// (VEC 8 / 4 / 1 = global memory; VEC 0 = a row staged in shared memory, three LDS.64 / STS.64)
template <int VEC>
__device__ __forceinline__ void load_row24(const uint8_t* __restrict__ p, uint32_t (&w)[6]) {
  if (VEC == 0) {
    const uint2* q = reinterpret_cast<const uint2*>(p);
    uint2 a = q[0], b = q[1], c = q[2];
    w[0] = a.x; w[1] = a.y; w[2] = b.x; w[3] = b.y; w[4] = c.x; w[5] = c.y;
  } else if (VEC == 8) {
    const uint2* q = reinterpret_cast<const uint2*>(p);
    uint2 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
    w[0] = a.x; w[1] = a.y; w[2] = b.x; w[3] = b.y; w[4] = c.x; w[5] = c.y;
  } else if (VEC == 4) {
    const uint32_t* q = reinterpret_cast<const uint32_t*>(p);
#pragma unroll
    for (int k = 0; k < 6; ++k) w[k] = __ldg(q + k);
  } else {
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      w[k] = (uint32_t)__ldg(p + 4 * k) | ((uint32_t)__ldg(p + 4 * k + 1) << 8) |
             ((uint32_t)__ldg(p + 4 * k + 2) << 16) | ((uint32_t)__ldg(p + 4 * k + 3) << 24);
    }
  }
}

template <int VEC>
__device__ __forceinline__ void store_row24(uint8_t* __restrict__ p, const uint32_t (&w)[6]) {
  if (VEC == 8 || VEC == 0) {
    uint2* q = reinterpret_cast<uint2*>(p);
    q[0] = make_uint2(w[0], w[1]); q[1] = make_uint2(w[2], w[3]); q[2] = make_uint2(w[4], w[5]);
  } else if (VEC == 4) {
    uint32_t* q = reinterpret_cast<uint32_t*>(p);
#pragma unroll
    for (int k = 0; k < 6; ++k) q[k] = w[k];
  } else {
#pragma unroll
    for (int k = 0; k < 24; ++k) p[k] = (uint8_t)(w[k >> 2] >> (8 * (k & 3)));
  }
}

// ---- block rows of any even block size N: 3N bytes in ceil(3N/4) words ----
template <int N> constexpr int kRowWords = (3 * N + 3) / 4;

// Load the 3N bytes of a block row into words.  AL = alignment every block row of the batch
// shares: 4 (N = 4, 8, 12, 16 with 4-byte aligned rows), 2 (any even N with 2-byte aligned rows),
// else single bytes.  For 3N % 4 == 2 the last word holds two bytes (upper half zero).
template <int N, int AL>
__device__ __forceinline__ void load_row_n(const uint8_t* __restrict__ p, uint32_t (&w)[kRowWords<N>]) {
  constexpr int NW = kRowWords<N>;
  if (AL == 4) {
#pragma unroll
    for (int k = 0; k < (3 * N) / 4; ++k) w[k] = __ldg(reinterpret_cast<const uint32_t*>(p) + k);
    if ((3 * N) % 4) w[NW - 1] = (uint32_t)__ldg(reinterpret_cast<const uint16_t*>(p) + (3 * N) / 2 - 1);
  } else if (AL == 2) {
    const uint16_t* q = reinterpret_cast<const uint16_t*>(p);
#pragma unroll
    for (int k = 0; k < NW; ++k) {
      uint32_t lo = (uint32_t)__ldg(q + 2 * k), hi = 0;
      if (2 * k + 1 < (3 * N) / 2) hi = (uint32_t)__ldg(q + 2 * k + 1);
      w[k] = lo | (hi << 16);
    }
  } else {
#pragma unroll
    for (int k = 0; k < NW; ++k) {
      uint32_t v = 0;
#pragma unroll
      for (int b = 0; b < 4; ++b)
        if (4 * k + b < 3 * N) v |= (uint32_t)__ldg(p + 4 * k + b) << (8 * b);
      w[k] = v;
    }
  }
}

template <int N, int AL>
__device__ __forceinline__ void store_row_n(uint8_t* __restrict__ p, const uint32_t (&w)[kRowWords<N>]) {
  constexpr int NW = kRowWords<N>;
  if (AL == 4) {
#pragma unroll
    for (int k = 0; k < (3 * N) / 4; ++k) reinterpret_cast<uint32_t*>(p)[k] = w[k];
    if ((3 * N) % 4) reinterpret_cast<uint16_t*>(p)[(3 * N) / 2 - 1] = (uint16_t)w[NW - 1];
  } else if (AL == 2) {
    uint16_t* q = reinterpret_cast<uint16_t*>(p);
#pragma unroll
    for (int k = 0; k < (3 * N) / 2; ++k) q[k] = (uint16_t)(w[k >> 1] >> (16 * (k & 1)));
  } else {
#pragma unroll
    for (int k = 0; k < 3 * N; ++k) p[k] = (uint8_t)(w[k >> 2] >> (8 * (k & 3)));
  }
}

// byte B (0..23) of a 24-byte row held in six words
#define TMF_BYTE(w, B) (((w)[(B) >> 2] >> (8 * ((B)&3))) & 0xffu)

// Ask for all rows of a block up front.  The row loops are rolled (small code), so without
// this each warp would have only one row (3 loads) in flight.
template <int ROWS = 8>
__device__ __forceinline__ void prefetch_block_rows(const uint8_t* __restrict__ base, uint32_t pitch) {
#pragma unroll
  for (int i = 0; i < ROWS; ++i) { asm volatile("prefetch.global.L1 [%0];" ::"l"(base)); base += pitch; }
}
#endif  // __CUDACC__

}  // namespace tmfi
