// libtmfwm: sm_100a kernels + C ABI for the DCT+SVD watermark path.
// See include/tmf_wm.h for the contract and DESIGN.md for the layout/rooflines.
//
// Mapping, every fused kernel: ONE THREAD OWNS ONE BLOCK; adjacent threads own adjacent
// blocks of one block-row, so a warp's accesses to image row r cover 32*24 = 768
// contiguous bytes.  All block arithmetic is in the owning thread's registers with
// compile-time indices: no shuffles, no exchange between threads, no redundant work.
//
//   FAST mode  (tmf_fast.cuh; k_embed_fast / k_extract_fast / k_sigma0_fast, and the
//              generic-N k_*_fast_n for the UI's other block sizes): two streaming row
//              passes around a certified power iteration on the 8x8 Gram matrix; packed
//              fp32 (FFMA2/FADD2/FMUL2); the block itself is never held.
//   FAITHFUL   (tmf_math.cuh; k_embed_faithful / k_extract_faithful / k_sigma0_faithful):
//              bit-exact colour, DCT, one-sided Jacobi SVD (A and V in registers, packed
//              rounds), U diag(S') V^T, IDCT.
//   taps       k_svd8x8, k_dct8x8, k_rgb2ycc, k_ycc2rgb, k_strip_roundtrip.
//   host       tmf_ctx_*: H2D / kernel / D2H pipeline over host buffers (end of file).
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <type_traits>

#include "../../include/tmf_wm.h"
#include "tmf_math.cuh"
#include "tmf_fast.cuh"
#include "tmf_resize.cuh"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return code;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(TMF_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return TMF_OK;
}

constexpr int kThreads = 128;

// tunables of the fast kernels (profiles/sweep_variants.py builds and times the alternatives)
#ifndef TMF_ROW_UNROLL
#define TMF_ROW_UNROLL 4      // rows per iteration of the rolled row loops (1, 2, 4, 8)
#endif
#ifndef TMF_L2_LOOKAHEAD
#define TMF_L2_LOOKAHEAD 0    // blocks ahead whose rows are prefetched into L2 (0 = off)
#endif
#ifndef TMF_FAST_MIN_CTAS
#define TMF_FAST_MIN_CTAS 6   // __launch_bounds__ minimum CTAs/SM of the fast extract / sigma0 kernels
#endif
#ifndef TMF_EMBED_MIN_CTAS
#define TMF_EMBED_MIN_CTAS 5  // ... of the fast embed kernel: 5 since the subnormal quantiser (96 registers, 160 KB of stash
                              // leave the L1 60 KB for the pass-2 re-reads; profiles/r01_sweep_variants.txt, tables 4-5, 14)
#endif
#ifndef TMF_STORE_HINT
#define TMF_STORE_HINT 0        // 8-byte row stores: 0 = plain, 1 = st.global.cs (streaming), 2 = st.global.wt
#endif
#ifndef TMF_EMBED_THREADS
#define TMF_EMBED_THREADS 32    // threads per CTA of the fast embed kernel (128, 64 or 32; same warps per SM): one-warp CTAs
                                // release their registers and stash as soon as their own warp ends (+1 %, table 16)
#endif
constexpr int kEmbedThreads = TMF_EMBED_THREADS;
#ifdef TMF_EMBED_MIN_CTAS_RAW
constexpr int kEmbedMinCtas = TMF_EMBED_MIN_CTAS_RAW;          // CTAs of kEmbedThreads threads per SM
#else
constexpr int kEmbedMinCtas = TMF_EMBED_MIN_CTAS * (128 / TMF_EMBED_THREADS);
#endif
#ifndef TMF_EXTRACT_THREADS
#define TMF_EXTRACT_THREADS 32  // threads per CTA of the fast extract kernel (+0.5 %)
#endif
constexpr int kExtractThreads = TMF_EXTRACT_THREADS;
#ifndef TMF_EMBED_PERSIST
#define TMF_EMBED_PERSIST 0     // > 0: persistent embed kernel with that many CTAs per SM and a next-block L2 prefetch
#endif
#ifndef TMF_EMBED_ROWPTR
#define TMF_EMBED_ROWPTR 0      // embed kernel's row addresses: 0 = base + i * pitch, 1 = running pointers
#endif
#ifndef TMF_EMBED_REPREFETCH
#define TMF_EMBED_REPREFETCH 0  // ask L1 for the block's rows again before (1) / after (2) the eigen-solve
#endif
#ifndef TMF_EMBED_STASH
#define TMF_EMBED_STASH 1     // 1: pass 1 parks the luma in shared memory for pass 2; 0: pass 2 recomputes it
#endif
constexpr int kRowUnroll = TMF_ROW_UNROLL;
#ifndef TMF_ROW_UNROLL_P2
#define TMF_ROW_UNROLL_P2 TMF_ROW_UNROLL   // the embed kernel's pass 2, separately
#endif
constexpr int kRowUnrollP2 = TMF_ROW_UNROLL_P2;

// ---------------------------------------------------------------------------
// 24-byte block-row load/store with the widest access the alignment allows
// ---------------------------------------------------------------------------
template <int VEC>
__device__ __forceinline__ void load_row24(const uint8_t* __restrict__ p, uint32_t (&w)[6]) {
  if (VEC == 0) {               // a row staged in shared memory (TMA tile): three LDS.64
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
  if (VEC == 8 && TMF_STORE_HINT == 1) {            // streaming stores: the output is never read again
    uint2* q = reinterpret_cast<uint2*>(p);
    __stcs(q, make_uint2(w[0], w[1])); __stcs(q + 1, make_uint2(w[2], w[3])); __stcs(q + 2, make_uint2(w[4], w[5]));
  } else if (VEC == 8 && TMF_STORE_HINT == 2) {
    uint2* q = reinterpret_cast<uint2*>(p);
    __stwt(q, make_uint2(w[0], w[1])); __stwt(q + 1, make_uint2(w[2], w[3])); __stwt(q + 2, make_uint2(w[4], w[5]));
  } else if (VEC == 8 || VEC == 0) {
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

// byte B (0..23) of a 24-byte row held in six words
#define TMF_BYTE(w, B) (((w)[(B) >> 2] >> (8 * ((B)&3))) & 0xffu)

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
__device__ __forceinline__ uint32_t fastdiv(uint32_t n, FastDiv f) {
  return (uint32_t)(((unsigned long long)n * f.mul) >> f.shift);
}

struct BlockGeom {
  int h, w, nbh, nbw;
  int bs;                     // block size (8 on the optimised path)
  long long blocks_per_img;   // nbh * nbw
  long long total_blocks;     // n * blocks_per_img  (< 2^31, make_geom)
  size_t img_stride;          // bytes between images
  size_t row_pitch;           // 3 * w
  uint32_t pitch32;           // the same in 32 bits (make_geom: 3 * w < 2^32): row addresses by 32-bit increments
  FastDiv div_bpi, div_nbw;   // block index -> (image, block row, block column) without divisions
};

__device__ __forceinline__ size_t block_origin(const BlockGeom& g, long long gb, long long& img, int& by, int& bx,
                                               uint32_t* in_img = nullptr) {
  const uint32_t n = (uint32_t)gb;
  const uint32_t im = fastdiv(n, g.div_bpi);
  const uint32_t r = n - im * (uint32_t)g.blocks_per_img;
  const uint32_t y = fastdiv(r, g.div_nbw);
  img = im;
  by = (int)y;
  bx = (int)(r - y * (uint32_t)g.nbw);
  if (in_img) *in_img = r;                            // block index inside its image (the shared map's index)
  // one 32x32->64 multiply per term (IMAD.WIDE.U32): no 64x64 products
  return (size_t)im * g.img_stride + (unsigned long long)(y * 8u) * g.pitch32 + (uint32_t)bx * 24u;
}

// Ask for all 8 rows of a block up front.  The row loops below are rolled (small
// code), so without this each warp would have only one row (3 loads) in flight.
#ifndef TMF_PREFETCH_FROM
#define TMF_PREFETCH_FROM 0   // first row asked for by the fast kernels' entry prefetch (the row loops load rows 0..3 at once anyway)
#endif
__device__ __forceinline__ void prefetch_block_rows(const uint8_t* __restrict__ base, uint32_t pitch, int from = 0) {
  base += (size_t)from * pitch;
#pragma unroll
  for (int i = from; i < 8; ++i) { asm volatile("prefetch.global.L1 [%0];" ::"l"(base)); base += pitch; }
}

// Faithful-mode block I/O.  The DCT / Jacobi need the whole block in registers
// with compile-time indices, but the per-pixel colour code is long (float64
// dots, exact division), so unrolling it over 64 pixels made the kernel 230 KB
// of instructions and it stalled on instruction fetch (profiles/r01_*).  The row
// loops are therefore rolled and exchange the block with the register file
// through a thread-private column of shared memory: sm[k * kThreads + tid]
// (conflict-free, no barrier needed).
template <int VEC>
__device__ __forceinline__ void luma_rows_to_smem(const uint8_t* __restrict__ base, size_t pitch, float* __restrict__ col) {
#pragma unroll 1
  for (int i = 0; i < 8; ++i) {
    uint32_t w[6];
    load_row24<VEC>(base + (size_t)i * pitch, w);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float r = tmf::unit_from_u8(TMF_BYTE(w, 3 * j));
      const float g = tmf::unit_from_u8(TMF_BYTE(w, 3 * j + 1));
      const float b = tmf::unit_from_u8(TMF_BYTE(w, 3 * j + 2));
      col[(8 * i + j) * kThreads] = tmf::luma_exact(r, g, b);
    }
  }
}

template <int VEC>
__device__ __forceinline__ void load_luma_block(const uint8_t* __restrict__ base, size_t pitch, float* __restrict__ col,
                                                float* a) {
  luma_rows_to_smem<VEC>(base, pitch, col);
#pragma unroll
  for (int k = 0; k < 64; ++k) a[k] = col[k * kThreads];
}

// ---------------------------------------------------------------------------
// fused embed, faithful mode (watermarking.py:163-219 in one launch)
// ---------------------------------------------------------------------------
#ifndef TMF_FAITHFUL_MIN_CTAS
#define TMF_FAITHFUL_MIN_CTAS 3
#endif
template <int VEC>
__global__ void __launch_bounds__(kThreads, TMF_FAITHFUL_MIN_CTAS)
k_embed_faithful(const uint8_t* __restrict__ rgb, uint8_t* __restrict__ out, BlockGeom g,
                 const uint8_t* __restrict__ wm, int wm_shared, double alpha) {
  __shared__ float sm[64 * kThreads];
  const long long gb = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (gb >= g.total_blocks) return;
  long long img; int by, bx;
  const size_t org = block_origin(g, gb, img, by, bx);
  const uint8_t* src = rgb + org;
  float* col = sm + threadIdx.x;
  prefetch_block_rows(src, g.pitch32);

  float a[64], v[64];
  load_luma_block<VEC>(src, g.row_pitch, col, a);
  const long long wi = (wm_shared ? 0 : img * g.blocks_per_img) + (long long)by * g.nbw + bx;
  tmf::embed_block_faithful(a, v, alpha, (uint32_t)__ldg(wm + wi), nullptr);
#pragma unroll
  for (int k = 0; k < 64; ++k) col[k * kThreads] = a[k];

  // colour out: chroma is recomputed from the (L1/L2-resident) input bytes
  uint8_t* dst = out + org;
#pragma unroll 1
  for (int i = 0; i < 8; ++i) {
    uint32_t w[6], o[6] = {0, 0, 0, 0, 0, 0};
    load_row24<VEC>(src + (size_t)i * g.row_pitch, w);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float r = tmf::unit_from_u8(TMF_BYTE(w, 3 * j));
      const float gg = tmf::unit_from_u8(TMF_BYTE(w, 3 * j + 1));
      const float b = tmf::unit_from_u8(TMF_BYTE(w, 3 * j + 2));
      float cb, cr;
      tmf::chroma_exact(r, gg, b, cb, cr);
      uint32_t R, G, B;
      tmf::ycc_to_rgb8_exact(col[(8 * i + j) * kThreads], cb, cr, R, G, B);
      o[(3 * j) >> 2] |= R << (8 * ((3 * j) & 3));
      o[(3 * j + 1) >> 2] |= G << (8 * ((3 * j + 1) & 3));
      o[(3 * j + 2) >> 2] |= B << (8 * ((3 * j + 2) & 3));
    }
    store_row24<VEC>(dst + (size_t)i * g.row_pitch, o);
  }
}

// pixels outside whole blocks: colour round trip only (watermarking.py:173-174, :216)
__global__ void __launch_bounds__(256)
k_strip_roundtrip(const uint8_t* __restrict__ rgb, uint8_t* __restrict__ out, BlockGeom g, int n,
                  long long strip_px_per_img) {
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  if (t >= strip_px_per_img * n) return;
  const long long img = t / strip_px_per_img;
  long long idx = t - img * strip_px_per_img;
  const int bw = g.nbw * g.bs, bh = g.nbh * g.bs, rw = g.w - bw;
  int y, x;
  if (idx < (long long)g.h * rw) {
    y = (int)(idx / rw);
    x = bw + (int)(idx - (long long)y * rw);
  } else {
    idx -= (long long)g.h * rw;
    y = bh + (int)(idx / bw);
    x = (int)(idx - (long long)(y - bh) * bw);
  }
  const size_t off = (size_t)img * g.img_stride + (size_t)y * g.row_pitch + (size_t)x * 3;
  const float r = tmf::unit_from_u8(__ldg(rgb + off));
  const float gg = tmf::unit_from_u8(__ldg(rgb + off + 1));
  const float b = tmf::unit_from_u8(__ldg(rgb + off + 2));
  float cb, cr;
  tmf::chroma_exact(r, gg, b, cb, cr);
  uint32_t R, G, B;
  tmf::ycc_to_rgb8_exact(tmf::luma_exact(r, gg, b), cb, cr, R, G, B);
  out[off] = (uint8_t)R; out[off + 1] = (uint8_t)G; out[off + 2] = (uint8_t)B;
}

// ---------------------------------------------------------------------------
// fused extract, faithful mode (watermarking.py:246-289 in one launch)
// ---------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(kThreads)
k_extract_faithful(const uint8_t* __restrict__ wmk, const uint8_t* __restrict__ orig, uint8_t* __restrict__ out_wm,
                   BlockGeom g, double alpha) {
  const long long gb = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (gb >= g.total_blocks) return;
  long long img; int by, bx;
  const size_t org = block_origin(g, gb, img, by, bx);
  __shared__ float sm[64 * kThreads];
  float* col = sm + threadIdx.x;
  prefetch_block_rows(wmk + org, g.pitch32);
  prefetch_block_rows(orig + org, g.pitch32);
  float sw = 0.0f, so = 0.0f;
#pragma unroll 1
  for (int which = 0; which < 2; ++which) {          // one copy of the code for both images
    float a[64];
    load_luma_block<VEC>((which == 0 ? wmk : orig) + org, g.row_pitch, col, a);
    const float sg = tmf::sigma0_block_faithful(a, nullptr);
    if (which == 0) sw = sg; else so = sg;
  }
  out_wm[gb] = (uint8_t)tmf::extract_level(sw, so, alpha);
}

template <int VEC>
__global__ void __launch_bounds__(kThreads)
k_sigma0_faithful(const uint8_t* __restrict__ rgb, float* __restrict__ sigma0, BlockGeom g) {
  const long long gb = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (gb >= g.total_blocks) return;
  long long img; int by, bx;
  const size_t org = block_origin(g, gb, img, by, bx);
  __shared__ float sm[64 * kThreads];
  float* col = sm + threadIdx.x;
  float a[64];
  load_luma_block<VEC>(rgb + org, g.row_pitch, col, a);
  sigma0[gb] = tmf::sigma0_block_faithful(a, nullptr);
}

// ---------------------------------------------------------------------------
// FAST mode kernels (tmf_fast.cuh): same one-thread-per-block mapping, spatial
// top-triplet instead of DCT + full SVD, fp32 colour in 0..255 units
// ---------------------------------------------------------------------------
// byte B of the 24-byte row as a float, via the 2^23 magic number (PRMT + FADD,
// both full-rate pipes; the I2F.U8 conversion pipe is much narrower)
// byte B of the row as the "magic" float 2^23 + k (the bit pattern 0x4B0000kk is exactly
// 8388608 + k).  Bytes 0-2 take one PRMT (ALU pipe, half rate); byte 3 takes one
// IMAD.HI - hi32(x * 256) + 0x4B000000 = (x >> 24) + magic - which runs on the FMA pipe:
// ncu shows the ALU pipe as the busier one (52 % vs 36 %, math_pipe_throttle stalls), so a
// quarter of the extractions is moved across.
#ifndef TMF_BYTE3_IMAD
#define TMF_BYTE3_IMAD 1
#endif
#ifndef TMF_EXTRACT_IMAD_MASK
#define TMF_EXTRACT_IMAD_MASK 0x8   // bit b set: byte b of a word is extracted on the FMA pipe (IMAD) instead of PRMT
#endif
// MAGIC = 0x4B000000 (2^23) or 0x4B400000 (1.5 * 2^23, the floor bias of the quantiser).
template <uint32_t MAGIC = 0x4B000000u>
__device__ __forceinline__ float byte_to_magic(const uint32_t (&w)[6], int B) {
  const uint32_t x = w[B >> 2];
  const int b = B & 3;
  uint32_t m;
  if ((TMF_EXTRACT_IMAD_MASK >> b) & 1) {
    if (b == 3) {
#if TMF_BYTE3_IMAD
      asm("mad.hi.u32 %0, %1, 256, %2;" : "=r"(m) : "r"(x), "n"(MAGIC));
#else
      m = __funnelshift_r(x, MAGIC >> 8, 24);         // (x >> 24) | MAGIC
#endif
    } else {
      uint32_t t;                                      // (x << (24 - 8b)) >> 24, both on the FMA pipe
      asm("mul.lo.u32 %0, %1, %2;" : "=r"(t) : "r"(x), "r"(1u << (24 - 8 * b)));
      asm("mad.hi.u32 %0, %1, 256, %2;" : "=r"(m) : "r"(t), "n"(MAGIC));
    }
  } else if (b == 3) {
    m = __funnelshift_r(x, MAGIC >> 8, 24);
  } else {
    m = __byte_perm(x, MAGIC, 0x7650u | (uint32_t)b);
  }
  return __uint_as_float(m);
}
__device__ __forceinline__ float byte_to_float(const uint32_t (&w)[6], int B) {
  return byte_to_magic(w, B) - 8388608.0f;
}

// one 24-byte block row -> r, g, b of its 8 pixels as floats in 0..255
template <int VEC>
__device__ __forceinline__ void load_row_rgb255(const uint8_t* __restrict__ p, float (&r)[8], float (&g)[8], float (&b)[8]) {
  uint32_t w[6];
  load_row24<VEC>(p, w);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    r[j] = byte_to_float(w, 3 * j);
    g[j] = byte_to_float(w, 3 * j + 1);
    b[j] = byte_to_float(w, 3 * j + 2);
  }
}

// Prefetch into L2 the rows of the block a later wave of CTAs will own, turning
// its DRAM latency into L2 latency.
__device__ __forceinline__ void prefetch_block_rows_l2(const uint8_t* __restrict__ base, size_t pitch) {
#pragma unroll
  for (int i = 0; i < 8; ++i) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (size_t)i * pitch));
}

// TMA-engine bulk prefetch into L2 (cp.async.bulk.prefetch.L2) of the 8 image rows of the
// tile that the CTA `ahead` CTAs further on will own: one thread issues at most 16
// instructions for 24 KB, so - unlike the per-thread look-ahead that lost throughput - it
// is free in issue slots.  Needs 16-byte aligned segments (even block offsets, 3W % 16 == 0);
// the caller checks `ok16`.
#ifndef TMF_BULK_AHEAD
#define TMF_BULK_AHEAD 0      // CTAs of look-ahead (0 = off)
#endif
__device__ __forceinline__ void bulk_prefetch_tile(const uint8_t* __restrict__ base, const BlockGeom& g, long long gb0) {
  long long first = gb0, left = kThreads;
  if (first >= g.total_blocks) return;
  if (first + left > g.total_blocks) left = g.total_blocks - first;
  while (left > 0) {
    long long img; int by, bx;
    const size_t org = block_origin(g, first, img, by, bx);
    long long run = g.nbw - bx;                       // blocks to the end of this block-row
    if (run > left) run = left;
    const unsigned bytes = (unsigned)(run * 24);
    if ((bytes & 15u) == 0 && ((org & 15) == 0)) {
#pragma unroll
      for (int r = 0; r < 8; ++r)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(base + org + (size_t)r * g.row_pitch), "r"(bytes));
    }
    first += run;
    left -= run;
  }
}

// pass 1 over the 8 rows of a block: Gram matrix of its luma (rolled loop: small code).
// With KEEP, row i's luma is parked in shared memory as two float4 at
// col[(2i) * kThreads] and col[(2i+1) * kThreads] (thread-private column,
// conflict-free 128-bit accesses) for pass 2.
// ---- packed fp32 (sm_100 FFMA2 / FADD2 / FMUL2) --------------------------------------
// Two adjacent pixels ride in one 64-bit register pair.  Each lane of a packed
// instruction is an ordinary round-to-nearest fp32 operation, so results are
// bit-identical to the scalar formulas in tmf_fast.cuh (which the host test harness
// runs); what changes is the issue cost: measured ~1.4 dispatch cycles per FFMA2 against
// 2 for two FFMAs (profiles/r01_ubench.txt), and fp32 math is over half of this kernel's
// issue slots.  Scalars broadcast and immediates come for free (FFMA2 Rd, Ra.F32, ...).
#ifndef TMF_USE_F32X2
#define TMF_USE_F32X2 1
#endif
__device__ __forceinline__ float2 bc2(float x) { return make_float2(x, x); }

// magic floats of bytes B and B + 3 (same channel of two adjacent pixels) -> (k0, k1)
__device__ __forceinline__ float2 bytes_to_float2(const uint32_t (&w)[6], int B) {
  return __fadd2_rn(make_float2(byte_to_magic(w, B), byte_to_magic(w, B + 3)), bc2(-8388608.0f));
}

// Exact integer luma 299 r + 587 g + 114 b (tmf::luma1000_exact) of the 8 pixels of a row
// without extracting a single byte: a pixel's three bytes sit in one or two of the row's six
// words, and two IDP.2A (16-bit weights x bytes 0-1 or 2-3 of a word, accumulate) cover them
// whatever the phase.  The accumulator starts at 0x4B000000, so the result already is the
// bit pattern of the float 2^23 + luma (luma < 2^18), and one packed FADD per pixel pair
// removes the 2^23.  Against PRMT extraction + fp32 FMAs this halves the issue cost of
// pass 1's colour step and moves it from the ALU pipe (the busier one) to the FMA pipe.
#ifndef TMF_LUMA_IDP
#define TMF_LUMA_IDP 1
#endif
#if TMF_LUMA_IDP && !TMF_USE_F32X2
#error "TMF_LUMA_IDP needs the packed (TMF_USE_F32X2) row code"
#endif
#if TMF_LUMA_IDP
#define TMF_LUMA_UNIT TMF_LUMA1000_UNIT
#else
#define TMF_LUMA_UNIT (1.0f / 255.0f)
#endif
__device__ __forceinline__ uint32_t dp2a_lo(uint32_t bytes, uint32_t w16x2, uint32_t acc) {
  uint32_t d;
  asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(w16x2), "r"(bytes), "r"(acc));
  return d;
}
__device__ __forceinline__ uint32_t dp2a_hi(uint32_t bytes, uint32_t w16x2, uint32_t acc) {
  uint32_t d;
  asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(w16x2), "r"(bytes), "r"(acc));
  return d;
}
// magic float 2^23 + (299 r + 587 g + 114 b) of pixel j of the row
__device__ __forceinline__ float pixel_luma_magic(const uint32_t (&w)[6], int j) {
  constexpr uint32_t kRG = 299u | (587u << 16), kB_ = 114u, k_R = 299u << 16, kGB = 587u | (114u << 16);
  const int k = (3 * j) >> 2;
  uint32_t m;
  switch ((3 * j) & 3) {
    case 0: m = dp2a_hi(w[k], kB_, dp2a_lo(w[k], kRG, 0x4B000000u)); break;          // [r g b .]
    case 1: m = dp2a_hi(w[k], kGB, dp2a_lo(w[k], k_R, 0x4B000000u)); break;          // [. r g b]
    case 2: m = dp2a_lo(w[k + 1], kB_, dp2a_hi(w[k], kRG, 0x4B000000u)); break;      // [. . r g][b . . .]
    default: m = dp2a_lo(w[k + 1], kGB, dp2a_hi(w[k], k_R, 0x4B000000u)); break;     // [. . . r][g b . .]
  }
  return __uint_as_float(m);
}

// luma of the 8 pixels of a row as four pairs; same values as tmf::luma1000_exact
// (TMF_LUMA_IDP) or tmf::luma255_fast
__device__ __forceinline__ void row_luma2(const uint32_t (&w)[6], float2 (&y2)[4]) {
#if TMF_LUMA_IDP
#pragma unroll
  for (int p = 0; p < 4; ++p)
    y2[p] = __fadd2_rn(make_float2(pixel_luma_magic(w, 2 * p), pixel_luma_magic(w, 2 * p + 1)), bc2(-8388608.0f));
  return;
#endif
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int B = 6 * p;                 // byte offset of pixel 2p
    // fma(c, 2^23 + b, -c 2^23) == RN(c b) exactly: blue needs no separate magic subtraction
    const float2 tb = __ffma2_rn(bc2(0.114f), make_float2(byte_to_magic(w, B + 2), byte_to_magic(w, B + 5)),
                                 bc2(-0.114f * 8388608.0f));
    y2[p] = __ffma2_rn(bc2(0.299f), bytes_to_float2(w, B), __ffma2_rn(bc2(0.587f), bytes_to_float2(w, B + 1), tb));
  }
}

// Gram matrix in paired form: gp[i][p] = (G[i][2p], G[i][2p+1]) for the pairs of the
// upper triangle that start at an even column, gd[i] = G[i][i] for odd i.
struct GramPairs {
  float2 gp[8][4];
  float gd[8];
};

__device__ __forceinline__ void gram_accumulate_row2(const float2 (&y2)[4], GramPairs& G) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float yi = (i & 1) ? y2[i >> 1].y : y2[i >> 1].x;
    if (i & 1) G.gd[i] = fmaf(yi, yi, G.gd[i]);
#pragma unroll
    for (int p = (i + 1) >> 1; p < 4; ++p) G.gp[i][p] = __ffma2_rn(bc2(yi), y2[p], G.gp[i][p]);
  }
}

__device__ __forceinline__ void gram_pairs_to_sym(const GramPairs& G, float (&gm)[36]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (i & 1) gm[tmf::sym_idx<8>(i, i)] = G.gd[i];
#pragma unroll
    for (int p = (i + 1) >> 1; p < 4; ++p) {
      gm[tmf::sym_idx<8>(i, 2 * p)] = G.gp[i][p].x;
      gm[tmf::sym_idx<8>(i, 2 * p + 1)] = G.gp[i][p].y;
    }
  }
}

// pass 1 over the 8 rows of a block: Gram matrix of its luma (rolled loop: small code).
// With KEEP, row i's luma is parked in shared memory as two float4 at
// col[(2i) * STRIDE] and col[(2i+1) * STRIDE] (thread-private column,
// conflict-free 128-bit accesses) for pass 2.
// Row addressing follows the type of `pitch`: uint32_t = a running pointer (one 64-bit add per row;
// k_extract_fast / k_sigma0_fast: +0.5 %), size_t = base + i * pitch (k_embed_fast, where the running
// pointer measured 1.7 % slower - profiles/r01_sweep_variants.txt table 15).
template <int VEC, bool KEEP, int STRIDE = kThreads, typename PITCH = uint32_t>
__device__ __forceinline__ void gram_of_block(const uint8_t* __restrict__ base0, PITCH pitch, float (&gm)[36],
                                              float4* __restrict__ col = nullptr) {
  constexpr bool kRunning = sizeof(PITCH) == 4;
  const uint8_t* __restrict__ base = base0;
#if TMF_USE_F32X2
  GramPairs G;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    G.gd[i] = 0.0f;
#pragma unroll
    for (int p = 0; p < 4; ++p) G.gp[i][p] = make_float2(0.0f, 0.0f);
  }
#pragma unroll kRowUnroll
  for (int i = 0; i < 8; ++i) {
    uint32_t w[6];
    float2 y2[4];
    if (kRunning) { load_row24<VEC>(base, w); base += pitch; }
    else load_row24<VEC>(base0 + (size_t)i * pitch, w);
    row_luma2(w, y2);
    if (KEEP) {
      col[(2 * i) * STRIDE] = make_float4(y2[0].x, y2[0].y, y2[1].x, y2[1].y);
      col[(2 * i + 1) * STRIDE] = make_float4(y2[2].x, y2[2].y, y2[3].x, y2[3].y);
    }
    gram_accumulate_row2(y2, G);
  }
  gram_pairs_to_sym(G, gm);
#else
#pragma unroll
  for (int k = 0; k < 36; ++k) gm[k] = 0.0f;
#pragma unroll kRowUnroll
  for (int i = 0; i < 8; ++i) {
    float y[8];
    uint32_t w[6];
    if (kRunning) { load_row24<VEC>(base, w); base += pitch; }
    else load_row24<VEC>(base0 + (size_t)i * pitch, w);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      // fma(c, 2^23 + b, -c 2^23) == RN(c b) exactly (one rounding of an exact sum), so the
      // blue byte needs no separate magic subtraction; same value as luma255_fast
      const float tb = fmaf(0.114f, byte_to_magic(w, 3 * j + 2), -0.114f * 8388608.0f);
      y[j] = fmaf(0.299f, byte_to_float(w, 3 * j), fmaf(0.587f, byte_to_float(w, 3 * j + 1), tb));
    }
    if (KEEP) {
      col[(2 * i) * STRIDE] = make_float4(y[0], y[1], y[2], y[3]);
      col[(2 * i + 1) * STRIDE] = make_float4(y[4], y[5], y[6], y[7]);
    }
    tmf::gram_accumulate_row(y, gm);
  }
#endif
}

// bits(floor-biased float) - 0x4B400000 = the integer level.  TMF_QUANT_IMAD picks the
// pipe: 0 = IADD3 (ALU pipe), 1 = IMAD x*1+c (FMA pipe), 2 = alternate.
#ifndef TMF_QUANT_IMAD
#define TMF_QUANT_IMAD 0
#endif
__device__ __forceinline__ int unbias_imad(float t) {
  int r;
  asm("mad.lo.s32 %0, %1, 1, %2;" : "=r"(r) : "r"(__float_as_int(t)), "r"(-0x4B400000));
  return r;
}
__device__ __forceinline__ int unbias(float t) {
#if TMF_QUANT_IMAD == 1
  return unbias_imad(t);
#else
  return __float_as_int(t) - 0x4B400000;
#endif
}

// Signed variant for the chroma rows of the reference's matrix (watermarking.py:37-39): the
// bit pattern of the float 1.5*2^23 + (cr*r + cg*g + cb*b) for pixel j, |sum| < 2^22.
__device__ __forceinline__ uint32_t dp2a_lo_s(uint32_t bytes, uint32_t w16x2, uint32_t acc) {
  int32_t d;
  asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"((int32_t)w16x2), "r"(bytes), "r"((int32_t)acc));
  return (uint32_t)d;
}
__device__ __forceinline__ uint32_t dp2a_hi_s(uint32_t bytes, uint32_t w16x2, uint32_t acc) {
  int32_t d;
  asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"((int32_t)w16x2), "r"(bytes), "r"((int32_t)acc));
  return (uint32_t)d;
}
template <int CR, int CG, int CB>
__device__ __forceinline__ float pixel_dot_magic_s(const uint32_t (&w)[6], int j) {
  constexpr uint32_t r = (uint32_t)(uint16_t)(int16_t)CR, g = (uint32_t)(uint16_t)(int16_t)CG,
                     b = (uint32_t)(uint16_t)(int16_t)CB;
  constexpr uint32_t kRG = r | (g << 16), kB_ = b, k_R = r << 16, kGB = g | (b << 16);
  const int k = (3 * j) >> 2;
  uint32_t m;
  switch ((3 * j) & 3) {
    case 0: m = dp2a_hi_s(w[k], kB_, dp2a_lo_s(w[k], kRG, 0x4B400000u)); break;
    case 1: m = dp2a_hi_s(w[k], kGB, dp2a_lo_s(w[k], k_R, 0x4B400000u)); break;
    case 2: m = dp2a_lo_s(w[k + 1], kB_, dp2a_hi_s(w[k], kRG, 0x4B400000u)); break;
    default: m = dp2a_lo_s(w[k + 1], kGB, dp2a_hi_s(w[k], k_R, 0x4B400000u)); break;
  }
  return __uint_as_float(m);
}
// Pixel pairs of a row whose pass 2 goes through exact integer chroma (4 IDP.2A per pixel on
// the FMA pipe, 5 packed FMAs per pair) instead of byte extraction (3 PRMT per pixel on the
// ALU pipe, 9 packed FMAs per pair): 0..4, a pipe-balance knob.  Needs the luma stash in
// 1/1000 units (TMF_LUMA_IDP).
#ifndef TMF_PASS2_IDP_PAIRS
#define TMF_PASS2_IDP_PAIRS 0   // measured slower for every value 1..4 (profiles/r01_sweep_variants.txt, table 9)
#endif
#ifndef TMF_PASS2_DELTA
#define TMF_PASS2_DELTA 1       // pass 2 as k_c + small term (see embed_row_fast2)
#endif
#ifndef TMF_QUANT_DENORM
#define TMF_QUANT_DENORM 1      // the quantiser floors straight onto the integer level (subnormal trick, embed_row_fast2)
#endif
#if TMF_QUANT_DENORM && TMF_PASS2_IDP_PAIRS > 0
#error "TMF_QUANT_DENORM scales du by 2^-49: not combined with the TMF_PASS2_IDP_PAIRS experiment"
#endif
#if TMF_PASS2_IDP_PAIRS > 0 && !TMF_LUMA_IDP
#error "TMF_PASS2_IDP_PAIRS needs TMF_LUMA_IDP"
#endif

// pass 2 for one row, packed: bytes of the row, its luma (4 pairs), w (4 pairs), f, c ->
// six output words.  Same arithmetic as tmf::embed_row_fast + pack4_sat_u8.
__device__ __forceinline__ void embed_row_fast2(const uint32_t (&w)[6], const float2 (&y2)[4], const float2 (&w2)[4],
                                                float f, float c, uint32_t (&o)[6]) {
  float2 acc = __fmul2_rn(y2[0], w2[0]);
#pragma unroll
  for (int p = 1; p < 4; ++p) acc = __ffma2_rn(y2[p], w2[p], acc);
  // tmf::dot8 accumulates sequentially; the pairwise order differs by rounding only in z,
  // which is scaled by f ~ 1e-3: far below the quantiser's resolution
  float du = fmaf(f, acc.x + acc.y, c);
  if (TMF_PASS2_DELTA && TMF_QUANT_DENORM) du *= 1.7763568394002505e-15f;   // 2^-49, see below (exact scaling)
  int q[24];
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int B = 6 * p;
    const float2 d2 = __fmul2_rn(bc2(du), w2[p]);
    float2 tR, tG, tB;
    if (p < TMF_PASS2_IDP_PAIRS) {
      // y' = y + d in 0..255 units from the stash (1000 * y255, exact); 1000 * chroma exact
      const float2 yd = __ffma2_rn(bc2(1.0e-3f), y2[p], d2);
      const float2 cb = __fadd2_rn(make_float2(pixel_dot_magic_s<-169, -331, 500>(w, 2 * p),
                                               pixel_dot_magic_s<-169, -331, 500>(w, 2 * p + 1)), bc2(-12582912.0f));
      const float2 cr = __fadd2_rn(make_float2(pixel_dot_magic_s<500, -419, -81>(w, 2 * p),
                                               pixel_dot_magic_s<500, -419, -81>(w, 2 * p + 1)), bc2(-12582912.0f));
      const float2 R = __ffma2_rn(bc2(1.403e-3f), cr, yd);                      // watermarking.py:61-67
      const float2 G = __ffma2_rn(bc2(-0.344e-3f), cb, __ffma2_rn(bc2(-0.714e-3f), cr, yd));
      const float2 Bv = __ffma2_rn(bc2(1.773e-3f), cb, yd);
      tR = __fadd2_rd(R, bc2(12582912.0f)); tG = __fadd2_rd(G, bc2(12582912.0f)); tB = __fadd2_rd(Bv, bc2(12582912.0f));
    } else {
#if TMF_PASS2_DELTA && TMF_QUANT_DENORM
      // As below (out_c = floor(k_c + s_c)), but the floor lands DIRECTLY on the integer level:
      // byte k extracted with a zero filler IS the bit pattern of the subnormal float k * 2^-149,
      // and FFMA.RM(s_c * 2^-49, 2^-100, k * 2^-149) is the exact sum (k_c + s_c) * 2^-149 rounded
      // toward -inf to a multiple of 2^-149, i.e. the float whose bits are floor(k_c + s_c) - or
      // a negative float (sign bit set = a hugely negative s32) when the level is below 0,
      // which the saturating pack clips to 0 like any other negative.  No bias to remove: 24
      // integer subtractions per row gone.  The 2^-49 rides in E's constants (scaled by 2^100:
      // the differences u, v are subnormal too, (r - g) * 2^-149, exact) and in du (caller);
      // power-of-two scalings are exact, so every value is the one the biased form computes.
      const float2 mr = make_float2(byte_to_magic<0u>(w, B), byte_to_magic<0u>(w, B + 3));
      const float2 mg = make_float2(byte_to_magic<0u>(w, B + 1), byte_to_magic<0u>(w, B + 4));
      const float2 mb = make_float2(byte_to_magic<0u>(w, B + 2), byte_to_magic<0u>(w, B + 5));
      const float2 u = __ffma2_rn(mg, bc2(-1.0f), mr), v = __ffma2_rn(mg, bc2(-1.0f), mb);
      constexpr float k2p100 = 1.2676506002282294e30f;      // 2^100
      const float2 sR = __ffma2_rn(bc2(5.00e-4f * k2p100), u, __ffma2_rn(bc2(3.57e-4f * k2p100), v, d2));
      const float2 sG = __ffma2_rn(bc2(1.36e-4f * k2p100), u, __ffma2_rn(bc2(-1.66e-4f * k2p100), v, d2));
      const float2 sB = __ffma2_rn(bc2(-6.37e-4f * k2p100), u, __ffma2_rn(bc2(5.00e-4f * k2p100), v, d2));
      constexpr float k2m100 = 7.888609052210118e-31f;      // 2^-100
      tR = __ffma2_rd(sR, bc2(k2m100), mr); tG = __ffma2_rd(sG, bc2(k2m100), mg); tB = __ffma2_rd(sB, bc2(k2m100), mb);
#elif TMF_PASS2_DELTA
      // M = I + E with E's rows summing to zero (a grey pixel maps to itself), so
      //   out_c = floor(k_c + s_c),  s_c = E_c0 (r - g) + E_c2 (b - g) + d   (|s_c| small).
      // The bytes are extracted straight onto the quantiser's bias, m_c = 1.5*2^23 + k_c
      // (exact); r - g and b - g are exact differences of those; and FADD.RD(s_c, m_c) is
      // 1.5*2^23 + floor(k_c + s_c) exactly.  Two FMAs per sample instead of three, no
      // separate removal of the bias, and the small term s_c is carried at full precision.
      const float2 mr = make_float2(byte_to_magic<0x4B400000u>(w, B), byte_to_magic<0x4B400000u>(w, B + 3));
      const float2 mg = make_float2(byte_to_magic<0x4B400000u>(w, B + 1), byte_to_magic<0x4B400000u>(w, B + 4));
      const float2 mb = make_float2(byte_to_magic<0x4B400000u>(w, B + 2), byte_to_magic<0x4B400000u>(w, B + 5));
      const float2 u = __ffma2_rn(mg, bc2(-1.0f), mr), v = __ffma2_rn(mg, bc2(-1.0f), mb);
      const float2 sR = __ffma2_rn(bc2(5.00e-4f), u, __ffma2_rn(bc2(3.57e-4f), v, d2));
      const float2 sG = __ffma2_rn(bc2(1.36e-4f), u, __ffma2_rn(bc2(-1.66e-4f), v, d2));
      const float2 sB = __ffma2_rn(bc2(-6.37e-4f), u, __ffma2_rn(bc2(5.00e-4f), v, d2));
      tR = __fadd2_rd(sR, mr); tG = __fadd2_rd(sG, mg); tB = __fadd2_rd(sB, mb);
#else
      const float2 r2 = bytes_to_float2(w, B), g2 = bytes_to_float2(w, B + 1), b2 = bytes_to_float2(w, B + 2);
      const float2 R = __ffma2_rn(bc2(1.0005f), r2, __ffma2_rn(bc2(-8.57e-4f), g2, __ffma2_rn(bc2(3.57e-4f), b2, d2)));
      const float2 G = __ffma2_rn(bc2(1.36e-4f), r2, __ffma2_rn(bc2(1.00003f), g2, __ffma2_rn(bc2(-1.66e-4f), b2, d2)));
      const float2 Bv = __ffma2_rn(bc2(-6.37e-4f), r2, __ffma2_rn(bc2(1.37e-4f), g2, __ffma2_rn(bc2(1.0005f), b2, d2)));
      tR = __fadd2_rd(R, bc2(12582912.0f)); tG = __fadd2_rd(G, bc2(12582912.0f)); tB = __fadd2_rd(Bv, bc2(12582912.0f));
#endif
    }
    if (TMF_PASS2_DELTA && TMF_QUANT_DENORM && p >= TMF_PASS2_IDP_PAIRS) {
      q[B] = __float_as_int(tR.x); q[B + 1] = __float_as_int(tG.x); q[B + 2] = __float_as_int(tB.x);
      q[B + 3] = __float_as_int(tR.y); q[B + 4] = __float_as_int(tG.y); q[B + 5] = __float_as_int(tB.y);
    } else {
      q[B] = unbias(tR.x); q[B + 1] = unbias(tG.x); q[B + 2] = unbias(tB.x);
      q[B + 3] = unbias(tR.y); q[B + 4] = unbias(tG.y); q[B + 5] = unbias(tB.y);
    }
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) o[k] = tmf::pack4_sat_u8(q[4 * k], q[4 * k + 1], q[4 * k + 2], q[4 * k + 3]);
}

// Fused embed, FAST mode.
//
// Blocks whose watermark value is 0 get d = f32(f64(s0) + alpha*0) - s0 = 0 exactly
// (watermarking.py:198): their output is the colour round trip alone, so a lane with a
// zero mark skips pass 1 and the eigenpair (a whole warp of zero marks - black areas of
// the map - saves the issue slots; in a mixed warp the lane just idles).  Sorting the
// blocks of a CTA so that warps become uniform was measured and is slower: the scattered
// 24-byte row pieces cost more L1 wavefronts than the skipped work saves
// (profiles/r01_sweep_variants.txt, third table).
template <int VEC>
__global__ void __launch_bounds__(kEmbedThreads, kEmbedMinCtas)
k_embed_fast(const uint8_t* __restrict__ rgb, uint8_t* __restrict__ out, BlockGeom g,
             const uint8_t* __restrict__ wm, int wm_shared, double alpha) {
#if TMF_EMBED_STASH
  __shared__ float4 lum[16 * kEmbedThreads];      // 32 KB: the block's luma, thread-private column
  float4* col = lum + threadIdx.x;
#else
  float4* col = nullptr;                     // nothing parked: pass 2 recomputes the luma (8 IDP.2A per row)
#endif
#if TMF_EMBED_PERSIST
  // persistent form: the grid is TMF_EMBED_PERSIST CTAs per SM, a thread walks blocks gb, gb + stride, ...
  // and asks L2 for its NEXT block's rows while it works on this one
  const long long stride = (long long)gridDim.x * kEmbedThreads;
  for (long long gb = (long long)blockIdx.x * kEmbedThreads + threadIdx.x; gb < g.total_blocks; gb += stride) {
#else
  const long long gb = (long long)blockIdx.x * kEmbedThreads + threadIdx.x;
  if (gb >= g.total_blocks) return;
  {
#endif
  long long img; int by, bx;
  uint32_t in_img;
  const size_t org = block_origin(g, gb, img, by, bx, &in_img);
  const uint8_t* src = rgb + org;
  prefetch_block_rows(src, g.pitch32, TMF_PREFETCH_FROM);
#if TMF_EMBED_PERSIST
  if (gb + stride < g.total_blocks) {
    long long i2; int y2, x2;
    const uint8_t* nxt = rgb + block_origin(g, gb + stride, i2, y2, x2);
#pragma unroll
    for (int i = 0; i < 8; ++i) { asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt)); nxt += g.pitch32; }
  }
#endif
  if (TMF_BULK_AHEAD > 0 && VEC == 8 && threadIdx.x == 0)
    bulk_prefetch_tile(rgb, g, ((long long)blockIdx.x + TMF_BULK_AHEAD) * kEmbedThreads);
  const uint32_t mark = (uint32_t)__ldg(wm + (wm_shared ? in_img : (uint32_t)gb));   // map index: 32 bits
  float w[8], f = 0.0f, c = 0.0f;
  if (mark != 0) {
    float gm[36];
#if TMF_EMBED_ROWPTR
    gram_of_block<VEC, TMF_EMBED_STASH != 0, kEmbedThreads>(src, g.pitch32, gm, col);
#else
    gram_of_block<VEC, TMF_EMBED_STASH != 0, kEmbedThreads, size_t>(src, g.row_pitch, gm, col);
#endif
    if (TMF_EMBED_REPREFETCH == 1) prefetch_block_rows(src, g.pitch32);
    tmf::embed_block_scalars_fast(gm, alpha, mark, w, f, c, nullptr, TMF_LUMA_UNIT);
    if (TMF_EMBED_REPREFETCH == 2) prefetch_block_rows(src, g.pitch32);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = 0.0f;
  }
  // pass 2: the rows again (L1/L2 hits), rank-1 update, colour out, quantise, store.
  // (TMF_EMBED_ROWPTR = 1: one running pointer, the output row at `input row + (out - rgb)`, a
  // warp-uniform distance: fewer address instructions, yet measured 1.7 % slower than base + i * pitch)
#if TMF_EMBED_ROWPTR
  const ptrdiff_t to_out = out - rgb;
#else
  uint8_t* dst = out + org;
#endif
#if TMF_USE_F32X2
  const float2 w2[4] = {make_float2(w[0], w[1]), make_float2(w[2], w[3]), make_float2(w[4], w[5]), make_float2(w[6], w[7])};
#endif
#pragma unroll kRowUnrollP2
  for (int i = 0; i < 8; ++i) {
    uint32_t o[6];
    float4 ya = make_float4(0.f, 0.f, 0.f, 0.f), yb = ya;
    if (TMF_EMBED_STASH && mark != 0) { ya = col[(2 * i) * kEmbedThreads]; yb = col[(2 * i + 1) * kEmbedThreads]; }
#if TMF_USE_F32X2
    uint32_t wd[6];
#if TMF_EMBED_ROWPTR
    const uint8_t* rowp = src;
#else
    const uint8_t* rowp = src + (size_t)i * g.row_pitch;
#endif
    load_row24<VEC>(rowp, wd);
    float2 y2[4] = {make_float2(ya.x, ya.y), make_float2(ya.z, ya.w), make_float2(yb.x, yb.y), make_float2(yb.z, yb.w)};
    if ((!TMF_EMBED_STASH && mark != 0) || (TMF_PASS2_IDP_PAIRS > 0 && mark == 0)) row_luma2(wd, y2);
    embed_row_fast2(wd, y2, w2, f, c, o);
#else
    float r[8], gg[8], b[8];
    int q[24];
    const uint8_t* rowp = src + (size_t)i * g.row_pitch;
    load_row_rgb255<VEC>(rowp, r, gg, b);
    const float y[8] = {ya.x, ya.y, ya.z, ya.w, yb.x, yb.y, yb.z, yb.w};
    tmf::embed_row_fast(r, gg, b, y, w, f, c, q);
#pragma unroll
    for (int k = 0; k < 6; ++k) o[k] = tmf::pack4_sat_u8(q[4 * k], q[4 * k + 1], q[4 * k + 2], q[4 * k + 3]);
#endif
#if TMF_EMBED_ROWPTR
    store_row24<VEC>(const_cast<uint8_t*>(rowp) + to_out, o);
    src += g.pitch32;
#else
    store_row24<VEC>(dst + (size_t)i * g.row_pitch, o);
#endif
  }
  }   // blocks of this thread
}


// ---------------------------------------------------------------------------
// Fused embed, FAST mode, rows staged through shared memory by the TMA engine.
//
// Why: with per-thread row accesses (k_embed_fast) a warp's 64-bit load covers 256 useful
// bytes spread over 768, i.e. 6-7 L1 wavefronts instead of 2, three times per row; the six
// resident CTAs own 144 KB of input rows but share < 64 KB of L1, so pass 2 re-reads most
// rows from L2 (ncu: L1 hit rate 56 %, 20 % of warp samples on long_scoreboard, L1 wavefronts
// 46 % busy), and every 32-byte output sector reaches L2 in three partial writes.  Here a
// WARP owns 32 adjacent blocks = 8 image-row pieces of 768 contiguous bytes:
//   * lanes 0-7 ask the TMA engine for one row piece each (cp.async.bulk global -> shared,
//     one mbarrier per warp; a warp that straddles the end of a block-row has two runs, the
//     second issued by lanes 8-15) - one warp instruction, not a loop in one lane,
//   * both row passes read the tile with conflict-free LDS.64 (thread stride 24 B: the 16
//     lanes of a half-warp hit 16 distinct bank pairs),
//   * pass 2 writes its bytes back IN PLACE (the 24 bytes of a row piece are private to
//     their lane), and the same lanes hand the tile to the TMA engine again
//     (cp.async.bulk shared -> global): full-line writes, no per-thread STG,
//   * nothing else is parked in shared memory: pass 2 recomputes the row's exact integer luma
//     from the staged bytes (8 IDP.2A) - the same values pass 1 had, so the results are
//     identical to k_embed_fast's - which keeps the footprint at 24 KB per CTA, 6 CTAs/SM.
// No __syncthreads: warps of a CTA run independently.  Needs 16-byte aligned runs: base
// pointers, image stride and 3*W multiples of 16, an even number (>= 32) of blocks per
// block-row; everything else takes k_embed_fast.
// ---------------------------------------------------------------------------
#ifndef TMF_TMA_MIN_CTAS
#define TMF_TMA_MIN_CTAS 6
#endif
constexpr int kTileRowBytes = 32 * 24;             // one image-row piece of a warp's 32 blocks
constexpr int kTileBytes = 8 * kTileRowBytes;      // 6 KB per warp
constexpr int kWarps = kThreads / 32;
constexpr int kEmbedTmaSmem = kWarps * kTileBytes + kWarps * 8;   // tiles + one mbarrier per warp

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "TMF_WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@!p bra TMF_WAIT_%=;\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}

// A warp's 32 consecutive blocks are one run, or two when they straddle the end of a
// block-row (nbw >= 32).  Lane L < 16 moves image row (L & 7) of run (L >> 3).
struct WarpRuns {
  size_t org;        // byte offset (in the batch) of this lane's run
  uint32_t soff;     // byte offset of the run inside a tile row
  uint32_t bytes;    // 0: this lane moves nothing
};
__device__ __forceinline__ WarpRuns warp_runs(const BlockGeom& g, int lane, int cnt, size_t my_org, int my_bx) {
  const int bx0 = __shfl_sync(0xffffffffu, my_bx, 0);
  int run0 = g.nbw - bx0;
  if (run0 > cnt) run0 = cnt;
  const size_t org0 = __shfl_sync(0xffffffffu, (unsigned long long)my_org, 0);
  const size_t org1 = __shfl_sync(0xffffffffu, (unsigned long long)my_org, run0 & 31);   // first block of run 1
  WarpRuns r;
  const int which = lane >> 3;
  r.org = which == 0 ? org0 : org1;
  r.soff = which == 0 ? 0u : (uint32_t)run0 * 24u;
  r.bytes = which == 0 ? (uint32_t)run0 * 24u : (which == 1 ? (uint32_t)(cnt - run0) * 24u : 0u);
  return r;
}

__global__ void __launch_bounds__(kThreads, TMF_TMA_MIN_CTAS)
k_embed_fast_tma(const uint8_t* __restrict__ rgb, uint8_t* __restrict__ out, BlockGeom g,
                 const uint8_t* __restrict__ wm, int wm_shared, double alpha) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long gw0 = (long long)blockIdx.x * kThreads + warp * 32;   // first block of this warp
  if (gw0 >= g.total_blocks) return;                                    // whole warp; nothing is CTA-wide here
  const long long rem = g.total_blocks - gw0;
  const int cnt = rem < 32 ? (int)rem : 32;
  const bool live = lane < cnt;
  uint8_t* tile = smem + warp * kTileBytes;
  const uint32_t tile_s = smem_u32(tile);
  const uint32_t bar = smem_u32(smem + kWarps * kTileBytes + warp * 8);
  if (lane == 0) mbar_init(bar, 1);
  long long img; int by, bx;
  const size_t org = block_origin(g, gw0 + (live ? lane : 0), img, by, bx);
  const WarpRuns mv = warp_runs(g, lane, cnt, org, bx);
  const int row = lane & 7;
  __syncwarp();                      // the barrier is initialised before anything signals or polls it
  if (lane == 0) mbar_expect_tx(bar, (uint32_t)cnt * 24u * 8u);
  if (mv.bytes) bulk_g2s(tile_s + row * kTileRowBytes + mv.soff, rgb + mv.org + (size_t)row * g.row_pitch, mv.bytes, bar);
  uint32_t mark = 0;
  if (live) mark = (uint32_t)__ldg(wm + (wm_shared ? 0 : img * g.blocks_per_img) + (long long)by * g.nbw + bx);
  mbar_wait(bar, 0);
  uint8_t* mine = tile + lane * 24;
  float w[8], f = 0.0f, c = 0.0f;
  if (live && mark != 0) {
    float gm[36];
    gram_of_block<0, false>(mine, kTileRowBytes, gm);
    tmf::embed_block_scalars_fast(gm, alpha, mark, w, f, c, nullptr, TMF_LUMA_UNIT);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = 0.0f;
  }
  if (live) {
    const float2 w2[4] = {make_float2(w[0], w[1]), make_float2(w[2], w[3]), make_float2(w[4], w[5]), make_float2(w[6], w[7])};
#pragma unroll kRowUnroll
    for (int i = 0; i < 8; ++i) {
      uint32_t o[6], wd[6];
      load_row24<0>(mine + i * kTileRowBytes, wd);
      float2 y2[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
      if (mark != 0 || TMF_PASS2_IDP_PAIRS > 0) row_luma2(wd, y2);
      embed_row_fast2(wd, y2, w2, f, c, o);
      store_row24<0>(mine + i * kTileRowBytes, o);
    }
  }
  // generic-proxy writes -> visible to the async proxy, then the moving lanes hand the tile over
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (mv.bytes) {
    bulk_s2g(out + mv.org + (size_t)row * g.row_pitch, tile_s + row * kTileRowBytes + mv.soff, mv.bytes);
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the tile must outlive the reads
  }
}

template <int VEC>
__global__ void __launch_bounds__(kExtractThreads, TMF_FAST_MIN_CTAS * (128 / TMF_EXTRACT_THREADS))
k_extract_fast(const uint8_t* __restrict__ wmk, const uint8_t* __restrict__ orig, uint8_t* __restrict__ out_wm,
               BlockGeom g, double alpha) {
  const long long gb = (long long)blockIdx.x * kExtractThreads + threadIdx.x;
  if (gb >= g.total_blocks) return;
  long long img; int by, bx;
  const size_t org = block_origin(g, gb, img, by, bx);
  prefetch_block_rows(wmk + org, g.pitch32, TMF_PREFETCH_FROM);
  prefetch_block_rows(orig + org, g.pitch32);
  if (TMF_BULK_AHEAD > 0 && VEC == 8 && threadIdx.x == 0) {
    bulk_prefetch_tile(wmk, g, ((long long)blockIdx.x + TMF_BULK_AHEAD) * kThreads);
    bulk_prefetch_tile(orig, g, ((long long)blockIdx.x + TMF_BULK_AHEAD) * kThreads);
  }
  // the two images go through ONE copy of the code (rolled loop): inlining pass 1 and the
  // eigen-solver twice made the kernel 44 KB and cost ~14 % in instruction-fetch stalls
  float sw = 0.0f, so = 0.0f;
#pragma unroll 1
  for (int which = 0; which < 2; ++which) {
    float gm[36];
    gram_of_block<VEC, false>((which == 0 ? wmk : orig) + org, g.pitch32, gm);
    const float sg = tmf::sigma0_from_gram_fast(gm, nullptr, TMF_LUMA_UNIT);
    if (which == 0) sw = sg; else so = sg;
  }
  out_wm[gb] = (uint8_t)tmf::extract_level(sw, so, alpha);
}

template <int VEC>
__global__ void __launch_bounds__(kThreads, TMF_FAST_MIN_CTAS)
k_sigma0_fast(const uint8_t* __restrict__ rgb, float* __restrict__ sigma0, BlockGeom g) {
  const long long gb = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (gb >= g.total_blocks) return;
  long long img; int by, bx;
  const size_t org = block_origin(g, gb, img, by, bx);
  float gm[36];
  prefetch_block_rows(rgb + org, g.pitch32);
  gram_of_block<VEC, false>(rgb + org, g.pitch32, gm);
  sigma0[gb] = tmf::sigma0_from_gram_fast(gm, nullptr, TMF_LUMA_UNIT);
}


// ---------------------------------------------------------------------------
// Other block sizes of the reference's UI (embed_watermark_page.py:324-331: 4..16,
// even).  Same streaming algebra (tmf_fast.cuh is templated on N), one thread per
// block, plain 32-bit or byte accesses; correctness first - the tuned path is N = 8.
// Both `mode`s take this path for N != 8.
// ---------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ size_t block_origin_n(const BlockGeom& g, long long gb, long long& img, int& by, int& bx) {
  const uint32_t n = (uint32_t)gb;                       // total_blocks < 2^31 (make_geom): division-free, as block_origin
  const uint32_t im = fastdiv(n, g.div_bpi);
  const uint32_t r = n - im * (uint32_t)g.blocks_per_img;
  const uint32_t y = fastdiv(r, g.div_nbw);
  img = im;
  by = (int)y;
  bx = (int)(r - y * (uint32_t)g.nbw);
  return (size_t)im * g.img_stride + (size_t)y * N * g.row_pitch + (size_t)bx * (3 * N);
}

// byte b (0..3) of w as a float, through the 2^23 magic number: one PRMT + one FADD on the
// full-rate pipes instead of I2F.U8 (16/clk/SM, profiles/r01_ubench.txt)
__device__ __forceinline__ float byte_of_word_f(uint32_t w, int b) {
  return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7650u | (uint32_t)b)) - 8388608.0f;
}

template <int N, int AL4>
__device__ __forceinline__ void load_row_rgb255_n(const uint8_t* __restrict__ p, float* r, float* g, float* b) {
  float v[3 * N];
  if (AL4 == 4) {
#pragma unroll
    for (int k = 0; k < (3 * N) / 4; ++k) {
      const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(p) + k);
#pragma unroll
      for (int q = 0; q < 4; ++q) v[4 * k + q] = byte_of_word_f(w, q);
    }
  } else if (AL4 == 2) {
#pragma unroll
    for (int k = 0; k < (3 * N) / 2; ++k) {
      const uint32_t w = __ldg(reinterpret_cast<const uint16_t*>(p) + k);     // I2F.U8 with a byte selector: measured
      v[2 * k] = (float)(uint8_t)w; v[2 * k + 1] = (float)(uint8_t)(w >> 8);    // faster than PRMT + FADD on this path
    }
  } else {
#pragma unroll
    for (int k = 0; k < 3 * N; ++k) v[k] = (float)__ldg(p + k);
  }
#pragma unroll
  for (int j = 0; j < N; ++j) { r[j] = v[3 * j]; g[j] = v[3 * j + 1]; b[j] = v[3 * j + 2]; }
}

template <int N, int AL4>
__device__ __forceinline__ void gram_of_block_n(const uint8_t* __restrict__ base, size_t pitch, float* gm) {
#pragma unroll
  for (int k = 0; k < N * (N + 1) / 2; ++k) gm[k] = 0.0f;
#pragma unroll 1
  for (int i = 0; i < N; ++i) {
    float r[N], g[N], b[N], y[N];
    load_row_rgb255_n<N, AL4>(base + (size_t)i * pitch, r, g, b);
#pragma unroll
    for (int j = 0; j < N; ++j) y[j] = tmf::luma255_fast(r[j], g[j], b[j]);
    tmf::gram_accumulate_row<N>(y, gm);
  }
}

template <int N, int AL4>
__global__ void __launch_bounds__(kThreads)
k_embed_fast_n(const uint8_t* __restrict__ rgb, uint8_t* __restrict__ out, BlockGeom g,
               const uint8_t* __restrict__ wm, int wm_shared, double alpha) {
  const long long gb = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (gb >= g.total_blocks) return;
  long long img; int by, bx;
  const size_t org = block_origin_n<N>(g, gb, img, by, bx);
  const uint8_t* src = rgb + org;
  uint8_t* dst = out + org;
  const long long wi = (wm_shared ? 0 : img * g.blocks_per_img) + (long long)by * g.nbw + bx;
  const uint32_t mark = (uint32_t)__ldg(wm + wi);
  float w[N], f = 0.0f, c = 0.0f;
#pragma unroll
  for (int i = 0; i < N; ++i) w[i] = 0.0f;
  if (mark != 0) {
    float gm[N * (N + 1) / 2];
    gram_of_block_n<N, AL4>(src, g.row_pitch, gm);
    tmf::embed_block_scalars_fast<N>(gm, alpha, mark, w, f, c, nullptr);
  }
#pragma unroll 1
  for (int i = 0; i < N; ++i) {
    float r[N], gg[N], b[N], y[N];
    int q[3 * N];
    load_row_rgb255_n<N, AL4>(src + (size_t)i * g.row_pitch, r, gg, b);
#pragma unroll
    for (int j = 0; j < N; ++j) y[j] = tmf::luma255_fast(r[j], gg[j], b[j]);
    tmf::embed_row_fast<N>(r, gg, b, y, w, f, c, q);
    uint8_t* d = dst + (size_t)i * g.row_pitch;
    if (AL4 == 4) {
#pragma unroll
      for (int k = 0; k < (3 * N) / 4; ++k)
        reinterpret_cast<uint32_t*>(d)[k] = tmf::pack4_sat_u8(q[4 * k], q[4 * k + 1], q[4 * k + 2], q[4 * k + 3]);
    } else if (AL4 == 2) {
#pragma unroll
      for (int k = 0; k < (3 * N) / 2; ++k)
        reinterpret_cast<uint16_t*>(d)[k] = (uint16_t)tmf::pack4_sat_u8(q[2 * k], q[2 * k + 1], 0, 0);
    } else {
#pragma unroll
      for (int k = 0; k < 3 * N; ++k) d[k] = (uint8_t)min(max(q[k], 0), 255);
    }
  }
}

template <int N, int AL4>
__device__ __forceinline__ float sigma0_of_block_n(const uint8_t* __restrict__ base, size_t pitch) {
  float gm[N * (N + 1) / 2];
  gram_of_block_n<N, AL4>(base, pitch, gm);
  return tmf::sigma0_from_gram_fast<N>(gm, nullptr);
}

template <int N, int AL4>
__global__ void __launch_bounds__(kThreads)
k_extract_fast_n(const uint8_t* __restrict__ wmk, const uint8_t* __restrict__ orig, uint8_t* __restrict__ out_wm,
                 BlockGeom g, double alpha) {
  const long long gb = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (gb >= g.total_blocks) return;
  long long img; int by, bx;
  const size_t org = block_origin_n<N>(g, gb, img, by, bx);
  const float sw = sigma0_of_block_n<N, AL4>(wmk + org, g.row_pitch);
  const float so = sigma0_of_block_n<N, AL4>(orig + org, g.row_pitch);
  out_wm[gb] = (uint8_t)tmf::extract_level(sw, so, alpha);
}

template <int N, int AL4>
__global__ void __launch_bounds__(kThreads)
k_sigma0_fast_n(const uint8_t* __restrict__ rgb, float* __restrict__ sigma0, BlockGeom g) {
  const long long gb = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (gb >= g.total_blocks) return;
  long long img; int by, bx;
  const size_t org = block_origin_n<N>(g, gb, img, by, bx);
  sigma0[gb] = sigma0_of_block_n<N, AL4>(rgb + org, g.row_pitch);
}

// ---------------------------------------------------------------------------
// standalone batched SVD / DCT / colour taps
// ---------------------------------------------------------------------------
constexpr int kPad = 65;   // smem row stride (floats) for a 64-float block: conflict-free both ways

__device__ __forceinline__ void cswap_cols(float* s, float* a, float* v, bool with_v, int i, int j) {
  // order so that s[i] >= s[j]
  const bool sw = s[i] < s[j];
  const float si = s[i], sj = s[j];
  s[i] = sw ? sj : si; s[j] = sw ? si : sj;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const float x = a[8 * r + i], y = a[8 * r + j];
    a[8 * r + i] = sw ? y : x; a[8 * r + j] = sw ? x : y;
  }
  if (with_v) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const float x = v[8 * r + i], y = v[8 * r + j];
      v[8 * r + i] = sw ? y : x; v[8 * r + j] = sw ? x : y;
    }
  }
}

// Complete the columns of U (row-major u[8*r+c] in LOCAL memory) flagged in
// `null_mask` to an orthonormal basis: twice-iterated Gram-Schmidt of the unit
// vectors e_0..e_7 against the columns already fixed.  Rare path (rank-deficient
// blocks only), so dynamic indexing / local memory is acceptable here.
__device__ __noinline__ void complete_u_columns(float* u, unsigned null_mask) {
  unsigned fixed = (~null_mask) & 0xffu;
  int cand = 0;
  for (int k = 0; k < 8; ++k) {
    if (!((null_mask >> k) & 1u)) continue;
    for (; cand < 8; ++cand) {
      float wv[8];
      for (int r = 0; r < 8; ++r) wv[r] = (r == cand) ? 1.0f : 0.0f;
      for (int pass = 0; pass < 2; ++pass) {
        for (int c = 0; c < 8; ++c) {
          if (!((fixed >> c) & 1u)) continue;
          float d = 0.f;
          for (int r = 0; r < 8; ++r) d = fmaf(u[8 * r + c], wv[r], d);
          for (int r = 0; r < 8; ++r) wv[r] = fmaf(-d, u[8 * r + c], wv[r]);
        }
      }
      float n2 = 0.f;
      for (int r = 0; r < 8; ++r) n2 = fmaf(wv[r], wv[r], n2);
      if (n2 > 0.25f) {
        const float inv = rsqrtf(n2);
        for (int r = 0; r < 8; ++r) u[8 * r + k] = wv[r] * inv;
        fixed |= 1u << k;
        ++cand;
        break;
      }
    }
  }
}

template <bool WITH_UV>
__global__ void __launch_bounds__(kThreads)
k_svd8x8(const float* __restrict__ blocks, long long nblocks, float* __restrict__ S, float* __restrict__ U,
         float* __restrict__ Vt, int* __restrict__ sweeps_out, int complete_u) {
  __shared__ float sm[kThreads * kPad];
  const long long b0 = (long long)blockIdx.x * kThreads;
  const int nb = (int)min((long long)kThreads, nblocks - b0);
  const int t = threadIdx.x;

  // coalesced stage-in
  const float* src = blocks + b0 * 64;
  for (int idx = t; idx < nb * 64; idx += kThreads) sm[(idx >> 6) * kPad + (idx & 63)] = __ldg(src + idx);
  __syncthreads();

  float a[64], v[WITH_UV ? 64 : 1], s[8];
  int sweeps = 0;
  float unscale = 1.0f;
  unsigned null_mask = 0;
  if (t < nb) {
#pragma unroll
    for (int k = 0; k < 64; ++k) a[k] = sm[t * kPad + k];
    sweeps = tmf::jacobi_svd8<WITH_UV>(a, v, unscale);
    float n2[8];
    tmf::column_norms2(a, n2);
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = tmf::f_sqrt(n2[j]);   // scaled domain
    // 19-comparator sorting network, descending
#define CS(i, j) cswap_cols(s, a, v, WITH_UV, i, j)
    CS(0, 1); CS(2, 3); CS(4, 5); CS(6, 7);
    CS(0, 2); CS(1, 3); CS(4, 6); CS(5, 7);
    CS(1, 2); CS(5, 6); CS(0, 4); CS(3, 7);
    CS(1, 5); CS(2, 6);
    CS(1, 4); CS(3, 6);
    CS(2, 4); CS(3, 5);
    CS(3, 4);
#undef CS
    if (WITH_UV) {
      // U = (A V) diag(1/sigma); columns at the noise floor are zeroed (or completed below)
      const float thr = 1.0e-6f * s[0];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const bool ok = s[j] > thr;
        null_mask |= ok ? 0u : (1u << j);
        const float inv = ok ? tmf::f_div(1.0f, s[j]) : 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) a[8 * i + j] *= inv;
      }
      if (s[0] == 0.0f) null_mask = 0xffu;
    }
  }
  __syncthreads();

  // S out
  if (t < nb) {
#pragma unroll
    for (int j = 0; j < 8; ++j) sm[t * 9 + j] = s[j] * unscale;
  }
  __syncthreads();
  for (int idx = t; idx < nb * 8; idx += kThreads) S[b0 * 8 + idx] = sm[(idx >> 3) * 9 + (idx & 7)];
  if (sweeps_out && t < nb) sweeps_out[b0 + t] = sweeps;
  if (!WITH_UV) return;
  __syncthreads();

  // U out
  if (t < nb) {
    if (complete_u && null_mask) {
      float u[64];
#pragma unroll
      for (int k = 0; k < 64; ++k) u[k] = a[k];
      complete_u_columns(u, null_mask);
#pragma unroll
      for (int k = 0; k < 64; ++k) a[k] = u[k];
    }
#pragma unroll
    for (int k = 0; k < 64; ++k) sm[t * kPad + k] = a[k];
  }
  __syncthreads();
  for (int idx = t; idx < nb * 64; idx += kThreads) U[b0 * 64 + idx] = sm[(idx >> 6) * kPad + (idx & 63)];
  __syncthreads();
  // Vt out: Vt[k][j] = V[j][k]
  if (t < nb) {
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) sm[t * kPad + 8 * k + j] = v[WITH_UV ? 8 * j + k : 0];
  }
  __syncthreads();
  for (int idx = t; idx < nb * 64; idx += kThreads) Vt[b0 * 64 + idx] = sm[(idx >> 6) * kPad + (idx & 63)];
}

__global__ void __launch_bounds__(kThreads)
k_dct8x8(const float* __restrict__ in, float* __restrict__ out, long long nblocks, int inverse) {
  __shared__ float sm[kThreads * kPad];
  const long long b0 = (long long)blockIdx.x * kThreads;
  const int nb = (int)min((long long)kThreads, nblocks - b0);
  const int t = threadIdx.x;
  for (int idx = t; idx < nb * 64; idx += kThreads) sm[(idx >> 6) * kPad + (idx & 63)] = __ldg(in + b0 * 64 + idx);
  __syncthreads();
  if (t < nb) {
    float a[64];
#pragma unroll
    for (int k = 0; k < 64; ++k) a[k] = sm[t * kPad + k];
    if (inverse) tmf::idct8x8(a); else tmf::dct8x8(a);
#pragma unroll
    for (int k = 0; k < 64; ++k) sm[t * kPad + k] = a[k];
  }
  __syncthreads();
  for (int idx = t; idx < nb * 64; idx += kThreads) out[b0 * 64 + idx] = sm[(idx >> 6) * kPad + (idx & 63)];
}

__global__ void __launch_bounds__(256)
k_rgb2ycc(const uint8_t* __restrict__ rgb, float* __restrict__ ycc, long long npx) {
  const long long p = (long long)blockIdx.x * 256 + threadIdx.x;
  if (p >= npx) return;
  const float r = tmf::unit_from_u8(__ldg(rgb + 3 * p));
  const float g = tmf::unit_from_u8(__ldg(rgb + 3 * p + 1));
  const float b = tmf::unit_from_u8(__ldg(rgb + 3 * p + 2));
  float cb, cr;
  tmf::chroma_exact(r, g, b, cb, cr);
  ycc[3 * p] = tmf::luma_exact(r, g, b);
  ycc[3 * p + 1] = cb;
  ycc[3 * p + 2] = cr;
}

__global__ void __launch_bounds__(256)
k_ycc2rgb(const float* __restrict__ ycc, uint8_t* __restrict__ rgb, long long npx) {
  const long long p = (long long)blockIdx.x * 256 + threadIdx.x;
  if (p >= npx) return;
  uint32_t R, G, B;
  tmf::ycc_to_rgb8_exact(__ldg(ycc + 3 * p), __ldg(ycc + 3 * p + 1), __ldg(ycc + 3 * p + 2), R, G, B);
  rgb[3 * p] = (uint8_t)R; rgb[3 * p + 1] = (uint8_t)G; rgb[3 * p + 2] = (uint8_t)B;
}

// ---------------------------------------------------------------------------
// Pixel-format taps for the PIL boundary.  PIL keeps an "RGB" image as 4 bytes per pixel
// (R, G, B, pad); packing it to 3 bytes on the host (Image.tobytes) and unpacking the result
// (Image.frombuffer) cost ~25 ms each for a 4K image - three orders of magnitude more than the
// embed kernel.  The single-image API therefore moves PIL's own 4-byte layout over PCIe and
// converts on the device: one thread per 4 pixels, 16 bytes <-> three words.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_rgbx_to_rgb(const uint8_t* __restrict__ rgbx, uint8_t* __restrict__ rgb, long long npx) {
  const long long q = (long long)blockIdx.x * 256 + threadIdx.x;          // group of 4 pixels
  const long long p0 = q * 4;
  if (p0 >= npx) return;
  if (p0 + 4 <= npx) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(rgbx) + q);      // x y z w = pixels 0..3, bytes R G B pad
    uint32_t* o = reinterpret_cast<uint32_t*>(rgb) + q * 3;
    o[0] = (v.x & 0x00ffffffu) | (v.y << 24);                              // R0 G0 B0 R1
    o[1] = ((v.y >> 8) & 0x0000ffffu) | (v.z << 16);                       // G1 B1 R2 G2
    o[2] = ((v.z >> 16) & 0x000000ffu) | (v.w << 8);                       // B2 R3 G3 B3
  } else {
    for (long long p = p0; p < npx; ++p)
      for (int c = 0; c < 3; ++c) rgb[p * 3 + c] = rgbx[p * 4 + c];
  }
}

__global__ void __launch_bounds__(256)
k_rgb_to_rgbx(const uint8_t* __restrict__ rgb, uint8_t* __restrict__ rgbx, long long npx, uint32_t pad) {
  const long long q = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long p0 = q * 4;
  if (p0 >= npx) return;
  if (p0 + 4 <= npx) {
    const uint32_t* i = reinterpret_cast<const uint32_t*>(rgb) + q * 3;
    const uint32_t a = __ldg(i), b = __ldg(i + 1), c = __ldg(i + 2), hi = pad << 24;
    uint4 v;
    v.x = (a & 0x00ffffffu) | hi;
    v.y = (a >> 24) | ((b & 0x0000ffffu) << 8) | hi;
    v.z = (b >> 16) | ((c & 0x000000ffu) << 16) | hi;
    v.w = (c >> 8) | hi;
    reinterpret_cast<uint4*>(rgbx)[q] = v;
  } else {
    for (long long p = p0; p < npx; ++p) {
      for (int ch = 0; ch < 3; ++ch) rgbx[p * 4 + ch] = rgb[p * 3 + ch];
      rgbx[p * 4 + 3] = (uint8_t)pad;
    }
  }
}

// ---------------------------------------------------------------------------
// host-side helpers
// ---------------------------------------------------------------------------
int make_geom(int n, int h, int w, size_t img_stride, int block, BlockGeom& g) {
  if (block < 4 || block > 16 || (block & 1))
    return fail(TMF_ERR_UNSUPPORTED_BLOCK,
                "block_size %d is not supported: this build implements the even sizes 4..16 the reference's UI "
                "offers (8, its BLOCK_SIZE, on the optimised path); there is no CPU fallback", block);
  if (n < 0 || h < 0 || w < 0) return fail(TMF_ERR_BAD_ARG, "negative dimension (n=%d h=%d w=%d)", n, h, w);
  if (n > 0 && img_stride < (size_t)h * w * 3)
    return fail(TMF_ERR_BAD_ARG, "img_stride %zu is smaller than one image (%zu bytes)", img_stride, (size_t)h * w * 3);
  g.h = h; g.w = w; g.bs = block; g.nbh = h / block; g.nbw = w / block;
  g.blocks_per_img = (long long)g.nbh * g.nbw;
  g.total_blocks = g.blocks_per_img * n;
  if (g.total_blocks > 0x7fffffffLL)   // 2^31 blocks of 192 B would be 412 GB of pixels
    return fail(TMF_ERR_BAD_ARG, "batch too large: %lld blocks (limit 2^31 - 1 per call)", g.total_blocks);
  g.div_bpi = make_fastdiv((uint32_t)(g.blocks_per_img > 0 ? g.blocks_per_img : 1));
  g.div_nbw = make_fastdiv((uint32_t)(g.nbw > 0 ? g.nbw : 1));
  g.img_stride = img_stride;
  g.row_pitch = (size_t)w * 3;
  if (g.row_pitch > 0xffffffffull) return fail(TMF_ERR_BAD_ARG, "image rows of %zu bytes are not supported (limit 2^32 - 1)", g.row_pitch);
  g.pitch32 = (uint32_t)g.row_pitch;
  return TMF_OK;
}

int pick_vec(const BlockGeom& g, const void* p0, const void* p1, const void* p2 = nullptr) {
  uintptr_t bits = (uintptr_t)p0 | (uintptr_t)p1 | (uintptr_t)p2 | (uintptr_t)g.img_stride | (uintptr_t)g.row_pitch;
  if ((bits & 7) == 0) return 8;
  if ((bits & 3) == 0) return 4;
  return 1;
}

// The TMA-staged embed kernel (k_embed_fast_tma) is an opt-in alternative: measured 4-5 % slower
// than the per-thread kernel on 1080p batches (it issues better, 72 % vs 66 %
// of the slots, but executes 14 % more instructions), so it runs only when TMF_EMBED_TMA=1 is in
// the environment (read once) - for A/B measurements and for testing the two code paths against
// each other (identical outputs).  It needs 16-byte aligned runs and >= 32 blocks per block-row.
bool tma_ok(const BlockGeom& g, const void* p0, const void* p1, const void* p2 = nullptr) {
  static const bool enabled = [] { const char* e = getenv("TMF_EMBED_TMA"); return e && e[0] == '1'; }();
  if (!enabled) return false;
  uintptr_t bits = (uintptr_t)p0 | (uintptr_t)p1 | (uintptr_t)p2 | (uintptr_t)g.img_stride | (uintptr_t)g.row_pitch;
  return (bits & 15) == 0 && (g.nbw & 1) == 0 && g.nbw >= 32 && g.bs == 8;
}

// opt in to > 48 KB of dynamic shared memory, once per device (idempotent, thread-safe)
int tma_kernel_attrs() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail(TMF_ERR_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
  static std::atomic<unsigned long long> done{0};
  if (dev < 64 && (done.load(std::memory_order_acquire) >> dev) & 1ull) return TMF_OK;
  e = cudaFuncSetAttribute(k_embed_fast_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, kEmbedTmaSmem);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(k_embed_fast_tma, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) { cudaGetLastError(); return fail(TMF_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e)); }
  if (dev < 64) done.fetch_or(1ull << dev, std::memory_order_release);
  return TMF_OK;
}

unsigned grid_for(long long items, int per_cta) { return (unsigned)((items + per_cta - 1) / per_cta); }

// grid of the fast embed kernel: one CTA per 128 blocks, or (TMF_EMBED_PERSIST) a fixed number per SM
unsigned embed_grid(unsigned tiles) {
#if TMF_EMBED_PERSIST
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const unsigned cap = (unsigned)sms * TMF_EMBED_PERSIST;
  return tiles < cap ? tiles : cap;
#else
  return tiles;
#endif
}


// dispatch over the block sizes other than 8
// Dispatch over the block sizes other than 8.  f(N, A) gets integral constants: N = block
// size, A = access width the block rows allow - 4 bytes (N = 4, 12, 16 with 4-byte aligned
// rows), 2 bytes (any even N with 2-byte aligned rows), else single bytes.
template <int N, typename F>
void with_access_width(int al, bool has4, F&& f) {
  if (has4 && al == 4) f(std::integral_constant<int, N>{}, std::integral_constant<int, 4>{});
  else if (al >= 2) f(std::integral_constant<int, N>{}, std::integral_constant<int, 2>{});
  else f(std::integral_constant<int, N>{}, std::integral_constant<int, 1>{});
}
template <typename F>
void for_block_size(int n, int al, F&& f) {
  switch (n) {
    case 4: with_access_width<4>(al, true, f); break;
    case 6: with_access_width<6>(al, false, f); break;
    case 10: with_access_width<10>(al, false, f); break;
    case 12: with_access_width<12>(al, true, f); break;
    case 14: with_access_width<14>(al, false, f); break;
    case 16: with_access_width<16>(al, true, f); break;
    default: break;
  }
}

int row_alignment(const BlockGeom& g, const void* p0, const void* p1, const void* p2 = nullptr) {
  uintptr_t bits = (uintptr_t)p0 | (uintptr_t)p1 | (uintptr_t)p2 | (uintptr_t)g.img_stride | (uintptr_t)g.row_pitch;
  return (bits & 3) == 0 ? 4 : ((bits & 1) == 0 ? 2 : 1);
}


}  // namespace

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
extern "C" {

int tmf_version(void) { return TMF_VERSION; }
const char* tmf_last_error(void) { return g_err; }

int tmf_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) { cudaGetLastError(); return fail(TMF_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e)); }
  return n;
}

int tmf_embed_rgb8(const uint8_t* rgb, uint8_t* out, int n, int h, int w, size_t img_stride, const uint8_t* wm,
                   int wm_shared, double alpha, int block, int mode, void* stream) {
  BlockGeom g;
  if (int rc = make_geom(n, h, w, img_stride, block, g)) return rc;
  if (mode != TMF_MODE_FAITHFUL && mode != TMF_MODE_FAST) return fail(TMF_ERR_BAD_ARG, "unknown mode %d", mode);
  if (n == 0 || h == 0 || w == 0) return TMF_OK;
  if (!rgb || !out) return fail(TMF_ERR_BAD_ARG, "null image pointer");
  if (rgb == out) return fail(TMF_ERR_BAD_ARG, "out must not alias rgb");
  if (g.total_blocks > 0 && !wm) return fail(TMF_ERR_BAD_ARG, "null watermark map");
  if (!(alpha == alpha)) return fail(TMF_ERR_BAD_ARG, "alpha is NaN");
  cudaStream_t st = (cudaStream_t)stream;
  if (g.total_blocks > 0 && block != 8) {
    const unsigned grid = grid_for(g.total_blocks, kThreads);
    for_block_size(block, row_alignment(g, rgb, out), [&](auto n_, auto a_) {
      k_embed_fast_n<decltype(n_)::value, decltype(a_)::value><<<grid, kThreads, 0, st>>>(rgb, out, g, wm, wm_shared, alpha);
    });
    if (int rc = check_launch("embed kernel launch")) return rc;
  } else if (g.total_blocks > 0) {
    const unsigned grid = grid_for(g.total_blocks, kThreads);
    const int vec = pick_vec(g, rgb, out);
    const unsigned egrid = grid_for(g.total_blocks, kEmbedThreads);
    if (mode == TMF_MODE_FAST && tma_ok(g, rgb, out)) {
      if (int rc = tma_kernel_attrs()) return rc;
      k_embed_fast_tma<<<grid, kThreads, kEmbedTmaSmem, st>>>(rgb, out, g, wm, wm_shared, alpha);
    } else if (mode == TMF_MODE_FAST) {
      switch (vec) {
        case 8: k_embed_fast<8><<<embed_grid(egrid), kEmbedThreads, 0, st>>>(rgb, out, g, wm, wm_shared, alpha); break;
        case 4: k_embed_fast<4><<<embed_grid(egrid), kEmbedThreads, 0, st>>>(rgb, out, g, wm, wm_shared, alpha); break;
        default: k_embed_fast<1><<<embed_grid(egrid), kEmbedThreads, 0, st>>>(rgb, out, g, wm, wm_shared, alpha); break;
      }
    } else {
      switch (vec) {
        case 8: k_embed_faithful<8><<<grid, kThreads, 0, st>>>(rgb, out, g, wm, wm_shared, alpha); break;
        case 4: k_embed_faithful<4><<<grid, kThreads, 0, st>>>(rgb, out, g, wm, wm_shared, alpha); break;
        default: k_embed_faithful<1><<<grid, kThreads, 0, st>>>(rgb, out, g, wm, wm_shared, alpha); break;
      }
    }
    if (int rc = check_launch("embed kernel launch")) return rc;
  }
  const long long strip = (long long)h * w - g.blocks_per_img * block * block;
  if (strip > 0) {
    k_strip_roundtrip<<<grid_for(strip * n, 256), 256, 0, st>>>(rgb, out, g, n, strip);
    if (int rc = check_launch("strip kernel launch")) return rc;
  }
  return TMF_OK;
}

int tmf_extract_rgb8(const uint8_t* wmk_rgb, const uint8_t* orig_rgb, uint8_t* out_wm, int n, int h, int w,
                     size_t img_stride, double alpha, int block, int mode, void* stream) {
  BlockGeom g;
  if (int rc = make_geom(n, h, w, img_stride, block, g)) return rc;
  if (mode != TMF_MODE_FAITHFUL && mode != TMF_MODE_FAST) return fail(TMF_ERR_BAD_ARG, "unknown mode %d", mode);
  if (g.total_blocks == 0) return TMF_OK;
  if (!wmk_rgb || !orig_rgb || !out_wm) return fail(TMF_ERR_BAD_ARG, "null pointer");
  if (!(alpha == alpha) || alpha == 0.0) return fail(TMF_ERR_BAD_ARG, "alpha must be a non-zero number");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = grid_for(g.total_blocks, kThreads);
  if (block != 8) {
    for_block_size(block, row_alignment(g, wmk_rgb, orig_rgb), [&](auto n_, auto a_) {
      k_extract_fast_n<decltype(n_)::value, decltype(a_)::value><<<grid, kThreads, 0, st>>>(wmk_rgb, orig_rgb, out_wm, g, alpha);
    });
    return check_launch("extract kernel launch");
  }
  const int vec = pick_vec(g, wmk_rgb, orig_rgb);
  const unsigned xgrid = grid_for(g.total_blocks, kExtractThreads);
  if (mode == TMF_MODE_FAST) {
    switch (vec) {
      case 8: k_extract_fast<8><<<xgrid, kExtractThreads, 0, st>>>(wmk_rgb, orig_rgb, out_wm, g, alpha); break;
      case 4: k_extract_fast<4><<<xgrid, kExtractThreads, 0, st>>>(wmk_rgb, orig_rgb, out_wm, g, alpha); break;
      default: k_extract_fast<1><<<xgrid, kExtractThreads, 0, st>>>(wmk_rgb, orig_rgb, out_wm, g, alpha); break;
    }
  } else {
    switch (vec) {
      case 8: k_extract_faithful<8><<<grid, kThreads, 0, st>>>(wmk_rgb, orig_rgb, out_wm, g, alpha); break;
      case 4: k_extract_faithful<4><<<grid, kThreads, 0, st>>>(wmk_rgb, orig_rgb, out_wm, g, alpha); break;
      default: k_extract_faithful<1><<<grid, kThreads, 0, st>>>(wmk_rgb, orig_rgb, out_wm, g, alpha); break;
    }
  }
  return check_launch("extract kernel launch");
}

int tmf_sigma0_rgb8(const uint8_t* rgb, float* sigma0, int n, int h, int w, size_t img_stride, int block, int mode,
                    void* stream) {
  BlockGeom g;
  if (int rc = make_geom(n, h, w, img_stride, block, g)) return rc;
  if (mode != TMF_MODE_FAITHFUL && mode != TMF_MODE_FAST) return fail(TMF_ERR_BAD_ARG, "unknown mode %d", mode);
  if (g.total_blocks == 0) return TMF_OK;
  if (!rgb || !sigma0) return fail(TMF_ERR_BAD_ARG, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = grid_for(g.total_blocks, kThreads);
  if (block != 8) {
    for_block_size(block, row_alignment(g, rgb, rgb), [&](auto n_, auto a_) {
      k_sigma0_fast_n<decltype(n_)::value, decltype(a_)::value><<<grid, kThreads, 0, st>>>(rgb, sigma0, g);
    });
    return check_launch("sigma0 kernel launch");
  }
  const int vec = pick_vec(g, rgb, rgb);
  if (mode == TMF_MODE_FAST) {
    switch (vec) {
      case 8: k_sigma0_fast<8><<<grid, kThreads, 0, st>>>(rgb, sigma0, g); break;
      case 4: k_sigma0_fast<4><<<grid, kThreads, 0, st>>>(rgb, sigma0, g); break;
      default: k_sigma0_fast<1><<<grid, kThreads, 0, st>>>(rgb, sigma0, g); break;
    }
  } else {
    switch (vec) {
      case 8: k_sigma0_faithful<8><<<grid, kThreads, 0, st>>>(rgb, sigma0, g); break;
      case 4: k_sigma0_faithful<4><<<grid, kThreads, 0, st>>>(rgb, sigma0, g); break;
      default: k_sigma0_faithful<1><<<grid, kThreads, 0, st>>>(rgb, sigma0, g); break;
    }
  }
  return check_launch("sigma0 kernel launch");
}

int tmf_svd8x8_f32(const float* blocks, int64_t nblocks, float* S, float* U, float* Vt, int32_t* sweeps,
                   int complete_u, void* stream) {
  if (nblocks < 0) return fail(TMF_ERR_BAD_ARG, "negative block count");
  if (nblocks == 0) return TMF_OK;
  if (!blocks || !S) return fail(TMF_ERR_BAD_ARG, "null pointer");
  if ((U == nullptr) != (Vt == nullptr)) return fail(TMF_ERR_BAD_ARG, "U and Vt must both be given or both be NULL");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = grid_for(nblocks, kThreads);
  if (U) k_svd8x8<true><<<grid, kThreads, 0, st>>>(blocks, nblocks, S, U, Vt, sweeps, complete_u);
  else k_svd8x8<false><<<grid, kThreads, 0, st>>>(blocks, nblocks, S, nullptr, nullptr, sweeps, 0);
  return check_launch("svd kernel launch");
}

int tmf_dct8x8_f32(const float* in, float* out, int64_t nblocks, int inverse, void* stream) {
  if (nblocks < 0) return fail(TMF_ERR_BAD_ARG, "negative block count");
  if (nblocks == 0) return TMF_OK;
  if (!in || !out) return fail(TMF_ERR_BAD_ARG, "null pointer");
  k_dct8x8<<<grid_for(nblocks, kThreads), kThreads, 0, (cudaStream_t)stream>>>(in, out, nblocks, inverse ? 1 : 0);
  return check_launch("dct kernel launch");
}

int tmf_rgb8_to_ycbcr_f32(const uint8_t* rgb, float* ycc, int64_t npixels, void* stream) {
  if (npixels < 0) return fail(TMF_ERR_BAD_ARG, "negative pixel count");
  if (npixels == 0) return TMF_OK;
  if (!rgb || !ycc) return fail(TMF_ERR_BAD_ARG, "null pointer");
  k_rgb2ycc<<<grid_for(npixels, 256), 256, 0, (cudaStream_t)stream>>>(rgb, ycc, npixels);
  return check_launch("rgb->ycbcr kernel launch");
}

int tmf_rgbx8_to_rgb8(const uint8_t* rgbx, uint8_t* rgb, int64_t npixels, void* stream) {
  if (npixels < 0) return fail(TMF_ERR_BAD_ARG, "negative pixel count");
  if (npixels == 0) return TMF_OK;
  if (!rgbx || !rgb) return fail(TMF_ERR_BAD_ARG, "null pointer");
  if (((uintptr_t)rgbx & 15) || ((uintptr_t)rgb & 3)) return fail(TMF_ERR_BAD_ARG, "rgbx must be 16-byte and rgb 4-byte aligned");
  k_rgbx_to_rgb<<<grid_for((npixels + 3) / 4, 256), 256, 0, (cudaStream_t)stream>>>(rgbx, rgb, npixels);
  return check_launch("rgbx->rgb kernel launch");
}

int tmf_rgb8_to_rgbx8(const uint8_t* rgb, uint8_t* rgbx, int64_t npixels, int pad, void* stream) {
  if (npixels < 0) return fail(TMF_ERR_BAD_ARG, "negative pixel count");
  if (npixels == 0) return TMF_OK;
  if (!rgbx || !rgb) return fail(TMF_ERR_BAD_ARG, "null pointer");
  if (((uintptr_t)rgbx & 15) || ((uintptr_t)rgb & 3)) return fail(TMF_ERR_BAD_ARG, "rgbx must be 16-byte and rgb 4-byte aligned");
  k_rgb_to_rgbx<<<grid_for((npixels + 3) / 4, 256), 256, 0, (cudaStream_t)stream>>>(rgb, rgbx, npixels, (uint32_t)pad & 0xffu);
  return check_launch("rgb->rgbx kernel launch");
}

int tmf_ycbcr_f32_to_rgb8(const float* ycc, uint8_t* rgb, int64_t npixels, void* stream) {
  if (npixels < 0) return fail(TMF_ERR_BAD_ARG, "negative pixel count");
  if (npixels == 0) return TMF_OK;
  if (!ycc || !rgb) return fail(TMF_ERR_BAD_ARG, "null pointer");
  k_ycc2rgb<<<grid_for(npixels, 256), 256, 0, (cudaStream_t)stream>>>(ycc, rgb, npixels);
  return check_launch("ycbcr->rgb kernel launch");
}

// ---------------------------------------------------------------------------
// Watermark map on the device (tmf_resize.cuh): resize_watermark after .convert("L").
// Workspace layout: [bounds_h][kT_h][bounds_v][k_v] int32 tables, then the uint8 image
// between the two passes (n x rows x new_w), every part 16-byte aligned.
// ---------------------------------------------------------------------------
}  // extern "C"
namespace {
struct MapPlan {
  tmf::MapGeometry g;
  bool need_h, need_v;
  int ksize_h, ksize_v;
  size_t off_bh, off_kh, off_bv, off_kv, off_tmp, table_bytes, total_bytes;
};
size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }
int lanczos_ksize(int in_size, int out_size) {
  const double scale = (double)(float)in_size / out_size;
  return (int)ceil(3.0 * (scale < 1.0 ? 1.0 : scale)) * 2 + 1;
}
int plan_map(int n, int src_h, int src_w, int target_h, int target_w, int preserve_ratio, MapPlan& p) {
  if (n < 0 || src_h <= 0 || src_w <= 0 || target_h <= 0 || target_w <= 0)
    return fail(TMF_ERR_BAD_ARG, "watermark map: sizes must be positive (n=%d src=%dx%d target=%dx%d)", n, src_h,
                src_w, target_h, target_w);
  p.g = tmf::watermark_map_geometry(src_h, src_w, target_h, target_w, preserve_ratio);
  if (p.g.new_h <= 0 || p.g.new_w <= 0)   // PIL: ValueError("height and width must be > 0")
    return fail(TMF_ERR_BAD_ARG, "watermark map: height and width must be > 0 (resized %dx%d)", p.g.new_h, p.g.new_w);
  p.need_h = p.g.new_w != src_w;
  p.need_v = p.g.new_h != src_h;
  if (src_h > (long long)src_w * 100 && p.g.new_h < src_h)
    return fail(TMF_ERR_BAD_ARG, "watermark map: sources more than 100x taller than wide are not supported");
  if (p.need_h && (size_t)src_w + 16 > (size_t)tmf::kResizeSmemBytes)
    return fail(TMF_ERR_BAD_ARG, "watermark map: source width %d exceeds %d", src_w, tmf::kResizeSmemBytes - 16);
  p.ksize_h = p.need_h ? lanczos_ksize(src_w, p.g.new_w) : 0;
  p.ksize_v = p.need_v ? lanczos_ksize(src_h, p.g.new_h) : 0;
  size_t o = 0;
  p.off_bh = o; o = align16(o + (size_t)(p.need_h ? 2 * p.g.new_w : 0) * 4);
  p.off_kh = o; o = align16(o + (size_t)p.ksize_h * p.g.new_w * 4);
  p.off_bv = o; o = align16(o + (size_t)(p.need_v ? 2 * p.g.new_h : 0) * 4);
  p.off_kv = o; o = align16(o + (size_t)p.ksize_v * p.g.new_h * 4);
  p.table_bytes = o;
  p.off_tmp = o;
  // upper bound: the vertical pass may need every source row
  o = align16(o + (p.need_h ? (size_t)n * src_h * p.g.new_w : 0));
  p.total_bytes = o;
  return TMF_OK;
}
template <int R>
void launch_rows(const uint8_t* src, size_t src_stride, int src_w, int row0, int rows, uint8_t* tmp, int out_w,
                 const int32_t* bounds, const int32_t* kT, int n, cudaStream_t st) {
  const unsigned gx = (unsigned)((rows + R - 1) / R);
  for (int i0 = 0; i0 < n; i0 += 65535) {   // gridDim.y limit
    const int cnt = n - i0 < 65535 ? n - i0 : 65535;
    tmf::k_resample_rows<R><<<dim3(gx, (unsigned)cnt), tmf::kResizeThreads, (size_t)R * src_w + 16, st>>>(
        src + (size_t)i0 * src_stride, src_stride, src_w, row0, rows, tmp + (size_t)i0 * rows * out_w, out_w, bounds,
        kT);
  }
}
}  // namespace
extern "C" {

size_t tmf_wm_map_workspace_bytes(int n, int src_h, int src_w, int target_h, int target_w, int preserve_ratio) {
  MapPlan p;
  if (plan_map(n, src_h, src_w, target_h, target_w, preserve_ratio, p)) return 0;
  return p.total_bytes > 0 ? p.total_bytes : 16;
}

int tmf_wm_map_l8(const uint8_t* src, int n, int src_h, int src_w, size_t src_stride, uint8_t* maps, int target_h,
                  int target_w, int preserve_ratio, void* workspace, size_t workspace_bytes, void* stream) {
  MapPlan p;
  if (int rc = plan_map(n, src_h, src_w, target_h, target_w, preserve_ratio, p)) return rc;
  if (n == 0) return TMF_OK;
  if (!src || !maps) return fail(TMF_ERR_BAD_ARG, "null pointer");
  if (src_stride < (size_t)src_h * src_w) return fail(TMF_ERR_BAD_ARG, "src_stride smaller than one image");
  if (p.total_bytes > 0 && (!workspace || workspace_bytes < p.total_bytes))
    return fail(TMF_ERR_BAD_ARG, "workspace too small: %zu bytes needed (tmf_wm_map_workspace_bytes)", p.total_bytes);
  if (((uintptr_t)workspace & 15) != 0) return fail(TMF_ERR_BAD_ARG, "workspace must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* ws = static_cast<uint8_t*>(workspace);

  // weight tables on the host (libm sin, as Pillow), one upload
  tmf::AxisTable th, tv;
  int row0 = 0, rows = src_h;
  std::vector<uint8_t> host(p.table_bytes, 0);
  if (p.need_v) {
    tmf::lanczos_axis_table(src_h, p.g.new_h, tv);
    if (p.need_h) {  // ImagingResampleInner: the horizontal pass covers only the rows the vertical one reads
      row0 = tv.bounds[0];
      rows = tv.bounds[2 * (p.g.new_h - 1)] + tv.bounds[2 * (p.g.new_h - 1) + 1] - row0;
      for (int i = 0; i < p.g.new_h; ++i) tv.bounds[2 * i] -= row0;
    }
    memcpy(host.data() + p.off_bv, tv.bounds.data(), tv.bounds.size() * 4);
    memcpy(host.data() + p.off_kv, tv.kk.data(), tv.kk.size() * 4);
  }
  if (p.need_h) {
    tmf::lanczos_axis_table(src_w, p.g.new_w, th);
    memcpy(host.data() + p.off_bh, th.bounds.data(), th.bounds.size() * 4);
    int32_t* kT = reinterpret_cast<int32_t*>(host.data() + p.off_kh);  // tap-major for coalesced loads
    for (int xx = 0; xx < p.g.new_w; ++xx)
      for (int i = 0; i < th.ksize; ++i) kT[(size_t)i * p.g.new_w + xx] = th.kk[(size_t)xx * th.ksize + i];
  }
  if (p.table_bytes > 0) {
    // pageable source: the runtime stages it before returning, so `host` may go out of scope
    cudaError_t e = cudaMemcpyAsync(ws, host.data(), p.table_bytes, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(TMF_ERR_CUDA, "weight table upload: %s", cudaGetErrorString(e)); }
  }
  const int32_t* d_bh = reinterpret_cast<const int32_t*>(ws + p.off_bh);
  const int32_t* d_kh = reinterpret_cast<const int32_t*>(ws + p.off_kh);
  const int32_t* d_bv = reinterpret_cast<const int32_t*>(ws + p.off_bv);
  const int32_t* d_kv = reinterpret_cast<const int32_t*>(ws + p.off_kv);
  uint8_t* tmp = ws + p.off_tmp;

  const uint8_t* in = src;
  size_t in_stride = src_stride;
  int in_pitch = src_w;
  if (p.need_h) {
    const size_t row_bytes = (size_t)src_w;
    if (8 * row_bytes + 16 <= (size_t)tmf::kResizeSmemBytes)
      launch_rows<8>(src, src_stride, src_w, row0, rows, tmp, p.g.new_w, d_bh, d_kh, n, st);
    else if (4 * row_bytes + 16 <= (size_t)tmf::kResizeSmemBytes)
      launch_rows<4>(src, src_stride, src_w, row0, rows, tmp, p.g.new_w, d_bh, d_kh, n, st);
    else if (2 * row_bytes + 16 <= (size_t)tmf::kResizeSmemBytes)
      launch_rows<2>(src, src_stride, src_w, row0, rows, tmp, p.g.new_w, d_bh, d_kh, n, st);
    else
      launch_rows<1>(src, src_stride, src_w, row0, rows, tmp, p.g.new_w, d_bh, d_kh, n, st);
    if (int rc = check_launch("watermark resize (rows) launch")) return rc;
    in = tmp;
    in_stride = (size_t)rows * p.g.new_w;
    in_pitch = p.g.new_w;
  }
  const long long total = (long long)n * target_h * target_w;
  tmf::k_compose_map<<<grid_for(total, 256), 256, 0, st>>>(in, in_stride, in_pitch, maps, target_h, target_w, p.g, d_bv,
                                                         d_kv, p.ksize_v, p.need_v ? 1 : 0, total);
  return check_launch("watermark resize (columns) launch");
}

// ---------------------------------------------------------------------------
// Host-buffer pipeline (C ABI).  A context owns three streams and `depth` device
// slots on one device; a batch of HOST images is cut into chunks and each chunk
// flows  H2D copy -> fused kernel -> D2H copy  with the three stages of different
// chunks overlapping.  Sharding over GPUs is by image: one context per device,
// enqueue on all of them, then synchronise each - no collective, no peer traffic.
// Host buffers should be page-locked (tmf_pin_host) for the copies to be truly
// asynchronous; pageable buffers work, staged by the driver.
// ---------------------------------------------------------------------------
struct tmf_ctx {
  int device;
  int depth;
  size_t chunk_bytes;
  cudaStream_t s_in, s_run, s_out;
  struct Slot {
    uint8_t *a, *b, *o, *wm;
    size_t cap_a, cap_b, cap_o, cap_wm;
    cudaEvent_t ev_in, ev_run, ev_out;
    bool used;
  } slot[TMF_CTX_MAX_DEPTH];
  uint8_t* wm_shared;
  size_t cap_wm_shared;
  long long launches, h2d_bytes, d2h_bytes;
};

namespace {

struct DeviceGuard {
  int prev;
  bool ok;
  explicit DeviceGuard(int dev) : prev(-1), ok(false) {
    if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); return; }
    ok = (cudaSetDevice(dev) == cudaSuccess);
    if (!ok) cudaGetLastError();
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int cuda_fail(const char* what, cudaError_t e) {
  cudaGetLastError();
  return fail(TMF_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

int grow(uint8_t** p, size_t* cap, size_t need) {
  if (*cap >= need) return TMF_OK;
  if (*p) cudaFree(*p);
  *p = nullptr; *cap = 0;
  cudaError_t e = cudaMalloc((void**)p, need);
  if (e != cudaSuccess) return cuda_fail("cudaMalloc", e);
  *cap = need;
  return TMF_OK;
}

#define TMF_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_fail(#call, e_); } while (0)

}  // namespace

int tmf_ctx_create(tmf_ctx** out, int device, size_t chunk_bytes, int depth) {
  if (!out) return fail(TMF_ERR_BAD_ARG, "null context pointer");
  *out = nullptr;
  if (depth < 1 || depth > TMF_CTX_MAX_DEPTH) return fail(TMF_ERR_BAD_ARG, "depth must be 1..%d", TMF_CTX_MAX_DEPTH);
  if (chunk_bytes == 0) chunk_bytes = (size_t)96 << 20;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess) return cuda_fail("cudaGetDeviceCount", e);
  if (device < 0 || device >= ndev) return fail(TMF_ERR_BAD_ARG, "device %d out of range (%d visible)", device, ndev);
  DeviceGuard guard(device);
  if (!guard.ok) return fail(TMF_ERR_CUDA, "cannot select device %d", device);
  tmf_ctx* c = new tmf_ctx();
  memset(c, 0, sizeof *c);
  c->device = device; c->depth = depth; c->chunk_bytes = chunk_bytes;
  auto init = [&]() -> int {
    TMF_CUDA(cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking));
    TMF_CUDA(cudaStreamCreateWithFlags(&c->s_run, cudaStreamNonBlocking));
    TMF_CUDA(cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking));
    for (int k = 0; k < depth; ++k) {
      TMF_CUDA(cudaEventCreateWithFlags(&c->slot[k].ev_in, cudaEventDisableTiming));
      TMF_CUDA(cudaEventCreateWithFlags(&c->slot[k].ev_run, cudaEventDisableTiming));
      TMF_CUDA(cudaEventCreateWithFlags(&c->slot[k].ev_out, cudaEventDisableTiming));
    }
    return TMF_OK;
  };
  if (int rc = init()) { delete c; return rc; }    // (streams/events of a half-built context die with the process)
  *out = c;
  return TMF_OK;
}

int tmf_ctx_destroy(tmf_ctx* c) {
  if (!c) return TMF_OK;
  DeviceGuard guard(c->device);
  cudaStreamSynchronize(c->s_in); cudaStreamSynchronize(c->s_run); cudaStreamSynchronize(c->s_out);
  for (int k = 0; k < c->depth; ++k) {
    cudaFree(c->slot[k].a); cudaFree(c->slot[k].b); cudaFree(c->slot[k].o); cudaFree(c->slot[k].wm);
    cudaEventDestroy(c->slot[k].ev_in); cudaEventDestroy(c->slot[k].ev_run); cudaEventDestroy(c->slot[k].ev_out);
  }
  cudaFree(c->wm_shared);
  cudaStreamDestroy(c->s_in); cudaStreamDestroy(c->s_run); cudaStreamDestroy(c->s_out);
  cudaGetLastError();
  delete c;
  return TMF_OK;
}

// kind 0 = embed (a = images, b unused), kind 1 = extract (a = watermarked, b = originals)
static int ctx_enqueue(tmf_ctx* c, int kind, const uint8_t* a, const uint8_t* b, uint8_t* out, int n, int h, int w,
                       const uint8_t* wm, int wm_shared, double alpha, int block, int mode) {
  if (!c) return fail(TMF_ERR_BAD_ARG, "null context");
  BlockGeom g;
  const size_t img = (size_t)h * w * 3;
  if (int rc = make_geom(n, h, w, img, block, g)) return rc;
  if (mode != TMF_MODE_FAITHFUL && mode != TMF_MODE_FAST) return fail(TMF_ERR_BAD_ARG, "unknown mode %d", mode);
  if (n == 0 || img == 0) return TMF_OK;
  if (!a || !out || (kind == 1 && !b)) return fail(TMF_ERR_BAD_ARG, "null host pointer");
  const size_t map = (size_t)g.blocks_per_img;
  if (kind == 0 && map > 0 && !wm) return fail(TMF_ERR_BAD_ARG, "null watermark map");
  if (kind == 1 && map == 0) return TMF_OK;
  DeviceGuard guard(c->device);
  if (!guard.ok) return fail(TMF_ERR_CUDA, "cannot select device %d", c->device);
  const size_t out_per_img = (kind == 0) ? img : map;
  size_t per = c->chunk_bytes / img;
  if (per < 1) per = 1;
  if (per > (size_t)n) per = (size_t)n;
  if (kind == 0 && wm_shared && map > 0) {
    // the previous call may still be reading the shared map: order the overwrite after it
    TMF_CUDA(cudaStreamSynchronize(c->s_run));
    if (int rc = grow(&c->wm_shared, &c->cap_wm_shared, map)) return rc;
    TMF_CUDA(cudaMemcpyAsync(c->wm_shared, wm, map, cudaMemcpyHostToDevice, c->s_in));
    c->h2d_bytes += (long long)map;
  }
  int step = 0;
  for (size_t lo = 0; lo < (size_t)n; lo += per, ++step) {
    const size_t k = ((size_t)n - lo < per) ? (size_t)n - lo : per;
    tmf_ctx::Slot& sl = c->slot[step % c->depth];
    // H2D (waits until the kernel that last read this slot's inputs has finished)
    if (sl.used) TMF_CUDA(cudaStreamWaitEvent(c->s_in, sl.ev_run, 0));
    if (int rc = grow(&sl.a, &sl.cap_a, per * img)) return rc;
    TMF_CUDA(cudaMemcpyAsync(sl.a, a + lo * img, k * img, cudaMemcpyHostToDevice, c->s_in));
    c->h2d_bytes += (long long)(k * img);
    if (kind == 1) {
      if (int rc = grow(&sl.b, &sl.cap_b, per * img)) return rc;
      TMF_CUDA(cudaMemcpyAsync(sl.b, b + lo * img, k * img, cudaMemcpyHostToDevice, c->s_in));
      c->h2d_bytes += (long long)(k * img);
    } else if (!wm_shared && map > 0) {
      if (int rc = grow(&sl.wm, &sl.cap_wm, per * map)) return rc;
      TMF_CUDA(cudaMemcpyAsync(sl.wm, wm + lo * map, k * map, cudaMemcpyHostToDevice, c->s_in));
      c->h2d_bytes += (long long)(k * map);
    }
    TMF_CUDA(cudaEventRecord(sl.ev_in, c->s_in));
    // kernel (waits for the inputs, and for the D2H that last read this slot's output)
    TMF_CUDA(cudaStreamWaitEvent(c->s_run, sl.ev_in, 0));
    if (sl.used) TMF_CUDA(cudaStreamWaitEvent(c->s_run, sl.ev_out, 0));
    if (int rc = grow(&sl.o, &sl.cap_o, per * out_per_img)) return rc;
    int rc = (kind == 0)
        ? tmf_embed_rgb8(sl.a, sl.o, (int)k, h, w, img, wm_shared ? c->wm_shared : sl.wm, wm_shared, alpha, block, mode, c->s_run)
        : tmf_extract_rgb8(sl.a, sl.b, sl.o, (int)k, h, w, img, alpha, block, mode, c->s_run);
    if (rc) return rc;
    c->launches += 1;
    TMF_CUDA(cudaEventRecord(sl.ev_run, c->s_run));
    // D2H
    TMF_CUDA(cudaStreamWaitEvent(c->s_out, sl.ev_run, 0));
    TMF_CUDA(cudaMemcpyAsync(out + lo * out_per_img, sl.o, k * out_per_img, cudaMemcpyDeviceToHost, c->s_out));
    c->d2h_bytes += (long long)(k * out_per_img);
    TMF_CUDA(cudaEventRecord(sl.ev_out, c->s_out));
    sl.used = true;
  }
  return TMF_OK;
}

int tmf_ctx_embed_host_async(tmf_ctx* c, const uint8_t* rgb, uint8_t* out, int n, int h, int w, const uint8_t* wm,
                             int wm_shared, double alpha, int block, int mode) {
  return ctx_enqueue(c, 0, rgb, nullptr, out, n, h, w, wm, wm_shared, alpha, block, mode);
}

int tmf_ctx_extract_host_async(tmf_ctx* c, const uint8_t* wmk_rgb, const uint8_t* orig_rgb, uint8_t* out_wm, int n,
                               int h, int w, double alpha, int block, int mode) {
  return ctx_enqueue(c, 1, wmk_rgb, orig_rgb, out_wm, n, h, w, nullptr, 0, alpha, block, mode);
}

int tmf_ctx_synchronize(tmf_ctx* c) {
  if (!c) return fail(TMF_ERR_BAD_ARG, "null context");
  DeviceGuard guard(c->device);
  TMF_CUDA(cudaStreamSynchronize(c->s_in));
  TMF_CUDA(cudaStreamSynchronize(c->s_run));
  TMF_CUDA(cudaStreamSynchronize(c->s_out));
  return TMF_OK;
}

int tmf_ctx_stats(tmf_ctx* c, long long* launches, long long* h2d_bytes, long long* d2h_bytes, int reset) {
  if (!c) return fail(TMF_ERR_BAD_ARG, "null context");
  if (launches) *launches = c->launches;
  if (h2d_bytes) *h2d_bytes = c->h2d_bytes;
  if (d2h_bytes) *d2h_bytes = c->d2h_bytes;
  if (reset) c->launches = c->h2d_bytes = c->d2h_bytes = 0;
  return TMF_OK;
}

int tmf_pin_host(void* p, size_t bytes) {
  if (!p || bytes == 0) return fail(TMF_ERR_BAD_ARG, "null pointer or zero size");
  TMF_CUDA(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
  return TMF_OK;
}

int tmf_unpin_host(void* p) {
  if (!p) return fail(TMF_ERR_BAD_ARG, "null pointer");
  TMF_CUDA(cudaHostUnregister(p));
  return TMF_OK;
}

}  // extern "C"
