// Taps and standalone kernels of libtmfwm: batched 8x8 SVD (BASELINE config 5), DCT, the
// bit-exact colour transforms, PIL's 4-byte pixel format, and the colour round trip of the
// pixels outside whole blocks.
#include "tmf_common.cuh"
#include "tmf_math.cuh"

namespace tmfi {
namespace {

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

}  // namespace

int launch_strip_roundtrip(const uint8_t* rgb, uint8_t* out, const BlockGeom& g, int n, cudaStream_t st) {
  const long long strip = (long long)g.h * g.w - g.blocks_per_img * g.bs * g.bs;
  if (strip <= 0 || n <= 0) return TMF_OK;
  k_strip_roundtrip<<<grid_for(strip * n, 256), 256, 0, st>>>(rgb, out, g, n, strip);
  return check_launch("strip kernel launch");
}

}  // namespace tmfi

using namespace tmfi;

extern "C" {

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

}  // extern "C"
