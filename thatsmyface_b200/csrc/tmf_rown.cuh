// Row arithmetic of the FAST kernels for the block sizes other than 8 (fast_n_kernels.cu: one thread per block).
//
// A block row is 3N bytes = 3N/2 halfwords, held in ceil(3N/4) 32-bit words (w[k] = bytes 4k..4k+3 of the
// row).  As in the block-8 kernels the luma is the exact integer 299 r + 587 g + 114 b, computed by two IDP.2A
// per pixel on the packed words (no byte extraction), and pass 2 floors straight onto the integer level with the
// subnormal quantiser (tmf_rowmath.cuh).
#pragma once
#include "tmf_common.cuh"
#include "tmf_rowmath.cuh"

namespace tmfi {

// exact integer luma of pixel j as the magic float 2^23 + (299 r + 587 g + 114 b): two IDP.2A
template <int NW>
__device__ __forceinline__ float pixel_luma_magic_n(const uint32_t (&w)[NW], int j) {
  constexpr uint32_t kRG = 299u | (587u << 16), kB_ = 114u, k_R = 299u << 16, kGB = 587u | (114u << 16);
  const int k = (3 * j) >> 2;
  uint32_t m;
  switch ((3 * j) & 3) {
    case 0: m = dp2a_hi(w[k], kB_, dp2a_lo(w[k], kRG, 0x4B000000u)); break;
    case 1: m = dp2a_hi(w[k], kGB, dp2a_lo(w[k], k_R, 0x4B000000u)); break;
    case 2: m = dp2a_lo(w[k + 1 < NW ? k + 1 : k], kB_, dp2a_hi(w[k], kRG, 0x4B000000u)); break;
    default: m = dp2a_lo(w[k + 1 < NW ? k + 1 : k], kGB, dp2a_hi(w[k], k_R, 0x4B000000u)); break;
  }
  return __uint_as_float(m);
}

template <int N>
__device__ __forceinline__ void row_luma_n(const uint32_t (&w)[kRowWords<N>], float (&y)[N]) {
#pragma unroll
  for (int p = 0; p < N / 2; ++p) {
    const float2 v = __fadd2_rn(make_float2(pixel_luma_magic_n<kRowWords<N>>(w, 2 * p),
                                            pixel_luma_magic_n<kRowWords<N>>(w, 2 * p + 1)), bc2(-8388608.0f));
    y[2 * p] = v.x; y[2 * p + 1] = v.y;
  }
}

// the same as pairs (y_2p, y_2p+1), as the packed Gram update below wants them
template <int N>
__device__ __forceinline__ void row_luma2_n(const uint32_t (&w)[kRowWords<N>], float2 (&y2)[N / 2]) {
#pragma unroll
  for (int p = 0; p < N / 2; ++p)
    y2[p] = __fadd2_rn(make_float2(pixel_luma_magic_n<kRowWords<N>>(w, 2 * p),
                                   pixel_luma_magic_n<kRowWords<N>>(w, 2 * p + 1)), bc2(-8388608.0f));
}

// Pass 1 in packed fp32 (GramPairs of tmf_rowmath.cuh for any even N): row i of the upper triangle as pairs
// (G[i][2p], G[i][2p+1]) for 2p >= i, one FFMA2 each; the diagonal of an odd i has no partner and stays scalar.
// N (N + 2) / 4 FFMA2 + N / 2 FFMA per row instead of N (N + 1) / 2 FFMA.  Only the entries named here exist.
template <int N>
struct GramPairsN {
  float2 gp[N][N / 2];
  float gd[N];
};
template <int N>
__device__ __forceinline__ void gram_clear_n(GramPairsN<N>& G) {
#pragma unroll
  for (int i = 0; i < N; ++i) {
    if (i & 1) G.gd[i] = 0.0f;
#pragma unroll
    for (int p = (i + 1) >> 1; p < N / 2; ++p) G.gp[i][p] = make_float2(0.0f, 0.0f);
  }
}
template <int N>
__device__ __forceinline__ void gram_accumulate_row2_n(const float2 (&y2)[N / 2], GramPairsN<N>& G) {
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const float yi = (i & 1) ? y2[i >> 1].y : y2[i >> 1].x;
    if (i & 1) G.gd[i] = fmaf(yi, yi, G.gd[i]);
#pragma unroll
    for (int p = (i + 1) >> 1; p < N / 2; ++p) G.gp[i][p] = __ffma2_rn(bc2(yi), y2[p], G.gp[i][p]);
  }
}
template <int N>
__device__ __forceinline__ void gram_pairs_to_sym_n(const GramPairsN<N>& G, float* gm) {
#pragma unroll
  for (int i = 0; i < N; ++i) {
    if (i & 1) gm[tmf::sym_idx<N>(i, i)] = G.gd[i];
#pragma unroll
    for (int p = (i + 1) >> 1; p < N / 2; ++p) {
      gm[tmf::sym_idx<N>(i, 2 * p)] = G.gp[i][p].x;
      gm[tmf::sym_idx<N>(i, 2 * p + 1)] = G.gp[i][p].y;
    }
  }
}

// byte B of the row as the subnormal float B * 2^-149 (bit pattern = the byte)
template <int NW>
__device__ __forceinline__ float byte_subnormal_n(const uint32_t (&w)[NW], int B) {
  const uint32_t x = w[B >> 2];
  const int b = B & 3;
  uint32_t m;
  if (b == 3) asm("mad.hi.u32 %0, %1, 256, %2;" : "=r"(m) : "r"(x), "n"(0));
  else m = __byte_perm(x, 0u, 0x7650u | (uint32_t)b);
  return __uint_as_float(m);
}

// pass 2 for one row: same arithmetic as embed_row_fast2 (tmf_rowmath.cuh) for any even N
template <int N>
__device__ __forceinline__ void embed_row_n(const uint32_t (&w)[kRowWords<N>], const float (&wv)[N], float f, float c,
                                            bool marked, uint32_t (&o)[kRowWords<N>]) {
  constexpr int NW = kRowWords<N>;
  float du = c;
  if (marked) {
    // z = row . w with two packed accumulators (four independent chains of N / 4 instead of one of N: with 2-3 warps
    // per scheduler the chain's latency is exposed).  The order differs from tmf::dotn by rounding only, in a term
    // that is scaled by f ~ 1e-3 - far below the quantiser's resolution, as in the block-8 kernels.
    float2 y2[N / 2];
    row_luma2_n<N>(w, y2);
    float2 a0 = __fmul2_rn(y2[0], make_float2(wv[0], wv[1]));
    float2 a1 = __fmul2_rn(y2[1], make_float2(wv[2], wv[3]));
#pragma unroll
    for (int p = 2; p < N / 2; ++p) {
      if (p & 1) a1 = __ffma2_rn(y2[p], make_float2(wv[2 * p], wv[2 * p + 1]), a1);
      else a0 = __ffma2_rn(y2[p], make_float2(wv[2 * p], wv[2 * p + 1]), a0);
    }
    du = fmaf(f, (a0.x + a0.y) + (a1.x + a1.y), c);
  }
  du *= 1.7763568394002505e-15f;   // 2^-49
  int q[4 * NW];
#pragma unroll
  for (int k = 3 * N; k < 4 * NW; ++k) q[k] = 0;
  constexpr float k2p100 = 1.2676506002282294e30f, k2m100 = 7.888609052210118e-31f;
#pragma unroll
  for (int p = 0; p < N / 2; ++p) {
    const int B = 6 * p;
    const float2 d2 = __fmul2_rn(bc2(du), make_float2(wv[2 * p], wv[2 * p + 1]));
    const float2 mr = make_float2(byte_subnormal_n<NW>(w, B), byte_subnormal_n<NW>(w, B + 3));
    const float2 mg = make_float2(byte_subnormal_n<NW>(w, B + 1), byte_subnormal_n<NW>(w, B + 4));
    const float2 mb = make_float2(byte_subnormal_n<NW>(w, B + 2), byte_subnormal_n<NW>(w, B + 5));
    const float2 u = __ffma2_rn(mg, bc2(-1.0f), mr), v = __ffma2_rn(mg, bc2(-1.0f), mb);
    const float2 sR = __ffma2_rn(bc2(5.00e-4f * k2p100), u, __ffma2_rn(bc2(3.57e-4f * k2p100), v, d2));
    const float2 sG = __ffma2_rn(bc2(1.36e-4f * k2p100), u, __ffma2_rn(bc2(-1.66e-4f * k2p100), v, d2));
    const float2 sB = __ffma2_rn(bc2(-6.37e-4f * k2p100), u, __ffma2_rn(bc2(5.00e-4f * k2p100), v, d2));
    const float2 tR = __ffma2_rd(sR, bc2(k2m100), mr), tG = __ffma2_rd(sG, bc2(k2m100), mg),
                 tB = __ffma2_rd(sB, bc2(k2m100), mb);
    q[B] = __float_as_int(tR.x); q[B + 1] = __float_as_int(tG.x); q[B + 2] = __float_as_int(tB.x);
    q[B + 3] = __float_as_int(tR.y); q[B + 4] = __float_as_int(tG.y); q[B + 5] = __float_as_int(tB.y);
  }
#pragma unroll
  for (int k = 0; k < NW; ++k) o[k] = tmf::pack4_sat_u8(q[4 * k], q[4 * k + 1], q[4 * k + 2], q[4 * k + 3]);
}

// threads per CTA: one-warp CTAs for the register-heavy sizes (10 ... 16: +1-3 %), 128 for 4 and 6 (one-warp CTAs
// cost size 4 8 % of its extract rate)
template <int N> __host__ __device__ constexpr int fastn_threads() { return N <= 6 ? 128 : TMF_FASTN_THREADS; }
template <int N> __host__ __device__ constexpr int fastn_min_ctas() {
  return N <= 6 ? TMF_FASTN_CTAS_SMALL : (N <= 10 ? TMF_FASTN_CTAS_10 : (N <= 12 ? TMF_FASTN_CTAS_12 : (N <= 14 ? TMF_FASTN_CTAS_14 : TMF_FASTN_CTAS_16)));
}

}  // namespace tmfi
