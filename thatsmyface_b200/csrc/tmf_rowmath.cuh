// Packed-fp32 row arithmetic of the FAST block-8 kernels (device only): what one thread does
// with one 24-byte row of its block in pass 1 (exact integer luma, Gram update) and in pass 2
// (rank-1 update, colour out, quantise, pack).  Shared by the per-thread kernels and the
// TMA-tiled kernels of fast_kernels.cu; the scalar statement of the same formulas is in
// tmf_fast.cuh (what tests/hostsim runs on the CPU).
//
// Packed fp32 (sm_100 FFMA2 / FADD2 / FMUL2): two adjacent pixels ride in one 64-bit register
// pair.  Each lane of a packed instruction is an ordinary round-to-nearest fp32 operation, so
// results are bit-identical to the scalar formulas; what changes is the dispatch cost
// (profiles/r01_ubench.txt: FFMA2 60/clk/SM = 120 FMA, against 112 for scalar FFMA) and the
// instruction count.  Scalars broadcast and immediates come for free (FFMA2 Rd, Ra.F32, ...).
#pragma once
#include "tmf_fast.cuh"

namespace tmfi {

__device__ __forceinline__ float2 bc2(float x) { return make_float2(x, x); }

// Byte B of the 24-byte row as the float whose BIT PATTERN is MAGIC | byte:
//   MAGIC = 0x4B000000: the float 2^23 + k (k exact in the mantissa);
//   MAGIC = 0:          the subnormal float k * 2^-149 (the quantiser's trick, embed_row_fast2).
// Bytes 0-2 take one PRMT (ALU pipe); byte 3 takes one IMAD.HI / LEA.HI - hi32(x * 256) + MAGIC
// = (x >> 24) + MAGIC - which moves a quarter of the extractions to the other pipe (ncu showed
// the ALU pipe as the busier one; profiles/r01_sweep_variants.txt).
template <uint32_t MAGIC = 0x4B000000u>
__device__ __forceinline__ float byte_to_magic(const uint32_t (&w)[6], int B) {
  const uint32_t x = w[B >> 2];
  const int b = B & 3;
  uint32_t m;
  if (b == 3) asm("mad.hi.u32 %0, %1, 256, %2;" : "=r"(m) : "r"(x), "n"(MAGIC));
  else m = __byte_perm(x, MAGIC, 0x7650u | (uint32_t)b);
  return __uint_as_float(m);
}

// Exact integer luma 299 r + 587 g + 114 b (tmf::luma1000_exact) of the 8 pixels of a row
// without extracting a single byte: a pixel's three bytes sit in one or two of the row's six
// words, and two IDP.2A (16-bit weights x bytes 0-1 or 2-3 of a word, accumulate) cover them
// whatever the phase.  The accumulator starts at 0x4B000000, so the result already is the
// bit pattern of the float 2^23 + luma (luma < 2^18), and one packed FADD per pixel pair
// removes the 2^23.
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
template <uint32_t ACC0 = 0x4B000000u>
__device__ __forceinline__ uint32_t pixel_luma_bits(const uint32_t (&w)[6], int j) {
  constexpr uint32_t kRG = 299u | (587u << 16), kB_ = 114u, k_R = 299u << 16, kGB = 587u | (114u << 16);
  const int k = (3 * j) >> 2;
  uint32_t m;
  switch ((3 * j) & 3) {
    case 0: m = dp2a_hi(w[k], kB_, dp2a_lo(w[k], kRG, ACC0)); break;          // [r g b .]
    case 1: m = dp2a_hi(w[k], kGB, dp2a_lo(w[k], k_R, ACC0)); break;          // [. r g b]
    case 2: m = dp2a_lo(w[k + 1], kB_, dp2a_hi(w[k], kRG, ACC0)); break;      // [. . r g][b . . .]
    default: m = dp2a_lo(w[k + 1], kGB, dp2a_hi(w[k], k_R, ACC0)); break;     // [. . . r][g b . .]
  }
  return m;
}
__device__ __forceinline__ float pixel_luma_magic(const uint32_t (&w)[6], int j) {
  return __uint_as_float(pixel_luma_bits(w, j));
}
#define TMF_LUMA_UNIT TMF_LUMA1000_UNIT

// luma of the 8 pixels of a row as four pairs; same values as tmf::luma1000_exact.  CONVERT: the integer goes
// through I2F (the conversion pipe, idle in the FAST kernels) instead of the magic-number FADD2 (the FMA pipe) -
// for a kernel bound by the FMA pipe rather than by issue slots.
template <bool CONVERT = false>
__device__ __forceinline__ void row_luma2(const uint32_t (&w)[6], float2 (&y2)[4]) {
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    if (CONVERT)
      y2[p] = make_float2(__uint2float_rn(pixel_luma_bits<0u>(w, 2 * p)), __uint2float_rn(pixel_luma_bits<0u>(w, 2 * p + 1)));
    else
      y2[p] = __fadd2_rn(make_float2(pixel_luma_magic(w, 2 * p), pixel_luma_magic(w, 2 * p + 1)), bc2(-8388608.0f));
  }
}

// Gram matrix in paired form: gp[i][p] = (G[i][2p], G[i][2p+1]) for the pairs of the
// upper triangle that start at an even column, gd[i] = G[i][i] for odd i.
struct GramPairs {
  float2 gp[8][4];
  float gd[8];
};

__device__ __forceinline__ void gram_clear(GramPairs& G) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    G.gd[i] = 0.0f;
#pragma unroll
    for (int p = 0; p < 4; ++p) G.gp[i][p] = make_float2(0.0f, 0.0f);
  }
}

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

// Pass 2 for one row, packed: bytes of the row, its luma (4 pairs), w (4 pairs), f * 2^-49, c * 2^-49
// -> six output words.  Same arithmetic as tmf::embed_row_fast + pack4_sat_u8.
//
// The reference maps (r, g, b) to (y, cb, cr), adds the mark to y and maps back with a matrix
// that is not the exact inverse (watermarking.py:37-39, :61): a pixel comes back as
// M k + d with M = Ti T = I + E, E's rows summing to zero, so
//     out_c = floor(k_c + s_c),   s_c = E_c0 (r - g) + E_c2 (b - g) + d     (|s_c| small).
// The floor lands DIRECTLY on the integer level: byte k extracted with a zero filler IS the bit
// pattern of the subnormal float k * 2^-149, and FFMA.RM(s_c * 2^-49, 2^-100, k * 2^-149) is the
// exact sum (k_c + s_c) * 2^-149 rounded toward -inf to a multiple of 2^-149, i.e. the float
// whose bits are floor(k_c + s_c) - or a negative float (sign bit set = a hugely negative s32)
// when the level is below 0, which the saturating pack clips to 0 like any other negative.  No
// bias to remove.  The 2^-49 rides in E's constants (scaled by 2^100: the differences u, v are
// subnormal too, (r - g) * 2^-149, exact) and in du; power-of-two scalings are exact.
// Subnormals cost nothing extra in FFMA on this hardware; the build must never use -ftz.
__device__ __forceinline__ void embed_row_fast2(const uint32_t (&w)[6], const float2 (&y2)[4], const float2 (&w2)[4],
                                                float f49, float c49, uint32_t (&o)[6]) {
  float2 acc = __fmul2_rn(y2[0], w2[0]);
#pragma unroll
  for (int p = 1; p < 4; ++p) acc = __ffma2_rn(y2[p], w2[p], acc);
  // tmf::dot8 accumulates sequentially; the pairwise order differs by rounding only in z,
  // which is scaled by f ~ 1e-3: far below the quantiser's resolution.
  // f49 = f * 2^-49, c49 = c * 2^-49 (scaled once per block by the caller; exact, so du is the same
  // float as fma(f, z, c) * 2^-49).  A block with a zero mark has f49 = c49 = 0 and w2 = 0: its y2 may be
  // anything finite (the tile kernel reads stale stash values) and du is exactly 0.
  const float du = fmaf(f49, acc.x + acc.y, c49);
  int q[24];
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int B = 6 * p;
    const float2 d2 = __fmul2_rn(bc2(du), w2[p]);
    const float2 mr = make_float2(byte_to_magic<0u>(w, B), byte_to_magic<0u>(w, B + 3));
    const float2 mg = make_float2(byte_to_magic<0u>(w, B + 1), byte_to_magic<0u>(w, B + 4));
    const float2 mb = make_float2(byte_to_magic<0u>(w, B + 2), byte_to_magic<0u>(w, B + 5));
    const float2 u = __ffma2_rn(mg, bc2(-1.0f), mr), v = __ffma2_rn(mg, bc2(-1.0f), mb);
    constexpr float k2p100 = 1.2676506002282294e30f;      // 2^100
    const float2 sR = __ffma2_rn(bc2(5.00e-4f * k2p100), u, __ffma2_rn(bc2(3.57e-4f * k2p100), v, d2));
    const float2 sG = __ffma2_rn(bc2(1.36e-4f * k2p100), u, __ffma2_rn(bc2(-1.66e-4f * k2p100), v, d2));
    const float2 sB = __ffma2_rn(bc2(-6.37e-4f * k2p100), u, __ffma2_rn(bc2(5.00e-4f * k2p100), v, d2));
    constexpr float k2m100 = 7.888609052210118e-31f;      // 2^-100
    const float2 tR = __ffma2_rd(sR, bc2(k2m100), mr), tG = __ffma2_rd(sG, bc2(k2m100), mg),
                 tB = __ffma2_rd(sB, bc2(k2m100), mb);
    q[B] = __float_as_int(tR.x); q[B + 1] = __float_as_int(tG.x); q[B + 2] = __float_as_int(tB.x);
    q[B + 3] = __float_as_int(tR.y); q[B + 4] = __float_as_int(tG.y); q[B + 5] = __float_as_int(tB.y);
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) o[k] = tmf::pack4_sat_u8(q[4 * k], q[4 * k + 1], q[4 * k + 2], q[4 * k + 3]);
}

}  // namespace tmfi
