// Per-block arithmetic of the DCT+SVD watermark path, written so that one
// thread owns one 8x8 block entirely in registers (every array index below is
// a compile-time constant after unrolling).
//
// Reference semantics (Rigelyon/ThatsMyFace, modules/watermarking.py):
//   colour in   :23-50    colour out :53-73    DCT :76-78    IDCT :81-83
//   SVD         :195, :279-282 (numpy.linalg.svd -> LAPACK sgesdd)
//   modulation  :198      inverse SVD :201     extract epilogue :285-289
//
// The functions are __host__ __device__ so the same source can be compiled for
// the host by tests/hostsim (a CPU *test harness* for the kernel arithmetic;
// the product library never takes that path).
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define TMF_HD __host__ __device__ __forceinline__
#else
#define TMF_HD inline
#endif

namespace tmf {

// ---------------------------------------------------------------------------
// rounding-explicit primitives (device intrinsics; plain C on the host, where
// the test harness is built with -ffp-contract=off)
// ---------------------------------------------------------------------------
TMF_HD float f_div(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fdiv_rn(a, b);
#else
  return a / b;
#endif
}
TMF_HD float f_add(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  volatile float r = a + b; return r;
#endif
}
TMF_HD float f_mul(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  volatile float r = a * b; return r;
#endif
}
TMF_HD double d_mul(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);
#else
  volatile double r = a * b; return r;
#endif
}
TMF_HD double d_add(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(a, b);
#else
  volatile double r = a + b; return r;
#endif
}
TMF_HD float f_sqrt(float x) {
#if defined(__CUDA_ARCH__)
  return __fsqrt_rn(x);
#else
  return sqrtf(x);
#endif
}
TMF_HD double d_fma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
  return __fma_rn(a, b, c);
#else
  return fma(a, b, c);
#endif
}
TMF_HD float f_rsqrt(float x) {   // 1/sqrt(x), a few ulp; refined by the caller where it matters
#if defined(__CUDA_ARCH__)
  return rsqrtf(x);
#else
  return 1.0f / sqrtf(x);
#endif
}
TMF_HD float f_rsqrt_fast(float x) {   // for x >= 2^-126 (no subnormal fix-up code around the MUFU)
#if defined(__CUDA_ARCH__)
  float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
#else
  return 1.0f / sqrtf(x);
#endif
}
TMF_HD float f_sqrt_fast(float x) {
#if defined(__CUDA_ARCH__)
  float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
#else
  return sqrtf(x);
#endif
}
TMF_HD float f_rcp_fast(float x) {
#if defined(__CUDA_ARCH__)
  float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
#else
  return 1.0f / x;
#endif
}

// ---------------------------------------------------------------------------
// colour transforms, bit-exact with the reference
//
// The reference computes np.dot(T64, x32) per pixel: float32 pixel promoted to
// float64, OpenBLAS dgemv accumulating fma(T[r][2],x2, fma(T[r][0],x0, T[r][1]*x1))
// (order pinned exhaustively over all 2^24 RGB triplets by oracle/make_golden.py),
// and the float64 result rounded to float32 on store.
// ---------------------------------------------------------------------------
TMF_HD float dot3_npdot(double t0, double t1, double t2, float x0, float x1, float x2) {
  double acc = d_mul(t1, (double)x1);
  acc = d_fma(t0, (double)x0, acc);
  acc = d_fma(t2, (double)x2, acc);
  return (float)acc;
}

// f32(u8) / 255.0 (below) from the byte already converted to float (callers that extract it from a packed word
// straight onto the 2^23 magic number)
TMF_HD float unit_from_float_byte(float k) {
  // 1/255 = r_hi + r_lo to 48 bits; k * r_hi is exact in the FMA (k has 8 bits), so the result is
  // k/255 rounded once, off by < 2^-50 relative before that rounding: the correctly rounded quotient
  // for every byte (all 256 checked with exact rational arithmetic, and by the bit-exact colour tests)
  const float r_hi = 0.00392156885936856270f;      // RN32(1/255)
  const float r_lo = -2.319175823606301e-10f;   // RN32(1/255 - r_hi)
  return fmaf(k, r_hi, f_mul(k, r_lo));
}
// watermarking.py:29 - f32(u8) / 255.0 (IEEE float32 division).  For the 256
// possible inputs the quotient is reproduced exactly by a two-term reciprocal
// (checked for every byte by tests/test_hostsim.py and by the bit-exact colour
// taps), which avoids the ~12-instruction division sequence.
TMF_HD float unit_from_u8(uint32_t v) {
#if defined(__CUDA_ARCH__)
  // the byte as a float through the 2^23 magic number (one logic op + one FADD on the full-rate
  // pipes; I2F runs at 16 lanes/clk/SM, profiles/r01_ubench.txt) - exact for 0..255
  const float k = __uint_as_float(v | 0x4B000000u) - 8388608.0f;
#else
  const float k = (float)v;
#endif
  return unit_from_float_byte(k);
}
// watermarking.py:37-48 (Y only)
TMF_HD float luma_exact(float r, float g, float b) {
  return dot3_npdot(0.299, 0.587, 0.114, r, g, b);
}
// watermarking.py:37-48 (Cb, Cr with the float32 "+= 0.5" of :48)
TMF_HD void chroma_exact(float r, float g, float b, float& cb, float& cr) {
  cb = f_add(dot3_npdot(-0.169, -0.331, 0.5, r, g, b), 0.5f);
  cr = f_add(dot3_npdot(0.5, -0.419, -0.081, r, g, b), 0.5f);
}
// watermarking.py:58-73: "-= 0.5" in float32, np.dot with Ti, clip, *255 in float32, truncation
// toward zero.  ycc_to_levels_exact returns floor(f32(channel * 255)) WITHOUT the clip: the clip
// commutes with the (monotonic) scaling and floor, clip(floor(255 c), 0, 255) = floor(255 clip(c, 0, 1)),
// so callers that pack with a saturating instruction (pack4_sat_u8) get it for free - six FMNMX per
// pixel less; ycc_to_rgb8_exact clips the levels itself.
TMF_HD void ycc_to_levels_exact(float y, float cb, float cr, int& R, int& G, int& B) {
  float zb = f_add(cb, -0.5f), zr = f_add(cr, -0.5f);
  // np.dot rows (1, 0, 1.403), (1, -0.344, -0.714), (1, 1.773, 0) in dot3_npdot's order
  // acc = t1 x1; acc = fma(t0, x0, acc); acc = fma(t2, x2, acc), with the exact steps folded:
  // 0 * zb = +-0 and fma(1, y, +-0) = y; fma(1, y, acc) = y + acc; fma(0, zr, acc) = acc
  // (identical bits for finite inputs; a NaN / Inf input is garbage in the reference too)
  const double yd = (double)y, zbd = (double)zb, zrd = (double)zr;
  const float r = (float)d_fma(1.403, zrd, yd);
  const float g = (float)d_fma(-0.714, zrd, d_add(yd, d_mul(-0.344, zbd)));
  const float b = (float)d_add(yd, d_mul(1.773, zbd));
#if defined(__CUDA_ARCH__)
  // floor with one FADD.RM onto 1.5 * 2^23 and an integer subtraction instead of F2I (16 lanes/clk/SM);
  // |255 c| is far below 2^22
  R = __float_as_int(__fadd_rd(f_mul(r, 255.0f), 12582912.0f)) - 0x4B400000;
  G = __float_as_int(__fadd_rd(f_mul(g, 255.0f), 12582912.0f)) - 0x4B400000;
  B = __float_as_int(__fadd_rd(f_mul(b, 255.0f), 12582912.0f)) - 0x4B400000;
#else
  R = (int)floorf(f_mul(r, 255.0f));
  G = (int)floorf(f_mul(g, 255.0f));
  B = (int)floorf(f_mul(b, 255.0f));
#endif
}
TMF_HD void ycc_to_rgb8_exact(float y, float cb, float cr, uint32_t& R, uint32_t& G, uint32_t& B) {
  int r, g, b;
  ycc_to_levels_exact(y, cb, cr, r, g, b);
  R = (uint32_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
  G = (uint32_t)(g < 0 ? 0 : (g > 255 ? 255 : g));
  B = (uint32_t)(b < 0 ? 0 : (b > 255 ? 255 : b));
}

// four floors -> four saturated bytes in one word (byte 0 = a0): two
// cvt.pack.sat.u8.s32 (SASS I2IP.U8.S32.SAT), each clamps and packs two values.
TMF_HD uint32_t pack4_sat_u8(int a0, int a1, int a2, int a3) {
#if defined(__CUDA_ARCH__)
  uint32_t hi, w;
  asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(a3), "r"(a2), "r"(0));
  asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(w) : "r"(a1), "r"(a0), "r"(hi));
  return w;
#else
  auto cl = [](int v) { return (uint32_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); };
  return cl(a0) | (cl(a1) << 8) | (cl(a2) << 16) | (cl(a3) << 24);
#endif
}


// ---------------------------------------------------------------------------
// 8-point orthonormal DCT-II / DCT-III on a strided register vector
// (even/odd split: 8 add + 32 fma per vector).  C[k][n] = s_k cos(pi(2n+1)k/16),
// s_0 = sqrt(1/8), s_k = 1/2  - scipy.fftpack.dct(norm="ortho").
// ---------------------------------------------------------------------------
#define TMF_C1 0.98078528040323044913f
#define TMF_C2 0.92387953251128675613f
#define TMF_C3 0.83146961230254523708f
#define TMF_C4 0.70710678118654752440f
#define TMF_C5 0.55557023301960222474f
#define TMF_C6 0.38268343236508977173f
#define TMF_C7 0.19509032201612826785f
#define TMF_G0 0.35355339059327376220f   // sqrt(1/8)

template <int STRIDE>
TMF_HD void dct8(float* x) {
  const float s0 = x[0 * STRIDE] + x[7 * STRIDE], d0 = x[0 * STRIDE] - x[7 * STRIDE];
  const float s1 = x[1 * STRIDE] + x[6 * STRIDE], d1 = x[1 * STRIDE] - x[6 * STRIDE];
  const float s2 = x[2 * STRIDE] + x[5 * STRIDE], d2 = x[2 * STRIDE] - x[5 * STRIDE];
  const float s3 = x[3 * STRIDE] + x[4 * STRIDE], d3 = x[3 * STRIDE] - x[4 * STRIDE];
  const float h = 0.5f;
  x[0 * STRIDE] = TMF_G0 * ((s0 + s3) + (s1 + s2));
  x[4 * STRIDE] = (h * TMF_C4) * ((s0 + s3) - (s1 + s2));
  x[2 * STRIDE] = (h * TMF_C2) * (s0 - s3) + (h * TMF_C6) * (s1 - s2);
  x[6 * STRIDE] = (h * TMF_C6) * (s0 - s3) - (h * TMF_C2) * (s1 - s2);
  x[1 * STRIDE] = (h * TMF_C1) * d0 + (h * TMF_C3) * d1 + (h * TMF_C5) * d2 + (h * TMF_C7) * d3;
  x[3 * STRIDE] = (h * TMF_C3) * d0 - (h * TMF_C7) * d1 - (h * TMF_C1) * d2 - (h * TMF_C5) * d3;
  x[5 * STRIDE] = (h * TMF_C5) * d0 - (h * TMF_C1) * d1 + (h * TMF_C7) * d2 + (h * TMF_C3) * d3;
  x[7 * STRIDE] = (h * TMF_C7) * d0 - (h * TMF_C5) * d1 + (h * TMF_C3) * d2 - (h * TMF_C1) * d3;
}

template <int STRIDE>
TMF_HD void idct8(float* X) {
  const float h = 0.5f;
  const float a = TMF_G0 * X[0 * STRIDE], b = (h * TMF_C4) * X[4 * STRIDE];
  const float p = a + b, q = a - b;
  const float u = (h * TMF_C2) * X[2 * STRIDE] + (h * TMF_C6) * X[6 * STRIDE];
  const float v = (h * TMF_C6) * X[2 * STRIDE] - (h * TMF_C2) * X[6 * STRIDE];
  const float e0 = p + u, e3 = p - u, e1 = q + v, e2 = q - v;
  const float x1 = X[1 * STRIDE], x3 = X[3 * STRIDE], x5 = X[5 * STRIDE], x7 = X[7 * STRIDE];
  const float o0 = (h * TMF_C1) * x1 + (h * TMF_C3) * x3 + (h * TMF_C5) * x5 + (h * TMF_C7) * x7;
  const float o1 = (h * TMF_C3) * x1 - (h * TMF_C7) * x3 - (h * TMF_C1) * x5 - (h * TMF_C5) * x7;
  const float o2 = (h * TMF_C5) * x1 - (h * TMF_C1) * x3 + (h * TMF_C7) * x5 + (h * TMF_C3) * x7;
  const float o3 = (h * TMF_C7) * x1 - (h * TMF_C5) * x3 + (h * TMF_C3) * x5 - (h * TMF_C1) * x7;
  X[0 * STRIDE] = e0 + o0; X[7 * STRIDE] = e0 - o0;
  X[1 * STRIDE] = e1 + o1; X[6 * STRIDE] = e1 - o1;
  X[2 * STRIDE] = e2 + o2; X[5 * STRIDE] = e2 - o2;
  X[3 * STRIDE] = e3 + o3; X[4 * STRIDE] = e3 - o3;
}

// 2-D transforms of a row-major 8x8 register block: column pass, then row pass
// (watermarking.py:78 transforms block.T first, i.e. along the columns).
TMF_HD void dct8x8(float* a) {
#pragma unroll
  for (int j = 0; j < 8; ++j) dct8<8>(a + j);
#pragma unroll
  for (int i = 0; i < 8; ++i) dct8<1>(a + 8 * i);
}
TMF_HD void idct8x8(float* a) {
#pragma unroll
  for (int j = 0; j < 8; ++j) idct8<8>(a + j);
#pragma unroll
  for (int i = 0; i < 8; ++i) idct8<1>(a + 8 * i);
}

// ---------------------------------------------------------------------------
// one-sided (Hestenes) Jacobi SVD of an 8x8 block held in registers.
//
// A (row-major, a[8*i+j]) is overwritten by A*V = U*diag(sigma): column k ends
// as sigma_k * u_k.  V (v[8*i+j]) accumulates the right rotations when WITH_V.
// Cyclic-by-rows pair order, rotation skipped when the two columns are
// orthogonal to TOL or when one of them is at the rounding-noise floor of the
// block.  The block is pre-scaled by an exact power of two so that ||A||_F is
// in [1, 2) - no underflow in alpha*beta, no overflow for huge inputs; the
// caller multiplies the column norms by `unscale`.
//
// Returns the number of sweeps that performed at least one rotation.
// ---------------------------------------------------------------------------
#define TMF_JACOBI_TOL 1.0e-6f
#define TMF_JACOBI_DONE 1.0e-3f        // see the stop rule below
#define TMF_JACOBI_MORE 1.5f          // jacobi_cs returned 2 for some pair of the sweep
#define TMF_JACOBI_FLOOR 1.0e-14f      // (eps * ||A||_F)^2 with ||A||_F ~ 1
#define TMF_JACOBI_MAX_SWEEPS 12

TMF_HD float pow2_scale_for(float frob2, float& unscale) {
  // frob2 = ||A||_F^2 > 0.  Pick s = 2^-e with e = floor(log2(frob2)/2) so
  // that frob2*s^2 in [1, 4).  Exponent arithmetic only (exact).
  int ebits;
#if defined(__CUDA_ARCH__)
  ebits = (__float_as_int(frob2) >> 23) & 0xff;
#else
  union { float f; int32_t i; } cv; cv.f = frob2; ebits = (cv.i >> 23) & 0xff;
#endif
  if (ebits == 0) ebits = 1;                     // denormal / flushed: treat as 2^-126
  int e2 = ebits - 127;                          // floor(log2(frob2))
  int e = (e2 >= 0) ? (e2 >> 1) : -((-e2 + 1) >> 1);   // floor(e2 / 2)
  int se = 127 - e, ue = 127 + e;                // biased exponents of 2^-e and 2^e
  se = se < 1 ? 1 : (se > 254 ? 254 : se);
  ue = ue < 1 ? 1 : (ue > 254 ? 254 : ue);
#if defined(__CUDA_ARCH__)
  unscale = __int_as_float(ue << 23);
  return __int_as_float(se << 23);
#else
  union { float f; int32_t i; } a, b; a.i = ue << 23; b.i = se << 23; unscale = a.f; return b.f;
#endif
}

// Rotation of one column pair from its Gram entries (al, be, ga).  Returns 0 if the pair is
// left alone ((c, s) = (1, 0)), 1 if it is rotated, 2 if it is rotated and its |cosine| was
// above TMF_JACOBI_DONE (another sweep is needed).  The cosine tests are done on squares,
// ga^2 against tol^2 al be: no division, no square root.  Branch-free.
TMF_HD float jacobi_cs(float al, float be, float ga, float& c, float& s, float* tan_out = nullptr,
                       float done2 = TMF_JACOBI_DONE * TMF_JACOBI_DONE) {
  const float ab = al * be;
  const float gg = ga * ga;
  const float g2 = ga + ga;
  const bool rot = (gg > (TMF_JACOBI_TOL * TMF_JACOBI_TOL) * ab) && (fminf(al, be) > TMF_JACOBI_FLOOR);
  // tan of the rotation angle, smaller root: t = 2g / (d + sign(d) sqrt(d^2 + 4g^2)); forced to 0 for a
  // pair that is left alone, so that c = 1 and s = 0 fall out of the same arithmetic (no selects)
  const float d = be - al;
  const float r = f_sqrt_fast(fmaf(g2, g2, d * d));
  const float t = rot ? g2 * f_rcp_fast(d + copysignf(r, d)) : 0.0f;
  const float tt = fmaf(t, t, 1.0f);
  float cc = f_rsqrt_fast(tt);                         // tt >= 1 (exactly 1 -> exactly 1)
  cc = fmaf(0.5f * cc, fmaf(-tt * cc, cc, 1.0f), cc);  // one Newton step: c^2 + s^2 = 1 to ~1 ulp
  c = cc;
  s = cc * t;
  if (tan_out) *tan_out = t;
  return rot ? ((gg > done2 * ab) ? 2.0f : 1.0f) : 0.0f;
}

// Round-robin ("chess tournament") ordering.  A round rotates the disjoint
// pairs (0,1) (2,3) (4,5) (6,7); the rotated columns are then written back to
// the permuted positions PI, so that after 7 rounds every one of the 28 pairs
// has met once and the columns are back in their original places.  The
// permutation costs nothing (the rotation outputs simply land in other
// registers) and it makes every round the SAME code, so a sweep is a rolled
// loop of 7 iterations: ~8 KB of instructions instead of the ~50 KB of a fully
// unrolled cyclic sweep, which stalled on instruction fetch (profiles/r01_*).
//   PI: 0->0 1->2 2->4 3->1 4->6 5->3 6->7 7->5
#define TMF_PI(p) ((p) == 0 ? 0 : (p) == 1 ? 2 : (p) == 2 ? 4 : (p) == 3 ? 1 : (p) == 4 ? 6 : (p) == 5 ? 3 : (p) == 6 ? 7 : 5)

TMF_HD void rr_apply_rows(float* m, const float* c, const float* s) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = m[8 * i + j];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float x = t[2 * k], y = t[2 * k + 1];
      m[8 * i + TMF_PI(2 * k)] = fmaf(c[k], x, -s[k] * y);
      m[8 * i + TMF_PI(2 * k + 1)] = fmaf(s[k], x, c[k] * y);
    }
  }
}

// one round; returns the largest |cosine| among the pairs it rotated
template <bool WITH_V>
TMF_HD float jacobi_round(float* a, float* v) {
  float c[4], s[4], worst = 0.0f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float al = 0.f, be = 0.f, ga = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      al = fmaf(a[8 * i + 2 * k], a[8 * i + 2 * k], al);
      be = fmaf(a[8 * i + 2 * k + 1], a[8 * i + 2 * k + 1], be);
      ga = fmaf(a[8 * i + 2 * k], a[8 * i + 2 * k + 1], ga);
    }
    worst = fmaxf(worst, jacobi_cs(al, be, ga, c[k], s[k]));
  }
  rr_apply_rows(a, c, s);
  if (WITH_V) rr_apply_rows(v, c, s);
  return worst;
}

#if defined(__CUDACC__)
// ---------------------------------------------------------------------------
// Packed-fp32 (sm_100 FFMA2/FMUL2) form of the round, device only.  Rows are
// paired - element (rp, j) holds rows 2rp and 2rp+1 of column j in one 64-bit
// register pair - which a column rotation / column permutation never separates,
// and (c, s) enter as broadcast scalars, so a rotation of a column pair costs 4
// packed instructions per row pair instead of 8 scalar ones per two rows.  Every
// lane is the same round-to-nearest operation as in the scalar code; only the
// summation order of the three dot products differs.
// ---------------------------------------------------------------------------
__device__ __forceinline__ float2 tmf_bc2(float x) { return make_float2(x, x); }

__device__ __forceinline__ void rr_apply_rows2(float2* m2, const float* c, const float* s) {
#pragma unroll
  for (int rp = 0; rp < 4; ++rp) {
    float2 t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = m2[8 * rp + j];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 x = t[2 * k], y = t[2 * k + 1];
      m2[8 * rp + TMF_PI(2 * k)] = __ffma2_rn(tmf_bc2(c[k]), x, __fmul2_rn(tmf_bc2(-s[k]), y));
      m2[8 * rp + TMF_PI(2 * k + 1)] = __ffma2_rn(tmf_bc2(s[k]), x, __fmul2_rn(tmf_bc2(c[k]), y));
    }
  }
}

// squared norms of the 8 columns (by position), from the packed block
__device__ __forceinline__ void column_norms2_packed(const float2* a2, float* n2) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int rp = 0; rp < 4; ++rp) acc = __ffma2_rn(a2[8 * rp + j], a2[8 * rp + j], acc);
    n2[j] = acc.x + acc.y;
  }
}

// One round.  The squared column norms are carried along instead of being recomputed for
// every pair: a rotation by tan t changes them by exactly -/+ t*gamma, so only the cross
// product gamma needs a pass over the column (4 packed FMAs instead of 12).  The carried
// values are refreshed from the data at the start of every sweep (jacobi_svd8), which
// bounds the drift; they only steer the rotation angle and the skip tests, never the
// result (singular values are taken from the columns at the end).
// (Used for the values-only solve; with V in registers as well the 8 extra live values
// cost more in spills than the saved FMAs bring - measured - so that variant recomputes.)
template <bool WITH_V>
__device__ __forceinline__ float jacobi_round2(float2* a2, float2* v2, float* n2) {
  float c[4], s[4], nn[8], worst = 0.0f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float2 ga = make_float2(0.f, 0.f), al2 = ga, be2 = ga;
#pragma unroll
    for (int rp = 0; rp < 4; ++rp) {
      const float2 x = a2[8 * rp + 2 * k], y = a2[8 * rp + 2 * k + 1];
      ga = __ffma2_rn(x, y, ga);
      if (WITH_V) { al2 = __ffma2_rn(x, x, al2); be2 = __ffma2_rn(y, y, be2); }
    }
    const float g = ga.x + ga.y;
    const float al = WITH_V ? al2.x + al2.y : n2[2 * k], be = WITH_V ? be2.x + be2.y : n2[2 * k + 1];
    float t;
    worst = fmaxf(worst, jacobi_cs(al, be, g, c[k], s[k], &t));
    if (!WITH_V) {
      const float tg = t * g;                                // t * gamma (0 when the pair is skipped)
      nn[TMF_PI(2 * k)] = fmaxf(al - tg, 0.0f);
      nn[TMF_PI(2 * k + 1)] = be + tg;
    }
  }
  if (!WITH_V) {
#pragma unroll
    for (int j = 0; j < 8; ++j) n2[j] = nn[j];
  }
  rr_apply_rows2(a2, c, s);
  if (WITH_V) rr_apply_rows2(v2, c, s);
  return worst;
}
#endif  // __CUDACC__

// Sweeps stop when the largest cosine met during a sweep is below
// TMF_JACOBI_DONE: Jacobi converges quadratically, so the sweep that just ran
// leaves cosines of order DONE^2 - no extra sweep is spent only to find out
// that nothing rotates.

template <bool WITH_V>
TMF_HD int jacobi_svd8(float* a, float* v, float& unscale) {
  float frob2 = 0.f;
#pragma unroll
  for (int k = 0; k < 64; ++k) frob2 = fmaf(a[k], a[k], frob2);
  unscale = 1.0f;
  if (WITH_V) {
#pragma unroll
    for (int k = 0; k < 64; ++k) v[k] = ((k >> 3) == (k & 7)) ? 1.0f : 0.0f;
  }
  // Inf/NaN input: frob2 is not finite; rotations would only spread NaNs - skip.
  const bool live = (frob2 > 0.0f) && (frob2 < INFINITY);
  if (live) {
    const float s = pow2_scale_for(frob2, unscale);
#pragma unroll
    for (int k = 0; k < 64; ++k) a[k] *= s;
  }
  int sweeps = 0;
  bool more = live;
#if defined(__CUDA_ARCH__)   // packed-fp32 rounds on the device, scalar ones on the host (tests/hostsim)
  // pack row pairs (register renaming), iterate, unpack
  float2 a2[32], v2[WITH_V ? 32 : 1];
#pragma unroll
  for (int rp = 0; rp < 4; ++rp)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      a2[8 * rp + j] = make_float2(a[16 * rp + j], a[16 * rp + 8 + j]);
      if (WITH_V) v2[8 * rp + j] = make_float2(v[16 * rp + j], v[16 * rp + 8 + j]);
    }
  for (int it = 0; it < TMF_JACOBI_MAX_SWEEPS && more; ++it) {
    float worst = 0.0f, n2[8];
    if (!WITH_V) column_norms2_packed(a2, n2);
#pragma unroll 1
    for (int r = 0; r < 7; ++r) worst = fmaxf(worst, jacobi_round2<WITH_V>(a2, v2, n2));
    sweeps += (worst > 0.0f) ? 1 : 0;
    more = worst > TMF_JACOBI_MORE;
  }
#pragma unroll
  for (int rp = 0; rp < 4; ++rp)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      a[16 * rp + j] = a2[8 * rp + j].x; a[16 * rp + 8 + j] = a2[8 * rp + j].y;
      if (WITH_V) { v[16 * rp + j] = v2[8 * rp + j].x; v[16 * rp + 8 + j] = v2[8 * rp + j].y; }
    }
#else
  for (int it = 0; it < TMF_JACOBI_MAX_SWEEPS && more; ++it) {
    float worst = 0.0f;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int r = 0; r < 7; ++r) worst = fmaxf(worst, jacobi_round<WITH_V>(a, v));
    sweeps += (worst > 0.0f) ? 1 : 0;
    more = worst > TMF_JACOBI_MORE;
  }
#endif
  return sweeps;
}

// squared column norms of A*V (sigma_k^2 in the scaled domain)
TMF_HD void column_norms2(const float* a, float* n2) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s = fmaf(a[8 * i + j], a[8 * i + j], s);
    n2[j] = s;
  }
}

// ---------------------------------------------------------------------------
// Dominant column only.  Embed and extract use sigma_0 and u_0 and nothing else of the SVD
// (watermarking.py:198 changes S[0]; :285 reads S[0]), and one-sided Jacobi delivers them as ONE
// column of A V: p = sigma_0 u_0.  Rotations between two OTHER columns move neither sigma_0 nor u_0,
// so once it is known which column will become p they can be left out: rotate p against each of
// the other seven, again and again, until p is orthogonal to all of them - then A A^T =
// p p^T + R R^T with p orthogonal to R's range, i.e. p is a left singular vector times its singular value.
//
// Certificate that it is the TOP one, checked before the first rotation: |p|^2 > sum of the other
// columns' |a_j|^2 =: r.  Rotations preserve the Frobenius norm and every rotation of (p, a_j) by the
// smaller angle lengthens the longer column, so the inequality only gets stronger; at the end
// sigma_max(R) <= ||R||_F < |p|.  The same two numbers bound the gap, sigma_1^2 / sigma_0^2 <= r / |p|^2,
// which sets the rate (the seven partners never become orthogonal to each other, so convergence is
// linear, about that ratio per sweep - with a 4 % gap the iteration crawls: measured).  The path is
// therefore taken only when |p|^2 > TMF_JACOBI_TOP_DOMINANCE * r (ratio below 1/4).  That holds whenever
// the block has a dominant DC term - any block of an ordinary photograph, noise around a mean (ratio
// 0.15 for uniform noise), flat and saturated areas - and fails for blocks that are black but for a
// few pixels, striped at the block period, or made of two comparable patches; those take the full
// cyclic sweeps of jacobi_svd8.
//
// Cost: 7 rotations per sweep instead of 28, and 2 (photographs) to 3-4 (noise, ratio near 1/4) sweeps
// instead of 5.  A sweep whose largest cosine was below TMF_JACOBI_TOP_DONE is the last (it still rotated
// those away; what is left is that times the ratio: u_0 to ~1e-4 of the quantiser's step, sigma_0 to
// second order).  Measured against float64 on 2 000 constructed blocks with ratios 0.05 ... 0.24: at
// most 4 sweeps, sigma_0 within 5e-7 relative (fp32 round-off of ~20 rotations of p, as in the full sweeps).
//
// a[64]: the block (destroyed).  u[8] <- u_0 (unit norm).  Returns sigma_0 in the block's units; 0 for an
// all-zero block (u undefined).  *sweeps: sweeps run, +100 if the full routine was used.
// ---------------------------------------------------------------------------
#define TMF_JACOBI_TOP_DONE 3.0e-4f
#define TMF_JACOBI_TOP_DOMINANCE 4.0f

TMF_HD float top_column8(float* a, float* u, int* sweeps) {
  float n2[8];
  column_norms2(a, n2);
  float frob2 = n2[0], best = n2[0];
  int top = 0;
#pragma unroll
  for (int j = 1; j < 8; ++j) {
    frob2 += n2[j];
    if (n2[j] > best) { best = n2[j]; top = j; }
  }
  const bool live = (frob2 > 0.0f) && (frob2 < INFINITY);
  float unscale = 1.0f, nrm2;
  if (live && best > TMF_JACOBI_TOP_DOMINANCE * (frob2 - best)) {
    const float sc = pow2_scale_for(frob2, unscale);
#pragma unroll
    for (int k = 0; k < 64; ++k) a[k] *= sc;
    if (top != 0) {                       // rare (the DC column dominates a DCT block): p to slot 0
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float x = a[8 * i];
#pragma unroll
        for (int j = 1; j < 8; ++j) {
          const float y = a[8 * i + j];
          a[8 * i + j] = (j == top) ? x : y;
          x = (j == top) ? y : x;
        }
        a[8 * i] = x;
      }
    }
    int sw = 0;
    bool more = true;
#if defined(__CUDA_ARCH__)   // packed-fp32 rounds on the device, scalar ones on the host (tests/hostsim)
    float2 a2[32];
#pragma unroll
    for (int rp = 0; rp < 4; ++rp)
#pragma unroll
      for (int j = 0; j < 8; ++j) a2[8 * rp + j] = make_float2(a[16 * rp + j], a[16 * rp + 8 + j]);
#pragma unroll 1
    for (int it = 0; it < TMF_JACOBI_MAX_SWEEPS && more; ++it) {
      float worst = 0.0f;
      column_norms2_packed(a2, n2);       // carried through the sweep, refreshed from the data here
#pragma unroll
      for (int j = 1; j < 8; ++j) {
        float2 ga = make_float2(0.f, 0.f);
#pragma unroll
        for (int rp = 0; rp < 4; ++rp) ga = __ffma2_rn(a2[8 * rp], a2[8 * rp + j], ga);
        const float g = ga.x + ga.y;
        float c, s, t;
        worst = fmaxf(worst, jacobi_cs(n2[0], n2[j], g, c, s, &t, TMF_JACOBI_TOP_DONE * TMF_JACOBI_TOP_DONE));
        const float tg = t * g;
        n2[0] = fmaxf(n2[0] - tg, 0.0f);
        n2[j] = n2[j] + tg;
#pragma unroll
        for (int rp = 0; rp < 4; ++rp) {
          const float2 x = a2[8 * rp], y = a2[8 * rp + j];
          a2[8 * rp] = __ffma2_rn(tmf_bc2(c), x, __fmul2_rn(tmf_bc2(-s), y));
          a2[8 * rp + j] = __ffma2_rn(tmf_bc2(s), x, __fmul2_rn(tmf_bc2(c), y));
        }
      }
      ++sw;
      more = worst > TMF_JACOBI_MORE;
    }
#pragma unroll
    for (int rp = 0; rp < 4; ++rp) { u[2 * rp] = a2[8 * rp].x; u[2 * rp + 1] = a2[8 * rp].y; }
#else
    for (int it = 0; it < TMF_JACOBI_MAX_SWEEPS && more; ++it) {
      float worst = 0.0f;
      column_norms2(a, n2);
      for (int j = 1; j < 8; ++j) {
        float g = 0.0f;
        for (int i = 0; i < 8; ++i) g = fmaf(a[8 * i], a[8 * i + j], g);
        float c, s, t;
        worst = fmaxf(worst, jacobi_cs(n2[0], n2[j], g, c, s, &t, TMF_JACOBI_TOP_DONE * TMF_JACOBI_TOP_DONE));
        const float tg = t * g;
        n2[0] = fmaxf(n2[0] - tg, 0.0f);
        n2[j] = n2[j] + tg;
        for (int i = 0; i < 8; ++i) {
          const float x = a[8 * i], y = a[8 * i + j];
          a[8 * i] = fmaf(c, x, -s * y);
          a[8 * i + j] = fmaf(s, x, c * y);
        }
      }
      ++sw;
      more = worst > TMF_JACOBI_MORE;
    }
    for (int i = 0; i < 8; ++i) u[i] = a[8 * i];
#endif
    if (sweeps) *sweeps = sw;
    nrm2 = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) nrm2 = fmaf(u[i], u[i], nrm2);
  } else {
    const int sw = jacobi_svd8<false>(a, nullptr, unscale);
    if (sweeps) *sweeps = sw + 100;
    column_norms2(a, n2);
    nrm2 = n2[0];
    top = 0;
#pragma unroll
    for (int j = 1; j < 8; ++j) {
      if (n2[j] > nrm2) { nrm2 = n2[j]; top = j; }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float x = 0.0f;
#pragma unroll
      for (int j = 0; j < 8; ++j) x = (j == top) ? a[8 * i + j] : x;
      u[i] = x;
    }
  }
  const float nrm = f_sqrt(nrm2);
  if (!(nrm > 0.0f)) return 0.0f;
  const float inv = f_div(1.0f, nrm);
#pragma unroll
  for (int i = 0; i < 8; ++i) u[i] *= inv;
  return nrm * unscale;
}

// ---------------------------------------------------------------------------
// whole-block stages shared by the fused kernels and the host test harness
// ---------------------------------------------------------------------------

// watermarking.py:198 - S[0] += alpha * (wm / 255.0): float32 + float64 product,
// float64 sum stored back to float32.
TMF_HD float modulate_sigma0(float s0, double alpha, uint32_t wm_u8) {
  // w = wm / 255.0 (float64 division, :198) without the division sequence (reciprocal seed, Newton steps, range
  // check, slow-path call: ~12 instructions per block): with r = RN(1 / 255), q = k r, e = fma(-q, 255, k),
  // q + e r is the correctly rounded quotient for every byte k (all 256 checked with exact rational arithmetic,
  // tests/test_host_logic.py; the plain product k r alone is wrong for 24 of them)
  const double k = (double)wm_u8, r = 1.0 / 255.0;
  const double q = d_mul(k, r);
  const double w = d_fma(d_fma(-q, 255.0, k), r, q);
  return (float)((double)s0 + d_mul(alpha, w));
}

// watermarking.py:285-289 - (S_w[0] - S_o[0]) / alpha in float32 (NumPy >= 2
// scalar rules), widened to float64, clip [0, 1], * 255 (float64), truncate.
TMF_HD uint32_t extract_level(float sw, float so, double alpha) {
  const float e32 = f_div(f_add(sw, -so), (float)alpha);
  double e = (double)e32;
  e = e < 0.0 ? 0.0 : (e > 1.0 ? 1.0 : e);
  return (uint32_t)d_mul(e, 255.0);
}

// Faithful embed of one luma block held in a[64] (row-major, values in [0,1]):
// DCT -> Jacobi SVD -> sigma_top += alpha*w -> U diag(S') V^T -> IDCT, in place.
// v[64] is scratch.  Returns sigma_top (before modulation); *sweeps optional.
TMF_HD float embed_block_faithful(float* a, float* v, double alpha, uint32_t wm_u8, int* sweeps) {
  dct8x8(a);
  float unscale;
  const int sw = jacobi_svd8<true>(a, v, unscale);
  if (sweeps) *sweeps = sw;
#pragma unroll
  for (int k = 0; k < 64; ++k) a[k] *= unscale;
  float n2[8];
  column_norms2(a, n2);
  float best = n2[0];
  int top = 0;
#pragma unroll
  for (int j = 1; j < 8; ++j) {
    if (n2[j] > best) { best = n2[j]; top = j; }
  }
  const float sig = f_sqrt(best);
  const float sig_new = modulate_sigma0(sig, alpha, wm_u8);
  if (sig > 0.0f) {
    const float f = f_div(sig_new, sig);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float sc = (j == top) ? f : 1.0f;
#pragma unroll
      for (int i = 0; i < 8; ++i) a[8 * i + j] *= sc;
    }
  } else {
    a[0] = sig_new;   // all-zero block: LAPACK returns U = V = I, so the mark lands on DC
  }
  // M = (A V diag(scale)) V^T, one row at a time, in place
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float t[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) t[k] = a[8 * i + k];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float m = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) m = fmaf(t[k], v[8 * j + k], m);
      a[8 * i + j] = m;
    }
  }
  idct8x8(a);
  return sig;
}

// Faithful pipeline up to the top singular triplet, without accumulating V:
// DCT -> one-sided Jacobi (values-only form: A <- A V = U diag(sigma)) -> the top column is
// sigma_top * u0, so u0 needs no V; the right vector follows from the block itself,
// v0 = D^T u0 / sigma_top, and the reference's U diag(S') V^T = D + d u0 v0^T exactly
// (watermarking.py:198-201 changes S[0] only).  Carried back to the spatial domain by the
// orthonormal DCT: block' = B + d (C^T u0)(C^T v0)^T with C^T v0 = B^T (C^T u0) / sigma_top.
// a[64] = the luma block (destroyed); uB[8] <- C^T u0 (unit norm).  Returns sigma_top in the
// block's units (0 for an all-zero block: then uB is the constant 1/sqrt(8) vector, which with
// v = uB reproduces LAPACK's U = V = I convention - the mark lands on the DC coefficient).
TMF_HD float top_left_vector_faithful(float* a, float* uB, int* sweeps) {
  dct8x8(a);
  const float sig = top_column8(a, uB, sweeps);
  if (!(sig > 0.0f)) {
#pragma unroll
    for (int i = 0; i < 8; ++i) uB[i] = TMF_G0;
    return 0.0f;
  }
  idct8<1>(uB);
  return sig;
}

// Largest singular value of the DCT of one luma block (extract side).
TMF_HD float sigma0_block_faithful(float* a, int* sweeps) {
  dct8x8(a);
  float u[8];
  return top_column8(a, u, sweeps);
}

}  // namespace tmf
