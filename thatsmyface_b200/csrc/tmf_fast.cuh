// FAST mode: algebraically reduced embed / extract of one 8x8 luma block,
// organised as two streaming passes over the block's rows.
//
// Algebra.  The orthonormal 2-D DCT preserves singular values, and for
// D = C B C^T with D = U S V^T the reference's output block is
//     IDCT(U diag(S + d e0) V^T) = B + d (C^T u0)(C^T v0)^T = B + d u0_B v0_B^T
// (modules/watermarking.py:192-204), where (s0, u0_B, v0_B) is the top singular
// triplet of the *spatial* block B.  Embed needs one triplet and a rank-1
// update, extract needs one singular value; no DCT, no full SVD.
//
// Streaming.  Nothing below ever holds the 64 luma values of a block.
//   pass 1 (per row):  luma of the row -> G += row^T row      (G = B^T B, 36 regs);
//                      the row's luma is parked in a thread-private shared-memory column
//   middle:            top eigenpair (mu, w) of M = G / tr(G)
//   pass 2 (per row):  z = row.w, d_j = f z w_j with f = d255 / (sigma255 w.w),
//                      pixel out = M_rgb k + d_j, clip, truncate
// so the row loops stay rolled (small code: the unrolled version stalled on
// instruction fetch, profiles/r01_*) and the register file holds G, not B.
//
// Top eigenpair, robustly.  B >= 0 (luma) makes G >= 0 entrywise, so by
// Perron-Frobenius v0 >= 0 and the all-ones start has tan(angle to v0) <= sqrt 7.
// M has unit trace, so for any x the Rayleigh quotient mu^ <= mu_0 gives
//     rho^ = (1 - mu^) / mu^ >= (sum_{i>=1} mu_i) / mu_0 >= mu_1 / mu_0:
// a certified upper bound on the power-iteration ratio that needs no knowledge
// of the gap.  After J products w_J = M^J 1 the vector error is <= sqrt7 rho^^J
// and mu~ = (w_J.w_J)/(w_{J-1}.w_J) <= mu_0 is off by <= 7 rho^^(2J-1); J <= 5 is
// read off precomputed thresholds.  Blocks whose bound is too weak for that
// (textured, near-tied) take the slow path: repeated squaring P = M M, where
// 1 - tr(P) bounds the distance from rank one for any gap.
//
// Pixels are carried in 0..255 units (exact small integers as floats); colour
// math is fp32 FMA with the reference's constants - within ~2 ulp of the
// reference's float64-dot-then-float32 values, not bit-identical (the faithful
// mode is).  All CUDA-core fp32: every contraction here is 8 wide.
#pragma once
#include "tmf_math.cuh"

namespace tmf {


// upper triangle of a symmetric NxN in N(N+1)/2 registers: index of (i, j), any order.
// N = 8 is the reference's BLOCK_SIZE; the templates also serve the other sizes its UI
// offers (4..16, embed_watermark_page.py:324-331).
template <int N>
TMF_HD constexpr int sym_idx(int i, int j) {
  return (i <= j) ? (i * N - (i * (i + 1)) / 2 + j) : (j * N - (j * (j + 1)) / 2 + i);
}
#define TMF_SYM(i, j) (tmf::sym_idx<8>((i), (j)))

// pass 1: G += y^T y for one row y[N] of the block
template <int N = 8>
TMF_HD void gram_accumulate_row(const float* y, float* g) {
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = i; j < N; ++j) g[sym_idx<N>(i, j)] = fmaf(y[i], y[j], g[sym_idx<N>(i, j)]);
}

template <int N = 8>
TMF_HD float sym_trace(const float* g) {
  float tr = 0.f;
#pragma unroll
  for (int i = 0; i < N; ++i) tr += g[sym_idx<N>(i, i)];
  return tr;
}

// p = m * m for symmetric m (upper triangles); returns tr(p)
template <int N = 8>
TMF_HD float sym_square(const float* m, float* p) {
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = i; j < N; ++j) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < N; ++k) s = fmaf(m[sym_idx<N>(i, k)], m[sym_idx<N>(k, j)], s);
      p[sym_idx<N>(i, j)] = s;
    }
  return sym_trace<N>(p);
}

// y = M x for symmetric M (upper triangle)
template <int N = 8>
TMF_HD void sym_matvec(const float* m, const float* x, float* y) {
#pragma unroll
  for (int i = 0; i < N; ++i) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < N; ++j) s = fmaf(m[sym_idx<N>(i, j)], x[j], s);
    y[i] = s;
  }
}

template <int N = 8>
TMF_HD float dotn(const float* a, const float* b) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < N; ++i) s = fmaf(a[i], b[i], s);
  return s;
}
TMF_HD float dot8(const float* a, const float* b) { return dotn<8>(a, b); }

// Top eigenpair of the unit-trace Gram matrix m (entrywise >= 0; DESTROYED) of a
// luma block: certified power iteration with adaptive squaring.
//
//   level 0: v = 1 (tan(angle to v0) <= sqrt(N-1) by Perron-Frobenius), x = M v, y = M x.
//   The Rayleigh quotient of x gives rho^ = (1 - mu^)/mu^ >= mu_1/mu_0 (unit trace), so
//   tan(angle(y)) <= theta * rho^^2, and every further product multiplies the bound by
//   rho^.  If at most 3 more products reach the tolerance, do them and stop.
//   Otherwise square: M <- M^2 / tr(M^2) (eigenvalue ratios are squared, trace is 1
//   again), keep y as the start vector with its bound, and repeat.  Textured blocks take
//   one squaring, near-tied ones a few; the bound is rigorous for any gap.
//
// On return w (not normalised, ww = w.w) spans v0 to `tol` and mu ~ mu_0 of the ORIGINAL
// m to ~1e-7 (mu_l = sqrt(mu_{l+1} tr(M_l^2)) unwinds the squarings; carried forward as one
// running product so that no per-level array - local memory - is needed).  Returns
// products + 100 * squarings.
#define TMF_FAST_TOL_VEC_EMBED 1.0e-6f     // u0 v0^T itself is used
#define TMF_FAST_TOL_VEC_EXTRACT 3.0e-4f   // sigma0 only: second order in the vector error
#define TMF_FAST_MAX_LEVELS 18

// ||m||_F^2 = tr(m^2) for symmetric m (upper triangle)
template <int N = 8>
TMF_HD float sym_frob2(const float* m) {
  float d = 0.f, o = 0.f;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    d = fmaf(m[sym_idx<N>(i, i)], m[sym_idx<N>(i, i)], d);
#pragma unroll
    for (int j = 0; j < N; ++j)          // fixed trip count: a `j = i + 1` start is left rolled for N >= 10, and the
      if (j > i) o = fmaf(m[sym_idx<N>(i, j)], m[sym_idx<N>(i, j)], o);   // dynamic index then pins m[] to local memory
  }
  return fmaf(2.0f, o, d);
}

// Second certificate, used only when the trace bound is too weak (textured / noise blocks,
// where the trace bound charges mu_0 with ALL the other eigenvalues):
//     mu_1^2 <= sum_{i>=1} mu_i^2 = tr(M^2) - mu_0^2 <= tr(M^2) - mu^^2     (mu^ <= mu_0)
// so mu_1/mu_0 <= sqrt(tr(M^2) - mu^^2) / mu^.  For a noise block with seven comparable small
// eigenvalues this is ~sqrt(7) times their ratio instead of 7 times it, which certifies a
// handful of plain products (64 FMAs each) where the trace bound forced a squaring (~330).
// The difference cancels when M is nearly rank one (both terms ~1, fp32 noise ~1e-6) - but
// then the trace bound has already certified the block on the fast path: this branch runs only
// when sum_{i>=1} mu_i / mu_0 is large (> ~0.06 for embed), hence tr(M^2) - mu_0^2 >= ~4e-4, and
// the slack of 8e-6 (the worst-case fp32 rounding of a 64-term sum near 1, two orders below the
// quantity) keeps the bound an upper bound.
#ifndef TMF_FAST_FROB_SLACK
#define TMF_FAST_FROB_SLACK 8.0e-6f
#endif
#ifndef TMF_FAST_MAX_MORE
#define TMF_FAST_MAX_MORE 6                // plain products allowed after the first two, per level
#endif

template <bool EMBED, int N = 8>
TMF_HD int top_pair(float* m, float* w, float& ww, float& mu) {
  constexpr int NS = N * (N + 1) / 2;
  const float tol = EMBED ? TMF_FAST_TOL_VEC_EMBED : TMF_FAST_TOL_VEC_EXTRACT;
  const float theta0 = f_sqrt((float)(N - 1));
  float x[N], y[N];
  float unwind = 1.0f;   // prod_l tr(M_l^2)^(1/2^(l+1)): mu_0 = unwind * mu_L^(1/2^L), kept in a register
  float theta = theta0;
  int level = 0, products = 0;
#pragma unroll
  for (int i = 0; i < N; ++i) y[i] = 1.0f;
  float xy, yy;
  for (;;) {
    sym_matvec<N>(m, y, x);              // x = M v      (v is y from the previous level, or ones)
    sym_matvec<N>(m, x, y);              // y = M x
    products += 2;
    const float xx = dotn<N>(x, x);
    xy = dotn<N>(x, y);
    yy = dotn<N>(y, y);
    float rho = fmaxf(xx - xy, 0.0f) * f_rcp_fast(xy);
    float err = theta * rho * rho;       // bound on tan(angle(y, v0))
    int more = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (err > tol) { err *= rho; ++more; }
    }
    if (err > tol && level < TMF_FAST_MAX_LEVELS - 1) {
      const float t2 = sym_frob2<N>(m);  // tr(M^2)
      const float muh = xy * f_rcp_fast(xx);                       // Rayleigh quotient of x, <= mu_0
      const float r2 = fminf(rho, f_sqrt(fmaxf(t2 - muh * muh, 0.0f) + TMF_FAST_FROB_SLACK) * f_rcp_fast(muh));
      float e2 = theta * r2 * r2;
      int m2 = 0;
#pragma unroll 1
      for (int k = 0; k < TMF_FAST_MAX_MORE; ++k) {
        if (e2 > tol) { e2 *= r2; ++m2; }
      }
      rho = r2;
      if (e2 <= tol) { err = e2; more = m2; }
    }
    if (err <= tol || level >= TMF_FAST_MAX_LEVELS - 1) {
#pragma unroll 1
      for (int k = 0; k < more; ++k) {   // rolled: one copy of the product (the predicated unrolled forms measured slower)
        sym_matvec<N>(m, y, x);          // x = M y
        xy = dotn<N>(y, x);              // w_{J-1}.w_J
        yy = dotn<N>(x, x);              // w_J.w_J
#pragma unroll
        for (int i = 0; i < N; ++i) y[i] = x[i];
      }
      products += more;
      break;
    }
    // square in place: M <- M^2 / tr(M^2)
    float p[NS];
    const float t = sym_square<N>(m, p);
    {
      float r = t;                        // t^(1/2^(level+1)); levels are few, square roots cheap
      for (int k = 0; k <= level; ++k) r = f_sqrt(r);
      unwind *= r;
    }
    const float inv = f_rcp_fast(t);
#pragma unroll
    for (int k = 0; k < NS; ++k) m[k] = p[k] * inv;
    // carry y over as the next start vector, rescaled so nothing underflows
    const float rn = f_rsqrt(yy);
#pragma unroll
    for (int i = 0; i < N; ++i) y[i] *= rn;
    theta = fminf(theta * rho * rho, theta0);
    ++level;
  }
#pragma unroll
  for (int i = 0; i < N; ++i) w[i] = y[i];
  ww = yy;
  mu = yy * f_rcp_fast(xy);              // (w_J.w_J)/(w_{J-1}.w_J) <= mu_0 of the current level
  for (int l = 0; l < level; ++l) mu = f_sqrt(mu);
  mu *= unwind;
  return products + 100 * level;
}

// Per-block scalars of the embed, from the Gram matrix accumulated in pass 1
// (g in 0..255 units, destroyed).  Output for pass 2:
//   w[N], and  y'_ij = y_ij + (f * z_i + c) * w_j,  z_i = row_i . w
// with (f, c) = (d255 / (sigma255 w.w), 0), or (0, d255/N) with w = 1 for an
// all-zero block (LAPACK's U = V = I puts the mark on the DC coefficient, which
// is the constant 1/N pattern in the spatial domain).  Returns sigma0 in the
// reference's units (luma in [0, 1]).
//
// `unit` = reference luma units per unit of the Gram's luma: 1/255 for luma in 0..255
// (luma255_fast), 1/255000 for the exact integer luma 299 r + 587 g + 114 b (luma1000_exact).
template <int N = 8>
TMF_HD float embed_block_scalars_fast(float* g, double alpha, uint32_t wm_u8, float* w, float& f, float& c,
                                      int* iters, float unit = 1.0f / 255.0f) {
  constexpr int NS = N * (N + 1) / 2;
  const float tr = sym_trace<N>(g);
  float sig255 = 0.0f, ww = (float)N;
  if (tr > 0.0f) {
    const float inv = f_rcp_fast(tr);
#pragma unroll
    for (int k = 0; k < NS; ++k) g[k] *= inv;
    float mu;
    const int it = top_pair<true, N>(g, w, ww, mu);
    if (iters) *iters = it;
    sig255 = f_sqrt(tr * mu);
  } else {
    if (iters) *iters = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) w[i] = 1.0f;
  }
  const float sig = sig255 * unit;
  const float d255 = (modulate_sigma0(sig, alpha, wm_u8) - sig) * 255.0f;   // watermarking.py:198
  if (tr > 0.0f) { f = f_div(d255, sig255 * ww); c = 0.0f; }
  else { f = 0.0f; c = d255 * (1.0f / (float)N); }
  return sig;
}

// Largest singular value (reference units) from the pass-1 Gram matrix (destroyed).
template <int N = 8>
TMF_HD float sigma0_from_gram_fast(float* g, int* iters, float unit = 1.0f / 255.0f) {
  constexpr int NS = N * (N + 1) / 2;
  const float tr = sym_trace<N>(g);
  if (iters) *iters = 0;
  if (!(tr > 0.0f)) return 0.0f;
  const float inv = f_rcp_fast(tr);
#pragma unroll
  for (int k = 0; k < NS; ++k) g[k] *= inv;
  float w[N], ww, mu;
  const int it = top_pair<false, N>(g, w, ww, mu);
  if (iters) *iters = it;
  return f_sqrt(tr * mu) * unit;
}

// The reference's luma weights are three-decimal constants (watermarking.py:37), so
// 1000 * 255 * Y = 299 r + 587 g + 114 b is an exact integer below 2^18: the N = 8 kernels
// compute it with two IDP.2A per pixel straight from the packed bytes (no byte extraction, no
// rounding) and carry the Gram matrix and the luma stash in these units.
TMF_HD float luma1000_exact(uint32_t r, uint32_t g, uint32_t b) { return (float)(299u * r + 587u * g + 114u * b); }
#define TMF_LUMA1000_UNIT (1.0f / 255000.0f)

// --- fp32 colour in 0..255 units -------------------------------------------
// luma of watermarking.py:37-45 times 255
TMF_HD float luma255_fast(float r, float g, float b) {
  return fmaf(0.299f, r, fmaf(0.587f, g, 0.114f * b));
}
// watermarking.py:37-48 and :58-73 composed: the reference maps (r, g, b) to
// (y, cb, cr), adds the mark to y, and maps back with a matrix that is not the
// exact inverse, so a pixel comes back as  M k + d  with  M = Ti T  (the two
// reference matrices multiplied out; every entry is an exact multiple of 1e-6) and d
// the luma change of that pixel.  M = I + E and E's rows sum to zero (a grey pixel maps
// to itself), so with u = r - g, v = b - g (exact):
//     out_c = k_c + s_c,   s_c = E_c0 u + E_c2 v + d        (|s_c| is small)
// All in 0..255 units; the caller floors k_c + s_c (exactly, floor_sum_to_int) and clips.
TMF_HD void rgb255_delta_fast(float r, float g, float b, float d, float& sR, float& sG, float& sB) {
  const float u = r - g, v = b - g;
  sR = fmaf(5.00e-4f, u, fmaf(3.57e-4f, v, d));
  sG = fmaf(1.36e-4f, u, fmaf(-1.66e-4f, v, d));
  sB = fmaf(-6.37e-4f, u, fmaf(5.00e-4f, v, d));
}
// the same pixel as one value per channel (k_c + s_c rounded to fp32): kept for callers that
// want the unquantised colour
TMF_HD void rgb255_out_fast(float r, float g, float b, float d, float& R, float& G, float& B) {
  float sR, sG, sB;
  rgb255_delta_fast(r, g, b, d, sR, sG, sB);
  R = r + sR; G = g + sG; B = b + sB;
}

// floor(x) as a signed integer for |x| < 2^22 with one FADD.RM (round toward
// -inf into the mantissa of 1.5*2^23) and one integer subtract - both full-rate
// pipes, unlike F2I (measured 16/clk/SM on B200, profiles/r01_ubench.txt).
// The clip to [0, 255] (watermarking.py:70) is done by the packing instruction;
// for the clipped value floor == the reference's truncation (:73).
TMF_HD int floor_to_int(float x) {
#if defined(__CUDA_ARCH__)
  return __float_as_int(__fadd_rd(x, 12582912.0f)) - 0x4B400000;
#else
  return (int)floorf(x);
#endif
}

// floor(k + s) for an integer-valued k in [0, 2^22) and |s| < 2^22, with no rounding of the
// sum: k sits on the quantiser's bias (1.5*2^23 + k is exact) and FADD.RM of s onto it is
// 1.5*2^23 + floor(k + s).
TMF_HD int floor_sum_to_int(float k, float s) {
#if defined(__CUDA_ARCH__)
  return __float_as_int(__fadd_rd(s, 12582912.0f + k)) - 0x4B400000;
#else
  return (int)floor((double)k + (double)s);
#endif
}

// pass 2 for one row: r, g, b in 0..255 units and the row's luma y (kept from
// pass 1) -> 3N output levels q[3j + c] (unclipped floors; pack4_sat_u8 clips)
template <int N = 8>
TMF_HD void embed_row_fast(const float* r, const float* g, const float* b, const float* y, const float* w, float f,
                           float c, int* q) {
  const float du = fmaf(f, dotn<N>(y, w), c);
#pragma unroll
  for (int j = 0; j < N; ++j) {
    float sR, sG, sB;
    rgb255_delta_fast(r[j], g[j], b[j], du * w[j], sR, sG, sB);
    q[3 * j] = floor_sum_to_int(r[j], sR);
    q[3 * j + 1] = floor_sum_to_int(g[j], sG);
    q[3 * j + 2] = floor_sum_to_int(b[j], sB);
  }
}

}  // namespace tmf
