// FAST mode: algebraically reduced embed / extract of one 8x8 luma block.
//
// The orthonormal 2-D DCT preserves singular values, and for D = C B C^T with
// D = U S V^T the reference's output block is
//     IDCT(U diag(S + d e0) V^T) = B + d * (C^T u0)(C^T v0)^T = B + d * u0_B v0_B^T
// (modules/watermarking.py:192-204), where (s0, u0_B, v0_B) is the top singular
// triplet of the *spatial* block B.  So embed needs one singular triplet and a
// rank-1 update, extract needs one singular value; no DCT, no full SVD.
//
// Top triplet, robustly: G = B^T B (symmetric PSD, 36 unique entries), scaled
// to unit trace, then repeated squaring P = M*M.  With tr(M) = 1,
//     1 - tr(P) = sum_i mu_i (1 - mu_i) >= mu_1/mu_0-ish
// is a rigorous (up to rounding) bound on how far M is from rank one, whatever
// the spectral gap, and every squaring squares the eigenvalue ratios.  When the
// bound is met, the column of P with the largest diagonal entry is v0 to
// ~(bound/2)^2; sigma0 = ||B v0|| (a Rayleigh quotient: second-order accurate)
// and u0 = B v0 / sigma0.
//
// Pixels are carried in 0..255 units (exact small integers as floats); colour
// math is fp32 FMA with the reference's constants folded - accurate to ~2 ulp
// of the reference's float64-dot-then-float32 values, not bit-identical (the
// faithful mode is).  All of it is CUDA-core fp32: the contractions are 8 wide.
#pragma once
#include "tmf_math.cuh"

namespace tmf {

// embed must resolve u0 v0^T to ~1e-6; extract only needs sigma0, whose error is
// second order in the vector error, so it can stop a squaring earlier.
#define TMF_FAST_TOL_EMBED 2.0e-3f
#define TMF_FAST_TOL_EXTRACT 3.0e-2f
#define TMF_FAST_MAX_SQUARINGS 18

// upper triangle of a symmetric 8x8 in 36 registers: index of (i, j), i <= j
#define TMF_SYM(i, j) ((i) * 8 - ((i) * ((i) + 1)) / 2 + (j))

// g = B^T B (upper triangle); returns trace
TMF_HD float gram_upper(const float* b, float* g) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = i; j < 8; ++j) {
      float s = 0.f;
#pragma unroll
      for (int r = 0; r < 8; ++r) s = fmaf(b[8 * r + i], b[8 * r + j], s);
      g[TMF_SYM(i, j)] = s;
    }
  float tr = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tr += g[TMF_SYM(i, i)];
  return tr;
}

// p = m * m for symmetric m (upper triangles); returns tr(p)
TMF_HD float sym_square(const float* m, float* p) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = i; j < 8; ++j) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float a = m[(i <= k) ? TMF_SYM(i, k) : TMF_SYM(k, i)];
        const float c = m[(k <= j) ? TMF_SYM(k, j) : TMF_SYM(j, k)];
        s = fmaf(a, c, s);
      }
      p[TMF_SYM(i, j)] = s;
    }
  float tr = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tr += p[TMF_SYM(i, i)];
  return tr;
}

// Top right-singular vector of B (unit norm) from its Gram matrix by repeated
// squaring.  g is destroyed.  Returns the number of squarings.  tr must be > 0.
TMF_HD int top_eigvec_by_squaring(float* g, float tr, float tol, float* v) {
  float p[36];
  float inv = f_rcp_fast(tr);
#pragma unroll
  for (int k = 0; k < 36; ++k) g[k] *= inv;
  int it = 0;
  for (;;) {
    const float t = sym_square(g, p);     // tr(g) == 1, so t = sum mu_i^2
    ++it;
    if (1.0f - t <= tol || it >= TMF_FAST_MAX_SQUARINGS) break;
    inv = f_rcp_fast(t);
#pragma unroll
    for (int k = 0; k < 36; ++k) g[k] = p[k] * inv;
  }
  // column of p with the largest diagonal entry (>= 1/8 of the trace)
  float best = p[TMF_SYM(0, 0)];
  int col = 0;
#pragma unroll
  for (int j = 1; j < 8; ++j) {
    const float d = p[TMF_SYM(j, j)];
    if (d > best) { best = d; col = j; }
  }
  float n2 = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float x = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float e = p[(i <= j) ? TMF_SYM(i, j) : TMF_SYM(j, i)];
      x = (j == col) ? e : x;
    }
    v[i] = x;
    n2 = fmaf(x, x, n2);
  }
  float rn = f_rsqrt(n2);
  rn = fmaf(0.5f * rn, fmaf(-n2 * rn, rn, 1.0f), rn);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] *= rn;
  return it;
}

// sigma0 and (optionally) u0 = B v0 / sigma0 from a unit v0
template <bool WITH_U>
TMF_HD float sigma_from_v(const float* b, const float* v, float* u) {
  float z[8], n2 = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s = fmaf(b[8 * i + j], v[j], s);
    z[i] = s;
    n2 = fmaf(s, s, n2);
  }
  const float sig = f_sqrt(n2);
  if (WITH_U) {
    const float inv = f_div(1.0f, sig);
#pragma unroll
    for (int i = 0; i < 8; ++i) u[i] = z[i] * inv;
  }
  return sig;
}

// Largest singular value of a spatial luma block given in 0..255 units; result
// in the reference's units (luma in [0, 1]).
TMF_HD float sigma0_block_fast(const float* b255, int* squarings) {
  float g[36], v[8];
  const float tr = gram_upper(b255, g);
  if (squarings) *squarings = 0;
  if (!(tr > 0.0f)) return 0.0f;
  const int it = top_eigvec_by_squaring(g, tr, TMF_FAST_TOL_EXTRACT, v);
  if (squarings) *squarings = it;
  return sigma_from_v<false>(b255, v, nullptr) * (1.0f / 255.0f);
}

// Embed on a spatial luma block in 0..255 units, in place: b += d255 * u0 v0^T,
// d255 = 255 * (f32(f64(s0) + alpha*w) - s0)   (watermarking.py:198).
// Returns sigma0 in the reference's units.
TMF_HD float embed_block_fast(float* b255, double alpha, uint32_t wm_u8, int* squarings) {
  float g[36], v[8], u[8];
  const float tr = gram_upper(b255, g);
  float sig = 0.0f;
  if (tr > 0.0f) {
    const int it = top_eigvec_by_squaring(g, tr, TMF_FAST_TOL_EMBED, v);
    if (squarings) *squarings = it;
    sig = sigma_from_v<true>(b255, v, u) * (1.0f / 255.0f);
  } else {
    // all-zero block: LAPACK returns U = V = I in the DCT domain, i.e. the DC
    // basis function, which is the constant 1/sqrt(8) vector in the spatial domain
    if (squarings) *squarings = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { u[i] = TMF_G0; v[i] = TMF_G0; }
  }
  const float sig_new = modulate_sigma0(sig, alpha, wm_u8);
  const float d255 = (sig_new - sig) * 255.0f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float du = d255 * u[i];
#pragma unroll
    for (int j = 0; j < 8; ++j) b255[8 * i + j] = fmaf(du, v[j], b255[8 * i + j]);
  }
  return sig;
}

// --- fp32 colour in 0..255 units -------------------------------------------
// luma of watermarking.py:37-45 times 255
TMF_HD float luma255_fast(float r, float g, float b) {
  return fmaf(0.299f, r, fmaf(0.587f, g, 0.114f * b));
}
// watermarking.py:37-48 and :58-73 composed: the reference maps (r, g, b) to
// (y, cb, cr), replaces y by y', and maps back with a matrix that is not the
// exact inverse.  out_c = y' + k_c . (cb, cr), all in 0..255 units, then clip
// and truncate.  `dy` = y' - y.
TMF_HD void rgb255_out_fast(float r, float g, float b, float y_new, float& R, float& G, float& B) {
  const float cb = fmaf(-0.169f, r, fmaf(-0.331f, g, 0.5f * b));
  const float cr = fmaf(0.5f, r, fmaf(-0.419f, g, -0.081f * b));
  R = fmaf(1.403f, cr, y_new);
  G = fmaf(-0.714f, cr, fmaf(-0.344f, cb, y_new));
  B = fmaf(1.773f, cb, y_new);
}
// clip [0, 255] and truncate toward zero (values are >= 0 after the clip)
TMF_HD uint32_t quant255(float x) {
  x = fminf(fmaxf(x, 0.0f), 255.0f);
  return (uint32_t)x;
}

}  // namespace tmf
