// FAITHFUL mode: the reference's pipeline executed as written - bit-exact colour (float64 FMA
// chain in the reference's accumulation order), orthonormal 2-D DCT, one-sided Jacobi SVD of the
// DCT block, S[0] += alpha * w, reconstruction, IDCT, bit-exact colour out
// (modules/watermarking.py:163-219, :246-289).  One thread per block.
//
// Reconstruction.  The reference changes S[0] only (:198), so U diag(S') V^T = D + d u0 v0^T with
// d = S'[0] - S[0]; the DCT is orthonormal, so IDCT of that is B + d (C^T u0)(C^T v0)^T.  The
// default faithful kernels therefore never accumulate V: the Jacobi runs in its values-only form
// (A <- A V = U diag(sigma); half the rotation work), u0 is the normalised top column, and
// C^T v0 = B^T (C^T u0) / sigma0 comes from the block itself, which is still parked in shared
// memory.  TMF_MODE_LITERAL (block size 8) keeps the literal U diag(S') V^T product and IDCT
// for comparison.
//
//   block size 8:    A (and V) in registers, packed-fp32 Jacobi rounds (tmf_math.cuh)
//   other sizes:     4..16 do not fit the register file (N^2 up to 256 values), so the block and
//                    the matrix under rotation live in thread-private shared-memory columns
//                    (conflict-free: element k of thread t at [k * T + t]); DCT by the N x N cosine
//                    matrix built once per CTA; cyclic one-sided Jacobi with runtime pair loops.
#include <atomic>
#include <type_traits>

#include "tmf_common.cuh"
#include "tmf_math.cuh"

namespace tmfi {
namespace {

// f32(byte) / 255.0 of byte B of a row held in words: the byte is extracted straight onto the 2^23
// magic number (one PRMT; a shift + mask + or would be two logic operations)
template <int NW>
__device__ __forceinline__ float unit_of_byte(const uint32_t (&w)[NW], int B) {
  const float km = __uint_as_float(__byte_perm(w[B >> 2], 0x4B000000u, 0x7650u | (uint32_t)(B & 3)));
  return tmf::unit_from_float_byte(km - 8388608.0f);
}

// ---------------------------------------------------------------------------
// block size 8
// ---------------------------------------------------------------------------
// The DCT / Jacobi need the whole block in registers with compile-time indices, but the
// per-pixel colour code is long (float64 dots, exact division): unrolled over 64 pixels it
// made the kernel 230 KB of instructions and it stalled on instruction fetch.  The row loops
// are therefore rolled and exchange the block with the register file through a thread-private
// column of shared memory: sm[k * kThreads + tid].
template <int VEC>
__device__ __forceinline__ void luma_rows_to_smem(const uint8_t* __restrict__ base, size_t pitch, float* __restrict__ col) {
#pragma unroll 1
  for (int i = 0; i < 8; ++i) {
    uint32_t w[6];
    load_row24<VEC>(base + (size_t)i * pitch, w);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float r = unit_of_byte(w, 3 * j);
      const float g = unit_of_byte(w, 3 * j + 1);
      const float b = unit_of_byte(w, 3 * j + 2);
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

// colour out: chroma is recomputed from the (L1/L2-resident) input bytes, luma from `col`
template <int VEC>
__device__ __forceinline__ void colour_rows_out(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, size_t pitch,
                                                const float* __restrict__ col) {
#pragma unroll 1
  for (int i = 0; i < 8; ++i) {
    uint32_t w[6], o[6];
    int q[24];
    load_row24<VEC>(src + (size_t)i * pitch, w);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float r = unit_of_byte(w, 3 * j);
      const float gg = unit_of_byte(w, 3 * j + 1);
      const float b = unit_of_byte(w, 3 * j + 2);
      float cb, cr;
      tmf::chroma_exact(r, gg, b, cb, cr);
      tmf::ycc_to_levels_exact(col[(8 * i + j) * kThreads], cb, cr, q[3 * j], q[3 * j + 1], q[3 * j + 2]);
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) o[k] = tmf::pack4_sat_u8(q[4 * k], q[4 * k + 1], q[4 * k + 2], q[4 * k + 3]);   // the clip of :70
    store_row24<VEC>(dst + (size_t)i * pitch, o);
  }
}

template <int VEC, bool LITERAL>
__global__ void __launch_bounds__(kThreads, LITERAL ? TMF_FAITHFUL_MIN_CTAS : TMF_FAITHFUL_R1_MIN_CTAS)
k_embed_faithful(const uint8_t* __restrict__ rgb, uint8_t* __restrict__ out, BlockGeom g,
                 const uint8_t* __restrict__ wm, int wm_shared, double alpha) {
  __shared__ float sm[64 * kThreads];
  const long long gb = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (gb >= g.total_blocks) return;
  long long img; int by, bx;
  uint32_t in_img;
  const size_t org = block_origin<8>(g, gb, img, by, bx, &in_img);
  const uint8_t* src = rgb + org;
  float* col = sm + threadIdx.x;
  prefetch_block_rows(src, g.pitch32);
  const uint32_t mark = (uint32_t)__ldg(wm + (wm_shared ? in_img : (uint32_t)gb));

  float a[64];
  load_luma_block<VEC>(src, g.row_pitch, col, a);
  if (LITERAL) {
    float v[64];
    tmf::embed_block_faithful(a, v, alpha, mark, nullptr);
#pragma unroll
    for (int k = 0; k < 64; ++k) col[k * kThreads] = a[k];
  } else {
    float uB[8];
    const float sig = tmf::top_left_vector_faithful(a, uB, nullptr);
    const float d = tmf::f_add(tmf::modulate_sigma0(sig, alpha, mark), -sig);
    if (d != 0.0f) {
      // vB = B^T uB / sigma0 from the parked block (all-zero block: vB = uB = 1/sqrt(8))
      float vB[8];
      if (sig > 0.0f) {
        const float inv = tmf::f_div(1.0f, sig);
#pragma unroll
        for (int j = 0; j < 8; ++j) vB[j] = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) vB[j] = fmaf(uB[i], col[(8 * i + j) * kThreads], vB[j]);
#pragma unroll
        for (int j = 0; j < 8; ++j) vB[j] *= inv;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) vB[j] = uB[j];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float du = d * uB[i];
#pragma unroll
        for (int j = 0; j < 8; ++j) col[(8 * i + j) * kThreads] = fmaf(du, vB[j], col[(8 * i + j) * kThreads]);
      }
    }
  }
  colour_rows_out<VEC>(src, out + org, g.row_pitch, col);
}

template <int VEC>
__global__ void __launch_bounds__(kThreads, TMF_FAITHFUL_SIGMA_MIN_CTAS)
k_extract_faithful(const uint8_t* __restrict__ wmk, const uint8_t* __restrict__ orig, uint8_t* __restrict__ out_wm,
                   BlockGeom g, double alpha) {
  const long long gb = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (gb >= g.total_blocks) return;
  long long img; int by, bx;
  const size_t org = block_origin<8>(g, gb, img, by, bx);
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
__global__ void __launch_bounds__(kThreads, TMF_FAITHFUL_SIGMA_MIN_CTAS)
k_sigma0_faithful(const uint8_t* __restrict__ rgb, float* __restrict__ sigma0, BlockGeom g) {
  const long long gb = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (gb >= g.total_blocks) return;
  long long img; int by, bx;
  const size_t org = block_origin<8>(g, gb, img, by, bx);
  __shared__ float sm[64 * kThreads];
  float* col = sm + threadIdx.x;
  float a[64];
  load_luma_block<VEC>(rgb + org, g.row_pitch, col, a);
  sigma0[gb] = tmf::sigma0_block_faithful(a, nullptr);
}

// ---------------------------------------------------------------------------
// other block sizes (4, 6, 10, 12, 14, 16): shared-memory resident
// ---------------------------------------------------------------------------
// threads per CTA: the embed kernel parks two N x N float matrices per thread
template <int N> __host__ __device__ constexpr int fn_threads() { return N <= 6 ? 128 : 32; }

// orthonormal DCT-II matrix C[k][n] = s_k cos(pi (2n+1) k / 2N), s_0 = sqrt(1/N), s_k = sqrt(2/N)
// (scipy.fftpack.dct(norm="ortho"), watermarking.py:78), built once per CTA in float64
template <int N, int T>
__device__ __forceinline__ void build_dct_matrix(float* __restrict__ C) {
  for (int idx = threadIdx.x; idx < N * N; idx += T) {
    const int k = idx / N, n = idx - k * N;
    const double s = (k == 0) ? sqrt(1.0 / N) : sqrt(2.0 / N);
    C[idx] = (float)(s * cospi((double)((2 * n + 1) * k) / (double)(2 * N)));
  }
  __syncthreads();
}

// x <- C x (forward) or C^T x (inverse) for the N-vector m[base + i * step], i = 0..N-1
template <int N, int T, bool INVERSE>
__device__ __forceinline__ void dct_vec_smem(float* __restrict__ m, int base, int step, const float* __restrict__ C) {
  float x[N];
#pragma unroll
  for (int i = 0; i < N; ++i) x[i] = m[(base + i * step) * T];
#pragma unroll 1
  for (int k = 0; k < N; ++k) {
    float s = 0.0f;
#pragma unroll
    for (int n = 0; n < N; ++n) s = fmaf(INVERSE ? C[n * N + k] : C[k * N + n], x[n], s);
    m[(base + k * step) * T] = s;
  }
}

// 2-D transform of the row-major block m (element (i, j) at m[(i * N + j) * T]): columns first,
// then rows, as watermarking.py:78
template <int N, int T, bool INVERSE>
__device__ __forceinline__ void dct2_smem(float* __restrict__ m, const float* __restrict__ C) {
#pragma unroll 1
  for (int j = 0; j < N; ++j) dct_vec_smem<N, T, INVERSE>(m, j, N, C);
#pragma unroll 1
  for (int i = 0; i < N; ++i) dct_vec_smem<N, T, INVERSE>(m, i * N, 1, C);
}

// Cyclic one-sided Jacobi on the row-major N x N matrix m in shared memory: m <- m V, columns
// mutually orthogonal on return (column k = sigma_k u_k) - or, when one column dominates, only the largest column
// is orthogonal to the rest (see inside).  Same rotation, skip rules, scaling
// and stop rule as tmf::jacobi_svd8.  Returns the power-of-two `unscale` of the column norms.
template <int N, int T>
__device__ __forceinline__ float jacobi_smem(float* __restrict__ m) {
  // column norms: their sum scales the block, the largest may certify a dominant column
  float frob2 = 0.0f, best = -1.0f;
  int top = 0;
#pragma unroll 1
  for (int j = 0; j < N; ++j) {
    float sj = 0.0f;
#pragma unroll
    for (int i = 0; i < N; ++i) sj = fmaf(m[(i * N + j) * T], m[(i * N + j) * T], sj);
    frob2 += sj;
    if (sj > best) { best = sj; top = j; }
  }
  float unscale = 1.0f;
  if (!((frob2 > 0.0f) && (frob2 < INFINITY))) return unscale;
  const float sc = tmf::pow2_scale_for(frob2, unscale);
#pragma unroll 1
  for (int k = 0; k < N * N; ++k) m[k * T] *= sc;

  if (best > TMF_JACOBI_TOP_DOMINANCE * (frob2 - best)) {
    // Dominant column only (tmf::top_column8 has the argument): p stays in registers and is rotated against
    // each of the other N - 1 columns, N - 1 rotations per sweep instead of N (N - 1) / 2, until it is
    // orthogonal to all of them.  The other columns are NOT mutually orthogonal on return; embed and
    // extract read the largest column only.
    float xp[N];
#pragma unroll
    for (int i = 0; i < N; ++i) xp[i] = m[(i * N + top) * T];
#pragma unroll 1
    for (int sweep = 0; sweep < TMF_JACOBI_MAX_SWEEPS + 4; ++sweep) {
      float worst = 0.0f, al = 0.0f;
#pragma unroll
      for (int i = 0; i < N; ++i) al = fmaf(xp[i], xp[i], al);
#pragma unroll 1
      for (int q = 0; q < N; ++q) {
        if (q == top) continue;
        float xq[N], be = 0.0f, ga = 0.0f;
#pragma unroll
        for (int i = 0; i < N; ++i) {
          xq[i] = m[(i * N + q) * T];
          be = fmaf(xq[i], xq[i], be);
          ga = fmaf(xp[i], xq[i], ga);
        }
        float c, s, t;
        const float code = tmf::jacobi_cs(al, be, ga, c, s, &t, TMF_JACOBI_TOP_DONE * TMF_JACOBI_TOP_DONE);
        worst = fmaxf(worst, code);
        if (code > 0.0f) {
#pragma unroll
          for (int i = 0; i < N; ++i) {
            m[(i * N + q) * T] = fmaf(s, xp[i], c * xq[i]);
            xp[i] = fmaf(c, xp[i], -s * xq[i]);
          }
          al = fmaxf(al - t * ga, 0.0f);
        }
      }
      if (!(worst > TMF_JACOBI_MORE)) break;
    }
#pragma unroll
    for (int i = 0; i < N; ++i) m[(i * N + top) * T] = xp[i];
    return unscale;
  }

#pragma unroll 1
  for (int sweep = 0; sweep < TMF_JACOBI_MAX_SWEEPS + 4; ++sweep) {     // larger N: a few more sweeps
    float worst = 0.0f;
#pragma unroll 1
    for (int p = 0; p < N - 1; ++p) {
#pragma unroll 1
      for (int q = p + 1; q < N; ++q) {
        float xp[N], xq[N], al = 0.0f, be = 0.0f, ga = 0.0f;
#pragma unroll
        for (int i = 0; i < N; ++i) {
          xp[i] = m[(i * N + p) * T];
          xq[i] = m[(i * N + q) * T];
          al = fmaf(xp[i], xp[i], al);
          be = fmaf(xq[i], xq[i], be);
          ga = fmaf(xp[i], xq[i], ga);
        }
        float c, s;
        const float cosv = tmf::jacobi_cs(al, be, ga, c, s);
        worst = fmaxf(worst, cosv);
        if (cosv > 0.0f) {
#pragma unroll
          for (int i = 0; i < N; ++i) {
            m[(i * N + p) * T] = fmaf(c, xp[i], -s * xq[i]);
            m[(i * N + q) * T] = fmaf(s, xp[i], c * xq[i]);
          }
        }
      }
    }
    if (!(worst > TMF_JACOBI_MORE)) break;
  }
  return unscale;
}

// squared norm of the largest column of m and its index
template <int N, int T>
__device__ __forceinline__ float top_column(const float* __restrict__ m, int& top) {
  float best = -1.0f;
  top = 0;
#pragma unroll 1
  for (int j = 0; j < N; ++j) {
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < N; ++i) s = fmaf(m[(i * N + j) * T], m[(i * N + j) * T], s);
    if (s > best) { best = s; top = j; }
  }
  return best;
}

// exact luma of the block at `base` into the row-major shared-memory matrix m
template <int N, int T, int AL>
__device__ __forceinline__ void luma_block_to_smem(const uint8_t* __restrict__ base, size_t pitch, float* __restrict__ m) {
#pragma unroll 1
  for (int i = 0; i < N; ++i) {
    uint32_t w[kRowWords<N>];
    load_row_n<N, AL>(base + (size_t)i * pitch, w);
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const float r = unit_of_byte(w, 3 * j);
      const float g = unit_of_byte(w, 3 * j + 1);
      const float b = unit_of_byte(w, 3 * j + 2);
      m[(i * N + j) * T] = tmf::luma_exact(r, g, b);
    }
  }
}

// largest singular value of the DCT of the luma block at `base` (m: N*N*T floats of scratch)
template <int N, int T, int AL>
__device__ __forceinline__ float sigma0_faithful_n(const uint8_t* __restrict__ base, size_t pitch, float* __restrict__ m,
                                                   const float* __restrict__ C) {
  luma_block_to_smem<N, T, AL>(base, pitch, m);
  dct2_smem<N, T, false>(m, C);
  const float unscale = jacobi_smem<N, T>(m);
  int top;
  const float best = top_column<N, T>(m, top);
  return tmf::f_sqrt(fmaxf(best, 0.0f)) * unscale;
}

template <int N, int AL>
__global__ void __launch_bounds__(fn_threads<N>())
k_embed_faithful_n(const uint8_t* __restrict__ rgb, uint8_t* __restrict__ out, BlockGeom g,
                   const uint8_t* __restrict__ wm, int wm_shared, double alpha) {
  constexpr int T = fn_threads<N>();
  extern __shared__ __align__(16) float fsm[];
  float* C = fsm;                                  // N*N
  float* Bm = fsm + N * N + threadIdx.x;           // the luma block, kept
  float* Am = Bm + N * N * T;                      // DCT block -> A V
  build_dct_matrix<N, T>(C);
  const long long gb = (long long)blockIdx.x * T + threadIdx.x;
  if (gb >= g.total_blocks) return;
  long long img; int by, bx;
  uint32_t in_img;
  const size_t org = block_origin<N>(g, gb, img, by, bx, &in_img);
  const uint8_t* src = rgb + org;
  const uint32_t mark = (uint32_t)__ldg(wm + (wm_shared ? in_img : (uint32_t)gb));

  luma_block_to_smem<N, T, AL>(src, g.row_pitch, Bm);
#pragma unroll 1
  for (int k = 0; k < N * N; ++k) Am[k * T] = Bm[k * T];
  dct2_smem<N, T, false>(Am, C);
  const float unscale = jacobi_smem<N, T>(Am);
  int top;
  const float best = top_column<N, T>(Am, top);
  const float nrm = tmf::f_sqrt(fmaxf(best, 0.0f));
  const float sig = nrm * unscale;
  const float d = tmf::f_add(tmf::modulate_sigma0(sig, alpha, mark), -sig);
  if (d != 0.0f) {
    float uB[N], vB[N];
    if (sig > 0.0f) {
      float u[N];
      const float inv = tmf::f_div(1.0f, nrm);
#pragma unroll
      for (int i = 0; i < N; ++i) u[i] = Am[(i * N + top) * T] * inv;
#pragma unroll
      for (int n = 0; n < N; ++n) {                 // uB = C^T u0
        float s = 0.0f;
#pragma unroll
        for (int k = 0; k < N; ++k) s = fmaf(C[k * N + n], u[k], s);
        uB[n] = s;
      }
      const float isig = tmf::f_div(1.0f, sig);
#pragma unroll
      for (int j = 0; j < N; ++j) vB[j] = 0.0f;
#pragma unroll 1
      for (int i = 0; i < N; ++i) {
        float ui = 0.0f;
#pragma unroll
        for (int k = 0; k < N; ++k) ui = (k == i) ? uB[k] : ui;
#pragma unroll
        for (int j = 0; j < N; ++j) vB[j] = fmaf(ui, Bm[(i * N + j) * T], vB[j]);
      }
#pragma unroll
      for (int j = 0; j < N; ++j) vB[j] *= isig;
    } else {                                        // all-zero block: LAPACK's U = V = I, the mark lands on DC
      const float dc = tmf::f_sqrt(1.0f / (float)N);
#pragma unroll
      for (int j = 0; j < N; ++j) { uB[j] = dc; vB[j] = dc; }
    }
#pragma unroll 1
    for (int i = 0; i < N; ++i) {
      float ui = 0.0f;
#pragma unroll
      for (int k = 0; k < N; ++k) ui = (k == i) ? uB[k] : ui;
      const float du = d * ui;
#pragma unroll
      for (int j = 0; j < N; ++j) Bm[(i * N + j) * T] = fmaf(du, vB[j], Bm[(i * N + j) * T]);
    }
  }
  // colour out
  uint8_t* dst = out + org;
#pragma unroll 1
  for (int i = 0; i < N; ++i) {
    uint32_t w[kRowWords<N>], o[kRowWords<N>];
    int q[4 * kRowWords<N>];
    load_row_n<N, AL>(src + (size_t)i * g.row_pitch, w);
#pragma unroll
    for (int k = 3 * N; k < 4 * kRowWords<N>; ++k) q[k] = 0;
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const float r = unit_of_byte(w, 3 * j);
      const float gg = unit_of_byte(w, 3 * j + 1);
      const float b = unit_of_byte(w, 3 * j + 2);
      float cb, cr;
      tmf::chroma_exact(r, gg, b, cb, cr);
      tmf::ycc_to_levels_exact(Bm[(i * N + j) * T], cb, cr, q[3 * j], q[3 * j + 1], q[3 * j + 2]);
    }
#pragma unroll
    for (int k = 0; k < kRowWords<N>; ++k) o[k] = tmf::pack4_sat_u8(q[4 * k], q[4 * k + 1], q[4 * k + 2], q[4 * k + 3]);
    store_row_n<N, AL>(dst + (size_t)i * g.row_pitch, o);
  }
}

template <int N, int AL>
__global__ void __launch_bounds__(fn_threads<N>())
k_extract_faithful_n(const uint8_t* __restrict__ wmk, const uint8_t* __restrict__ orig, uint8_t* __restrict__ out_wm,
                     BlockGeom g, double alpha) {
  constexpr int T = fn_threads<N>();
  extern __shared__ __align__(16) float fsm[];
  float* C = fsm;
  float* m = fsm + N * N + threadIdx.x;
  build_dct_matrix<N, T>(C);
  const long long gb = (long long)blockIdx.x * T + threadIdx.x;
  if (gb >= g.total_blocks) return;
  long long img; int by, bx;
  const size_t org = block_origin<N>(g, gb, img, by, bx);
  float sw = 0.0f, so = 0.0f;
#pragma unroll 1
  for (int which = 0; which < 2; ++which) {
    const float sg = sigma0_faithful_n<N, T, AL>((which == 0 ? wmk : orig) + org, g.row_pitch, m, C);
    if (which == 0) sw = sg; else so = sg;
  }
  out_wm[gb] = (uint8_t)tmf::extract_level(sw, so, alpha);
}

template <int N, int AL>
__global__ void __launch_bounds__(fn_threads<N>())
k_sigma0_faithful_n(const uint8_t* __restrict__ rgb, float* __restrict__ sigma0, BlockGeom g) {
  constexpr int T = fn_threads<N>();
  extern __shared__ __align__(16) float fsm[];
  float* C = fsm;
  float* m = fsm + N * N + threadIdx.x;
  build_dct_matrix<N, T>(C);
  const long long gb = (long long)blockIdx.x * T + threadIdx.x;
  if (gb >= g.total_blocks) return;
  long long img; int by, bx;
  const size_t org = block_origin<N>(g, gb, img, by, bx);
  sigma0[gb] = sigma0_faithful_n<N, T, AL>(rgb + org, g.row_pitch, m, C);
}

template <int N> __host__ __device__ constexpr int fn_smem(int matrices) { return (N * N + matrices * N * N * fn_threads<N>()) * 4; }

template <int N, typename F>
void with_access_width(int al, F&& f) {
  if (al == 4 && (3 * N) % 4 == 0) f(std::integral_constant<int, N>{}, std::integral_constant<int, 4>{});
  else if (al >= 2) f(std::integral_constant<int, N>{}, std::integral_constant<int, 2>{});
  else f(std::integral_constant<int, N>{}, std::integral_constant<int, 1>{});
}
template <typename F>
void for_block_size(int n, int al, F&& f) {
  switch (n) {
    case 4: with_access_width<4>(al, f); break;
    case 6: with_access_width<6>(al, f); break;
    case 10: with_access_width<10>(al, f); break;
    case 12: with_access_width<12>(al, f); break;
    case 14: with_access_width<14>(al, f); break;
    case 16: with_access_width<16>(al, f); break;
    default: break;
  }
}

// dynamic shared memory above 48 KB needs an opt-in per kernel (and device); cheap, so done per call
template <typename K>
int allow_smem(K kernel, int bytes) {
  if (bytes <= 48 * 1024) return TMF_OK;
  TMF_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return TMF_OK;
}

}  // namespace

int launch_embed_faithful(const uint8_t* rgb, uint8_t* out, const BlockGeom& g, const uint8_t* wm, int wm_shared,
                          double alpha, int literal, cudaStream_t st) {
  if (g.bs == 8) {
    const unsigned grid = grid_for(g.total_blocks, kThreads);
    const int vec = pick_vec(g, rgb, out);
    if (literal) {
      switch (vec) {
        case 8: k_embed_faithful<8, true><<<grid, kThreads, 0, st>>>(rgb, out, g, wm, wm_shared, alpha); break;
        case 4: k_embed_faithful<4, true><<<grid, kThreads, 0, st>>>(rgb, out, g, wm, wm_shared, alpha); break;
        default: k_embed_faithful<1, true><<<grid, kThreads, 0, st>>>(rgb, out, g, wm, wm_shared, alpha); break;
      }
    } else {
      switch (vec) {
        case 8: k_embed_faithful<8, false><<<grid, kThreads, 0, st>>>(rgb, out, g, wm, wm_shared, alpha); break;
        case 4: k_embed_faithful<4, false><<<grid, kThreads, 0, st>>>(rgb, out, g, wm, wm_shared, alpha); break;
        default: k_embed_faithful<1, false><<<grid, kThreads, 0, st>>>(rgb, out, g, wm, wm_shared, alpha); break;
      }
    }
    return check_launch("embed (faithful) kernel launch");
  }
  if (literal)
    return fail(TMF_ERR_BAD_ARG, "TMF_MODE_LITERAL is implemented for block size 8 only (block %d: use TMF_MODE_FAITHFUL)", g.bs);
  int rc = TMF_OK;
  for_block_size(g.bs, row_alignment(g, rgb, out), [&](auto n_, auto a_) {
    constexpr int N = decltype(n_)::value, A = decltype(a_)::value;
    constexpr int T = fn_threads<N>(), smem = fn_smem<N>(2);
    auto kernel = k_embed_faithful_n<N, A>;
    if ((rc = allow_smem(kernel, smem)) != TMF_OK) return;
    kernel<<<grid_for(g.total_blocks, T), T, smem, st>>>(rgb, out, g, wm, wm_shared, alpha);
  });
  if (rc) return rc;
  return check_launch("embed (faithful) kernel launch");
}

int launch_extract_faithful(const uint8_t* wmk, const uint8_t* orig, uint8_t* out_wm, const BlockGeom& g,
                            double alpha, cudaStream_t st) {
  if (g.bs == 8) {
    const unsigned grid = grid_for(g.total_blocks, kThreads);
    switch (pick_vec(g, wmk, orig)) {
      case 8: k_extract_faithful<8><<<grid, kThreads, 0, st>>>(wmk, orig, out_wm, g, alpha); break;
      case 4: k_extract_faithful<4><<<grid, kThreads, 0, st>>>(wmk, orig, out_wm, g, alpha); break;
      default: k_extract_faithful<1><<<grid, kThreads, 0, st>>>(wmk, orig, out_wm, g, alpha); break;
    }
    return check_launch("extract (faithful) kernel launch");
  }
  int rc = TMF_OK;
  for_block_size(g.bs, row_alignment(g, wmk, orig), [&](auto n_, auto a_) {
    constexpr int N = decltype(n_)::value, A = decltype(a_)::value;
    constexpr int T = fn_threads<N>(), smem = fn_smem<N>(1);
    auto kernel = k_extract_faithful_n<N, A>;
    if ((rc = allow_smem(kernel, smem)) != TMF_OK) return;
    kernel<<<grid_for(g.total_blocks, T), T, smem, st>>>(wmk, orig, out_wm, g, alpha);
  });
  if (rc) return rc;
  return check_launch("extract (faithful) kernel launch");
}

int launch_sigma0_faithful(const uint8_t* rgb, float* sigma0, const BlockGeom& g, cudaStream_t st) {
  if (g.bs == 8) {
    const unsigned grid = grid_for(g.total_blocks, kThreads);
    switch (pick_vec(g, rgb, rgb)) {
      case 8: k_sigma0_faithful<8><<<grid, kThreads, 0, st>>>(rgb, sigma0, g); break;
      case 4: k_sigma0_faithful<4><<<grid, kThreads, 0, st>>>(rgb, sigma0, g); break;
      default: k_sigma0_faithful<1><<<grid, kThreads, 0, st>>>(rgb, sigma0, g); break;
    }
    return check_launch("sigma0 (faithful) kernel launch");
  }
  int rc = TMF_OK;
  for_block_size(g.bs, row_alignment(g, rgb, rgb), [&](auto n_, auto a_) {
    constexpr int N = decltype(n_)::value, A = decltype(a_)::value;
    constexpr int T = fn_threads<N>(), smem = fn_smem<N>(1);
    auto kernel = k_sigma0_faithful_n<N, A>;
    if ((rc = allow_smem(kernel, smem)) != TMF_OK) return;
    kernel<<<grid_for(g.total_blocks, T), T, smem, st>>>(rgb, sigma0, g);
  });
  if (rc) return rc;
  return check_launch("sigma0 (faithful) kernel launch");
}

}  // namespace tmfi
