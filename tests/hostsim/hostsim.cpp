// CPU TEST HARNESS for the kernel arithmetic in thatsmyface_b200/csrc/tmf_math.cuh.
// Built by tests/test_hostsim.py with g++ (-ffp-contract=off) so the per-block
// math the CUDA kernels run can be checked against the oracle without a GPU.
// It is test infrastructure: the product library (libtmfwm.so) never links or
// calls it, and the product has no CPU path.
#include <cstdint>
#include <cstring>
#include "../../thatsmyface_b200/csrc/tmf_math.cuh"
#include "../../thatsmyface_b200/csrc/tmf_fast.cuh"
#include "../../thatsmyface_b200/csrc/tmf_resize.cuh"

using namespace tmf;

static void load_luma(const uint8_t* rgb, int w, int by, int bx, float* a) {
  for (int i = 0; i < 8; ++i)
    for (int j = 0; j < 8; ++j) {
      const uint8_t* p = rgb + ((size_t)(by * 8 + i) * w + bx * 8 + j) * 3;
      a[8 * i + j] = luma_exact(unit_from_u8(p[0]), unit_from_u8(p[1]), unit_from_u8(p[2]));
    }
}

extern "C" {

int hostsim_embed(const uint8_t* rgb, uint8_t* out, int h, int w, const uint8_t* wm, double alpha,
                  float* sigma_out, int* sweeps_out) {
  const int nbh = h / 8, nbw = w / 8;
  // colour round trip for every pixel first (strips keep this value)
  for (size_t p = 0; p < (size_t)h * w; ++p) {
    float r = unit_from_u8(rgb[3 * p]), g = unit_from_u8(rgb[3 * p + 1]), b = unit_from_u8(rgb[3 * p + 2]);
    float cb, cr; chroma_exact(r, g, b, cb, cr);
    uint32_t R, G, B; ycc_to_rgb8_exact(luma_exact(r, g, b), cb, cr, R, G, B);
    out[3 * p] = (uint8_t)R; out[3 * p + 1] = (uint8_t)G; out[3 * p + 2] = (uint8_t)B;
  }
  for (int by = 0; by < nbh; ++by)
    for (int bx = 0; bx < nbw; ++bx) {
      float a[64], v[64];
      load_luma(rgb, w, by, bx, a);
      int sw;
      float sig = embed_block_faithful(a, v, alpha, wm[by * nbw + bx], &sw);
      if (sigma_out) sigma_out[by * nbw + bx] = sig;
      if (sweeps_out) sweeps_out[by * nbw + bx] = sw;
      for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 8; ++j) {
          size_t p = (size_t)(by * 8 + i) * w + bx * 8 + j;
          float r = unit_from_u8(rgb[3 * p]), g = unit_from_u8(rgb[3 * p + 1]), b = unit_from_u8(rgb[3 * p + 2]);
          float cb, cr; chroma_exact(r, g, b, cb, cr);
          uint32_t R, G, B; ycc_to_rgb8_exact(a[8 * i + j], cb, cr, R, G, B);
          out[3 * p] = (uint8_t)R; out[3 * p + 1] = (uint8_t)G; out[3 * p + 2] = (uint8_t)B;
        }
    }
  return 0;
}

// The product's FAITHFUL mode (k_embed_faithful<.., false>): DCT -> values-only Jacobi -> top column,
// reconstruction as B + d (C^T u0)(B^T C^T u0 / sigma0)^T - same order of operations as the kernel.
int hostsim_embed_rank1(const uint8_t* rgb, uint8_t* out, int h, int w, const uint8_t* wm, double alpha,
                        float* sigma_out, int* sweeps_out) {
  const int nbh = h / 8, nbw = w / 8;
  for (size_t p = 0; p < (size_t)h * w; ++p) {
    float r = unit_from_u8(rgb[3 * p]), g = unit_from_u8(rgb[3 * p + 1]), b = unit_from_u8(rgb[3 * p + 2]);
    float cb, cr; chroma_exact(r, g, b, cb, cr);
    uint32_t R, G, B; ycc_to_rgb8_exact(luma_exact(r, g, b), cb, cr, R, G, B);
    out[3 * p] = (uint8_t)R; out[3 * p + 1] = (uint8_t)G; out[3 * p + 2] = (uint8_t)B;
  }
  for (int by = 0; by < nbh; ++by)
    for (int bx = 0; bx < nbw; ++bx) {
      float a[64], blk[64], uB[8], vB[8];
      load_luma(rgb, w, by, bx, a);
      for (int k = 0; k < 64; ++k) blk[k] = a[k];
      int sw;
      const float sig = top_left_vector_faithful(a, uB, &sw);
      if (sigma_out) sigma_out[by * nbw + bx] = sig;
      if (sweeps_out) sweeps_out[by * nbw + bx] = sw;
      const float d = f_add(modulate_sigma0(sig, alpha, wm[by * nbw + bx]), -sig);
      if (d != 0.0f) {
        if (sig > 0.0f) {
          const float inv = f_div(1.0f, sig);
          for (int j = 0; j < 8; ++j) vB[j] = 0.0f;
          for (int i = 0; i < 8; ++i)
            for (int j = 0; j < 8; ++j) vB[j] = fmaf(uB[i], blk[8 * i + j], vB[j]);
          for (int j = 0; j < 8; ++j) vB[j] = f_mul(vB[j], inv);
        } else {
          for (int j = 0; j < 8; ++j) vB[j] = uB[j];
        }
        for (int i = 0; i < 8; ++i) {
          const float du = f_mul(d, uB[i]);
          for (int j = 0; j < 8; ++j) blk[8 * i + j] = fmaf(du, vB[j], blk[8 * i + j]);
        }
      }
      for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 8; ++j) {
          size_t p = (size_t)(by * 8 + i) * w + bx * 8 + j;
          float r = unit_from_u8(rgb[3 * p]), g = unit_from_u8(rgb[3 * p + 1]), b = unit_from_u8(rgb[3 * p + 2]);
          float cb, cr; chroma_exact(r, g, b, cb, cr);
          uint32_t R, G, B; ycc_to_rgb8_exact(blk[8 * i + j], cb, cr, R, G, B);
          out[3 * p] = (uint8_t)R; out[3 * p + 1] = (uint8_t)G; out[3 * p + 2] = (uint8_t)B;
        }
    }
  return 0;
}

int hostsim_extract(const uint8_t* wmk, const uint8_t* orig, uint8_t* out, int h, int w, double alpha) {
  const int nbh = h / 8, nbw = w / 8;
  for (int by = 0; by < nbh; ++by)
    for (int bx = 0; bx < nbw; ++bx) {
      float a[64];
      load_luma(wmk, w, by, bx, a);
      float sw = sigma0_block_faithful(a, nullptr);
      load_luma(orig, w, by, bx, a);
      float so = sigma0_block_faithful(a, nullptr);
      out[by * nbw + bx] = (uint8_t)extract_level(sw, so, alpha);
    }
  return 0;
}

static void row_rgb255(const uint8_t* rgb, int w, int y, int x0, float* r, float* g, float* b) {
  for (int j = 0; j < 8; ++j) {
    const uint8_t* p = rgb + ((size_t)y * w + x0 + j) * 3;
    r[j] = (float)p[0]; g[j] = (float)p[1]; b[j] = (float)p[2];
  }
}

static void gram_of_block(const uint8_t* rgb, int w, int by, int bx, float* gm, float* keep = nullptr) {
  for (int k = 0; k < 36; ++k) gm[k] = 0.0f;
  for (int i = 0; i < 8; ++i) {
    float r[8], g[8], b[8], y[8];
    row_rgb255(rgb, w, by * 8 + i, bx * 8, r, g, b);
    // the N = 8 kernels carry the exact integer luma 299 r + 587 g + 114 b (IDP.2A on the device)
    for (int j = 0; j < 8; ++j) y[j] = luma1000_exact((uint32_t)r[j], (uint32_t)g[j], (uint32_t)b[j]);
    if (keep) for (int j = 0; j < 8; ++j) keep[8 * i + j] = y[j];
    gram_accumulate_row(y, gm);
  }
}

int hostsim_embed_fast(const uint8_t* rgb, uint8_t* out, int h, int w, const uint8_t* wm, double alpha,
                       float* sigma_out, int* sweeps_out) {
  const int nbh = h / 8, nbw = w / 8;
  for (size_t p = 0; p < (size_t)h * w; ++p) {   // strips: exact colour round trip, as in the library
    float r = unit_from_u8(rgb[3 * p]), g = unit_from_u8(rgb[3 * p + 1]), b = unit_from_u8(rgb[3 * p + 2]);
    float cb, cr; chroma_exact(r, g, b, cb, cr);
    uint32_t R, G, B; ycc_to_rgb8_exact(luma_exact(r, g, b), cb, cr, R, G, B);
    out[3 * p] = (uint8_t)R; out[3 * p + 1] = (uint8_t)G; out[3 * p + 2] = (uint8_t)B;
  }
  for (int by = 0; by < nbh; ++by)
    for (int bx = 0; bx < nbw; ++bx) {
      float gm[36], wv[8], f, c, lum[64];
      gram_of_block(rgb, w, by, bx, gm, lum);
      int it;
      float sig = embed_block_scalars_fast(gm, alpha, wm[by * nbw + bx], wv, f, c, &it, TMF_LUMA1000_UNIT);
      if (sigma_out) sigma_out[by * nbw + bx] = sig;
      if (sweeps_out) sweeps_out[by * nbw + bx] = it;
      for (int i = 0; i < 8; ++i) {
        float r[8], g[8], b[8];
        int q[24];
        row_rgb255(rgb, w, by * 8 + i, bx * 8, r, g, b);
        embed_row_fast(r, g, b, lum + 8 * i, wv, f, c, q);
        uint8_t* dst = out + ((size_t)(by * 8 + i) * w + bx * 8) * 3;
        for (int k = 0; k < 6; ++k) {
          uint32_t word = pack4_sat_u8(q[4 * k], q[4 * k + 1], q[4 * k + 2], q[4 * k + 3]);
          for (int t = 0; t < 4; ++t) dst[4 * k + t] = (uint8_t)(word >> (8 * t));
        }
      }
    }
  return 0;
}

int hostsim_extract_fast(const uint8_t* wmk, const uint8_t* orig, uint8_t* out, int h, int w, double alpha) {
  const int nbh = h / 8, nbw = w / 8;
  for (int by = 0; by < nbh; ++by)
    for (int bx = 0; bx < nbw; ++bx) {
      float gm[36];
      gram_of_block(wmk, w, by, bx, gm);
      float sw = sigma0_from_gram_fast(gm, nullptr, TMF_LUMA1000_UNIT);
      gram_of_block(orig, w, by, bx, gm);
      float so = sigma0_from_gram_fast(gm, nullptr, TMF_LUMA1000_UNIT);
      out[by * nbw + bx] = (uint8_t)extract_level(sw, so, alpha);
    }
  return 0;
}

}  // extern "C"

// ---- other block sizes (generic-N templates of tmf_fast.cuh)
template <int N>
static void gram_of_block_n(const uint8_t* rgb, int w, int by, int bx, float* gm) {
  for (int k = 0; k < N * (N + 1) / 2; ++k) gm[k] = 0.0f;
  for (int i = 0; i < N; ++i) {
    float y[N];
    for (int j = 0; j < N; ++j) {
      const uint8_t* p = rgb + ((size_t)(by * N + i) * w + bx * N + j) * 3;
      y[j] = luma255_fast((float)p[0], (float)p[1], (float)p[2]);
    }
    gram_accumulate_row<N>(y, gm);
  }
}

template <int N>
static void embed_n(const uint8_t* rgb, uint8_t* out, int h, int w, const uint8_t* wm, double alpha) {
  const int nbh = h / N, nbw = w / N;
  for (int by = 0; by < nbh; ++by)
    for (int bx = 0; bx < nbw; ++bx) {
      float gm[N * (N + 1) / 2], wv[N], f = 0.f, c = 0.f;
      for (int i = 0; i < N; ++i) wv[i] = 0.f;
      const uint32_t mark = wm[by * nbw + bx];
      if (mark) {
        gram_of_block_n<N>(rgb, w, by, bx, gm);
        embed_block_scalars_fast<N>(gm, alpha, mark, wv, f, c, nullptr);
      }
      for (int i = 0; i < N; ++i) {
        float r[N], g[N], b[N], y[N];
        int q[3 * N];
        for (int j = 0; j < N; ++j) {
          const uint8_t* p = rgb + ((size_t)(by * N + i) * w + bx * N + j) * 3;
          r[j] = (float)p[0]; g[j] = (float)p[1]; b[j] = (float)p[2];
          y[j] = luma255_fast(r[j], g[j], b[j]);
        }
        embed_row_fast<N>(r, g, b, y, wv, f, c, q);
        uint8_t* d = out + ((size_t)(by * N + i) * w + bx * N) * 3;
        for (int k = 0; k < 3 * N; ++k) d[k] = (uint8_t)(q[k] < 0 ? 0 : (q[k] > 255 ? 255 : q[k]));
      }
    }
}

template <int N>
static void extract_n(const uint8_t* wmk, const uint8_t* orig, uint8_t* out, int h, int w, double alpha) {
  const int nbh = h / N, nbw = w / N;
  for (int by = 0; by < nbh; ++by)
    for (int bx = 0; bx < nbw; ++bx) {
      float gm[N * (N + 1) / 2];
      gram_of_block_n<N>(wmk, w, by, bx, gm);
      float sw = sigma0_from_gram_fast<N>(gm, nullptr);
      gram_of_block_n<N>(orig, w, by, bx, gm);
      float so = sigma0_from_gram_fast<N>(gm, nullptr);
      out[by * nbw + bx] = (uint8_t)extract_level(sw, so, alpha);
    }
}

extern "C" {

#define HOSTSIM_FOR_N(n, CALL) switch (n) { case 4: { CALL(4); } break; case 6: { CALL(6); } break; case 8: { CALL(8); } break; \
  case 10: { CALL(10); } break; case 12: { CALL(12); } break; case 14: { CALL(14); } break; case 16: { CALL(16); } break; default: return -2; }

int hostsim_embed_n(const uint8_t* rgb, uint8_t* out, int h, int w, const uint8_t* wm, double alpha, int bs) {
  for (size_t p = 0; p < (size_t)h * w; ++p) {   // strips: exact colour round trip, as in the library
    float r = unit_from_u8(rgb[3 * p]), g = unit_from_u8(rgb[3 * p + 1]), b = unit_from_u8(rgb[3 * p + 2]);
    float cb, cr; chroma_exact(r, g, b, cb, cr);
    uint32_t R, G, B; ycc_to_rgb8_exact(luma_exact(r, g, b), cb, cr, R, G, B);
    out[3 * p] = (uint8_t)R; out[3 * p + 1] = (uint8_t)G; out[3 * p + 2] = (uint8_t)B;
  }
#define CALL_E(NN) embed_n<NN>(rgb, out, h, w, wm, alpha)
  HOSTSIM_FOR_N(bs, CALL_E)
  return 0;
}

int hostsim_extract_n(const uint8_t* wmk, const uint8_t* orig, uint8_t* out, int h, int w, double alpha, int bs) {
#define CALL_X(NN) extract_n<NN>(wmk, orig, out, h, w, alpha)
  HOSTSIM_FOR_N(bs, CALL_X)
  return 0;
}

// full SVD of N row-major 8x8 blocks: S unsorted column norms, AV and V returned raw
int hostsim_svd(const float* blocks, int64_t n, float* AV, float* V, float* S, int* sweeps) {
  for (int64_t b = 0; b < n; ++b) {
    float a[64], v[64], un;
    std::memcpy(a, blocks + 64 * b, sizeof a);
    sweeps[b] = jacobi_svd8<true>(a, v, un);
    float n2[8]; column_norms2(a, n2);
    for (int j = 0; j < 8; ++j) S[8 * b + j] = f_sqrt(n2[j]) * un;
    for (int k = 0; k < 64; ++k) { AV[64 * b + k] = a[k] * un; V[64 * b + k] = v[k]; }
  }
  return 0;
}

// top singular triplet's left half of N row-major 8x8 blocks, as embed / extract get it (tmf::top_column8):
// sigma0, u0, and sweeps (+100 when the full cyclic routine was used instead of the dominant-column one)
int hostsim_top_column(const float* blocks, int64_t n, float* sigma0, float* u0, int* sweeps) {
  for (int64_t b = 0; b < n; ++b) {
    float a[64];
    std::memcpy(a, blocks + 64 * b, sizeof a);
    sigma0[b] = top_column8(a, u0 + 8 * b, sweeps + b);
  }
  return 0;
}

int hostsim_dct(const float* in, float* out, int64_t n, int inverse) {
  for (int64_t b = 0; b < n; ++b) {
    float a[64];
    std::memcpy(a, in + 64 * b, sizeof a);
    if (inverse) idct8x8(a); else dct8x8(a);
    std::memcpy(out + 64 * b, a, sizeof a);
  }
  return 0;
}

int hostsim_rgb2ycc(const uint8_t* rgb, float* ycc, int64_t npx) {
  for (int64_t p = 0; p < npx; ++p) {
    float r = unit_from_u8(rgb[3 * p]), g = unit_from_u8(rgb[3 * p + 1]), b = unit_from_u8(rgb[3 * p + 2]);
    ycc[3 * p] = luma_exact(r, g, b);
    chroma_exact(r, g, b, ycc[3 * p + 1], ycc[3 * p + 2]);
  }
  return 0;
}
}

// tmf_wm_map_l8's arithmetic on the host: the library's own table builder and per-sample
// function (tmf_resize.cuh) driven the way the two kernels drive them.
extern "C" int hostsim_wm_map_l8(const uint8_t* src, int src_h, int src_w, uint8_t* map, int target_h, int target_w,
                                 int preserve_ratio) {
  const MapGeometry g = watermark_map_geometry(src_h, src_w, target_h, target_w, preserve_ratio);
  if (g.new_h <= 0 || g.new_w <= 0) return -1;
  const bool need_h = g.new_w != src_w, need_v = g.new_h != src_h;
  AxisTable th, tv;
  int row0 = 0, rows = src_h;
  if (need_v) {
    lanczos_axis_table(src_h, g.new_h, tv);
    if (need_h) {
      row0 = tv.bounds[0];
      rows = tv.bounds[2 * (g.new_h - 1)] + tv.bounds[2 * (g.new_h - 1) + 1] - row0;
      for (int i = 0; i < g.new_h; ++i) tv.bounds[2 * i] -= row0;
    }
  }
  std::vector<uint8_t> tmp;
  const uint8_t* in = src;
  int pitch = src_w;
  if (need_h) {
    lanczos_axis_table(src_w, g.new_w, th);
    tmp.resize((size_t)rows * g.new_w);
    for (int r = 0; r < rows; ++r)
      for (int xx = 0; xx < g.new_w; ++xx)
        tmp[(size_t)r * g.new_w + xx] = resample_sample(src + (size_t)(row0 + r) * src_w + th.bounds[2 * xx], 1,
                                                        &th.kk[(size_t)xx * th.ksize], 1, th.bounds[2 * xx + 1]);
    in = tmp.data();
    pitch = g.new_w;
  }
  for (int y = 0; y < target_h; ++y)
    for (int x = 0; x < target_w; ++x) {
      const int yy = y - g.paste_y, xx = x - g.paste_x;
      uint8_t v = 255;
      if (yy >= 0 && yy < g.new_h && xx >= 0 && xx < g.new_w)
        v = need_v ? resample_sample(in + (size_t)tv.bounds[2 * yy] * pitch + xx, pitch, &tv.kk[(size_t)yy * tv.ksize], 1,
                                     tv.bounds[2 * yy + 1])
                   : in[(size_t)yy * pitch + xx];
      map[(size_t)y * target_w + x] = v;
    }
  return 0;
}
