// Watermark-map preparation on the device: resize_watermark (modules/watermarking.py:86-132)
// after `.convert("L")` - PIL's LANCZOS resize of a mode-"L" image plus the centred paste on a
// white canvas - bit-exact with Pillow's 8-bit resampler.
//
// Pillow is a third-party dependency of the reference (requirements.txt:3, unpinned; 12.2.0 in
// this image), so this restates its published algorithm (src/libImaging/Resample.c:
// precompute_coeffs, normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc,
// ImagingResampleVertical_8bpc): float64 Lanczos-3 weights per output sample, normalised, rounded
// to 22-bit fixed point; a horizontal pass then a vertical pass, each
// clip8((2^21 + sum pixel*k) >> 22) in int32 with a uint8 image between them.
//
// The weight tables are built on the HOST with the C library's sin() - the one Pillow's own
// extension calls - so the fixed-point tables are identical; everything after that is integer
// arithmetic and therefore exact on the device.  The per-sample function is __host__ __device__
// so tests/hostsim can run it without a GPU (test harness only).
#pragma once
#include <math.h>
#include <stdint.h>
#include <vector>

#if defined(__CUDACC__)
#define TMF_RS_HD __host__ __device__ __forceinline__
#else
#define TMF_RS_HD inline
#endif

namespace tmf {

constexpr int kResampleBits = 32 - 8 - 2;  // Resample.c PRECISION_BITS

// ---- host: weight tables --------------------------------------------------------------------
inline double sinc_pi(double x) {
  if (x == 0.0) return 1.0;
  x = x * 3.14159265358979323846;  // M_PI
  return sin(x) / x;
}
inline double lanczos3(double x) { return (-3.0 <= x && x < 3.0) ? sinc_pi(x) * sinc_pi(x / 3) : 0.0; }

struct AxisTable {
  int in_size = 0, out_size = 0, ksize = 0;
  std::vector<int32_t> bounds;  // out_size x {first, count}
  std::vector<int32_t> kk;      // out_size x ksize (zero padded)
};

// precompute_coeffs + normalize_coeffs_8bpc for the whole-image box (0, in_size).
inline void lanczos_axis_table(int in_size, int out_size, AxisTable& t) {
  const double scale = (double)(float)in_size / out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 3.0 * filterscale;
  const int ksize = (int)ceil(support) * 2 + 1;
  const double ss = 1.0 / filterscale;
  t.in_size = in_size; t.out_size = out_size; t.ksize = ksize;
  t.bounds.assign((size_t)out_size * 2, 0);
  t.kk.assign((size_t)out_size * ksize, 0);
  std::vector<double> w((size_t)ksize);
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = 0.0 + (xx + 0.5) * scale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      w[x] = lanczos3((x + xmin - center + 0.5) * ss);
      ww += w[x];
    }
    int32_t* k = &t.kk[(size_t)xx * ksize];
    for (int x = 0; x < xmax; ++x) {
      const double v = ww != 0.0 ? w[x] / ww : w[x];
      k[x] = v < 0 ? (int32_t)(-0.5 + v * (1 << kResampleBits)) : (int32_t)(0.5 + v * (1 << kResampleBits));
    }
    t.bounds[2 * xx] = xmin;
    t.bounds[2 * xx + 1] = xmax;
  }
}

// Sizes and paste offset of resize_watermark (:105-123).  Python's `tw / ow` is a float64
// division and int() truncates, as here.
struct MapGeometry {
  int new_h, new_w, paste_y, paste_x;
};
inline int floordiv2(int a) { return a >= 0 ? a / 2 : -((-a + 1) / 2); }
inline MapGeometry watermark_map_geometry(int src_h, int src_w, int target_h, int target_w, int preserve_ratio) {
  MapGeometry g;
  if (preserve_ratio) {
    const double rw = (double)target_w / src_w, rh = (double)target_h / src_h;
    const double ratio = rw < rh ? rw : rh;
    g.new_w = (int)(src_w * ratio);
    g.new_h = (int)(src_h * ratio);
    g.paste_x = floordiv2(target_w - g.new_w);
    g.paste_y = floordiv2(target_h - g.new_h);
  } else {
    g.new_h = target_h; g.new_w = target_w; g.paste_y = 0; g.paste_x = 0;
  }
  return g;
}

// ---- host + device: one output sample -------------------------------------------------------
TMF_RS_HD uint8_t resample_clip8(int32_t acc) {
  const int32_t v = acc >> kResampleBits;  // arithmetic shift, as clip8() in Resample.c
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// sum over `count` taps of src[i*stride] * k[i*kstride], rounded and clipped
TMF_RS_HD uint8_t resample_sample(const uint8_t* src, long long stride, const int32_t* k, long long kstride,
                                  int count) {
  int32_t acc = 1 << (kResampleBits - 1);
  for (int i = 0; i < count; ++i) acc += (int32_t)src[i * stride] * k[i * kstride];
  return resample_clip8(acc);
}

#if defined(__CUDACC__)
// ---- device: the two passes -----------------------------------------------------------------
constexpr int kResizeThreads = 128;
constexpr int kResizeSmemBytes = 48 * 1024;  // static limit, no opt-in needed

// Horizontal pass.  One CTA stages R consecutive source rows (a contiguous byte range, the
// images are tightly packed) in shared memory with 16-byte loads and produces their out_w
// samples; a thread owns output column(s) xx and keeps R accumulators, so every weight is
// loaded once per R multiply-adds.  kT is the weight table transposed (tap-major) so a warp's
// weight loads are contiguous.  tmp: n x rows x out_w.
template <int R>
__global__ void __launch_bounds__(kResizeThreads)
k_resample_rows(const uint8_t* __restrict__ src, size_t src_stride, int src_w, int row0, int rows,
                uint8_t* __restrict__ tmp, int out_w, const int32_t* __restrict__ bounds,
                const int32_t* __restrict__ kT) {
  extern __shared__ __align__(16) uint8_t s_raw[];
  const int img = blockIdx.y;
  const int r_first = blockIdx.x * R;
  const int r_cnt = min(R, rows - r_first);
  const uint8_t* base = src + (size_t)img * src_stride + (size_t)(row0 + r_first) * src_w;
  const int total = r_cnt * src_w;
  const int mis = (int)((uintptr_t)base & 15);
  uint8_t* s = s_raw + mis;  // s[i] <-> base[i]; s + head is 16-byte aligned
  const int head = min(total, (16 - mis) & 15);
  const int nvec = (total - head) >> 4;
  for (int i = threadIdx.x; i < head; i += kResizeThreads) s[i] = __ldg(base + i);
  for (int v = threadIdx.x; v < nvec; v += kResizeThreads)
    *reinterpret_cast<uint4*>(s + head + 16 * v) = __ldg(reinterpret_cast<const uint4*>(base + head) + v);
  for (int i = head + 16 * nvec + threadIdx.x; i < total; i += kResizeThreads) s[i] = __ldg(base + i);
  __syncthreads();

  for (int xx = threadIdx.x; xx < out_w; xx += kResizeThreads) {
    const int first = __ldg(bounds + 2 * xx), cnt = __ldg(bounds + 2 * xx + 1);
    int32_t acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 1 << (kResampleBits - 1);
    const uint8_t* p = s + first;
    for (int i = 0; i < cnt; ++i) {
      const int32_t k = __ldg(kT + (size_t)i * out_w + xx);
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] += (int32_t)p[r * src_w + i] * k;  // rows >= r_cnt: unused
    }
    uint8_t* o = tmp + ((size_t)img * rows + r_first) * out_w + xx;
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (r < r_cnt) o[(size_t)r * out_w] = resample_clip8(acc[r]);
  }
}

// Vertical pass + paste: one thread per pixel of the target_h x target_w canvas.  Inside the
// pasted rectangle it resamples column xx of `in` (the horizontal pass's output, or the source
// itself when the widths already match); outside it writes the white border (:116).
__global__ void __launch_bounds__(256)
k_compose_map(const uint8_t* __restrict__ in, size_t in_img_stride, int in_pitch, uint8_t* __restrict__ maps,
              int target_h, int target_w, MapGeometry g, const int32_t* __restrict__ bounds_v,
              const int32_t* __restrict__ kv, int ksize_v, int need_v, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int x = (int)(idx % target_w);
  const long long t = idx / target_w;
  const int y = (int)(t % target_h);
  const long long img = t / target_h;
  const int yy = y - g.paste_y, xx = x - g.paste_x;
  uint8_t v = 255;
  if (yy >= 0 && yy < g.new_h && xx >= 0 && xx < g.new_w) {
    const uint8_t* col = in + (size_t)img * in_img_stride + xx;
    if (need_v) {
      const int first = __ldg(bounds_v + 2 * yy), cnt = __ldg(bounds_v + 2 * yy + 1);
      v = resample_sample(col + (size_t)first * in_pitch, in_pitch, kv + (size_t)yy * ksize_v, 1, cnt);
    } else {
      v = __ldg(col + (size_t)yy * in_pitch);
    }
  }
  maps[idx] = v;
}
#endif  // __CUDACC__

}  // namespace tmf
