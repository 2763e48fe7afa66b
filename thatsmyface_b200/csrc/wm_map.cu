// Watermark map on the device (tmf_resize.cuh): resize_watermark after .convert("L")
// (modules/watermarking.py:86-132) - PIL's LANCZOS resize restated, bit-exact.
// Workspace layout: [bounds_h][kT_h][bounds_v][k_v] int32 tables, then the uint8 image
// between the two passes (n x rows x new_w), every part 16-byte aligned.
#include <math.h>
#include <string.h>
#include <vector>

#include "tmf_common.cuh"
#include "tmf_resize.cuh"

#include <memory>
#include <mutex>

using namespace tmfi;

namespace {
// host-side weight tables of one (source, resized) size pair, kept for the next call
struct TableKey {
  int src_h, src_w, new_h, new_w;
  bool operator==(const TableKey& o) const { return src_h == o.src_h && src_w == o.src_w && new_h == o.new_h && new_w == o.new_w; }
};
struct HostTables {
  std::vector<uint8_t> bytes;   // [bounds_h][kT_h][bounds_v][k_v], laid out as in the device workspace
  int row0 = 0, rows = 0;       // source rows the horizontal pass covers
};
std::mutex g_tables_mutex;
std::vector<std::pair<TableKey, std::shared_ptr<const HostTables>>> g_tables;   // most recent last, at most 8

std::shared_ptr<const HostTables> cached_tables(const TableKey& k) {
  std::lock_guard<std::mutex> lock(g_tables_mutex);
  for (auto& e : g_tables)
    if (e.first == k) return e.second;
  return nullptr;
}
std::shared_ptr<const HostTables> remember_tables(const TableKey& k, std::shared_ptr<const HostTables> t) {
  std::lock_guard<std::mutex> lock(g_tables_mutex);
  for (auto& e : g_tables)
    if (e.first == k) return e.second;          // another thread was faster: same content
  if (g_tables.size() >= 8) g_tables.erase(g_tables.begin());
  g_tables.emplace_back(k, t);
  return t;
}
}  // namespace

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

// Host only (no device is touched): the weight table of one axis, as the kernels get it.  A GPU-free pin of this
// library's host floating point (libm sin, -ffp-contract=off) against Pillow's tables.
int tmf_wm_map_axis_table(int in_size, int out_size, int* ksize, int32_t* bounds, int32_t* kk, size_t kk_capacity) {
  if (in_size < 1 || out_size < 1 || !ksize) return fail(TMF_ERR_BAD_ARG, "axis table: sizes must be positive");
  tmf::AxisTable t;
  tmf::lanczos_axis_table(in_size, out_size, t);
  *ksize = t.ksize;
  if (!bounds && !kk) return TMF_OK;                      // size query
  if (!bounds || !kk || kk_capacity < t.kk.size()) return fail(TMF_ERR_BAD_ARG, "axis table: %zu coefficients needed", t.kk.size());
  memcpy(bounds, t.bounds.data(), t.bounds.size() * sizeof(int32_t));
  memcpy(kk, t.kk.data(), t.kk.size() * sizeof(int32_t));
  return TMF_OK;
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

  // weight tables on the host (libm sin, as Pillow), one upload.  The tables depend on the four sizes
  // only and cost ~0.2 ms of sin() per call (most of a single map's latency), so the last few are kept.
  int row0 = 0, rows = src_h;
  const TableKey key{src_h, src_w, p.g.new_h, p.g.new_w};
  std::shared_ptr<const HostTables> cached = cached_tables(key);
  if (!cached) {
    auto t = std::make_shared<HostTables>();
    tmf::AxisTable th, tv;
    t->bytes.assign(p.table_bytes, 0);
    t->rows = src_h;
    if (p.need_v) {
      tmf::lanczos_axis_table(src_h, p.g.new_h, tv);
      if (p.need_h) {  // ImagingResampleInner: the horizontal pass covers only the rows the vertical one reads
        t->row0 = tv.bounds[0];
        t->rows = tv.bounds[2 * (p.g.new_h - 1)] + tv.bounds[2 * (p.g.new_h - 1) + 1] - t->row0;
        for (int i = 0; i < p.g.new_h; ++i) tv.bounds[2 * i] -= t->row0;
      }
      memcpy(t->bytes.data() + p.off_bv, tv.bounds.data(), tv.bounds.size() * 4);
      memcpy(t->bytes.data() + p.off_kv, tv.kk.data(), tv.kk.size() * 4);
    }
    if (p.need_h) {
      tmf::lanczos_axis_table(src_w, p.g.new_w, th);
      memcpy(t->bytes.data() + p.off_bh, th.bounds.data(), th.bounds.size() * 4);
      int32_t* kT = reinterpret_cast<int32_t*>(t->bytes.data() + p.off_kh);  // tap-major for coalesced loads
      for (int xx = 0; xx < p.g.new_w; ++xx)
        for (int i = 0; i < th.ksize; ++i) kT[(size_t)i * p.g.new_w + xx] = th.kk[(size_t)xx * th.ksize + i];
    }
    cached = remember_tables(key, t);
  }
  row0 = cached->row0;
  rows = cached->rows;
  const std::vector<uint8_t>& host = cached->bytes;
  if (p.table_bytes > 0) {
    // pageable source: the runtime stages it before returning, so `host` may go out of scope
    cudaError_t e = cudaMemcpyAsync(ws, host.data(), p.table_bytes, cudaMemcpyHostToDevice, st);   // (`cached` keeps `host` alive)
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

}  // extern "C"
