// libtmfwm C ABI of the fused entry points (include/tmf_wm.h): argument checks and dispatch
// to the kernel families; plus the helpers every translation unit shares (tmf_common.cuh).
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>

#include "tmf_common.cuh"

namespace tmfi {

namespace {
thread_local char g_err[512] = "";
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return code;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(TMF_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return TMF_OK;
}

int cuda_fail(const char* what, cudaError_t e) {
  cudaGetLastError();
  return fail(TMF_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

const char* last_error_message() { return g_err; }

int make_geom(int n, int h, int w, size_t img_stride, int block, BlockGeom& g) {
  if (block < 4 || block > 16 || (block & 1))
    return fail(TMF_ERR_UNSUPPORTED_BLOCK,
                "block_size %d is not supported: this build implements the even sizes 4..16 the reference's UI "
                "offers (8, its BLOCK_SIZE, on the optimised path); there is no CPU fallback", block);
  if (n < 0 || h < 0 || w < 0) return fail(TMF_ERR_BAD_ARG, "negative dimension (n=%d h=%d w=%d)", n, h, w);
  if (n > 0 && img_stride < (size_t)h * w * 3)
    return fail(TMF_ERR_BAD_ARG, "img_stride %zu is smaller than one image (%zu bytes)", img_stride, (size_t)h * w * 3);
  g.h = h; g.w = w; g.bs = block; g.nbh = h / block; g.nbw = w / block;
  g.blocks_per_img = (long long)g.nbh * g.nbw;
  g.total_blocks = g.blocks_per_img * n;
  if (g.total_blocks > 0x7fffffffLL)   // 2^31 blocks of 192 B would be 412 GB of pixels
    return fail(TMF_ERR_BAD_ARG, "batch too large: %lld blocks (limit 2^31 - 1 per call)", g.total_blocks);
  g.div_bpi = make_fastdiv((uint32_t)(g.blocks_per_img > 0 ? g.blocks_per_img : 1));
  g.div_nbw = make_fastdiv((uint32_t)(g.nbw > 0 ? g.nbw : 1));
  g.img_stride = img_stride;
  g.row_pitch = (size_t)w * 3;
  if (g.row_pitch > 0xffffffffull)
    return fail(TMF_ERR_BAD_ARG, "image rows of %zu bytes are not supported (limit 2^32 - 1)", g.row_pitch);
  g.pitch32 = (uint32_t)g.row_pitch;
  return TMF_OK;
}

int pick_vec(const BlockGeom& g, const void* p0, const void* p1, const void* p2) {
  uintptr_t bits = (uintptr_t)p0 | (uintptr_t)p1 | (uintptr_t)p2 | (uintptr_t)g.img_stride | (uintptr_t)g.row_pitch;
  if ((bits & 7) == 0) return 8;
  if ((bits & 3) == 0) return 4;
  return 1;
}

int row_alignment(const BlockGeom& g, const void* p0, const void* p1, const void* p2) {
  uintptr_t bits = (uintptr_t)p0 | (uintptr_t)p1 | (uintptr_t)p2 | (uintptr_t)g.img_stride | (uintptr_t)g.row_pitch;
  return (bits & 3) == 0 ? 4 : ((bits & 1) == 0 ? 2 : 1);
}

int sm_count() {
  static std::atomic<int> cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 148; }
  if (dev >= 0 && dev < 64) {
    const int c = cache[dev].load(std::memory_order_relaxed);
    if (c > 0) return c;
  }
  int sms = 148;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) { cudaGetLastError(); sms = 148; }
  if (dev >= 0 && dev < 64) cache[dev].store(sms, std::memory_order_relaxed);
  return sms;
}

}  // namespace tmfi

using namespace tmfi;

namespace {
int check_mode(int mode) {
  if (mode != TMF_MODE_FAITHFUL && mode != TMF_MODE_FAST && mode != TMF_MODE_LITERAL)
    return fail(TMF_ERR_BAD_ARG, "unknown mode %d", mode);
  return TMF_OK;
}
}  // namespace

extern "C" {

int tmf_version(void) { return TMF_VERSION; }
const char* tmf_last_error(void) { return last_error_message(); }

int tmf_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) return cuda_fail("cudaGetDeviceCount", e);
  return n;
}

int tmf_last_fast_path(void) { return last_fast_path(); }

int tmf_embed_rgb8(const uint8_t* rgb, uint8_t* out, int n, int h, int w, size_t img_stride, const uint8_t* wm,
                   int wm_shared, double alpha, int block, int mode, void* stream) {
  BlockGeom g;
  if (int rc = make_geom(n, h, w, img_stride, block, g)) return rc;
  if (int rc = check_mode(mode)) return rc;
  if (n == 0 || h == 0 || w == 0) return TMF_OK;
  if (!rgb || !out) return fail(TMF_ERR_BAD_ARG, "null image pointer");
  if (rgb == out) return fail(TMF_ERR_BAD_ARG, "out must not alias rgb");
  if (g.total_blocks > 0 && !wm) return fail(TMF_ERR_BAD_ARG, "null watermark map");
  if (!(alpha == alpha)) return fail(TMF_ERR_BAD_ARG, "alpha is NaN");
  cudaStream_t st = (cudaStream_t)stream;
  if (g.total_blocks > 0) {
    int rc;
    if (mode != TMF_MODE_FAST) rc = launch_embed_faithful(rgb, out, g, wm, wm_shared, alpha, mode == TMF_MODE_LITERAL, st);
    else if (block == 8) rc = launch_embed_fast8(rgb, out, g, wm, wm_shared, alpha, st);
    else rc = launch_embed_fast_n(rgb, out, g, wm, wm_shared, alpha, st);
    if (rc) return rc;
  }
  return launch_strip_roundtrip(rgb, out, g, n, st);
}

int tmf_extract_rgb8(const uint8_t* wmk_rgb, const uint8_t* orig_rgb, uint8_t* out_wm, int n, int h, int w,
                     size_t img_stride, double alpha, int block, int mode, void* stream) {
  BlockGeom g;
  if (int rc = make_geom(n, h, w, img_stride, block, g)) return rc;
  if (int rc = check_mode(mode)) return rc;
  if (g.total_blocks == 0) return TMF_OK;
  if (!wmk_rgb || !orig_rgb || !out_wm) return fail(TMF_ERR_BAD_ARG, "null pointer");
  if (!(alpha == alpha) || alpha == 0.0) return fail(TMF_ERR_BAD_ARG, "alpha must be a non-zero number");
  cudaStream_t st = (cudaStream_t)stream;
  if (mode != TMF_MODE_FAST) return launch_extract_faithful(wmk_rgb, orig_rgb, out_wm, g, alpha, st);
  if (block == 8) return launch_extract_fast8(wmk_rgb, orig_rgb, out_wm, g, alpha, st);
  return launch_extract_fast_n(wmk_rgb, orig_rgb, out_wm, g, alpha, st);
}

int tmf_sigma0_rgb8(const uint8_t* rgb, float* sigma0, int n, int h, int w, size_t img_stride, int block, int mode,
                    void* stream) {
  BlockGeom g;
  if (int rc = make_geom(n, h, w, img_stride, block, g)) return rc;
  if (int rc = check_mode(mode)) return rc;
  if (g.total_blocks == 0) return TMF_OK;
  if (!rgb || !sigma0) return fail(TMF_ERR_BAD_ARG, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (mode != TMF_MODE_FAST) return launch_sigma0_faithful(rgb, sigma0, g, st);
  if (block == 8) return launch_sigma0_fast8(rgb, sigma0, g, st);
  return launch_sigma0_fast_n(rgb, sigma0, g, st);
}

}  // extern "C"
