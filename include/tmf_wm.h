/* libtmfwm - B200 (sm_100a) DCT+SVD watermark embed / extract.  C ABI.
 *
 * Drop-in boundary for the hot path of Rigelyon/ThatsMyFace
 * `modules/watermarking.py`.  The reference is pure Python, so there is no FFI
 * in it to mirror; these entry points are what a ctypes binding placed inside
 * the reference's two public functions calls:
 *
 *   embed_watermark()   modules/watermarking.py:135-221  -> tmf_embed_rgb8
 *   extract_watermark() modules/watermarking.py:224-294  -> tmf_extract_rgb8
 *   rgb_to_ycbcr()      :23-50                           -> tmf_rgb8_to_ycbcr_f32
 *   ycbcr_to_rgb()      :53-73                           -> tmf_ycbcr_f32_to_rgb8
 *   apply_dct_to_block / apply_idct_to_block :76-83      -> tmf_dct8x8_f32
 *   np.linalg.svd(block, full_matrices=True) :195,:279   -> tmf_svd8x8_f32
 *   resize_watermark() after .convert("L")  :105-132     -> tmf_wm_map_l8
 *
 * Conventions
 *   - every pointer is a DEVICE pointer on the current CUDA device unless the
 *     function name ends in `_host`;
 *   - images are interleaved 8-bit RGB, rows tightly packed (3*w bytes),
 *     image k of a batch starts at `base + k*img_stride` bytes;
 *   - the caller owns every buffer; the library allocates nothing persistent
 *     and keeps no reference after the call's work on `stream` has completed;
 *   - `stream` is a cudaStream_t (NULL = default stream); calls are
 *     asynchronous with respect to the host and re-entrant;
 *   - return value 0 = OK, negative = error (enum below); a description of the
 *     last error on the calling thread is available from tmf_last_error();
 *   - there is no CPU fallback: without a usable CUDA device every compute
 *     entry point returns TMF_ERR_CUDA.
 */
#ifndef TMF_WM_H
#define TMF_WM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TMF_VERSION 201 /* 0.2.1: + tmf_wm_map_axis_table (0.2.0: TMF_MODE_LITERAL, tmf_last_fast_path, faithful mode at every block size) */

enum {
  TMF_OK = 0,
  TMF_ERR_BAD_ARG = -1,           /* null pointer, negative size, bad stride, bad mode */
  TMF_ERR_UNSUPPORTED_BLOCK = -2, /* block_size not one of 4, 6, 8, 10, 12, 14, 16 (the UI's range,
                                     embed_watermark_page.py:324-331); 8 = BLOCK_SIZE, constants.py:7 */
  TMF_ERR_CUDA = -3               /* CUDA runtime / launch failure (message has the CUDA error string) */
};

/* `mode` of the fused kernels */
enum {
  TMF_MODE_FAITHFUL = 0, /* DCT -> one-sided Jacobi for sigma0, u0 (rotations against the dominant column only
                            when one is certified dominant, full cyclic sweeps otherwise) -> sigma0 += alpha*w ->
                            U S' V^T (as the rank-1 identity below) -> IDCT; colour math bit-exact with the
                            reference's float64 dot */
  TMF_MODE_FAST = 1,     /* algebraically reduced: top singular triplet of the spatial block and a
                            rank-1 update (orthonormal DCT preserves singular values; NO DCT and NO SVD
                            are executed); fp32 colour */
  TMF_MODE_LITERAL = 2   /* block size 8 only: FAITHFUL with the literal U diag(S') V^T product and IDCT
                            (V accumulated by the Jacobi).  FAITHFUL itself uses the identity
                            U diag(S') V^T = D + (S'[0] - S[0]) u0 v0^T, v0 = D^T u0 / S[0] - the same
                            matrix, without V.  For extract / sigma0 LITERAL equals FAITHFUL. */
};

int tmf_version(void);
const char* tmf_last_error(void);

/* Number of CUDA devices visible, or a negative error. */
int tmf_device_count(void);

/* Which kernel the calling thread's last FAST block-8 embed / extract took:
 * 1 = TMA-tiled persistent kernel, 0 = per-thread kernel. */
int tmf_last_fast_path(void);

/* `block` is the reference's block_size (any even size 4..16, the range of its UI); both modes
 * honour every size (8 = BLOCK_SIZE runs the tuned kernels).  Below, B = block.
 * Limits per call: n * (h/B) * (w/B) < 2^31 blocks, 3*w < 2^32 bytes per row (TMF_ERR_BAD_ARG).
 *
 * embed_watermark on a batch.  rgb/out: n images of h x w x 3 bytes (out may not
 * alias rgb).  wm: watermark map(s), (h/B) x (w/B) bytes each, already resized
 * (resize_watermark stays on the host, watermarking.py:86-132); one map per
 * image, or a single shared map when wm_shared != 0.  Pixels outside whole
 * blocks (h%B, w%B strips) take the colour round trip only, as in the
 * reference.  alpha is double because the reference adds alpha*w in float64
 * (watermarking.py:198).
 * FAST mode, block 8: batches whose pointers, img_stride and 3*w are multiples of 16 and whose rows
 * hold a multiple of 16 blocks (w % 128 == 0: 512, 1280, 1920, 3840, 7680 ...) run the TMA-tiled
 * persistent kernel; everything else the per-thread kernel - same results (tmf_last_fast_path()).
 * TMF_NO_TILE=1 in the environment (read once) forces the per-thread kernel, for A/B tests. */
int tmf_embed_rgb8(const uint8_t* rgb, uint8_t* out, int n, int h, int w, size_t img_stride,
                   const uint8_t* wm, int wm_shared, double alpha, int block, int mode, void* stream);

/* extract_watermark on a batch.  out_wm: n maps of (h/B) x (w/B) bytes. */
int tmf_extract_rgb8(const uint8_t* wmk_rgb, const uint8_t* orig_rgb, uint8_t* out_wm, int n, int h, int w,
                     size_t img_stride, double alpha, int block, int mode, void* stream);

/* Tap: largest singular value of every whole luma block, float32, n x (h/B) x (w/B). */
int tmf_sigma0_rgb8(const uint8_t* rgb, float* sigma0, int n, int h, int w, size_t img_stride, int block,
                    int mode, void* stream);

/* Batched SVD of nblocks row-major 8x8 float32 matrices (one-sided Jacobi).
 * S: nblocks x 8, descending.  U, Vt: nblocks x 64 row-major, or both NULL for
 * values only; A = U diag(S) Vt.  Columns of U whose singular value is below
 * 1e-6*S[0] are completed to an orthonormal basis only if `complete_u` != 0
 * (LAPACK's full_matrices=True contract); otherwise they are returned as zero
 * columns (the reference only ever multiplies them by that singular value).
 * sweeps (nullable): nblocks int32, Jacobi sweeps that rotated. */
int tmf_svd8x8_f32(const float* blocks, int64_t nblocks, float* S, float* U, float* Vt, int32_t* sweeps,
                   int complete_u, void* stream);

/* Batched orthonormal 2-D DCT-II (inverse = 0) or its inverse (inverse = 1) of
 * row-major 8x8 float32 blocks; in may equal out. */
int tmf_dct8x8_f32(const float* in, float* out, int64_t nblocks, int inverse, void* stream);

/* rgb_to_ycbcr / ycbcr_to_rgb taps on npixels interleaved pixels, bit-exact
 * with the reference (float64 dot, float32 storage, truncating quantiser). */
int tmf_rgb8_to_ycbcr_f32(const uint8_t* rgb, float* ycc, int64_t npixels, void* stream);
int tmf_ycbcr_f32_to_rgb8(const float* ycc, uint8_t* rgb, int64_t npixels, void* stream);

/* Pixel-format taps for the PIL boundary (image.convert("RGB"), watermarking.py:154,242-243, and
 * Image.fromarray, :219): PIL stores an "RGB" image as 4 bytes per pixel (R, G, B, pad), and packing /
 * unpacking that on the host costs ~25 ms per 4K image each way.  The host binding moves PIL's own
 * layout and converts here.  rgbx: 16-byte aligned, rgb: 4-byte aligned; pad = the 4th byte written. */
int tmf_rgbx8_to_rgb8(const uint8_t* rgbx, uint8_t* rgb, int64_t npixels, void* stream);
int tmf_rgb8_to_rgbx8(const uint8_t* rgb, uint8_t* rgbx, int64_t npixels, int pad, void* stream);

/* ---- watermark map on the device -------------------------------------------------------
 * resize_watermark (watermarking.py:86-132) after `.convert("L")`: PIL's LANCZOS resize of n
 * mode-"L" images (src_h x src_w bytes each, tightly packed rows, image k at src + k*src_stride)
 * to the block grid, target_h x target_w = (h/B) x (w/B).  preserve_ratio != 0 keeps the aspect
 * ratio (sizes truncated as :107-110) and pastes the result centred on a white canvas
 * (:116-123); 0 stretches to the target (:128-130).  maps: n x target_h x target_w bytes.
 * Bit-exact with Pillow's 8-bit resampler (Resample.c: float64 Lanczos-3 weights rounded to
 * 22-bit fixed point, horizontal pass then vertical pass with a uint8 image in between); the
 * weight tables are built on the host with libm's sin(), the passes run on the device.
 * `workspace`: device scratch of at least tmf_wm_map_workspace_bytes(...) bytes, 16-byte
 * aligned, caller-owned, reusable once the work on `stream` has completed.
 * Not supported (TMF_ERR_BAD_ARG): a resized side of 0 pixels (PIL raises too), sources more
 * than 100x taller than wide (PIL switches pass order there), src_w above 49136. */
size_t tmf_wm_map_workspace_bytes(int n, int src_h, int src_w, int target_h, int target_w, int preserve_ratio);
int tmf_wm_map_l8(const uint8_t* src, int n, int src_h, int src_w, size_t src_stride, uint8_t* maps, int target_h,
                  int target_w, int preserve_ratio, void* workspace, size_t workspace_bytes, void* stream);

/* Host only, no device needed: the weight table of one resampling axis as the kernels receive it
 * (Resample.c precompute_coeffs + normalize_coeffs_8bpc): ksize, bounds = out_size x {first, count},
 * kk = out_size x ksize 22-bit fixed-point weights.  bounds == kk == NULL queries ksize only. */
int tmf_wm_map_axis_table(int in_size, int out_size, int* ksize, int32_t* bounds, int32_t* kk, size_t kk_capacity);

/* ---- host-buffer pipeline -------------------------------------------------------------
 * The per-image loop of the embed page (embed_watermark_page.py:492-558) as one call on
 * HOST memory.  A context owns 3 streams and `depth` device slots on one device; the
 * batch is cut into chunks of ~chunk_bytes (0 = 96 MiB) and H2D copy, fused kernel and D2H
 * copy of successive chunks overlap.  Calls enqueue and return (no host-side wait, except when a
 * device buffer has to grow; if an enqueue fails the context is drained before the error is
 * returned, so the host buffers are no longer in use); results are in the host
 * buffers after tmf_ctx_synchronize().  For several GPUs: one context per device, split the
 * batch by image, enqueue on all, synchronise each (no collective).  A context is not
 * thread-safe; different contexts are independent.  Host buffers must stay valid until the
 * synchronise; page-lock them (tmf_pin_host, or any pinned allocator) for real overlap. */
#define TMF_CTX_MAX_DEPTH 8
typedef struct tmf_ctx tmf_ctx;

int tmf_ctx_create(tmf_ctx** ctx, int device, size_t chunk_bytes, int depth);
int tmf_ctx_destroy(tmf_ctx* ctx);

/* images: n x h x w x 3 bytes, tightly packed; wm: (h/B) x (w/B) bytes, one shared map
 * (wm_shared != 0) or n maps. */
int tmf_ctx_embed_host_async(tmf_ctx* ctx, const uint8_t* rgb, uint8_t* out, int n, int h, int w,
                             const uint8_t* wm, int wm_shared, double alpha, int block, int mode);
/* out_wm: n x (h/B) x (w/B) bytes. */
int tmf_ctx_extract_host_async(tmf_ctx* ctx, const uint8_t* wmk_rgb, const uint8_t* orig_rgb, uint8_t* out_wm,
                               int n, int h, int w, double alpha, int block, int mode);
int tmf_ctx_synchronize(tmf_ctx* ctx);
/* kernel launches and bytes copied since creation / the last reset (any pointer may be NULL) */
int tmf_ctx_stats(tmf_ctx* ctx, long long* launches, long long* h2d_bytes, long long* d2h_bytes, int reset);

/* cudaHostRegister / cudaHostUnregister for callers without a CUDA toolchain */
int tmf_pin_host(void* p, size_t bytes);
int tmf_unpin_host(void* p);

#ifdef __cplusplus
}
#endif
#endif /* TMF_WM_H */
