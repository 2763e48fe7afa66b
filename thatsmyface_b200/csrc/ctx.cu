// Host-buffer pipeline of libtmfwm (C ABI, tmf_ctx_*).
#include <string.h>

#include "tmf_common.cuh"

using namespace tmfi;

extern "C" {

// ---------------------------------------------------------------------------
// Host-buffer pipeline (C ABI).  A context owns three streams and `depth` device
// slots on one device; a batch of HOST images is cut into chunks and each chunk
// flows  H2D copy -> fused kernel -> D2H copy  with the three stages of different
// chunks overlapping.  Sharding over GPUs is by image: one context per device,
// enqueue on all of them, then synchronise each - no collective, no peer traffic.
// Host buffers should be page-locked (tmf_pin_host) for the copies to be truly
// asynchronous; pageable buffers work, staged by the driver.
// ---------------------------------------------------------------------------
struct tmf_ctx {
  int device;
  int depth;
  size_t chunk_bytes;
  cudaStream_t s_in, s_run, s_out;
  struct Slot {
    uint8_t *a, *b, *o, *wm;
    size_t cap_a, cap_b, cap_o, cap_wm;
    cudaEvent_t ev_in, ev_run, ev_out;
    bool used;
  } slot[TMF_CTX_MAX_DEPTH];
  uint8_t* wm_shared;
  size_t cap_wm_shared;
  cudaEvent_t ev_wm;         // recorded on s_run after the last kernel that reads wm_shared
  bool wm_used;
  long long launches, h2d_bytes, d2h_bytes;
};

namespace {

struct DeviceGuard {
  int prev;
  bool ok;
  explicit DeviceGuard(int dev) : prev(-1), ok(false) {
    if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); return; }
    ok = (cudaSetDevice(dev) == cudaSuccess);
    if (!ok) cudaGetLastError();
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int grow(uint8_t** p, size_t* cap, size_t need) {
  if (*cap >= need) return TMF_OK;
  if (*p) cudaFree(*p);
  *p = nullptr; *cap = 0;
  cudaError_t e = cudaMalloc((void**)p, need);
  if (e != cudaSuccess) return cuda_fail("cudaMalloc", e);
  *cap = need;
  return TMF_OK;
}

// streams / events / buffers of a context (also of a half-built one: null handles are skipped)
void release(tmf_ctx* c) {
  for (int k = 0; k < TMF_CTX_MAX_DEPTH; ++k) {
    tmf_ctx::Slot& sl = c->slot[k];
    if (sl.a) cudaFree(sl.a);
    if (sl.b) cudaFree(sl.b);
    if (sl.o) cudaFree(sl.o);
    if (sl.wm) cudaFree(sl.wm);
    if (sl.ev_in) cudaEventDestroy(sl.ev_in);
    if (sl.ev_run) cudaEventDestroy(sl.ev_run);
    if (sl.ev_out) cudaEventDestroy(sl.ev_out);
  }
  if (c->wm_shared) cudaFree(c->wm_shared);
  if (c->ev_wm) cudaEventDestroy(c->ev_wm);
  if (c->s_in) cudaStreamDestroy(c->s_in);
  if (c->s_run) cudaStreamDestroy(c->s_run);
  if (c->s_out) cudaStreamDestroy(c->s_out);
  cudaGetLastError();
  delete c;
}

// after a failed enqueue: nothing of this context may still be reading or writing the caller's
// host buffers when the error is returned
int drain_and_return(tmf_ctx* c, int rc) {
  cudaStreamSynchronize(c->s_in); cudaStreamSynchronize(c->s_run); cudaStreamSynchronize(c->s_out);
  cudaGetLastError();
  return rc;
}

}  // namespace

int tmf_ctx_create(tmf_ctx** out, int device, size_t chunk_bytes, int depth) {
  if (!out) return fail(TMF_ERR_BAD_ARG, "null context pointer");
  *out = nullptr;
  if (depth < 1 || depth > TMF_CTX_MAX_DEPTH) return fail(TMF_ERR_BAD_ARG, "depth must be 1..%d", TMF_CTX_MAX_DEPTH);
  if (chunk_bytes == 0) chunk_bytes = (size_t)96 << 20;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess) return cuda_fail("cudaGetDeviceCount", e);
  if (device < 0 || device >= ndev) return fail(TMF_ERR_BAD_ARG, "device %d out of range (%d visible)", device, ndev);
  DeviceGuard guard(device);
  if (!guard.ok) return fail(TMF_ERR_CUDA, "cannot select device %d", device);
  tmf_ctx* c = new tmf_ctx();
  memset(c, 0, sizeof *c);
  c->device = device; c->depth = depth; c->chunk_bytes = chunk_bytes;
  auto init = [&]() -> int {
    TMF_CUDA(cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking));
    TMF_CUDA(cudaStreamCreateWithFlags(&c->s_run, cudaStreamNonBlocking));
    TMF_CUDA(cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking));
    TMF_CUDA(cudaEventCreateWithFlags(&c->ev_wm, cudaEventDisableTiming));
    for (int k = 0; k < depth; ++k) {
      TMF_CUDA(cudaEventCreateWithFlags(&c->slot[k].ev_in, cudaEventDisableTiming));
      TMF_CUDA(cudaEventCreateWithFlags(&c->slot[k].ev_run, cudaEventDisableTiming));
      TMF_CUDA(cudaEventCreateWithFlags(&c->slot[k].ev_out, cudaEventDisableTiming));
    }
    return TMF_OK;
  };
  if (int rc = init()) { release(c); return rc; }
  *out = c;
  return TMF_OK;
}

int tmf_ctx_destroy(tmf_ctx* c) {
  if (!c) return TMF_OK;
  DeviceGuard guard(c->device);
  cudaStreamSynchronize(c->s_in); cudaStreamSynchronize(c->s_run); cudaStreamSynchronize(c->s_out);
  release(c);
  return TMF_OK;
}

// kind 0 = embed (a = images, b unused), kind 1 = extract (a = watermarked, b = originals)
static int ctx_enqueue_body(tmf_ctx* c, int kind, const uint8_t* a, const uint8_t* b, uint8_t* out, int n, int h, int w,
                       const uint8_t* wm, int wm_shared, double alpha, int block, int mode) {
  if (!c) return fail(TMF_ERR_BAD_ARG, "null context");
  BlockGeom g;
  const size_t img = (size_t)h * w * 3;
  if (int rc = make_geom(n, h, w, img, block, g)) return rc;
  if (mode != TMF_MODE_FAITHFUL && mode != TMF_MODE_FAST && mode != TMF_MODE_LITERAL)
    return fail(TMF_ERR_BAD_ARG, "unknown mode %d", mode);
  if (n == 0 || img == 0) return TMF_OK;
  if (!a || !out || (kind == 1 && !b)) return fail(TMF_ERR_BAD_ARG, "null host pointer");
  const size_t map = (size_t)g.blocks_per_img;
  if (kind == 0 && map > 0 && !wm) return fail(TMF_ERR_BAD_ARG, "null watermark map");
  if (kind == 1 && map == 0) return TMF_OK;
  DeviceGuard guard(c->device);
  if (!guard.ok) return fail(TMF_ERR_CUDA, "cannot select device %d", c->device);
  const size_t out_per_img = (kind == 0) ? img : map;
  size_t per = c->chunk_bytes / img;
  if (per < 1) per = 1;
  if (per > (size_t)n) per = (size_t)n;
  if (kind == 0 && wm_shared && map > 0) {
    // a kernel of the previous call may still be reading the shared map: the copy stream waits for
    // it on the device (no host synchronisation; growing the buffer goes through cudaFree, which
    // waits for the device by itself)
    if (c->wm_used) TMF_CUDA(cudaStreamWaitEvent(c->s_in, c->ev_wm, 0));
    if (int rc = grow(&c->wm_shared, &c->cap_wm_shared, map)) return rc;
    TMF_CUDA(cudaMemcpyAsync(c->wm_shared, wm, map, cudaMemcpyHostToDevice, c->s_in));
    c->h2d_bytes += (long long)map;
  }
  int step = 0;
  for (size_t lo = 0; lo < (size_t)n; lo += per, ++step) {
    const size_t k = ((size_t)n - lo < per) ? (size_t)n - lo : per;
    tmf_ctx::Slot& sl = c->slot[step % c->depth];
    // H2D (waits until the kernel that last read this slot's inputs has finished)
    if (sl.used) TMF_CUDA(cudaStreamWaitEvent(c->s_in, sl.ev_run, 0));
    if (int rc = grow(&sl.a, &sl.cap_a, per * img)) return rc;
    TMF_CUDA(cudaMemcpyAsync(sl.a, a + lo * img, k * img, cudaMemcpyHostToDevice, c->s_in));
    c->h2d_bytes += (long long)(k * img);
    if (kind == 1) {
      if (int rc = grow(&sl.b, &sl.cap_b, per * img)) return rc;
      TMF_CUDA(cudaMemcpyAsync(sl.b, b + lo * img, k * img, cudaMemcpyHostToDevice, c->s_in));
      c->h2d_bytes += (long long)(k * img);
    } else if (!wm_shared && map > 0) {
      if (int rc = grow(&sl.wm, &sl.cap_wm, per * map)) return rc;
      TMF_CUDA(cudaMemcpyAsync(sl.wm, wm + lo * map, k * map, cudaMemcpyHostToDevice, c->s_in));
      c->h2d_bytes += (long long)(k * map);
    }
    TMF_CUDA(cudaEventRecord(sl.ev_in, c->s_in));
    // kernel (waits for the inputs, and for the D2H that last read this slot's output)
    TMF_CUDA(cudaStreamWaitEvent(c->s_run, sl.ev_in, 0));
    if (sl.used) TMF_CUDA(cudaStreamWaitEvent(c->s_run, sl.ev_out, 0));
    if (int rc = grow(&sl.o, &sl.cap_o, per * out_per_img)) return rc;
    int rc = (kind == 0)
        ? tmf_embed_rgb8(sl.a, sl.o, (int)k, h, w, img, wm_shared ? c->wm_shared : sl.wm, wm_shared, alpha, block, mode, c->s_run)
        : tmf_extract_rgb8(sl.a, sl.b, sl.o, (int)k, h, w, img, alpha, block, mode, c->s_run);
    if (rc) return rc;
    c->launches += 1;
    TMF_CUDA(cudaEventRecord(sl.ev_run, c->s_run));
    // D2H
    TMF_CUDA(cudaStreamWaitEvent(c->s_out, sl.ev_run, 0));
    TMF_CUDA(cudaMemcpyAsync(out + lo * out_per_img, sl.o, k * out_per_img, cudaMemcpyDeviceToHost, c->s_out));
    c->d2h_bytes += (long long)(k * out_per_img);
    TMF_CUDA(cudaEventRecord(sl.ev_out, c->s_out));
    sl.used = true;
  }
  if (kind == 0 && wm_shared && map > 0) {
    TMF_CUDA(cudaEventRecord(c->ev_wm, c->s_run));
    c->wm_used = true;
  }
  return TMF_OK;
}

static int ctx_enqueue(tmf_ctx* c, int kind, const uint8_t* a, const uint8_t* b, uint8_t* out, int n, int h, int w,
                       const uint8_t* wm, int wm_shared, double alpha, int block, int mode) {
  const int rc = ctx_enqueue_body(c, kind, a, b, out, n, h, w, wm, wm_shared, alpha, block, mode);
  if (rc != TMF_OK && c) {
    // chunks enqueued before the failure are still in flight: join them before the caller sees the
    // error (it may free its host buffers at once); the message of the failure is kept
    DeviceGuard guard(c->device);
    return drain_and_return(c, rc);
  }
  return rc;
}

int tmf_ctx_embed_host_async(tmf_ctx* c, const uint8_t* rgb, uint8_t* out, int n, int h, int w, const uint8_t* wm,
                             int wm_shared, double alpha, int block, int mode) {
  return ctx_enqueue(c, 0, rgb, nullptr, out, n, h, w, wm, wm_shared, alpha, block, mode);
}

int tmf_ctx_extract_host_async(tmf_ctx* c, const uint8_t* wmk_rgb, const uint8_t* orig_rgb, uint8_t* out_wm, int n,
                               int h, int w, double alpha, int block, int mode) {
  return ctx_enqueue(c, 1, wmk_rgb, orig_rgb, out_wm, n, h, w, nullptr, 0, alpha, block, mode);
}

int tmf_ctx_synchronize(tmf_ctx* c) {
  if (!c) return fail(TMF_ERR_BAD_ARG, "null context");
  DeviceGuard guard(c->device);
  TMF_CUDA(cudaStreamSynchronize(c->s_in));
  TMF_CUDA(cudaStreamSynchronize(c->s_run));
  TMF_CUDA(cudaStreamSynchronize(c->s_out));
  return TMF_OK;
}

int tmf_ctx_stats(tmf_ctx* c, long long* launches, long long* h2d_bytes, long long* d2h_bytes, int reset) {
  if (!c) return fail(TMF_ERR_BAD_ARG, "null context");
  if (launches) *launches = c->launches;
  if (h2d_bytes) *h2d_bytes = c->h2d_bytes;
  if (d2h_bytes) *d2h_bytes = c->d2h_bytes;
  if (reset) c->launches = c->h2d_bytes = c->d2h_bytes = 0;
  return TMF_OK;
}

int tmf_pin_host(void* p, size_t bytes) {
  if (!p || bytes == 0) return fail(TMF_ERR_BAD_ARG, "null pointer or zero size");
  TMF_CUDA(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
  return TMF_OK;
}

int tmf_unpin_host(void* p) {
  if (!p) return fail(TMF_ERR_BAD_ARG, "null pointer");
  TMF_CUDA(cudaHostUnregister(p));
  return TMF_OK;
}

}  // extern "C"
