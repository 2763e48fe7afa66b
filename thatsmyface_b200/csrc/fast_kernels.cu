// FAST mode, block size 8 (tmf_fast.cuh, tmf_rowmath.cuh): two streaming row passes around a
// certified power iteration on the 8x8 Gram matrix of the block's luma; the block itself is
// never held in registers.
//
//   k_embed_tile   persistent embed kernel fed by the TMA engine: a warp owns a tile of 32 blocks
//       (two boxes of 16 blocks x 8 image rows); ONE cp.async.bulk.tensor per box brings it into
//       shared memory, both row passes run out of shared memory, pass 2 writes in place and ONE
//       cp.async.bulk.tensor per box stores it.  The default wherever the batch is 16-byte aligned
//       and an image row holds a multiple of 16 blocks (512^2, 720p, 1080p, 4K, 8K).
//   k_embed_fast / k_extract_fast / k_sigma0_fast   one thread per block with per-thread global
//       accesses: any size, any alignment.  Extract reads every byte once, so staging it through
//       shared memory only costs occupancy: the per-thread kernel is its fastest form
//       (a TMA-tiled extract measured 0.81-0.89 M MP/s against 1.02 M, profiles/r02_sweep_a_tile_shapes.txt).
#include <cuda.h>   // CUtensorMap, cuTensorMapEncodeTiled's prototype (the entry point comes from the runtime)

#include <atomic>

#include "tmf_common.cuh"
#include "tmf_rowmath.cuh"

namespace tmfi {
namespace {

constexpr float kTwoM49 = 1.7763568394002505e-15f;     // 2^-49: the quantiser's scaling (tmf_rowmath.cuh), exact
constexpr int kRowUnroll = TMF_ROW_UNROLL;
constexpr int kRowUnrollP2 = TMF_ROW_UNROLL_P2;
constexpr int kEmbedThreads = TMF_EMBED_THREADS;
constexpr int kEmbedMinCtas = TMF_EMBED_MIN_CTAS * (128 / TMF_EMBED_THREADS);
constexpr int kExtractThreads = TMF_EXTRACT_THREADS;
constexpr int kExtractMinCtas = TMF_FAST_MIN_CTAS * (128 / TMF_EXTRACT_THREADS);

thread_local int g_last_path = 0;

// ---------------------------------------------------------------------------
// pass 1 over the 8 rows of a block: Gram matrix of its luma (rolled loop: small code).
// With KEEP, row i's luma is parked in shared memory for pass 2, in a thread-private column
// (conflict-free): KEEP = 2: four float2 at col[(4i + p) * STRIDE] (the tile kernels: the packed
// FADD2 results are register PAIRS, two float4 cost 8 MOVs per row to assemble register quads);
// KEEP = 4: two float4 at col4[(2i + q) * STRIDE] (the per-thread kernel, where the two wider
// accesses measured 3 % faster than four narrow ones).
// VEC 8 / 4 / 1: rows in global memory at base + i * pitch; VEC 0: rows of a staged tile in
// shared memory.
// ---------------------------------------------------------------------------
template <int VEC, int KEEP, int STRIDE, typename PITCH>
__device__ __forceinline__ void gram_of_block(const uint8_t* __restrict__ base0, PITCH pitch, float (&gm)[36],
                                              float2* __restrict__ col = nullptr) {
  // uint32_t pitch = a running pointer (one 64-bit add per row; extract / sigma0: +0.5 %);
  // size_t pitch = base + i * pitch (k_embed_fast, where the running pointer measured 1.7 % slower;
  // profiles/r01_sweep_variants.txt table 15); VEC 0 = shared memory, constant offsets
  constexpr bool kRunning = sizeof(PITCH) == 4 && VEC != 0;
  const uint8_t* __restrict__ base = base0;
  GramPairs G;
  gram_clear(G);
#pragma unroll kRowUnroll
  for (int i = 0; i < 8; ++i) {
    uint32_t w[6];
    float2 y2[4];
    if (kRunning) { load_row24<VEC>(base, w); base += pitch; }
    else load_row24<VEC>(base0 + (size_t)i * pitch, w);
    row_luma2<KEEP == 0 && TMF_EXTRACT_I2F>(w, y2);
    if (KEEP == 2) {
#pragma unroll
      for (int p = 0; p < 4; ++p) col[(4 * i + p) * STRIDE] = y2[p];
    } else if (KEEP == 4) {
      float4* col4 = reinterpret_cast<float4*>(col);
      col4[(2 * i) * STRIDE] = make_float4(y2[0].x, y2[0].y, y2[1].x, y2[1].y);
      col4[(2 * i + 1) * STRIDE] = make_float4(y2[2].x, y2[2].y, y2[3].x, y2[3].y);
    }
    gram_accumulate_row2(y2, G);
  }
  gram_pairs_to_sym(G, gm);
}

// ---------------------------------------------------------------------------
// Fused embed, per-thread accesses.
//
// Blocks whose watermark value is 0 get d = f32(f64(s0) + alpha*0) - s0 = 0 exactly
// (watermarking.py:198): their output is the colour round trip alone, so a lane with a zero
// mark skips pass 1 and the eigenpair.
// ---------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(kEmbedThreads, kEmbedMinCtas)
k_embed_fast(const uint8_t* __restrict__ rgb, uint8_t* __restrict__ out, BlockGeom g,
             const uint8_t* __restrict__ wm, int wm_shared, double alpha) {
  __shared__ float4 lum[16 * kEmbedThreads];      // the block's luma, thread-private column
  float4* col = lum + threadIdx.x;
  const long long gb = (long long)blockIdx.x * kEmbedThreads + threadIdx.x;
  if (gb >= g.total_blocks) return;
  long long img; int by, bx;
  uint32_t in_img;
  const size_t org = block_origin<8>(g, gb, img, by, bx, &in_img);
  const uint8_t* src = rgb + org;
  prefetch_block_rows(src, g.pitch32);
  const uint32_t mark = (uint32_t)__ldg(wm + (wm_shared ? in_img : (uint32_t)gb));   // map index: 32 bits
  float w[8], f = 0.0f, c = 0.0f;
  if (mark != 0) {
    float gm[36];
    gram_of_block<VEC, 4, kEmbedThreads, size_t>(src, g.row_pitch, gm, reinterpret_cast<float2*>(col));
    tmf::embed_block_scalars_fast(gm, alpha, mark, w, f, c, nullptr, TMF_LUMA_UNIT);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = 0.0f;
  }
  // pass 2: the rows again (L1/L2 hits), rank-1 update, colour out, quantise, store
  uint8_t* dst = out + org;
  const float2 w2[4] = {make_float2(w[0], w[1]), make_float2(w[2], w[3]), make_float2(w[4], w[5]), make_float2(w[6], w[7])};
  const float f49 = f * kTwoM49, c49 = c * kTwoM49;
#pragma unroll kRowUnrollP2
  for (int i = 0; i < 8; ++i) {
    uint32_t o[6], wd[6];
    float4 ya = make_float4(0.f, 0.f, 0.f, 0.f), yb = ya;
    if (mark != 0) { ya = col[(2 * i) * kEmbedThreads]; yb = col[(2 * i + 1) * kEmbedThreads]; }
    load_row24<VEC>(src + (size_t)i * g.row_pitch, wd);
    const float2 y2[4] = {make_float2(ya.x, ya.y), make_float2(ya.z, ya.w), make_float2(yb.x, yb.y), make_float2(yb.z, yb.w)};
    embed_row_fast2(wd, y2, w2, f49, c49, o);
    store_row24<VEC>(dst + (size_t)i * g.row_pitch, o);
  }
}

template <int VEC>
__global__ void __launch_bounds__(kExtractThreads, kExtractMinCtas)
k_extract_fast(const uint8_t* __restrict__ wmk, const uint8_t* __restrict__ orig, uint8_t* __restrict__ out_wm,
               BlockGeom g, double alpha) {
  const long long gb = (long long)blockIdx.x * kExtractThreads + threadIdx.x;
  if (gb >= g.total_blocks) return;
  long long img; int by, bx;
  const size_t org = block_origin<8>(g, gb, img, by, bx);
  prefetch_block_rows(wmk + org, g.pitch32);
  prefetch_block_rows(orig + org, g.pitch32);
  // the two images go through ONE copy of the code (rolled loop): inlining pass 1 and the
  // eigen-solver twice made the kernel 44 KB and cost ~14 % in instruction-fetch stalls
  float sw = 0.0f, so = 0.0f;
#pragma unroll 1
  for (int which = 0; which < 2; ++which) {
    float gm[36];
    gram_of_block<VEC, 0, 1, uint32_t>((which == 0 ? wmk : orig) + org, g.pitch32, gm);
    const float sg = tmf::sigma0_from_gram_fast(gm, nullptr, TMF_LUMA_UNIT);
    if (which == 0) sw = sg; else so = sg;
  }
  out_wm[gb] = (uint8_t)tmf::extract_level(sw, so, alpha);
}

template <int VEC>
__global__ void __launch_bounds__(kThreads, TMF_FAST_MIN_CTAS)
k_sigma0_fast(const uint8_t* __restrict__ rgb, float* __restrict__ sigma0, BlockGeom g) {
  const long long gb = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (gb >= g.total_blocks) return;
  long long img; int by, bx;
  const size_t org = block_origin<8>(g, gb, img, by, bx);
  float gm[36];
  prefetch_block_rows(rgb + org, g.pitch32);
  gram_of_block<VEC, 0, 1, uint32_t>(rgb + org, g.pitch32, gm);
  sigma0[gb] = tmf::sigma0_from_gram_fast(gm, nullptr, TMF_LUMA_UNIT);
}

// ---------------------------------------------------------------------------
// TMA-tiled persistent embed kernel.
//
// Why: with per-thread row accesses a warp's 64-bit load covers 256 useful bytes spread over
// 768, i.e. 6-7 L1 wavefronts instead of 2, three times per row; the resident CTAs' rows do not
// fit in the L1 left beside the stash, so pass 2 re-reads most rows from L2 (ncu, round 1:
// 33 M load sectors for 6.2 M of input, L1 hit 58 %, 18 % of the warp samples on long_scoreboard),
// and every 32-byte output sector reaches L2 in three partial writes.  Round 1's answer, eight
// per-row cp.async.bulk per warp each way, fixed the memory side and lost on copy-issue
// instructions.  Here:
//
//   * the batch is ONE 3-D tensor (3W/4 words, H rows, N images; strides 3W and img_stride
//     bytes), and a BOX is 16 blocks x 8 rows = 96 words x 8 rows = 3 KB;
//   * a warp's TILE is two boxes, consecutive in the linear box order (same block-row, or
//     wrapping to the next block-row / image): lane L owns block (L & 15) of box (L >> 4), so
//     no lane idles whenever an image row holds a multiple of 16 blocks;
//   * one elected lane issues ONE cp.async.bulk.tensor.3d per box (SASS UTMALDG.3D) onto the
//     warp's mbarrier; the warp waits, runs pass 1, the eigen-solve and pass 2 out of shared
//     memory (three conflict-free LDS.64 / STS.64 per row: thread stride 24 B, constant offsets),
//     writing its output bytes IN PLACE, and ONE cp.async.bulk.tensor.3d store per box
//     (UTMASTG.3D) hands them back: full-line writes, no per-thread STG;
//   * the kernel is persistent: a warp walks tiles t, t + W, t + 2W, ... (W = warps in the grid);
//     warps never synchronise with each other.  Warps per CTA must be a multiple of 4: the walk
//     is static, so a scheduler with one warp more than its neighbours sets the pace;
//   * one 6 KB buffer per warp + the 8 KB luma stash = 16 warps per SM.  A single buffer cannot be
//     refilled before it has been stored; what hides that turn-around is an L2 prefetch
//     (cp.async.bulk.prefetch.tensor, UTMAPF.L2.3D) of the warp's NEXT tile, issued at the start of
//     the current one: the load that follows the store finds its data in L2 (mbarrier wait 16 % ->
//     2.6 % of the warp samples, 794 k -> 842 k MP/s).  Measured and dropped (tables in
//     profiles/r02_sweep_*.txt): a second buffer without the stash (pass 2 recomputes the luma,
//     +9 % instructions: no gain), the buffer handed over in halves (twice the copies and
//     barriers: -7 %), 20-28 warps without stash (dispatch-bound: flat).
// ---------------------------------------------------------------------------
constexpr int kBoxBlocks = 16;                        // blocks per box row
constexpr int kBoxRowBytes = kBoxBlocks * 24;         // 384
constexpr int kBoxBytes = 8 * kBoxRowBytes;           // 3072
constexpr int kTileBytes = 2 * kBoxBytes;             // 6144 per warp and stage
constexpr int kBoxWords = kBoxRowBytes / 4;           // 96: box width in tensor elements (u32)

struct TileGeom {
  uint32_t total_boxes, total_tiles;
  uint32_t boxes_per_row, boxes_per_img, nbw, blocks_per_img;
  FastDiv div_bpi, div_bpr;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "TMF_WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@!p bra TMF_WAIT_%=;\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_box(uint32_t dst, const CUtensorMap* map, uint32_t c0, uint32_t c1, uint32_t c2,
                                             uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store_box(const CUtensorMap* map, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t src) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
               ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(src) : "memory");
}
__device__ __forceinline__ void tma_prefetch_box(const CUtensorMap* map, uint32_t c0, uint32_t c1, uint32_t c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];"
               ::"l"(map), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Tile control lives on the UNIFORM datapath: everything that is the same for the whole warp - the
// tile index, both boxes' coordinates, the barrier phase - is computed once per tile, warp-uniformly
// (the warp index comes from a shuffle, which tells the compiler so), and ONE elected lane issues
// the loads / stores / prefetches of both boxes.  (A first version let lanes 0 and 16 compute their
// own box and issue their own copies: the compiler serialised the two lanes and moved every operand
// with R2UR, 157 of 1 904 instructions per tile; this form is +1.8 %, profiles/r02_sweep_i_uniform.txt.)
// Lane-specific: only the map index of the lane's block.
template <int WARPS>
struct EmbedTileSmem {
  static constexpr int kTiles = WARPS * kTileBytes;
  static constexpr int kStash = WARPS * 32 * 32 * 8;                     // 32 float2 per thread
  static constexpr int kBars = WARPS * 8;
  static constexpr int kTotal = kTiles + kStash + kBars;
};

struct TileAt {          // warp-uniform: tensor coordinates of the tile's two boxes
  uint32_t c0a, c1a, c2a, c0b, c1b, c2b;
  uint32_t bxa, bxb;     // first block column of each box (for the map index)
  uint32_t nb;           // 2, or 1 for the last tile of an odd box count
};
__device__ __forceinline__ TileAt tile_at(const TileGeom& tg, uint32_t t) {
  TileAt a;
  const uint32_t b0 = 2u * t;
  const uint32_t img = fastdiv(b0, tg.div_bpi);
  const uint32_t r = b0 - img * tg.boxes_per_img;
  const uint32_t by = fastdiv(r, tg.div_bpr);
  const uint32_t bcol = r - by * tg.boxes_per_row;      // box column inside the block-row
  a.c0a = bcol * (uint32_t)kBoxWords; a.c1a = by * 8u; a.c2a = img;
  a.bxa = bcol * (uint32_t)kBoxBlocks;
  a.nb = (b0 + 1u < tg.total_boxes) ? 2u : 1u;
  const bool same_row = bcol + 1u < tg.boxes_per_row;
  const bool same_img = r + 1u < tg.boxes_per_img;
  a.c0b = same_row ? a.c0a + (uint32_t)kBoxWords : 0u;
  a.bxb = same_row ? a.bxa + (uint32_t)kBoxBlocks : 0u;
  a.c1b = same_row ? a.c1a : (same_img ? a.c1a + 8u : 0u);
  a.c2b = same_img ? img : img + 1u;
  return a;
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

template <int WARPS, int CTAS_PER_SM>
__global__ void __launch_bounds__(WARPS * 32, CTAS_PER_SM)
k_embed_tile(const __grid_constant__ CUtensorMap src_map, const __grid_constant__ CUtensorMap dst_map, TileGeom tg,
               const uint8_t* __restrict__ wm, int wm_shared, double alpha) {
  using L = EmbedTileSmem<WARPS>;
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);       // warp-uniform, and the compiler knows it
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t half = lane >> 4, l16 = lane & 15u;
  uint8_t* tiles = smem + warp * kTileBytes;
  const uint32_t tiles_s = smem_u32(smem) + warp * kTileBytes;
  const uint32_t bar = smem_u32(smem) + L::kTiles + L::kStash + warp * 8;
  float2* col = reinterpret_cast<float2*>(smem + L::kTiles) + threadIdx.x;
  constexpr int kStride = WARPS * 32;

  const uint32_t nwarps = gridDim.x * WARPS;
  uint32_t t = blockIdx.x * WARPS + warp;
  if (t >= tg.total_tiles) return;                    // whole warp; nothing below is CTA-wide
  if (elect_one()) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
#pragma unroll
  for (int k = 0; k < 32; ++k) col[k * kStride] = make_float2(0.0f, 0.0f);   // lanes with a zero mark read their stash as it is:
  __syncwarp();                                                               // finite values, times f = c = w = 0

  auto load_tile = [&](const TileAt& a) {              // by ONE lane: both boxes onto the warp's barrier
    mbar_expect_tx(bar, a.nb * kBoxBytes);
    tma_load_box(tiles_s, &src_map, a.c0a, a.c1a, a.c2a, bar);
    if (a.nb == 2u) tma_load_box(tiles_s + kBoxBytes, &src_map, a.c0b, a.c1b, a.c2b, bar);
  };
  auto lane_mark = [&](const TileAt& a) -> uint32_t {  // the lane's watermark value (0 for a missing second box)
    const uint32_t bx = half ? a.bxb : a.bxa, c1 = half ? a.c1b : a.c1a, c2 = half ? a.c2b : a.c2a;
    if (half && a.nb != 2u) return 0u;
    const uint32_t idx = (c1 >> 3) * tg.nbw + bx + l16;
    return (uint32_t)__ldg(wm + (wm_shared ? 0u : c2 * tg.blocks_per_img) + idx);
  };

  TileAt at = tile_at(tg, t);
  if (elect_one()) load_tile(at);
  uint32_t mark = lane_mark(at);
  uint8_t* mine = tiles + half * kBoxBytes + l16 * 24u;

  for (uint32_t it = 0;; ++it) {
    const uint32_t tn = t + nwarps;
    const bool more = tn < tg.total_tiles;
    const bool valid = !(half && at.nb != 2u);
    TileAt nx = at;
    uint32_t mark_n = 0;
    if (more) {
      nx = tile_at(tg, tn);
      if (elect_one()) {                               // ask L2 for the next tile now: the load after the store finds it there
        tma_prefetch_box(&src_map, nx.c0a, nx.c1a, nx.c2a);
        if (nx.nb == 2u) tma_prefetch_box(&src_map, nx.c0b, nx.c1b, nx.c2b);
      }
      mark_n = lane_mark(nx);
    }
    mbar_wait(bar, it & 1u);

    float w[8], f = 0.0f, c = 0.0f;
    if (valid && mark != 0) {
      float gm[36];
      gram_of_block<0, 2, kStride, uint32_t>(mine, (uint32_t)kBoxRowBytes, gm, col);
      tmf::embed_block_scalars_fast(gm, alpha, mark, w, f, c, nullptr, TMF_LUMA_UNIT);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) w[i] = 0.0f;
    }
    if (valid) {
      const float2 w2[4] = {make_float2(w[0], w[1]), make_float2(w[2], w[3]), make_float2(w[4], w[5]), make_float2(w[6], w[7])};
      const float f49 = f * kTwoM49, c49 = c * kTwoM49;
#pragma unroll kRowUnrollP2
      for (int i = 0; i < 8; ++i) {
        uint32_t o[6], wd[6];
        load_row24<0>(mine + i * kBoxRowBytes, wd);
        float2 y2[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) y2[p] = col[(4 * i + p) * kStride];       // zero-mark lanes: an earlier tile's values, times 0
        embed_row_fast2(wd, y2, w2, f49, c49, o);
        store_row24<0>(mine + i * kBoxRowBytes, o);
      }
    }
    // generic-proxy writes -> visible to the async proxy, then one lane hands both boxes over
    fence_async_smem();
    __syncwarp();
    if (elect_one()) {
      tma_store_box(&dst_map, at.c0a, at.c1a, at.c2a, tiles_s);
      if (at.nb == 2u) tma_store_box(&dst_map, at.c0b, at.c1b, at.c2b, tiles_s + kBoxBytes);
      bulk_commit();
      bulk_wait_read0();                               // the buffer has been read out: refill it (or leave)
      if (more) load_tile(nx);
    }
    if (!more) break;
    t = tn;
    at = nx;
    mark = mark_n;
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static std::atomic<EncodeTiledFn> cached{nullptr};
  EncodeTiledFn f = cached.load(std::memory_order_acquire);
  if (f) return f;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess || !p) {
    cudaGetLastError();
    return nullptr;
  }
  f = reinterpret_cast<EncodeTiledFn>(p);
  cached.store(f, std::memory_order_release);
  return f;
}

// The tile kernels need: block size 8, every image row a multiple of 16 blocks wide... of the
// part that holds whole blocks; 16-byte aligned base, row pitch and image stride (TMA's rules).
bool tile_ok(const BlockGeom& g, const void* p0, const void* p1, const void* p2 = nullptr) {
  static const bool disabled = [] { const char* e = getenv("TMF_NO_TILE"); return e && e[0] == '1'; }();
  if (disabled) return false;
  const uintptr_t bits = (uintptr_t)p0 | (uintptr_t)p1 | (uintptr_t)p2 | (uintptr_t)g.img_stride | (uintptr_t)g.row_pitch;
  return g.bs == 8 && (bits & 15) == 0 && g.nbw >= kBoxBlocks && (g.nbw % kBoxBlocks) == 0 && g.nbh > 0 &&
         g.img_stride < (1ull << 40);
}

int make_tile_geom(const BlockGeom& g, TileGeom& tg) {
  tg.nbw = (uint32_t)g.nbw;
  tg.blocks_per_img = (uint32_t)g.blocks_per_img;
  tg.boxes_per_row = (uint32_t)(g.nbw / kBoxBlocks);
  tg.boxes_per_img = tg.boxes_per_row * (uint32_t)g.nbh;
  tg.total_boxes = (uint32_t)(g.total_blocks / kBoxBlocks);
  tg.total_tiles = (tg.total_boxes + 1u) / 2u;
  tg.div_bpi = make_fastdiv(tg.boxes_per_img);
  tg.div_bpr = make_fastdiv(tg.boxes_per_row);
  return TMF_OK;
}

int make_map(CUtensorMap* m, const void* base, const BlockGeom& g, int n) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return fail(TMF_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t dims[3] = {(cuuint64_t)(g.row_pitch / 4), (cuuint64_t)g.nbh * 8u, (cuuint64_t)n};
  const cuuint64_t strides[2] = {(cuuint64_t)g.row_pitch, (cuuint64_t)g.img_stride};
  const cuuint32_t box[3] = {(cuuint32_t)kBoxWords, 8u, 1u};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<void*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(TMF_ERR_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
  return TMF_OK;
}

// opt in to > 48 KB of dynamic shared memory, once per device and kernel (idempotent, thread-safe)
template <typename K>
int set_smem(K kernel, int bytes, std::atomic<unsigned long long>& done) {
  int dev = 0;
  TMF_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && (done.load(std::memory_order_acquire) >> dev) & 1ull) return TMF_OK;
  TMF_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  TMF_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  if (dev < 64) done.fetch_or(1ull << dev, std::memory_order_release);
  return TMF_OK;
}

}  // namespace

int last_fast_path() { return g_last_path; }

int launch_embed_fast8(const uint8_t* rgb, uint8_t* out, const BlockGeom& g, const uint8_t* wm, int wm_shared,
                       double alpha, cudaStream_t st) {
  if (tile_ok(g, rgb, out)) {
    constexpr int W = TMF_TILE_WARPS, C = TMF_TILE_CTAS_PER_SM;
    auto kernel = k_embed_tile<W, C>;
    constexpr int smem = EmbedTileSmem<W>::kTotal;
    static std::atomic<unsigned long long> done{0};
    if (int rc = set_smem(kernel, smem, done)) return rc;
    TileGeom tg;
    make_tile_geom(g, tg);
    const int n = (int)(g.total_blocks / g.blocks_per_img);
    CUtensorMap ms, md;
    if (make_map(&ms, rgb, g, n) == TMF_OK && make_map(&md, out, g, n) == TMF_OK) {
      const unsigned cap = (unsigned)sm_count() * C;
      const unsigned need = grid_for(tg.total_tiles, W);
      kernel<<<need < cap ? need : cap, W * 32, smem, st>>>(ms, md, tg, wm, wm_shared, alpha);
      g_last_path = 1;
      return check_launch("embed (tile) kernel launch");
    }
    // no tensor maps from this driver (or a shape it refuses): the per-thread kernel computes the same bytes
  }
  const unsigned grid = grid_for(g.total_blocks, kEmbedThreads);
  switch (pick_vec(g, rgb, out)) {
    case 8: k_embed_fast<8><<<grid, kEmbedThreads, 0, st>>>(rgb, out, g, wm, wm_shared, alpha); break;
    case 4: k_embed_fast<4><<<grid, kEmbedThreads, 0, st>>>(rgb, out, g, wm, wm_shared, alpha); break;
    default: k_embed_fast<1><<<grid, kEmbedThreads, 0, st>>>(rgb, out, g, wm, wm_shared, alpha); break;
  }
  g_last_path = 0;
  return check_launch("embed kernel launch");
}

int launch_extract_fast8(const uint8_t* wmk, const uint8_t* orig, uint8_t* out_wm, const BlockGeom& g, double alpha,
                         cudaStream_t st) {
  const unsigned grid = grid_for(g.total_blocks, kExtractThreads);
  switch (pick_vec(g, wmk, orig)) {
    case 8: k_extract_fast<8><<<grid, kExtractThreads, 0, st>>>(wmk, orig, out_wm, g, alpha); break;
    case 4: k_extract_fast<4><<<grid, kExtractThreads, 0, st>>>(wmk, orig, out_wm, g, alpha); break;
    default: k_extract_fast<1><<<grid, kExtractThreads, 0, st>>>(wmk, orig, out_wm, g, alpha); break;
  }
  g_last_path = 0;
  return check_launch("extract kernel launch");
}

int launch_sigma0_fast8(const uint8_t* rgb, float* sigma0, const BlockGeom& g, cudaStream_t st) {
  const unsigned grid = grid_for(g.total_blocks, kThreads);
  switch (pick_vec(g, rgb, rgb)) {
    case 8: k_sigma0_fast<8><<<grid, kThreads, 0, st>>>(rgb, sigma0, g); break;
    case 4: k_sigma0_fast<4><<<grid, kThreads, 0, st>>>(rgb, sigma0, g); break;
    default: k_sigma0_fast<1><<<grid, kThreads, 0, st>>>(rgb, sigma0, g); break;
  }
  return check_launch("sigma0 kernel launch");
}

}  // namespace tmfi
