// FAST mode for the other block sizes of the reference's UI (embed_watermark_page.py:324-331:
// 4..16, even; 8 has its own tuned kernels in fast_kernels.cu).  Same streaming algebra
// (tmf_fast.cuh is templated on N): pass 1 accumulates the Gram matrix of the block's luma
// row by row, certified power iteration, pass 2 applies the rank-1 update and the colour
// round trip.  One thread per block.
//
// A block row is 3N bytes = 3N/2 halfwords, held in ceil(3N/4) 32-bit words (w[k] = bytes
// 4k..4k+3 of the row).  As in the block-8 kernels the luma is the exact integer
// 299 r + 587 g + 114 b, computed by two IDP.2A per pixel on the packed words (no byte
// extraction), and pass 2 floors straight onto the integer level with the subnormal
// quantiser (tmf_rowmath.cuh).
//
// Access width AL.  16: N = 16 with 16-byte aligned rows (three LDG.128 / STG.128 per 48-byte block row
// instead of twelve 32-bit accesses).  4: block rows start 4-byte aligned (3N % 4 == 0: N = 4, 12, 16).  3: sizes with
// 3N % 4 == 2 (N = 6, 10, 14) in rows that start 4-byte aligned: blocks alternate between offsets
// 0 and 2 (mod 4) along a row, so every lane reads the aligned words that cover its row and
// funnel-shifts by 0 or 16 bits, and writes aligned words too - the one word an even block shares
// with its right-hand neighbour is completed with a shuffle and written by the even lane
// (2-byte stores cost 16 partial writes per 32-byte sector: bs 10 embed 132k -> 441k MP/s,
// bs 6 300k -> 593k).  The two bytes an even block reads past its end belong to the next block or
// to the W % N strip: 3N % 4 == 2 with 4-byte rows and no strip implies an even block count per row.
// 2 / 1: halfword / byte accesses for everything else.
#include <type_traits>

#include "tmf_common.cuh"
#include "tmf_rown.cuh"

namespace tmfi {
namespace {

// ---- row I/O: AL = 3 is the aligned-words-with-parity path, the others are tmf_common.cuh's ----
// load_row = fetch_row (the loads, nothing that waits for them) + settle_row (the parity shift of AL = 3): the row
// loops fetch row i + 1 before they work on row i, so the split keeps the shift from dragging the wait forward.
template <int N, int AL>
__device__ __forceinline__ void fetch_row(const uint8_t* __restrict__ p, bool odd, uint32_t (&r)[kRowWords<N>]) {
  if (AL == 16) {                       // N = 16: a block row is 48 bytes = three 16-byte words
    const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
    for (int k = 0; k < (3 * N) / 16; ++k) {
      const uint4 v = __ldg(q + k);
      r[4 * k] = v.x; r[4 * k + 1] = v.y; r[4 * k + 2] = v.z; r[4 * k + 3] = v.w;
    }
  } else if (AL == 3) {
    const uint32_t* q = reinterpret_cast<const uint32_t*>(p - (odd ? 2 : 0));
#pragma unroll
    for (int k = 0; k < kRowWords<N>; ++k) r[k] = __ldg(q + k);
  } else {
    load_row_n<N, (AL == 3 || AL == 16 ? 2 : AL)>(p, r);
  }
}
template <int N, int AL>
__device__ __forceinline__ void settle_row(bool odd, const uint32_t (&r)[kRowWords<N>], uint32_t (&w)[kRowWords<N>]) {
  constexpr int NW = kRowWords<N>;
  if (AL == 3) {
    const uint32_t sh = odd ? 16u : 0u;
#pragma unroll
    for (int k = 0; k < NW - 1; ++k) w[k] = __funnelshift_r(r[k], r[k + 1], sh);
    w[NW - 1] = r[NW - 1] >> sh;          // even lanes: the upper half is the neighbour's (never used)
  } else {
#pragma unroll
    for (int k = 0; k < NW; ++k) w[k] = r[k];
  }
}

template <int N, int AL>
__device__ __forceinline__ void store_row(uint8_t* __restrict__ p, bool odd, bool paired, const uint32_t (&o)[kRowWords<N>]) {
  if (AL == 16) {
    uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
    for (int k = 0; k < (3 * N) / 16; ++k) q[k] = make_uint4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
  } else if (AL == 3) {
    constexpr int NW = kRowWords<N>;
    // The word an even block shares with its right-hand neighbour is completed with a shuffle and
    // written by the even lane.  `paired`: that neighbour is the next lane (same warp, same row) for an
    // even block / the previous lane for an odd one; an unpaired lane (lane 31, lane 0, the last block
    // of a row with an odd block count) writes its own halfword instead.
    const uint32_t nb = __shfl_down_sync(__activemask(), o[0], 1);
    uint32_t* q = reinterpret_cast<uint32_t*>(p + (odd ? 2 : 0));
    const uint32_t sh = odd ? 16u : 0u;
#pragma unroll
    for (int k = 0; k < NW - 1; ++k) q[k] = __funnelshift_r(o[k], o[k + 1], sh);
    if (!odd) {
      if (paired) q[NW - 1] = (o[NW - 1] & 0xffffu) | (nb << 16);
      else *reinterpret_cast<uint16_t*>(p + 4 * (NW - 1)) = (uint16_t)o[NW - 1];
    } else if (!paired) {
      *reinterpret_cast<uint16_t*>(p) = (uint16_t)o[0];
    }
  } else {
    store_row_n<N, (AL == 3 || AL == 16 ? 2 : AL)>(p, o);
  }
}

// Row loops fetch one row ahead (ping-pong, two rows per trip) where the extra row of registers does not spill.
template <int N> constexpr bool kRowsAhead = N <= TMF_FASTN_ROWS_AHEAD_MAX_N;

template <int N, int AL>
__device__ __forceinline__ void gram_row_n(bool odd, const uint32_t (&r)[kRowWords<N>], GramPairsN<N>& G) {
  uint32_t w[kRowWords<N>];
  float2 y2[N / 2];
  settle_row<N, AL>(odd, r, w);
  row_luma2_n<N>(w, y2);
  gram_accumulate_row2_n<N>(y2, G);
}

// Rows two at a time, ping-pong: the loads of the next row are in flight while this one is worked on (one
// thread per block leaves few warps per SM - 12-23 % occupancy - so the latency has to be covered in the thread).
template <int N, int AL>
__device__ __forceinline__ void gram_of_block_n(const uint8_t* __restrict__ base, size_t pitch, bool odd, float* gm) {
  GramPairsN<N> G;
  gram_clear_n<N>(G);
  uint32_t ra[kRowWords<N>];
  if constexpr (kRowsAhead<N>) {
    uint32_t rb[kRowWords<N>];
    fetch_row<N, AL>(base, odd, ra);
#pragma unroll 1
    for (int i = 0; i < N; i += 2) {
      fetch_row<N, AL>(base + (size_t)(i + 1) * pitch, odd, rb);
      gram_row_n<N, AL>(odd, ra, G);
      if (i + 2 < N) fetch_row<N, AL>(base + (size_t)(i + 2) * pitch, odd, ra);
      gram_row_n<N, AL>(odd, rb, G);
    }
  } else {
#pragma unroll 1
    for (int i = 0; i < N; ++i) {
      fetch_row<N, AL>(base + (size_t)i * pitch, odd, ra);
      gram_row_n<N, AL>(odd, ra, G);
    }
  }
  gram_pairs_to_sym_n<N>(G, gm);
}

template <int N, int AL>
__global__ void __launch_bounds__(fastn_threads<N>(), fastn_min_ctas<N>())
k_embed_fast_n(const uint8_t* __restrict__ rgb, uint8_t* __restrict__ out, BlockGeom g,
               const uint8_t* __restrict__ wm, int wm_shared, double alpha) {
  const long long gb = (long long)blockIdx.x * fastn_threads<N>() + threadIdx.x;
  if (gb >= g.total_blocks) return;
  long long img; int by, bx;
  uint32_t in_img;
  const size_t org = block_origin<N>(g, gb, img, by, bx, &in_img);
  const uint8_t* src = rgb + org;
  uint8_t* dst = out + org;
  prefetch_block_rows<N>(src, g.pitch32);
  const uint32_t mark = (uint32_t)__ldg(wm + (wm_shared ? in_img : (uint32_t)gb));
  const bool odd = (bx & 1) != 0;
  const int lane = threadIdx.x & 31;
  const bool paired = odd ? lane > 0 : (lane < 31 && bx + 1 < g.nbw && gb + 1 < g.total_blocks);
  float w[N], f = 0.0f, c = 0.0f;
#pragma unroll
  for (int i = 0; i < N; ++i) w[i] = 0.0f;
  if (mark != 0) {
    float gm[N * (N + 1) / 2];
    gram_of_block_n<N, AL>(src, g.row_pitch, odd, gm);
    tmf::embed_block_scalars_fast<N>(gm, alpha, mark, w, f, c, nullptr, TMF_LUMA_UNIT);
  }
  uint32_t ra[kRowWords<N>];
  if constexpr (kRowsAhead<N>) {
    uint32_t rb[kRowWords<N>];
    fetch_row<N, AL>(src, odd, ra);
#pragma unroll 1
    for (int i = 0; i < N; i += 2) {
      uint32_t wd[kRowWords<N>], o[kRowWords<N>];
      fetch_row<N, AL>(src + (size_t)(i + 1) * g.row_pitch, odd, rb);
      settle_row<N, AL>(odd, ra, wd);
      embed_row_n<N>(wd, w, f, c, mark != 0, o);
      store_row<N, AL>(dst + (size_t)i * g.row_pitch, odd, paired, o);
      if (i + 2 < N) fetch_row<N, AL>(src + (size_t)(i + 2) * g.row_pitch, odd, ra);
      settle_row<N, AL>(odd, rb, wd);
      embed_row_n<N>(wd, w, f, c, mark != 0, o);
      store_row<N, AL>(dst + (size_t)(i + 1) * g.row_pitch, odd, paired, o);
    }
  } else {
#pragma unroll 1
    for (int i = 0; i < N; ++i) {
      uint32_t wd[kRowWords<N>], o[kRowWords<N>];
      fetch_row<N, AL>(src + (size_t)i * g.row_pitch, odd, ra);
      settle_row<N, AL>(odd, ra, wd);
      embed_row_n<N>(wd, w, f, c, mark != 0, o);
      store_row<N, AL>(dst + (size_t)i * g.row_pitch, odd, paired, o);
    }
  }
}

template <int N, int AL>
__device__ __forceinline__ float sigma0_of_block_n(const uint8_t* __restrict__ base, size_t pitch, bool odd) {
  float gm[N * (N + 1) / 2];
  gram_of_block_n<N, AL>(base, pitch, odd, gm);
  return tmf::sigma0_from_gram_fast<N>(gm, nullptr, TMF_LUMA_UNIT);
}

template <int N, int AL>
__global__ void __launch_bounds__(fastn_threads<N>(), fastn_min_ctas<N>())
k_extract_fast_n(const uint8_t* __restrict__ wmk, const uint8_t* __restrict__ orig, uint8_t* __restrict__ out_wm,
                 BlockGeom g, double alpha) {
  const long long gb = (long long)blockIdx.x * fastn_threads<N>() + threadIdx.x;
  if (gb >= g.total_blocks) return;
  long long img; int by, bx;
  const size_t org = block_origin<N>(g, gb, img, by, bx);
  prefetch_block_rows<N>(wmk + org, g.pitch32);
  prefetch_block_rows<N>(orig + org, g.pitch32);
  float sw = 0.0f, so = 0.0f;
#pragma unroll 1
  for (int which = 0; which < 2; ++which) {          // one copy of the code for both images
    const float sg = sigma0_of_block_n<N, AL>((which == 0 ? wmk : orig) + org, g.row_pitch, (bx & 1) != 0);
    if (which == 0) sw = sg; else so = sg;
  }
  out_wm[gb] = (uint8_t)tmf::extract_level(sw, so, alpha);
}

template <int N, int AL>
__global__ void __launch_bounds__(fastn_threads<N>(), fastn_min_ctas<N>())
k_sigma0_fast_n(const uint8_t* __restrict__ rgb, float* __restrict__ sigma0, BlockGeom g) {
  const long long gb = (long long)blockIdx.x * fastn_threads<N>() + threadIdx.x;
  if (gb >= g.total_blocks) return;
  long long img; int by, bx;
  const size_t org = block_origin<N>(g, gb, img, by, bx);
  sigma0[gb] = sigma0_of_block_n<N, AL>(rgb + org, g.row_pitch, (bx & 1) != 0);
}

// Dispatch over the block sizes other than 8.  f(N, A) gets integral constants: N = block size,
// A = access width the block rows allow (4 needs 3N % 4 == 0 || the trailing halfword: any even N
// has 3N % 2 == 0, so rows that start 4-byte aligned can use word loads plus one halfword).
template <int N, typename F>
void with_access_width(int al, bool all16, F&& f) {
  if constexpr (N == 16) {
    if (all16) { f(std::integral_constant<int, N>{}, std::integral_constant<int, 16>{}); return; }
  }
  if (al == 4 && (3 * N) % 4 == 0) f(std::integral_constant<int, N>{}, std::integral_constant<int, 4>{});
  else if (al == 4) f(std::integral_constant<int, N>{}, std::integral_constant<int, 3>{});
  else if (al >= 2) f(std::integral_constant<int, N>{}, std::integral_constant<int, 2>{});
  else f(std::integral_constant<int, N>{}, std::integral_constant<int, 1>{});
}
template <typename F>
void for_block_size(const BlockGeom& g, int al, F&& f, bool even_rows) {
  switch (g.bs) {
    case 4: with_access_width<4>(al, false, f); break;
    case 6: with_access_width<6>(al, false, f); break;
    case 10: with_access_width<10>(al, false, f); break;
    case 12: with_access_width<12>(al, false, f); break;
    case 14: with_access_width<14>(al, false, f); break;
    case 16: with_access_width<16>(al, even_rows, f); break;
    default: break;
  }
}

// pointers, image stride and row pitch all multiples of 16 (block rows of N = 16 then are, too)
bool all16(const BlockGeom& g, const void* p0, const void* p1) {
  return (((uintptr_t)p0 | (uintptr_t)p1 | (uintptr_t)g.img_stride | (uintptr_t)g.row_pitch) & 15) == 0;
}

}  // namespace

int launch_embed_fast_n(const uint8_t* rgb, uint8_t* out, const BlockGeom& g, const uint8_t* wm, int wm_shared,
                        double alpha, cudaStream_t st) {
  for_block_size(g, row_alignment(g, rgb, out), [&](auto n_, auto a_) {
    constexpr int N = decltype(n_)::value, T = fastn_threads<N>();
    k_embed_fast_n<N, decltype(a_)::value><<<grid_for(g.total_blocks, T), T, 0, st>>>(rgb, out, g, wm, wm_shared, alpha);
  }, all16(g, rgb, out));
  return check_launch("embed kernel launch");
}

int launch_extract_fast_n(const uint8_t* wmk, const uint8_t* orig, uint8_t* out_wm, const BlockGeom& g, double alpha,
                          cudaStream_t st) {
  for_block_size(g, row_alignment(g, wmk, orig), [&](auto n_, auto a_) {
    constexpr int N = decltype(n_)::value, T = fastn_threads<N>();
    k_extract_fast_n<N, decltype(a_)::value><<<grid_for(g.total_blocks, T), T, 0, st>>>(wmk, orig, out_wm, g, alpha);
  }, all16(g, wmk, orig));
  return check_launch("extract kernel launch");
}

int launch_sigma0_fast_n(const uint8_t* rgb, float* sigma0, const BlockGeom& g, cudaStream_t st) {
  for_block_size(g, row_alignment(g, rgb, rgb), [&](auto n_, auto a_) {
    constexpr int N = decltype(n_)::value, T = fastn_threads<N>();
    k_sigma0_fast_n<N, decltype(a_)::value><<<grid_for(g.total_blocks, T), T, 0, st>>>(rgb, sigma0, g);
  }, all16(g, rgb, rgb));
  return check_launch("sigma0 kernel launch");
}

}  // namespace tmfi
