// Launch-shape knobs of the fused kernels - the ONLY compile-time tunables of the library.
// profiles/sweep_variants.py rebuilds the library with -D overrides of these and times the
// alternatives on the GPU; the defaults below are the measured winners
// (profiles/r01_sweep_variants.txt, profiles/r02_sweep_*.txt).  Every knob whose sweep
// verdict was "loses" has been deleted together with its code path; the tables are the record.
#pragma once

// rows per iteration of the rolled row loops of the per-thread fast kernels (1, 2, 4, 8)
#ifndef TMF_ROW_UNROLL
#define TMF_ROW_UNROLL 8
#endif
#ifndef TMF_ROW_UNROLL_P2
#define TMF_ROW_UNROLL_P2 4                // the embed kernels' pass 2, separately
#endif

// per-thread fast kernels (k_embed_fast / k_extract_fast / k_sigma0_fast): CTA size and the
// __launch_bounds__ minimum CTAs per SM, given per 128 threads
#ifndef TMF_EMBED_THREADS
#define TMF_EMBED_THREADS 32     // one-warp CTAs give registers and stash back as soon as their warp ends
#endif
#ifndef TMF_EMBED_MIN_CTAS
#define TMF_EMBED_MIN_CTAS 5     // 96 registers: no spill; 160 KB of stash leave 60 KB of L1
#endif
#ifndef TMF_EXTRACT_THREADS
#define TMF_EXTRACT_THREADS 32
#endif
#ifndef TMF_EXTRACT_I2F
#define TMF_EXTRACT_I2F 1        // extract / sigma0: luma integers to float by I2F (conversion pipe) instead of FADD2
#endif
#ifndef TMF_FAST_MIN_CTAS
#define TMF_FAST_MIN_CTAS 6      // extract / sigma0: 80 registers
#endif

// TMA-tiled persistent embed kernel (k_embed_tile): warps per CTA (a multiple of 4) and CTAs per SM;
// a warp needs 14 KB of shared memory (tile + luma stash), so 16 warps fill an SM
#ifndef TMF_TILE_WARPS
#define TMF_TILE_WARPS 16
#endif
#ifndef TMF_TILE_CTAS_PER_SM
#define TMF_TILE_CTAS_PER_SM 1
#endif

// faithful kernels, block size 8: minimum CTAs of 128 threads per SM - the literal form (A and V in registers),
// the default embed (A only; 4 CTAs = 128 registers with 196 spilled bytes measured 186 k against 181 k MP/s alone but
// 161 k against 169-180 k inside bench.py under the power cap: left at 3) and extract / sigma0 (80 registers: 210 k -> 253 k MP/s; 248 k at 5 x 128 threads, 238 k at 4)
#ifndef TMF_FAITHFUL_MIN_CTAS
#define TMF_FAITHFUL_MIN_CTAS 3
#endif
#ifndef TMF_FAITHFUL_R1_MIN_CTAS
#define TMF_FAITHFUL_R1_MIN_CTAS 3
#endif
#ifndef TMF_FAITHFUL_SIGMA_MIN_CTAS
#define TMF_FAITHFUL_SIGMA_MIN_CTAS 6
#endif

// generic-N fast kernels (fast_n_kernels.cu): threads per CTA for sizes 10 ... 16 (4 and 6 run 128) and minimum
// CTAs per SM by block size (4, 6 | 10 | 12 | 14 | 16); the Gram matrix alone is N (N + 1) / 2 registers
#ifndef TMF_FASTN_THREADS
#define TMF_FASTN_THREADS 32
#endif
#ifndef TMF_FASTN_CTAS_SMALL
#define TMF_FASTN_CTAS_SMALL 6
#endif
#ifndef TMF_FASTN_CTAS_10
#define TMF_FASTN_CTAS_10 16       // 128 registers
#endif
#ifndef TMF_FASTN_CTAS_12
#define TMF_FASTN_CTAS_12 12       // 168 registers
#endif
#ifndef TMF_FASTN_CTAS_14
#define TMF_FASTN_CTAS_14 12       // 168 registers and a few spilled words; at 8 (223 registers, no spill) extract is 10 % slower
                                   // (the register file is per scheduler: 9-12 warps per SM all mean 168 registers)
#endif
#ifndef TMF_FASTN_CTAS_16
#define TMF_FASTN_CTAS_16 8        // 255 registers: the 136-entry Gram matrix and a row ahead, no spill
#endif
// ... and the largest block size whose row loops fetch one row ahead (a second row of registers)
#ifndef TMF_FASTN_ROWS_AHEAD_MAX_N
#define TMF_FASTN_ROWS_AHEAD_MAX_N 16
#endif
