// Instruction-throughput microbenchmarks used to choose the per-pixel
// conversion sequences of the fused kernels (results: profiles/r01_ubench.txt).
// One CTA of 1024 threads per SM; every thread runs ITER iterations of 8
// independent chains of the instruction under test; ops/clk/SM from clock64.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 4096

#define BENCH(NAME, DECL, BODY, SINK)                                                        \
  __global__ void __launch_bounds__(1024) NAME(long long* clk, float* sink) {                  \
    DECL;                                                                                    \
    __syncthreads();                                                                         \
    long long t0 = clock64();                                                                \
    _Pragma("unroll 4") for (int it = 0; it < ITER; ++it) { BODY; }                                              \
    long long t1 = clock64();                                                                \
    __syncthreads();                                                                         \
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;                                         \
    SINK;                                                                                    \
  }

#define F8 float a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7
#define SINKF if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 12345.678f) sink[0] = a0
#define R8(OP) OP(a0) OP(a1) OP(a2) OP(a3) OP(a4) OP(a5) OP(a6) OP(a7)

#define OP_FFMA(x) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x) : "f"(1.0000001f), "f"(0.5f));
BENCH(k_ffma, F8, R8(OP_FFMA), SINKF)
#define OP_FADD_RM(x) asm volatile("add.rm.f32 %0, %0, %1;" : "+f"(x) : "f"(0.25f));
BENCH(k_fadd_rm, F8, R8(OP_FADD_RM), SINKF)
#define OP_FMNMX(x) asm volatile("max.f32 %0, %0, %1;" : "+f"(x) : "f"(0.25f));
BENCH(k_fmnmx, F8, R8(OP_FMNMX), SINKF)
#define OP_CLAMP(x) asm volatile("max.f32 %0, %0, %1; min.f32 %0, %0, %2;" : "+f"(x) : "f"(0.0f), "f"(255.0f));
BENCH(k_clamp2, F8, R8(OP_CLAMP), SINKF)
#define OP_RSQ(x) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(x));
BENCH(k_mufu_rsq, F8, R8(OP_RSQ), SINKF)
#define OP_RCP(x) asm volatile("rcp.approx.ftz.f32 %0, %0; add.f32 %0, %0, 0f3F800000;" : "+f"(x));
BENCH(k_mufu_rcp, F8, R8(OP_RCP), SINKF)
#define OP_F2I(x) { unsigned u; asm volatile("cvt.rzi.u32.f32 %0, %1;" : "=r"(u) : "f"(x)); asm volatile("mov.b32 %0, %1;" : "=f"(x) : "r"(u | 0x3f800000u)); }
BENCH(k_f2i_trunc_plus_lop, F8, R8(OP_F2I), SINKF)
#define OP_F2I_SAT8(x) { unsigned u; asm volatile("{ .reg .u8 t; cvt.rzi.sat.u8.f32 t, %1; cvt.u32.u8 %0, t; }" : "=r"(u) : "f"(x)); asm volatile("mov.b32 %0, %1;" : "=f"(x) : "r"(u | 0x3f800000u)); }
BENCH(k_f2i_sat_u8_plus_lop, F8, R8(OP_F2I_SAT8), SINKF)
#define OP_LOP(x) { unsigned u = __float_as_uint(x); asm volatile("or.b32 %0, %0, %1;" : "+r"(u) : "r"(0x3f800000u)); x = __uint_as_float(u); }
BENCH(k_lop_only, F8, R8(OP_LOP), SINKF)
#define OP_I2F(x) { unsigned u = __float_as_uint(x); asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(x) : "r"(u & 0xffu)); }
BENCH(k_i2f_plus_lop, F8, R8(OP_I2F), SINKF)
#define OP_PRMT_MAGIC(x) { unsigned u = __float_as_uint(x); asm volatile("prmt.b32 %0, %0, %1, 0x7650;" : "+r"(u) : "r"(0x4B000000u)); asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(x) : "f"(__uint_as_float(u)), "f"(-8388608.0f)); }
BENCH(k_prmt_fadd_magic, F8, R8(OP_PRMT_MAGIC), SINKF)
#define OP_PRMT(x) { unsigned u = __float_as_uint(x); asm volatile("prmt.b32 %0, %0, %1, 0x7650;" : "+r"(u) : "r"(0x4B000000u)); x = __uint_as_float(u); }
BENCH(k_prmt_only, F8, R8(OP_PRMT), SINKF)
#define OP_I2IP(x) { int u = __float_as_int(x); asm volatile("cvt.pack.sat.u8.s32.b32 %0, %0, %1, %2;" : "+r"(u) : "r"(77), "r"(0)); x = __int_as_float(u); }
BENCH(k_cvt_pack_sat_u8, F8, R8(OP_I2IP), SINKF)

#define D8 double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7
#define SINKD if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 12345.678) sink[0] = (float)a0
#define OP_DFMA(x) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x) : "d"(1.0000001), "d"(0.5));
BENCH(k_dfma, D8, R8(OP_DFMA), SINKD)
#define OP_F2F_UPDOWN(x) { double d; asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d) : "f"(x)); asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(x) : "d"(d)); }
BENCH(k_f2f_up_and_down, F8, R8(OP_F2F_UPDOWN), SINKF)
#define OP_FDIV(x) x = __fdiv_rn(x, 255.0f) + 1.0f;
BENCH(k_fdiv_rn_by_255_plus_fadd, F8, R8(OP_FDIV), SINKF)


#define U8 unsigned a0 = threadIdx.x, a1 = a0 * 3 + 1, a2 = a0 * 5 + 2, a3 = a0 * 7 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7; unsigned kb = clk[0] & 0xff, kc = 0x4B000000u | (unsigned)(clk[1] & 1)
#define SINKU if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 12345u) sink[0] = a0
#define OP_LOP3(x) asm volatile("lop3.b32 %0, %0, %1, %2, 0xEA;" : "+r"(x) : "r"(kb), "r"(kc));
BENCH(k_lop3, U8, R8(OP_LOP3), SINKU)
#define OP_SHF(x) asm volatile("shf.r.clamp.b32 %0, %0, %1, 24;" : "+r"(x) : "r"(kc));
BENCH(k_shf_funnel, U8, R8(OP_SHF), SINKU)
#define OP_IADD(x) asm volatile("add.s32 %0, %0, %1;" : "+r"(x) : "r"(kc));
BENCH(k_iadd, U8, R8(OP_IADD), SINKU)
#define OP_IMAD(x) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x) : "r"(kb), "r"(kc));
BENCH(k_imad, U8, R8(OP_IMAD), SINKU)
#define OP_I2FP(x) { float f; asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(f) : "r"(x)); x = __float_as_uint(f) >> 3; }
BENCH(k_i2fp_u32_plus_shr, U8, R8(OP_I2FP), SINKU)
#define OP_SHR(x) x = x >> 3; asm volatile("" : "+r"(x));
BENCH(k_shr_only, U8, R8(OP_SHR), SINKU)
#define OP_HADD2(x) asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(x) : "r"(kc));
BENCH(k_hadd2, U8, R8(OP_HADD2), SINKU)
#define OP_H2F(x) { float f; asm volatile("{ .reg .f16 lo, hi; mov.b32 {lo, hi}, %1; cvt.f32.f16 %0, lo; }" : "=f"(f) : "r"(x)); x = __float_as_uint(f) | 1u; }
BENCH(k_cvt_f32_f16_plus_lop, U8, R8(OP_H2F), SINKU)


// ---- packed fp32 (sm_100 FFMA2 / FADD2) and co-issue with the ALU pipe
#define P8 float2 a0 = make_float2(threadIdx.x, 1.f), a1 = a0, a2 = a0, a3 = a0, a4 = a0, a5 = a0, a6 = a0, a7 = a0; const float2 m2 = make_float2(1.0000001f, 0.9999999f), c2 = make_float2(0.5f, 0.25f)
#define SINKP if (a0.x + a1.y + a2.x + a3.y + a4.x + a5.y + a6.x + a7.y == 12345.678f) sink[0] = a0.x
#define OP_FFMA2(x) x = __ffma2_rn(x, m2, c2);
BENCH(k_ffma2, P8, R8(OP_FFMA2), SINKP)
#define OP_FADD2RD(x) x = __fadd2_rd(x, c2);
BENCH(k_fadd2_rd, P8, R8(OP_FADD2RD), SINKP)
#define MIXDECL float a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7; \
  unsigned u0 = threadIdx.x, u1 = u0 * 3, u2 = u0 * 5, u3 = u0 * 7, u4 = u0 + 4, u5 = u0 + 5, u6 = u0 + 6, u7 = u0 + 7; unsigned kc = 0x4B000000u | (unsigned)(clk[1] & 1)
#define MIXSINK if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + (float)(u0 + u1 + u2 + u3 + u4 + u5 + u6 + u7) == 12345.678f) sink[0] = a0
#define OP_PR(u) asm volatile("prmt.b32 %0, %0, %1, 0x7651;" : "+r"(u) : "r"(kc));
#define FF8 OP_FFMA(a0) OP_FFMA(a1) OP_FFMA(a2) OP_FFMA(a3) OP_FFMA(a4) OP_FFMA(a5) OP_FFMA(a6) OP_FFMA(a7)
#define PR8 OP_PR(u0) OP_PR(u1) OP_PR(u2) OP_PR(u3) OP_PR(u4) OP_PR(u5) OP_PR(u6) OP_PR(u7)
#define PR4 OP_PR(u0) OP_PR(u1) OP_PR(u2) OP_PR(u3)
BENCH(k_mix_8ffma_8prmt, MIXDECL, FF8 PR8, MIXSINK)
BENCH(k_mix_8ffma_4prmt, MIXDECL, FF8 PR4, MIXSINK)
#define MIX2DECL float2 b0 = make_float2(threadIdx.x, 1.f), b1 = b0, b2 = b0, b3 = b0; const float2 m2 = make_float2(1.0000001f, 0.9999999f), c2 = make_float2(0.5f, 0.25f); \
  unsigned u0 = threadIdx.x, u1 = u0 * 3, u2 = u0 * 5, u3 = u0 * 7, u4 = u0 + 4, u5 = u0 + 5, u6 = u0 + 6, u7 = u0 + 7; unsigned kc = 0x4B000000u | (unsigned)(clk[1] & 1)
#define MIX2SINK if (b0.x + b1.y + b2.x + b3.y + (float)(u0 + u1 + u2 + u3 + u4 + u5 + u6 + u7) == 12345.678f) sink[0] = b0.x
#define F24 OP_FFMA2(b0) OP_FFMA2(b1) OP_FFMA2(b2) OP_FFMA2(b3)
BENCH(k_mix_4ffma2_8prmt, MIX2DECL, F24 PR8, MIX2SINK)
BENCH(k_mix_4ffma2_4prmt, MIX2DECL, F24 PR4, MIX2SINK)


// does a slow-pipe conversion overlap with FMA-pipe work? (8 ffma + n conversions per iteration)
#define OP_F2IP(u, a) asm volatile("{ .reg .u8 t; cvt.rzi.sat.u8.f32 t, %1; cvt.u32.u8 %0, t; }" : "=r"(u) : "f"(a));
#define OP_I2FB(a, u) asm volatile("{ .reg .u8 t; .reg .b8 x0,x1,x2; mov.b32 {t,x0,x1,x2}, %1; cvt.rn.f32.u8 %0, t; }" : "=f"(a) : "r"(u));
BENCH(k_mix_8ffma_2f2ip, MIXDECL, FF8 OP_F2IP(u0, a0) OP_F2IP(u1, a1), MIXSINK)
BENCH(k_mix_8ffma_4f2ip, MIXDECL, FF8 OP_F2IP(u0, a0) OP_F2IP(u1, a1) OP_F2IP(u2, a2) OP_F2IP(u3, a3), MIXSINK)
BENCH(k_mix_8ffma_2mufu, MIXDECL, FF8 OP_RSQ(a0) OP_RSQ(a1), MIXSINK)
BENCH(k_mix_8ffma_only, MIXDECL, FF8, MIXSINK)


// u8 -> f32 conversion alternatives inside an FMA-heavy loop: per iteration 8 FFMA plus NCONV
// conversions of byte 1 of a word, each added into a live accumulator.
#define CONV_DECL float a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7, acc = 0.f; \
  unsigned u0 = threadIdx.x * 2654435761u, u1 = u0 * 3u, u2 = u0 * 5u, u3 = u0 * 7u; const unsigned kmag = 0x4B000000u | (unsigned)(clk[1] & 1)
#define CONV_SINK if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + acc + (float)(u0 ^ u1 ^ u2 ^ u3) == 12345.678f) sink[0] = a0
#define CV_MAGIC(u) { unsigned m; asm volatile("prmt.b32 %0, %1, %2, 0x7651;" : "=r"(m) : "r"(u), "r"(kmag)); acc += __uint_as_float(m) - 8388608.0f; u += 0x01010101u; }
#define CV_I2F(u) { float f; asm volatile("{ .reg .b8 x0,x1,x2,x3; mov.b32 {x0,x1,x2,x3}, %1; cvt.rn.f32.u8 %0, x1; }" : "=f"(f) : "r"(u)); acc += f; u += 0x01010101u; }
BENCH(k_conv_magic_x2, CONV_DECL, FF8 CV_MAGIC(u0) CV_MAGIC(u1), CONV_SINK)
BENCH(k_conv_i2f_x2, CONV_DECL, FF8 CV_I2F(u0) CV_I2F(u1), CONV_SINK)
BENCH(k_conv_magic_x4, CONV_DECL, FF8 CV_MAGIC(u0) CV_MAGIC(u1) CV_MAGIC(u2) CV_MAGIC(u3), CONV_SINK)
BENCH(k_conv_i2f_x4, CONV_DECL, FF8 CV_I2F(u0) CV_I2F(u1) CV_I2F(u2) CV_I2F(u3), CONV_SINK)
BENCH(k_conv_mixed_3m1i, CONV_DECL, FF8 CV_MAGIC(u0) CV_MAGIC(u1) CV_MAGIC(u2) CV_I2F(u3), CONV_SINK)

#define OP_DP2A(x) asm volatile("dp2a.lo.u32.u32 %0, %1, %0, %2;" : "+r"(x) : "r"(kb | 0x024B012Bu), "r"(kc));
BENCH(k_dp2a, U8, R8(OP_DP2A), SINKU)
#define OP_DP4A(x) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(kb | 0x01020304u), "r"(kc));
BENCH(k_dp4a, U8, R8(OP_DP4A), SINKU)
#define OP_DP(u) asm volatile("dp2a.lo.u32.u32 %0, %1, %0, %2;" : "+r"(u) : "r"(kc | 0x024B012Bu), "r"(kc));
#define DP8 OP_DP(u0) OP_DP(u1) OP_DP(u2) OP_DP(u3) OP_DP(u4) OP_DP(u5) OP_DP(u6) OP_DP(u7)
#define DP4 OP_DP(u0) OP_DP(u1) OP_DP(u2) OP_DP(u3)
BENCH(k_mix_8ffma_8dp2a, MIXDECL, FF8 DP8, MIXSINK)
BENCH(k_mix_8ffma_4dp2a, MIXDECL, FF8 DP4, MIXSINK)
BENCH(k_mix_4dp2a_4prmt, MIXDECL, DP4 OP_PR(u4) OP_PR(u5) OP_PR(u6) OP_PR(u7), MIXSINK)
struct T { const char* name; void (*k)(long long*, float*); int ops; };

int main() {
  int dev = 0, sms = 0, clk_khz = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev);
  long long* clk; float* sink;
  cudaMalloc(&clk, sms * sizeof(long long)); cudaMalloc(&sink, 4);
  T tests[] = {
    {"ffma", k_ffma, 8}, {"fadd.rm", k_fadd_rm, 8}, {"fmnmx", k_fmnmx, 8}, {"clamp(max+min)", k_clamp2, 8},
    {"mufu.rsq", k_mufu_rsq, 8}, {"mufu.rcp", k_mufu_rcp, 8}, {"lop (baseline for the +lop rows)", k_lop_only, 8},
    {"f2i.trunc + lop", k_f2i_trunc_plus_lop, 8}, {"f2i.sat.u8 (+cvt) + lop", k_f2i_sat_u8_plus_lop, 8},
    {"i2f.u32 + lop", k_i2f_plus_lop, 8}, {"prmt only", k_prmt_only, 8}, {"prmt + fadd (magic u8->f32)", k_prmt_fadd_magic, 8},
    {"cvt.pack.sat.u8.s32", k_cvt_pack_sat_u8, 8}, {"dfma", k_dfma, 8}, {"f2f f32->f64->f32 pair", k_f2f_up_and_down, 8},
    {"fdiv.rn(x,255)+fadd", k_fdiv_rn_by_255_plus_fadd, 8},
    {"lop3 (x&a)|b", k_lop3, 8}, {"shf.r funnel", k_shf_funnel, 8}, {"iadd", k_iadd, 8}, {"imad", k_imad, 8},
    {"shr only (baseline)", k_shr_only, 8}, {"i2fp.f32.u32 + shr", k_i2fp_u32_plus_shr, 8}, {"hadd2", k_hadd2, 8},
    {"cvt.f32.f16 + lop", k_cvt_f32_f16_plus_lop, 8},
    {"ffma2 (packed f32x2; 2 FMA each)", k_ffma2, 8}, {"fadd2.rd (packed)", k_fadd2_rd, 8},
    {"mix: 8 ffma + 8 prmt  (per iter = 1 seq)", k_mix_8ffma_8prmt, 1}, {"mix: 8 ffma + 4 prmt", k_mix_8ffma_4prmt, 1},
    {"mix: 8 ffma only", k_mix_8ffma_only, 1}, {"mix: 8 ffma + 2 f2ip.sat.u8", k_mix_8ffma_2f2ip, 1},
    {"mix: 8 ffma + 4 f2ip.sat.u8", k_mix_8ffma_4f2ip, 1}, {"mix: 8 ffma + 2 mufu.rsq", k_mix_8ffma_2mufu, 1},
    {"conv: 8 ffma + 2 x (prmt+fadd magic, +fadd acc, +iadd)", k_conv_magic_x2, 1},
    {"conv: 8 ffma + 2 x (i2f.u8 byte1, +fadd acc, +iadd)", k_conv_i2f_x2, 1},
    {"conv: 8 ffma + 4 x magic", k_conv_magic_x4, 1}, {"conv: 8 ffma + 4 x i2f.u8", k_conv_i2f_x4, 1},
    {"conv: 8 ffma + 3 x magic + 1 x i2f.u8", k_conv_mixed_3m1i, 1},
    {"dp2a.lo.u32.u32", k_dp2a, 8}, {"dp4a.u32.u32", k_dp4a, 8},
    {"mix: 8 ffma + 8 dp2a", k_mix_8ffma_8dp2a, 1}, {"mix: 8 ffma + 4 dp2a", k_mix_8ffma_4dp2a, 1},
    {"mix: 4 dp2a + 4 prmt", k_mix_4dp2a_4prmt, 1},
    {"mix: 4 ffma2 + 8 prmt (same flops)", k_mix_4ffma2_8prmt, 1}, {"mix: 4 ffma2 + 4 prmt", k_mix_4ffma2_4prmt, 1},
  };
  printf("device SMs=%d clockRate=%d kHz; rows are iterations of the listed sequence per clk per SM (1024 thr/SM)\n", sms, clk_khz);
  for (auto& t : tests) {
    t.k<<<sms, 1024>>>(clk, sink);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    t.k<<<sms, 1024>>>(clk, sink);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    long long h[256]; cudaMemcpy(h, clk, sms * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms; ++i) avg += (double)h[i]; avg /= sms;
    double per_clk = 1024.0 * ITER * t.ops / avg;
    printf("%-36s %8.2f seq/clk/SM   (%.0f clk, %.3f ms, eff clk %.0f MHz) %s\n", t.name, per_clk, avg, ms,
           avg / (ms * 1e-3) * 1e-6, err == cudaSuccess ? "" : cudaGetErrorString(err));
  }
  return 0;
}
