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
#define OP_RCP(x) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x));
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
