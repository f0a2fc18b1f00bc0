// Instruction-throughput microbenchmark for the LK kernel design (sm_100a).
// Measures warp-instructions per clock per SM for the instruction kinds the LK inner loop is built from, with
// 32 warps resident per SM and 8 independent dependency chains per thread.  Build: see tools/Makefile.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <string>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

constexpr int ITERS = 2048;
constexpr int CHAINS = 8;

#define BODY_KERNEL(NAME, DECL, STMT)                                                        \
    __global__ void __launch_bounds__(1024) NAME(unsigned* out, long long* cyc, unsigned seed) \
    {                                                                                        \
        __shared__ unsigned sm[4096];                                                        \
        for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * seed;               \
        __syncthreads();                                                                     \
        unsigned a[CHAINS];                                                                  \
        for (int k = 0; k < CHAINS; k++) a[k] = threadIdx.x * 7 + k + seed;                  \
        unsigned b = seed | 1, c = seed * 3 + threadIdx.x;                                   \
        const unsigned smb = (unsigned)__cvta_generic_to_shared(sm);                         \
        DECL;                                                                                \
        long long t0 = clock64();                                                            \
        for (int it = 0; it < ITERS; it++) {                                                 \
            _Pragma("unroll") for (int k = 0; k < CHAINS; k++) { STMT; }                     \
        }                                                                                    \
        long long t1 = clock64();                                                            \
        unsigned r = 0;                                                                      \
        for (int k = 0; k < CHAINS; k++) r ^= a[k];                                          \
        out[blockIdx.x * blockDim.x + threadIdx.x] = r + b + c;                              \
        if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;                                     \
    }

#define BODY_KERNEL_POST(NAME, DECL, STMT, POST)                                                        \
    __global__ void __launch_bounds__(1024) NAME(unsigned* out, long long* cyc, unsigned seed) \
    {                                                                                        \
        __shared__ unsigned sm[4096];                                                        \
        for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * seed;               \
        __syncthreads();                                                                     \
        unsigned a[CHAINS];                                                                  \
        for (int k = 0; k < CHAINS; k++) a[k] = threadIdx.x * 7 + k + seed;                  \
        unsigned b = seed | 1, c = seed * 3 + threadIdx.x;                                   \
        const unsigned smb = (unsigned)__cvta_generic_to_shared(sm);                         \
        DECL;                                                                                \
        long long t0 = clock64();                                                            \
        for (int it = 0; it < ITERS; it++) {                                                 \
            _Pragma("unroll") for (int k = 0; k < CHAINS; k++) { STMT; }                     \
        }                                                                                    \
        long long t1 = clock64();                                                            \
        POST;                                                                                \
        unsigned r = 0;                                                                      \
        for (int k = 0; k < CHAINS; k++) r ^= a[k];                                          \
        out[blockIdx.x * blockDim.x + threadIdx.x] = r + b + c;                              \
        if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;                                     \
    }

BODY_KERNEL(k_imad, , asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(b), "r"(c)))
BODY_KERNEL(k_iadd3, , asm volatile("add.s32 %0, %0, %1;" : "+r"(a[k]) : "r"(b)))
BODY_KERNEL(k_lop3, , asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[k]) : "r"(b), "r"(c)))
BODY_KERNEL(k_shf, , asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(b), "r"(c)))
BODY_KERNEL(k_shr, , asm volatile("shr.s32 %0, %0, 9;" : "+r"(a[k])))
BODY_KERNEL(k_prmt, , asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(b), "r"(c)))
BODY_KERNEL(k_dp2a, , asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(a[k]) : "r"(b), "r"(c)))
BODY_KERNEL(k_dp4a, , asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(a[k]) : "r"(b), "r"(c)))
BODY_KERNEL(k_ffma, float fb = __uint_as_float(0x3f800001u); float fc = 1e-9f, asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(a[k]) : "f"(fb), "f"(fc)))
BODY_KERNEL(k_i2f, , asm volatile("cvt.rn.f32.s32 %0, %0;" : "+r"(a[k])))
BODY_KERNEL(k_f2i, , asm volatile("cvt.rni.s32.f32 %0, %0;" : "+r"(a[k])))
BODY_KERNEL(k_redux, , asm volatile("redux.sync.add.s32 %0, %0, 0xffffffff;" : "+r"(a[k])))
BODY_KERNEL(k_shfl, , asm volatile("shfl.sync.bfly.b32 %0, %0, 1, 0x1f, 0xffffffff;" : "+r"(a[k])))
BODY_KERNEL(k_lds32, , asm volatile("{ .reg .u32 t; and.b32 t, %0, 0x3ffc; add.u32 t, t, %1; ld.shared.u32 %0, [t]; }" : "+r"(a[k]) : "r"(smb)))
BODY_KERNEL(k_lds32_lin, unsigned base = smb + (threadIdx.x & 31) * 4, asm volatile("{ .reg .u32 t, u; add.u32 t, %1, %2; ld.shared.u32 u, [t]; xor.b32 %0, %0, u; }" : "+r"(a[k]) : "r"(base), "r"((unsigned)(k * 128))))
BODY_KERNEL(k_lds8, unsigned base = smb + (threadIdx.x & 31) * 4, asm volatile("{ .reg .u32 t, u; add.u32 t, %1, %2; ld.shared.u8 u, [t]; xor.b32 %0, %0, u; }" : "+r"(a[k]) : "r"(base), "r"((unsigned)(k * 128 + 1))))
BODY_KERNEL(k_lds64_lin, unsigned base = smb + (threadIdx.x & 31) * 8, asm volatile("{ .reg .u32 t, u, v; add.u32 t, %1, %2; ld.shared.v2.u32 {u, v}, [t]; xor.b32 %0, %0, u; xor.b32 %0, %0, v; }" : "+r"(a[k]) : "r"(base), "r"((unsigned)(k * 256))))
BODY_KERNEL(k_dfma, double da = 1.0000001; double dacc = seed, asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(dacc) : "d"(da)); a[k] += (unsigned)__double2loint(dacc))
// mixes representative of the inner loop
BODY_KERNEL(k_mix_imad_shf, , asm volatile("mad.lo.s32 %0, %0, %1, %2; shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(b), "r"(c)))
BODY_KERNEL(k_mix_dp2a_shr_imad, , asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %0; dp2a.hi.u32.u32 %0, %1, %2, %0; shr.s32 %0, %0, 9; mad.lo.s32 %0, %0, %1, %2; mad.lo.s32 %0, %0, %2, %1;" : "+r"(a[k]) : "r"(b), "r"(c)))
BODY_KERNEL(k_mix_4imad_shr_2imad, , asm volatile("mad.lo.s32 %0, %0, %1, %2; mad.lo.s32 %0, %0, %1, %2; mad.lo.s32 %0, %0, %1, %2; mad.lo.s32 %0, %0, %1, %2; shr.s32 %0, %0, 9; mad.lo.s32 %0, %0, %1, %2; mad.lo.s32 %0, %0, %2, %1;" : "+r"(a[k]) : "r"(b), "r"(c)))
BODY_KERNEL(k_mix_ffma4_imad2, float fb = __uint_as_float(0x3f800001u); float fc = 1e-9f, asm volatile("fma.rn.f32 %0, %0, %1, %2; fma.rn.f32 %0, %0, %1, %2; fma.rn.f32 %0, %0, %1, %2; fma.rn.f32 %0, %0, %1, %2;" : "+r"(a[k]) : "f"(fb), "f"(fc)); asm volatile("mad.lo.s32 %0, %0, %1, %2; mad.lo.s32 %0, %0, %2, %1;" : "+r"(a[k]) : "r"(b), "r"(c)))

BODY_KERNEL(k_mix_imad_ffma, float fb = __uint_as_float(0x3f800001u); float fc = 1e-9f, asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(b), "r"(c)); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(a[(k + 1) & 7]) : "f"(fb), "f"(fc)))
BODY_KERNEL(k_mix_imad_ffma2, float fb = __uint_as_float(0x3f800001u); float fc = 1e-9f, asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(b), "r"(c)); asm volatile("fma.rn.f32 %0, %0, %1, %2; fma.rn.f32 %3, %3, %1, %2;" : "+r"(a[(k + 1) & 7]), "+r"(a[(k + 3) & 7]) : "f"(fb), "f"(fc)))
BODY_KERNEL(k_mix_idp_prmt, , asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(a[k]) : "r"(b), "r"(c)); asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a[(k + 1) & 7]) : "r"(b), "r"(c)))
BODY_KERNEL(k_mix_idp_prmt_iadd, , asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(a[k]) : "r"(b), "r"(c)); asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a[(k + 1) & 7]) : "r"(b), "r"(c)); asm volatile("add.s32 %0, %0, %1;" : "+r"(a[(k + 2) & 7]) : "r"(b)))
BODY_KERNEL(k_f2i_floor, , asm volatile("cvt.rmi.s32.f32 %0, %0;" : "+r"(a[k])))
BODY_KERNEL(k_i2f64, , asm volatile("{ .reg .f64 d; cvt.rn.f64.s32 d, %0; cvt.rn.f32.f64 %0, d; }" : "+r"(a[k])))
BODY_KERNEL(k_fadd, float fc = 1e-9f, asm volatile("add.rn.f32 %0, %0, %1;" : "+r"(a[k]) : "f"(fc)))
BODY_KERNEL(k_isetp_sel, , asm volatile("{ .reg .pred p; setp.gt.s32 p, %0, %1; selp.s32 %0, %1, %2, p; }" : "+r"(a[k]) : "r"(b), "r"(c)))

BODY_KERNEL_POST(k_fadd2, unsigned long long fa2[CHAINS]; for (int k = 0; k < CHAINS; k++) fa2[k] = ((unsigned long long)a[k] << 32) | 0x3f800000u; unsigned long long fc2 = 0x3089705f3089705full, asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(fa2[k]) : "l"(fc2)), for (int k = 0; k < CHAINS; k++) a[k] ^= (unsigned)(fa2[k] >> 32) ^ (unsigned)fa2[k])
BODY_KERNEL_POST(k_ffma2, unsigned long long fa2[CHAINS]; for (int k = 0; k < CHAINS; k++) fa2[k] = ((unsigned long long)a[k] << 32) | 0x3f800000u; unsigned long long fc2 = 0x3089705f3089705full; unsigned long long fb2 = 0x3f8000013f800001ull, asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(fa2[k]) : "l"(fb2), "l"(fc2)), for (int k = 0; k < CHAINS; k++) a[k] ^= (unsigned)(fa2[k] >> 32) ^ (unsigned)fa2[k])
BODY_KERNEL(k_fadd_rm, float fc = 12582912.f, asm volatile("add.rm.f32 %0, %0, %1;" : "+r"(a[k]) : "f"(fc)))
BODY_KERNEL_POST(k_imad_wide, unsigned long long wa[CHAINS]; for (int k = 0; k < CHAINS; k++) wa[k] = a[k], asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(wa[k]) : "r"((int)wa[k]), "r"(c)), for (int k = 0; k < CHAINS; k++) a[k] ^= (unsigned)(wa[k] >> 32) ^ (unsigned)wa[k])
BODY_KERNEL(k_lea, , asm volatile("{ .reg .u32 t; shl.b32 t, %0, 7; add.u32 %0, t, %1; }" : "+r"(a[k]) : "r"(c)))
BODY_KERNEL(k_mix_idp_lop, , asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(a[k]) : "r"(b), "r"(c)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[(k + 1) & 7]) : "r"(b), "r"(c)))
BODY_KERNEL(k_mix_idp_iadd, , asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(a[k]) : "r"(b), "r"(c)); asm volatile("add.s32 %0, %0, %1;" : "+r"(a[(k + 1) & 7]) : "r"(b)))
BODY_KERNEL(k_mix_idp_fadd, float fc = 1e-9f, asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(a[k]) : "r"(b), "r"(c)); asm volatile("add.rn.f32 %0, %0, %1;" : "+r"(a[(k + 1) & 7]) : "f"(fc)))
BODY_KERNEL(k_mix_idp_2fadd, float fc = 1e-9f, asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(a[k]) : "r"(b), "r"(c)); asm volatile("add.rn.f32 %0, %0, %2; add.rn.f32 %1, %1, %2;" : "+r"(a[(k + 1) & 7]), "+r"(a[(k + 3) & 7]) : "f"(fc)))
BODY_KERNEL(k_mix_lop_iadd, , asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[k]) : "r"(b), "r"(c)); asm volatile("add.s32 %0, %0, %1;" : "+r"(a[(k + 1) & 7]) : "r"(b)))
BODY_KERNEL(k_sgxt, , asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; cvt.s32.s16 %0, t; }" : "+r"(a[k]) : "r"(b)))
BODY_KERNEL(k_lds64_x, unsigned base = smb + (threadIdx.x & 31) * 8, asm volatile("{ .reg .u32 t, u, v; add.u32 t, %1, %2; ld.shared.v2.u32 {u, v}, [t]; xor.b32 %0, %0, u; xor.b32 %0, %0, v; }" : "+r"(a[k]) : "r"(base), "r"((unsigned)(k * 256))))

struct Entry { const char* name; void (*fn)(unsigned*, long long*, unsigned); int instr_per_stmt; };

int main()
{
    cudaDeviceProp prop;
    CHECK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("device %s, %d SMs, max clock %d MHz\n", prop.name, sms, clk_khz / 1000);
    unsigned* out; long long* cyc;
    CHECK(cudaMalloc(&out, sizeof(unsigned) * sms * 2 * 1024));
    CHECK(cudaMalloc(&cyc, sizeof(long long) * sms * 2));
    Entry tests[] = {
        {"imad", k_imad, 1}, {"iadd", k_iadd3, 1}, {"lop3", k_lop3, 1}, {"shf", k_shf, 1}, {"shr", k_shr, 1}, {"prmt", k_prmt, 1},
        {"dp2a", k_dp2a, 1}, {"dp4a", k_dp4a, 1}, {"ffma", k_ffma, 1}, {"i2f", k_i2f, 1}, {"f2i", k_f2i, 1},
        {"redux", k_redux, 1}, {"shfl", k_shfl, 1}, {"lds32_dep", k_lds32, 3}, {"lds32_lin(+2alu)", k_lds32_lin, 3},
        {"lds8_lin(+2alu)", k_lds8, 3}, {"lds64_lin(+3alu)", k_lds64_lin, 4}, {"dfma(+2)", k_dfma, 3},
        {"mix imad+ffma", k_mix_imad_ffma, 2}, {"mix imad+2ffma", k_mix_imad_ffma2, 3}, {"mix idp+prmt", k_mix_idp_prmt, 2},
        {"mix idp+prmt+iadd", k_mix_idp_prmt_iadd, 3}, {"f2i floor", k_f2i_floor, 1}, {"i2f64+f2f", k_i2f64, 2}, {"fadd", k_fadd, 1},
        {"isetp+sel", k_isetp_sel, 2},
        {"mix imad+shf", k_mix_imad_shf, 2}, {"mix 2dp2a+shr+2imad", k_mix_dp2a_shr_imad, 5},
        {"mix 4imad+shr+2imad", k_mix_4imad_shr_2imad, 7}, {"mix 4ffma+2imad", k_mix_ffma4_imad2, 6},
        {"fadd2", k_fadd2, 1}, {"ffma2", k_ffma2, 1}, {"fadd.rm", k_fadd_rm, 1}, {"imad.wide", k_imad_wide, 1},
        {"shl+add (lea?)", k_lea, 1}, {"mix idp+lop3", k_mix_idp_lop, 2}, {"mix idp+iadd", k_mix_idp_iadd, 2},
        {"mix idp+fadd", k_mix_idp_fadd, 2}, {"mix idp+2fadd", k_mix_idp_2fadd, 3}, {"mix lop3+iadd", k_mix_lop_iadd, 2}, {"add+cvt.s32.s16", k_sgxt, 2},
    };
    for (auto& t : tests) {
        for (int warps : {32, 16}) {
            const int threads = warps * 32;
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            t.fn<<<sms, threads>>>(out, cyc, 12345u);  // warm-up
            CHECK(cudaDeviceSynchronize());
            cudaEventRecord(e0);
            t.fn<<<sms, threads>>>(out, cyc, 12345u);
            cudaEventRecord(e1);
            CHECK(cudaDeviceSynchronize());
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            std::vector<long long> h(sms);
            CHECK(cudaMemcpy(h.data(), cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
            double avg = 0;
            for (auto v : h) avg += (double)v;
            avg /= sms;
            const double winstr = (double)warps * ITERS * CHAINS * t.instr_per_stmt;
            printf("%-24s warps/SM %2d  cycles %9.0f  warp-instr/clk/SM %6.3f  (%.3f ms, eff clock %.0f MHz)\n", t.name, warps, avg,
                   winstr / avg, ms, avg / (ms * 1e3));
        }
    }
    return 0;
}
