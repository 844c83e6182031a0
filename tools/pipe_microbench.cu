// pipe_microbench.cu -- issue cost of the instructions the coder is made of, on the GPU it runs on.
// Development aid (not part of the product):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pipe_microbench tools/pipe_microbench.cu && /tmp/pipe_microbench
// Every test runs ITER iterations of 8 independent chains of one instruction in every warp of a
// full-occupancy launch (one 1024-thread CTA per SM = 8 warps per scheduler) and reports
// scheduler cycles per warp-instruction (1.0 = one per clock per SM sub-partition), plus the
// dependent-issue latency from a single warp running a single chain.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define ITER 4096

#define CHAINS8(OP)          \
    OP(0) OP(1) OP(2) OP(3) OP(4) OP(5) OP(6) OP(7)

template <int KIND>
__global__ void __launch_bounds__(1024) bench(double* out, long long* cycles, int chains) {
    double d[8];
    float f[8];
    unsigned u[8];
    for (int i = 0; i < 8; ++i) {
        d[i] = 1.0 + 1e-3 * (threadIdx.x + i);
        f[i] = 1.0f + 1e-3f * (threadIdx.x + i);
        u[i] = threadIdx.x * 7 + i;
    }
    const double c1 = 0.999999, c2 = 1e-7;
    const float g1 = 0.999999f, g2 = 1e-7f;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
        if (chains == 8) {
#define DFMA(i) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(c1), "d"(c2));
#define DADD(i) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(c2));
#define DADDRZ(i) asm volatile("add.rz.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(c2));
#define DMUL(i) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(c1));
#define F2F64(i) asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d[i]) : "f"(f[i]));
#define F2F32(i) asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f[i]) : "d"(d[i]));
#define D2I(i) asm volatile("cvt.rzi.s32.f64 %0, %1;" : "=r"(u[i]) : "d"(d[i]));
#define F2I(i) asm volatile("cvt.rzi.s32.f32 %0, %1;" : "=r"(u[i]) : "f"(f[i]));
#define I2F(i) asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(f[i]) : "r"(u[i]));
#define I2D(i) asm volatile("cvt.rn.f64.s32 %0, %1;" : "=d"(d[i]) : "r"(u[i]));
#define RCP64(i) asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(d[i]));
#define RCP32(i) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
#define LG2(i) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
#define FFMA(i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(g1), "f"(g2));
#define IMAD(i) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(u[(i + 1) & 7] | 1u), "r"(12345u));
#define IADD(i) asm volatile("add.u32 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]));
#define LOP(i) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]), "r"(0x5a5a5a5au));
#define SHF(i) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]));
#define FMNMX(i) asm volatile("min.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(g1));
#define IMNMX(i) asm volatile("min.u32 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]));
#define FADD(i) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(g2));
#define SEL(i) asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %0, %1, p;}" : "+r"(u[i]) : "r"(u[(i + 1) & 7]));
            if (KIND == 0) { CHAINS8(DFMA) }
            if (KIND == 1) { CHAINS8(DADD) }
            if (KIND == 2) { CHAINS8(DMUL) }
            if (KIND == 3) { CHAINS8(F2F64) }
            if (KIND == 4) { CHAINS8(F2F32) }
            if (KIND == 5) { CHAINS8(D2I) }
            if (KIND == 6) { CHAINS8(F2I) }
            if (KIND == 7) { CHAINS8(I2F) }
            if (KIND == 8) { CHAINS8(I2D) }
            if (KIND == 9) { CHAINS8(RCP64) }
            if (KIND == 10) { CHAINS8(RCP32) }
            if (KIND == 11) { CHAINS8(LG2) }
            if (KIND == 12) { CHAINS8(FFMA) }
            if (KIND == 13) { CHAINS8(IMAD) }
            if (KIND == 14) { CHAINS8(IADD) }
            if (KIND == 15) { CHAINS8(LOP) }
            if (KIND == 16) { CHAINS8(SHF) }
            if (KIND == 17) { CHAINS8(FMNMX) }
            if (KIND == 18) { CHAINS8(IMNMX) }
            if (KIND == 19) { CHAINS8(FADD) }
            if (KIND == 20) { CHAINS8(DADDRZ) }
            if (KIND == 21) { CHAINS8(SEL) }
            // mixes: does FP64 issue overlap with other pipes?
            if (KIND == 22) { CHAINS8(DFMA) CHAINS8(IADD) CHAINS8(LOP) }
            if (KIND == 23) { CHAINS8(DFMA) CHAINS8(FFMA) CHAINS8(IMAD) }
            if (KIND == 24) { CHAINS8(DFMA) CHAINS8(DFMA) CHAINS8(DFMA) CHAINS8(DFMA) CHAINS8(F2F32) }
            if (KIND == 25) { CHAINS8(DFMA) CHAINS8(DADD) }
            if (KIND == 26) { CHAINS8(DFMA) CHAINS8(FFMA) }
            if (KIND == 27) { CHAINS8(DFMA) CHAINS8(IMAD) }
            if (KIND == 28) { CHAINS8(DFMA) CHAINS8(IADD) }
            if (KIND == 29) { CHAINS8(DFMA) CHAINS8(LOP) }
            if (KIND == 30) { CHAINS8(DFMA) CHAINS8(FFMA) CHAINS8(FFMA) CHAINS8(FFMA) }
            if (KIND == 31) { CHAINS8(DFMA) CHAINS8(LOP) CHAINS8(FFMA) CHAINS8(F2F64) }
            if (KIND == 32) { CHAINS8(FFMA) CHAINS8(LOP) }
            if (KIND == 33) { CHAINS8(FFMA) CHAINS8(IMAD) }
            if (KIND == 34) { CHAINS8(DFMA) CHAINS8(FMNMX) CHAINS8(SHF) }
        } else {
            if (KIND == 0) { DFMA(0) DFMA(0) DFMA(0) DFMA(0) DFMA(0) DFMA(0) DFMA(0) DFMA(0) }
            if (KIND == 1) { DADD(0) DADD(0) DADD(0) DADD(0) DADD(0) DADD(0) DADD(0) DADD(0) }
            if (KIND == 2) { DMUL(0) DMUL(0) DMUL(0) DMUL(0) DMUL(0) DMUL(0) DMUL(0) DMUL(0) }
            if (KIND == 9) { RCP64(0) RCP64(0) RCP64(0) RCP64(0) RCP64(0) RCP64(0) RCP64(0) RCP64(0) }
            if (KIND == 12) { FFMA(0) FFMA(0) FFMA(0) FFMA(0) FFMA(0) FFMA(0) FFMA(0) FFMA(0) }
            if (KIND == 14) { IADD(0) IADD(0) IADD(0) IADD(0) IADD(0) IADD(0) IADD(0) IADD(0) }
            if (KIND == 13) { IMAD(0) IMAD(0) IMAD(0) IMAD(0) IMAD(0) IMAD(0) IMAD(0) IMAD(0) }
        }
    }
    const long long t1 = clock64();
    double acc = 0;
    for (int i = 0; i < 8; ++i) acc += d[i] + f[i] + u[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int KIND>
void run(const char* name, int per_iter, bool latency) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out;
    long long* cyc;
    cudaMalloc(&out, sizeof(double) * sms * 1024);
    cudaMalloc(&cyc, sizeof(long long) * sms);
    long long h[1024];
    bench<KIND><<<sms, 1024>>>(out, cyc, 8);
    bench<KIND><<<sms, 1024>>>(out, cyc, 8);
    cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < sms; ++i) mean += (double)h[i];
    mean /= sms;
    // 8 warps per scheduler, each issuing per_iter instructions per iteration
    const double thr = mean / ((double)ITER * per_iter * 8);
    double lat = 0;
    if (latency) {
        bench<KIND><<<1, 32>>>(out, cyc, 1);
        cudaMemcpy(h, cyc, sizeof(long long), cudaMemcpyDeviceToHost);
        lat = (double)h[0] / ((double)ITER * 8);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("%-34s %7.3f cycles/warp-instr/scheduler", name, thr);
    if (latency) printf("   dependent latency %6.2f cycles", lat);
    printf("%s\n", e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    run<0>("DFMA", 8, true);
    run<1>("DADD", 8, true);
    run<2>("DMUL", 8, true);
    run<20>("DADD.RZ", 8, false);
    run<3>("F2F.F64.F32", 8, false);
    run<4>("F2F.F32.F64", 8, false);
    run<5>("F2I.S32.F64", 8, false);
    run<6>("F2I.S32.F32", 8, false);
    run<7>("I2F.F32.S32", 8, false);
    run<8>("I2F.F64.S32", 8, false);
    run<9>("MUFU.RCP64H", 8, true);
    run<10>("MUFU.RCP", 8, false);
    run<11>("MUFU.LG2", 8, false);
    run<12>("FFMA", 8, true);
    run<19>("FADD", 8, false);
    run<13>("IMAD", 8, true);
    run<14>("IADD3", 8, true);
    run<15>("LOP3", 8, false);
    run<16>("SHF", 8, false);
    run<17>("FMNMX", 8, false);
    run<18>("IMNMX", 8, false);
    run<21>("ISETP+SEL (2 instr)", 16, false);
    run<22>("mix DFMA+IADD3+LOP3 (per instr)", 24, false);
    run<23>("mix DFMA+FFMA+IMAD (per instr)", 24, false);
    run<24>("mix 4 DFMA + 1 F2F (per instr)", 40, false);
    run<25>("mix DFMA+DADD (per instr)", 16, false);
    run<26>("mix DFMA+FFMA (per instr)", 16, false);
    run<27>("mix DFMA+IMAD (per instr)", 16, false);
    run<28>("mix DFMA+IADD3 (per instr)", 16, false);
    run<29>("mix DFMA+LOP3 (per instr)", 16, false);
    run<30>("mix DFMA+3 FFMA (per instr)", 32, false);
    run<31>("mix DFMA+LOP3+FFMA+F2F (per instr)", 32, false);
    run<32>("mix FFMA+LOP3 (per instr)", 16, false);
    run<33>("mix FFMA+IMAD (per instr)", 16, false);
    run<34>("mix DFMA+FMNMX+SHF (per instr)", 24, false);
    return 0;
}
