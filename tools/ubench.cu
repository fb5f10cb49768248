// ubench.cu — issue-rate microbenchmarks for the instruction mix of the splat kernel (sm_100a).
// For each op: 8 independent dependency chains per thread, 1024 threads per SM, all 148 SMs;
// prints warp-instructions per clock per SM.  Used to decide which pipe (fma / alu) each op of
// the gather loop occupies and whether packed f32x2 ops issue at full rate.  Not part of the product.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define ITER 4096
#define CH 8

template <int OP>
__global__ void __launch_bounds__(1024) k(float *out, float a, float b, int ia, int ib) {
    float f[CH], g[CH]; u64 p[CH]; int n[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { f[i] = a + i + threadIdx.x; g[i] = b + i; n[i] = ia + i + threadIdx.x; p[i] = ((u64)__float_as_uint(f[i]) << 32) | __float_as_uint(b + i); }
    u64 pb = ((u64)__float_as_uint(b) << 32) | __float_as_uint(a);
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if (OP == 0) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(b));
            if (OP == 1) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(b));
            if (OP == 2) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(b), "f"(a));
            if (OP == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb));
            if (OP == 4) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb));
            if (OP == 5) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(pb));
            if (OP == 6) asm volatile("add.s32 %0, %0, %1;" : "+r"(n[i]) : "r"(ib));
            if (OP == 7) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(n[i]) : "r"(ib), "r"(ia));
            if (OP == 8) asm volatile("prmt.b32 %0, %0, %1, 0x4441;" : "+r"(n[i]) : "r"(ib));
            if (OP == 9) asm volatile("shf.l.wrap.b32 %0, %0, %1, 3;" : "+r"(n[i]) : "r"(ib));
            if (OP == 10) asm volatile("min.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(b));
            if (OP == 11) asm volatile("add.rm.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(b));
            if (OP == 12) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(n[i]) : "r"(ib), "r"(ia));
            if (OP == 13) asm volatile("{ .reg .pred q; setp.ne.s32 q, %0, %1; selp.s32 %0, %0, %1, q; }" : "+r"(n[i]) : "r"(ib));
            if (OP == 14) { asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(b)); asm volatile("add.s32 %0, %0, %1;" : "+r"(n[i]) : "r"(ib)); }   // fma + alu mix
            if (OP == 15) { asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(b)); asm volatile("add.s32 %0, %0, %1;" : "+r"(n[i]) : "r"(ib)); }   // fadd + iadd mix
            if (OP == 16) { asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb)); asm volatile("add.s32 %0, %0, %1;" : "+r"(n[i]) : "r"(ib)); } // fadd2 + iadd
            if (OP == 17) { asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb)); asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(b)); } // fadd2 + fmul
            if (OP == 18) asm volatile("shl.b32 %0, %0, 2; add.s32 %0, %0, %1;" : "+r"(n[i]) : "r"(ib));   // lea-like
            if (OP == 20) { asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(b)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(n[i]) : "r"(ib), "r"(ia)); }   // fma pipe + alu pipe
            if (OP == 21) { asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(n[i]) : "r"(ib), "r"(ia)); }
            if (OP == 22) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(n[i]) : "r"(ib), "r"(ia));
            if (OP == 23) { asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(b)); asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(n[i]) : "r"(ib), "r"(ia)); }
            if (OP == 24) { asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(b)); asm volatile("prmt.b32 %0, %0, %1, 0x4441;" : "+r"(n[i]) : "r"(ib)); }
            if (OP == 25) { asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(b)); asm volatile("min.f32 %0, %0, %1;" : "+f"(f[(i + 4) % CH]) : "f"(a)); }   // fmul + fmnmx
            if (OP == 26) { asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(b)); asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(a)); asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(b));
                            asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb)); asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb)); }   // the tap's FP part: 3 FMUL + 2 FADD2
            if (OP == 27) { asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(b)); asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(a)); asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(b));
                            asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb)); asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb));
                            asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(n[i]) : "r"(ib), "r"(ia)); }   // + the address IDP
            if (OP == 28 || OP == 29) { asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(b)); asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(a)); asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(b));
                            asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(g[i]) : "f"(b)); asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(g[i]) : "f"(a));
                            asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(g[i]) : "f"(b)); asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(g[i]) : "f"(a));
                            if (OP == 29) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(n[i]) : "r"(ib), "r"(ia)); }   // scalar tap: 3 FMUL + 4 FADD (+ IDP)
            if (OP == 30) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(b), "f"(a)); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(g[i]) : "f"(b), "f"(a));
                            asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(a), "f"(b)); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(g[i]) : "f"(a), "f"(b));
                            asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(n[i]) : "r"(ib), "r"(ia)); }   // fma-mode tap, scalar: 4 FFMA + IDP
            if (OP == 31) { asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(pb)); asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(pb));
                            asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(n[i]) : "r"(ib), "r"(ia)); }   // fma-mode tap, packed: 2 FFMA2 + IDP
            if (OP == 19) { asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(pb)); asm volatile("add.s32 %0, %0, %1;" : "+r"(n[i]) : "r"(ib)); }
        }
    }
    float s = 0; int t = 0; u64 q = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) { s += f[i] + g[i]; t += n[i]; q ^= p[i]; }
    if ((int)threadIdx.x == ib * 1000) out[0] = s + (float)t + (float)q;  // ib is a runtime value: never true, never provably so
}

template <int OP> void run(const char *name, int per_iter) {
    float *d; cudaMalloc(&d, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<148, 1024>>>(d, 1.0001f, 0.9999f, 3, 5);
    cudaEventRecord(e0);
    k<OP><<<148, 1024>>>(d, 1.0001f, 0.9999f, 3, 5);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double winst = 32.0 * ITER * CH * per_iter;  // warp-instructions per SM (32 warps)
    double cycles = ms * 1e-3 * clk * 1e3;
    printf("%-28s %.3f ms  %.2f warp-inst/clk/SM (at %d MHz nominal)\n", name, ms, winst / cycles, clk / 1000);
    cudaFree(d);
}

int main() {
    run<0>("FADD", 1); run<1>("FMUL", 1); run<2>("FFMA", 1); run<3>("FADD2", 1); run<4>("FMUL2", 1); run<5>("FFMA2", 1);
    run<6>("IADD", 1); run<7>("IMAD", 1); run<8>("PRMT", 1); run<9>("SHF", 1); run<10>("FMNMX", 1); run<11>("FADD.RM", 1);
    run<12>("LOP3", 1); run<13>("ISETP+SEL", 2); run<14>("FMUL+IADD", 2); run<15>("FADD+IADD", 2); run<16>("FADD2+IADD", 2);
    run<17>("FADD2+FMUL", 2); run<18>("SHL+IADD (LEA)", 1); run<19>("FFMA2+IADD", 2);
    run<20>("FMUL+LOP3", 2); run<21>("FADD2+LOP3", 2); run<22>("IDP.4A", 1); run<23>("FMUL+IDP", 2); run<24>("FMUL+PRMT", 2);
    run<25>("FMUL+FMNMX", 2); run<26>("3 FMUL + 2 FADD2", 5); run<27>("3 FMUL + 2 FADD2 + IDP", 6);
    run<28>("3 FMUL + 4 FADD", 7); run<29>("3 FMUL + 4 FADD + IDP", 8); run<30>("4 FFMA + IDP", 5); run<31>("2 FFMA2 + IDP", 3);
    return 0;
}
