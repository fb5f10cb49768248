// ubench_gather.cu — what can the gather's instruction mix reach?  The interior column visit of
// splat_window_kernel<2,128,false> (record LDS.128 + LDS.64, table column index, per row IDP.4A + LDS +
// 3 FMUL + 2 FADD2) over a static shared-memory record buffer: no pre-pass, no barriers, no votes, no
// global memory.  Run at 1..8 resident CTAs per SM (128 threads each) to separate "too few warps" from
// "this mix cannot issue faster".  Not part of the product.  Build: nvcc -O3 -std=c++17 -fmad=false
// -gencode arch=compute_100a,code=sm_100a tools/ubench_gather.cu -o ubench_gather
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;

__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 v; asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(lo), "f"(hi)); return v; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 v; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(v) : "l"(a), "l"(b)); return v; }
__device__ __forceinline__ float lds_f32(unsigned addr) { float v; asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr)); return v; }
__device__ __forceinline__ unsigned bin_bits(float v) { return (unsigned)__float_as_int(__fadd_rd(fminf(fabsf(v), 15.f), 8388608.f)); }

constexpr int TW = 128, H = 2, ROWS = 5, SPP = 16, PITCH = 17, NPX = TW + 2 * H;
constexpr int TAB_ROW_BYTES = 21 * 4, TAB_BYTES = 1536;

template <int UNROLL, int MODE>
__global__ void __launch_bounds__(TW) gather_only(float *out, int reps, float irx16, int never, const unsigned *flagp) {
    extern __shared__ __align__(16) unsigned char smem[];
    float4 *s_a = reinterpret_cast<float4 *>(smem + TAB_BYTES);
    uint2 *s_b = reinterpret_cast<uint2 *>(smem + TAB_BYTES + NPX * PITCH * 16);
    const int tid = threadIdx.x;
    for (int i = tid; i < TAB_BYTES / 4; i += TW) reinterpret_cast<float *>(smem)[i] = 1.f / (float)(1 + i);
    for (int i = tid; i < NPX * PITCH; i += TW) {
        const int px = i / PITCH, s = i % PITCH;
        s_a[i] = make_float4(0.1f * s, 0.2f, 0.3f, (float)px + 0.03f * s);  // pdx inside its pixel
        const unsigned r = (s * 5 + px) & 7;
        s_b[i] = make_uint2(r | ((r + 8u) << 8) | ((r + 1u) << 16) | ((15u - r) << 24), 16u);
    }
    __syncthreads();
    unsigned tab_base = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("mov.u32 %0, %0;" : "+r"(tab_base));
    const float fx = (float)(tid + H);
    const unsigned flags = flagp[0];
    u64 acc_rg[ROWS], acc_bw[ROWS];
    float sr[ROWS], sg[ROWS], sb[ROWS], sw[ROWS];
#pragma unroll
    for (int j = 0; j < ROWS; ++j) { acc_rg[j] = acc_bw[j] = 0ull; sr[j] = sg[j] = sb[j] = sw[j] = 0.f; }
    for (int rep = 0; rep < reps; ++rep) {
#pragma unroll 1
        for (int d = -H + 1; d <= H - 1; ++d) {
            const float4 *pa = s_a + (tid + H + d) * PITCH;
            const uint2 *pb = s_b + (tid + H + d) * PITCH;
            for (int s0 = 0; s0 < SPP; s0 += 8) {
            const bool upper = MODE == 4 ? ((flags >> (s0 >> 3)) & 1) != 0 : false;   // warp-uniform, from memory
            if (upper) {
#pragma unroll UNROLL
            for (int s = s0; s < s0 + 8; ++s) {
                const bool upper = true;
                const float4 a = pa[s];
                const uint2 yb = pb[s];
                const unsigned ifx = bin_bits((fx - a.w) * irx16) & 0xFu;
                const unsigned xcol = tab_base + ifx * 4;
#pragma unroll
                for (int j = 0; j < ROWS; ++j) {
                    if (MODE == 4 && j == (upper ? ROWS - 1 : 0)) continue;
                    const unsigned waddr = __dp4a(j < 4 ? yb.x : yb.y, (unsigned)TAB_ROW_BYTES << (8 * (j & 3)), xcol);
                    const float w = lds_f32(waddr);
                    acc_rg[j] = add2(acc_rg[j], pack2(a.x * w, a.y * w));
                    acc_bw[j] = add2(acc_bw[j], pack2(a.z * w, w));
                }
            }
            } else {
#pragma unroll UNROLL
            for (int s = s0; s < s0 + 8; ++s) {
                const float4 a = pa[s];
                const uint2 yb = pb[s];
                const unsigned ifx = bin_bits((fx - a.w) * irx16) & 0xFu;
                const unsigned xcol = tab_base + ifx * 4;
#pragma unroll
                for (int j = 0; j < ROWS; ++j) {
                    if (MODE == 4 && j == (upper ? ROWS - 1 : 0)) continue;   // the row this sample type never reaches
                    const unsigned waddr = __dp4a(j < 4 ? yb.x : yb.y, (unsigned)TAB_ROW_BYTES << (8 * (j & 3)), xcol);
                    const float w = lds_f32(waddr);
                    if (MODE == 0 || MODE == 4) {          // exact, packed adds (the kernel today)
                        acc_rg[j] = add2(acc_rg[j], pack2(a.x * w, a.y * w));
                        acc_bw[j] = add2(acc_bw[j], pack2(a.z * w, w));
                    } else if (MODE == 1) {   // exact, scalar adds
                        sr[j] += a.x * w; sg[j] += a.y * w; sb[j] += a.z * w; sw[j] += w;
                    } else if (MODE == 2) {   // fused, scalar
                        sr[j] = __fmaf_rn(a.x, w, sr[j]); sg[j] = __fmaf_rn(a.y, w, sg[j]); sb[j] = __fmaf_rn(a.z, w, sb[j]); sw[j] += w;
                    } else {                  // fused, packed
                        const u64 ww = pack2(w, w);
                        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc_rg[j]) : "l"(pack2(a.x, a.y)), "l"(ww));
                        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc_bw[j]) : "l"(pack2(a.z, 1.f)), "l"(ww));
                    }
                }
            }
            }
            }
        }
    }
    u64 q = 0;
#pragma unroll
    for (int j = 0; j < ROWS; ++j) q ^= acc_rg[j] ^ acc_bw[j] ^ (u64)__float_as_uint(sr[j] + sg[j] + sb[j] + sw[j]);
    if (tid == never * 1000 + 4096) out[0] = (float)q;
}

template <int UNROLL, int MODE>
static void run(int ctas_per_sm) {
    const size_t need = TAB_BYTES + (size_t)NPX * PITCH * 24;
    // pad shared memory so that exactly ctas_per_sm CTAs fit
    size_t smem = (size_t)(227 * 1024) / ctas_per_sm - 1024;
    if (smem < need) { printf("%d CTAs/SM do not fit\n", ctas_per_sm); return; }
    if (ctas_per_sm >= 5) smem = need;  // cannot force more than the records allow; report what fits
    cudaFuncSetAttribute(gather_only<UNROLL, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gather_only<UNROLL, MODE>, TW, smem);
    float *d; cudaMalloc(&d, 4);
    unsigned *dflags; cudaMalloc(&dflags, 4); unsigned hf = 1u; cudaMemcpy(dflags, &hf, 4, cudaMemcpyHostToDevice);
    const int reps = 400;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    gather_only<UNROLL, MODE><<<148 * per_sm, TW, smem>>>(d, reps, 8.f, 7, dflags);
    cudaEventRecord(e0);
    gather_only<UNROLL, MODE><<<148 * per_sm, TW, smem>>>(d, reps, 8.f, 7, dflags);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double visits = (double)reps * 3 * SPP;             // sample-visits per thread
    const double cycles = ms * 1e-3 * clk * 1e3;
    // per SM: per_sm CTAs x 4 warps, each `visits` sample-visits
    printf("mode %d unroll %d  CTAs/SM %d (%2d warps)  %.3f ms  %.1f clk per warp-sample-visit per SM  = %.2e sample-visits/s/GPU\n", MODE, UNROLL,
           per_sm, per_sm * 4, ms, cycles / (visits * per_sm * 4), visits * per_sm * 4 * 32 * 148 / (ms * 1e-3));
    cudaFree(d);
}

int main() {
    for (int c : {1, 2, 3, 4}) run<8, 0>(c);
    for (int c : {2, 4}) run<4, 0>(c);
    for (int c : {2, 4}) run<16, 0>(c);
    for (int c : {2, 4}) run<8, 1>(c);
    for (int c : {2, 4}) run<8, 2>(c);
    for (int c : {2, 4}) run<8, 3>(c);
    for (int c : {2, 4}) run<8, 4>(c);
    for (int c : {4}) run<4, 4>(c);
    return 0;
}
