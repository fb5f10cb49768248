// splat.cu — [T2, extension] FilmTile::add_sample for a pixel-major sample stream, fused with
// merge_film_tile.  The reference declares the fields this needs (src/core/film.rs:428-436) but
// has no add_sample; the algorithm is pbrt-v3 7.9.2, restated on the CPU in
// oracle/pbrt_oracle.c:orc_ext_tile_add_sample.
//
// Three implementations of the same contract:
//
//   window gather (hot path, radius with h = floor(r + .5) in 1..4, equal on both axes)
//     A sample at nominal pixel n only ever touches pixels n-h .. n+h, so an output pixel can
//     GATHER: walk the samples of its (2h+1)^2 neighbouring pixels in stream order and add the
//     ones whose footprint covers it.  No atomics, fixed order => bit-reproducible, and with
//     mul-then-add identical to the CPU restatement.
//     A thread owns one pixel COLUMN of a CTA-wide strip and marches down the rows holding a
//     2h+1-row window of RGBW accumulators in registers.  Per sample row the CTA first runs a
//     pre-pass, one thread per sample (coalesced float2 + float4 loads): luminance clamp,
//     L*sample_weight, and the table ROW (ify) of each of the 2h+1 window rows packed one per byte
//     (16 = the table's zero row: row outside the footprint).  Records go to shared memory with a padded
//     pixel pitch so that lanes reading sample s of consecutive pixels hit distinct banks.
//     The gather then costs, per (sample, column): one LDS.128 + one LDS.32/64, the ifx index
//     (5 FP ops), and per window row one IDP.4A (table address) + LDS(table) + 3 FMUL + 2 FADD2
//     (exact) or 3 FFMA + FADD (PBRT_SPLAT_FMA).  ptxas fuses a packed mul feeding a packed add into one
//     FFMA2 even under --fmad=false, so the exact variant multiplies in scalar registers and only
//     adds packed (tests/test_abi.py checks the SASS).
//     When the oldest window row can no longer be reached it is converted (rgb_to_xyz) and
//     added to the film: one float4 read-modify-write per pixel per call.
//     The kernel is bound by instruction dispatch, not by memory (tools/ubench_gather.cu):
//     what pays is removing instructions, not hiding latency.
//
//   generic gather (any radius): thread per output pixel, samples read through L1/L2.
//
//   scatter with shared-memory atomics (PBRT_SPLAT_ATOMIC): the textbook formulation, kept for
//     the ncu comparison in profiles/; the order of additions is not deterministic.
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <type_traits>

#include "common.cuh"

namespace pb {



// ---- pieces shared by all variants -------------------------------------------------------

// min(floor(|v|), 15) as an int, v already multiplied out; |v| < 2^23 or NaN.
// Adding 2^23 with round-toward-minus-infinity leaves floor(t) in the low mantissa bits.
__device__ __forceinline__ int table_index(float v) {
    float t = fminf(fabsf(v), 15.f);
    return __float_as_int(__fadd_rd(t, 8388608.f)) & 0xF;
}

__device__ __forceinline__ void clamp_luminance(float4 &L, float max_lum) {
    float ly = luminance(L.x, L.y, L.z);
    if (ly > max_lum) {
        float s = max_lum / ly;
        L.x *= s; L.y *= s; L.z *= s;
    }
}

// film[p] += to_xyz(rgb sum), weight: the body of merge_film_tile (film.rs:318-324)
__device__ __forceinline__ void flush_pixel(float4 *film, const Bounds &owned, int x, int y, float r, float g, float b,
                                            float w) {
    size_t fo = (size_t)(y - owned.y0) * (owned.x1 - owned.x0) + (x - owned.x0);
    float4 p = film[fo];
    float X, Y, Z;
    rgb_to_xyz(r, g, b, X, Y, Z);
    p.x += X; p.y += Y; p.z += Z; p.w += w;
    film[fo] = p;
}

// ---- generic gather ----------------------------------------------------------------------

template <bool FMA>
__global__ void __launch_bounds__(256) splat_gather_generic_kernel(SplatParams P, int hx, int hy) {
    __shared__ float s_table[256];
    s_table[threadIdx.x] = P.table[threadIdx.x];
    __syncthreads();
    const int x = P.tb.x0 + blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = P.tb.y0 + blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= P.tb.x1 || y >= P.tb.y1) return;
    const int W = P.sb.x1 - P.sb.x0;
    const float fx = (float)x, fy = (float)y;
    float ar = 0.f, ag = 0.f, ab = 0.f, aw = 0.f;
    const int ny0 = max(P.sb.y0, y - hy), ny1 = min(P.sb.y1 - 1, y + hy);
    const int nx0 = max(P.sb.x0, x - hx), nx1 = min(P.sb.x1 - 1, x + hx);
    for (int ny = ny0; ny <= ny1; ++ny)
        for (int nx = nx0; nx <= nx1; ++nx) {
            const size_t base = ((size_t)(ny - P.sb.y0) * W + (nx - P.sb.x0)) * (size_t)P.spp;
            for (int s = 0; s < P.spp; ++s) {
                const float2 p = P.xy[base + s];
                if (!(p.x >= (float)nx && p.x <= (float)(nx + 1) && p.y >= (float)ny && p.y <= (float)(ny + 1)))
                    atomicOr(P.err, ERRBIT_NOT_PIXEL_MAJOR);
                const float dx = p.x - 0.5f, dy = p.y - 0.5f;
                // x in [ceil(dx - r), floor(dx + r)]  <=>  dx - r <= x <= dx + r for integer x
                if (!(fx >= dx - P.rx && fx <= dx + P.rx && fy >= dy - P.ry && fy <= dy + P.ry)) continue;
                float4 L = P.rgbw[base + s];
                clamp_luminance(L, P.max_lum);
                const int ix = table_index((fx - dx) * P.irx * 16.f);
                const int iy = table_index((fy - dy) * P.iry * 16.f);
                const float w = s_table[iy * 16 + ix];
                if (FMA) {
                    ar = __fmaf_rn(L.x * L.w, w, ar); ag = __fmaf_rn(L.y * L.w, w, ag); ab = __fmaf_rn(L.z * L.w, w, ab);
                } else {
                    ar += L.x * L.w * w; ag += L.y * L.w * w; ab += L.z * L.w * w;
                }
                aw += w;
            }
        }
    // non-finite radiance (a contract violation) shows in the sums: 0 * x is NaN for x = inf or NaN
    const float z = ar * 0.f + ag * 0.f + ab * 0.f + aw * 0.f;
    if (z != z) atomicOr(P.err, ERRBIT_NONFINITE);
    flush_pixel(P.film, P.owned, x, y, ar, ag, ab, aw);
}

// ---- window gather -----------------------------------------------------------------------

#ifndef PBRT_WIN_UNROLL
#define PBRT_WIN_UNROLL 8
#endif
#ifndef PBRT_PREPASS_BATCH
#define PBRT_PREPASS_BATCH 6
#endif
constexpr int kPrepassBatch = PBRT_PREPASS_BATCH;
constexpr int kWinUnroll = PBRT_WIN_UNROLL;
#ifndef PBRT_WIN_INTERIOR_UNROLL
#define PBRT_WIN_INTERIOR_UNROLL 1
#endif
constexpr int kWinInteriorUnroll = PBRT_WIN_INTERIOR_UNROLL;  // 1 = keep the interior-column loop rolled (smaller code)  // samples per trip of the gather's inner loop

typedef unsigned long long u64;

__device__ __forceinline__ u64 pack2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(u64 v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
    u64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// Row bytes of a sample: for each of the ROWS = 2h+1 output rows of the window, the filter-table
// row (ify, 0..15) it reads — or 16, the all-zero table row, when the sample's footprint does not
// reach that row (only the two outermost rows can be unreachable).  One byte per row, 1..3 words,
// fetched with one LDS.
template <int NW> struct RowBytes;
template <> struct RowBytes<1> { typedef unsigned T; };
template <> struct RowBytes<2> { typedef uint2 T; };
template <> struct RowBytes<3> { typedef uint4 T; };

template <int H>
struct WinCfg {
    static constexpr int ROWS = 2 * H + 1;
    static constexpr int NW = (ROWS + 3) / 4;
    typedef typename RowBytes<NW>::T RB;
};

// Filter table in shared memory: 17 rows x 17 entries.  Row 16 and column 16 are zero: a sample that does not reach a row /
// column adds an exact zero there (x + 0 == x for every finite x), which keeps the inner loop free of
// branches.  The odd row pitch also spreads equal columns of different rows over different banks.
constexpr int TAB_ZERO = 16;
// Row pitch.  The lanes of a warp hold sample s of neighbouring pixels; for stratified streams their
// (row, column) table indices differ by at most 2 each.  With a pitch of 21 entries, 21*drow + dcol is
// never 0 modulo the 32 four-byte (or 16 eight-byte) bank slots in that range, so the lookups of a warp
// fall on distinct banks (a pitch of 17 made (row, col) collide with (row+1, col-1): 30 % extra wavefronts).
// Entries are single 4-byte weights in both variants (measured: scalar multiplies / FFMAs beat the packed
// forms here, tools/ubench_gather.cu).
template <int H, bool FMA> struct TabCfg {
    static constexpr int ENTRY = 4;
    static constexpr int ROW_BYTES = 21 * ENTRY;
    static constexpr int BYTES = (17 * ROW_BYTES + 127) / 128 * 128;
};

// shared-memory layout for one sample row of a CTA strip.  The pixel pitch is odd (in elements)
// so that lanes reading sample s of consecutive pixels fall on distinct banks for 4/8/16-byte loads.
template <int H, int TW, bool FMA>
struct WinSmem {
    static constexpr int NPX = TW + 2 * H;
    __host__ __device__ static int pitch(int spp) { return spp | 1; }
    __host__ __device__ static size_t bytes(int spp) {
        return TabCfg<H, FMA>::BYTES + (size_t)NPX * pitch(spp) * (16 + sizeof(typename WinCfg<H>::RB));
    }
};

__device__ __forceinline__ unsigned rb_word(unsigned v, int) { return v; }
__device__ __forceinline__ unsigned rb_word(const uint2 &v, int k) { return k == 0 ? v.x : v.y; }
__device__ __forceinline__ unsigned rb_word(const uint4 &v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }
__device__ __forceinline__ void rb_make(unsigned &o, const unsigned *w) { o = w[0]; }
__device__ __forceinline__ void rb_make(uint2 &o, const unsigned *w) { o = make_uint2(w[0], w[1]); }
__device__ __forceinline__ void rb_make(uint4 &o, const unsigned *w) { o = make_uint4(w[0], w[1], w[2], 0u); }
template <typename RB>
__device__ __forceinline__ unsigned rb_byte(const RB &v, int i) {
    return __byte_perm(rb_word(v, i >> 2), 0, 0x4440 + (i & 3));
}

// float bits whose low byte is min(floor(|v|), 15): see table_index()
__device__ __forceinline__ unsigned bin_bits(float v) {
    return (unsigned)__float_as_int(__fadd_rd(fminf(fabsf(v), 15.f), 8388608.f));
}

// Table fetch.  Deliberately not `volatile`: the table is constant once the kernel's first barrier has
// passed and the address depends on data read after it, so the compiler may schedule these loads
// early across the unrolled samples without being able to hoist them too far.
__device__ __forceinline__ float lds_f32(unsigned addr) {
    float v;
    asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}

template <int H, int TW, bool FMA>
__global__ void __launch_bounds__(TW) splat_window_kernel(SplatParams P) {
    constexpr int ROWS = WinCfg<H>::ROWS;
    constexpr int NW = WinCfg<H>::NW;
    typedef typename WinCfg<H>::RB RB;
    constexpr int NPX = WinSmem<H, TW, FMA>::NPX;
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int TAB_ROW_BYTES = TabCfg<H, FMA>::ROW_BYTES;
    constexpr int TAB_ENTRY = TabCfg<H, FMA>::ENTRY;
    constexpr int WIN_TABLE_BYTES = TabCfg<H, FMA>::BYTES;
    float4 *s_a = reinterpret_cast<float4 *>(smem + WIN_TABLE_BYTES);
    const int pitch = WinSmem<H, TW, FMA>::pitch(P.spp);
    RB *s_b = reinterpret_cast<RB *>(smem + WIN_TABLE_BYTES + (size_t)NPX * pitch * 16);

    const int tid = threadIdx.x;
    float4 *tile_out = nullptr;
    if (P.tiles) {  // batched: this CTA works on tile blockIdx.z (uniform across the grid's z slice)
        const SplatTile t = P.tiles[blockIdx.z];
        P.sb = t.sb;
        P.tb = t.tb;
        P.xy += t.sample_offset;
        P.rgbw += t.sample_offset;
        tile_out = P.tile_out + t.pixel_offset;
        if (P.tb.x1 <= P.tb.x0 || P.tb.y1 <= P.tb.y0 || P.sb.x1 <= P.sb.x0 || P.sb.y1 <= P.sb.y0) return;
        if (P.tb.x0 + (int)blockIdx.x * TW >= P.tb.x1) return;  // the grid is sized for the widest tile
    }
    for (int i = tid; i < 17 * 17; i += TW) {
        const int ty = i / 17, tx = i - ty * 17;
        const float w = (ty < 16 && tx < 16) ? P.table[ty * 16 + tx] : 0.f;
        *reinterpret_cast<float *>(smem + ty * TAB_ROW_BYTES + tx * 4) = w;
    }
    // shared-window address of the table, kept opaque so it lives in a register instead of being
    // rematerialised (S2UR/ULEA) at every use
    unsigned tab_base = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("mov.u32 %0, %0;" : "+r"(tab_base));

    const int cx0 = P.tb.x0 + blockIdx.x * TW;             // first output column of the strip
    const int cy0 = P.tb.y0 + blockIdx.y * P.rows_per_cta; // first output row
    const int cy1 = min(cy0 + P.rows_per_cta, P.tb.y1);
    const int x = cx0 + tid;
    const bool col_ok = x < P.tb.x1;
    const float fx = (float)x;
    const int spp = P.spp;
    const int W = P.sb.x1 - P.sb.x0;
    const float inf = __int_as_float(0x7f800000);
    const bool clamp_on = P.max_lum < inf;
    // (d * inv_radius) * 16 == d * (inv_radius * 16) bit for bit: scaling by a power of two commutes with rounding
    const float irx16 = P.irx * 16.f, iry16 = P.iry * 16.f;

    // staged nominal pixels of a row: [sx0, sx1); local index = nx - (cx0 - H)
    const int sx0 = max(cx0 - H, P.sb.x0), sx1 = min(cx0 + TW + H, P.sb.x1);
    const int nstaged = max(sx1 - sx0, 0) * spp;
    // element e = tid + k*TW of the staged run is sample s of staged pixel q: advance (q, s) without dividing
    const int q0 = tid / spp, r0 = tid - q0 * spp;
    const int dq = TW / spp, dr = TW - dq * spp;
    const int pl_base = sx0 - (cx0 - H);

    u64 acc_rg[ROWS], acc_bw[ROWS];
#pragma unroll
    for (int j = 0; j < ROWS; ++j) acc_rg[j] = acc_bw[j] = 0ull;
    unsigned errbits = 0;  // contract violations seen by this thread; reported once at the end
    float vmax = 0.f, fin = 0.f;  // see the pre-pass
    const int slot_step = dq * pitch + dr;
    const float fdq = (float)dq;
    // lanes that gather (the last strip of a row may be narrower than the CTA)
    const unsigned gather_mask = __ballot_sync(0xffffffffu, col_ok);

    for (int ny = cy0 - H; ny < cy1 + H; ++ny) {
        const bool row_has_samples = ny >= P.sb.y0 && ny < P.sb.y1 && nstaged > 0;
        if (row_has_samples) {
            __syncthreads();  // previous row fully consumed (also orders the table fill)
            // ---------------- pre-pass: one thread per sample of the row ----------------
            const size_t row_base = ((size_t)(ny - P.sb.y0) * W + (sx0 - P.sb.x0)) * (size_t)spp;
            const float2 *gxy = P.xy + row_base;
            const float4 *grgbw = P.rgbw + row_base;
            const float fny = (float)ny;
            // element e of the staged run = (staged pixel q, sample sidx); its record slot and the float of its
            // pixel coordinate are carried along instead of being recomputed per sample
            int sidx = r0;
            int slot = (pl_base + q0) * pitch + r0;
            float fnx = (float)(sx0 + q0);
            constexpr int U = kPrepassBatch;  // samples per thread per trip: all loads issued before any is consumed
            for (int e0 = tid; e0 < nstaged; e0 += U * TW) {
                float2 p[U];
                float4 L[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (e0 + u * TW < nstaged) {
                        p[u] = ldg_stream(&gxy[e0 + u * TW]);
                        L[u] = ldg_stream(&grgbw[e0 + u * TW]);
                    }
                }
                // WRAP = false when spp divides the strip width: a thread then keeps its sample index for the whole row
                auto process = [&](auto wrap_tag) {
                    constexpr bool WRAP = decltype(wrap_tag)::value;
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        if (e0 + u * TW < nstaged) {
                            if (clamp_on) clamp_luminance(L[u], P.max_lum);
                            const float cr = L[u].x * L[u].w, cg = L[u].y * L[u].w, cb = L[u].z * L[u].w;
                            const float pdx = p[u].x - 0.5f, pdy = p[u].y - 0.5f;
                            // Contract: the sample lies in its nominal pixel (closed) and everything is finite.  Checked on
                            // the discrete position every later step uses: |pd - n| <= 0.5; the running maximum and a
                            // NaN-propagating sum (0 * x is NaN for x = inf or NaN) are tested once at the end.
                            const float tx = pdx - fnx, ty = pdy - fny;
                            vmax = fmaxf(vmax, fmaxf(fabsf(tx), fabsf(ty)));
                            fin = __fmaf_rn(fabsf(cr) + fabsf(cg) + fabsf(cb) + fabsf(tx) + fabsf(ty), 0.f, fin);
                            unsigned bits[12];
#pragma unroll
                            for (int j = 0; j < ROWS; ++j) bits[j] = bin_bits((fny + (float)(j - H) - pdy) * iry16);
#pragma unroll
                            for (int j = ROWS; j < 12; ++j) bits[j] = 0u;
                            // only the outermost rows can fall outside [ceil(pdy - r), floor(pdy + r)]
                            if (!(fny - (float)H >= pdy - P.ry)) bits[0] = TAB_ZERO;
                            if (!(fny + (float)H <= pdy + P.ry)) bits[ROWS - 1] = TAB_ZERO;
                            unsigned words[NW];
#pragma unroll
                            for (int k = 0; k < NW; ++k)
                                words[k] = __byte_perm(__byte_perm(bits[4 * k], bits[4 * k + 1], 0x0040),
                                                       __byte_perm(bits[4 * k + 2], bits[4 * k + 3], 0x0040), 0x5410);
                            s_a[slot] = make_float4(cr, cg, cb, pdx);
                            rb_make(s_b[slot], words);
                        }
                        slot += slot_step;
                        fnx += fdq;
                        if (WRAP) {
                            sidx += dr;
                            if (sidx >= spp) { sidx -= spp; slot += pitch - spp; fnx += 1.f; }
                        }
                    }
                };
                if (dr == 0) process(std::false_type{});
                else process(std::true_type{});
            }
            __syncthreads();
            // pull the next sample row of this strip into L2 while this one is gathered, so that its
            // pre-pass loads pay L2 latency instead of DRAM latency
#ifndef PBRT_NO_PREFETCH
            if (ny + 1 < P.sb.y1 && ny + 1 < cy1 + H) {
                const char *nxy = reinterpret_cast<const char *>(gxy + (size_t)W * spp);
                const char *nrgbw = reinterpret_cast<const char *>(grgbw + (size_t)W * spp);
                for (int o = tid * 128; o < nstaged * 8; o += TW * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(nxy + o));
                for (int o = tid * 128; o < nstaged * 16; o += TW * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(nrgbw + o));
            }
#endif
            // ---------------- gather: this thread's column against the row ----------------
            if (col_ok) {
                // EDGE = -1 / +1: the outermost columns, whose samples may not reach this pixel; 0: interior
                auto visit = [&](const int d, auto edge_tag) {
                    constexpr int EDGE = decltype(edge_tag)::value;
                    const int nx = x + d;
                    const bool in_range = nx >= P.sb.x0 && nx < P.sb.x1;
                    // lanes that take part in this visit's votes (columns at the film edge drop out)
                    const unsigned visit_mask = __ballot_sync(gather_mask, in_range);
                    if (!in_range) return;
                    const int pl = nx - (cx0 - H);
                    const float4 *pa = s_a + pl * pitch;
                    const RB *pb = s_b + pl * pitch;
                    // all window rows of one sample against this column; ifx = table column (TAB_ZERO adds exact zeros)
                    // `ifx_bits`: a word whose low byte is the table column and whose other bytes do not matter
                    auto taps = [&](const float4 a, const RB yb, const unsigned ifx_bits) {
                        // shared address of the table column: low byte x 4 + base, in one dot-product instruction
                        const unsigned xcol = __dp4a(ifx_bits, (unsigned)TAB_ENTRY, tab_base);
#pragma unroll
                        for (int j = 0; j < ROWS; ++j) {
                            // row byte j times the table's row pitch, plus the column address, in one dot-product
                            // instruction: dp4a(bytes, pitch in byte lane j, xcol)
                            const unsigned waddr = __dp4a(rb_word(yb, j >> 2), (unsigned)TAB_ROW_BYTES << (8 * (j & 3)), xcol);
                            const float w = lds_f32(waddr);
                            if (FMA) {
                                // three scalar FFMAs and one FADD on the halves of the accumulator pairs: cheaper to
                                // dispatch than two packed FFMA2 (profiles/r1_ubench_gather.txt)
                                float r, g, b, ws;
                                unpack2(acc_rg[j], r, g);
                                unpack2(acc_bw[j], b, ws);
                                acc_rg[j] = pack2(__fmaf_rn(a.x, w, r), __fmaf_rn(a.y, w, g));
                                acc_bw[j] = pack2(__fmaf_rn(a.z, w, b), ws + w);
                            } else {
                                // products by scalar multiplies (rounded like the CPU's), sums packed: (r*w, g*w) and (b*w, w)
                                acc_rg[j] = add2(acc_rg[j], pack2(a.x * w, a.y * w));
                                acc_bw[j] = add2(acc_bw[j], pack2(a.z * w, w));
                            }
                        }
                    };
                    if (EDGE == 0) {
#pragma unroll kWinUnroll
                        for (int s = 0; s < spp; ++s) {
                            const float4 a = pa[s];
                            taps(a, pb[s], bin_bits((fx - a.w) * irx16));
                        }
                    } else {
                        // Outermost columns: x must lie in [ceil(pdx - r), floor(pdx + r)].  Lanes hold sample s of
                        // neighbouring pixels, which for stratified streams sit in the same stratum and agree: vote
                        // and skip the sample for the whole warp when no lane needs it; a lane that does not need
                        // it reads the zero column.
#pragma unroll kWinUnroll
                        for (int s = 0; s < spp; ++s) {
                            const float4 a = pa[s];
                            const bool reach = EDGE > 0 ? fx >= a.w - P.rx : fx <= a.w + P.rx;
                            if (!__any_sync(visit_mask, reach)) continue;
                            taps(a, pb[s], reach ? bin_bits((fx - a.w) * irx16) : (unsigned)TAB_ZERO);
                        }
                    }
                };
                // columns are visited left to right: the accumulation order of the oracle (and the reference's add_sample)
                visit(-H, std::integral_constant<int, -1>{});
#pragma unroll kWinInteriorUnroll
                for (int d = -H + 1; d <= H - 1; ++d) visit(d, std::integral_constant<int, 0>{});
                visit(H, std::integral_constant<int, 1>{});
            }
        }
        // output row ny - H is complete: no later sample row reaches it
        const int yo = ny - H;
        if (col_ok && yo >= cy0 && yo < cy1) {
            float r, g, b, w;
            unpack2(acc_rg[0], r, g);
            unpack2(acc_bw[0], b, w);
            if (tile_out)  // FilmTilePixel {contrib_sum, filter_weight_sum} of this tile (film.rs:39-42)
                tile_out[(size_t)(yo - P.tb.y0) * (P.tb.x1 - P.tb.x0) + (x - P.tb.x0)] = make_float4(r, g, b, w);
            else
                flush_pixel(P.film, P.owned, x, yo, r, g, b, w);
        }
#pragma unroll
        for (int j = 0; j + 1 < ROWS; ++j) { acc_rg[j] = acc_rg[j + 1]; acc_bw[j] = acc_bw[j + 1]; }
        acc_rg[ROWS - 1] = acc_bw[ROWS - 1] = 0ull;
    }
    if (!(vmax <= 0.5f)) errbits |= ERRBIT_NOT_PIXEL_MAJOR;
    if (fin != fin) errbits |= ERRBIT_NONFINITE;
    if (errbits) atomicOr(P.err, (int)errbits);
}

// ---- scatter with shared-memory atomics --------------------------------------------------

// CTA = 32x8 nominal pixels; accumulates into a (32+2h)x(8+2h) shared tile with atomics and
// flushes it with global atomics into a scratch RGBW tile (overlapping halos of neighbouring CTAs).
constexpr int AT_W = 32, AT_H = 8;

__global__ void __launch_bounds__(256) splat_atomic_kernel(SplatParams P, int hx, int hy, float4 *__restrict__ scratch) {
    extern __shared__ __align__(16) unsigned char smem[];
    float *s_table = reinterpret_cast<float *>(smem);
    float *s_tile = s_table + 256;
    const int tw = AT_W + 2 * hx, th = AT_H + 2 * hy;
    for (int i = threadIdx.x; i < 256; i += 256) s_table[i] = P.table[i];
    for (int i = threadIdx.x; i < tw * th * 4; i += 256) s_tile[i] = 0.f;
    __syncthreads();
    const int bx0 = P.sb.x0 + blockIdx.x * AT_W, by0 = P.sb.y0 + blockIdx.y * AT_H;
    const int nx = bx0 + (threadIdx.x & 31), ny = by0 + (threadIdx.x >> 5);
    const int W = P.sb.x1 - P.sb.x0;
    const int ox = bx0 - hx, oy = by0 - hy;  // film coords of smem tile origin
    if (nx < P.sb.x1 && ny < P.sb.y1) {
        const size_t base = ((size_t)(ny - P.sb.y0) * W + (nx - P.sb.x0)) * (size_t)P.spp;
        for (int s = 0; s < P.spp; ++s) {
            // lanes walk consecutive pixels: a 2-D strided read, spp*8 B apart
            const float2 p = P.xy[base + s];
            float4 L = P.rgbw[base + s];
            if (!(p.x >= (float)nx && p.x <= (float)(nx + 1) && p.y >= (float)ny && p.y <= (float)(ny + 1)))
                atomicOr(P.err, ERRBIT_NOT_PIXEL_MAJOR);
            clamp_luminance(L, P.max_lum);
            const float dx = p.x - 0.5f, dy = p.y - 0.5f;
            const int p0x = max(max(__float2int_ru(dx - P.rx), P.tb.x0), ox);
            const int p0y = max(max(__float2int_ru(dy - P.ry), P.tb.y0), oy);
            const int p1x = min(min(__float2int_rd(dx + P.rx) + 1, P.tb.x1), ox + tw);
            const int p1y = min(min(__float2int_rd(dy + P.ry) + 1, P.tb.y1), oy + th);
            const float cr = L.x * L.w, cg = L.y * L.w, cb = L.z * L.w;
            const float z = cr * 0.f + cg * 0.f + cb * 0.f;
            if (z != z) atomicOr(P.err, ERRBIT_NONFINITE);
            for (int y = p0y; y < p1y; ++y) {
                const int iy = table_index(((float)y - dy) * P.iry * 16.f);
                for (int x = p0x; x < p1x; ++x) {
                    const int ix = table_index(((float)x - dx) * P.irx * 16.f);
                    const float w = s_table[iy * 16 + ix];
                    float *px = s_tile + ((y - oy) * tw + (x - ox)) * 4;
                    atomicAdd(px + 0, cr * w);
                    atomicAdd(px + 1, cg * w);
                    atomicAdd(px + 2, cb * w);
                    atomicAdd(px + 3, w);
                }
            }
        }
    }
    __syncthreads();
    const int ttw = P.tb.x1 - P.tb.x0;
    for (int i = threadIdx.x; i < tw * th; i += 256) {
        const int x = ox + i % tw, y = oy + i / tw;
        if (x < P.tb.x0 || x >= P.tb.x1 || y < P.tb.y0 || y >= P.tb.y1) continue;
        const float *v = s_tile + i * 4;
        if (v[3] == 0.f && v[0] == 0.f && v[1] == 0.f && v[2] == 0.f) continue;
        float *o = reinterpret_cast<float *>(&scratch[(size_t)(y - P.tb.y0) * ttw + (x - P.tb.x0)]);
        atomicAdd(o + 0, v[0]); atomicAdd(o + 1, v[1]); atomicAdd(o + 2, v[2]); atomicAdd(o + 3, v[3]);
    }
}

__global__ void __launch_bounds__(256) merge_scratch_kernel(float4 *__restrict__ film, Bounds owned, Bounds tb,
                                                            const float4 *__restrict__ tile) {
    const int tw = tb.x1 - tb.x0;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= tw) return;
    for (int y = blockIdx.y; y < tb.y1 - tb.y0; y += gridDim.y) {
        float4 t = tile[(size_t)y * tw + x];
        flush_pixel(film, owned, tb.x0 + x, tb.y0 + y, t.x, t.y, t.z, t.w);
    }
}

// ---- launch ------------------------------------------------------------------------------

static int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

template <int H, int TW, bool FMA>
static int launch_window(const SplatParams &P0) {
    SplatParams P = P0;
    // PBRT_B200_SMEM_PAD: experiment knob — extra dynamic shared memory per CTA, i.e. fewer resident CTAs
    const size_t smem = WinSmem<H, TW, FMA>::bytes(P.spp) + (size_t)env_int("PBRT_B200_SMEM_PAD", 0);
    static bool attr_set = false;
    if (!attr_set) {
        PB_CUDA(cudaFuncSetAttribute(splat_window_kernel<H, TW, FMA>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     227 * 1024));
        attr_set = true;
    }
    int per_sm = 0;
    PB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, splat_window_kernel<H, TW, FMA>, TW, smem));
    if (per_sm < 1) return -1;
    const int cols = (bw(P.tb) + TW - 1) / TW;
    const int rows = bh(P.tb);
    // One wave: at most as many CTAs as are resident at once (a ragged second wave would double the
    // run time), but strips no shorter than 4h rows so the 2h halo rows stay a modest overhead.
    const int resident = ctx().sm_count * per_sm;
    int segs = std::max(1, resident / cols);
    segs = std::min(segs, std::max(1, rows / (4 * H)));
    int rpc = env_int("PBRT_B200_ROWS_PER_CTA", 0);
    P.rows_per_cta = rpc > 0 ? rpc : (rows + segs - 1) / segs;
    segs = (rows + P.rows_per_cta - 1) / P.rows_per_cta;
    dim3 grid(cols, segs);
    splat_window_kernel<H, TW, FMA><<<grid, TW, smem, ctx().stream>>>(P);
    PB_LAUNCH_CHECK("splat_window_kernel");
    return PBRT_OK;
}

template <int H, bool FMA>
static int pick_width(const SplatParams &P) {
    // widest strip whose row of records leaves room for >= 2 CTAs per SM, else whatever fits
    const int force = env_int("PBRT_B200_TW", 0);
    if (force == 128) return launch_window<H, 128, FMA>(P);
    if (force == 64) return launch_window<H, 64, FMA>(P);
    if (force == 32) return launch_window<H, 32, FMA>(P);
    if (WinSmem<H, 128, FMA>::bytes(P.spp) * 2 <= 227 * 1024) return launch_window<H, 128, FMA>(P);
    if (WinSmem<H, 64, FMA>::bytes(P.spp) * 2 <= 227 * 1024) return launch_window<H, 64, FMA>(P);
    if (WinSmem<H, 32, FMA>::bytes(P.spp) <= 227 * 1024) return launch_window<H, 32, FMA>(P);
    return -1;
}

template <bool FMA>
static int pick_window(const SplatParams &P, int h) {
    switch (h) {
    case 1: return pick_width<1, FMA>(P);
    case 2: return pick_width<2, FMA>(P);
    case 3: return pick_width<3, FMA>(P);
    case 4: return pick_width<4, FMA>(P);
    }
    return -1;
}

// The same scatter with warp-aggregated atomics (PBRT_B200_ATOMIC_AGG=1; kept for the ncu comparison of DESIGN.md 5.5):
// a thread per SAMPLE, so that the lanes of a warp hold consecutive samples of one or two pixels and every tap of theirs
// goes to the same destination pixel; a segmented shuffle reduction over the run of lanes of a pixel leaves one shared
// atomic per run, tap and channel instead of one per lane.  The order of additions is a tree: tolerance mode only.
__global__ void __launch_bounds__(256) splat_atomic_agg_kernel(SplatParams P, int hx, int hy, float4 *__restrict__ scratch) {
    extern __shared__ __align__(16) unsigned char smem[];
    float *s_table = reinterpret_cast<float *>(smem);
    float *s_tile = s_table + 256;
    const int tw = AT_W + 2 * hx, th = AT_H + 2 * hy;
    for (int i = threadIdx.x; i < 256; i += 256) s_table[i] = P.table[i];
    for (int i = threadIdx.x; i < tw * th * 4; i += 256) s_tile[i] = 0.f;
    __syncthreads();
    const int bx0 = P.sb.x0 + blockIdx.x * AT_W, by0 = P.sb.y0 + blockIdx.y * AT_H;
    const int nw = min(AT_W, P.sb.x1 - bx0), nh = min(AT_H, P.sb.y1 - by0);
    const int W = P.sb.x1 - P.sb.x0, spp = P.spp;
    const int ox = bx0 - hx, oy = by0 - hy;  // film coords of smem tile origin
    const int nsamp = nw * nh * spp;
    const int lane = threadIdx.x & 31;
    for (int e0 = 0; e0 < nsamp; e0 += 256) {
        const int e = e0 + threadIdx.x;
        const bool valid = e < nsamp;
        const int pix = valid ? e / spp : -1 - lane;  // invalid lanes: runs of their own
        const int s = e - (valid ? pix : 0) * spp;
        const int nx = bx0 + (valid ? pix % nw : 0), ny = by0 + (valid ? pix / nw : 0);
        float cr = 0.f, cg = 0.f, cb = 0.f, dx = 0.f, dy = 0.f;
        int p0x = 0, p1x = 0, p0y = 0, p1y = 0;
        if (valid) {
            const size_t idx = ((size_t)(ny - P.sb.y0) * W + (nx - P.sb.x0)) * (size_t)spp + s;
            const float2 p = P.xy[idx];
            float4 L = P.rgbw[idx];
            if (!(p.x >= (float)nx && p.x <= (float)(nx + 1) && p.y >= (float)ny && p.y <= (float)(ny + 1)))
                atomicOr(P.err, ERRBIT_NOT_PIXEL_MAJOR);
            clamp_luminance(L, P.max_lum);
            dx = p.x - 0.5f; dy = p.y - 0.5f;
            p0x = max(max(__float2int_ru(dx - P.rx), P.tb.x0), ox);
            p0y = max(max(__float2int_ru(dy - P.ry), P.tb.y0), oy);
            p1x = min(min(__float2int_rd(dx + P.rx) + 1, P.tb.x1), ox + tw);
            p1y = min(min(__float2int_rd(dy + P.ry) + 1, P.tb.y1), oy + th);
            cr = L.x * L.w; cg = L.y * L.w; cb = L.z * L.w;
            const float z = cr * 0.f + cg * 0.f + cb * 0.f;
            if (z != z) atomicOr(P.err, ERRBIT_NONFINITE);
        }
        // lanes of one pixel are a contiguous run: which partners of the doubling steps belong to this lane's run
        unsigned same = 0;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const int other = __shfl_down_sync(0xffffffffu, pix, 1 << k);
            if (lane + (1 << k) < 32 && other == pix) same |= 1u << k;
        }
        const int before = __shfl_up_sync(0xffffffffu, pix, 1);
        const bool head = valid && (lane == 0 || before != pix);
        for (int jy = 0; jy <= 2 * hy; ++jy) {
            const int y = ny - hy + jy;
            const bool in_y = y >= p0y && y < p1y;
            const int iy = table_index(((float)y - dy) * P.iry * 16.f);
            for (int jx = 0; jx <= 2 * hx; ++jx) {
                const int x = nx - hx + jx;
                const bool inside = in_y && x >= p0x && x < p1x;
                const int ix = table_index(((float)x - dx) * P.irx * 16.f);
                const float w = inside ? s_table[iy * 16 + ix] : 0.f;
                float v0 = cr * w, v1 = cg * w, v2 = cb * w, v3 = w;
                const unsigned any = __ballot_sync(0xffffffffu, inside);
                if (!any) continue;
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    const float t0 = __shfl_down_sync(0xffffffffu, v0, 1 << k), t1 = __shfl_down_sync(0xffffffffu, v1, 1 << k);
                    const float t2 = __shfl_down_sync(0xffffffffu, v2, 1 << k), t3 = __shfl_down_sync(0xffffffffu, v3, 1 << k);
                    if (same & (1u << k)) { v0 += t0; v1 += t1; v2 += t2; v3 += t3; }
                }
                // the destination pixel is the same for the whole run; it lies in the tile whenever some lane is inside
                if (head && x >= max(P.tb.x0, ox) && x < min(P.tb.x1, ox + tw) && y >= max(P.tb.y0, oy) && y < min(P.tb.y1, oy + th) &&
                    (v3 != 0.f || v0 != 0.f || v1 != 0.f || v2 != 0.f)) {
                    float *px = s_tile + ((y - oy) * tw + (x - ox)) * 4;
                    atomicAdd(px + 0, v0); atomicAdd(px + 1, v1); atomicAdd(px + 2, v2); atomicAdd(px + 3, v3);
                }
            }
        }
    }
    __syncthreads();
    const int ttw = P.tb.x1 - P.tb.x0;
    for (int i = threadIdx.x; i < tw * th; i += 256) {
        const int x = ox + i % tw, y = oy + i / tw;
        if (x < P.tb.x0 || x >= P.tb.x1 || y < P.tb.y0 || y >= P.tb.y1) continue;
        const float *v = s_tile + i * 4;
        if (v[3] == 0.f && v[0] == 0.f && v[1] == 0.f && v[2] == 0.f) continue;
        float *o = reinterpret_cast<float *>(&scratch[(size_t)(y - P.tb.y0) * ttw + (x - P.tb.x0)]);
        atomicAdd(o + 0, v[0]); atomicAdd(o + 1, v[1]); atomicAdd(o + 2, v[2]); atomicAdd(o + 3, v[3]);
    }
}

// ---- batched tiles ---------------------------------------------------------------------------

template <int H, bool FMA>
static int launch_window_batched(const SplatParams &P0, int ntiles, int max_w, int max_h) {
    constexpr int TW = 32;  // renderer tiles are small (16x16 samples -> 20x20 pixels at r = 2)
    SplatParams P = P0;
    const size_t smem = WinSmem<H, TW, FMA>::bytes(P.spp);
    if (smem > 227 * 1024) return -1;
    static bool attr_set = false;
    if (!attr_set) {
        PB_CUDA(cudaFuncSetAttribute(splat_window_kernel<H, TW, FMA>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     227 * 1024));
        attr_set = true;
    }
    P.rows_per_cta = std::max(max_h, 1);  // one CTA column strip walks the whole tile height
    for (int z0 = 0; z0 < ntiles; z0 += 65535) {
        const int nz = std::min(65535, ntiles - z0);
        SplatParams Q = P;
        Q.tiles = P.tiles + z0;
        dim3 grid((max_w + TW - 1) / TW, 1, nz);
        splat_window_kernel<H, TW, FMA><<<grid, TW, smem, ctx().stream>>>(Q);
        PB_LAUNCH_CHECK("splat_window_kernel(batched)");
    }
    return PBRT_OK;
}

// tiles: device array of ntiles SplatTile; tile_out: device RGBW scratch the tiles' pixels go to
int launch_splat_tiles(PbrtFilm *f, int ntiles, const SplatTile *d_tiles, int max_w, int max_h, int spp, const float2 *xy,
                       const float4 *rgbw, float4 *tile_out, int mode) {
    SplatParams P;
    P.sb = P.tb = Bounds{0, 0, 0, 0};
    P.owned = f->owned;
    P.spp = spp;
    P.rx = f->radius[0]; P.ry = f->radius[1];
    P.irx = f->inv_radius[0]; P.iry = f->inv_radius[1];
    P.max_lum = f->max_lum;
    P.xy = xy; P.rgbw = rgbw;
    P.table = f->d_table;
    P.film = f->d_xyzw;
    P.err = f->d_err;
    P.rows_per_cta = 0;
    P.tiles = d_tiles;
    P.tile_out = tile_out;
    const int hx = (int)floorf(P.rx + 0.5f), hy = (int)floorf(P.ry + 0.5f);
    if (hx != hy || hx < 1 || hx > 4) return -1;  // caller falls back to one launch per tile
    {   // radius 2 or 4 and spp dividing 32: the phase-class gather (splat_class.cu); -1 = not served
        const int rc = launch_splat_class_tiles(f, P, ntiles, max_w, mode);
        if (rc >= 0) return rc;
    }
    const bool fma = mode == PBRT_SPLAT_FMA;
    switch (hx) {
    case 1: return fma ? launch_window_batched<1, true>(P, ntiles, max_w, max_h) : launch_window_batched<1, false>(P, ntiles, max_w, max_h);
    case 2: return fma ? launch_window_batched<2, true>(P, ntiles, max_w, max_h) : launch_window_batched<2, false>(P, ntiles, max_w, max_h);
    case 3: return fma ? launch_window_batched<3, true>(P, ntiles, max_w, max_h) : launch_window_batched<3, false>(P, ntiles, max_w, max_h);
    case 4: return fma ? launch_window_batched<4, true>(P, ntiles, max_w, max_h) : launch_window_batched<4, false>(P, ntiles, max_w, max_h);
    }
    return -1;
}

static int g_force_generic = 0;

int launch_splat_tile(PbrtFilm *f, const Bounds &sb, const Bounds &tb, int spp, const float2 *xy, const float4 *rgbw,
                      int mode) {
    SplatParams P;
    P.sb = sb; P.tb = tb; P.owned = f->owned;
    P.spp = spp;
    P.rx = f->radius[0]; P.ry = f->radius[1];
    P.irx = f->inv_radius[0]; P.iry = f->inv_radius[1];
    P.max_lum = f->max_lum;
    P.xy = xy; P.rgbw = rgbw;
    P.table = f->d_table;
    P.film = f->d_xyzw;
    P.err = f->d_err;
    P.rows_per_cta = 0;
    P.tiles = nullptr;
    P.tile_out = nullptr;
    if (!(P.rx > 0.f) || !(P.ry > 0.f) || P.rx > 1024.f || P.ry > 1024.f)
        return fail(PBRT_E_UNSUPPORTED, "filter radius (%g, %g) outside (0, 1024]", P.rx, P.ry);
    // a sample in nominal pixel n reaches pixels n - h .. n + h, h = floor(r + .5)
    const int hx = (int)floorf(P.rx + 0.5f), hy = (int)floorf(P.ry + 0.5f);

    if (mode == PBRT_SPLAT_ATOMIC) {
        const size_t px = (size_t)bw(tb) * bh(tb);
        const size_t smem = 1024 + (size_t)(AT_W + 2 * hx) * (AT_H + 2 * hy) * 16;
        if (smem > 227 * 1024) return fail(PBRT_E_UNSUPPORTED, "radius too large for the shared-memory tile");
        if (px > f->scratch_tile_px) {
            cudaFree(f->d_scratch_tile);
            f->d_scratch_tile = nullptr;
            f->scratch_tile_px = 0;
            PB_CUDA(cudaMalloc(&f->d_scratch_tile, px * sizeof(float4)));
            f->scratch_tile_px = px;
        }
        PB_CUDA(cudaMemsetAsync(f->d_scratch_tile, 0, px * sizeof(float4), ctx().stream));
        static bool attr_set = false;
        if (!attr_set) {
            PB_CUDA(cudaFuncSetAttribute(splat_atomic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            attr_set = true;
        }
        dim3 grid((bw(sb) + AT_W - 1) / AT_W, (bh(sb) + AT_H - 1) / AT_H);
        const char *agg = getenv("PBRT_B200_ATOMIC_AGG");
        if (agg && *agg == '1') {
            static bool agg_attr_set = false;
            if (!agg_attr_set) {
                PB_CUDA(cudaFuncSetAttribute(splat_atomic_agg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
                agg_attr_set = true;
            }
            splat_atomic_agg_kernel<<<grid, 256, smem, ctx().stream>>>(P, hx, hy, f->d_scratch_tile);
            PB_LAUNCH_CHECK("splat_atomic_agg_kernel");
        } else {
            splat_atomic_kernel<<<grid, 256, smem, ctx().stream>>>(P, hx, hy, f->d_scratch_tile);
            PB_LAUNCH_CHECK("splat_atomic_kernel");
        }
        dim3 mgrid((bw(tb) + 255) / 256, std::min(bh(tb), 4096));
        merge_scratch_kernel<<<mgrid, 256, 0, ctx().stream>>>(f->d_xyzw, f->owned, tb, f->d_scratch_tile);
        PB_LAUNCH_CHECK("merge_scratch_kernel");
        return PBRT_OK;
    }

    const bool fma = mode == PBRT_SPLAT_FMA;
    if (!g_force_generic) {
        // radius 2 or 4: the phase-class gather (splat_class.cu); -1 = not served, fall through
        int rc = launch_splat_class(f, P, mode);
        if (rc >= 0) return rc;
    }
    if (!g_force_generic && hx == hy && hx >= 1 && hx <= 4) {
        int rc = fma ? pick_window<true>(P, hx) : pick_window<false>(P, hx);
        if (rc >= 0) return rc;
    }
    dim3 grid((bw(tb) + 31) / 32, (bh(tb) + 7) / 8);
    if (fma)
        splat_gather_generic_kernel<true><<<grid, 256, 0, ctx().stream>>>(P, hx, hy);
    else
        splat_gather_generic_kernel<false><<<grid, 256, 0, ctx().stream>>>(P, hx, hy);
    PB_LAUNCH_CHECK("splat_gather_generic_kernel");
    return PBRT_OK;
}

}  // namespace pb

// test hook: route add_samples_tile through the generic gather regardless of radius
extern "C" int pbrt_b200_debug_force_generic_splat(int on) {
    pb::g_force_generic = on;
    return PBRT_OK;
}
