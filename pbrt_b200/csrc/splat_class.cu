// splat_class.cu — [T2, extension] the phase-class gather: FilmTile::add_sample for a pixel-major stream when the
// filter radius is 2 or 4 pixels on both axes (the radii of triangle / gaussian / mitchell and of lanczos-sinc).
// Same contract and the same results, bit for bit, as splat_window_kernel (splat.cu) and the CPU restatement
// (oracle/pbrt_oracle.c:orc_ext_tile_add_sample, pbrt-v3 7.9.2 over the fields at src/core/film.rs:428-436).
//
// Why it is faster.  The window gather is bound by instruction dispatch: per (sample, pixel) tap it
// pays a table-address IDP and an LDS, per (sample, column) visit five instructions for the table column, and it
// computes every row of the 2h+1 window although a sample reaches only 2h of them.  All of that is a function of the
// sample's sub-pixel PHASE w = (p - 0.5) - nominal pixel, in [-0.5, 0.5], and for an integer radius r the function is a
// step function with few steps: with K = 16 / r table cells per pixel, every table index floor(|(d - w) * K|) is constant
// on each of the K open intervals between the points m / K - 0.5, and takes its own values on those K + 1 points.
// So the pre-pass classifies the phase on each axis once per sample (2K + 1 classes: one LUT fetch and two comparisons
// against float bounds the host found by bisection over the very float expression the CPU path evaluates), and the
// record carries the shared-memory address of the class pair's block of PRECOMBINED weights
//        block[column d][row i] = table[ify(i)][ifx(d)],  2h live rows per column, zero for an unreached column.
// The gather then costs, per (sample, column): LDS.128 record + LDS.128 weights (two at h = 4) and per live row
// 3 FMUL + 2 FADD2 — no index arithmetic, no per-tap load, no dead row.  Which 2h of the 2h+1 window rows are live
// ("up": rows 0..2h-1, phase < 0; "down": rows 1..2h, phase > 0, stored mirrored) and whether an outermost column is
// reached are properties of the class; the pre-pass reduces them over the strip to one bit per sample index, and the
// gather walks runs of sample indices that are uniformly "up" or "down" with branch-free bodies (stratified streams:
// two runs).  Anything else — a phase exactly 0 (2h+1 live rows), a sample whose floor(pd + r) rounds across an
// integer at a power-of-two coordinate, a stream that is not stratified — goes through a per-lane general body, and a
// sample no class describes through a slow path that evaluates the CPU expressions directly.  Nothing is approximated.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <type_traits>
#include <vector>

#include "common.cuh"

namespace pb {

// ---- table geometry (host and device) ------------------------------------------------------------------------------

// Two LUTs per axis, indexed by a phase cell.
//   fast    cell = floor(K * (w + 0.5)), entry {point, offset(interval), offset(point)}: valid when the phase is a multiple
//           of 2^-22 (any sample with |pd| >= 2): every d - w is then exact and the classes are the ideal ones —
//           the single float `point` and the open interval above it.
//   careful cell = nearest point, entry {glo, ghi, point, -, offset(interval below), offset(point), offset(interval above)}:
//           for the pixels next to the origin, where a finer-grained phase within a few ulps of a point rounds
//           differently in d - w for different d.  [glo, ghi] is the range around the point outside of which the
//           neighbouring interval's profile holds; inside it only the point itself is a class.
constexpr int CT_LUT_ENTRIES = 32;                 // fast cells; entries above K are empty
constexpr int CT_LUT_BYTES = CT_LUT_ENTRIES * 16;
constexpr int CT_CARE_ENTRIES = 16;
constexpr int CT_CARE_BYTES = CT_CARE_ENTRIES * 32;
constexpr int CT_LUT_Y = 0, CT_LUT_X = CT_LUT_BYTES;
constexpr int CT_CARE_Y = 2 * CT_LUT_BYTES, CT_CARE_X = CT_CARE_Y + CT_CARE_BYTES;
constexpr int CT_Q_OFFSET = CT_CARE_X + CT_CARE_BYTES;
// an offset no class has: the sample takes the slow path (x and y markers add without cancelling)
constexpr unsigned CT_SLOW_X = 0x40000000u, CT_SLOW_Y = 0x80000000u;
constexpr int CT_ZERO = 16;  // "not reached" in a profile
// flag bits in the low nibble of a LUT offset (block addresses are multiples of 16)
// CF_UP: the class can be gathered as "up" (window rows 0..2h-1 in storage order), CF_DOWN: as "down" (rows 1..2h, the
// mirrored block in reverse).  A phase of exactly 0 reaches all 2h+1 rows and carries both: its block is its own mirror
// image, so it runs with whichever kind its sample index has in the row and gets one more tap — on window row 2h after
// an "up" run, on row 0 after a "down" run (CF_BOTH).  CF_SLOW carries neither: a classless sample makes its sample
// index neither "up" nor "down" in the per-index masks.
constexpr unsigned CF_UP = 1, CF_DOWN = 2, CF_BOTH = 3, CF_LEFT = 4, CF_RIGHT = 8, CF_SLOW = 12;

template <int H> struct ClassCfg {
    static constexpr int ROWS = 2 * H + 1;
    static constexpr int LIVE = 2 * H;
    static constexpr int K = 16 / H;
    static constexpr int NX = 2 * K + 1;
    static constexpr int COLB = LIVE * 4;    // bytes per column of a block
    static constexpr int BLK = ROWS * COLB;  // bytes per class-pair block
};

struct ClassGeom {
    int H, K, NX, NXP, COLB, BLK, ROWP, EOFF, BYTES;
};

static bool class_geom(float rx, float ry, ClassGeom *g) {
    if (rx != ry || (rx != 2.f && rx != 4.f)) return false;
    g->H = (int)rx;
    g->K = 16 / g->H;
    g->NX = 2 * g->K + 1;
    g->COLB = 2 * g->H * 4;
    g->BLK = (2 * g->H + 1) * g->COLB;
    // Row pitch in blocks.  The lanes of a warp hold sample s of neighbouring pixels; for a stratified stream they fall into
    // a few neighbouring classes on each axis.  A pitch whose 16-byte count is 2 or 6 modulo 8 puts the four interval
    // classes of such a neighbourhood on four different bank groups.
    g->NXP = g->NX;
    for (int n = g->NX; n < g->NX + 8; ++n)
        if (((n * g->BLK / 16) % 8) == 2 || ((n * g->BLK / 16) % 8) == 6) { g->NXP = n; break; }
    g->ROWP = g->NXP * g->BLK;
    g->EOFF = CT_Q_OFFSET + (g->K + 1) * g->ROWP;  // last-row weights of the phase-0 class: [class x][column]
    g->BYTES = (g->EOFF + g->NX * (2 * g->H + 1) * 4 + 15) / 16 * 16;
    return true;
}

// ---- host: profiles, classes, tables -----------------------------------------------------------------------------

// One axis of add_sample for a sample of phase w: for pixel offset d in -H..H the filter-table index the CPU path
// computes — min(floor(|((x - pd) * inv_radius) * 16|), 15) with x - pd == d - w exactly, the scaling by 16 / r a power
// of two — or CT_ZERO when [ceil(pd - r), floor(pd + r)] excludes the pixel (in exact arithmetic; the kernel checks
// the one place where the float expression can differ, see class_prepass).
static void axis_profile(float w, int H, int *bins) {
    const float c16 = (1.f / (float)H) * 16.f;
    for (int d = -H; d <= H; ++d) {
        const bool reached = d == -H ? w <= 0.f : (d == H ? w >= 0.f : true);
        const float t = fabsf(((float)d - w) * c16);
        int b = t >= 16.f ? 15 : (int)floorf(t);
        if (b > 15) b = 15;
        bins[d + H] = reached ? b : CT_ZERO;
    }
}

static int32_t float_order(float f) {
    int32_t i;
    memcpy(&i, &f, 4);
    return i ^ ((i >> 31) & 0x7fffffff);
}
static float order_float(int32_t o) {
    int32_t i = o ^ ((o >> 31) & 0x7fffffff);
    float f;
    memcpy(&f, &i, 4);
    return f;
}

static float class_point(int kk, int K) { return (float)kk / (float)K - 0.5f; }
static float class_mid(int kk, int K) { return ((float)kk + 0.5f) / (float)K - 0.5f; }
// representative phase of class c (even: the point c/2, odd: the interval (c-1)/2)
static float class_rep(int c, int K) { return (c & 1) ? class_mid(c >> 1, K) : class_point(c >> 1, K); }

// Builds the blob the kernel copies into shared memory: two LUTs and the weight blocks.  Returns false when the radius is
// not served or the float step function does not have the expected shape (then the window kernel runs instead).
static bool class_tables_host(const float table[256], float rx, float ry, std::vector<unsigned char> *blob,
                              ClassGeom *geom) {
    ClassGeom g;
    if (!class_geom(rx, ry, &g)) return false;
    const int H = g.H, K = g.K, ROWS = 2 * H + 1;
    std::vector<int> want(ROWS), prof(ROWS);
    auto same = [&](float w) {
        axis_profile(w, H, prof.data());
        return std::equal(prof.begin(), prof.end(), want.begin());
    };
    // first float from `good` towards `bad` (both in float order) whose profile is still `want`
    auto boundary = [&](float bad, float good) {
        int32_t a = float_order(bad), b = float_order(good);
        const int32_t step = a < b ? 1 : -1;
        while ((b - a) * step > 1) {
            const int32_t m = a + (b - a) / 2;
            if (same(order_float(m))) b = m; else a = m;
        }
        return order_float(b);
    };
    // interval kk = [ilo, ihi]: every float in it has the profile of the interval's midpoint
    std::vector<float> ilo(K), ihi(K);
    const float grain = 1.f / 4194304.f;  // 2^-22
    for (int kk = 0; kk < K; ++kk) {
        axis_profile(class_mid(kk, K), H, want.data());
        ilo[kk] = boundary(class_point(kk, K), class_mid(kk, K));
        ihi[kk] = boundary(class_point(kk + 1, K), class_mid(kk, K));
        // the fast LUT's premise: on the 2^-22 grid the interval is everything strictly between its two points
        if (!(ilo[kk] <= class_point(kk, K) + grain && ihi[kk] >= class_point(kk + 1, K) - grain)) return false;
    }
    // "down" classes are stored as their mirror image: profile(w) reversed must be profile(-w)
    for (int c = 0; c <= 2 * K; ++c) {
        std::vector<int> a(ROWS), b(ROWS);
        axis_profile(class_rep(c, K), H, a.data());
        axis_profile(class_rep(2 * K - c, K), H, b.data());
        std::reverse(b.begin(), b.end());
        if (a != b) return false;
    }
    blob->assign(g.BYTES, 0);
    auto offx = [&](int c) { return (uint32_t)(c * g.BLK) | (c <= K ? CF_LEFT : 0u) | (c >= K ? CF_RIGHT : 0u); };
    auto offy = [&](int c) {
        const int row = c <= K ? c : 2 * K - c;
        return (uint32_t)(CT_Q_OFFSET + row * g.ROWP) | (c <= K ? CF_UP : 0u) | (c >= K ? CF_DOWN : 0u);
    };
    struct Fast { float point; uint32_t off_interval, off_point, pad; };
    struct Care { float glo, ghi, point, pad0; uint32_t off_below, off_point, off_above, pad1; };
    for (int axis = 0; axis < 2; ++axis) {
        const uint32_t slow = axis ? CT_SLOW_X : CT_SLOW_Y;
        auto off = [&](int c) { return axis ? offx(c) : offy(c); };
        Fast *fast = reinterpret_cast<Fast *>(blob->data() + (axis ? CT_LUT_X : CT_LUT_Y));
        for (int kk = 0; kk < CT_LUT_ENTRIES; ++kk) {
            Fast e = {NAN, slow, slow, 0u};
            if (kk <= K) {
                e.point = class_point(kk, K);
                e.off_point = off(2 * kk);
                if (kk < K) e.off_interval = off(2 * kk + 1);
            }
            fast[kk] = e;
        }
        Care *care = reinterpret_cast<Care *>(blob->data() + (axis ? CT_CARE_X : CT_CARE_Y));
        for (int m = 0; m < CT_CARE_ENTRIES; ++m) {
            Care e = {INFINITY, -INFINITY, NAN, 0.f, slow, slow, slow, 0u};  // w < glo: slow
            if (m <= K) {
                e.point = class_point(m, K);
                e.off_point = off(2 * m);
                // below the point: interval m-1 up to its last float; above: interval m from its first float
                e.glo = m > 0 ? order_float(float_order(ihi[m - 1]) + 1) : e.point;
                e.ghi = m < K ? order_float(float_order(ilo[m]) - 1) : e.point;
                if (m > 0) e.off_below = off(2 * m - 1);
                if (m < K) e.off_above = off(2 * m + 1);
            }
            care[m] = e;
        }
    }
    std::vector<int> px(ROWS), py(ROWS);
    for (int cy = 0; cy <= K; ++cy) {
        axis_profile(class_rep(cy, K), H, py.data());
        for (int cx = 0; cx <= 2 * K; ++cx) {
            axis_profile(class_rep(cx, K), H, px.data());
            for (int v = 0; v < ROWS; ++v) {
                auto weight = [&](int i) {
                    return (px[v] == CT_ZERO || py[i] == CT_ZERO) ? 0.f : table[py[i] * 16 + px[v]];
                };
                float *col = reinterpret_cast<float *>(blob->data() + CT_Q_OFFSET + cy * g.ROWP + cx * g.BLK + v * g.COLB);
                for (int i = 0; i < 2 * H; ++i) col[i] = weight(i);
                if (cy == K) reinterpret_cast<float *>(blob->data() + g.EOFF)[cx * ROWS + v] = weight(2 * H);
            }
        }
    }
    *geom = g;
    return true;
}

int class_tables_create(PbrtFilm *f) {
    std::vector<unsigned char> blob;
    ClassGeom g;
    f->class_bytes = 0;
    if (!class_tables_host(f->table, f->radius[0], f->radius[1], &blob, &g)) return PBRT_OK;
    PB_CUDA(cudaMalloc(&f->d_class, blob.size()));
    PB_CUDA(cudaMemcpyAsync(f->d_class, blob.data(), blob.size(), cudaMemcpyHostToDevice, ctx().stream));
    PB_CUDA(cudaStreamSynchronize(ctx().stream));  // blob is a local
    f->class_bytes = (int)blob.size();
    f->class_h = g.H;
    f->class_k = g.K;
    f->class_rowp = g.ROWP;
    return PBRT_OK;
}

// ---- device ----------------------------------------------------------------------------------------------------

typedef unsigned long long u64;

struct ClassParams {
    SplatParams S;
    const uint4 *blob;
    int blob_bytes;
    int rowp;
    int eoff;
    const int *seg_start;  // first output row of each row segment (blockIdx.y), one past the last at [gridDim.y]
    int wait_first;        // wait for the preceding grid before anything is read (see launch_class)
#ifdef PBRT_CLASS_TRACE
    unsigned long long *trace;  // per CTA {smid, start ns, end ns, rows} (tools/cta_trace.py)
#endif
};

__device__ __forceinline__ u64 cpack2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void cunpack2(u64 v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
// in place: the accumulator keeps its register pair across the unrolled loop (no copies at the back edge)
__device__ __forceinline__ void cadd2(u64 &a, u64 b) {
    asm("add.rn.f32x2 %0, %0, %1;" : "+l"(a) : "l"(b));
}
// Loads from the table blob, which is constant once the barrier after its fill has passed.  Not volatile: the address
// depends on data read after that barrier, so the compiler may schedule these freely across the unrolled samples.
__device__ __forceinline__ void lds_blob4(unsigned addr, float &a, float &b, float &c, float &d) {
    asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a), "=f"(b), "=f"(c), "=f"(d) : "r"(addr));
}
// record / flag stores by shared-window address (a generic pointer makes the compiler rebuild the window base per store)
__device__ __forceinline__ void sts_rec(unsigned addr, float a, float b, float c, unsigned d) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void sts_flag(unsigned addr, unsigned v) {
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ float lds_blob1(unsigned addr) {
    float v;
    asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}

// min(floor(|v|), 15) — the CPU path's table index (slow path only)
__device__ __forceinline__ int class_table_index(float v) {
    return __float_as_int(__fadd_rd(fminf(fabsf(v), 15.f), 8388608.f)) & 0xF;
}

// Window accumulators: ROWS x {r, g, b, weight}.  LAYOUT 0: two packed pairs (r,g) (b,w) per row, sums by FADD2 — the
// (b*w, w) addend pair costs a MOV per tap since w arrives in a register of its own; 1: (r,g) packed, b and w scalar;
// 2: four scalars.  The fused variant (PBRT_SPLAT_FMA) always uses 2.  Indices are compile-time after unrolling.
#ifndef PBRT_CLASS_ACC
#define PBRT_CLASS_ACC 3
#endif
// packed add on two scalar accumulators that the register allocator is asked to keep as an aligned pair
__device__ __forceinline__ void cadd2s(float &a0, float &a1, float b0, float b1) {
    asm("{\n\t.reg .b64 pa, pb;\n\tmov.b64 pa, {%0, %1};\n\tmov.b64 pb, {%2, %3};\n\tadd.rn.f32x2 pa, pa, pb;\n\tmov.b64 {%0, %1}, pa;\n\t}"
        : "+f"(a0), "+f"(a1) : "f"(b0), "f"(b1));
}
template <int ROWS, int LAYOUT>
struct ClassAcc;
template <int ROWS>
struct ClassAcc<ROWS, 3> {
    float r[ROWS], g[ROWS], b[ROWS], w[ROWS];
    __device__ __forceinline__ void clear(const int j) { r[j] = g[j] = b[j] = w[j] = 0.f; }
    __device__ __forceinline__ void move(const int to, const int from) { r[to] = r[from]; g[to] = g[from]; b[to] = b[from]; w[to] = w[from]; }
    __device__ __forceinline__ void get(const int j, float &R, float &G, float &B, float &Wt) const { R = r[j]; G = g[j]; B = b[j]; Wt = w[j]; }
    __device__ __forceinline__ void tap(const int j, const float cr, const float cg, const float cb, const float wt) {
        cadd2s(r[j], g[j], cr * wt, cg * wt);
        cadd2s(b[j], w[j], cb * wt, wt);
    }
    __device__ __forceinline__ void tap_fma(const int j, const float cr, const float cg, const float cb, const float wt) {
        r[j] = __fmaf_rn(cr, wt, r[j]); g[j] = __fmaf_rn(cg, wt, g[j]); b[j] = __fmaf_rn(cb, wt, b[j]);
        w[j] += wt;
    }
};
template <int ROWS>
struct ClassAcc<ROWS, 4> {  // (r, g) packed, b and weight scalar
    float r[ROWS], g[ROWS], b[ROWS], w[ROWS];
    __device__ __forceinline__ void clear(const int j) { r[j] = g[j] = b[j] = w[j] = 0.f; }
    __device__ __forceinline__ void move(const int to, const int from) { r[to] = r[from]; g[to] = g[from]; b[to] = b[from]; w[to] = w[from]; }
    __device__ __forceinline__ void get(const int j, float &R, float &G, float &B, float &Wt) const { R = r[j]; G = g[j]; B = b[j]; Wt = w[j]; }
    __device__ __forceinline__ void tap(const int j, const float cr, const float cg, const float cb, const float wt) {
        cadd2s(r[j], g[j], cr * wt, cg * wt);
        b[j] += cb * wt;
        w[j] += wt;
    }
    __device__ __forceinline__ void tap_fma(const int j, const float cr, const float cg, const float cb, const float wt) {
        r[j] = __fmaf_rn(cr, wt, r[j]); g[j] = __fmaf_rn(cg, wt, g[j]); b[j] = __fmaf_rn(cb, wt, b[j]);
        w[j] += wt;
    }
};
template <int ROWS, int LAYOUT>
struct ClassAcc {
    u64 rg[LAYOUT <= 1 ? ROWS : 1];
    u64 bw[LAYOUT == 0 ? ROWS : 1];
    float r[LAYOUT == 2 ? ROWS : 1], g[LAYOUT == 2 ? ROWS : 1];
    float b[LAYOUT >= 1 ? ROWS : 1], w[LAYOUT >= 1 ? ROWS : 1];
    __device__ __forceinline__ void clear(const int j) {
        if (LAYOUT <= 1) rg[j] = 0ull; else r[j] = g[j] = 0.f;
        if (LAYOUT == 0) bw[j] = 0ull; else b[j] = w[j] = 0.f;
    }
    __device__ __forceinline__ void move(const int to, const int from) {
        if (LAYOUT <= 1) rg[to] = rg[from]; else { r[to] = r[from]; g[to] = g[from]; }
        if (LAYOUT == 0) bw[to] = bw[from]; else { b[to] = b[from]; w[to] = w[from]; }
    }
    __device__ __forceinline__ void get(const int j, float &R, float &G, float &B, float &Wt) const {
        if (LAYOUT <= 1) cunpack2(rg[j], R, G); else { R = r[j]; G = g[j]; }
        if (LAYOUT == 0) cunpack2(bw[j], B, Wt); else { B = b[j]; Wt = w[j]; }
    }
    // exact: products by scalar multiplies (rounded like the CPU's), then added — ptxas would fuse a packed multiply
    // feeding a packed add into one FFMA2 even under --fmad=false
    __device__ __forceinline__ void tap(const int j, const float cr, const float cg, const float cb, const float wt) {
        if (LAYOUT <= 1) cadd2(rg[j], cpack2(cr * wt, cg * wt));
        else { r[j] += cr * wt; g[j] += cg * wt; }
        if (LAYOUT == 0) cadd2(bw[j], cpack2(cb * wt, wt));
        else { b[j] += cb * wt; w[j] += wt; }
    }
    // same order, one rounding per accumulate (LAYOUT 2 only)
    __device__ __forceinline__ void tap_fma(const int j, const float cr, const float cg, const float cb, const float wt) {
        r[j] = __fmaf_rn(cr, wt, r[j]); g[j] = __fmaf_rn(cg, wt, g[j]); b[j] = __fmaf_rn(cb, wt, b[j]);
        w[j] += wt;
    }
};

// Slow path: a sample no class describes.  Evaluates the CPU path's expressions for one (sample, output column) from
// scratch — out of line and self-contained, so that nothing it needs stays live in the gather loops.  Returns the weight
// of each window row (0 where the sample does not reach: adding (L * 0, 0) changes nothing) and reports a sample outside
// its nominal pixel.
template <int ROWS>
struct SlowWeights { float w[ROWS]; int out_of_pixel; };

template <int H>
__device__ __noinline__ SlowWeights<2 * H + 1> class_slow_weights(const float2 *xy, const float *table, size_t index, int nx,
                                                                  int ny, int x, float rx, float ry) {
    SlowWeights<2 * H + 1> r;
    const float2 p = __ldg(&xy[index]);
    const float pdx = p.x - 0.5f, pdy = p.y - 0.5f, fx = (float)x, fny = (float)ny;
    const float c16 = (1.f / (float)H) * 16.f;
    r.out_of_pixel = !(fabsf(pdx - (float)nx) <= 0.5f && fabsf(pdy - fny) <= 0.5f);
    const bool reach_x = fx >= pdx - rx && fx <= pdx + rx;
    const int ix = class_table_index((fx - pdx) * c16);
#pragma unroll
    for (int j = 0; j < 2 * H + 1; ++j) {
        const float fy = fny + (float)(j - H);
        const bool reach = reach_x && fy >= pdy - ry && fy <= pdy + ry;
        const int iy = class_table_index((fy - pdy) * c16);
        r.w[j] = reach ? __ldg(&table[iy * 16 + ix]) : 0.f;
    }
    return r;
}

template <int H, int TW>
struct ClassSmem {
    static constexpr int NPX = TW + 2 * H;
    static constexpr int MASK_BYTES = 128;  // 2 parities x 8 masks of up to 64 bits
    __host__ __device__ static int pitch(int spp) { return spp | 1; }
    __host__ __device__ static size_t bytes(int blob_bytes, int spp) {
        return (size_t)blob_bytes + MASK_BYTES + ((size_t)NPX * pitch(spp) * 17 + 15) / 16 * 16;
    }
};

#ifndef PBRT_CLASS_PREPASS_BATCH
#define PBRT_CLASS_PREPASS_BATCH 4
#endif
#ifndef PBRT_CLASS_UNROLL
#define PBRT_CLASS_UNROLL 4
#endif
#ifndef PBRT_CLASS_UNROLL4
#define PBRT_CLASS_UNROLL4 2  // 4: +8 % instruction-cache misses on the 8K / Lanczos config (stall_no_inst 14.6 % -> measured 3.42e10 vs 3.71e10)
#endif
constexpr int kClassUnroll = PBRT_CLASS_UNROLL;    // samples per trip of a run's loop, h = 2
constexpr int kClassUnroll4 = PBRT_CLASS_UNROLL4;  // h = 4 (twice the taps per sample)

#ifndef PBRT_CLASS_RESIDENT_THREADS
#define PBRT_CLASS_RESIDENT_THREADS 512  // 4 CTAs of 128 threads per SM: 128 registers per thread
#endif
// TILES: the batched form (pbrt_film_add_samples_tiles): blockIdx.z selects a renderer tile, whose bounds and streams
// replace sb / tb / xy / rgbw; its finished pixels go to the tile's own FilmTilePixel buffer instead of the film.
// WIDE: 64-bit per-index masks (33..64 samples per pixel); the 32-bit form is the tuned one and stays as it is.
__device__ __forceinline__ int mask_popc(unsigned m) { return __popc(m); }
__device__ __forceinline__ int mask_popc(u64 m) { return __popcll(m); }
__device__ __forceinline__ int mask_ffs(unsigned m) { return __ffs((int)m); }
__device__ __forceinline__ int mask_ffs(u64 m) { return __ffsll((long long)m); }

template <int H, int TW, bool FMA, bool TILES = false, bool WIDE = false>
__global__ void __launch_bounds__(TW, TILES ? 1 : PBRT_CLASS_RESIDENT_THREADS / TW) splat_class_kernel(ClassParams CP) {
    typedef typename std::conditional<WIDE, u64, unsigned>::type M;  // one bit per sample index of a pixel
    constexpr int MB = WIDE ? 64 : 32;
    typedef ClassCfg<H> C;
    constexpr int ROWS = C::ROWS, LIVE = C::LIVE, K = C::K, COLB = C::COLB, BLK = C::BLK;
    constexpr int NPX = ClassSmem<H, TW>::NPX;
    const SplatParams &P = CP.S;
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x;
    Bounds sb = P.sb, tb = P.tb;
    const float2 *sxy = CP.S.xy;
    const float4 *srgbw = CP.S.rgbw;
    float4 *tile_out = nullptr;
    if (TILES) {
        const SplatTile t = P.tiles[blockIdx.z];
        sb = t.sb;
        tb = t.tb;
        sxy += t.sample_offset;
        srgbw += t.sample_offset;
        tile_out = P.tile_out + t.pixel_offset;
        if (tb.x1 <= tb.x0 || tb.y1 <= tb.y0 || sb.x1 <= sb.x0 || sb.y1 <= sb.y0) return;
        if (tb.x0 + (int)blockIdx.x * TW >= tb.x1) return;  // the grid is sized for the widest tile
    }
    // Programmatic dependent launch: the next splat launch of the stream may fill the SM slots this grid's early finishers
    // leave (the grid is one wave, its tail would idle otherwise).  Nothing a splat launch reads before its first flush
    // is written by the splat launch ahead of it; the film is, so its first access waits for that grid (film_wait).
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    bool film_ready = false;
    if (CP.wait_first) { asm volatile("griddepcontrol.wait;" ::: "memory"); film_ready = true; }
#ifdef PBRT_CLASS_TRACE
    unsigned long long trace_t0 = 0;
    if (tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(trace_t0));
#endif
    const int spp = P.spp;
    const int pitch = ClassSmem<H, TW>::pitch(spp);
    M *s_mask = reinterpret_cast<M *>(smem + CP.blob_bytes);
    float4 *s_rec = reinterpret_cast<float4 *>(smem + CP.blob_bytes + ClassSmem<H, TW>::MASK_BYTES);
    unsigned char *s_flag = reinterpret_cast<unsigned char *>(s_rec + (size_t)NPX * pitch);

    for (int i = tid; i < CP.blob_bytes / 16; i += TW) reinterpret_cast<uint4 *>(smem)[i] = CP.blob[i];
    if (tid < 16) s_mask[tid] = M(0);
    __syncthreads();
    // shared-window address of the blob, kept opaque so that it lives in a register
    unsigned sbase = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("mov.u32 %0, %0;" : "+r"(sbase));
    const unsigned a_rec = sbase + CP.blob_bytes + ClassSmem<H, TW>::MASK_BYTES;
    const unsigned a_flag = a_rec + (unsigned)(NPX * pitch) * 16u;

    const int cx0 = tb.x0 + blockIdx.x * TW;              // first output column of the strip
    const int cy0 = TILES ? tb.y0 : CP.seg_start[blockIdx.y];  // first output row
    const int cy1 = TILES ? tb.y1 : CP.seg_start[blockIdx.y + 1];
    const int x = cx0 + tid;
    const bool col_ok = x < tb.x1;
    const int W = sb.x1 - sb.x0;
    const bool clamp_on = P.max_lum < __int_as_float(0x7f800000);
    const float rH = (float)H;                 // == P.rx == P.ry

    // staged nominal pixels of a row: [sx0, sx1); local index = nx - (cx0 - H)
    const int sx0 = max(cx0 - H, sb.x0), sx1 = min(cx0 + TW + H, sb.x1);
    const int nstaged = max(sx1 - sx0, 0) * spp;
    // element e = tid + k*TW of the staged run is sample `sidx` of staged pixel q: advance (q, sidx) without dividing
    const int q0 = tid / spp, r0 = tid - q0 * spp;
    const int dq = TW / spp, dr = TW - dq * spp;
    const int pl_base = sx0 - (cx0 - H);
    const int slot_step = dq * pitch + dr;
    const float fdq = (float)dq;
    // a thread keeps its sample index for the whole row when spp divides the strip width: the per-index class masks
    // are then reduced in registers; otherwise (and above 32 spp) every sample takes the general body
    const bool mask_mode = dr == 0 && spp <= MB;
    const M sppmask = spp >= MB ? ~M(0) : ((M(1) << spp) - M(1));
    constexpr int U = PBRT_CLASS_PREPASS_BATCH;
    // Which of this thread's batches need more than the plain classification.  The pixels of batch k are the same in
    // every row, so both are found once, bit k per batch:
    //   near_x:  a nominal pixel n < 3 — a phase finer than 2^-22 near the origin, p - 0.5 rounded at negative
    //            coordinates: the careful classification;
    //   check_x: a power of two in (n, n + H] — there floor(pd + r) can round up across the integer (see `one`):
    //            the plain classification plus that test.
    // Rows likewise (near_y, check_y below).  Keeping these batches cheap matters: the grid is one wave, so the
    // strips and row segments that hold such pixels set the kernel's time.
    unsigned near_x = 0xffffffffu, check_x = 0u;
    if (dr == 0 && nstaged <= 32 * U * TW) {
        near_x = 0u;
        for (int k = 0; tid + k * U * TW < nstaged; ++k)
            for (int u = 0; u < U; ++u) {
                const int n = sx0 + q0 + (k * U + u) * dq;
                if (n < 3) near_x |= 1u << k;
                else if (__clz(n) != __clz(n + H)) check_x |= 1u << k;
            }
    }

    ClassAcc<ROWS, FMA ? 2 : PBRT_CLASS_ACC> acc;
#pragma unroll
    for (int j = 0; j < ROWS; ++j) acc.clear(j);
    unsigned errbits = 0;
    float vmax = 0.f;  // largest |phase| this thread has seen
    int parity = 0;

    auto tap = [&](const int j, const float cr, const float cg, const float cb, const float w) {
        if (FMA) acc.tap_fma(j, cr, cg, cb, w);
        else acc.tap(j, cr, cg, cb, w);
    };

    for (int ny = cy0 - H; ny < cy1 + H; ++ny) {
        const bool row_has_samples = ny >= sb.y0 && ny < sb.y1 && nstaged > 0;
        // the film pixel this row's flush adds to: loaded ahead of the gather, which hides the latency
        const int yo = ny - H;
        const bool flush = col_ok && yo >= cy0 && yo < cy1;
        const size_t fo = (size_t)(yo - P.owned.y0) * (P.owned.x1 - P.owned.x0) + (x - P.owned.x0);
        float4 px = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row_has_samples) {
            __syncthreads();  // previous row fully consumed
            // ---------------- pre-pass: one thread per sample of the row ----------------
            const size_t row_base = ((size_t)(ny - sb.y0) * W + (sx0 - sb.x0)) * (size_t)spp;
            const float2 *gxy = sxy + row_base;
            const float4 *grgbw = srgbw + row_base;
            const float fny = (float)ny, fnyH = fny + rH, fnyh = fny + 0.5f;
            const bool near_y = ny < 3, check_y = !near_y && __clz(ny) != __clz(ny + H);
            int sidx = r0;
            int slot = (pl_base + q0) * pitch + r0;
            float fnxh = (float)(sx0 + q0) + 0.5f;  // nominal pixel of the thread's next sample, plus one half
            unsigned andf = 15u, orf = 0u;
            unsigned near_bits = near_x, check_bits = check_x;
            // max_sample_luminance: off (infinite) in every BASELINE config
            auto clamp_lum = [&](float4 &Lv) {
                const float ly = luminance(Lv.x, Lv.y, Lv.z);
                if (ly > P.max_lum) {
                    const float sc = P.max_lum / ly;
                    Lv.x *= sc; Lv.y *= sc; Lv.z *= sc;
                }
            };
            // One sample: phase, class, record.  TIER 0: the classes are the ideal ones; 1: also tests whether
            // floor(pd + r) rounds up; 2 ("careful"): the pixels next to the origin, whose phase can be finer than
            // 2^-22 (bounds found on the host decide), and negative coordinates.
            auto one = [&](const float2 pv, const float4 Lv, const int slot_, const float fnxh_, auto tier_tag) {
                constexpr int TIER = decltype(tier_tag)::value;
                constexpr bool CAREFUL = TIER == 2;
                const float cr = Lv.x * Lv.w, cg = Lv.y * Lv.w, cb = Lv.z * Lv.w;
                // phase = pd - n with pd = p - 0.5 as the CPU path rounds it: exact (Sterbenz) for a sample inside its
                // nominal pixel.  For n >= 1 pd itself is exact, so the phase is p - (n + 0.5) in one subtraction.
                const float pdx = pv.x - 0.5f, pdy = pv.y - 0.5f, fnx = fnxh_ - 0.5f;
                const float wx = CAREFUL ? pdx - fnx : pv.x - fnxh_;
                const float wy = CAREFUL ? pdy - fny : pv.y - fnyh;
                // contract: the sample lies in its nominal pixel (a NaN phase reaches the slow path instead)
                vmax = fmaxf(vmax, fmaxf(fabsf(wx), fabsf(wy)));
                unsigned offx, offy;
                if (!CAREFUL) {
                    auto classify = [&](const float w, const unsigned lut) {
                        float t = __fmaf_rn(w, (float)K, 0.5f * (float)K);
                        t = fminf(fabsf(t), (float)(CT_LUT_ENTRIES - 1));
                        const unsigned cell = (unsigned)__float_as_int(__fadd_rd(t, 8388608.f));
                        float point, oi, op, pad;
                        lds_blob4(__dp4a(cell, 16u, lut), point, oi, op, pad);
                        return __float_as_uint(w == point ? op : oi);
                    };
                    offx = classify(wx, sbase + CT_LUT_X);
                    offy = classify(wy, sbase + CT_LUT_Y);
                } else {
                    auto classify = [&](const float w, const unsigned lut, const unsigned slow) {
                        float t = __fmaf_rn(w, (float)K, 0.5f * (float)K + 0.5f);
                        t = fminf(fabsf(t), (float)(CT_CARE_ENTRIES - 1));
                        const unsigned cell = (unsigned)__float_as_int(__fadd_rd(t, 8388608.f));
                        const unsigned e = __dp4a(cell, 32u, lut);
                        float glo, ghi, point, pad0, ob, op, oa, pad1;
                        lds_blob4(e, glo, ghi, point, pad0);
                        lds_blob4(e + 16, ob, op, oa, pad1);
                        return w < glo ? __float_as_uint(ob)
                                       : (w > ghi ? __float_as_uint(oa) : (w == point ? __float_as_uint(op) : slow));
                    };
                    offx = classify(wx, sbase + CT_CARE_X, CT_SLOW_X);
                    offy = classify(wy, sbase + CT_CARE_Y, CT_SLOW_Y);
                }
                // The classes take "pixel n + H is reached" to mean w >= 0.  The CPU path asks whether
                // n + H <= floor(pd + r) in floats, and pd + r can round up to the integer when the sum crosses a
                // power of two: such a sample is none of the classes.  Only pixels with a power of two in
                // (n, n + H] can do that (elsewhere pd + r is exact): tiers 1 and 2.  (ceil(pd - r) has no such
                // case: the difference is exact wherever the result is a pixel coordinate >= 0.)
                const unsigned sum = offx + offy;
                bool ok = sum < CT_SLOW_X;
                if (TIER >= 1) ok = ok && !(wx < 0.f && fnx + rH <= pdx + rH) && !(wy < 0.f && fnyH <= pdy + rH);
                const unsigned fl = ok ? (sum & 15u) : CF_SLOW;
                sts_rec(a_rec + 16u * (unsigned)slot_, cr, cg, cb, sum & ~15u);
                sts_flag(a_flag + (unsigned)slot_, fl);
                andf &= fl;
                orf |= fl;
            };
            typedef std::integral_constant<int, 0> T_PLAIN;
            typedef std::integral_constant<int, 1> T_CHECK;
            typedef std::integral_constant<int, 2> T_CAREFUL;
            // the batch of sample e0 (this thread's), e0 + TW, ...: loads, then classification
            auto load_batch = [&](float2 *p, float4 *L, const int e0) {
                const float2 *lxy = gxy + e0;
                const float4 *lrgbw = grgbw + e0;
                if (e0 + (U - 1) * TW < nstaged) {
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        p[u] = ldg_stream(lxy + u * TW);
                        L[u] = ldg_stream(lrgbw + u * TW);
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        if (e0 + u * TW < nstaged) {
                            p[u] = ldg_stream(lxy + u * TW);
                            L[u] = ldg_stream(lrgbw + u * TW);
                        }
                    }
                }
            };
            auto process_batch = [&](float2 *p, float4 *L, const int e0) {
                const bool full = e0 + (U - 1) * TW < nstaged;
                if (clamp_on) {
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        if (full || e0 + u * TW < nstaged) clamp_lum(L[u]);
                }
                const bool careful = near_y || (near_bits & 1u) || dr != 0;
                const bool check = check_y || (check_bits & 1u);
                near_bits >>= 1;
                check_bits >>= 1;
                if (careful) {
                    // pixels next to the origin, or spp not dividing the strip width: bounds-checked, with the careful
                    // classification (valid for every phase, a dozen instructions longer)
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        if (e0 + u * TW < nstaged) {
                            // only the samples of those pixels: the rest of the batch takes the checked plain form
                            if (near_y || dr != 0 || fnxh < 3.f) one(p[u], L[u], slot, fnxh, T_CAREFUL{});
                            else one(p[u], L[u], slot, fnxh, T_CHECK{});
                        }
                        slot += slot_step;
                        fnxh += fdq;
                        sidx += dr;
                        if (sidx >= spp) { sidx -= spp; slot += pitch - spp; fnxh += 1.f; }
                    }
                } else if (full && !check) {
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        one(p[u], L[u], slot, fnxh, T_PLAIN{});
                        slot += slot_step;
                        fnxh += fdq;
                    }
                } else {
                    // a pixel below a power of two, or the tail of the row (bounds-checked)
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        if (full || e0 + u * TW < nstaged) one(p[u], L[u], slot, fnxh, T_CHECK{});
                        slot += slot_step;
                        fnxh += fdq;
                    }
                }
            };
            for (int e0 = tid; e0 < nstaged; e0 += U * TW) {
                float2 p[U];
                float4 L[U];
                load_batch(p, L, e0);
                process_batch(p, L, e0);
            }
            // per sample index: is every sample of the strip "up" / "down", does every / no sample reach the outermost
            // column on the left / right.  Words: 0 !up 1 !down 2 !left-all 3 !left-none 4 !right-all 5 !right-none
            // 6 some sample may have phase 0 on y (its extra tap follows the run it is in)  7 ... on x (it reaches both
            // outermost columns: taken per lane among the uniform samples of an edge-column visit)
            M *mk = s_mask + parity * 8;
            if (mask_mode && tid < nstaged) {
                const M bit = M(1) << r0;
                if (!(andf & CF_UP)) atomicOr(&mk[0], bit);
                if (!(andf & CF_DOWN)) atomicOr(&mk[1], bit);
                if ((orf & CF_BOTH) == CF_BOTH) atomicOr(&mk[6], bit);
                if ((orf & (CF_LEFT | CF_RIGHT)) == (CF_LEFT | CF_RIGHT)) atomicOr(&mk[7], bit);
                if (!(andf & CF_LEFT)) atomicOr(&mk[2], bit);
                if (orf & CF_LEFT) atomicOr(&mk[3], bit);
                if (!(andf & CF_RIGHT)) atomicOr(&mk[4], bit);
                if (orf & CF_RIGHT) atomicOr(&mk[5], bit);
            }
            __syncthreads();
            if (tid < 8) s_mask[(parity ^ 1) * 8 + tid] = M(0);  // last read before this row's first barrier
            // pull the next sample row of this strip into L2 while this one is gathered
#ifndef PBRT_NO_PREFETCH
            if (tid == 0 && ny + 1 < sb.y1 && ny + 1 < cy1 + H) {
                // one bulk prefetch per stream, in whole 16-byte granules inside the run
                const size_t bxy = reinterpret_cast<size_t>(gxy + (size_t)W * spp);
                const size_t nxy = (bxy + 15) & ~(size_t)15;
                const size_t nrgbw = reinterpret_cast<size_t>(grgbw + (size_t)W * spp);
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(nxy), "r"((unsigned)(bxy + (size_t)nstaged * 8 - nxy) & ~15u) : "memory");
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(nrgbw), "r"((unsigned)(nstaged * 16)) : "memory");
            }
#endif
            if (flush && !TILES) {
                if (!film_ready) { asm volatile("griddepcontrol.wait;" ::: "memory"); film_ready = true; }
                px = P.film[fo];
            }
            // ---------------- gather: this thread's column against the row ----------------
#ifdef PBRT_CLASS_TIMING_NO_GATHER  // timing probe (results wrong): pre-pass and flush only
            if (col_ok && ny == cy0 - H) {
#else
            if (col_ok) {
#endif
                const M m_up = mask_mode ? ~mk[0] & sppmask : M(0), m_down = mask_mode ? ~mk[1] & sppmask : M(0);
                const M m_lall = ~mk[2] & sppmask, m_lnone = mask_mode ? ~mk[3] & sppmask : M(0);
                const M m_rall = ~mk[4] & sppmask, m_rnone = mask_mode ? ~mk[5] & sppmask : M(0);
                const M m_both = mk[6] & sppmask, m_xboth = mk[7] & sppmask;
                // A "simple" row (every stratified row without a phase-0 or classless sample): sample indices 0..n_up-1 are
                // "up" everywhere in the strip, the rest "down", and each index either always or never reaches an
                // outermost column.  Its visits are two loops with no per-sample decisions.
                const int n_up = mask_popc(m_up);
                const bool simple_rows = mask_mode && m_up == ((M(1) << n_up) - M(1)) && (m_up | m_down) == sppmask;
                // outermost columns: every index reaches its left one throughout, or its right one throughout (a sample on
                // x phase 0 reaches both and fits either)
                const bool simple_edges = (m_lall | m_rall) == sppmask;
                // Halo rows (sample rows above / below the output rows of this CTA) reach only a few owned rows: window
                // rows j >= H + t for the t-th row above, j <= H - b for the b-th row below.  Computing more is harmless
                // (rows that are not owned are never flushed), so this only selects cheaper bodies.
                const int band = ny < cy0 ? cy0 - ny : (ny >= cy1 ? -(ny - cy1 + 1) : 0);
                // Columns are visited left to right: the accumulation order of the CPU path.  The visit of nominal pixel
                // x - H (d = -H) sees only samples that reach their rightmost column, x + H only their leftmost.
#pragma unroll 1
                for (int d = -H; d <= H; ++d) {
                    const int nx = x + d;
                    if (nx < sb.x0 || nx >= sb.x1) continue;
                    const int pl = nx - (cx0 - H);
                    const float4 *pa = s_rec + pl * pitch;
                    const unsigned char *pf = s_flag + pl * pitch;
                    // shared-window address of this column inside the block at offset 0
                    const unsigned qcol = sbase + (unsigned)((H - d) * COLB);
                    const bool interior = d != -H && d != H;
                    M run_up = m_up, run_down = m_down, live = sppmask;
                    if (d == -H) { run_up &= m_rall; run_down &= m_rall; live &= ~m_rnone; }
                    if (d == H) { run_up &= m_lall; run_down &= m_lall; live &= ~m_lnone; }
                    // taps I0..I1-1 of one sample against this column: DOWN = false: tap i is window row i, weights in
                    // order; true: window row i + 1, the mirrored class's weights in reverse
                    auto body = [&](const float4 a, auto down_tag, auto i0_tag, auto i1_tag) {
                        constexpr bool DOWN = decltype(down_tag)::value;
                        constexpr int I0 = decltype(i0_tag)::value, I1 = decltype(i1_tag)::value;
                        const unsigned waddr = qcol + __float_as_uint(a.w);
                        float w[LIVE];
                        if (!FMA) {
                            // one 4-byte load per weight: each lands next to its b * w product (the (b * w, w) addend pair)
                            // without a copy — a 16-byte load pins four weights to one aligned quad and costs a MOV each
#pragma unroll
                            for (int k = 0; k < LIVE; ++k) w[k] = lds_blob1(waddr + 4 * k);
                        } else {
#pragma unroll
                            for (int k4 = 0; k4 < LIVE / 4; ++k4)
                                lds_blob4(waddr + 16 * k4, w[4 * k4], w[4 * k4 + 1], w[4 * k4 + 2], w[4 * k4 + 3]);
                        }
#pragma unroll
                        for (int i = I0; i < I1; ++i) {
                            if (DOWN) tap(i + 1, a.x, a.y, a.z, w[LIVE - 1 - i]);
                            else tap(i, a.x, a.y, a.z, w[i]);
                        }
                    };
                    // the further tap of a phase-0 sample, on window row 0 or 2h: the weight of the class pair's last row (equal
                    // to its first: the phase-0 profile is symmetric)
                    auto extra_tap = [&](const float4 a, const int row) {
                        const unsigned cx = (__float_as_uint(a.w) - (unsigned)(CT_Q_OFFSET + K * CP.rowp)) / (unsigned)BLK;
                        const float we = lds_blob1(sbase + CP.eoff + (cx * ROWS + (unsigned)(H - d)) * 4u);
                        if (row == 0) tap(0, a.x, a.y, a.z, we);
                        else tap(ROWS - 1, a.x, a.y, a.z, we);
                    };
                    typedef std::integral_constant<int, 0> I_0;
                    typedef std::integral_constant<int, LIVE> I_LIVE;
                    // n consecutive samples / the samples whose bits are set, all of one kind
                    auto run = [&](const float4 *q, const int n, auto down_tag, auto i0_tag, auto i1_tag) {
                        if (decltype(i0_tag)::value >= decltype(i1_tag)::value) return;
#pragma unroll (H == 4 ? kClassUnroll4 : kClassUnroll)
                        for (int i = 0; i < n; ++i) body(q[i], down_tag, i0_tag, i1_tag);
                    };
                    // the samples whose bits are set; those of `lane` per lane: only a sample that itself reaches this
                    // outermost column (x phase 0 among samples that reach the other one)
                    const unsigned reach = d == -H ? CF_RIGHT : (d == H ? CF_LEFT : 0u);
                    auto run_bits = [&](M bits, auto down_tag, auto i0_tag, auto i1_tag) {
                        if (decltype(i0_tag)::value >= decltype(i1_tag)::value) return;
                        while (bits) {  // two at a time where neighbours are set (stratified streams: always)
                            const int s0 = mask_ffs(bits) - 1;
                            const M two = M(3) << s0;
                            if ((bits & two) == two) {
                                const float4 a0 = pa[s0], a1 = pa[s0 + 1];
                                bits &= ~two;
                                body(a0, down_tag, i0_tag, i1_tag);
                                body(a1, down_tag, i0_tag, i1_tag);
                            } else {
                                bits &= bits - M(1);
                                body(pa[s0], down_tag, i0_tag, i1_tag);
                            }
                        }
                    };
                    // the same with some indices (`lane`) taken per lane (rows that hold an x phase-0 sample)
                    auto run_bits_lane = [&](M bits, const M lane, auto down_tag, auto i0_tag, auto i1_tag) {
                        if (decltype(i0_tag)::value >= decltype(i1_tag)::value) return;
                        while (bits) {
                            const int s0 = mask_ffs(bits) - 1;
                            bits &= bits - M(1);
                            if (!((lane >> s0) & M(1)) || (pf[s0] & reach)) body(pa[s0], down_tag, i0_tag, i1_tag);
                        }
                    };
                    if (simple_rows && (interior || simple_edges)) {
                        // up-run with taps [U0, U1), then down-run with taps [D0, D1).  A phase-0 sample runs as the kind of
                        // its index; its one further tap keeps the stream's order per pixel: window row 2h (which only
                        // "down" samples touch otherwise) right after the up-run, row 0 (only "up" samples) after it too.
                        const M uni = d == -H ? m_rall : m_lall;             // edge column: every sample of the index reaches it
                        const M lane = interior ? M(0) : m_xboth & ~uni;        // ... only samples on x phase 0 do
                        const M eb_all = m_both & (interior ? sppmask : (uni | lane));
                        auto extras = [&](M eb, const int row) {
                            while (eb) {
                                const int s0 = mask_ffs(eb) - 1;
                                eb &= eb - M(1);
                                const unsigned fl = pf[s0];
                                if ((fl & CF_BOTH) == CF_BOTH && (fl & reach) == reach) extra_tap(pa[s0], row);
                            }
                        };
                        auto visit = [&](auto u0, auto u1, auto d0, auto d1) {
                            if (interior) run(pa, n_up, std::false_type{}, u0, u1);
                            else if (lane) run_bits_lane(m_up & (uni | lane), lane, std::false_type{}, u0, u1);
                            else run_bits(m_up & uni, std::false_type{}, u0, u1);
                            if (eb_all) {
                                if (band >= 0) extras(eb_all & m_up, ROWS - 1);
                                if (band <= 0) extras(eb_all & ~m_up, 0);
                            }
                            if (interior) run(pa + n_up, spp - n_up, std::true_type{}, d0, d1);
                            else if (lane) run_bits_lane(m_down & (uni | lane), lane, std::true_type{}, d0, d1);
                            else run_bits(m_down & uni, std::true_type{}, d0, d1);
                        };
                        // Halo rows: the b-th sample row above the segment reaches window rows >= H + b only — taps
                        // [H + b, LIVE) of "up", [H + b - 1, LIVE) of "down"; the b-th below, window rows <= H - b —
                        // taps [0, H - b + 1) of "up", [0, H - b) of "down".  One instantiation per band.
                        auto above = [&](auto b_tag) {
                            constexpr int B = decltype(b_tag)::value;
                            visit(std::integral_constant<int, (H + B < LIVE ? H + B : LIVE)>{}, I_LIVE{},
                                  std::integral_constant<int, H + B - 1>{}, I_LIVE{});
                        };
                        auto below = [&](auto b_tag) {
                            constexpr int B = decltype(b_tag)::value;
                            visit(I_0{}, std::integral_constant<int, H - B + 1>{}, I_0{}, std::integral_constant<int, H - B>{});
                        };
                        typedef std::integral_constant<int, 1> B_1;
                        typedef std::integral_constant<int, 2> B_2;
                        if (band == 0 || band > H || band < -H) visit(I_0{}, I_LIVE{}, I_0{}, I_LIVE{});
                        else if (band == 1) above(B_1{});
                        else if (band == 2) above(B_2{});
                        else if (band == -1) below(B_1{});
                        else if (band == -2) below(B_2{});
                        else if (H == 4 && band == 3) above(std::integral_constant<int, (H == 4 ? 3 : 1)>{});
                        else if (H == 4 && band == 4) above(std::integral_constant<int, (H == 4 ? 4 : 1)>{});
                        else if (H == 4 && band == -3) below(std::integral_constant<int, (H == 4 ? 3 : 1)>{});
                        else below(std::integral_constant<int, (H == 4 ? 4 : 1)>{});
                        continue;
                    }
                    // (from here on the uniform runs leave out the indices that may hold a phase-0 sample: `general` adds
                    // its extra tap)
                    run_up &= ~m_both;
                    run_down &= ~m_both;
                    // a sample of any kind, decided per lane
                    auto general = [&](const int s, const int c0) {
                        const float4 a = pa[s];
                        const unsigned fl = pf[s] & 15u;
                        if (fl == CF_SLOW) {
                            // no class: the CPU path's expressions for this (sample, column)
                            const size_t index = ((size_t)(ny - sb.y0) * (sb.x1 - sb.x0) + (nx - sb.x0)) * (size_t)spp + c0 + s;
                            const SlowWeights<ROWS> sw = class_slow_weights<H>(sxy, P.table, index, nx, ny, x, P.rx, P.ry);
                            if (sw.out_of_pixel) errbits |= ERRBIT_NOT_PIXEL_MAJOR;
#pragma unroll
                            for (int j = 0; j < ROWS; ++j) tap(j, a.x, a.y, a.z, sw.w[j]);
                            return;
                        }
                        if (d == -H && !(fl & CF_RIGHT)) return;
                        if (d == H && !(fl & CF_LEFT)) return;
                        if (fl & CF_UP) {
                            body(a, std::false_type{}, I_0{}, I_LIVE{});
                            if (fl & CF_DOWN) extra_tap(a, ROWS - 1);  // phase 0: the last window row as well
                        } else {
                            body(a, std::true_type{}, I_0{}, I_LIVE{});
                        }
                    };
                    // one segment: a run of consecutive indices (always, at an interior column) or scattered ones
                    auto segment = [&](const M seg, auto down_tag) {
                        const int first = mask_ffs(seg) - 1;
                        const M shifted = seg >> first;
                        if ((shifted & (shifted + M(1))) == M(0)) run(pa + first, mask_popc(seg), down_tag, I_0{}, I_LIVE{});
                        else run_bits(seg, down_tag, I_0{}, I_LIVE{});
                    };
                    // Any other row: the live samples in stream order, cut into segments of one kind.  Skipped samples do
                    // not end a segment.  Sample indices go in chunks of 32 (the masks describe the first chunk; above 32
                    // spp every sample is general).
                    for (int c0 = 0; c0 < spp; c0 += MB, pa += MB, pf += MB) {
                        M rem = c0 == 0 ? live : (spp - c0 >= MB ? ~M(0) : (M(1) << (spp - c0)) - M(1));
                        while (rem) {
                            const M low = rem & (M(0) - rem);
                            if (low & run_up) {
                                const M nb = rem & ~run_up;  // live samples that are not "up": the first one ends the segment
                                const M seg = nb ? rem & ((nb & (M(0) - nb)) - M(1)) : rem;
                                segment(seg, std::false_type{});
                                rem &= ~seg;
                            } else if (low & run_down) {
                                const M nb = rem & ~run_down;
                                const M seg = nb ? rem & ((nb & (M(0) - nb)) - M(1)) : rem;
                                segment(seg, std::true_type{});
                                rem &= ~seg;
                            } else {
                                general(mask_ffs(low) - 1, c0);
                                rem &= ~low;
                            }
                        }
                    }
                }
            }
            parity ^= 1;
        }
        // output row ny - H is complete: no later sample row reaches it
        if (flush) {
            float r, g, b, w;
            acc.get(0, r, g, b, w);
            // non-finite radiance (a contract violation) shows in the sums: 0 * x is NaN for x = inf or NaN
            const float z = r * 0.f + g * 0.f + b * 0.f + w * 0.f;
            if (z != z) errbits |= ERRBIT_NONFINITE;
            if (TILES) {  // FilmTilePixel {contrib_sum, filter_weight_sum} of this tile (film.rs:39-42)
                tile_out[(size_t)(yo - tb.y0) * (tb.x1 - tb.x0) + (x - tb.x0)] = make_float4(r, g, b, w);
            } else {
                if (!row_has_samples) {
                    if (!film_ready) { asm volatile("griddepcontrol.wait;" ::: "memory"); film_ready = true; }
                    px = P.film[fo];
                }
                float X, Y, Z;
                rgb_to_xyz(r, g, b, X, Y, Z);
                px.x += X; px.y += Y; px.z += Z; px.w += w;
                P.film[fo] = px;
            }
        }
#pragma unroll
        for (int j = 0; j + 1 < ROWS; ++j) acc.move(j, j + 1);
        acc.clear(ROWS - 1);
    }
    if (!(vmax <= 0.5f)) errbits |= ERRBIT_NOT_PIXEL_MAJOR;
    if (errbits) atomicOr(P.err, (int)errbits);
    // a CTA that never touched the film must not let the grid complete ahead of the one before it
    if (!film_ready) asm volatile("griddepcontrol.wait;" ::: "memory");
#ifdef PBRT_CLASS_TRACE
    __syncthreads();
    if (tid == 0 && CP.trace) {
        unsigned long long t1;
        unsigned smid;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        unsigned long long *t = CP.trace + 4 * (size_t)(blockIdx.y * gridDim.x + blockIdx.x);
        t[0] = smid; t[1] = trace_t0; t[2] = t1; t[3] = (unsigned long long)(cy1 - cy0);
    }
#endif
}

// ---- launch ------------------------------------------------------------------------------------------------------

static int class_env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

#ifdef PBRT_CLASS_TRACE
static unsigned long long *g_trace = nullptr;
static int g_trace_n = 0, g_trace_cols = 0;
#endif

// Row segments of the one-wave grid.  The CTAs resident on an SM do not run at the same pace: the warp scheduler favours
// the older ones, so with equal segments the first CTA of an SM finishes ~15 % ahead of the fourth and its slot then
// idles to the end of the kernel (per-CTA trace: tools/cta_trace.py, profiles/r2_cta_trace.txt).  CTAs are placed
// round-robin by linear block id, so the residency rank of a segment's CTAs is (blockIdx.y * cols + x) / SMs; segments
// of a slower rank get fewer rows: rows + c = T * w[rank], c = the cost of the 2h halo sample rows in row units.
static double g_rank_w[8] = {1.0, 0.97, 0.93, 0.885, 0.85, 0.82, 0.79, 0.76};  // measured on C2 / C3 / C5, back-to-back launches
static double g_halo_c = -1.0;

// host only: first row of each of `segs` segments of [y0, y0 + rows), one past the last at [segs]
static void class_segment_rows(int y0, int rows, int cols, int segs, int per_sm, int H, int nsm, std::vector<int> *out) {
    if (g_halo_c < 0.0) {
        g_halo_c = 1.2;
        if (const char *v = getenv("PBRT_B200_HALO_C")) g_halo_c = atof(v);
        if (const char *v = getenv("PBRT_B200_RANK_W")) {
            int i = 0;
            for (const char *q = v; *q && i < 8; ++i) {
                g_rank_w[i] = atof(q);
                while (*q && *q != ',') ++q;
                if (*q == ',') ++q;
            }
            for (; i < 8 && i > 0; ++i) g_rank_w[i] = g_rank_w[i - 1];
        }
    }
    std::vector<double> ws(segs);
    double wsum = 0.0;
    for (int s = 0; s < segs; ++s) {
        const int rank = std::min(std::max(std::min(per_sm, 8), 1) - 1, (s * cols + cols / 2) / std::max(nsm, 1));
        ws[s] = g_rank_w[rank];
        wsum += ws[s];
    }
    const double c = g_halo_c * H, T = (rows + segs * c) / wsum;
    std::vector<int> &host = *out;
    host.assign(segs + 1, 0);
    double acc = 0.0;
    for (int s = 0; s < segs; ++s) {
        host[s] = (int)(acc + 0.5);
        acc += std::max(1.0, T * ws[s] - c);
    }
    host[segs] = rows;
    // every segment keeps at least one row (segs <= rows)
    for (int s = 1; s < segs; ++s) host[s] = std::max(host[s], host[s - 1] + 1);
    for (int s = segs - 1; s > 0; --s) host[s] = std::min(host[s], host[s + 1] - 1);
    for (int s = 0; s <= segs; ++s) host[s] += y0;
}

static cudaError_t class_segments(int y0, int rows, int cols, int segs, int per_sm, int H, const int **d_out) {
    static int *d_seg = nullptr;
    static int cap = 0;
    static std::vector<int> host;
    static int key[7] = {-1, -1, -1, -1, -1, -1, -1};
    const int k[7] = {y0, rows, cols, segs, per_sm, H, ctx().device};
    if (memcmp(k, key, sizeof k) != 0 || !d_seg) {
        if (segs + 1 > cap) {
            if (d_seg) cudaFree(d_seg);
            d_seg = nullptr;
            cap = 0;
            cudaError_t e = cudaMalloc(&d_seg, (size_t)(segs + 1 + 64) * sizeof(int));
            if (e != cudaSuccess) return e;
            cap = segs + 1 + 64;
        }
        class_segment_rows(y0, rows, cols, segs, per_sm, H, ctx().sm_count, &host);
        cudaError_t e = cudaMemcpyAsync(d_seg, host.data(), (size_t)(segs + 1) * sizeof(int), cudaMemcpyHostToDevice, ctx().stream);
        if (e != cudaSuccess) return e;
        memcpy(key, k, sizeof k);
    }
    *d_out = d_seg;
    return cudaSuccess;
}

static int g_overlap_any = 0;  // pbrt_b200_overlap_passes

template <int H, int TW, bool FMA, bool WIDE = false>
static int launch_class(const ClassParams &CP0) {
    ClassParams CP = CP0;
    SplatParams &P = CP.S;
    const size_t smem = ClassSmem<H, TW>::bytes(CP.blob_bytes, P.spp) + (size_t)class_env_int("PBRT_B200_SMEM_PAD", 0);
    if (smem > 227 * 1024) return -1;
    static int attr_device = -1;
    if (attr_device != ctx().device) {
        PB_CUDA(cudaFuncSetAttribute(splat_class_kernel<H, TW, FMA, false, WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_device = ctx().device;
    }
    int per_sm = 0;
    PB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, splat_class_kernel<H, TW, FMA, false, WIDE>, TW, smem));
    if (per_sm < 1) return -1;
    const int cols = (bw(P.tb) + TW - 1) / TW;
    const int rows = bh(P.tb);
    // one wave: at most as many CTAs as are resident at once, strips no shorter than 4h rows
    const int resident = ctx().sm_count * per_sm;
    int segs = std::max(1, resident / cols);
    segs = std::min(segs, std::max(1, rows / (4 * H)));
    const int rpc = class_env_int("PBRT_B200_ROWS_PER_CTA", 0);
    if (rpc > 0) segs = (rows + rpc - 1) / rpc;
    P.rows_per_cta = (rows + segs - 1) / segs;
    PB_CUDA(class_segments(P.tb.y0, rows, cols, segs, per_sm, H, &CP.seg_start));
    dim3 grid(cols, segs);
#ifdef PBRT_CLASS_TRACE
    static unsigned long long *d_trace = nullptr;
    if (!d_trace) PB_CUDA(cudaMalloc(&d_trace, 4 * sizeof(unsigned long long) * 65536));
    CP.trace = cols * segs <= 65536 ? d_trace : nullptr;
    g_trace = d_trace; g_trace_n = cols * segs; g_trace_cols = cols;
#endif
    // The sample streams are read before the film wait.  That is safe when they were complete before the preceding grid
    // began — taken to hold when this launch reads the very buffers the preceding splat launch read (a multi-pass
    // render over resident samples); any other launch waits first and overlaps only its table staging.
    static const void *last_xy = nullptr, *last_rgbw = nullptr;
    static uint64_t last_launch = ~0ull;
    // Overlap pays where segments are short (the tail it fills is a larger share, and the early CTAs' scrambled placement
    // upsets the rank weights less): +10 % at 14 rows per CTA, +7 % at 28 (C2), +4 % at 60 (a C5 shard of 8), but
    // -2.4 % at 113 (C3) and -1.2 % at 480 (C5) — so only up to PBRT_B200_PDL_MAX_ROWS (80) rows per CTA.
    const bool short_segments = P.rows_per_cta <= class_env_int("PBRT_B200_PDL_MAX_ROWS", 80);
    CP.wait_first = !((g_overlap_any || (last_xy == P.xy && last_rgbw == P.rgbw)) && last_launch == ctx().launches && !P.tiles &&
                      short_segments) ||
                    class_env_int("PBRT_B200_PDL_WAIT_FIRST", 0);
    last_xy = P.xy;
    last_rgbw = P.rgbw;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(TW);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx().stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    // (a launch that must wait before its first load gains nothing from starting early — its CTAs would only hold SM
    // slots the grid ahead still needs: 0.366 ms per pass against 0.350 ms — so it is launched the ordinary way)
    attr[0].val.programmaticStreamSerializationAllowed =
        (class_env_int("PBRT_B200_NO_PDL", 0) || (CP.wait_first && !class_env_int("PBRT_B200_PDL_WAIT_FIRST", 0))) ? 0 : 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    PB_CUDA(cudaLaunchKernelEx(&cfg, splat_class_kernel<H, TW, FMA, false, WIDE>, CP));
    PB_LAUNCH_CHECK("splat_class_kernel");
    last_launch = ctx().launches;  // the library launched nothing else since, if this still equals ctx().launches next time
    return PBRT_OK;
}

template <int H, bool FMA>
static int class_pick_width(const ClassParams &CP) {
#ifdef PBRT_CLASS_PROBE  // SASS probes: one instantiation only
    return launch_class<H, 128, FMA>(CP);
#else
    const int force = class_env_int("PBRT_B200_TW", 0);
    if (force == 128) return launch_class<H, 128, FMA>(CP);
    if (force == 96) return launch_class<H, 96, FMA>(CP);
    if (force == 64) return launch_class<H, 64, FMA>(CP);
    if (force == 32) return launch_class<H, 32, FMA>(CP);
    const int spp = CP.S.spp;
    if (spp > 32) {
        // 33..64 spp: 64-bit index masks, and only the shapes that run on the uniform path (spp divides the strip and
        // two CTAs fit an SM: 64 spp on 64 columns, 48 on 96) — two warps per CTA, four per SM; anything else is
        // better off with the window kernel
        if (64 % spp == 0 && ClassSmem<H, 64>::bytes(CP.blob_bytes, spp) * 2 <= 226 * 1024) return launch_class<H, 64, FMA, true>(CP);
        if (96 % spp == 0 && ClassSmem<H, 96>::bytes(CP.blob_bytes, spp) * 2 <= 226 * 1024) return launch_class<H, 96, FMA, true>(CP);
        return -1;
    }
    // widest strip of which two CTAs fit an SM — among the widths spp divides, if any: a thread then keeps its sample
    // index along a row and the row can run on the uniform path (24 spp: 96 columns 2.2 x the rate of 128)
    const bool fit128 = ClassSmem<H, 128>::bytes(CP.blob_bytes, spp) * 2 <= 226 * 1024;
    const bool fit96 = ClassSmem<H, 96>::bytes(CP.blob_bytes, spp) * 2 <= 226 * 1024;
    const bool fit64 = ClassSmem<H, 64>::bytes(CP.blob_bytes, spp) * 2 <= 226 * 1024;
    if (fit128 && 128 % spp == 0) return launch_class<H, 128, FMA>(CP);
    if (fit96 && 96 % spp == 0) return launch_class<H, 96, FMA>(CP);
    if (fit64 && 64 % spp == 0) return launch_class<H, 64, FMA>(CP);
    if (fit128) return launch_class<H, 128, FMA>(CP);
    if (fit64) return launch_class<H, 64, FMA>(CP);
    return launch_class<H, 32, FMA>(CP);
#endif
}

// Batched tiles (pbrt_film_add_samples_tiles): 32-column strips (a renderer tile of 16x16 samples is 20 pixels wide
// at radius 2), one CTA column per tile walking the whole tile height, blockIdx.z = tile.  Launched without the
// programmatic attribute: the kernel ahead may be the merge that still reads the tile buffers this one overwrites.
template <int H, bool FMA>
static int launch_class_tiles(const ClassParams &CP0, int ntiles, int max_w) {
    constexpr int TW = 32;
    ClassParams CP = CP0;
    const size_t smem = ClassSmem<H, TW>::bytes(CP.blob_bytes, CP.S.spp);
    if (smem > 227 * 1024) return -1;
    static int attr_device = -1;
    if (attr_device != ctx().device) {
        PB_CUDA(cudaFuncSetAttribute(splat_class_kernel<H, TW, FMA, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_device = ctx().device;
    }
    CP.seg_start = nullptr;
    CP.wait_first = 0;
#ifdef PBRT_CLASS_TRACE
    CP.trace = nullptr;
#endif
    for (int z0 = 0; z0 < ntiles; z0 += 65535) {
        const int nz = std::min(65535, ntiles - z0);
        ClassParams Q = CP;
        Q.S.tiles = CP.S.tiles + z0;
        dim3 grid((max_w + TW - 1) / TW, 1, nz);
        splat_class_kernel<H, TW, FMA, true><<<grid, TW, smem, ctx().stream>>>(Q);
        PB_LAUNCH_CHECK("splat_class_kernel(batched)");
    }
    return PBRT_OK;
}

int launch_splat_class_tiles(PbrtFilm *f, const SplatParams &P, int ntiles, int max_w, int mode) {
    if (!f->class_bytes || !P.tiles || class_env_int("PBRT_B200_NO_CLASS", 0)) return -1;
    if (mode != PBRT_SPLAT_EXACT && mode != PBRT_SPLAT_FMA) return -1;
    if (P.spp > 32 || 32 % P.spp != 0) return -1;  // the uniform path needs spp to divide the strip width
    ClassGeom g;
    if (!class_geom(P.rx, P.ry, &g) || g.H != f->class_h) return -1;
    ClassParams CP;
    CP.S = P;
    CP.blob = reinterpret_cast<const uint4 *>(f->d_class);
    CP.blob_bytes = f->class_bytes;
    CP.rowp = g.ROWP;
    CP.eoff = g.EOFF;
#ifdef PBRT_CLASS_PROBE
    return -1;
#else
    const bool fma = mode == PBRT_SPLAT_FMA;
    if (g.H == 2) return fma ? launch_class_tiles<2, true>(CP, ntiles, max_w) : launch_class_tiles<2, false>(CP, ntiles, max_w);
    return fma ? launch_class_tiles<4, true>(CP, ntiles, max_w) : launch_class_tiles<4, false>(CP, ntiles, max_w);
#endif
}

int launch_splat_class(PbrtFilm *f, const SplatParams &P, int mode) {
    if (!f->class_bytes || P.tiles || class_env_int("PBRT_B200_NO_CLASS", 0)) return -1;
    if (mode != PBRT_SPLAT_EXACT && mode != PBRT_SPLAT_FMA) return -1;
    // the per-index class masks cover 64 sample indices; longer pixel runs keep the window kernel's narrower strips
    // (33..64 spp with 64-bit masks is built and tested but measured slower than the window kernel — two CTAs of two
    // warps per SM: 3.0e10 against 3.4e10 samples/s at 64 spp — so it runs only on request)
    if (P.spp > 64 || (P.spp > 32 && !class_env_int("PBRT_B200_WIDE", 0))) return -1;
    ClassGeom g;
    if (!class_geom(P.rx, P.ry, &g) || g.H != f->class_h) return -1;
    ClassParams CP;
    CP.S = P;
    CP.blob = reinterpret_cast<const uint4 *>(f->d_class);
    CP.blob_bytes = f->class_bytes;
    CP.rowp = g.ROWP;
    CP.eoff = g.EOFF;
    const bool fma = mode == PBRT_SPLAT_FMA;
#ifdef PBRT_CLASS_PROBE
    (void)fma;
    return class_pick_width<PBRT_CLASS_PROBE, false>(CP);
#else
    if (g.H == 2) return fma ? class_pick_width<2, true>(CP) : class_pick_width<2, false>(CP);
    return fma ? class_pick_width<4, true>(CP) : class_pick_width<4, false>(CP);
#endif
}

}  // namespace pb

// [UTIL, test hook] the phase-class tables of a filter table, built on the host (no device needed): the blob the
// kernel stages in shared memory and its geometry {H, K, NX, NXP, COLB, BLK, ROWP, EOFF, BYTES}.  Returns the blob's
// size in bytes, 0 when the radius is not served by the class kernel; copies min(size, cap) bytes.
extern "C" int pbrt_b200_debug_class_tables(const float table[256], float rx, float ry, uint8_t *out, int cap,
                                            int32_t geom[9]) {
    std::vector<unsigned char> blob;
    pb::ClassGeom g;
    if (!table || !pb::class_tables_host(table, rx, ry, &blob, &g)) return 0;
    if (geom) {
        const int v[9] = {g.H, g.K, g.NX, g.NXP, g.COLB, g.BLK, g.ROWP, g.EOFF, g.BYTES};
        for (int i = 0; i < 9; ++i) geom[i] = v[i];
    }
    if (out && cap > 0) memcpy(out, blob.data(), std::min<size_t>(blob.size(), (size_t)cap));
    return (int)blob.size();
}

#ifdef PBRT_CLASS_TRACE
// [UTIL, debug builds only] the per-CTA trace of the last splat_class_kernel launch: n x {smid, start ns, end ns, rows}
extern "C" int pbrt_b200_debug_cta_trace(unsigned long long *out, int cap, int *cols) {
    if (!pb::g_trace) return 0;
    cudaDeviceSynchronize();
    const int n = pb::g_trace_n < cap ? pb::g_trace_n : cap;
    cudaMemcpy(out, pb::g_trace, (size_t)n * 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    if (cols) *cols = pb::g_trace_cols;
    return n;
}
#endif

// [UTIL, test hook] the row segments launch_class would give a grid of `cols` strips x `segs` segments over rows
// [y0, y0 + rows) on a device of `nsm` SMs holding `per_sm` CTAs each (host only): segs + 1 ints
extern "C" int pbrt_b200_debug_class_segments(int y0, int rows, int cols, int segs, int per_sm, int h, int nsm, int32_t *out) {
    if (!out || segs < 1 || rows < segs) return 0;
    std::vector<int> v;
    pb::class_segment_rows(y0, rows, cols, segs, per_sm, h, nsm, &v);
    for (int i = 0; i <= segs; ++i) out[i] = v[i];
    return segs + 1;
}

// [UTIL] include/pbrt_b200.h
extern "C" int pbrt_b200_overlap_passes(int on) {
    const int was = pb::g_overlap_any;
    pb::g_overlap_any = on ? 1 : 0;
    return was;
}
