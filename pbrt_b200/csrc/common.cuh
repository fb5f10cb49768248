// common.cuh — shared host/device pieces of libpbrt_b200 (sm_100a only).
//
// The whole library is compiled with -fmad=false: the reference is Rust, which never contracts
// a*b+c, so no kernel may either unless it asks for it explicitly (__fmaf_rn / fma.rn.f32x2 in
// the PBRT_SPLAT_FMA path).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pbrt_b200.h"

namespace pb {

struct Bounds { int x0, y0, x1, y1; };

// bits of the film's sticky device-side error word (PbrtFilm::d_err), mapped to PbrtStatus by pbrt_film_check
enum { ERRBIT_NOT_PIXEL_MAJOR = 1, ERRBIT_NONFINITE = 2 };

__host__ __device__ inline int bw(const Bounds &b) { return b.x1 - b.x0; }
__host__ __device__ inline int bh(const Bounds &b) { return b.y1 - b.y0; }

// ---- colour matrices: src/core/spectrum.rs:129-145, evaluated left to right ----
__device__ __forceinline__ void rgb_to_xyz(float r, float g, float b, float &x, float &y, float &z) {
    x = 0.412453f * r + 0.357580f * g + 0.180423f * b;
    y = 0.212671f * r + 0.715160f * g + 0.072169f * b;
    z = 0.019334f * r + 0.119193f * g + 0.950227f * b;
}
__device__ __forceinline__ void xyz_to_rgb(float x, float y, float z, float &r, float &g, float &b) {
    r = 3.240479f * x - 1.537150f * y - 0.498535f * z;
    g = -0.969256f * x + 1.875991f * y + 0.041556f * z;
    b = 0.055648f * x - 0.204043f * y + 1.057311f * z;
}
// RGBSpectrum::y() of pbrt-v3; weights = row 2 of rgb_to_xyz (spectrum.rs:142)
__device__ __forceinline__ float luminance(float r, float g, float b) {
    return 0.212671f * r + 0.715160f * g + 0.072169f * b;
}

// ---- PCG32: src/core/rng.rs:19-93 ----
struct Pcg32 {
    uint64_t state, inc;
    __host__ __device__ uint32_t next_u32() {
        uint64_t old = state;
        state = old * 0x5851f42d4c957f2dULL + inc;
        uint32_t xorshifted = (uint32_t)(((old >> 18) ^ old) >> 27);
        uint32_t rot = (uint32_t)(old >> 59);
        return (xorshifted >> rot) | (xorshifted << ((~rot + 1u) & 31));
    }
    __host__ __device__ void set_sequence(uint64_t seq) {
        state = 0;
        inc = (seq << 1) | 1;
        next_u32();
        state += 0x853c49e6748fea9bULL;
        next_u32();
    }
    __device__ float next_float() {
        float v = __uint2float_rn(next_u32()) * 2.3283064365386963e-10f;
        const float one_minus_eps = 1.f - 1.1920929e-07f;
        return one_minus_eps < v ? one_minus_eps : v;
    }
};

// ---- streaming loads / stores (data touched once: keep it out of L1) ----
__device__ __forceinline__ float4 ldg_stream(const float4 *p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float2 ldg_stream(const float2 *p) {
    float2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ void stg_stream(float4 *p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

}  // namespace pb

// ---- the film object behind the opaque handle ----
struct PbrtFilm {
    int xres, yres;
    float crop[4];
    float radius[2], inv_radius[2];
    float diagonal_m, scale, max_lum;
    pb::Bounds cropped;  // cropped_pixel_bounds of the whole film
    pb::Bounds owned;    // rows stored by this handle (== cropped unless sharded)
    int64_t npix;
    float table[256];
    float4 *d_xyzw;   // {xyz, filter_weight_sum} per owned pixel
    float *d_splat;   // splat_xyz, 3 floats per owned pixel
    float *d_table;   // 256 floats
    int *d_err;       // sticky async error word
    int device;
    // grow-only staging for host-pointer calls
    void *d_stage[2];
    size_t stage_bytes[2];
    float4 *d_scratch_tile;  // add_samples (arbitrary order) scratch
    size_t scratch_tile_px;
    // merge_tiles: the per-cell tile index of the last batch, reused while the tiling repeats
    // (a renderer merges the same tile grid every pass)
    int32_t *idx_bounds;     // host copy of the batch's tile bounds (4 per tile)
    int64_t *idx_offsets;    // host copy of the batch's pixel offsets
    int idx_ntiles;
    void *d_idx;             // device blob [cell_start | cell_tiles | tile bounds | tile offsets]
    size_t d_idx_bytes;
    pb::Bounds idx_box;
    int idx_cells_x;
    size_t idx_off[5];
    int64_t idx_need;        // pixels the batch's rgbw buffer must hold
    void *d_tile_desc;       // add_samples_tiles: SplatTile array (grow-only)
    size_t tile_desc_bytes;
    void *h_tile_desc;       // its page-locked host staging copy: the upload is asynchronous, ...
    cudaEvent_t ev_tile_desc;  // ... and this event says when the staging copy may be overwritten
    bool tile_desc_event;
    // PBRT_MEM_PINNED_ASYNC inputs: two staging sets filled on the copy stream while the other is consumed
    void *d_pipe[2][4];      // [set][0 = xy, 1 = rgbw, 2 = rgb, 3 = sample weights]
    size_t pipe_bytes[2][4];
    cudaEvent_t ev_staged[2], ev_consumed[2];
    bool pipe_ready;
    int pipe_turn;
    // splat_class.cu: phase-class tables of the filter (LUTs + precombined weight blocks), built when the radius is
    // 1, 2 or 4 on both axes; class_bytes == 0 otherwise
    void *d_class;
    int class_bytes;
    int class_h, class_k, class_rowp;
};

namespace pb {

struct Ctx {
    bool ready;
    int device;
    int sm_count;
    cudaStream_t own_stream;
    cudaStream_t copy_stream;  // uploads of PBRT_MEM_PINNED_ASYNC inputs
    cudaStream_t stream;
    uint64_t launches;
};
Ctx &ctx();
int fail(int code, const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);
int ensure_ready();
int stage_in(PbrtFilm *f, int slot, const void *host, size_t bytes, void **dev_out);

#define PB_CUDA(expr)                                          \
    do {                                                       \
        cudaError_t _e = (expr);                               \
        if (_e != cudaSuccess) return pb::cuda_fail(_e, #expr); \
    } while (0)

#define PB_LAUNCH_CHECK(name)                                     \
    do {                                                          \
        pb::ctx().launches++;                                     \
        cudaError_t _e = cudaGetLastError();                      \
        if (_e != cudaSuccess) return pb::cuda_fail(_e, name);    \
    } while (0)

// one tile of a batched splat (pbrt_film_add_samples_tiles)
struct SplatTile {
    Bounds sb, tb;            // sample bounds; tile pixel bounds = get_film_tile(sb)
    long long sample_offset;  // first sample of the tile in xy / rgbw
    long long pixel_offset;   // first pixel of the tile in the RGBW scratch buffer
};

// arguments of the pixel-major splat kernels (splat.cu, splat_class.cu)
struct SplatParams {
    Bounds sb;      // sample bounds (nominal pixels that carry samples)
    Bounds tb;      // tile pixel bounds = get_film_tile(sb), already clipped to the film
    Bounds owned;   // film rows/cols stored
    int spp;
    float rx, ry, irx, iry;
    float max_lum;
    const float2 *xy;
    const float4 *rgbw;
    const float *table;  // 256 floats, device
    float4 *film;
    int *err;
    int rows_per_cta;
    // batched mode (pbrt_film_add_samples_tiles): blockIdx.z selects a tile; its bounds and streams replace
    // sb / tb / xy / rgbw, and finished pixels go to the tile's own RGBW buffer instead of the film
    const SplatTile *tiles;
    float4 *tile_out;
};

// kernels implemented in splat.cu
int launch_splat_tiles(PbrtFilm *f, int ntiles, const SplatTile *d_tiles, int max_w, int max_h, int spp, const float2 *xy,
                       const float4 *rgbw, float4 *tile_out, int mode);
int launch_splat_tile(PbrtFilm *f, const Bounds &sb, const Bounds &tb, int spp, const float2 *xy, const float4 *rgbw,
                      int mode);

// splat_class.cu: the phase-class gather (radius 1, 2 or 4).  class_tables_create uploads the tables of f->table
// (no-op for other radii); launch_splat_class returns -1 when the film / stream shape is not one it serves.
int class_tables_create(PbrtFilm *f);
int launch_splat_class(PbrtFilm *f, const SplatParams &P, int mode);
int launch_splat_class_tiles(PbrtFilm *f, const SplatParams &P, int ntiles, int max_w, int mode);

}  // namespace pb
