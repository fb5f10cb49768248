// film.cu — C ABI, host-side film logic and the streaming kernels of libpbrt_b200.
//
// Reference lines are relative to the wathiede/pbrt tree.  Host-side bounds arithmetic restates
// src/core/film.rs in f32 exactly (no contraction: built with -Xcompiler -ffp-contract=off);
// the device kernels are the loop bodies of merge_film_tile / write_image / ConstantTexture.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <vector>

#include "common.cuh"

// ===================================================================== runtime

namespace pb {

static thread_local char g_err[512] = "";

Ctx &ctx() {
    static Ctx c = {false, -1, 0, nullptr, nullptr, nullptr, 0};
    return c;
}

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(cudaError_t e, const char *what) {
    return fail(PBRT_E_CUDA, "CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
}

static std::mutex g_init_mu;
// One lock for every compute entry point: the reference serialises all pixel access through one Mutex
// (film.rs:73) and FilmTile is Send, so worker threads may call merge_film_tile concurrently.  Handles carry
// mutable host-side state (staging buffers, cached tile index) and all work goes to one stream anyway.
static std::recursive_mutex g_api_mu;
// The CUDA current device is a per-host-thread setting: a worker thread that never called pbrt_b200_init would launch on
// device 0 against streams and buffers of the bound device.  Every locked entry point therefore re-binds its thread.
static void bind_thread() {
    static thread_local int bound = -1;
    if (ctx().ready && bound != ctx().device) {
        cudaSetDevice(ctx().device);
        bound = ctx().device;
    }
}
struct ApiGuard {
    std::lock_guard<std::recursive_mutex> lock;
    ApiGuard() : lock(g_api_mu) { bind_thread(); }
};
#define PB_API_LOCK pb::ApiGuard pb_api_lock_

int ensure_ready() {
    if (ctx().ready) {
        bind_thread();
        return PBRT_OK;
    }
    return pbrt_b200_init(0);
}

// grow-only per-film staging buffer for host-pointer calls
int stage_in(PbrtFilm *f, int slot, const void *host, size_t bytes, void **dev_out) {
    if (bytes > f->stage_bytes[slot]) {
        if (f->d_stage[slot]) cudaFree(f->d_stage[slot]);
        f->d_stage[slot] = nullptr;
        f->stage_bytes[slot] = 0;
        size_t want = bytes + (bytes >> 3) + 256;
        PB_CUDA(cudaMalloc(&f->d_stage[slot], want));
        f->stage_bytes[slot] = want;
    }
    if (bytes && host) PB_CUDA(cudaMemcpyAsync(f->d_stage[slot], host, bytes, cudaMemcpyHostToDevice, ctx().stream));
    *dev_out = f->d_stage[slot];
    return PBRT_OK;
}

// process-wide scratch for host-destination results that have no film (textures)
static void *g_out_stage = nullptr;
static size_t g_out_stage_bytes = 0;
static int out_stage(size_t bytes, void **dev_out) {
    if (bytes > g_out_stage_bytes) {
        if (g_out_stage) cudaFree(g_out_stage);
        g_out_stage = nullptr;
        g_out_stage_bytes = 0;
        PB_CUDA(cudaMalloc(&g_out_stage, bytes + 256));
        g_out_stage_bytes = bytes + 256;
    }
    *dev_out = g_out_stage;
    return PBRT_OK;
}

// Rust `as isize` for a float: saturating, NaN -> 0 (src/core/geometry/point.rs:323-330)
static int64_t f2i(float v) {
    if (!(v == v)) return 0;
    if (v >= 9223372036854775808.f) return INT64_MAX;
    if (v <= -9223372036854775808.f) return INT64_MIN;
    return (int64_t)v;
}

static bool fits_i32(int64_t v) { return v >= -(1LL << 30) && v <= (1LL << 30); }

}  // namespace pb

using pb::Bounds;
using pb::ctx;
using pb::fail;

extern "C" int pbrt_b200_version(void) { return 100; }
extern "C" const char *pbrt_b200_last_error(void) { return pb::g_err; }

extern "C" int pbrt_b200_init(int device) {
    std::lock_guard<std::mutex> lk(pb::g_init_mu);
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(PBRT_E_CUDA, "no CUDA device (%s); libpbrt_b200 has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(PBRT_E_INVALID, "device %d out of range (have %d)", device, n);
    PB_CUDA(cudaSetDevice(device));
    cudaDeviceProp p;
    PB_CUDA(cudaGetDeviceProperties(&p, device));
    if (p.major != 10 || p.minor != 0)
        return fail(PBRT_E_CUDA, "device %d is sm_%d%d; this library carries sm_100a code only", device, p.major,
                    p.minor);
    pb::Ctx &c = ctx();
    if (c.ready && c.device == device) return PBRT_OK;
    // one device per process (one process per GPU): streams, films and per-kernel attributes belong to it
    if (c.ready)
        return fail(PBRT_E_INVALID, "already bound to device %d; libpbrt_b200 serves one device per process", c.device);
    PB_CUDA(cudaStreamCreateWithFlags(&c.own_stream, cudaStreamNonBlocking));
    PB_CUDA(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
    c.stream = c.own_stream;
    c.device = device;
    c.sm_count = p.multiProcessorCount;
    c.ready = true;
    return PBRT_OK;
}

extern "C" int pbrt_b200_set_stream(void *s) {
    PB_API_LOCK;
    if (int rc = pb::ensure_ready()) return rc;
    ctx().stream = s ? (cudaStream_t)s : ctx().own_stream;
    return PBRT_OK;
}

extern "C" int pbrt_b200_synchronize(void) {
    PB_API_LOCK;
    if (int rc = pb::ensure_ready()) return rc;
    PB_CUDA(cudaStreamSynchronize(ctx().copy_stream));
    PB_CUDA(cudaStreamSynchronize(ctx().stream));
    return PBRT_OK;
}

extern "C" int pbrt_b200_device_info(int *device, int *sm_count, int *cc_major, int *cc_minor, uint64_t *hbm_bytes) {
    if (int rc = pb::ensure_ready()) return rc;
    cudaDeviceProp p;
    PB_CUDA(cudaGetDeviceProperties(&p, ctx().device));
    if (device) *device = ctx().device;
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (hbm_bytes) *hbm_bytes = (uint64_t)p.totalGlobalMem;
    return PBRT_OK;
}

extern "C" uint64_t pbrt_b200_launch_count(void) { return ctx().launches; }

extern "C" int pbrt_b200_malloc(uint64_t bytes, void **out) {
    if (int rc = pb::ensure_ready()) return rc;
    if (!out) return fail(PBRT_E_INVALID, "null out pointer");
    cudaError_t e = cudaMalloc(out, bytes ? bytes : 1);
    if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); return fail(PBRT_E_NOMEM, "cudaMalloc(%llu) failed", (unsigned long long)bytes); }
    PB_CUDA(e);
    return PBRT_OK;
}
extern "C" int pbrt_b200_free(void *p) {
    if (p) PB_CUDA(cudaFree(p));
    return PBRT_OK;
}
extern "C" int pbrt_b200_host_alloc(uint64_t bytes, void **out) {
    if (int rc = pb::ensure_ready()) return rc;
    if (!out) return fail(PBRT_E_INVALID, "null out pointer");
    PB_CUDA(cudaMallocHost(out, bytes ? bytes : 1));
    return PBRT_OK;
}
extern "C" int pbrt_b200_host_free(void *p) {
    if (p) PB_CUDA(cudaFreeHost(p));
    return PBRT_OK;
}
extern "C" int pbrt_b200_memcpy_h2d(void *dev, const void *host, uint64_t bytes) {
    if (int rc = pb::ensure_ready()) return rc;
    PB_CUDA(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, ctx().stream));
    PB_CUDA(cudaStreamSynchronize(ctx().stream));
    return PBRT_OK;
}
extern "C" int pbrt_b200_memcpy_d2h(void *host, const void *dev, uint64_t bytes) {
    if (int rc = pb::ensure_ready()) return rc;
    PB_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx().stream));
    PB_CUDA(cudaStreamSynchronize(ctx().stream));
    return PBRT_OK;
}
extern "C" int pbrt_b200_ipc_export(void *dev, uint8_t handle[64]) {
    if (int rc = pb::ensure_ready()) return rc;
    if (!dev || !handle) return fail(PBRT_E_INVALID, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    PB_CUDA(cudaIpcGetMemHandle(&h, dev));
    memcpy(handle, &h, 64);
    return PBRT_OK;
}
extern "C" int pbrt_b200_ipc_import(const uint8_t handle[64], void **out) {
    if (int rc = pb::ensure_ready()) return rc;
    if (!handle || !out) return fail(PBRT_E_INVALID, "null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    PB_CUDA(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
    return PBRT_OK;
}
extern "C" int pbrt_b200_ipc_close(void *dev) {
    if (dev) PB_CUDA(cudaIpcCloseMemHandle(dev));
    return PBRT_OK;
}
extern "C" int pbrt_b200_memset(void *dev, int byte, uint64_t bytes) {
    if (int rc = pb::ensure_ready()) return rc;
    PB_CUDA(cudaMemsetAsync(dev, byte, bytes, ctx().stream));
    return PBRT_OK;
}

// ===================================================================== filters (host)

struct PbrtFilter {
    int kind;
    float radius[2], inv_radius[2];
    float p0, p1;
    float exp_x, exp_y;
};

namespace {

// [T2] pbrt-v3 7.8 closed forms; only ever consumed through the 16x16 table
float gaussian_1d(float alpha, float d, float expv) {
    float g = expf(-alpha * d * d) - expv;
    return g > 0.f ? g : 0.f;
}
float mitchell_1d(float B, float C, float x) {
    x = fabsf(2.f * x);
    if (x > 1.f)
        return ((-B - 6.f * C) * x * x * x + (6.f * B + 30.f * C) * x * x + (-12.f * B - 48.f * C) * x +
                (8.f * B + 24.f * C)) * (1.f / 6.f);
    return ((12.f - 9.f * B - 6.f * C) * x * x * x + (-18.f + 12.f * B + 6.f * C) * x * x + (6.f - 2.f * B)) *
           (1.f / 6.f);
}
const float kPi = 3.14159265358979323846f;
float sinc_1d(float x) {
    x = fabsf(x);
    if (x < 1e-5f) return 1.f;
    return sinf(kPi * x) / (kPi * x);
}
float windowed_sinc(float x, float radius, float tau) {
    x = fabsf(x);
    if (x > radius) return 0.f;
    return sinc_1d(x) * sinc_1d(x / tau);
}

}  // namespace

extern "C" int pbrt_filter_create(int kind, float rx, float ry, float p0, float p1, PbrtFilter **out) {
    if (!out) return fail(PBRT_E_INVALID, "null out pointer");
    if (kind < PBRT_FILTER_BOX || kind > PBRT_FILTER_LANCZOS) return fail(PBRT_E_INVALID, "unknown filter kind %d", kind);
    PbrtFilter *f = new PbrtFilter();
    f->kind = kind;
    f->radius[0] = rx; f->radius[1] = ry;
    f->inv_radius[0] = 1.f / rx; f->inv_radius[1] = 1.f / ry;  // box.rs:40
    f->p0 = p0; f->p1 = p1;
    f->exp_x = f->exp_y = 0.f;
    if (kind == PBRT_FILTER_GAUSSIAN) {
        f->exp_x = expf(-p0 * rx * rx);
        f->exp_y = expf(-p0 * ry * ry);
    }
    *out = f;
    return PBRT_OK;
}

extern "C" int pbrt_box_filter_create_from_params(int has_xw, float xw, int has_yw, float yw, PbrtFilter **out) {
    return pbrt_filter_create(PBRT_FILTER_BOX, has_xw ? xw : 0.5f, has_yw ? yw : 0.5f, 0.f, 0.f, out);  // box.rs:58-59
}

extern "C" void pbrt_filter_destroy(PbrtFilter *f) { delete f; }

extern "C" float pbrt_filter_evaluate(const PbrtFilter *f, float x, float y) {
    switch (f->kind) {
    case PBRT_FILTER_BOX: return 1.f;  // box.rs:66-68
    case PBRT_FILTER_TRIANGLE: {
        float a = f->radius[0] - fabsf(x), b = f->radius[1] - fabsf(y);
        return (a > 0.f ? a : 0.f) * (b > 0.f ? b : 0.f);
    }
    case PBRT_FILTER_GAUSSIAN: return gaussian_1d(f->p0, x, f->exp_x) * gaussian_1d(f->p0, y, f->exp_y);
    case PBRT_FILTER_MITCHELL:
        return mitchell_1d(f->p0, f->p1, x * f->inv_radius[0]) * mitchell_1d(f->p0, f->p1, y * f->inv_radius[1]);
    case PBRT_FILTER_LANCZOS: return windowed_sinc(x, f->radius[0], f->p0) * windowed_sinc(y, f->radius[1], f->p0);
    }
    return 0.f;
}
extern "C" void pbrt_filter_radius(const PbrtFilter *f, float out[2]) { out[0] = f->radius[0]; out[1] = f->radius[1]; }
extern "C" void pbrt_filter_inv_radius(const PbrtFilter *f, float out[2]) {
    out[0] = f->inv_radius[0];
    out[1] = f->inv_radius[1];
}

// film.rs:113-123
extern "C" int pbrt_filter_table(const PbrtFilter *f, float table[256]) {
    if (!f || !table) return fail(PBRT_E_INVALID, "null argument");
    const float w = (float)PBRT_FILTER_TABLE_WIDTH;
    int k = 0;
    for (int y = 0; y < PBRT_FILTER_TABLE_WIDTH; ++y)
        for (int x = 0; x < PBRT_FILTER_TABLE_WIDTH; ++x) {
            float fx = ((float)x + 0.5f) * f->radius[0] / w;
            float fy = ((float)y + 0.5f) * f->radius[1] / w;
            table[k++] = pbrt_filter_evaluate(f, fx, fy);
        }
    return PBRT_OK;
}

// ===================================================================== film: host logic

// Film::new's geometry (film.rs:92-101, :129, :450) into *f; no device involved
static int film_geometry_init(PbrtFilm *f, int32_t xres, int32_t yres, const float crop[4], const float radius[2],
                              float diagonal_mm, int rank, int nranks) {
    if (!crop || !radius) return fail(PBRT_E_INVALID, "null argument");
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(PBRT_E_INVALID, "rank %d of %d", rank, nranks);
    // film.rs:92-101 — ceil(res * crop) per corner, then Bounds2i::from sorts each axis
    int64_t ax = pb::f2i(ceilf((float)xres * crop[0])), ay = pb::f2i(ceilf((float)yres * crop[1]));
    int64_t bx = pb::f2i(ceilf((float)xres * crop[2])), by = pb::f2i(ceilf((float)yres * crop[3]));
    int64_t x0 = std::min(ax, bx), x1 = std::max(ax, bx), y0 = std::min(ay, by), y1 = std::max(ay, by);
    if (!pb::fits_i32(x0) || !pb::fits_i32(x1) || !pb::fits_i32(y0) || !pb::fits_i32(y1))
        return fail(PBRT_E_RANGE, "cropped pixel bounds do not fit the device's 32-bit coordinates");
    memset(f, 0, sizeof *f);
    f->xres = xres; f->yres = yres;
    memcpy(f->crop, crop, sizeof f->crop);
    f->radius[0] = radius[0]; f->radius[1] = radius[1];
    f->inv_radius[0] = 1.f / radius[0]; f->inv_radius[1] = 1.f / radius[1];  // film.rs:450
    f->diagonal_m = diagonal_mm * 0.001f;                                    // film.rs:129
    f->cropped = Bounds{(int)x0, (int)y0, (int)x1, (int)y1};
    f->owned = f->cropped;
    if (nranks > 1) {
        int64_t H = y1 - y0;
        f->owned.y0 = (int)(y0 + H * rank / nranks);
        f->owned.y1 = (int)(y0 + H * (rank + 1) / nranks);
    }
    return PBRT_OK;
}

static int film_create_impl(int32_t xres, int32_t yres, const float crop[4], const float radius[2],
                            const float table[256], float diagonal_mm, float scale, float max_lum, int rank,
                            int nranks, PbrtFilm **out) {
    if (int rc = pb::ensure_ready()) return rc;
    if (!crop || !radius || !table || !out) return fail(PBRT_E_INVALID, "null argument");
    PbrtFilm geo;
    if (int rc = film_geometry_init(&geo, xres, yres, crop, radius, diagonal_mm, rank, nranks)) return rc;
    PbrtFilm *f = new PbrtFilm(geo);
    f->scale = scale; f->max_lum = max_lum;
    int64_t area = (int64_t)pb::bw(f->owned) * pb::bh(f->owned);
    f->npix = area > 0 ? area : 0;
    memcpy(f->table, table, sizeof f->table);
    f->device = ctx().device;
    size_t n = (size_t)(f->npix ? f->npix : 1);
    cudaError_t e = cudaMalloc(&f->d_xyzw, n * sizeof(float4));
    if (e == cudaSuccess) e = cudaMalloc(&f->d_splat, n * 3 * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&f->d_table, 256 * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&f->d_err, sizeof(int));
    if (e != cudaSuccess) {
        cudaGetLastError();
        const long long npix = (long long)f->npix;
        pbrt_film_destroy(f);
        return e == cudaErrorMemoryAllocation ? fail(PBRT_E_NOMEM, "film of %lld pixels does not fit", npix)
                                              : pb::cuda_fail(e, "film alloc");
    }
    cudaStream_t s = ctx().stream;
    PB_CUDA(cudaMemsetAsync(f->d_xyzw, 0, n * sizeof(float4), s));  // Pixel::default(), film.rs:106-110
    PB_CUDA(cudaMemsetAsync(f->d_splat, 0, n * 3 * sizeof(float), s));
    PB_CUDA(cudaMemsetAsync(f->d_err, 0, sizeof(int), s));
    PB_CUDA(cudaMemcpyAsync(f->d_table, f->table, sizeof f->table, cudaMemcpyHostToDevice, s));
    PB_CUDA(cudaStreamSynchronize(s));
    if (int rc = pb::class_tables_create(f)) {
        pbrt_film_destroy(f);
        return rc;
    }
    *out = f;
    return PBRT_OK;
}

extern "C" int pbrt_film_create(int32_t xres, int32_t yres, const float crop[4], const float radius[2],
                                const float table[256], float diagonal_mm, float scale, float max_lum,
                                PbrtFilm **out) {
    return film_create_impl(xres, yres, crop, radius, table, diagonal_mm, scale, max_lum, 0, 1, out);
}

extern "C" int pbrt_film_create_sharded(int32_t xres, int32_t yres, const float crop[4], const float radius[2],
                                        const float table[256], float diagonal_mm, float scale, float max_lum,
                                        int rank, int nranks, PbrtFilm **out) {
    return film_create_impl(xres, yres, crop, radius, table, diagonal_mm, scale, max_lum, rank, nranks, out);
}

extern "C" int pbrt_film_destroy(PbrtFilm *f) {
    PB_API_LOCK;
    if (!f) return PBRT_OK;
    if (ctx().ready) { cudaStreamSynchronize(ctx().copy_stream); cudaStreamSynchronize(ctx().stream); }
    cudaFree(f->d_xyzw);
    cudaFree(f->d_splat);
    cudaFree(f->d_table);
    cudaFree(f->d_err);
    cudaFree(f->d_stage[0]);
    cudaFree(f->d_stage[1]);
    cudaFree(f->d_scratch_tile);
    cudaFree(f->d_idx);
    cudaFree(f->d_tile_desc);
    if (f->h_tile_desc) cudaFreeHost(f->h_tile_desc);
    if (f->tile_desc_event) cudaEventDestroy(f->ev_tile_desc);
    cudaFree(f->d_class);
    for (int i = 0; i < 2; ++i) {
        for (int k = 0; k < 4; ++k) cudaFree(f->d_pipe[i][k]);
        if (f->pipe_ready) { cudaEventDestroy(f->ev_staged[i]); cudaEventDestroy(f->ev_consumed[i]); }
    }
    free(f->idx_bounds);
    free(f->idx_offsets);
    cudaGetLastError();
    delete f;
    return PBRT_OK;
}

static void put_bounds(const Bounds &b, int32_t out[4]) { out[0] = b.x0; out[1] = b.y0; out[2] = b.x1; out[3] = b.y1; }

extern "C" int pbrt_film_cropped_pixel_bounds(const PbrtFilm *f, int32_t out[4]) {
    if (!f || !out) return fail(PBRT_E_INVALID, "null argument");
    put_bounds(f->cropped, out);
    return PBRT_OK;
}
extern "C" int pbrt_film_owned_pixel_bounds(const PbrtFilm *f, int32_t out[4]) {
    if (!f || !out) return fail(PBRT_E_INVALID, "null argument");
    put_bounds(f->owned, out);
    return PBRT_OK;
}

// film.rs:166-175
extern "C" int pbrt_film_get_sample_bounds(const PbrtFilm *f, int32_t out[4]) {
    if (!f || !out) return fail(PBRT_E_INVALID, "null argument");
    float ax = floorf((float)f->cropped.x0 + 0.5f - f->radius[0]);
    float ay = floorf((float)f->cropped.y0 + 0.5f - f->radius[1]);
    float bx = ceilf((float)f->cropped.x1 - 0.5f + f->radius[0]);
    float by = ceilf((float)f->cropped.y1 - 0.5f + f->radius[1]);
    int64_t v[4] = {pb::f2i(ax < bx ? ax : bx), pb::f2i(ay < by ? ay : by), pb::f2i(ax > bx ? ax : bx),
                    pb::f2i(ay > by ? ay : by)};
    for (int i = 0; i < 4; ++i) {
        if (!pb::fits_i32(v[i])) return fail(PBRT_E_RANGE, "sample bounds do not fit 32-bit coordinates");
        out[i] = (int32_t)v[i];
    }
    return PBRT_OK;
}

// film.rs:218-227
extern "C" int pbrt_film_get_physical_extent(const PbrtFilm *f, float out[4]) {
    if (!f || !out) return fail(PBRT_E_INVALID, "null argument");
    float aspect = (float)f->yres / (float)f->xres;
    float x = sqrtf(f->diagonal_m * f->diagonal_m / (1.f + aspect * aspect));
    float y = aspect * x;
    float ax = -x / 2.f, ay = -y / 2.f, bx = x / 2.f, by = y / 2.f;
    out[0] = ax < bx ? ax : bx; out[1] = ay < by ? ay : by;
    out[2] = ax > bx ? ax : bx; out[3] = ay > by ? ay : by;
    return PBRT_OK;
}

// film.rs:264-273 against `clip` (cropped bounds, or the owned rows of a shard)
static int tile_bounds_impl(const PbrtFilm *f, const int32_t sb[4], const Bounds &clip, Bounds *out) {
    int64_t p0x = pb::f2i(ceilf((float)sb[0] - 0.5f - f->radius[0]));
    int64_t p0y = pb::f2i(ceilf((float)sb[1] - 0.5f - f->radius[1]));
    int64_t p1x = pb::f2i(floorf((float)sb[2] - 0.5f + f->radius[0]) + 1.f);
    int64_t p1y = pb::f2i(floorf((float)sb[3] - 0.5f + f->radius[1]) + 1.f);
    int64_t x0 = std::min(p0x, p1x), x1 = std::max(p0x, p1x), y0 = std::min(p0y, p1y), y1 = std::max(p0y, p1y);
    x0 = std::max<int64_t>(x0, clip.x0); y0 = std::max<int64_t>(y0, clip.y0);   // Bounds2i::intersect: no re-sort
    x1 = std::min<int64_t>(x1, clip.x1); y1 = std::min<int64_t>(y1, clip.y1);
    if (!pb::fits_i32(x0) || !pb::fits_i32(x1) || !pb::fits_i32(y0) || !pb::fits_i32(y1))
        return fail(PBRT_E_RANGE, "tile bounds do not fit 32-bit coordinates");
    *out = Bounds{(int)x0, (int)y0, (int)x1, (int)y1};
    return PBRT_OK;
}

extern "C" int pbrt_film_tile_bounds(const PbrtFilm *f, const int32_t sb[4], int32_t out[4], int64_t *pixel_count) {
    if (!f || !sb || !out) return fail(PBRT_E_INVALID, "null argument");
    Bounds b;
    if (int rc = tile_bounds_impl(f, sb, f->owned, &b)) return rc;
    put_bounds(b, out);
    if (pixel_count) {
        int64_t area = (int64_t)pb::bw(b) * pb::bh(b);  // bounds.rs:195-198
        *pixel_count = area > 0 ? area : 0;             // film.rs:446
    }
    return PBRT_OK;
}

// [T1] the same computations without a film object or a device (tests; hosts that only need the bounds)
extern "C" int pbrt_film_geometry(int32_t xres, int32_t yres, const float crop[4], const float radius[2],
                                  float diagonal_mm, int rank, int nranks, int32_t cropped[4], int32_t owned[4],
                                  int32_t sample_bounds[4], float physical_extent[4]) {
    PbrtFilm geo;
    if (int rc = film_geometry_init(&geo, xres, yres, crop, radius, diagonal_mm, rank, nranks)) return rc;
    if (cropped) put_bounds(geo.cropped, cropped);
    if (owned) put_bounds(geo.owned, owned);
    if (sample_bounds)
        if (int rc = pbrt_film_get_sample_bounds(&geo, sample_bounds)) return rc;
    if (physical_extent)
        if (int rc = pbrt_film_get_physical_extent(&geo, physical_extent)) return rc;
    return PBRT_OK;
}

extern "C" int pbrt_film_geometry_tile_bounds(const int32_t clip[4], const float radius[2], const int32_t sb[4],
                                              int32_t out[4], int64_t *pixel_count) {
    if (!clip || !radius || !sb || !out) return fail(PBRT_E_INVALID, "null argument");
    PbrtFilm geo;
    memset(&geo, 0, sizeof geo);
    geo.radius[0] = radius[0]; geo.radius[1] = radius[1];
    geo.owned = Bounds{clip[0], clip[1], clip[2], clip[3]};
    return pbrt_film_tile_bounds(&geo, sb, out, pixel_count);
}

// [UTIL] Routing of a pixel-major stream to row shards.  A source that holds the nominal sample rows
// [src_rows[0], src_rows[1]) owes shard g the rows of its own block plus the halo h = floor(r.y + .5) on either side,
// clipped to the sample bounds: in a pixel-major stream that is ONE contiguous run of samples, and a row within h of a
// shard edge simply belongs to two runs.  No device work: the exchange is nranks sends of slices of the source buffer.
extern "C" int pbrt_film_route_plan(const int32_t sample_bounds[4], const int32_t cropped[4], const float radius[2],
                                    int32_t nranks, const int32_t src_rows[2], int32_t out_rows[]) {
    if (!sample_bounds || !cropped || !radius || !src_rows || !out_rows) return fail(PBRT_E_INVALID, "null argument");
    if (nranks < 1) return fail(PBRT_E_INVALID, "nranks %d", nranks);
    const int64_t y0 = cropped[1], H = (int64_t)cropped[3] - cropped[1];
    const int h = (int)floorf(radius[1] + 0.5f);
    for (int g = 0; g < nranks; ++g) {
        // the same split pbrt_film_create_sharded makes
        const int64_t oy0 = y0 + H * g / nranks, oy1 = y0 + H * (g + 1) / nranks;
        int64_t a = std::max<int64_t>(std::max<int64_t>(oy0 - h, sample_bounds[1]), src_rows[0]);
        int64_t b = std::min<int64_t>(std::min<int64_t>(oy1 + h, sample_bounds[3]), src_rows[1]);
        if (oy1 <= oy0 || b < a) b = a;  // an empty shard needs nothing
        out_rows[2 * g] = (int32_t)a;
        out_rows[2 * g + 1] = (int32_t)b;
    }
    return PBRT_OK;
}

// ===================================================================== kernels: merge

// One tile: film.rs:317-325.  16 B tile read + 16 B film read + 16 B film write per pixel.
__global__ void __launch_bounds__(256) merge_tile_kernel(float4 *__restrict__ film, Bounds owned, Bounds tb,
                                                         const float4 *__restrict__ tile) {
    const int tw = tb.x1 - tb.x0;
    const int fw = owned.x1 - owned.x0;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= tw) return;
    for (int y = blockIdx.y; y < tb.y1 - tb.y0; y += gridDim.y) {
        float4 t = pb::ldg_stream(&tile[(size_t)y * tw + x]);
        size_t fo = (size_t)(tb.y0 + y - owned.y0) * fw + (tb.x0 + x - owned.x0);
        float4 p = film[fo];
        float X, Y, Z;
        pb::rgb_to_xyz(t.x, t.y, t.z, X, Y, Z);
        p.x += X; p.y += Y; p.z += Z; p.w += t.w;
        film[fo] = p;
    }
}

// Many tiles, one launch, tile order preserved per pixel.  A thread owns one film pixel of the
// union box and walks the tiles recorded for its 16x16 cell in ascending tile index.
struct MergeIndex {
    Bounds box;       // union of the tiles, clipped to the film
    int cells_x;      // cells across
    const int *cell_start;  // CSR, cells_x*cells_y + 1
    const int *cell_tiles;
    const int4 *tile_bounds;
    const long long *tile_offset;
    const struct CellHead *cell_head;
};

// The first four tiles of a cell, inline: one 128-byte fetch tells a CTA everything it needs for the usual
// case (a cell of a regular tiling is covered by at most 2x2 tiles), instead of the three dependent global
// loads of the CSR walk (cell_start -> cell_tiles -> tile_bounds/offset) ahead of the first pixel load.
struct __align__(16) CellHead {
    int count, pad0, pad1, pad2;
    int4 bounds[4];
    long long offset[4];
    long long pad3[2];
};
static_assert(sizeof(CellHead) == 128, "one cache line per cell");

// One CTA per 16x16 cell: the cell's tile list (bounds, offsets) is the same for all 256 pixels, so it
// is fetched into shared memory once and read back as broadcasts; the per-pixel work is then one
// float4 film read, one float4 tile read per covering tile, one float4 film write.
constexpr int MERGE_CHUNK = 32;

__global__ void __launch_bounds__(256) merge_tiles_kernel(float4 *__restrict__ film, Bounds owned, MergeIndex ix,
                                                          const float4 *__restrict__ tiles) {
    __shared__ CellHead s_head;
    __shared__ int4 s_bounds[MERGE_CHUNK];
    __shared__ long long s_offset[MERGE_CHUNK];
    const int cell = blockIdx.y * ix.cells_x + blockIdx.x;
    if (threadIdx.x < 8)
        reinterpret_cast<int4 *>(&s_head)[threadIdx.x] = reinterpret_cast<const int4 *>(ix.cell_head + cell)[threadIdx.x];
    const int x = ix.box.x0 + blockIdx.x * 16 + (threadIdx.x & 15);
    const int y = ix.box.y0 + blockIdx.y * 16 + (threadIdx.x >> 4);
    const bool in_box = x < ix.box.x1 && y < ix.box.y1;
    const size_t fo = (size_t)(y - owned.y0) * (owned.x1 - owned.x0) + (x - owned.x0);
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (in_box) p = film[fo];  // in flight while the head arrives
    __syncthreads();
    const int total = s_head.count;
    if (total == 0) return;
    bool touched = false;
    auto add_tile = [&](const int4 b, long long off) {  // ascending tile index: the order sequential merge_film_tile calls would use
        if (x < b.x || x >= b.z || y < b.y || y >= b.w) return;
        const float4 v = pb::ldg_stream(&tiles[off + (size_t)(y - b.y) * (b.z - b.x) + (x - b.x)]);
        float X, Y, Z;
        pb::rgb_to_xyz(v.x, v.y, v.z, X, Y, Z);
        p.x += X; p.y += Y; p.z += Z; p.w += v.w;
        touched = true;
    };
    if (in_box) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (k < total) add_tile(s_head.bounds[k], s_head.offset[k]);
    }
    if (total > 4) {  // the rest of a crowded cell through the CSR, a chunk at a time
        const int beg = ix.cell_start[cell] + 4, end = ix.cell_start[cell + 1];
        for (int k0 = beg; k0 < end; k0 += MERGE_CHUNK) {
            const int n = min(MERGE_CHUNK, end - k0);
            __syncthreads();
            if ((int)threadIdx.x < n) {
                const int t = ix.cell_tiles[k0 + threadIdx.x];
                s_bounds[threadIdx.x] = ix.tile_bounds[t];
                s_offset[threadIdx.x] = ix.tile_offset[t];
            }
            __syncthreads();
            if (in_box)
                for (int k = 0; k < n; ++k) add_tile(s_bounds[k], s_offset[k]);
        }
    }
    if (touched) film[fo] = p;
}

static bool inside(const Bounds &outer, const Bounds &b) {
    return b.x0 >= outer.x0 && b.y0 >= outer.y0 && b.x1 <= outer.x1 && b.y1 <= outer.y1;
}

// entry points that take host or device memory only (PBRT_MEM_PINNED_ASYNC is for add_samples_tile[_rgb] and resolve)
static int host_or_device(int kind, const char *what) {
    if (kind == PBRT_MEM_HOST || kind == PBRT_MEM_DEVICE) return PBRT_OK;
    return fail(PBRT_E_INVALID, "%s: memory kind %d not supported here (PBRT_MEM_HOST or PBRT_MEM_DEVICE)", what, kind);
}

extern "C" int pbrt_film_merge_tile(PbrtFilm *f, const int32_t tbv[4], const float *rgbw, int src_is_device) {
    PB_API_LOCK;
    if (int rc = host_or_device(src_is_device, "merge_film_tile")) return rc;
    if (!f || !tbv) return fail(PBRT_E_INVALID, "null argument");
    Bounds tb{tbv[0], tbv[1], tbv[2], tbv[3]};
    // Bounds2i::iter over an inverted or empty box yields nothing (bounds.rs:284-288)
    if (tb.x1 <= tb.x0 || tb.y1 <= tb.y0) return PBRT_OK;
    if (!inside(f->owned, tb))
        return fail(PBRT_E_RANGE, "tile [%d,%d)x[%d,%d) outside film [%d,%d)x[%d,%d)", tb.x0, tb.x1, tb.y0, tb.y1,
                    f->owned.x0, f->owned.x1, f->owned.y0, f->owned.y1);
    if (!rgbw) return fail(PBRT_E_INVALID, "null tile pixels");
    const float4 *d_tile = (const float4 *)rgbw;
    if (!src_is_device) {
        void *d;
        if (int rc = pb::stage_in(f, 0, rgbw, (size_t)pb::bw(tb) * pb::bh(tb) * sizeof(float4), &d)) return rc;
        d_tile = (const float4 *)d;
    }
    dim3 block(256);
    dim3 grid((pb::bw(tb) + 255) / 256, std::min(pb::bh(tb), 4096));
    merge_tile_kernel<<<grid, block, 0, ctx().stream>>>(f->d_xyzw, f->owned, tb, d_tile);
    PB_LAUNCH_CHECK("merge_tile_kernel");
    return PBRT_OK;
}

extern "C" int pbrt_film_merge_tiles(PbrtFilm *f, int32_t ntiles, const int32_t *tbs, const int64_t *offsets,
                                     const float *rgbw, int64_t total_pixels, int src_is_device) {
    PB_API_LOCK;
    if (int rc = host_or_device(src_is_device, "merge_film_tiles")) return rc;
    if (!f) return fail(PBRT_E_INVALID, "null film");
    if (ntiles <= 0) return PBRT_OK;
    if (!tbs || !offsets || !rgbw) return fail(PBRT_E_INVALID, "null argument");
    const bool cached = f->d_idx && f->idx_ntiles == ntiles &&
                        memcmp(f->idx_bounds, tbs, (size_t)ntiles * 4 * sizeof(int32_t)) == 0 &&
                        memcmp(f->idx_offsets, offsets, (size_t)ntiles * sizeof(int64_t)) == 0;
    if (!cached) {
        // union box and range checks
        Bounds box{INT32_MAX, INT32_MAX, INT32_MIN, INT32_MIN};
        std::vector<int4> tb(ntiles);
        std::vector<long long> off(ntiles);
        bool any = false;
        int64_t need = 0;
        for (int i = 0; i < ntiles; ++i) {
            Bounds b{tbs[4 * i], tbs[4 * i + 1], tbs[4 * i + 2], tbs[4 * i + 3]};
            off[i] = offsets[i];
            if (b.x1 <= b.x0 || b.y1 <= b.y0) { tb[i] = make_int4(0, 0, 0, 0); continue; }
            if (!inside(f->owned, b)) return fail(PBRT_E_RANGE, "tile %d outside the film", i);
            if (offsets[i] < 0 || offsets[i] + (int64_t)pb::bw(b) * pb::bh(b) > total_pixels)
                return fail(PBRT_E_INVALID, "tile %d: pixels outside the rgbw buffer", i);
            tb[i] = make_int4(b.x0, b.y0, b.x1, b.y1);
            need = std::max<int64_t>(need, offsets[i] + (int64_t)pb::bw(b) * pb::bh(b));
            box.x0 = std::min(box.x0, b.x0); box.y0 = std::min(box.y0, b.y0);
            box.x1 = std::max(box.x1, b.x1); box.y1 = std::max(box.y1, b.y1);
            any = true;
        }
        if (!any) return PBRT_OK;
        // CSR of tiles per 16x16 cell, ascending tile index
        const int cx = (pb::bw(box) + 15) / 16, cy = (pb::bh(box) + 15) / 16;
        std::vector<int> start((size_t)cx * cy + 1, 0);
        for (int i = 0; i < ntiles; ++i) {
            if (tb[i].z <= tb[i].x) continue;
            for (int gy = (tb[i].y - box.y0) >> 4; gy <= (tb[i].w - 1 - box.y0) >> 4; ++gy)
                for (int gx = (tb[i].x - box.x0) >> 4; gx <= (tb[i].z - 1 - box.x0) >> 4; ++gx) start[(size_t)gy * cx + gx + 1]++;
        }
        for (size_t c = 0; c < (size_t)cx * cy; ++c) start[c + 1] += start[c];
        std::vector<int> fill(start.begin(), start.end() - 1);
        std::vector<int> list((size_t)start.back());
        for (int i = 0; i < ntiles; ++i) {
            if (tb[i].z <= tb[i].x) continue;
            for (int gy = (tb[i].y - box.y0) >> 4; gy <= (tb[i].w - 1 - box.y0) >> 4; ++gy)
                for (int gx = (tb[i].x - box.x0) >> 4; gx <= (tb[i].z - 1 - box.x0) >> 4; ++gx) list[(size_t)fill[(size_t)gy * cx + gx]++] = i;
        }
        // the first four tiles of every cell, inline
        std::vector<CellHead> heads((size_t)cx * cy);
        memset(heads.data(), 0, heads.size() * sizeof(CellHead));
        for (size_t c = 0; c < heads.size(); ++c) {
            heads[c].count = start[c + 1] - start[c];
            for (int k = 0; k < std::min(heads[c].count, 4); ++k) {
                const int t = list[(size_t)start[c] + k];
                heads[c].bounds[k] = tb[t];
                heads[c].offset[k] = off[t];
            }
        }
        // one upload: [start | list | bounds | offsets | heads], each part 16-byte aligned
        auto align16 = [](size_t v) { return (v + 15) & ~size_t(15); };
        const size_t b0 = 0, b1 = align16(start.size() * sizeof(int)), b2 = b1 + align16(list.size() * sizeof(int));
        const size_t b3 = b2 + tb.size() * sizeof(int4);
        const size_t b4 = (b3 + off.size() * sizeof(long long) + 127) & ~size_t(127);
        const size_t b5 = b4 + heads.size() * sizeof(CellHead);
        std::vector<unsigned char> blob(b5);
        memcpy(&blob[b4], heads.data(), heads.size() * sizeof(CellHead));
        memcpy(&blob[b0], start.data(), start.size() * sizeof(int));
        if (!list.empty()) memcpy(&blob[b1], list.data(), list.size() * sizeof(int));
        memcpy(&blob[b2], tb.data(), tb.size() * sizeof(int4));
        memcpy(&blob[b3], off.data(), off.size() * sizeof(long long));
        if (blob.size() > f->d_idx_bytes) {
            cudaFree(f->d_idx);
            f->d_idx = nullptr;
            f->d_idx_bytes = 0;
            PB_CUDA(cudaMalloc(&f->d_idx, blob.size() + 256));
            f->d_idx_bytes = blob.size() + 256;
        }
        f->idx_ntiles = 0;  // invalid until everything below succeeded
        PB_CUDA(cudaMemcpyAsync(f->d_idx, blob.data(), blob.size(), cudaMemcpyHostToDevice, ctx().stream));
        PB_CUDA(cudaStreamSynchronize(ctx().stream));  // blob is a local
        free(f->idx_bounds);
        free(f->idx_offsets);
        f->idx_bounds = (int32_t *)malloc((size_t)ntiles * 4 * sizeof(int32_t));
        f->idx_offsets = (int64_t *)malloc((size_t)ntiles * sizeof(int64_t));
        if (!f->idx_bounds || !f->idx_offsets) return fail(PBRT_E_NOMEM, "host allocation failed");
        memcpy(f->idx_bounds, tbs, (size_t)ntiles * 4 * sizeof(int32_t));
        memcpy(f->idx_offsets, offsets, (size_t)ntiles * sizeof(int64_t));
        f->idx_box = box;
        f->idx_cells_x = cx;
        f->idx_off[0] = b0; f->idx_off[1] = b1; f->idx_off[2] = b2; f->idx_off[3] = b3; f->idx_off[4] = b4;
        f->idx_need = need;
        f->idx_ntiles = ntiles;
    }
    if (total_pixels < f->idx_need) return fail(PBRT_E_INVALID, "rgbw buffer smaller than the tiles it must hold");
    const Bounds box = f->idx_box;
    void *d_blob = f->d_idx;
    const size_t b0 = f->idx_off[0], b1 = f->idx_off[1], b2 = f->idx_off[2], b3 = f->idx_off[3];
    const int cx = f->idx_cells_x;
    const float4 *d_tiles = (const float4 *)rgbw;
    if (!src_is_device) {
        void *d;
        if (int rc = pb::stage_in(f, 0, rgbw, (size_t)total_pixels * sizeof(float4), &d)) return rc;
        d_tiles = (const float4 *)d;
    }
    MergeIndex ix;
    ix.box = box;
    ix.cells_x = cx;
    ix.cell_start = (const int *)((char *)d_blob + b0);
    ix.cell_tiles = (const int *)((char *)d_blob + b1);
    ix.tile_bounds = (const int4 *)((char *)d_blob + b2);
    ix.tile_offset = (const long long *)((char *)d_blob + b3);
    ix.cell_head = (const CellHead *)((char *)d_blob + f->idx_off[4]);
    dim3 grid(cx, (pb::bh(box) + 15) / 16);  // one CTA per 16x16 cell of the index
    merge_tiles_kernel<<<grid, 256, 0, ctx().stream>>>(f->d_xyzw, f->owned, ix, d_tiles);
    PB_LAUNCH_CHECK("merge_tiles_kernel");
    return PBRT_OK;
}

// ===================================================================== kernels: resolve

// Rust f32::max(v, 0.): NaN -> 0 (film.rs:359-361).  max(-0., 0.) is unspecified there; +0 here.
__device__ __forceinline__ float max0(float v) { return v > 0.f ? v : 0.f; }

// film.rs:346-372 for one pixel
__device__ __forceinline__ void resolve_pixel(float4 p, float sx, float sy, float sz, float splat_scale, float scale,
                                              float &r, float &g, float &b) {
    pb::xyz_to_rgb(p.x, p.y, p.z, r, g, b);
    if (p.w != 0.f) {
        float inv = 1.f / p.w;
        r = max0(r * inv); g = max0(g * inv); b = max0(b * inv);
    }
    float sr, sg, sb;
    pb::xyz_to_rgb(sx, sy, sz, sr, sg, sb);
    r += splat_scale * sr; g += splat_scale * sg; b += splat_scale * sb;
    r *= scale; g *= scale; b *= scale;
}

// src/lib.rs:93-99 + src/core/imageio.rs:66-68: byte = clamp(255 * gamma_correct(v) + .5, 0, 255) as u8.
// to_byte is monotone in v, so it is fully described by 255 thresholds: kToByteThreshold[k] is the smallest
// f32 with to_byte >= k under a correctly rounded powf (tools/make_to_byte_table.py, no libm involved).
// The kernel needs no power function at all: the top bits of v (exponent + 7 mantissa bits, "slice") index a
// byte table holding to_byte of the slice's first value; a slice is narrow enough to contain at most one
// threshold (tests/test_host.py checks that on the table), so one comparison against the next threshold
// finishes the job.  Exact for every input, no slow path, no branch.
__device__ const unsigned kToByteThreshold[256] = {
#include "to_byte_table.inc"
};
constexpr unsigned SLICE_SHIFT = 16;                      // 7 mantissa bits per slice
constexpr unsigned SLICE_FIRST_BITS = 0x39000000u;        // 2^-13 < threshold[1]
constexpr int SLICE_COUNT = ((0x3f800000u - SLICE_FIRST_BITS) >> SLICE_SHIFT) + 1;  // up to and including 1.0
constexpr int SLICE_WORDS = (SLICE_COUNT + 3) / 4;
__device__ unsigned kToByteSlice[SLICE_WORDS];            // filled once by to_byte_slices_kernel

__global__ void to_byte_slices_kernel() {
    unsigned char *out = reinterpret_cast<unsigned char *>(kToByteSlice);
    for (int i = threadIdx.x; i < SLICE_WORDS * 4; i += blockDim.x) {
        const unsigned start = SLICE_FIRST_BITS + ((unsigned)min(i, SLICE_COUNT - 1) << SLICE_SHIFT);
        int k = 0;  // number of thresholds <= start (positive floats order like their bit patterns)
        for (int step = 128; step > 0; step >>= 1)
            if (k + step <= 255 && kToByteThreshold[k + step] <= start) k += step;
        out[i] = (unsigned char)k;
    }
}

// thr = [unused, thresholds 1..255, NaN...]: index 256 never steps up
constexpr int THR_WORDS = 260;
__device__ __forceinline__ void load_to_byte_tables(float *s_thr, unsigned *s_slice, int tid, int nthreads) {
    for (int i = tid; i < THR_WORDS; i += nthreads)
        s_thr[i] = __uint_as_float(i < 256 ? kToByteThreshold[i] : 0x7fc00000u);
    for (int i = tid; i < SLICE_WORDS; i += nthreads) s_slice[i] = kToByteSlice[i];
}

// src/lib.rs:93-99 + src/core/imageio.rs:66-68: byte = clamp(255 * gamma_correct(v) + .5, 0, 255) as u8
__device__ __forceinline__ unsigned to_byte(float v, const float *thr, const unsigned *slice) {
    // below 2^-13 (also negative, and NaN: `as u8` maps NaN to 0) every input gives 0, above 1.0 every input 255
    const float vc = fminf(fmaxf(v, 1.220703125e-4f), 1.f);
    const unsigned idx = (__float_as_uint(vc) >> SLICE_SHIFT) - (SLICE_FIRST_BITS >> SLICE_SHIFT);
    const unsigned k0 = reinterpret_cast<const unsigned char *>(slice)[idx];
    return k0 + (vc >= thr[k0 + 1]);
}

constexpr int RES_PIX = 256;  // resolve_to_frames_kernel: pixels per block; 256*12 B = 3072 B, a multiple of 16
constexpr int RES_THREADS = 256;

// PPT pixels per thread, pixel j of thread t = base + j*RES_THREADS + t: coalesced float4 loads, PPT of them in flight.
template <bool BYTES, int PPT>
__global__ void __launch_bounds__(RES_THREADS) resolve_kernel(const float4 *__restrict__ xyzw,
                                                              const float *__restrict__ splat, long long npix,
                                                              float splat_scale, float scale, void *__restrict__ out) {
    constexpr int TILE = RES_THREADS * PPT;
    __shared__ __align__(16) float s_in[TILE * 3];
    __shared__ __align__(16) float s_out[BYTES ? TILE * 3 / 4 : TILE * 3];
    __shared__ float s_thr[BYTES ? THR_WORDS : 1];
    __shared__ unsigned s_slice[BYTES ? SLICE_WORDS : 1];
    const long long base = (long long)blockIdx.x * TILE;
    const int n = (int)min((long long)TILE, npix - base);
    const int tid = threadIdx.x;
    const bool full = n == TILE;
    // 28 B / pixel in: float4 xyzw straight to registers, splat through shared memory
    float4 p[PPT];
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
        p[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (full || j * RES_THREADS + tid < n) p[j] = pb::ldg_stream(&xyzw[base + j * RES_THREADS + tid]);
    }
    if (full) {
        const float4 *src = reinterpret_cast<const float4 *>(splat + base * 3);
#pragma unroll
        for (int i = 0; i < (TILE * 3 / 4 + RES_THREADS - 1) / RES_THREADS; ++i)
            if (i * RES_THREADS + tid < TILE * 3 / 4)
                reinterpret_cast<float4 *>(s_in)[i * RES_THREADS + tid] = pb::ldg_stream(&src[i * RES_THREADS + tid]);
    } else {
        for (int i = tid; i < n * 3; i += RES_THREADS) s_in[i] = splat[base * 3 + i];
    }
    if (BYTES) load_to_byte_tables(s_thr, s_slice, tid, RES_THREADS);
    __syncthreads();
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
        const int q = j * RES_THREADS + tid;
        if (full || q < n) {
            float r, g, b;
            resolve_pixel(p[j], s_in[3 * q], s_in[3 * q + 1], s_in[3 * q + 2], splat_scale, scale, r, g, b);
            if (BYTES) {
                unsigned char *o = reinterpret_cast<unsigned char *>(s_out);
                o[3 * q] = (unsigned char)to_byte(r, s_thr, s_slice);
                o[3 * q + 1] = (unsigned char)to_byte(g, s_thr, s_slice);
                o[3 * q + 2] = (unsigned char)to_byte(b, s_thr, s_slice);
            } else {
                s_out[3 * q] = r; s_out[3 * q + 1] = g; s_out[3 * q + 2] = b;
            }
        }
    }
    __syncthreads();
    if (BYTES) {
        unsigned char *dst = reinterpret_cast<unsigned char *>(out) + base * 3;
        if (full) {
            if (tid < TILE * 3 / 16) reinterpret_cast<uint4 *>(dst)[tid] = reinterpret_cast<uint4 *>(s_out)[tid];
        } else {
            const unsigned char *o = reinterpret_cast<const unsigned char *>(s_out);
            for (int i = tid; i < n * 3; i += RES_THREADS) dst[i] = o[i];
        }
    } else {
        float *dst = reinterpret_cast<float *>(out) + base * 3;
        if (full) {
#pragma unroll
            for (int i = 0; i < (TILE * 3 / 4 + RES_THREADS - 1) / RES_THREADS; ++i)
                if (i * RES_THREADS + tid < TILE * 3 / 4)
                    pb::stg_stream(&reinterpret_cast<float4 *>(dst)[i * RES_THREADS + tid],
                                   reinterpret_cast<float4 *>(s_out)[i * RES_THREADS + tid]);
        } else {
            for (int i = tid; i < n * 3; i += RES_THREADS) dst[i] = s_out[i];
        }
    }
}

template <bool BYTES>
static int resolve_impl(const PbrtFilm *f, float splat_scale, void *out, int dst_is_device) {
    if (!f || !out) return fail(PBRT_E_INVALID, "null argument");
    if (f->npix == 0) return PBRT_OK;
    const size_t bytes = (size_t)f->npix * 3 * (BYTES ? 1 : sizeof(float));
    void *d_out = out;
    const bool to_host = dst_is_device != PBRT_MEM_DEVICE;
    if (to_host) {
        if (int rc = pb::out_stage(bytes, &d_out)) return rc;
    } else if (((uintptr_t)out & 15) != 0) {
        return fail(PBRT_E_INVALID, "device output must be 16-byte aligned");
    }
    if (BYTES) {
        static bool slices_ready = false;  // one device per process (pbrt_b200_init)
        if (!slices_ready) {
            to_byte_slices_kernel<<<1, 256, 0, ctx().stream>>>();
            PB_LAUNCH_CHECK("to_byte_slices_kernel");
            slices_ready = true;
        }
    }
    static int ppt_env = getenv("PBRT_B200_RES_PPT") ? atoi(getenv("PBRT_B200_RES_PPT")) : 0;
    if (ppt_env != 0 && ppt_env != 1 && ppt_env != 2 && ppt_env != 4)
        return fail(PBRT_E_INVALID, "PBRT_B200_RES_PPT=%d: pixels per thread must be 1, 2 or 4", ppt_env);
    const int PPT = ppt_env ? ppt_env : (BYTES ? 4 : 2);
    int blocks = (int)((f->npix + RES_THREADS * PPT - 1) / (RES_THREADS * PPT));
#define RES_LAUNCH(N) resolve_kernel<BYTES, N><<<blocks, RES_THREADS, 0, ctx().stream>>>(f->d_xyzw, f->d_splat, (long long)f->npix, splat_scale, f->scale, d_out)
    if (PPT == 1) RES_LAUNCH(1); else if (PPT == 2) RES_LAUNCH(2); else RES_LAUNCH(4);
    PB_LAUNCH_CHECK("resolve_kernel");
    if (to_host) {
        PB_CUDA(cudaMemcpyAsync(out, d_out, bytes, cudaMemcpyDeviceToHost, ctx().stream));
        // PBRT_MEM_PINNED_ASYNC: the read-back is only enqueued; pbrt_b200_synchronize() completes it
        if (dst_is_device != PBRT_MEM_PINNED_ASYNC) PB_CUDA(cudaStreamSynchronize(ctx().stream));
    }
    return PBRT_OK;
}

extern "C" int pbrt_film_resolve_rgb(const PbrtFilm *f, float splat_scale, float *out, int dst_is_device) {
    PB_API_LOCK;
    return resolve_impl<false>(f, splat_scale, out, dst_is_device);
}
extern "C" int pbrt_film_resolve_rgb8(const PbrtFilm *f, float splat_scale, uint8_t *out, int dst_is_device) {
    PB_API_LOCK;
    return resolve_impl<true>(f, splat_scale, out, dst_is_device);
}

// resolve fused with the all-gather of a row-sharded film: every block stores its 256 resolved pixels
// into each rank's full frame (local HBM for this rank, NVLink peer stores for the others)
constexpr int MAX_FRAMES = 16;
struct FrameList {
    float *p[MAX_FRAMES];
    int n;
};

// PPT pixels per thread (pixel j of thread t = base + j*RES_PIX + t), as resolve_kernel: PPT film loads in flight per thread
template <int PPT>
__global__ void __launch_bounds__(RES_PIX) resolve_to_frames_kernel(const float4 *__restrict__ xyzw,
                                                                    const float *__restrict__ splat, long long npix,
                                                                    float splat_scale, float scale, FrameList frames,
                                                                    long long frame_offset_px) {
    constexpr int TILE = RES_PIX * PPT;
    __shared__ __align__(16) float s_in[TILE * 3];
    __shared__ __align__(16) float s_out[TILE * 3];
    const long long base = (long long)blockIdx.x * TILE;
    const int n = (int)min((long long)TILE, npix - base);
    const int tid = threadIdx.x;
    const bool full = n == TILE;
    float4 p[PPT];
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
        p[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (full || j * RES_PIX + tid < n) p[j] = pb::ldg_stream(&xyzw[base + j * RES_PIX + tid]);
    }
    if (full) {
        const float4 *src = reinterpret_cast<const float4 *>(splat + base * 3);
#pragma unroll
        for (int i = 0; i < (TILE * 3 / 4 + RES_PIX - 1) / RES_PIX; ++i)
            if (i * RES_PIX + tid < TILE * 3 / 4)
                reinterpret_cast<float4 *>(s_in)[i * RES_PIX + tid] = pb::ldg_stream(&src[i * RES_PIX + tid]);
    } else {
        for (int i = tid; i < n * 3; i += RES_PIX) s_in[i] = splat[base * 3 + i];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
        const int q = j * RES_PIX + tid;
        if (full || q < n) {
            float r, g, b;
            resolve_pixel(p[j], s_in[3 * q], s_in[3 * q + 1], s_in[3 * q + 2], splat_scale, scale, r, g, b);
            s_out[3 * q] = r; s_out[3 * q + 1] = g; s_out[3 * q + 2] = b;
        }
    }
    __syncthreads();
    const long long o = (frame_offset_px + base) * 3;  // float index in a frame; 16-byte aligned for full blocks
    for (int f = 0; f < frames.n; ++f) {
        float *dst = frames.p[f] + o;
        if (full && ((o & 3) == 0)) {
#pragma unroll
            for (int i = 0; i < (TILE * 3 / 4 + RES_PIX - 1) / RES_PIX; ++i)
                if (i * RES_PIX + tid < TILE * 3 / 4)
                    reinterpret_cast<float4 *>(dst)[i * RES_PIX + tid] = reinterpret_cast<float4 *>(s_out)[i * RES_PIX + tid];
        } else {
            for (int i = tid; i < n * 3; i += RES_PIX) dst[i] = s_out[i];
        }
    }
}

extern "C" int pbrt_film_resolve_rgb_to_frames(const PbrtFilm *f, float splat_scale, int32_t nframes, void *const *frames) {
    PB_API_LOCK;
    if (!f || !frames) return fail(PBRT_E_INVALID, "null argument");
    if (nframes < 1 || nframes > MAX_FRAMES) return fail(PBRT_E_INVALID, "between 1 and %d frames", MAX_FRAMES);
    if (f->npix == 0) return PBRT_OK;
    if (f->owned.x0 != f->cropped.x0 || f->owned.x1 != f->cropped.x1)
        return fail(PBRT_E_UNSUPPORTED, "row shards must span the full width");
    FrameList fl;
    fl.n = nframes;
    for (int i = 0; i < nframes; ++i) {
        if (!frames[i] || ((uintptr_t)frames[i] & 15)) return fail(PBRT_E_INVALID, "frame %d null or not 16-byte aligned", i);
        fl.p[i] = (float *)frames[i];
    }
    const long long off = (long long)(f->owned.y0 - f->cropped.y0) * pb::bw(f->cropped);
    constexpr int PPT = 2;
    int blocks = (int)((f->npix + RES_PIX * PPT - 1) / (RES_PIX * PPT));
    resolve_to_frames_kernel<PPT><<<blocks, RES_PIX, 0, ctx().stream>>>(f->d_xyzw, f->d_splat, (long long)f->npix, splat_scale,
                                                                        f->scale, fl, off);
    PB_LAUNCH_CHECK("resolve_to_frames_kernel");
    return PBRT_OK;
}

extern "C" int pbrt_film_get_pixel_xyz(const PbrtFilm *f, int32_t x, int32_t y, float out[3]) {
    PB_API_LOCK;
    if (!f || !out) return fail(PBRT_E_INVALID, "null argument");
    if (x < f->owned.x0 || x >= f->owned.x1 || y < f->owned.y0 || y >= f->owned.y1)
        return fail(PBRT_E_RANGE, "p [%d, %d] outside film", x, y);  // film.rs:391-396
    size_t off = (size_t)(y - f->owned.y0) * pb::bw(f->owned) + (x - f->owned.x0);  // film.rs:397-401
    float4 p;
    PB_CUDA(cudaMemcpyAsync(&p, f->d_xyzw + off, sizeof p, cudaMemcpyDeviceToHost, ctx().stream));
    PB_CUDA(cudaStreamSynchronize(ctx().stream));
    out[0] = p.x; out[1] = p.y; out[2] = p.z;
    return PBRT_OK;
}

__global__ void read_pixels_kernel(const float4 *__restrict__ xyzw, const float *__restrict__ splat, long long npix,
                                   float *__restrict__ out7) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    float4 p = xyzw[i];
    float *o = out7 + i * 7;
    o[0] = p.x; o[1] = p.y; o[2] = p.z; o[3] = p.w;
    o[4] = splat[3 * i]; o[5] = splat[3 * i + 1]; o[6] = splat[3 * i + 2];
}

extern "C" int pbrt_film_read_pixels(const PbrtFilm *f, float *out7, int dst_is_device) {
    PB_API_LOCK;
    if (!f || !out7) return fail(PBRT_E_INVALID, "null argument");
    if (f->npix == 0) return PBRT_OK;
    size_t bytes = (size_t)f->npix * 7 * sizeof(float);
    void *d_out = out7;
    if (!dst_is_device)
        if (int rc = pb::out_stage(bytes, &d_out)) return rc;
    read_pixels_kernel<<<(unsigned)((f->npix + 255) / 256), 256, 0, ctx().stream>>>(f->d_xyzw, f->d_splat,
                                                                                     (long long)f->npix, (float *)d_out);
    PB_LAUNCH_CHECK("read_pixels_kernel");
    if (!dst_is_device) {
        PB_CUDA(cudaMemcpyAsync(out7, d_out, bytes, cudaMemcpyDeviceToHost, ctx().stream));
        PB_CUDA(cudaStreamSynchronize(ctx().stream));
    }
    return PBRT_OK;
}

extern "C" int pbrt_film_device_buffers(const PbrtFilm *f, void **xyzw, void **splat, int64_t *npixels) {
    if (!f) return fail(PBRT_E_INVALID, "null film");
    if (xyzw) *xyzw = f->d_xyzw;
    if (splat) *splat = f->d_splat;
    if (npixels) *npixels = f->npix;
    return PBRT_OK;
}

extern "C" int pbrt_film_check(PbrtFilm *f) {
    PB_API_LOCK;
    if (!f) return fail(PBRT_E_INVALID, "null film");
    int e = 0;
    PB_CUDA(cudaMemcpyAsync(&e, f->d_err, sizeof e, cudaMemcpyDeviceToHost, ctx().stream));
    PB_CUDA(cudaStreamSynchronize(ctx().stream));
    if (e) {
        PB_CUDA(cudaMemsetAsync(f->d_err, 0, sizeof(int), ctx().stream));
        if (e & pb::ERRBIT_NOT_PIXEL_MAJOR)
            return fail(PBRT_E_NOT_PIXEL_MAJOR, "add_samples_tile: a sample lies outside its nominal pixel; film contents are undefined");
        if (e & pb::ERRBIT_NONFINITE)
            return fail(PBRT_E_NONFINITE, "add_samples_tile: non-finite radiance; film contents are undefined");
        return fail(PBRT_E_CUDA, "asynchronous kernel error word %d", e);
    }
    return PBRT_OK;
}

// ===================================================================== kernels: arbitrary-order samples, splats, set_image

// [T2] scatter with global float atomics into a scratch tile (no ordering contract)
__global__ void __launch_bounds__(256) scatter_samples_kernel(float4 *__restrict__ tile, Bounds tb, unsigned long long n,
                                                              const float2 *__restrict__ xy,
                                                              const float4 *__restrict__ rgbw,
                                                              const float *__restrict__ table, float rx, float ry,
                                                              float irx, float iry, float max_lum, int *__restrict__ err) {
    __shared__ float s_table[256];
    s_table[threadIdx.x] = table[threadIdx.x];
    __syncthreads();
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        float2 p = pb::ldg_stream(&xy[i]);
        float4 L = pb::ldg_stream(&rgbw[i]);
        float ly = pb::luminance(L.x, L.y, L.z);
        if (ly > max_lum) {
            float s = max_lum / ly;
            L.x *= s; L.y *= s; L.z *= s;
        }
        float dx = p.x - 0.5f, dy = p.y - 0.5f;
        int p0x = max(__float2int_ru(dx - rx), tb.x0), p0y = max(__float2int_ru(dy - ry), tb.y0);
        int p1x = min(__float2int_rd(dx + rx) + 1, tb.x1), p1y = min(__float2int_rd(dy + ry) + 1, tb.y1);
        float cr = L.x * L.w, cg = L.y * L.w, cb = L.z * L.w;
        const float z = cr * 0.f + cg * 0.f + cb * 0.f + dx * 0.f + dy * 0.f;  // NaN iff anything is inf or NaN
        if (z != z) {
            atomicOr(err, pb::ERRBIT_NONFINITE);
            continue;
        }
        const int tw = tb.x1 - tb.x0;
        for (int y = p0y; y < p1y; ++y) {
            int iy = min(__float2int_rd(fabsf(((float)y - dy) * iry * 16.f)), 15);
            for (int x = p0x; x < p1x; ++x) {
                int ix = min(__float2int_rd(fabsf(((float)x - dx) * irx * 16.f)), 15);
                float w = s_table[iy * 16 + ix];
                float *px = reinterpret_cast<float *>(&tile[(size_t)(y - tb.y0) * tw + (x - tb.x0)]);
                atomicAdd(px + 0, cr * w);
                atomicAdd(px + 1, cg * w);
                atomicAdd(px + 2, cb * w);
                atomicAdd(px + 3, w);
            }
        }
    }
}

extern "C" int pbrt_film_add_samples(PbrtFilm *f, const int32_t sbv[4], uint64_t n, const float *xy, const float *rgbw,
                                     int src_is_device) {
    PB_API_LOCK;
    if (int rc = host_or_device(src_is_device, "add_samples")) return rc;
    if (!f || !sbv) return fail(PBRT_E_INVALID, "null argument");
    Bounds tb;
    if (int rc = tile_bounds_impl(f, sbv, f->owned, &tb)) return rc;
    if (n == 0 || tb.x1 <= tb.x0 || tb.y1 <= tb.y0) return PBRT_OK;
    if (!xy || !rgbw) return fail(PBRT_E_INVALID, "null sample stream");
    const size_t px = (size_t)pb::bw(tb) * pb::bh(tb);
    if (px > f->scratch_tile_px) {
        cudaFree(f->d_scratch_tile);
        f->d_scratch_tile = nullptr;
        f->scratch_tile_px = 0;
        PB_CUDA(cudaMalloc(&f->d_scratch_tile, px * sizeof(float4)));
        f->scratch_tile_px = px;
    }
    const float2 *d_xy = (const float2 *)xy;
    const float4 *d_rgbw = (const float4 *)rgbw;
    if (!src_is_device) {
        void *a, *b;
        if (int rc = pb::stage_in(f, 0, xy, n * sizeof(float2), &a)) return rc;
        if (int rc = pb::stage_in(f, 1, rgbw, n * sizeof(float4), &b)) return rc;
        d_xy = (const float2 *)a; d_rgbw = (const float4 *)b;
    }
    cudaStream_t s = ctx().stream;
    PB_CUDA(cudaMemsetAsync(f->d_scratch_tile, 0, px * sizeof(float4), s));  // FilmTilePixel::default()
    int blocks = (int)std::min<uint64_t>((n + 255) / 256, (uint64_t)ctx().sm_count * 16);
    scatter_samples_kernel<<<blocks, 256, 0, s>>>(f->d_scratch_tile, tb, n, d_xy, d_rgbw, f->d_table, f->radius[0],
                                                  f->radius[1], f->inv_radius[0], f->inv_radius[1], f->max_lum, f->d_err);
    PB_LAUNCH_CHECK("scatter_samples_kernel");
    dim3 grid((pb::bw(tb) + 255) / 256, std::min(pb::bh(tb), 4096));
    merge_tile_kernel<<<grid, 256, 0, s>>>(f->d_xyzw, f->owned, tb, f->d_scratch_tile);
    PB_LAUNCH_CHECK("merge_tile_kernel");
    return PBRT_OK;
}

// [T2] pbrt-v3 Film::AddSplat; float atomics stand in for AtomicFloat (src/core/parallel.rs:85-99)
__global__ void __launch_bounds__(256) add_splats_kernel(float *__restrict__ splat, Bounds owned, unsigned long long n,
                                                         const float2 *__restrict__ xy, const float *__restrict__ rgb,
                                                         float max_lum) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float2 p = xy[i];
    float r = rgb[3 * i], g = rgb[3 * i + 1], b = rgb[3 * i + 2];
    if (r != r || g != g || b != b) return;
    float ly = pb::luminance(r, g, b);
    if (ly < 0.f || isinf(ly)) return;
    int ix = __float2int_rz(p.x), iy = __float2int_rz(p.y);
    if (p.x != p.x) ix = 0;
    if (p.y != p.y) iy = 0;
    if (ix < owned.x0 || ix >= owned.x1 || iy < owned.y0 || iy >= owned.y1) return;
    if (ly > max_lum) {
        float s = max_lum / ly;
        r *= s; g *= s; b *= s;
    }
    float X, Y, Z;
    pb::rgb_to_xyz(r, g, b, X, Y, Z);
    float *o = splat + 3 * ((size_t)(iy - owned.y0) * (owned.x1 - owned.x0) + (ix - owned.x0));
    atomicAdd(o, X); atomicAdd(o + 1, Y); atomicAdd(o + 2, Z);
}

extern "C" int pbrt_film_add_splats(PbrtFilm *f, uint64_t n, const float *xy, const float *rgb, int src_is_device) {
    PB_API_LOCK;
    if (int rc = host_or_device(src_is_device, "add_splats")) return rc;
    if (!f) return fail(PBRT_E_INVALID, "null film");
    if (n == 0 || f->npix == 0) return PBRT_OK;
    if (!xy || !rgb) return fail(PBRT_E_INVALID, "null argument");
    const float2 *d_xy = (const float2 *)xy;
    const float *d_rgb = rgb;
    if (!src_is_device) {
        void *a, *b;
        if (int rc = pb::stage_in(f, 0, xy, n * sizeof(float2), &a)) return rc;
        if (int rc = pb::stage_in(f, 1, rgb, n * 3 * sizeof(float), &b)) return rc;
        d_xy = (const float2 *)a; d_rgb = (const float *)b;
    }
    add_splats_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx().stream>>>(f->d_splat, f->owned, n, d_xy, d_rgb, f->max_lum);
    PB_LAUNCH_CHECK("add_splats_kernel");
    return PBRT_OK;
}

// [T2] pbrt-v3 Film::SetImage: xyz = to_xyz(img), weight = 1, splat = 0
__global__ void __launch_bounds__(256) set_image_kernel(float4 *__restrict__ xyzw, float *__restrict__ splat,
                                                        long long npix, const float *__restrict__ rgb) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    float X, Y, Z;
    pb::rgb_to_xyz(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2], X, Y, Z);
    xyzw[i] = make_float4(X, Y, Z, 1.f);
    splat[3 * i] = 0.f; splat[3 * i + 1] = 0.f; splat[3 * i + 2] = 0.f;
}

extern "C" int pbrt_film_set_image(PbrtFilm *f, const float *rgb, int src_is_device) {
    PB_API_LOCK;
    if (int rc = host_or_device(src_is_device, "set_image")) return rc;
    if (!f || !rgb) return fail(PBRT_E_INVALID, "null argument");
    if (f->npix == 0) return PBRT_OK;
    const float *d_rgb = rgb;
    if (!src_is_device) {
        void *a;
        if (int rc = pb::stage_in(f, 0, rgb, (size_t)f->npix * 3 * sizeof(float), &a)) return rc;
        d_rgb = (const float *)a;
    }
    set_image_kernel<<<(unsigned)((f->npix + 255) / 256), 256, 0, ctx().stream>>>(f->d_xyzw, f->d_splat, (long long)f->npix, d_rgb);
    PB_LAUNCH_CHECK("set_image_kernel");
    return PBRT_OK;
}

// [T2] pbrt-v3 Film::Clear
extern "C" int pbrt_film_clear(PbrtFilm *f) {
    PB_API_LOCK;
    if (!f) return fail(PBRT_E_INVALID, "null film");
    if (f->npix == 0) return PBRT_OK;
    PB_CUDA(cudaMemsetAsync(f->d_xyzw, 0, (size_t)f->npix * sizeof(float4), ctx().stream));
    PB_CUDA(cudaMemsetAsync(f->d_splat, 0, (size_t)f->npix * 3 * sizeof(float), ctx().stream));
    return PBRT_OK;
}

// {r, g, b} + optional weight stream -> the {r, g, b, sample_weight} records the splat kernels read
__global__ void pack_rgbw_kernel(const float *__restrict__ rgb, const float *__restrict__ sw, size_t n,
                                 float4 *__restrict__ out) {
    // 4 samples per thread: three float4 loads of rgb (48 B), one of sw, four float4 stores
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t i = g * 4;
    if (i + 4 <= n && (((uintptr_t)rgb | (uintptr_t)sw) & 15) == 0) {
        const float4 a = pb::ldg_stream(reinterpret_cast<const float4 *>(rgb) + 3 * g);
        const float4 b = pb::ldg_stream(reinterpret_cast<const float4 *>(rgb) + 3 * g + 1);
        const float4 c = pb::ldg_stream(reinterpret_cast<const float4 *>(rgb) + 3 * g + 2);
        float4 w = make_float4(1.f, 1.f, 1.f, 1.f);
        if (sw) w = pb::ldg_stream(reinterpret_cast<const float4 *>(sw) + g);
        out[i] = make_float4(a.x, a.y, a.z, w.x);
        out[i + 1] = make_float4(a.w, b.x, b.y, w.y);
        out[i + 2] = make_float4(b.z, b.w, c.x, w.z);
        out[i + 3] = make_float4(c.y, c.z, c.w, w.w);
    } else {
        for (size_t k = i; k < n && k < i + 4; ++k)
            out[k] = make_float4(rgb[3 * k], rgb[3 * k + 1], rgb[3 * k + 2], sw ? sw[k] : 1.f);
    }
}

// grow-only device buffer `k` of staging set `set`
static int pipe_buffer(PbrtFilm *f, int set, int k, size_t bytes) {
    if (bytes > f->pipe_bytes[set][k]) {
        cudaFree(f->d_pipe[set][k]);  // synchronises the device: nothing is using the old buffer
        f->d_pipe[set][k] = nullptr;
        f->pipe_bytes[set][k] = 0;
        PB_CUDA(cudaMalloc(&f->d_pipe[set][k], bytes + 256));
        f->pipe_bytes[set][k] = bytes + 256;
    }
    return PBRT_OK;
}

// [T2] the pixel-major splat: bounds on the host, kernels in splat.cu.  The radiance arrives either as one
// {r,g,b,sample_weight} stream (rgbw) or as separate rgb / optional weight streams (rgb3, sw).
static int add_samples_tile_impl(PbrtFilm *f, const int32_t sbv[4], int32_t spp, const float *xy, const float *rgbw,
                                 const float *rgb3, const float *sw, int src_is_device, int mode) {
    if (!f || !sbv) return fail(PBRT_E_INVALID, "null argument");
    if (spp < 1) return fail(PBRT_E_INVALID, "spp must be >= 1");
    if (mode < PBRT_SPLAT_EXACT || mode > PBRT_SPLAT_ATOMIC) return fail(PBRT_E_INVALID, "unknown splat mode %d", mode);
    if (src_is_device < PBRT_MEM_HOST || src_is_device > PBRT_MEM_PINNED_ASYNC)
        return fail(PBRT_E_INVALID, "unknown memory kind %d", src_is_device);
    Bounds sb{sbv[0], sbv[1], sbv[2], sbv[3]};
    Bounds tb;
    if (int rc = tile_bounds_impl(f, sbv, f->owned, &tb)) return rc;
    if (sb.x1 <= sb.x0 || sb.y1 <= sb.y0) return PBRT_OK;          // no samples
    if (tb.x1 <= tb.x0 || tb.y1 <= tb.y0) return PBRT_OK;          // tile misses the film
    const bool split = rgb3 != nullptr;
    if (!xy || (!rgbw && !rgb3)) return fail(PBRT_E_INVALID, "null sample stream");
    const size_t n = (size_t)pb::bw(sb) * pb::bh(sb) * (size_t)spp;
    // host streams: [0] xy, [1] rgbw, [2] rgb, [3] weights
    const void *src[4] = {xy, rgbw, rgb3, sw};
    const size_t need[4] = {n * sizeof(float2), (split || rgbw) ? n * sizeof(float4) : 0, split ? n * 3 * sizeof(float) : 0,
                            (split && sw) ? n * sizeof(float) : 0};
    const void *dev[4] = {xy, rgbw, rgb3, sw};
    int set = 0;
    const bool async = src_is_device == PBRT_MEM_PINNED_ASYNC;
    if ((async || split) && !f->pipe_ready) {
        // the staging sets are about to be used for the first time: their events must exist before any kernel reads them
        for (int i = 0; i < 2; ++i) {
            PB_CUDA(cudaEventCreateWithFlags(&f->ev_staged[i], cudaEventDisableTiming));
            PB_CUDA(cudaEventCreateWithFlags(&f->ev_consumed[i], cudaEventDisableTiming));
        }
        f->pipe_ready = true;
    }
    if (async) {
        // upload on the copy stream into the staging set the previous-but-one call used; the kernel waits for
        // the upload, the next upload into this set waits for the kernel
        set = f->pipe_turn;
        f->pipe_turn ^= 1;
        for (int k = 0; k < 4; ++k)
            if (need[k])
                if (int rc = pipe_buffer(f, set, k, need[k])) return rc;
        cudaStream_t cs = ctx().copy_stream;
        PB_CUDA(cudaStreamWaitEvent(cs, f->ev_consumed[set], 0));
        for (int k = 0; k < 4; ++k) {
            if (!need[k] || !src[k]) continue;
            PB_CUDA(cudaMemcpyAsync(f->d_pipe[set][k], src[k], need[k], cudaMemcpyHostToDevice, cs));
            dev[k] = f->d_pipe[set][k];
        }
        PB_CUDA(cudaEventRecord(f->ev_staged[set], cs));
        PB_CUDA(cudaStreamWaitEvent(ctx().stream, f->ev_staged[set], 0));
    } else if (src_is_device == PBRT_MEM_HOST) {
        // synchronous-in-stream staging; slot 0 = xy, slot 1 = radiance (rgbw, or rgb followed by the weights)
        void *a, *b;
        if (int rc = pb::stage_in(f, 0, xy, need[0], &a)) return rc;
        dev[0] = a;
        if (!split) {
            if (int rc = pb::stage_in(f, 1, rgbw, need[1], &b)) return rc;
            dev[1] = b;
        } else {
            const size_t off = (need[2] + 255) & ~(size_t)255;
            if (int rc = pb::stage_in(f, 1, nullptr, off + need[3], &b)) return rc;  // reserve, no copy
            PB_CUDA(cudaMemcpyAsync(b, rgb3, need[2], cudaMemcpyHostToDevice, ctx().stream));
            dev[2] = b;
            if (sw) {
                PB_CUDA(cudaMemcpyAsync((char *)b + off, sw, need[3], cudaMemcpyHostToDevice, ctx().stream));
                dev[3] = (char *)b + off;
            }
        }
    } else if (((uintptr_t)xy & 7) || ((uintptr_t)rgbw & 15) || ((uintptr_t)rgb3 & 3) || ((uintptr_t)sw & 3)) {
        return fail(PBRT_E_INVALID, "device sample streams must be 8- (xy), 16- (rgbw) and 4-byte (rgb, weights) aligned");
    }
    if (split) {
        // interleave on the device (28 B/sample of HBM traffic, ~0.1 ms for 33 M samples) so that every mode runs
        // the same splat kernels
        if (!async) set = 0;
        if (int rc = pipe_buffer(f, set, 1, n * sizeof(float4))) return rc;
        const size_t threads = (n + 3) / 4;
        pack_rgbw_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, ctx().stream>>>(
            (const float *)dev[2], (const float *)dev[3], n, (float4 *)f->d_pipe[set][1]);
        PB_LAUNCH_CHECK("pack_rgbw_kernel");
        dev[1] = f->d_pipe[set][1];
    }
    int rc = pb::launch_splat_tile(f, sb, tb, spp, (const float2 *)dev[0], (const float4 *)dev[1], mode);
    // also after a synchronous call that borrowed set 0's interleave buffer: the next upload into it must wait
    if (async || split) PB_CUDA(cudaEventRecord(f->ev_consumed[set], ctx().stream));
    return rc;
}

extern "C" int pbrt_film_add_samples_tile(PbrtFilm *f, const int32_t sbv[4], int32_t spp, const float *xy,
                                          const float *rgbw, int src_is_device, int mode) {
    PB_API_LOCK;
    if (!rgbw) return fail(PBRT_E_INVALID, "null sample stream");
    return add_samples_tile_impl(f, sbv, spp, xy, rgbw, nullptr, nullptr, src_is_device, mode);
}

extern "C" int pbrt_film_add_samples_tile_rgb(PbrtFilm *f, const int32_t sbv[4], int32_t spp, const float *xy,
                                              const float *rgb, const float *sample_weight, int src_is_device,
                                              int mode) {
    PB_API_LOCK;
    if (!rgb) return fail(PBRT_E_INVALID, "null sample stream");
    return add_samples_tile_impl(f, sbv, spp, xy, nullptr, rgb, sample_weight, src_is_device, mode);
}

// [T2] many tiles per call: splat every tile into its own RGBW buffer, then the ordered batched merge
extern "C" int pbrt_film_add_samples_tiles(PbrtFilm *f, int32_t ntiles, const int32_t *sbs, const int64_t *sample_offsets,
                                           int32_t spp, const float *xy, const float *rgbw, int64_t total_samples,
                                           int src_is_device, int mode) {
    PB_API_LOCK;
    if (int rc = host_or_device(src_is_device, "add_samples_tiles")) return rc;
    if (!f) return fail(PBRT_E_INVALID, "null film");
    if (ntiles <= 0) return PBRT_OK;
    if (!sbs || !sample_offsets || !xy || !rgbw) return fail(PBRT_E_INVALID, "null argument");
    if (spp < 1) return fail(PBRT_E_INVALID, "spp must be >= 1");
    if (mode != PBRT_SPLAT_EXACT && mode != PBRT_SPLAT_FMA) return fail(PBRT_E_INVALID, "batched tiles support PBRT_SPLAT_EXACT and PBRT_SPLAT_FMA");
    std::vector<pb::SplatTile> desc(ntiles);
    std::vector<int32_t> tbs((size_t)ntiles * 4);
    std::vector<int64_t> poffs(ntiles);
    int64_t total_px = 0;
    int max_w = 0, max_h = 0;
    for (int i = 0; i < ntiles; ++i) {
        Bounds sb{sbs[4 * i], sbs[4 * i + 1], sbs[4 * i + 2], sbs[4 * i + 3]};
        Bounds tb;
        if (int rc = tile_bounds_impl(f, &sbs[4 * i], f->owned, &tb)) return rc;
        const bool empty = sb.x1 <= sb.x0 || sb.y1 <= sb.y0 || tb.x1 <= tb.x0 || tb.y1 <= tb.y0;
        if (empty) tb = Bounds{0, 0, 0, 0};  // a tile without samples merges zeros: nothing to do
        const int64_t ns = empty ? 0 : (int64_t)pb::bw(sb) * pb::bh(sb) * spp;
        if (sample_offsets[i] < 0 || sample_offsets[i] + ns > total_samples)
            return fail(PBRT_E_INVALID, "tile %d: samples outside the stream", i);
        desc[i].sb = sb; desc[i].tb = tb;
        desc[i].sample_offset = sample_offsets[i];
        desc[i].pixel_offset = total_px;
        poffs[i] = total_px;
        tbs[4 * i] = tb.x0; tbs[4 * i + 1] = tb.y0; tbs[4 * i + 2] = tb.x1; tbs[4 * i + 3] = tb.y1;
        total_px += (int64_t)pb::bw(tb) * pb::bh(tb);
        max_w = std::max(max_w, pb::bw(tb)); max_h = std::max(max_h, pb::bh(tb));
    }
    if (total_px == 0) return PBRT_OK;
    const float2 *d_xy = (const float2 *)xy;
    const float4 *d_rgbw = (const float4 *)rgbw;
    if (!src_is_device) {
        void *a, *b;
        if (int rc = pb::stage_in(f, 0, xy, (size_t)total_samples * sizeof(float2), &a)) return rc;
        if (int rc = pb::stage_in(f, 1, rgbw, (size_t)total_samples * sizeof(float4), &b)) return rc;
        d_xy = (const float2 *)a; d_rgbw = (const float4 *)b;
    }
    if ((size_t)total_px > f->scratch_tile_px) {
        cudaFree(f->d_scratch_tile);
        f->d_scratch_tile = nullptr;
        f->scratch_tile_px = 0;
        PB_CUDA(cudaMalloc(&f->d_scratch_tile, (size_t)total_px * sizeof(float4)));
        f->scratch_tile_px = (size_t)total_px;
    }
    const size_t dbytes = desc.size() * sizeof(pb::SplatTile);
    // descriptors go through a page-locked staging copy owned by the film, so the upload is asynchronous; the event
    // tells the next call when it may overwrite the staging copy (normally long since)
    if (f->tile_desc_event) PB_CUDA(cudaEventSynchronize(f->ev_tile_desc));
    if (dbytes > f->tile_desc_bytes) {
        cudaFree(f->d_tile_desc);
        if (f->h_tile_desc) cudaFreeHost(f->h_tile_desc);
        f->d_tile_desc = f->h_tile_desc = nullptr;
        f->tile_desc_bytes = 0;
        PB_CUDA(cudaMalloc(&f->d_tile_desc, dbytes + 256));
        PB_CUDA(cudaMallocHost(&f->h_tile_desc, dbytes + 256));
        f->tile_desc_bytes = dbytes + 256;
    }
    if (!f->tile_desc_event) {
        PB_CUDA(cudaEventCreateWithFlags(&f->ev_tile_desc, cudaEventDisableTiming));
        f->tile_desc_event = true;
    }
    memcpy(f->h_tile_desc, desc.data(), dbytes);
    PB_CUDA(cudaMemcpyAsync(f->d_tile_desc, f->h_tile_desc, dbytes, cudaMemcpyHostToDevice, ctx().stream));
    PB_CUDA(cudaEventRecord(f->ev_tile_desc, ctx().stream));
    int rc = pb::launch_splat_tiles(f, ntiles, (const pb::SplatTile *)f->d_tile_desc, max_w, max_h, spp, d_xy, d_rgbw,
                                    f->d_scratch_tile, mode);
    if (rc < 0) {
        // radius outside the window kernel's range: one tile at a time, which is the same thing by definition
        for (int i = 0; i < ntiles; ++i) {
            if (desc[i].tb.x1 <= desc[i].tb.x0) continue;
            if (int r2 = pb::launch_splat_tile(f, desc[i].sb, desc[i].tb, spp, d_xy + desc[i].sample_offset,
                                               d_rgbw + desc[i].sample_offset, mode))
                return r2;
        }
        return PBRT_OK;
    }
    if (rc != PBRT_OK) return rc;
    return pbrt_film_merge_tiles(f, ntiles, tbs.data(), poffs.data(), (const float *)f->d_scratch_tile, total_px, 1);
}

// ===================================================================== kernels: textures, LUT, synthetic inputs

// ConstantTexture<T>::evaluate x n (constant.rs:139-141): 4 B (Float) or 12 B (Spectrum) per lookup, write-only.
// No loop: one CTA of 128 threads per 8 KB, four float4 stores per thread, 512 contiguous bytes per warp-store, and the
// hardware block scheduler does the striding — the shape torch's vectorised fill has, and the one that reaches its
// bandwidth (a persistent grid-stride version of the same stores stayed 4-6 % below it; tools/fill_bench.py).
// PERIOD3: the float stream of an rgb fill has period 3, so the float4 at vector index k holds value[(head + 4k + j) % 3]
// = value[(head + k + j) % 3].  Block 0 also writes the unaligned head (to 16-byte alignment) and the tail.
template <bool PERIOD3>
__global__ void __launch_bounds__(128) fill_kernel(float *__restrict__ out, unsigned long long nf, float c0, float c1, float c2) {
    const unsigned long long head = min(nf, (unsigned long long)((16 - ((uintptr_t)out & 15)) & 15) / 4);
    const unsigned long long nvec = (nf - head) / 4;
    float4 *body = reinterpret_cast<float4 *>(out + head);
    const float4 v0 = make_float4(c0, c1, c2, c0), v1 = make_float4(c1, c2, c0, c1), v2 = make_float4(c2, c0, c1, c2);
    const unsigned long long base = (unsigned long long)blockIdx.x * 512 + threadIdx.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const unsigned long long k = base + j * 128;
        if (k < nvec) {
            float4 v = v0;
            if (PERIOD3) {
                const unsigned ph = (unsigned)((head + k) % 3);
                v = ph == 0 ? v0 : (ph == 1 ? v1 : v2);
            }
            body[k] = v;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < 8) {
        const float c[3] = {c0, c1, c2};
        const unsigned long long i = threadIdx.x, done = head + nvec * 4;
        if (i < head) out[i] = c[PERIOD3 ? i % 3 : 0];
        if (done + i < nf) out[done + i] = c[PERIOD3 ? (done + i) % 3 : 0];
    }
}

static unsigned fill_grid(unsigned long long nf) { return (unsigned)std::max<unsigned long long>(1, (nf / 4 + 511) / 512); }

extern "C" int pbrt_texture_constant_eval_f32(float value, uint64_t n, float *out, int dst_is_device) {
    PB_API_LOCK;
    if (int rc = pb::ensure_ready()) return rc;
    if (n == 0) return PBRT_OK;
    if (!out) return fail(PBRT_E_INVALID, "null output");
    void *d_out = out;
    if (!dst_is_device)
        if (int rc = pb::out_stage(n * sizeof(float), &d_out)) return rc;
    fill_kernel<false><<<fill_grid(n), 128, 0, ctx().stream>>>((float *)d_out, n, value, value, value);
    PB_LAUNCH_CHECK("fill_kernel<f32>");
    if (!dst_is_device) {
        PB_CUDA(cudaMemcpyAsync(out, d_out, n * sizeof(float), cudaMemcpyDeviceToHost, ctx().stream));
        PB_CUDA(cudaStreamSynchronize(ctx().stream));
    }
    return PBRT_OK;
}

extern "C" int pbrt_texture_constant_eval_rgb(const float value[3], uint64_t n, float *out, int dst_is_device) {
    PB_API_LOCK;
    if (int rc = pb::ensure_ready()) return rc;
    if (n == 0) return PBRT_OK;
    if (!out || !value) return fail(PBRT_E_INVALID, "null argument");
    void *d_out = out;
    if (!dst_is_device)
        if (int rc = pb::out_stage(n * 3 * sizeof(float), &d_out)) return rc;
    fill_kernel<true><<<fill_grid(n * 3), 128, 0, ctx().stream>>>((float *)d_out, n * 3, value[0], value[1], value[2]);
    PB_LAUNCH_CHECK("fill_kernel<rgb>");
    if (!dst_is_device) {
        PB_CUDA(cudaMemcpyAsync(out, d_out, n * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx().stream));
        PB_CUDA(cudaStreamSynchronize(ctx().stream));
    }
    return PBRT_OK;
}

// mipmap.rs:43-52.  expf on the device is not glibc's expf: agreement is to 1 ulp, not bit-exact.
__global__ void weight_lut_kernel(float *out) {
    int i = threadIdx.x;
    const float alpha = 2.f;
    float r2 = (float)i / (float)(128 - 1);
    out[i] = expf(-alpha * r2) - expf(-alpha);
}

extern "C" int pbrt_mipmap_weight_lut(float out[128]) {
    PB_API_LOCK;
    if (int rc = pb::ensure_ready()) return rc;
    if (!out) return fail(PBRT_E_INVALID, "null output");
    void *d;
    if (int rc = pb::out_stage(128 * sizeof(float), &d)) return rc;
    weight_lut_kernel<<<1, 128, 0, ctx().stream>>>((float *)d);
    PB_LAUNCH_CHECK("weight_lut_kernel");
    PB_CUDA(cudaMemcpyAsync(out, d, 128 * sizeof(float), cudaMemcpyDeviceToHost, ctx().stream));
    PB_CUDA(cudaStreamSynchronize(ctx().stream));
    return PBRT_OK;
}

// one thread per pixel, the pixel's PCG32 stream drawn in order
__global__ void __launch_bounds__(128) synth_samples_kernel(Bounds b, Bounds ib, int spp, int n, unsigned long long seed,
                                                            float2 *__restrict__ xy, float4 *__restrict__ rgbw) {
    const int W = b.x1 - b.x0;
    const long long npx = (long long)W * (b.y1 - b.y0);
    long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npx) return;
    const int px = b.x0 + (int)(p % W), py = b.y0 + (int)(p / W);
    const unsigned long long idx = (unsigned long long)(py - ib.y0) * (unsigned long long)(ib.x1 - ib.x0) + (unsigned long long)(px - ib.x0);
    pb::Pcg32 rng;
    rng.set_sequence((seed << 32) + idx);
    const float fn = (float)n;
    for (int s = 0; s < spp; ++s) {
        const int sx = s % n, sy = s / n;
        float jx = rng.next_float(), jy = rng.next_float();
        float r = rng.next_float(), g = rng.next_float(), bl = rng.next_float();
        const size_t k = (size_t)p * spp + s;
        xy[k] = make_float2((float)px + ((float)sx + jx) / fn, (float)py + ((float)sy + jy) / fn);
        rgbw[k] = make_float4(r, g, bl, 1.f);
    }
}

extern "C" int pbrt_synth_samples(const int32_t bv[4], const int32_t ibv[4], int32_t spp, uint64_t seed, float *xy_dev,
                                  float *rgbw_dev) {
    PB_API_LOCK;
    if (int rc = pb::ensure_ready()) return rc;
    if (!bv || !xy_dev || !rgbw_dev || spp < 1) return fail(PBRT_E_INVALID, "bad argument");
    Bounds b{bv[0], bv[1], bv[2], bv[3]};
    Bounds ib = ibv ? Bounds{ibv[0], ibv[1], ibv[2], ibv[3]} : b;
    if (b.x1 <= b.x0 || b.y1 <= b.y0) return PBRT_OK;
    int n = 1;
    while (n * n < spp) ++n;
    long long npx = (long long)pb::bw(b) * pb::bh(b);
    synth_samples_kernel<<<(unsigned)((npx + 127) / 128), 128, 0, ctx().stream>>>(b, ib, spp, n, seed, (float2 *)xy_dev,
                                                                                   (float4 *)rgbw_dev);
    PB_LAUNCH_CHECK("synth_samples_kernel");
    return PBRT_OK;
}

// App. C tile fill: pixel p of tile t draws rgb from Rng(seed<<32 + t<<20 + p), weight 1
__global__ void __launch_bounds__(256) synth_tiles_kernel(int ntiles, const long long *__restrict__ offsets,
                                                          const long long *__restrict__ counts, unsigned long long seed,
                                                          float4 *__restrict__ rgbw) {
    const int t = blockIdx.y;
    if (t >= ntiles) return;
    const long long cnt = counts[t];
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < cnt; p += (long long)gridDim.x * blockDim.x) {
        pb::Pcg32 rng;
        rng.set_sequence((seed << 32) + ((unsigned long long)t << 20) + (unsigned long long)p);
        float r = rng.next_float(), g = rng.next_float(), b = rng.next_float();
        rgbw[offsets[t] + p] = make_float4(r, g, b, 1.f);
    }
}

extern "C" int pbrt_synth_tiles(int32_t ntiles, const int64_t *offsets, const int64_t *counts, uint64_t seed,
                                float *rgbw_dev, int64_t total_pixels) {
    PB_API_LOCK;
    if (int rc = pb::ensure_ready()) return rc;
    if (ntiles <= 0) return PBRT_OK;
    if (!offsets || !counts || !rgbw_dev) return fail(PBRT_E_INVALID, "null argument");
    int64_t maxc = 0;
    for (int i = 0; i < ntiles; ++i) {
        if (offsets[i] < 0 || counts[i] < 0 || offsets[i] + counts[i] > total_pixels)
            return fail(PBRT_E_INVALID, "tile %d outside the buffer", i);
        maxc = std::max(maxc, counts[i]);
    }
    if (maxc == 0) return PBRT_OK;
    long long *d = nullptr;
    PB_CUDA(cudaMalloc(&d, (size_t)ntiles * 2 * sizeof(long long)));
    cudaStream_t s = ctx().stream;
    cudaMemcpyAsync(d, offsets, (size_t)ntiles * sizeof(long long), cudaMemcpyHostToDevice, s);
    cudaMemcpyAsync(d + ntiles, counts, (size_t)ntiles * sizeof(long long), cudaMemcpyHostToDevice, s);
    int rc = PBRT_OK;
    if (ntiles > 65535) {
        rc = fail(PBRT_E_UNSUPPORTED, "more than 65535 tiles per call");
    } else {
        dim3 grid((unsigned)std::min<int64_t>((maxc + 255) / 256, 64), ntiles);
        synth_tiles_kernel<<<grid, 256, 0, s>>>(ntiles, d, d + ntiles, seed, (float4 *)rgbw_dev);
        ctx().launches++;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) rc = pb::cuda_fail(e, "synth_tiles_kernel");
    }
    cudaStreamSynchronize(s);
    cudaFree(d);
    return rc;
}
