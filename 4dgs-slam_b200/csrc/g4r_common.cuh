// g4r_common.cuh -- shared device helpers + scratch layouts for the sm_100a rasterizer.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/g4r.h"

#define G4R_BLOCK 256              // threads per CTA for every kernel in this library
#define G4R_TILE_PIX (G4R_TILE * G4R_TILE)
#define G4R_HEADER_WORDS 8
// Per-tile instance counters live one per 128-byte line: atomics on neighbouring words of one line serialise in a
// single L2 slice (measured: 1.26 M atomics on 1200 packed counters = 95 us; one counter per line = ~10x faster).
#define G4R_COUNT_STRIDE 32

// SH basis constants (values identical to DGR/cuda_rasterizer/auxiliary.h:22-39 and
// gaussian_splatting/utils/sh_utils.py:24-41 -- they are the real SH normalisation constants).
#define G4R_SH_C0 0.28209479177387814f
#define G4R_SH_C1 0.4886025119029199f
#define G4R_SH_C2_0 1.0925484305920792f
#define G4R_SH_C2_1 (-1.0925484305920792f)
#define G4R_SH_C2_2 0.31539156525252005f
#define G4R_SH_C2_3 (-1.0925484305920792f)
#define G4R_SH_C2_4 0.5462742152960396f
#define G4R_SH_C3_0 (-0.5900435899266435f)
#define G4R_SH_C3_1 2.890611442640554f
#define G4R_SH_C3_2 (-0.4570457994644658f)
#define G4R_SH_C3_3 0.3731763325901154f
#define G4R_SH_C3_4 (-0.4570457994644658f)
#define G4R_SH_C3_5 1.445305721320277f
#define G4R_SH_C3_6 (-0.5900435899266435f)

static __host__ __device__ __forceinline__ size_t g4r_align(size_t x, size_t a = 256) {
    return (x + a - 1) / a * a;
}

// ---- scratch layouts (host+device agree through these) --------------------------------
struct GeomLayout {      // per-Gaussian state, saved for backward
    size_t rec, clamped, total;
    __host__ __device__ explicit GeomLayout(int P) {
        size_t o = 0;
        rec = o;      o = g4r_align(o + (size_t)P * 48);
        clamped = o;  o = g4r_align(o + (size_t)P);
        total = o + 256;
    }
};
struct ImageLayout {     // per-pixel + per-tile state, saved for backward
    size_t header, counts, ranges, final_T, n_contrib, order, total;
    int tiles_x, tiles_y, tiles;
    __host__ __device__ ImageLayout(int W, int H) {
        tiles_x = (W + G4R_TILE - 1) / G4R_TILE;
        tiles_y = (H + G4R_TILE - 1) / G4R_TILE;
        tiles = tiles_x * tiles_y;
        size_t o = 0;
        header = o;    o = g4r_align(o + G4R_HEADER_WORDS * 4);
        counts = o;    o = g4r_align(o + (size_t)tiles * 4 * G4R_COUNT_STRIDE);   // histogram, then scatter cursors
        ranges = o;    o = g4r_align(o + (size_t)tiles * 8);
        final_T = o;   o = g4r_align(o + (size_t)W * H * 4);
        n_contrib = o; o = g4r_align(o + (size_t)W * H * 4);
        order = o;     o = g4r_align(o + (size_t)tiles * 4);                       // tiles, heaviest first (launch order)
        total = o + 256;
    }
};
// Per-instance state SAVED for backward (the reference's binningBuffer): the sorted id list and, in stream mode
// (g4r_stream_mode()), the sorted splat STREAM -- the 48-byte records of every tile's list laid out contiguously in list
// order (id in the spare slot), which is what lets the composite kernels stage a whole span with one TMA bulk copy.
struct BinLayout {
    size_t point_list, stream, total;
    __host__ __device__ BinLayout(int64_t cap, bool with_stream) {
        size_t c = (size_t)(cap < 1 ? 1 : cap);
        // whichever array the BACKWARD reads sits at offset 0, so that g4r_backward needs no capacity argument
        size_t o = 0;
        stream = o;     o = g4r_align(o + (with_stream ? c * 48 : 0));
        point_list = o; o = g4r_align(o + c * 4);
        total = o + 256;
    }
};
bool g4r_stream_mode();      // process-wide: G4R_TUNE_STREAM (api.cu)
struct SortLayout {      // per-instance scratch of the forward only (dies with the call): unsorted (depth bits, id) pairs
    size_t pairs, pairs_alt, total;
    __host__ __device__ explicit SortLayout(int64_t cap) {
        size_t c = (size_t)(cap < 1 ? 1 : cap);
        size_t o = 0;
        pairs = o;      o = g4r_align(o + c * 8);
        pairs_alt = o;  o = g4r_align(o + c * 8);      // ping-pong buffer of the oversized-tile radix sort
        total = o + 256;
    }
};
// backward accumulators: 12 floats per Gaussian
//   [0]=dL/dmean2D.x [1]=dL/dmean2D.y [2]=dL/dconic.x [3]=dL/dconic.y [4]=dL/dconic.w
//   [5]=dL/dopacity  [6]=dL/dcolor.r  [7]=dL/dcolor.g  [8]=dL/dcolor.b [9]=dL/ddepth
#define G4R_ACC_STRIDE 12

// ---- small device helpers ---------------------------------------------------------------
// float -> int32 with the semantics of PTX cvt.rzi.s32.f32 (truncate, saturate, NaN -> 0),
// which is what a C cast compiles to on the GPU (and what the oracle emulates on the CPU).
static __device__ __forceinline__ int f2i_rz(float v) { return __float2int_rz(v); }

struct TileRect { uint32_t x0, y0, x1, y1; };

// Tile rectangle of a splat, bit-exact to getRect (DGR/cuda_rasterizer/auxiliary.h:46-56):
// all arithmetic in float, one rounding per operation, in this order.
static __device__ __forceinline__ TileRect tile_rect(float px, float py, int radius, uint32_t gx, uint32_t gy) {
    const float rf = (float)radius;
    TileRect r;
    r.x0 = min(gx, (uint32_t)max(0, f2i_rz(__fmul_rn(__fsub_rn(px, rf), 0.0625f))));
    r.y0 = min(gy, (uint32_t)max(0, f2i_rz(__fmul_rn(__fsub_rn(py, rf), 0.0625f))));
    r.x1 = min(gx, (uint32_t)max(0, f2i_rz(__fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(px, rf), 16.0f), -1.0f), 0.0625f))));
    r.y1 = min(gy, (uint32_t)max(0, f2i_rz(__fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(py, rf), 16.0f), -1.0f), 0.0625f))));
    return r;
}

// 128-bit read-only loads (LDG.E.128.CONSTANT) for the record gathers.
static __device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }

// Host-side launch check.
int g4r_set_error(int code, const char* fmt, ...);
#define G4R_CUDA_OK(expr)                                                                 \
    do {                                                                                  \
        cudaError_t e__ = (expr);                                                         \
        if (e__ != cudaSuccess)                                                           \
            return g4r_set_error(G4R_ECUDA, "%s failed: %s", #expr, cudaGetErrorString(e__)); \
    } while (0)
#define G4R_LAUNCH_OK(name)                                                               \
    do {                                                                                  \
        cudaError_t e__ = cudaGetLastError();                                             \
        if (e__ != cudaSuccess)                                                           \
            return g4r_set_error(G4R_ECUDA, "launch of %s failed: %s", name, cudaGetErrorString(e__)); \
    } while (0)

// ---- optional per-stage timing (CUDA events on the launching stream; bench.py's roofline numbers) -------
enum G4RStage { ST_PROJECT = 0, ST_TILE_SCAN, ST_SCATTER, ST_TILE_SORT, ST_COMPOSITE_FWD, ST_COMPOSITE_BWD, ST_GAUSSIAN_BWD, ST_COUNT };
void g4r_stage_begin(int stage, cudaStream_t s);
void g4r_stage_end(int stage, cudaStream_t s);

// ---- raw-parameter mode: the activations of GaussianModel (gaussian_model.py:100-128), one fixed operation order that
// the CPU restatement under oracle/ follows: sigmoid = 1/(1+exp(-x)) (torch's CUDA sigmoid), exp = expf, normalize = q / max(|q|, 1e-12)
// (torch.nn.functional.normalize), |q|^2 accumulated r,x,y,z with fused multiply-adds.
#ifdef __CUDACC__
static __device__ __forceinline__ float g4r_sigmoid(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }
static __device__ __forceinline__ float g4r_quat_norm(float qr, float qx, float qy, float qz) {
    const float n2 = __fmaf_rn(qz, qz, __fmaf_rn(qy, qy, __fmaf_rn(qx, qx, __fmul_rn(qr, qr))));
    return fmaxf(__fsqrt_rn(n2), 1e-12f);
}
#endif

// Experiment switches (environment, read once per process): G4R_TUNE_<NAME>=<int>.  Defaults are the measured best.
int g4r_tunable(const char* name, int dflt);

// Tile ownership of the sharded render: interleaved (t % world == rank) or a contiguous strip of tile rows.
struct TileOwner {
    uint32_t rank, world, row_begin, row_end;      // world > 1: modulo; else row_end > row_begin: strip; else everything
    __host__ __device__ __forceinline__ bool all() const { return world <= 1u && row_end <= row_begin; }
    __host__ __device__ __forceinline__ bool owns(uint32_t tile, uint32_t gx) const {
        if (world > 1u) return tile % world == rank;
        if (row_end > row_begin) { const uint32_t ty = tile / gx; return ty >= row_begin && ty < row_end; }
        return true;
    }
};
static inline TileOwner g4r_owner(const G4RFrame& f) {
    TileOwner o;
    o.world = f.tile_world > 0 ? (uint32_t)f.tile_world : 1u;
    o.rank = f.tile_world > 0 ? (uint32_t)f.tile_rank : 0u;
    o.row_begin = f.tile_row_begin > 0 ? (uint32_t)f.tile_row_begin : 0u;
    o.row_end = f.tile_row_end > 0 ? (uint32_t)f.tile_row_end : 0u;
    return o;
}

// ---- kernel launchers (one translation unit each) ----------------------------------------
int launch_project(const G4RFrame& f, const G4RGaussians& g, void* geom, void* img, int32_t* radii, int32_t* n_touched,
                   cudaStream_t s);
int launch_mark_visible(int P, const float* means3D, const float* viewmatrix, uint8_t* present, cudaStream_t s);
int launch_tile_scan(const G4RFrame& f, void* img, cudaStream_t s, uint32_t* mirror = nullptr);
int launch_count_tiles(const G4RFrame& f, int P, const int32_t* radii, const void* geom, void* img, cudaStream_t s);
int launch_tile_rows(const G4RFrame& f, int P, const int32_t* radii, const void* geom, int32_t* rows, cudaStream_t s);
int launch_scatter_sort(const G4RFrame& f, int P, const int32_t* radii, const void* geom, void* img, void* binning,
                        void* sort_scratch, int64_t capacity, bool record_overflow, cudaStream_t s);
int g4r_overflow_read(int reset, unsigned int* out);
// world x world int matrix, row r at src + r * row_stride: asynchronous copy into the context's pinned buffer + an event
int g4r_context_fetch_matrix(G4RContext* ctx, const void* src, size_t row_stride, int world, cudaStream_t s);
int g4r_context_wait_matrix(G4RContext* ctx, int world, int32_t* out);
int launch_composite_forward(const G4RFrame& f, int P, const void* geom, void* img, const void* binning, int64_t capacity,
                             const G4RForwardOut& out, cudaStream_t s);
int launch_composite_backward(const G4RFrame& f, int P, const void* geom, const void* img, const void* binning,
                              const float* dL_dcolor, const float* dL_ddepth, float* acc, cudaStream_t s);
int launch_gaussian_backward(const G4RFrame& f, const G4RGaussians& g, const int32_t* radii, const void* geom,
                             const float* acc, const G4RBackwardIO& io, cudaStream_t s);
