// api.cu -- extern "C" entry points declared in include/g4r.h.
#include "g4r_common.cuh"
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <atomic>
#include <map>
#include <mutex>
#include <new>
#include <string>

static thread_local char g_err[512] = "";

int g4r_set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

// ---- per-stage profiling: process-wide, off by default -----------------------------------------------------------------------------
// Autograd runs the backward on its own thread(s), so stage_begin / stage_end / read can race: the flag is an atomic (the
// fast path when profiling is off) and everything else is serialised by one mutex.
struct StageProfile {
    std::atomic<bool> enabled{false};
    bool created = false;
    cudaEvent_t ev[ST_COUNT][2];
    bool pending[ST_COUNT] = {};
    double ms[ST_COUNT] = {};
    int64_t count[ST_COUNT] = {};
    std::mutex mu;
};
static StageProfile g_prof;
static const char* const k_stage_names[ST_COUNT] = {"project", "tile_scan", "scatter", "tile_sort", "composite_forward",
                                                    "composite_backward", "gaussian_backward"};

static void prof_collect_locked(int only_stage = -1) {
    for (int i = 0; i < ST_COUNT; ++i)
        if (g_prof.pending[i] && (only_stage < 0 || i == only_stage)) {
            // a stage whose end event has not fired yet stays pending (it is folded in by a later read)
            if (cudaEventQuery(g_prof.ev[i][1]) != cudaSuccess) { cudaGetLastError(); continue; }
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, g_prof.ev[i][0], g_prof.ev[i][1]) == cudaSuccess) { g_prof.ms[i] += ms; g_prof.count[i]++; }
            g_prof.pending[i] = false;
        }
}
void g4r_stage_begin(int stage, cudaStream_t s) {
    if (!g_prof.enabled.load(std::memory_order_relaxed)) return;
    std::lock_guard<std::mutex> lock(g_prof.mu);
    if (g_prof.pending[stage]) {            // the previous launch of this stage was not collected: fold it in now
        cudaEventSynchronize(g_prof.ev[stage][1]);
        prof_collect_locked(stage);
    }
    cudaEventRecord(g_prof.ev[stage][0], s);
}
void g4r_stage_end(int stage, cudaStream_t s) {
    if (!g_prof.enabled.load(std::memory_order_relaxed)) return;
    std::lock_guard<std::mutex> lock(g_prof.mu);
    cudaEventRecord(g_prof.ev[stage][1], s);
    g_prof.pending[stage] = true;
}

struct G4RContext {
    int32_t* host_n;        // pinned: [0..3] instance count read-back, [4..4+16*16) the sharded render's count matrix
    uint32_t* dev_n;        // the device's address of host_n[0] (mapped pinned memory; NULL if the mapping is unavailable)
    cudaEvent_t ev;
    cudaEvent_t ev_matrix;
    bool pending;
    bool matrix_pending;
    int renders_since_project;   // > 0: the scatter cursors of the current image state are dirty
};
#define CTX_MATRIX_OFFSET 4
#define CTX_PINNED_INTS (CTX_MATRIX_OFFSET + 16 * 16)

int g4r_context_fetch_matrix(G4RContext* ctx, const void* src, size_t row_stride, int world, cudaStream_t s) {
    G4R_CUDA_OK(cudaMemcpy2DAsync(ctx->host_n + CTX_MATRIX_OFFSET, sizeof(int32_t) * world, src, row_stride, sizeof(int32_t) * world, world,
                                  cudaMemcpyDeviceToHost, s));
    G4R_CUDA_OK(cudaEventRecord(ctx->ev_matrix, s));
    ctx->matrix_pending = true;
    return G4R_OK;
}
int g4r_context_wait_matrix(G4RContext* ctx, int world, int32_t* out) {
    if (!ctx->matrix_pending) return g4r_set_error(G4R_EINVAL, "no count matrix was requested");
    G4R_CUDA_OK(cudaEventSynchronize(ctx->ev_matrix));
    ctx->matrix_pending = false;
    memcpy(out, ctx->host_n + CTX_MATRIX_OFFSET, sizeof(int32_t) * world * world);
    return G4R_OK;
}

// Experiment switches: G4R_TUNE_<NAME> in the environment, read once per name.
int g4r_tunable(const char* name, int dflt) {
    static std::mutex mu;
    static std::map<std::string, int> cache;
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(name);
    if (it != cache.end()) return it->second;
    const std::string key = std::string("G4R_TUNE_") + name;
    const char* v = getenv(key.c_str());
    const int val = (v && *v) ? atoi(v) : dflt;
    cache[name] = val;
    return val;
}

// Sorted splat stream + TMA staging in the composite kernels (csrc/composite.cu): on / off for the whole process, because the
// saved `binning` buffer has to have the same layout in forward and backward.
bool g4r_stream_mode() {
    static const bool on = g4r_tunable("STREAM", 0) != 0;
    return on;
}

static int check_frame(const G4RFrame* f, bool backward) {
    if (!f) return g4r_set_error(G4R_EINVAL, "frame is NULL");
    if (f->width <= 0 || f->height <= 0) return g4r_set_error(G4R_EINVAL, "image size %dx%d is not positive", f->width, f->height);
    if (!f->bg || !f->viewmatrix || !f->projmatrix || !f->campos) return g4r_set_error(G4R_EINVAL, "bg/viewmatrix/projmatrix/campos must be device pointers");
    if (backward && !f->projmatrix_raw) return g4r_set_error(G4R_EINVAL, "projmatrix_raw is required for backward");
    if (f->sh_degree < 0 || f->sh_degree > 3) return g4r_set_error(G4R_EINVAL, "sh_degree %d outside 0..3", f->sh_degree);
    if (f->tile_world > 0 && (f->tile_rank < 0 || f->tile_rank >= f->tile_world))
        return g4r_set_error(G4R_EINVAL, "tile_rank %d outside [0, %d)", f->tile_rank, f->tile_world);
    return G4R_OK;
}

static int check_gaussians(const G4RFrame* f, const G4RGaussians* g) {
    if (!g) return g4r_set_error(G4R_EINVAL, "gaussians is NULL");
    if (g->P < 0) return g4r_set_error(G4R_EINVAL, "P is negative");
    if (g->P == 0) return G4R_OK;
    if (!g->means3D || !g->opacities) return g4r_set_error(G4R_EINVAL, "means3D and opacities are required");
    if ((g->shs == nullptr) == (g->colors_precomp == nullptr))
        return g4r_set_error(G4R_EINVAL, "Please provide excatly one of either SHs or precomputed colors!");
    const bool has_sr = g->scales != nullptr && g->rotations != nullptr;
    if (((g->scales == nullptr) != (g->rotations == nullptr)) || (has_sr == (g->cov3D_precomp != nullptr)))
        return g4r_set_error(G4R_EINVAL, "Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!");
    if (g->shs) {
        const int K = (f->sh_degree + 1) * (f->sh_degree + 1);
        if (f->sh_coeffs < K) return g4r_set_error(G4R_EINVAL, "sh_degree %d needs %d coefficients but sh has %d", f->sh_degree, K, f->sh_coeffs);
    }
    if (g->activation != G4R_ACT_NONE && g->activation != G4R_ACT_RAW) return g4r_set_error(G4R_EINVAL, "unknown activation mode %d", g->activation);
    if (g->activation == G4R_ACT_RAW) {
        if (!g->shs || !has_sr) return g4r_set_error(G4R_EINVAL, "raw-parameter mode needs features_dc, scaling and rotation");
        if (f->sh_coeffs > 1 && !g->shs_rest) return g4r_set_error(G4R_EINVAL, "raw-parameter mode with %d SH coefficients needs features_rest", f->sh_coeffs);
        if (g->scale_dim != 0 && g->scale_dim != 1 && g->scale_dim != 3) return g4r_set_error(G4R_EINVAL, "scale_dim must be 1 or 3");
    }
    return G4R_OK;
}

extern "C" {

const char* g4r_last_error(void) { return g_err; }
int g4r_version(void) { return 5; }
void g4r_struct_sizes(int32_t* out5) {
    out5[0] = (int32_t)sizeof(G4RFrame); out5[1] = (int32_t)sizeof(G4RGaussians); out5[2] = (int32_t)sizeof(G4RForwardOut);
    out5[3] = (int32_t)sizeof(G4RBackwardIO); out5[4] = (int32_t)sizeof(G4RLayout);
}

int g4r_context_create(G4RContext** out) {
    if (!out) return g4r_set_error(G4R_EINVAL, "out is NULL");
    G4RContext* c = new (std::nothrow) G4RContext();
    if (!c) return g4r_set_error(G4R_EINVAL, "out of host memory");
    c->host_n = nullptr; c->ev = nullptr; c->ev_matrix = nullptr; c->pending = false; c->matrix_pending = false; c->renders_since_project = 0;
    cudaError_t e = cudaMallocHost((void**)&c->host_n, sizeof(int32_t) * CTX_PINNED_INTS);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_matrix, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        if (c->host_n) cudaFreeHost(c->host_n);
        delete c;
        return g4r_set_error(G4R_ECUDA, "context creation failed: %s", cudaGetErrorString(e));
    }
    c->host_n[0] = 0;
    c->dev_n = nullptr;
    if (cudaHostGetDevicePointer((void**)&c->dev_n, c->host_n, 0) != cudaSuccess) { c->dev_n = nullptr; (void)cudaGetLastError(); }
    *out = c;
    return G4R_OK;
}

void g4r_context_destroy(G4RContext* c) {
    if (!c) return;
    if (c->ev) cudaEventDestroy(c->ev);
    if (c->ev_matrix) cudaEventDestroy(c->ev_matrix);
    if (c->host_n) cudaFreeHost(c->host_n);
    delete c;
}

size_t g4r_geom_bytes(int32_t P) { return GeomLayout(P < 0 ? 0 : P).total; }
size_t g4r_image_bytes(int32_t W, int32_t H) { return ImageLayout(W < 1 ? 1 : W, H < 1 ? 1 : H).total; }
size_t g4r_binning_bytes(int64_t capacity) { return BinLayout(capacity, g4r_stream_mode()).total; }
size_t g4r_sort_scratch_bytes(int64_t capacity) { return SortLayout(capacity).total; }
size_t g4r_backward_scratch_bytes(int32_t P) { return g4r_align((size_t)(P < 1 ? 1 : P) * G4R_ACC_STRIDE * sizeof(float)) + 256; }

int g4r_layout(int32_t P, int32_t W, int32_t H, int64_t capacity, G4RLayout* out) {
    if (!out) return g4r_set_error(G4R_EINVAL, "out is NULL");
    const GeomLayout gl(P < 0 ? 0 : P);
    const ImageLayout il(W < 1 ? 1 : W, H < 1 ? 1 : H);
    const BinLayout bl(capacity, g4r_stream_mode());
    out->geom_rec = gl.rec; out->geom_clamped = gl.clamped;
    out->img_final_T = il.final_T; out->img_n_contrib = il.n_contrib; out->img_ranges = il.ranges;
    out->img_counts = il.counts; out->img_header = il.header;
    out->bin_point_list = bl.point_list; out->bin_pairs = SortLayout(capacity).pairs;
    return G4R_OK;
}

int g4r_forward_project(G4RContext* ctx, const G4RFrame* f, const G4RGaussians* g, void* geom, void* img, int32_t* radii,
                        int32_t* n_touched, void* stream) {
    int rc;
    // ctx == NULL: no read-back of N (CUDA-graph capture: the caller guarantees the capacity handed to phase 2)
    if ((rc = check_frame(f, false)) != G4R_OK) return rc;
    if ((rc = check_gaussians(f, g)) != G4R_OK) return rc;
    if (!img) return g4r_set_error(G4R_EINVAL, "img buffer is NULL");
    if (g->P > 0 && (!geom || !radii || !n_touched)) return g4r_set_error(G4R_EINVAL, "geom/radii/n_touched buffers are NULL");
    if (((uintptr_t)geom | (uintptr_t)img) & 15u) return g4r_set_error(G4R_EINVAL, "scratch buffers must be 16-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    const ImageLayout il(f->width, f->height);
    char* ib = (char*)img;
    // header and the per-tile counters (histogram, later scatter cursors) are contiguous at the start of the image state
    G4R_CUDA_OK(cudaMemsetAsync(ib + il.header, 0, il.ranges - il.header, s));
    if (g->P > 0) {
        if ((rc = launch_project(*f, *g, geom, img, radii, n_touched, s)) != G4R_OK) return rc;
    }
    // N reaches the host through mapped pinned memory written by the scan kernel itself (no copy in the stream); the event
    // behind the kernel tells the host when to read it
    if ((rc = launch_tile_scan(*f, img, s, ctx ? ctx->dev_n : nullptr)) != G4R_OK) return rc;
    if (ctx) {
        if (!ctx->dev_n) G4R_CUDA_OK(cudaMemcpyAsync(ctx->host_n, ib + il.header, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        G4R_CUDA_OK(cudaEventRecord(ctx->ev, s));
        ctx->pending = true;
        ctx->renders_since_project = 0;
    }
    return G4R_OK;
}

int64_t g4r_wait_num_rendered(G4RContext* ctx) {
    if (!ctx) return g4r_set_error(G4R_EINVAL, "context is NULL");
    if (!ctx->pending) return g4r_set_error(G4R_EINVAL, "g4r_wait_num_rendered without a preceding g4r_forward_project");
    cudaError_t e = cudaEventSynchronize(ctx->ev);
    if (e != cudaSuccess) return g4r_set_error(G4R_ECUDA, "cudaEventSynchronize failed: %s", cudaGetErrorString(e));
    ctx->pending = false;
    return (int64_t)(uint32_t)*(volatile int32_t*)ctx->host_n;
}

int g4r_forward_render(G4RContext* ctx, const G4RFrame* f, const G4RGaussians* g, void* geom, void* img, void* binning,
                       void* sort_scratch, int64_t capacity, const G4RForwardOut* out, void* stream) {
    int rc;
    if ((rc = check_frame(f, false)) != G4R_OK) return rc;
    if (!g || g->P < 0) return g4r_set_error(G4R_EINVAL, "gaussians is NULL or P is negative");   // phase 2 only needs P
    if (!out || !out->color || !out->depth || !out->opacity) return g4r_set_error(G4R_EINVAL, "output images are NULL");
    if (!img || !binning || !sort_scratch) return g4r_set_error(G4R_EINVAL, "img/binning/sort_scratch buffers are NULL");
    if (g->P > 0 && (!geom || !out->radii || !out->n_touched)) return g4r_set_error(G4R_EINVAL, "geom/radii/n_touched are NULL");
    if (capacity < 0) return g4r_set_error(G4R_EINVAL, "capacity is negative");
    if (((uintptr_t)geom | (uintptr_t)img | (uintptr_t)binning | (uintptr_t)sort_scratch) & 15u)
        return g4r_set_error(G4R_EINVAL, "scratch buffers must be 16-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    if (g->P > 0) {
        if (ctx && ctx->renders_since_project > 0) {
            // re-run after a capacity overflow: the aborted attempt touched nothing (every phase-2 kernel
            // exits when N > capacity), but be robust to a caller re-rendering a completed frame too.
            const ImageLayout il(f->width, f->height);
            G4R_CUDA_OK(cudaMemsetAsync((char*)img + il.counts, 0, il.ranges - il.counts, s));
            G4R_CUDA_OK(cudaMemsetAsync(out->n_touched, 0, sizeof(int32_t) * (size_t)g->P, s));
        }
        if (ctx) ctx->renders_since_project++;
        if ((rc = launch_scatter_sort(*f, g->P, out->radii, geom, img, binning, sort_scratch, capacity, ctx == nullptr, s)) != G4R_OK) return rc;
    }
    return launch_composite_forward(*f, g->P, geom, img, binning, capacity, *out, s);
}

int g4r_backward_composite(const G4RFrame* f, int32_t P_all, const void* geom_all, const void* img, const void* binning,
                           const float* dL_dcolor, const float* dL_ddepth, void* scratch, void* stream) {
    int rc;
    if ((rc = check_frame(f, false)) != G4R_OK) return rc;
    if (P_all <= 0) return G4R_OK;
    if (!dL_dcolor || !dL_ddepth) return g4r_set_error(G4R_EINVAL, "dL_dcolor / dL_ddepth are NULL");
    if (!geom_all || !img || !binning || !scratch) return g4r_set_error(G4R_EINVAL, "saved state / scratch is NULL");
    if (((uintptr_t)geom_all | (uintptr_t)img | (uintptr_t)binning | (uintptr_t)scratch) & 15u)
        return g4r_set_error(G4R_EINVAL, "scratch buffers must be 16-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    float* acc = (float*)scratch;
    G4R_CUDA_OK(cudaMemsetAsync(acc, 0, (size_t)P_all * G4R_ACC_STRIDE * sizeof(float), s));
    return launch_composite_backward(*f, P_all, geom_all, img, binning, dL_dcolor, dL_ddepth, acc, s);
}

int g4r_backward_gaussians(const G4RFrame* f, const G4RGaussians* g, const int32_t* radii, const void* geom, const void* acc,
                           const G4RBackwardIO* io, void* stream) {
    int rc;
    if ((rc = check_frame(f, true)) != G4R_OK) return rc;
    if ((rc = check_gaussians(f, g)) != G4R_OK) return rc;
    if (!io || !io->dL_dtau) return g4r_set_error(G4R_EINVAL, "io / dL_dtau is NULL");
    cudaStream_t s = (cudaStream_t)stream;
    G4R_CUDA_OK(cudaMemsetAsync(io->dL_dtau, 0, sizeof(float) * 8, s));
    if (g->P == 0) return G4R_OK;
    // Every per-Gaussian output is optional: a NULL pointer means the caller does not need that gradient (autograd's
    // needs_input_grad; e.g. pose tracking consumes dL_dtau only) and the kernel skips those stores.
    if (!radii || !geom || !acc) return g4r_set_error(G4R_EINVAL, "saved state / accumulators are NULL");
    if (((uintptr_t)geom | (uintptr_t)acc) & 15u) return g4r_set_error(G4R_EINVAL, "scratch buffers must be 16-byte aligned");
    if (io->dL_drotations && ((uintptr_t)io->dL_drotations & 15u)) return g4r_set_error(G4R_EINVAL, "dL_drotations must be 16-byte aligned");
    return launch_gaussian_backward(*f, *g, radii, geom, (const float*)acc, *io, s);
}

int g4r_backward(const G4RFrame* f, const G4RGaussians* g, const int32_t* radii, const void* geom, const void* img,
                 const void* binning, void* scratch, const G4RBackwardIO* io, void* stream) {
    int rc;
    if ((rc = check_frame(f, true)) != G4R_OK) return rc;
    if ((rc = check_gaussians(f, g)) != G4R_OK) return rc;
    if (!io || !io->dL_dtau) return g4r_set_error(G4R_EINVAL, "io / dL_dtau is NULL");
    if (g->P > 0) {
        if ((rc = g4r_backward_composite(f, g->P, geom, img, binning, io->dL_dcolor, io->dL_ddepth, scratch, stream)) != G4R_OK) return rc;
    }
    return g4r_backward_gaussians(f, g, radii, geom, scratch, io, stream);
}

int g4r_tile_rows(const G4RFrame* f, int32_t P, const int32_t* radii, const void* geom, int32_t* rows, void* stream) {
    int rc;
    if ((rc = check_frame(f, false)) != G4R_OK) return rc;
    if (P <= 0) return G4R_OK;
    if (!radii || !geom || !rows) return g4r_set_error(G4R_EINVAL, "radii/geom/rows are NULL");
    if (((uintptr_t)geom & 15u) || ((uintptr_t)rows & 7u)) return g4r_set_error(G4R_EINVAL, "geom must be 16-byte and rows 8-byte aligned");
    return launch_tile_rows(*f, P, radii, geom, rows, (cudaStream_t)stream);
}

int g4r_project_only(const G4RFrame* f, const G4RGaussians* g, void* geom, int32_t* radii, int32_t* n_touched, void* stream) {
    int rc;
    if ((rc = check_frame(f, false)) != G4R_OK) return rc;
    if ((rc = check_gaussians(f, g)) != G4R_OK) return rc;
    if (g->P == 0) return G4R_OK;
    if (!geom || !radii || !n_touched) return g4r_set_error(G4R_EINVAL, "geom/radii/n_touched buffers are NULL");
    if ((uintptr_t)geom & 15u) return g4r_set_error(G4R_EINVAL, "scratch buffers must be 16-byte aligned");
    return launch_project(*f, *g, geom, nullptr, radii, n_touched, (cudaStream_t)stream);
}

int g4r_count_tiles(G4RContext* ctx, const G4RFrame* f, int32_t P_all, const int32_t* radii_all, const void* geom_all, void* img,
                    void* stream) {
    int rc;
    if (!ctx) return g4r_set_error(G4R_EINVAL, "context is NULL");
    if ((rc = check_frame(f, false)) != G4R_OK) return rc;
    if (!img) return g4r_set_error(G4R_EINVAL, "img buffer is NULL");
    if (P_all > 0 && (!radii_all || !geom_all)) return g4r_set_error(G4R_EINVAL, "radii/geom are NULL");
    if (((uintptr_t)geom_all | (uintptr_t)img) & 15u) return g4r_set_error(G4R_EINVAL, "scratch buffers must be 16-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    const ImageLayout il(f->width, f->height);
    char* ib = (char*)img;
    G4R_CUDA_OK(cudaMemsetAsync(ib + il.header, 0, il.ranges - il.header, s));
    if (P_all > 0 && (rc = launch_count_tiles(*f, P_all, radii_all, geom_all, img, s)) != G4R_OK) return rc;
    if ((rc = launch_tile_scan(*f, img, s, ctx->dev_n)) != G4R_OK) return rc;
    if (!ctx->dev_n) G4R_CUDA_OK(cudaMemcpyAsync(ctx->host_n, ib + il.header, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    G4R_CUDA_OK(cudaEventRecord(ctx->ev, s));
    ctx->pending = true;
    ctx->renders_since_project = 0;
    return G4R_OK;
}

int64_t g4r_overflow_status(int reset) {
    unsigned int v = 0;
    const int rc = g4r_overflow_read(reset, &v);
    return rc != G4R_OK ? (int64_t)rc : (int64_t)v;
}

int g4r_profile_enable(int on) {
    std::lock_guard<std::mutex> lock(g_prof.mu);
    if (on && !g_prof.created) {
        for (int i = 0; i < ST_COUNT; ++i)
            for (int j = 0; j < 2; ++j) G4R_CUDA_OK(cudaEventCreate(&g_prof.ev[i][j]));
        g_prof.created = true;
    }
    g_prof.enabled.store(on != 0);
    return G4R_OK;
}
int g4r_profile_stage_count(void) { return ST_COUNT; }
const char* g4r_profile_stage_name(int i) { return (i >= 0 && i < ST_COUNT) ? k_stage_names[i] : ""; }
int g4r_profile_read(double* ms_out, int64_t* count_out, int reset) {
    // caller must have synchronised the stream(s) the stages ran on
    std::lock_guard<std::mutex> lock(g_prof.mu);
    prof_collect_locked();
    for (int i = 0; i < ST_COUNT; ++i) {
        if (ms_out) ms_out[i] = g_prof.ms[i];
        if (count_out) count_out[i] = g_prof.count[i];
        if (reset) { g_prof.ms[i] = 0.0; g_prof.count[i] = 0; }
    }
    return G4R_OK;
}

int g4r_mark_visible(int32_t P, const float* means3D, const float* viewmatrix, const float* projmatrix, uint8_t* present, void* stream) {
    (void)projmatrix;   // the reference's frustum test only uses the view-space depth (auxiliary.h:154)
    if (P < 0) return g4r_set_error(G4R_EINVAL, "P is negative");
    if (P == 0) return G4R_OK;
    if (!means3D || !viewmatrix || !present) return g4r_set_error(G4R_EINVAL, "NULL argument");
    return launch_mark_visible(P, means3D, viewmatrix, present, (cudaStream_t)stream);
}

}  // extern "C"
