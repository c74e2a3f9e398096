// shard_nccl.cu -- native runtime of the Gaussian-sharded render: the whole frame (kernels AND collectives) is enqueued by three
// C calls per forward and one per backward, so the host cost of a frame no longer scales with the ~25 launches and 5
// collectives it contains (the torch.distributed + ctypes orchestration of sharded.py cost ~1.1 ms of host time per frame; the
// device work of a 2 M-Gaussian 1280x960 frame on 8 B200s is ~0.5 ms).
//
// NCCL is not linked: the library that torch already loaded (libnccl.so.2) is opened at run time and eight entry points are
// resolved by name, so libg4r.so keeps loading (and the single-GPU path keeps working) on a machine without NCCL.  The
// communicator is this library's own (ncclCommInitRank with an id created on rank 0 and distributed by the caller).
//
// Collectives of one frame (all on the caller's stream, in this order):
//   forward_a : all-reduce(max) of the largest (source, destination) pair count        4 B      overflow check of the slabs
//               all-to-all of the slabs = grouped ncclSend / ncclRecv                   world x (cap+1) x 48 B per rank
//   forward_b : grouped send/recv: every plane of the strip into its place in every peer's image, the n_touched segments
//               back to their owners                                                     (5 x strip + cap+1) x 4 B per peer
//   backward  : all-to-all of the accumulator rows (reverse direction)                  world x (cap+1) x 48 B per rank
//               all-reduce(sum) of the pose gradient                                    32 B
#include "g4r_common.cuh"
#include <dlfcn.h>
#include <nccl.h>
#include <mutex>
#include <new>

namespace {
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;
std::mutex g_nccl_mu;

template <typename F> bool sym(void* h, const char* name, F& out) {
    out = reinterpret_cast<F>(dlsym(h, name));
    return out != nullptr;
}
}  // namespace

#define G4R_NCCL_OK(expr)                                                                             \
    do {                                                                                              \
        ncclResult_t r__ = (expr);                                                                    \
        if (r__ != ncclSuccess)                                                                       \
            return g4r_set_error(G4R_ECUDA, "%s failed: %s", #expr, g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "?"); \
    } while (0)

struct G4RShardComm {
    ncclComm_t comm;
    int rank, world;
};

extern "C" {

int g4r_shard_buffers_size(void) { return (int)sizeof(G4RShardBuffers); }

int g4r_shard_nccl_load(const char* path) {
    std::lock_guard<std::mutex> lock(g_nccl_mu);
    if (g_nccl.handle) return G4R_OK;
    void* h = dlopen(path && *path ? path : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return g4r_set_error(G4R_EINVAL, "cannot open NCCL (%s): %s", path ? path : "libnccl.so.2", dlerror());
    NcclApi a;
    a.handle = h;
    const bool ok = sym(h, "ncclGetUniqueId", a.GetUniqueId) && sym(h, "ncclCommInitRank", a.CommInitRank) && sym(h, "ncclCommDestroy", a.CommDestroy) &&
                    sym(h, "ncclGroupStart", a.GroupStart) && sym(h, "ncclGroupEnd", a.GroupEnd) && sym(h, "ncclSend", a.Send) &&
                    sym(h, "ncclRecv", a.Recv) && sym(h, "ncclAllGather", a.AllGather) && sym(h, "ncclAllReduce", a.AllReduce) &&
                    sym(h, "ncclGetErrorString", a.GetErrorString);
    if (!ok) return g4r_set_error(G4R_EINVAL, "NCCL library lacks a required entry point");
    g_nccl = a;
    return G4R_OK;
}

int g4r_shard_nccl_unique_id(uint8_t* out128) {
    if (!g_nccl.handle) return g4r_set_error(G4R_EINVAL, "g4r_shard_nccl_load has not been called");
    if (!out128) return g4r_set_error(G4R_EINVAL, "out is NULL");
    ncclUniqueId id;
    G4R_NCCL_OK(g_nccl.GetUniqueId(&id));
    memcpy(out128, id.internal, NCCL_UNIQUE_ID_BYTES);
    return G4R_OK;
}

int g4r_shard_comm_create(const uint8_t* id128, int32_t rank, int32_t world, G4RShardComm** out) {
    if (!g_nccl.handle) return g4r_set_error(G4R_EINVAL, "g4r_shard_nccl_load has not been called");
    if (!id128 || !out || world < 1 || rank < 0 || rank >= world) return g4r_set_error(G4R_EINVAL, "bad arguments");
    G4RShardComm* c = new (std::nothrow) G4RShardComm();
    if (!c) return g4r_set_error(G4R_EINVAL, "out of host memory");
    ncclUniqueId id;
    memcpy(id.internal, id128, NCCL_UNIQUE_ID_BYTES);
    c->rank = rank; c->world = world;
    const ncclResult_t r = g_nccl.CommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess) { delete c; return g4r_set_error(G4R_ECUDA, "ncclCommInitRank failed: %s", g_nccl.GetErrorString(r)); }
    *out = c;
    return G4R_OK;
}

void g4r_shard_comm_destroy(G4RShardComm* c) {
    if (!c) return;
    if (g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    delete c;
}

static int all_to_all_rows(G4RShardComm* c, const void* send, void* recv, size_t floats_per_peer, cudaStream_t s) {
    G4R_NCCL_OK(g_nccl.GroupStart());
    for (int r = 0; r < c->world; ++r) {
        G4R_NCCL_OK(g_nccl.Send((const float*)send + (size_t)r * floats_per_peer, floats_per_peer, ncclFloat, r, c->comm, s));
        G4R_NCCL_OK(g_nccl.Recv((float*)recv + (size_t)r * floats_per_peer, floats_per_peer, ncclFloat, r, c->comm, s));
    }
    G4R_NCCL_OK(g_nccl.GroupEnd());
    return G4R_OK;
}

// ---- forward, part a: project -> pack -> max pair count -> all-to-all -> unpack -> count / scan -> scatter / sort / composite ----
// Leaves two early read-backs pending on the context: the largest pair count (overflow check of the slab capacity) and N
// (overflow check of the instance capacity); g4r_shard_forward_wait returns both.
int g4r_shard_forward_a(G4RShardComm* c, G4RContext* ctx, const G4RFrame* full, const G4RFrame* strip, const G4RGaussians* g,
                        const G4RShardBuffers* b, void* stream) {
    if (!c || !ctx || !full || !strip || !g || !b) return g4r_set_error(G4R_EINVAL, "NULL argument");
    cudaStream_t s = (cudaStream_t)stream;
    int rc;
    const int world = c->world;
    const int64_t rows = b->cap + 1;
    if (g->P > 0 && (rc = g4r_project_only(full, g, b->geom_local, b->radii_local, b->n_touched_local, stream)) != G4R_OK) return rc;
    int32_t* counts = (int32_t*)(b->payload + b->counts_offset);
    if ((rc = g4r_shard_pack(full, g->P, b->radii_local, b->geom_local, world, b->cap, b->send_slab, counts, b->slots, b->pack_scratch, stream)) != G4R_OK)
        return rc;
    if ((rc = g4r_shard_max_count(counts, world, b->worst, stream)) != G4R_OK) return rc;
    G4R_NCCL_OK(g_nccl.AllReduce(b->worst, b->worst, 1, ncclInt32, ncclMax, c->comm, s));
    if ((rc = g4r_shard_fetch_counts(ctx, b->worst, 4, 0, 1, stream)) != G4R_OK) return rc;
    if ((rc = all_to_all_rows(c, b->send_slab, b->recv_slab, (size_t)rows * 12, s)) != G4R_OK) return rc;
    int32_t* nt_all = (int32_t*)(b->payload + b->strip_elems);
    if ((rc = g4r_shard_unpack(world, b->cap, b->recv_slab, b->radii_all, nt_all, stream)) != G4R_OK) return rc;
    const int32_t P_all = (int32_t)(world * rows);
    if ((rc = g4r_count_tiles(ctx, strip, P_all, b->radii_all, b->recv_slab, b->img_state, stream)) != G4R_OK) return rc;
    return g4r_shard_render_owned(ctx, strip, b, stream);
}

// Phase 2 of the owned strip (scatter / sort / composite into the payload's strip region); also the local re-run after an
// instance-capacity overflow (new binning / sort_scratch / cap_n in `b`).
int g4r_shard_render_owned(G4RContext* ctx, const G4RFrame* strip, const G4RShardBuffers* b, void* stream) {
    if (!ctx || !strip || !b) return g4r_set_error(G4R_EINVAL, "NULL argument");
    const int64_t rows = b->cap + 1;
    const int32_t P_all = (int32_t)(b->world * rows);
    G4RGaussians ga;
    memset(&ga, 0, sizeof(ga));
    ga.P = P_all;
    float* base = b->payload - (int64_t)16 * strip->tile_row_begin * strip->width;      // image row y = strip row y - 16 * begin
    const int64_t plane = b->maxh * strip->width;
    G4RForwardOut out;
    out.color = base; out.depth = base + 3 * plane; out.opacity = base + 4 * plane;
    out.radii = b->radii_all; out.n_touched = (int32_t*)(b->payload + b->strip_elems);
    out.color_plane_stride = plane;
    return g4r_forward_render(ctx, strip, &ga, b->recv_slab, b->img_state, b->binning, b->sort_scratch, b->cap_n, &out, stream);
}

int g4r_shard_forward_wait(G4RContext* ctx, int64_t* N, int64_t* worst) {
    if (!ctx || !N || !worst) return g4r_set_error(G4R_EINVAL, "NULL argument");
    int32_t w = 0;
    int rc;
    if ((rc = g4r_shard_wait_counts(ctx, 1, &w)) != G4R_OK) return rc;
    *worst = w;
    const int64_t n = g4r_wait_num_rendered(ctx);
    if (n < 0) return (int)n;
    *N = n;
    return G4R_OK;
}

// ---- forward, part b: every rank's strip straight into every rank's image, n_touched back to the owners ----------------
// The strips have unequal heights (tiles_y is rarely a multiple of world), so instead of an all-gather of padded strips plus an
// assembly pass, each plane of the strip is SENT to its final place in every peer's [5][H][W] image (one grouped
// ncclSend / ncclRecv, 6 messages per peer: 5 planes + the n_touched segment); the own strip is copied locally.
int g4r_shard_forward_b(G4RShardComm* c, const G4RFrame* full, int32_t P, const G4RShardBuffers* b, void* stream) {
    if (!c || !full || !b) return g4r_set_error(G4R_EINVAL, "NULL argument");
    cudaStream_t s = (cudaStream_t)stream;
    int rc;
    const int W = full->width, H = full->height;
    const int gy = (H + G4R_TILE - 1) / G4R_TILE;
    const size_t X = (size_t)W * H;
    const int64_t rows = b->cap + 1;
    auto y0_of = [&](int r) { return (int)(((int64_t)r * gy) / c->world) * G4R_TILE; };
    auto y1_of = [&](int r) { const int y = (int)(((int64_t)(r + 1) * gy) / c->world) * G4R_TILE; return y < H ? y : H; };
    const int my0 = y0_of(c->rank), my1 = y1_of(c->rank);
    const size_t my_n = (size_t)(my1 - my0) * W;
    int32_t* nt_gather = (int32_t*)b->gathered;                           // [world][rows]: n_touched segments that came back
    const int32_t* nt_all = (const int32_t*)(b->payload + b->strip_elems);
    G4R_NCCL_OK(g_nccl.GroupStart());
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) continue;
        const int ry0 = y0_of(r), ry1 = y1_of(r);
        const size_t r_n = (size_t)(ry1 - ry0) * W;
        for (int pl = 0; pl < 5; ++pl) {
            if (my_n) G4R_NCCL_OK(g_nccl.Send(b->payload + (size_t)pl * b->maxh * W, my_n, ncclFloat, r, c->comm, s));
            if (r_n) G4R_NCCL_OK(g_nccl.Recv(b->images + (size_t)pl * X + (size_t)ry0 * W, r_n, ncclFloat, r, c->comm, s));
        }
        G4R_NCCL_OK(g_nccl.Send(nt_all + (size_t)r * rows, (size_t)rows, ncclInt32, r, c->comm, s));
        G4R_NCCL_OK(g_nccl.Recv(nt_gather + (size_t)r * rows, (size_t)rows, ncclInt32, r, c->comm, s));
    }
    G4R_NCCL_OK(g_nccl.GroupEnd());
    // own strip and own n_touched segment: local copies
    if (my_n)
        G4R_CUDA_OK(cudaMemcpy2DAsync(b->images + (size_t)my0 * W, X * sizeof(float), b->payload, (size_t)b->maxh * W * sizeof(float),
                                      my_n * sizeof(float), 5, cudaMemcpyDeviceToDevice, s));
    G4R_CUDA_OK(cudaMemcpyAsync(nt_gather + (size_t)c->rank * rows, nt_all + (size_t)c->rank * rows, (size_t)rows * sizeof(int32_t),
                                cudaMemcpyDeviceToDevice, s));
    if (P > 0 && (rc = g4r_shard_gather(P, c->world, b->cap, b->slots, nullptr, 0, nullptr, nt_gather, rows, b->n_touched_local, stream)) != G4R_OK)
        return rc;
    return G4R_OK;
}

// ---- backward: composite of the owned strip -> reverse all-to-all -> per-Gaussian sums -> per-Gaussian backward -> pose all-reduce --
int g4r_shard_backward(G4RShardComm* c, const G4RFrame* full, const G4RFrame* strip, const G4RGaussians* g, const G4RShardBuffers* b,
                       const float* dL_dcolor, const float* dL_ddepth, void* acc_all, void* acc_back, void* acc_local, const G4RBackwardIO* io,
                       int32_t reduce_pose, void* stream) {
    if (!c || !full || !strip || !g || !b || !io) return g4r_set_error(G4R_EINVAL, "NULL argument");
    cudaStream_t s = (cudaStream_t)stream;
    int rc;
    const int64_t rows = b->cap + 1;
    const int32_t P_all = (int32_t)(c->world * rows);
    if ((rc = g4r_backward_composite(strip, P_all, b->recv_slab, b->img_state, b->binning, dL_dcolor, dL_ddepth, acc_all, stream)) != G4R_OK) return rc;
    if ((rc = all_to_all_rows(c, acc_all, acc_back, (size_t)rows * G4R_ACC_STRIDE, s)) != G4R_OK) return rc;
    if (g->P > 0 && (rc = g4r_shard_gather(g->P, c->world, b->cap, b->slots, acc_back, rows, acc_local, nullptr, 0, nullptr, stream)) != G4R_OK) return rc;
    if ((rc = g4r_backward_gaussians(full, g, b->radii_local, b->geom_local, acc_local, io, stream)) != G4R_OK) return rc;
    if (reduce_pose) G4R_NCCL_OK(g_nccl.AllReduce(io->dL_dtau, io->dL_dtau, 8, ncclFloat, ncclSum, c->comm, s));
    return G4R_OK;
}

}  // extern "C"
