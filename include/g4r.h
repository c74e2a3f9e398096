/*
 * g4r.h -- C ABI of the B200-native differentiable Gaussian rasterizer ("g4r").
 *
 * This is the drop-in boundary for the ONE hot path of yanyan-li/4DGS-SLAM: the
 * diff_gaussian_rasterization extension.  Every entry point below replaces a
 * function of the reference's native layer (citations relative to
 * /root/reference/submodules/diff-gaussian-rasterization/, "DGR/"):
 *
 *   g4r_forward_project + g4r_forward_render
 *        == CudaRasterizer::Rasterizer::forward      DGR/cuda_rasterizer/rasterizer_impl.cu:198-344
 *           (called from RasterizeGaussiansCUDA        DGR/rasterize_points.cu:35-122)
 *   g4r_backward
 *        == CudaRasterizer::Rasterizer::backward     DGR/cuda_rasterizer/rasterizer_impl.cu:348-455
 *           (called from RasterizeGaussiansBackwardCUDA DGR/rasterize_points.cu:124-211)
 *           plus the (P,6)->(6) pose-gradient sum of  DGR/diff_gaussian_rasterization/__init__.py:152-154
 *   g4r_mark_visible
 *        == CudaRasterizer::Rasterizer::markVisible  DGR/cuda_rasterizer/rasterizer_impl.cu:141-153
 *   g4r_geom_bytes / g4r_image_bytes / g4r_binning_bytes / g4r_backward_scratch_bytes
 *        == required<GeometryState|ImageState|BinningState>() DGR/cuda_rasterizer/rasterizer_impl.h:64-72
 *
 * Conventions
 *   - plain C: raw DEVICE pointers, sizes, a cudaStream_t passed as void*. No torch
 *     types, no C++ exceptions cross this boundary.
 *   - every function returns 0 on success or a negative G4R_E* code; the text of
 *     the last error on the calling thread is available from g4r_last_error().
 *   - all buffers are owned by the caller (the Python side hands in torch tensors
 *     so autograd keeps them alive until backward, like the reference's
 *     geomBuffer/binningBuffer/imgBuffer).  Optional inputs are NULL, mirroring
 *     the reference's empty-tensor sentinels (DGR/rasterize_points.cu:85-117).
 *   - all work is enqueued on `stream`; nothing here synchronises the device
 *     except g4r_wait_num_rendered(), which waits on one event.
 *   - float32 everywhere; matrices are 16 floats in the reference's layout
 *     (column-major world->view / full projection).
 */
#ifndef G4R_H_INCLUDED
#define G4R_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define G4R_OK            0
#define G4R_EINVAL       -1   /* bad argument (NULL where required, negative size, ...) */
#define G4R_ECUDA        -2   /* a CUDA runtime call or kernel launch failed            */
#define G4R_EOVERFLOW    -3   /* binning capacity too small (informational)              */

#define G4R_ACT_NONE      0   /* inputs are activated parameters (the reference rasterizer's contract) */
#define G4R_ACT_RAW       1   /* inputs are GaussianModel's raw parameters; activations are applied in-kernel */

#define G4R_TILE          16  /* tile edge in pixels: BLOCK_X/BLOCK_Y of DGR/cuda_rasterizer/config.h:16-17 */
#define G4R_CHANNELS      3   /* NUM_CHANNELS of DGR/cuda_rasterizer/config.h:15 */

/* Per-frame camera + raster settings: the numeric fields of
 * GaussianRasterizationSettings (DGR/diff_gaussian_rasterization/__init__.py:173-186).
 * Pointers are DEVICE pointers, read by the kernels (never by the host). */
typedef struct G4RFrame {
    int32_t width;                /* image_width  */
    int32_t height;               /* image_height */
    float   tan_fovx;
    float   tan_fovy;
    float   scale_modifier;
    int32_t sh_degree;            /* active degree D (0..3) */
    int32_t sh_coeffs;            /* M = coefficients allocated per Gaussian (sh.size(1)); 0 when no SH */
    int32_t prefiltered;          /* accepted for API parity; a culled point is simply skipped */
    const float* bg;              /* [3]  */
    const float* viewmatrix;      /* [16] */
    const float* projmatrix;      /* [16] */
    const float* projmatrix_raw;  /* [16] backward only (pose Jacobian); may be NULL in forward */
    const float* campos;          /* [3]  */
    /* Tile ownership for the Gaussian-sharded multi-GPU render: phase 2 and the backward composite only touch tiles t
     * with t % tile_world == tile_rank.  Single GPU: tile_rank = 0, tile_world = 1 (a zeroed tile_world means 1). */
    int32_t tile_rank;
    int32_t tile_world;
    /* Alternative ownership by contiguous tile-row strip [tile_row_begin, tile_row_end) (used when tile_world <= 1 and
     * tile_row_end > tile_row_begin): the all-to-all exchange sends a splat only to the ranks whose strip it touches. */
    int32_t tile_row_begin;
    int32_t tile_row_end;
} G4RFrame;

/* Per-Gaussian inputs (all DEVICE pointers, contiguous, read-only). */
typedef struct G4RGaussians {
    int32_t P;                    /* number of Gaussians */
    const float* means3D;         /* [P,3] */
    const float* opacities;       /* [P]   (post-sigmoid) */
    const float* shs;             /* [P,M,3] or NULL */
    const float* colors_precomp;  /* [P,3]   or NULL  (exactly one of shs/colors_precomp) */
    const float* scales;          /* [P,3] (post-exp)  or NULL */
    const float* rotations;       /* [P,4] (w,x,y,z)   or NULL */
    const float* cov3D_precomp;   /* [P,6] or NULL  (exactly one of scales+rotations / cov3D_precomp) */
    /* ---- raw-parameter mode (opt-in; folds the activation prelude of the caller into the kernels) ----------------
     * The reference's render() (gaussian_splatting/gaussian_renderer/__init__.py:108-131) feeds the rasterizer
     * GaussianModel.get_opacity = sigmoid(_opacity), get_scaling = exp(_scaling) (repeated x3 when isotropic),
     * get_rotation = normalize(_rotation) and get_features = cat(_features_dc, _features_rest)
     * (gaussian_splatting/scene/gaussian_model.py:100-128): ~7 element-wise torch kernels per call plus their
     * autograd twins.  With activation = G4R_ACT_RAW the pointers above are the RAW parameters: opacities = _opacity,
     * scales = _scaling [P,scale_dim], rotations = _rotation, shs = _features_dc [P,1,3], shs_rest = _features_rest
     * [P,M-1,3] (NULL when M == 1); the kernels apply the activations, and the backward returns gradients w.r.t. the
     * raw parameters (chain rule of exp / sigmoid / normalize applied in the per-Gaussian backward kernel). */
    const float* shs_rest;        /* [P,M-1,3] or NULL; only read when activation == G4R_ACT_RAW */
    int32_t activation;           /* G4R_ACT_NONE (reference surface) or G4R_ACT_RAW */
    int32_t scale_dim;            /* raw mode: 3, or 1 for an isotropic _scaling [P,1]; 0 means 3 */
    /* ---- static mask + dynamic offsets (opt-in; the rest of render()'s prelude, gaussian_renderer/__init__.py:159-191) ----
     * The reference gathers every per-Gaussian tensor with a boolean mask before the call (tracking renders only the static
     * Gaussians, utils/slam_frontend.py:412-414) and scatters per-dynamic-Gaussian offsets into zero tensors that it adds to the
     * positions / activated scales / activated rotations (:163-174).  Here the kernels do both in place:
     *   mask[i] == 0        -> Gaussian i is skipped (radius 0, zero gradients); outputs keep the FULL length P
     *   dyn_slot[i] = k >= 0 -> means3D[i] += dx[k], scale[i] += ds[k] (after exp), rotation[i] += dr[k] (after normalize);
     *                          k = rank of i among the dynamic Gaussians (cumsum(dygs) - 1), -1 for static ones. */
    const uint8_t* mask;          /* [P] or NULL */
    const int32_t* dyn_slot;      /* [P] or NULL */
    const float* dx;              /* [Pd,3] or NULL */
    const float* ds;              /* [Pd,3] or NULL */
    const float* dr;              /* [Pd,4] or NULL */
} G4RGaussians;

/* Forward outputs (DEVICE pointers, written in full; no pre-zeroing needed). */
typedef struct G4RForwardOut {
    float*   color;               /* [3,H,W] */
    float*   depth;               /* [1,H,W] */
    float*   opacity;             /* [1,H,W] */
    int32_t* radii;               /* [P]     */
    int32_t* n_touched;           /* [P]     */
    int64_t  color_plane_stride;  /* floats between the three colour planes; 0 = width * height.  The sharded render lets a rank
                                   * composite its strip of tile rows straight into an all-gather send buffer [5][rows][W]: it
                                   * passes pointers offset by -first_row * W and the strip's plane stride. */
} G4RForwardOut;

/* Incoming image gradients + outgoing per-Gaussian gradients (DEVICE pointers).
 * Every non-NULL output is written in full (zeros for invisible Gaussians), so the
 * caller may hand in uninitialised memory.  ANY per-Gaussian output may be NULL: the
 * gradient is then not needed (autograd's needs_input_grad -- pose tracking consumes
 * dL_dtau only, utils/slam_frontend.py:441-448) and its stores are skipped. */
typedef struct G4RBackwardIO {
    const float* dL_dcolor;       /* [3,H,W] */
    const float* dL_ddepth;       /* [1,H,W] */
    float* dL_dmeans3D;           /* [P,3]   */
    float* dL_dmeans2D;           /* [P,3]   (x,y screen-space NDC-scaled; z = 0) */
    float* dL_dopacity;           /* [P]     */
    float* dL_dshs;               /* [P,M,3] or NULL */
    float* dL_dcolors_precomp;    /* [P,3]   or NULL */
    float* dL_dscales;            /* [P,3]   or NULL */
    float* dL_drotations;         /* [P,4]   or NULL */
    float* dL_dcov3D;             /* [P,6]   or NULL (only when cov3D_precomp was given) */
    float* dL_dtau;               /* [8]: [0:3] = grad_rho, [3:6] = grad_theta, [6:8] padding */
    float* dL_dshs_rest;          /* [P,M-1,3] or NULL; raw mode only (then dL_dshs is [P,1,3], dL_dscales [P,scale_dim]) */
    float* dL_ddx;                /* [Pd,3] or NULL: gradients of the dynamic offsets (rows of Gaussians with dyn_slot >= 0) */
    float* dL_dds;                /* [Pd,3] or NULL */
    float* dL_ddr;                /* [Pd,4] or NULL */
} G4RBackwardIO;

typedef struct G4RContext G4RContext;   /* owns one pinned int + one event; one per host thread/device */

/* ---- library / context ------------------------------------------------------------ */
const char* g4r_last_error(void);
int  g4r_version(void);                               /* ABI version, currently 5 */
void g4r_struct_sizes(int32_t* out5);                 /* sizeof {G4RFrame, G4RGaussians, G4RForwardOut, G4RBackwardIO, G4RLayout}: FFI self-check */
int  g4r_context_create(G4RContext** out);
void g4r_context_destroy(G4RContext* ctx);

/* ---- scratch sizing (bytes) --------------------------------------------------------- */
size_t g4r_geom_bytes(int32_t P);                     /* per-Gaussian splat records, saved for backward */
size_t g4r_image_bytes(int32_t width, int32_t height);/* per-pixel final_T/n_contrib + per-tile ranges  */
size_t g4r_binning_bytes(int64_t capacity);           /* sorted id list for `capacity` (tile,Gaussian) instances, saved for backward */
size_t g4r_sort_scratch_bytes(int64_t capacity);      /* unsorted (depth,id) pairs: forward-only scratch, free after the call */
size_t g4r_backward_scratch_bytes(int32_t P);         /* per-Gaussian gradient accumulators (not saved) */

/* ---- forward, phase 1: projection + tile histogram + tile offsets --------------------
 * Enqueues: per-Gaussian projection (writes radii, zeroes n_touched, fills geom), per-tile instance
 * counts and their exclusive scan (writes ranges into img), then an async copy of the
 * instance total N into the context's pinned int and an event record.
 * Never blocks.  ctx may be NULL: nothing is read back (CUDA-graph capture; the caller guarantees phase 2's capacity and
 * can read N later from the image state, G4RLayout.img_header). */
int g4r_forward_project(G4RContext* ctx, const G4RFrame* frame, const G4RGaussians* g,
                        void* geom, void* img, int32_t* radii, int32_t* n_touched, void* stream);

/* ---- forward, phase 2: instance scatter + per-tile depth sort + composite -------------
 * `capacity` = number of instances `binning` (the sorted id list, saved for backward) and `sort_scratch` (the unsorted
 * pairs, needed only until this call's kernels have run) have room for.  If the device-side N
 * turns out larger, every phase-2 kernel exits without touching memory and the
 * caller must call again with a larger buffer (see g4r_wait_num_rendered).
 * Never blocks. */
int g4r_forward_render(G4RContext* ctx, const G4RFrame* frame, const G4RGaussians* g,
                       void* geom, void* img, void* binning, void* sort_scratch, int64_t capacity,
                       const G4RForwardOut* out, void* stream);

/* Waits for the event recorded by g4r_forward_project and returns N (>= 0), the
 * reference's `num_rendered` (DGR/cuda_rasterizer/rasterizer_impl.cu:283-284), or a
 * negative error code.  By the time phase 2 has been enqueued the event has normally
 * already fired, so the GPU is never idle waiting for the host. */
int64_t g4r_wait_num_rendered(G4RContext* ctx);

/* Forwards that run without host read-back (ctx == NULL: captured in a CUDA graph and replayed) cannot re-run phase 2 when
 * N outgrows the capacity fixed at capture: every kernel exits untouched, the backward returns zero gradients, and the
 * kernel that detects it records N in a per-device word.  Returns the largest such N since the last reset on the CURRENT
 * device (0 = no overflow) or a negative error code; synchronises the device. */
int64_t g4r_overflow_status(int reset);

/* ---- backward ------------------------------------------------------------------------ */
int g4r_backward(const G4RFrame* frame, const G4RGaussians* g,
                 const int32_t* radii, const void* geom, const void* img, const void* binning,
                 void* scratch, const G4RBackwardIO* io, void* stream);

/* ---- building blocks of the Gaussian-sharded multi-GPU render ------------
 * One process per GPU.  Rank r owns a contiguous shard of the Gaussians and the strip of tile rows
 * [r*tiles_y/world, (r+1)*tiles_y/world).  Per frame and rank (all enqueued on `stream`, no host synchronisation):
 *   g4r_project_only      project the local shard (48-byte splat records + radii)
 *   g4r_shard_pack        per destination rank d, stably compact the records whose tile rectangle touches d's strip into
 *                         send_slab[d][0..counts[d]) (fixed capacity `cap` per pair; the radius rides in the record's spare
 *                         slot; the count in the slab's header row) and write slots[d][i] = position of Gaussian i in slab d, or -1
 *   (caller)              all-to-all of the slabs (NCCL, equal splits)
 *   g4r_shard_unpack      radii of all world*(cap+1) received slots (0 for unused ones)
 *   g4r_count_tiles + g4r_forward_render   bin / sort / composite the OWNED strip over the received records (frame->tile_row_*),
 *                         writing into a strip buffer through G4RForwardOut.color_plane_stride
 *   (caller)              all-gather of the strips; g4r_shard_assemble builds the [planes,H,W] images
 *   g4r_backward_composite   accumulator rows of the received records; (caller) reverse all-to-all
 *   g4r_shard_gather      acc_local[i] = sum of the rows that came back for Gaussian i (same for n_touched)
 *   g4r_backward_gaussians   per-Gaussian backward of the local shard; (caller) all-reduce of dL_dtau
 * g4r_backward == g4r_backward_composite + g4r_backward_gaussians on one GPU.  counts[d] > cap means records were dropped:
 * the caller (which sees the count matrix on the host after everything is enqueued) redoes the frame with a larger cap. */
int g4r_project_only(const G4RFrame* frame, const G4RGaussians* g, void* geom, int32_t* radii, int32_t* n_touched, void* stream);
size_t g4r_shard_scratch_bytes(int32_t P, int32_t world);
int g4r_shard_pack(const G4RFrame* frame, int32_t P, const int32_t* radii, const void* geom, int32_t world, int64_t cap,
                   void* send_slab, int32_t* counts, int32_t* slots, void* scratch, void* stream);
/* Slabs are [world][cap + 1] records: row `cap` of slab d is its header {count, 0, ...}, so the counts travel inside the
 * all-to-all.  Received slots are numbered j = s * (cap + 1) + k; g4r_shard_unpack writes radii_all[world * (cap + 1)]. */
int g4r_shard_unpack(int32_t world, int64_t cap, const void* recv_slab, int32_t* radii_all, int32_t* n_touched_all /* zeroed; may be NULL */,
                     void* stream);
/* The world x world count matrix (row r = what rank r sent to everybody) sits at gathered + offset_bytes + r * rank_stride_bytes
 * after the all-gather: g4r_shard_fetch_counts enqueues its copy into the context's pinned buffer plus an event,
 * g4r_shard_wait_counts waits for THAT event only and copies the matrix out. */
int g4r_shard_fetch_counts(G4RContext* ctx, const void* gathered, int64_t rank_stride_bytes, int64_t offset_bytes, int32_t world, void* stream);
int g4r_shard_wait_counts(G4RContext* ctx, int32_t world, int32_t* out);
/* acc_back / n_touched_back hold, for destination d, the rows that came back for the records this rank sent to d, starting at
 * row (element) d * stride: stride = cap + 1 rows after the reverse all-to-all, or the payload size when the data arrives
 * inside the strips' all-gather. */
int g4r_shard_gather(int32_t P, int32_t world, int64_t cap, const int32_t* slots, const void* acc_back, int64_t acc_stride_rows,
                     void* acc_local, const int32_t* n_touched_back, int64_t n_touched_stride, int32_t* n_touched, void* stream);
/* strips: for rank r, [planes][maxh][W] floats starting at strips + r * rank_stride. */
int g4r_shard_assemble(const G4RFrame* frame, int32_t world, int32_t planes, int32_t maxh, int64_t rank_stride, const float* strips,
                       float* images, void* stream);
int g4r_shard_max_count(const int32_t* counts, int32_t world, int32_t* out, void* stream);

/* ---- native runtime of the sharded render: kernels AND collectives of a frame enqueued by a handful of C calls --------------
 * NCCL is resolved at run time from the library the process already loaded (g4r_shard_nccl_load(path or NULL)); the communicator
 * is this library's own: rank 0 calls g4r_shard_nccl_unique_id, the caller distributes the 128 bytes, every rank calls
 * g4r_shard_comm_create (collective).  Per frame:
 *   g4r_shard_forward_a     project, pack, all-reduce(max pair count), all-to-all of the slabs, unpack, bin, sort, composite the
 *                           owned strip into payload[0 .. strip_elems)
 *   g4r_shard_forward_wait  waits for two EARLY events: the largest pair count (> cap: every rank redoes part a with larger slabs)
 *                           and N of the owned strip (> cap_n: this rank re-runs g4r_shard_render_owned with larger buffers)
 *   g4r_shard_forward_b     every plane of the strip sent to its place in every peer's image (grouped send/recv, no padding and no
 *                           assembly pass), n_touched segments back to their owners and summed per Gaussian
 *   g4r_shard_backward      composite backward of the owned strip, reverse all-to-all of the accumulator rows, per-Gaussian sums,
 *                           per-Gaussian backward, all-reduce of dL_dtau (when reduce_pose != 0) */
typedef struct G4RShardComm G4RShardComm;
typedef struct G4RShardBuffers {      /* DEVICE pointers, all owned by the caller */
    int32_t world, rank;
    int64_t cap;                      /* records per (source, destination) pair; slabs have cap + 1 rows (header row last) */
    int64_t cap_n;                    /* instance capacity of binning / sort_scratch */
    void*    geom_local;              /* g4r_geom_bytes(P) */
    int32_t* radii_local;             /* [P] */
    int32_t* n_touched_local;         /* [P] */
    void*    send_slab;               /* [world][cap+1][12] floats */
    void*    recv_slab;               /* [world][cap+1][12] floats; saved for backward */
    int32_t* slots;                   /* [world][P]; saved for backward */
    void*    pack_scratch;            /* g4r_shard_scratch_bytes(P, world) */
    float*   payload;                 /* [payload_elems]: strip [5][maxh][W] | n_touched [world][cap+1] | counts [world] | pad */
    int64_t  strip_elems, maxh, counts_offset, payload_elems;
    float*   gathered;                /* native runtime: [world][cap+1] int32 n_touched segments; Python-orchestrated path: [world][payload_elems] */
    float*   images;                  /* [5][H][W] */
    void*    img_state;               /* g4r_image_bytes; saved */
    void*    binning;                 /* g4r_binning_bytes(cap_n); saved */
    void*    sort_scratch;            /* g4r_sort_scratch_bytes(cap_n) */
    int32_t* radii_all;               /* [world * (cap+1)] */
    int32_t* worst;                   /* [1] */
} G4RShardBuffers;
int g4r_shard_buffers_size(void);                      /* sizeof(G4RShardBuffers): FFI self-check */
int g4r_shard_nccl_load(const char* path);
int g4r_shard_nccl_unique_id(uint8_t* out128);
int g4r_shard_comm_create(const uint8_t* id128, int32_t rank, int32_t world, G4RShardComm** out);
void g4r_shard_comm_destroy(G4RShardComm* comm);
int g4r_shard_forward_a(G4RShardComm* comm, G4RContext* ctx, const G4RFrame* full, const G4RFrame* strip, const G4RGaussians* g,
                        const G4RShardBuffers* b, void* stream);
int g4r_shard_render_owned(G4RContext* ctx, const G4RFrame* strip, const G4RShardBuffers* b, void* stream);
int g4r_shard_forward_wait(G4RContext* ctx, int64_t* N, int64_t* worst);
int g4r_shard_forward_b(G4RShardComm* comm, const G4RFrame* full, int32_t P, const G4RShardBuffers* b, void* stream);
int g4r_shard_backward(G4RShardComm* comm, const G4RFrame* full, const G4RFrame* strip, const G4RGaussians* g, const G4RShardBuffers* b,
                       const float* dL_dcolor, const float* dL_ddepth, void* acc_all, void* acc_back, void* acc_local,
                       const G4RBackwardIO* io, int32_t reduce_pose, void* stream);

/* First / last tile row touched by each Gaussian (rows[2*i], rows[2*i+1]; 1,0 when invisible).  Same rectangle arithmetic as
 * the binning kernels (inspection / tests). */
int g4r_tile_rows(const G4RFrame* frame, int32_t P, const int32_t* radii, const void* geom, int32_t* rows, void* stream);
int g4r_count_tiles(G4RContext* ctx, const G4RFrame* frame, int32_t P_all, const int32_t* radii_all, const void* geom_all,
                    void* img, void* stream);
int g4r_backward_composite(const G4RFrame* frame, int32_t P_all, const void* geom_all, const void* img, const void* binning,
                           const float* dL_dcolor, const float* dL_ddepth, void* acc, void* stream);
int g4r_backward_gaussians(const G4RFrame* frame, const G4RGaussians* g, const int32_t* radii, const void* geom,
                           const void* acc, const G4RBackwardIO* io, void* stream);

/* ---- fused RGB-D losses of the tracking / mapping loops (opt-in) -----------------------------------
 * get_loss_tracking_rgbd (utils/slam_utils.py:57-173; mode 0) and get_loss_mapping_rgbd (:252-364, static non-split branch;
 * mode 1) with their image gradients in one kernel: writes dL_dimage [3,H,W], dL_ddepth [1,H,W] and out4 = {loss, dL/d
 * exposure_a, dL/d exposure_b, 0}.  Masks are uint8 [H,W] (non-zero = keep) or NULL; exposure pointers may be NULL (a = b = 0).
 * scratch32 = 32 bytes of device scratch.  The rendered opacity weights the tracking RGB term as a constant (the rasterizer
 * drops its gradient, DGR/diff_gaussian_rasterization/__init__.py:108). */
typedef struct G4RLossIn {
    int32_t width, height;
    int32_t mode;                 /* 0 tracking, 1 mapping */
    float   alpha;                /* weight of the RGB term (config Training.alpha, default 0.95) */
    float   rgb_boundary_threshold;
    const float* image;           /* [3,H,W] rendered colour */
    const float* depth;           /* [1,H,W] rendered depth */
    const float* opacity;         /* [1,H,W] rendered opacity (tracking) or NULL */
    const float* gt_image;        /* [3,H,W] */
    const float* gt_depth;        /* [1,H,W] */
    const float* exposure_a;      /* [1] or NULL */
    const float* exposure_b;      /* [1] or NULL */
    const uint8_t* motion_mask;   /* [H,W] or NULL */
    const uint8_t* grad_mask;     /* [H,W] or NULL (tracking only) */
} G4RLossIn;
int g4r_slam_loss(const G4RLossIn* in, float* dL_dimage, float* dL_ddepth, float* out4, void* scratch32, void* stream);

/* ---- simple-knn distCUDA2 ----------------------------------------------------------------------
 * Replaces SimpleKNN::knn (submodules/simple-knn/simple_knn.cu:185-220) behind distCUDA2 (spatial.cu:15-26): mean_dist2[i] = mean
 * of the squared distances from points[i] to its 3 nearest other points (FLT_MAX stands in for a missing neighbour when P < 4,
 * giving FLT_MAX / 3 or +inf like the reference).  points = [P,3] float32, mean_dist2 = [P] float32, both DEVICE pointers;
 * scratch = g4r_knn_scratch_bytes(P) bytes of device memory owned by the caller.  Enqueued on `stream`, no host synchronisation
 * and no allocation (the reference allocates, frees and copies the bounding box to the host twice per call).  Bit-identical to
 * the reference build's output. */
size_t g4r_knn_scratch_bytes(int32_t P);
int g4r_knn_mean_dist2(int32_t P, const float* points, float* mean_dist2, void* scratch, size_t scratch_bytes, void* stream);

/* ---- control-node warp of the deformation step (opt-in) ------------------------------------------
 * ControlNodeWarp.forward (utils/time_utils.py:1192-1275) after the node MLP, with cal_nn_weight (:981-1015): K nearest control
 * nodes per Gaussian (replaces pytorch3d.ops.knn_points, :998), Gaussian-kernel weights from exp(log_radius) and
 * sigmoid(weight_logit), blended translation (optionally in the nodes' local frames, :1208-1214), rotation and scale residuals,
 * times motion_mask.  x [N,3] and nodes [M,node_stride] (first 3 columns = position) are constants for autograd like in the
 * reference (:1196, :994).  Forward also writes the neighbour lists nn_idx / nn_dist (squared) / nn_weight [N,K], which the
 * backward reads.  Backward writes the gradients of the node tensors (any output may be NULL); scratch =
 * g4r_warp_scratch_bytes(M) bytes.  1 <= K <= 8. */
typedef struct G4RWarpIn {
    int32_t N, M, K;
    int32_t node_stride;          /* floats per row of `nodes` (3 + hyper_dim) */
    int32_t d_rot_as_res;         /* 1: rotation = sum w rot * mask (:1252); 0: ((sum w (rot + (1,0,0,0))) - (1,0,0,0)) * mask + (1,0,0,0) (:1232) */
    int32_t local_frame;          /* 1: translate through the nodes' local rotations (:1208-1214); 0: sum w d_xyz (:1216) */
    const float* x;               /* [N,3] */
    const float* nodes;           /* [M,node_stride] */
    const float* log_radius;      /* [M]   _node_radius (node_radius = exp, :893) */
    const float* weight_logit;    /* [M]   _node_weight (node_weight = sigmoid, :897) or NULL (with_node_weight off) */
    const float* d_xyz;           /* [M,3] node MLP outputs */
    const float* d_rotation;      /* [M,4] */
    const float* d_scaling;       /* [M,3] */
    const float* local_rotation;  /* [M,4] or NULL when local_frame == 0 */
    const float* motion_mask;     /* [N] or NULL (= 1) */
} G4RWarpIn;
int g4r_warp_forward(const G4RWarpIn* in, float* translate, float* rotation, float* scale, int32_t* nn_idx, float* nn_dist, float* nn_weight,
                     void* stream);
size_t g4r_warp_scratch_bytes(int32_t M);
int g4r_warp_backward(const G4RWarpIn* in, const int32_t* nn_idx, const float* nn_dist, const float* nn_weight, const float* dL_dtranslate,
                      const float* dL_drotation, const float* dL_dscale, float* dL_dd_xyz, float* dL_dd_rotation, float* dL_dd_scaling,
                      float* dL_dlocal_rotation, float* dL_dlog_radius, float* dL_dweight_logit, void* scratch, void* stream);

/* ---- misc ---------------------------------------------------------------------------- */
int g4r_mark_visible(int32_t P, const float* means3D, const float* viewmatrix,
                     const float* projmatrix, uint8_t* present, void* stream);

/* ---- optional per-stage device timing ---------------------------------------------------
 * When enabled (process-wide; single-stream use), every kernel launch of this library is bracketed by CUDA events on the
 * launching stream.  g4r_profile_read() accumulates elapsed milliseconds and launch counts per stage since the
 * last reset; call it after synchronising the stream.  Used by bench.py for the roofline numbers. */
int g4r_profile_enable(int on);
int g4r_profile_stage_count(void);
const char* g4r_profile_stage_name(int stage);
int g4r_profile_read(double* ms_out, int64_t* count_out, int reset);

/* Test / inspection hooks: byte offsets of the saved state inside the caller's buffers
 * (the parity tests read radii, point_list, ranges, n_contrib through these). */
typedef struct G4RLayout {
    size_t geom_rec;        /* float4[3*P]: {mx,my,conic.x,conic.y} {conic.z,opacity,depth,r} {g,b,cull_q,0} */
    size_t geom_clamped;    /* uint8[P]: bit c set when SH colour channel c was clamped at 0 */
    size_t img_final_T;     /* float[H*W]    */
    size_t img_n_contrib;   /* uint32[H*W]   */
    size_t img_ranges;      /* uint2[tiles]  */
    size_t img_counts;      /* uint32[tiles*32]: one counter per 128-byte line */
    size_t img_header;      /* uint32[8]: [0] = N */
    size_t bin_point_list;  /* uint32[capacity] sorted Gaussian ids == reference point_list */
    size_t bin_pairs;       /* uint2[capacity]  unsorted (depth bits, id), offset inside sort_scratch */
} G4RLayout;
int g4r_layout(int32_t P, int32_t width, int32_t height, int64_t capacity, G4RLayout* out);

#ifdef __cplusplus
}
#endif
#endif /* G4R_H_INCLUDED */
