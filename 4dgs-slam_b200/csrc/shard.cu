// shard.cu -- device side of the exchange of the Gaussian-sharded multi-GPU render (BASELINE.json config 4; no
// counterpart in the reference, which is single-GPU).
//
// Rank r owns a contiguous shard of the Gaussians and the contiguous strip of tile rows [r*gy/world, (r+1)*gy/world).
// Forward:  every rank projects its shard, then PACKS, per destination rank d, the 48-byte splat records whose tile rectangle
//           touches d's strip into slab[d][0..count_d) -- a stable compaction, so a record's position is monotone in its local
//           index and (source rank, position) orders the received records like the global Gaussian id; ties in depth therefore
//           sort exactly as on one GPU.  The slabs have a FIXED capacity per (source, destination) pair, so the all-to-all
//           (NCCL, equal splits, over NVLink) needs no host-side sizes and nothing on this path synchronises the host; the
//           true counts travel in a world x world int matrix that the host inspects after everything is enqueued (overflow ->
//           the frame is redone with a larger capacity, like the instance capacity of the single-GPU path).
// Backward: the accumulator rows of the received records travel back the same way and are summed into the owner's rows through
//           the slot table the pack kernel wrote.
#include "g4r_common.cuh"

#define SHARD_MAX_WORLD 16

struct ShardGeom {
    uint32_t gx, gy, world, cap;
    uint32_t row_begin[SHARD_MAX_WORLD + 1];       // strip r = tile rows [row_begin[r], row_begin[r+1])
};

static ShardGeom make_geom(const G4RFrame& f, int world, int64_t cap) {
    const ImageLayout il(f.width, f.height);
    ShardGeom g;
    g.gx = (uint32_t)il.tiles_x; g.gy = (uint32_t)il.tiles_y; g.world = (uint32_t)world;
    g.cap = (uint32_t)(cap > 0x7fffffffll ? 0x7fffffffll : cap);
    for (int r = 0; r <= SHARD_MAX_WORLD; ++r) g.row_begin[r] = r <= world ? (uint32_t)(((uint64_t)r * g.gy) / (uint32_t)world) : g.gy;
    return g;
}

// First / last+1 tile row of Gaussian i (empty range when invisible).
static __device__ __forceinline__ void row_range(int i, int P, const int32_t* __restrict__ radii, const float4* __restrict__ rec,
                                                 uint32_t gx, uint32_t gy, uint32_t& y0, uint32_t& y1) {
    y0 = 1; y1 = 0;
    if (i >= P) return;
    const int radius = radii[i];
    if (radius <= 0) return;
    const float4 a = ldg4(rec + (size_t)i * 3);
    const TileRect t = tile_rect(a.x, a.y, radius, gx, gy);
    if (t.x1 > t.x0 && t.y1 > t.y0) { y0 = t.y0; y1 = t.y1; }
}

// Pass 1: per CTA and destination, how many of the CTA's Gaussians go there.
__global__ void __launch_bounds__(G4R_BLOCK) shard_count_kernel(int P, const int32_t* __restrict__ radii, const float4* __restrict__ rec,
                                                                const ShardGeom g, uint32_t* __restrict__ block_counts, int nblocks) {
    __shared__ uint32_t s_cnt[G4R_BLOCK / 32][SHARD_MAX_WORLD];
    const int i = blockIdx.x * G4R_BLOCK + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t y0, y1;
    row_range(i, P, radii, rec, g.gx, g.gy, y0, y1);
    for (uint32_t d = 0; d < g.world; ++d) {
        const bool to_d = y0 < y1 && y0 < g.row_begin[d + 1] && y1 > g.row_begin[d];
        const uint32_t m = __ballot_sync(0xffffffffu, to_d);
        if (lane == 0) s_cnt[warp][d] = __popc(m);
    }
    __syncthreads();
    if (threadIdx.x < g.world) {
        uint32_t s = 0;
#pragma unroll
        for (int w = 0; w < G4R_BLOCK / 32; ++w) s += s_cnt[w][threadIdx.x];
        block_counts[(size_t)threadIdx.x * nblocks + blockIdx.x] = s;
    }
}

// Pass 2 (one CTA): exclusive scan of the per-CTA counts of every destination, in place; totals -> counts[d].
__global__ void __launch_bounds__(1024) shard_scan_kernel(uint32_t* __restrict__ block_counts, int nblocks, uint32_t world,
                                                          int32_t* __restrict__ counts, float4* __restrict__ slab, uint32_t cap) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t d = 0; d < world; ++d) {
        uint32_t* row = block_counts + (size_t)d * nblocks;
        if (threadIdx.x == 0) s_carry = 0;
        __syncthreads();
        for (int base = 0; base < nblocks; base += 1024) {
            const int j = base + threadIdx.x;
            const uint32_t v = j < nblocks ? row[j] : 0u;
            uint32_t incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) s_warp[warp] = incl;
            __syncthreads();
            if (warp == 0) {
                uint32_t w = s_warp[lane], wi = w;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
                    if (lane >= o) wi += t;
                }
                s_warp[lane] = wi - w;
            }
            __syncthreads();
            const uint32_t carry = s_carry;
            if (j < nblocks) row[j] = carry + s_warp[warp] + incl - v;
            __syncthreads();
            if (threadIdx.x == 1023) s_carry = carry + s_warp[31] + incl;
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            counts[d] = (int32_t)s_carry;
            // header row of slab d (row `cap`, behind the records): the receiver reads the count from here, so the counts
            // travel inside the all-to-all and need no collective of their own
            if (slab) slab[((size_t)d * (cap + 1) + cap) * 3] = make_float4(__int_as_float((int)s_carry), 0.f, 0.f, 0.f);
        }
        __syncthreads();
    }
}

// Pass 3: stable placement.  slots[d][i] = position of Gaussian i in slab d (or -1); records beyond the capacity are dropped
// (the host sees count > capacity and redoes the frame).
__global__ void __launch_bounds__(G4R_BLOCK) shard_pack_kernel(int P, const int32_t* __restrict__ radii, const float4* __restrict__ rec,
                                                               const ShardGeom g, const uint32_t* __restrict__ block_base, int nblocks,
                                                               float4* __restrict__ slab, int32_t* __restrict__ slots) {
    __shared__ uint32_t s_cnt[G4R_BLOCK / 32][SHARD_MAX_WORLD];
    const int i = blockIdx.x * G4R_BLOCK + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t y0, y1;
    row_range(i, P, radii, rec, g.gx, g.gy, y0, y1);
    uint32_t rank_in_warp[SHARD_MAX_WORLD];
    uint32_t dest_mask = 0;
    for (uint32_t d = 0; d < g.world; ++d) {
        const bool to_d = y0 < y1 && y0 < g.row_begin[d + 1] && y1 > g.row_begin[d];
        const uint32_t m = __ballot_sync(0xffffffffu, to_d);
        rank_in_warp[d] = __popc(m & ((1u << lane) - 1u));
        if (to_d) dest_mask |= 1u << d;
        if (lane == 0) s_cnt[warp][d] = __popc(m);
    }
    __syncthreads();
    if (i >= P) return;
    float4 r0, r1, r2;
    if (dest_mask) {
        r0 = ldg4(rec + (size_t)i * 3); r1 = ldg4(rec + (size_t)i * 3 + 1); r2 = ldg4(rec + (size_t)i * 3 + 2);
        r2.w = __int_as_float(radii[i]);                       // the radius rides in the record's spare slot
    }
    for (uint32_t d = 0; d < g.world; ++d) {
        int32_t slot = -1;
        if (dest_mask & (1u << d)) {
            uint32_t s = block_base[(size_t)d * nblocks + blockIdx.x] + rank_in_warp[d];
            for (int w = 0; w < warp; ++w) s += s_cnt[w][d];
            slot = (int32_t)s;
            if (s < g.cap) {
                float4* dst = slab + ((size_t)d * (g.cap + 1) + s) * 3;
                dst[0] = r0; dst[1] = r1; dst[2] = r2;
            }
        }
        slots[(size_t)d * P + i] = slot;
    }
}

// Receiver: radius of every slot of the received slabs (0 for the unused tail of each source's slab).
// Slot j = s * (cap + 1) + k of the received slabs (row `cap` of every slab is its header and counts as unused).
__global__ void __launch_bounds__(G4R_BLOCK) shard_unpack_kernel(uint32_t world, uint32_t cap, const float4* __restrict__ slab,
                                                                 int32_t* __restrict__ radii_all, int32_t* __restrict__ n_touched_all) {
    const size_t j = (size_t)blockIdx.x * G4R_BLOCK + threadIdx.x;
    if (j >= (size_t)world * (cap + 1)) return;
    if (n_touched_all) n_touched_all[j] = 0;                       // accumulated by composite_forward_kernel
    const uint32_t s = (uint32_t)(j / (cap + 1)), k = (uint32_t)(j % (cap + 1));
    const int32_t n = min(__float_as_int(__ldg(reinterpret_cast<const float*>(slab + ((size_t)s * (cap + 1) + cap) * 3))), (int32_t)cap);
    radii_all[j] = (int32_t)k < n ? __float_as_int(__ldg(reinterpret_cast<const float*>(slab + j * 3 + 2) + 3)) : 0;
}

// Owner: acc_local[i] = sum over the destinations Gaussian i was sent to of the accumulator row that came back.
__global__ void __launch_bounds__(G4R_BLOCK) shard_gather_acc_kernel(int P, uint32_t world, uint32_t cap, const int32_t* __restrict__ slots,
                                                                     const float4* __restrict__ acc_back, size_t stride_rows,
                                                                     float4* __restrict__ acc_local) {
    const int i = blockIdx.x * G4R_BLOCK + threadIdx.x;
    if (i >= P) return;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a, c = a;
    for (uint32_t d = 0; d < world; ++d) {
        const int32_t s = slots[(size_t)d * P + i];
        if (s >= 0 && (uint32_t)s < cap) {
            const float4* row = acc_back + ((size_t)d * stride_rows + s) * 3;
            const float4 x = __ldg(row), y = __ldg(row + 1), z = __ldg(row + 2);
            a.x += x.x; a.y += x.y; a.z += x.z; a.w += x.w;
            b.x += y.x; b.y += y.y; b.z += y.z; b.w += y.w;
            c.x += z.x; c.y += z.y; c.z += z.z; c.w += z.w;
        }
    }
    acc_local[(size_t)i * 3] = a; acc_local[(size_t)i * 3 + 1] = b; acc_local[(size_t)i * 3 + 2] = c;
}

__global__ void __launch_bounds__(G4R_BLOCK) shard_gather_int_kernel(int P, uint32_t world, uint32_t cap, const int32_t* __restrict__ slots,
                                                                     const int32_t* __restrict__ back, size_t stride, int32_t* __restrict__ out) {
    const int i = blockIdx.x * G4R_BLOCK + threadIdx.x;
    if (i >= P) return;
    int32_t v = 0;
    for (uint32_t d = 0; d < world; ++d) {
        const int32_t s = slots[(size_t)d * P + i];
        if (s >= 0 && (uint32_t)s < cap) v += __ldg(back + (size_t)d * stride + s);
    }
    out[i] = v;
}

// All-gathered strips (rank r: [planes][maxh][W] at strips + r * rank_stride) -> images [planes][H][W].
// blockIdx.y = plane * H + y; threads cover the row in float4 steps when W % 4 == 0 and the bases are 16-byte aligned.
template <bool kVec>
__global__ void __launch_bounds__(G4R_BLOCK) shard_assemble_kernel(int W, int H, int maxh, size_t rank_stride, const ShardGeom g,
                                                                   const float* __restrict__ strips, float* __restrict__ images) {
    const int pl = blockIdx.y / H, y = blockIdx.y % H;
    const uint32_t ty = (uint32_t)y / G4R_TILE;
    uint32_t r = 0;
    while (r + 1 < g.world && ty >= g.row_begin[r + 1]) ++r;
    const int yy = y - (int)g.row_begin[r] * G4R_TILE;
    const float* src = strips + (size_t)r * rank_stride + ((size_t)pl * maxh + yy) * W;
    float* dst = images + ((size_t)pl * H + y) * W;
    if (kVec) {
        for (int x = blockIdx.x * G4R_BLOCK + threadIdx.x; x < W / 4; x += gridDim.x * G4R_BLOCK)
            reinterpret_cast<float4*>(dst)[x] = __ldg(reinterpret_cast<const float4*>(src) + x);
    } else {
        for (int x = blockIdx.x * G4R_BLOCK + threadIdx.x; x < W; x += gridDim.x * G4R_BLOCK) dst[x] = __ldg(src + x);
    }
}

__global__ void shard_max_count_kernel(const int32_t* __restrict__ counts, int world, int32_t* __restrict__ out) {
    int32_t v = threadIdx.x < world ? counts[threadIdx.x] : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (threadIdx.x == 0) *out = v;
}

static int check_world(const G4RFrame* f, int world) {
    if (!f) return g4r_set_error(G4R_EINVAL, "frame is NULL");
    if (f->width <= 0 || f->height <= 0) return g4r_set_error(G4R_EINVAL, "image size %dx%d is not positive", f->width, f->height);
    if (world < 1 || world > SHARD_MAX_WORLD) return g4r_set_error(G4R_EINVAL, "world %d outside [1, %d]", world, SHARD_MAX_WORLD);
    const ImageLayout il(f->width, f->height);
    if (world > il.tiles_y) return g4r_set_error(G4R_EINVAL, "world %d exceeds the %d tile rows of a %dx%d image: a rank would own an empty strip",
                                                 world, il.tiles_y, f->width, f->height);
    return G4R_OK;
}

extern "C" {

size_t g4r_shard_scratch_bytes(int32_t P, int32_t world) {
    const size_t nblocks = (size_t)((P < 1 ? 1 : P) + G4R_BLOCK - 1) / G4R_BLOCK;
    return g4r_align(nblocks * (size_t)(world < 1 ? 1 : world) * sizeof(uint32_t)) + 256;
}

int g4r_shard_pack(const G4RFrame* f, int32_t P, const int32_t* radii, const void* geom, int32_t world, int64_t cap, void* send_slab,
                   int32_t* counts, int32_t* slots, void* scratch, void* stream) {
    int rc;
    if ((rc = check_world(f, world)) != G4R_OK) return rc;
    if (P < 0 || cap < 1) return g4r_set_error(G4R_EINVAL, "P is negative or the slab capacity is not positive");
    if (!counts) return g4r_set_error(G4R_EINVAL, "counts is NULL");
    cudaStream_t s = (cudaStream_t)stream;
    if (!send_slab) return g4r_set_error(G4R_EINVAL, "send_slab is NULL");
    if (P == 0) {          // nothing to send: zero counts and zero headers
        G4R_CUDA_OK(cudaMemsetAsync(counts, 0, sizeof(int32_t) * world, s));
        for (int d = 0; d < world; ++d)
            G4R_CUDA_OK(cudaMemsetAsync((char*)send_slab + ((size_t)d * (cap + 1) + cap) * 48, 0, 48, s));
        return G4R_OK;
    }
    if (!radii || !geom || !slots || !scratch) return g4r_set_error(G4R_EINVAL, "radii/geom/slots/scratch are NULL");
    if (((uintptr_t)geom | (uintptr_t)send_slab) & 15u) return g4r_set_error(G4R_EINVAL, "geom and send_slab must be 16-byte aligned");
    const GeomLayout gl(P);
    const float4* rec = (const float4*)((const char*)geom + gl.rec);
    const ShardGeom g = make_geom(*f, world, cap);
    const int nblocks = (P + G4R_BLOCK - 1) / G4R_BLOCK;
    uint32_t* bc = (uint32_t*)scratch;
    shard_count_kernel<<<nblocks, G4R_BLOCK, 0, s>>>(P, radii, rec, g, bc, nblocks);
    G4R_LAUNCH_OK("shard_count_kernel");
    shard_scan_kernel<<<1, 1024, 0, s>>>(bc, nblocks, (uint32_t)world, counts, (float4*)send_slab, g.cap);
    G4R_LAUNCH_OK("shard_scan_kernel");
    shard_pack_kernel<<<nblocks, G4R_BLOCK, 0, s>>>(P, radii, rec, g, bc, nblocks, (float4*)send_slab, slots);
    G4R_LAUNCH_OK("shard_pack_kernel");
    return G4R_OK;
}

int g4r_shard_unpack(int32_t world, int64_t cap, const void* recv_slab, int32_t* radii_all, int32_t* n_touched_all, void* stream) {
    if (world < 1 || world > SHARD_MAX_WORLD || cap < 1) return g4r_set_error(G4R_EINVAL, "bad world / capacity");
    if (!recv_slab || !radii_all) return g4r_set_error(G4R_EINVAL, "NULL argument");
    const size_t n = (size_t)world * (size_t)(cap + 1);
    shard_unpack_kernel<<<(unsigned)((n + G4R_BLOCK - 1) / G4R_BLOCK), G4R_BLOCK, 0, (cudaStream_t)stream>>>(
        (uint32_t)world, (uint32_t)cap, (const float4*)recv_slab, radii_all, n_touched_all);
    G4R_LAUNCH_OK("shard_unpack_kernel");
    return G4R_OK;
}

int g4r_shard_gather(int32_t P, int32_t world, int64_t cap, const int32_t* slots, const void* acc_back, int64_t acc_stride_rows,
                     void* acc_local, const int32_t* n_touched_back, int64_t n_touched_stride, int32_t* n_touched, void* stream) {
    if (world < 1 || world > SHARD_MAX_WORLD || cap < 1 || P < 0) return g4r_set_error(G4R_EINVAL, "bad world / capacity / P");
    if (P == 0) return G4R_OK;
    if (!slots) return g4r_set_error(G4R_EINVAL, "slots is NULL");
    cudaStream_t s = (cudaStream_t)stream;
    const int blocks = (P + G4R_BLOCK - 1) / G4R_BLOCK;
    if (acc_back && acc_local) {
        if (((uintptr_t)acc_back | (uintptr_t)acc_local) & 15u) return g4r_set_error(G4R_EINVAL, "accumulator buffers must be 16-byte aligned");
        shard_gather_acc_kernel<<<blocks, G4R_BLOCK, 0, s>>>(P, (uint32_t)world, (uint32_t)cap, slots, (const float4*)acc_back, (size_t)acc_stride_rows,
                                                           (float4*)acc_local);
        G4R_LAUNCH_OK("shard_gather_acc_kernel");
    }
    if (n_touched_back && n_touched) {
        shard_gather_int_kernel<<<blocks, G4R_BLOCK, 0, s>>>(P, (uint32_t)world, (uint32_t)cap, slots, n_touched_back, (size_t)n_touched_stride, n_touched);
        G4R_LAUNCH_OK("shard_gather_int_kernel");
    }
    return G4R_OK;
}

int g4r_shard_assemble(const G4RFrame* f, int32_t world, int32_t planes, int32_t maxh, int64_t rank_stride, const float* strips, float* images,
                       void* stream) {
    int rc;
    if ((rc = check_world(f, world)) != G4R_OK) return rc;
    if (!strips || !images || planes < 1 || maxh < 1) return g4r_set_error(G4R_EINVAL, "bad strips / images / planes / maxh");
    const ShardGeom g = make_geom(*f, world, 1);
    const bool vec = f->width % 4 == 0 && rank_stride % 4 == 0 && ((size_t)maxh * f->width) % 4 == 0 &&
                     (((uintptr_t)strips | (uintptr_t)images) & 15u) == 0;
    const dim3 grid((unsigned)((f->width / (vec ? 4 : 1) + G4R_BLOCK - 1) / G4R_BLOCK), (unsigned)(planes * f->height));
    if (vec) shard_assemble_kernel<true><<<grid, G4R_BLOCK, 0, (cudaStream_t)stream>>>(f->width, f->height, maxh, (size_t)rank_stride, g, strips, images);
    else shard_assemble_kernel<false><<<grid, G4R_BLOCK, 0, (cudaStream_t)stream>>>(f->width, f->height, maxh, (size_t)rank_stride, g, strips, images);
    G4R_LAUNCH_OK("shard_assemble_kernel");
    return G4R_OK;
}

int g4r_shard_max_count(const int32_t* counts, int32_t world, int32_t* out, void* stream) {
    if (!counts || !out || world < 1 || world > 32) return g4r_set_error(G4R_EINVAL, "bad arguments");
    shard_max_count_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(counts, world, out);
    G4R_LAUNCH_OK("shard_max_count_kernel");
    return G4R_OK;
}

int g4r_shard_fetch_counts(G4RContext* ctx, const void* gathered, int64_t rank_stride_bytes, int64_t offset_bytes, int32_t world, void* stream) {
    if (!ctx || !gathered || world < 1 || world > SHARD_MAX_WORLD) return g4r_set_error(G4R_EINVAL, "bad arguments");
    return g4r_context_fetch_matrix(ctx, (const char*)gathered + offset_bytes, (size_t)rank_stride_bytes, world, (cudaStream_t)stream);
}

int g4r_shard_wait_counts(G4RContext* ctx, int32_t world, int32_t* out) {
    if (!ctx || !out || world < 1 || world > SHARD_MAX_WORLD) return g4r_set_error(G4R_EINVAL, "bad arguments");
    return g4r_context_wait_matrix(ctx, world, out);
}

}  // extern "C"
