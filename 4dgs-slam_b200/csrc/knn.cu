// knn.cu -- distCUDA2 of simple-knn: mean squared distance of every point to its 3 nearest neighbours,
// the value the reference seeds Gaussian scales with once per keyframe (gaussian_splatting/scene/gaussian_model.py:237,381 ->
// submodules/simple-knn/spatial.cu:15-26 -> simple_knn.cu:185-220).
//
// Pipeline (all on `stream`, no host synchronisation, no allocation -- the reference does 1 cudaMalloc, 2 blocking cudaMemcpy and
// 5 thrust allocations per call):
//   knn_bbox_kernel      bounding box with ordered-integer atomics                (reference: 2x cub::DeviceReduce + 2 memcpy D2H)
//   knn_code_kernel      63-bit Morton codes over the cubic box
//   cub::DeviceRadixSort (code, index) pairs -- the same library sort the reference calls (simple_knn.cu:207-211), 64-bit keys
//   knn_gather_kernel    16-byte (x, y, z, index) records in Morton order
//   knn_table_kernel     dense prefix table of the first T octree levels
//   knn_query_kernel     one thread per point: the adaptive 3x3x3 search of knn_search.cuh
// Bytes per point: 12 read + 8+4 written (codes) ; sort ~8 passes x 24 ; gather 12 + 12 read, 16 written ; query 16 + 4 written plus
// ~100-300 neighbour records that hit L1 / L2 (neighbouring threads are neighbours in space).
#include "g4r_common.cuh"
#include "knn_search.cuh"
#include <cub/device/device_radix_sort.cuh>

struct KnnHeader {            // 64 bytes at the start of the scratch buffer; zeroed per call
    uint32_t bb[6];           // [0..2] ~ordered(min), [3..5] ordered(max): both grow under atomicMax from zero
    uint32_t pad[10];
};

static __device__ __forceinline__ uint32_t knn_ordered(float f) {
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
static __device__ __forceinline__ float knn_unordered(uint32_t o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}
static __device__ __forceinline__ KnnGrid knn_load_grid(const KnnHeader* h) {
    return knn_make_grid(knn_unordered(~h->bb[0]), knn_unordered(~h->bb[1]), knn_unordered(~h->bb[2]),
                         knn_unordered(h->bb[3]), knn_unordered(h->bb[4]), knn_unordered(h->bb[5]));
}

__global__ void __launch_bounds__(G4R_BLOCK) knn_bbox_kernel(int P, const float* __restrict__ pts, KnnHeader* hdr) {
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = blockIdx.x * G4R_BLOCK + threadIdx.x; i < P; i += gridDim.x * G4R_BLOCK) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float v = __ldg(pts + 3 * (size_t)i + a);
            mn[a] = fminf(mn[a], v);
            mx[a] = fmaxf(mx[a], v);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], d));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], d));
        }
    }
    __shared__ float s_mn[G4R_BLOCK / 32][3], s_mx[G4R_BLOCK / 32][3];
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) { s_mn[threadIdx.x >> 5][a] = mn[a]; s_mx[threadIdx.x >> 5][a] = mx[a]; }
    }
    __syncthreads();
    if (threadIdx.x < 3) {                           // one thread per axis: 6 atomics per CTA
        const int a = threadIdx.x;
        float lo = s_mn[0][a], hi = s_mx[0][a];
#pragma unroll
        for (int w = 1; w < G4R_BLOCK / 32; ++w) { lo = fminf(lo, s_mn[w][a]); hi = fmaxf(hi, s_mx[w][a]); }
        atomicMax(&hdr->bb[a], ~knn_ordered(lo));
        atomicMax(&hdr->bb[3 + a], knn_ordered(hi));
    }
}

__global__ void __launch_bounds__(G4R_BLOCK) knn_code_kernel(int P, const float* __restrict__ pts, const KnnHeader* __restrict__ hdr,
                                                             uint64_t* __restrict__ code, uint32_t* __restrict__ idx) {
    const int i = blockIdx.x * G4R_BLOCK + threadIdx.x;
    if (i >= P) return;
    const KnnGrid g = knn_load_grid(hdr);
    float u[3];
    uint32_t c[3];
    knn_cell(g, __ldg(pts + 3 * (size_t)i), __ldg(pts + 3 * (size_t)i + 1), __ldg(pts + 3 * (size_t)i + 2), u, c);
    code[i] = knn_morton(c[0], c[1], c[2]);
    idx[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(G4R_BLOCK) knn_gather_kernel(int P, const float* __restrict__ pts, const uint32_t* __restrict__ idx,
                                                               KnnPoint* __restrict__ sorted) {
    const int j = blockIdx.x * G4R_BLOCK + threadIdx.x;
    if (j >= P) return;
    const uint32_t i = idx[j];
    float4 r;
    r.x = __ldg(pts + 3 * (size_t)i); r.y = __ldg(pts + 3 * (size_t)i + 1); r.z = __ldg(pts + 3 * (size_t)i + 2);
    r.w = __uint_as_float(i);
    reinterpret_cast<float4*>(sorted)[j] = r;
}

// table[t] = first sorted position whose T-level prefix is >= t, for t in [0, 8^T]: one binary search per entry (neighbouring
// entries walk the same cache lines).  Filling the gaps from the sorted side instead -- thread j writes the entries between its
// predecessor's prefix and its own -- serialises millions of stores in one thread on clouds with large empty regions (measured:
// 12.8 ms instead of 0.3 ms on a 500 k depth-map cloud).
__global__ void __launch_bounds__(G4R_BLOCK) knn_table_kernel(int P, int T, const uint64_t* __restrict__ code, uint32_t* __restrict__ table) {
    const uint32_t t = blockIdx.x * G4R_BLOCK + threadIdx.x;
    if (t > (1u << (3 * T))) return;
    table[t] = knn_lower_bound(code, 0u, (uint32_t)P, (uint64_t)t << (63 - 3 * T));
}

__global__ void __launch_bounds__(G4R_BLOCK) knn_query_kernel(KnnIndex ix, const KnnHeader* __restrict__ hdr, float* __restrict__ out) {
    const int j = blockIdx.x * G4R_BLOCK + threadIdx.x;
    if (j >= ix.P) return;
    const KnnGrid g = knn_load_grid(hdr);
    const float v = knn_query(ix, g, j, nullptr);
    out[ix.pts[j].idx] = v;
}

namespace {
struct KnnLayout {
    size_t hdr, code_a, code_b, idx_a, idx_b, sorted, table, cub, total, cub_bytes;
    int T;
};
int knn_table_levels(int P) {            // 8^T ~ P: the table costs about as much as one index array
    int T = 1;
    while (T < 8 && (1ll << (3 * T)) < (long long)P) ++T;
    return T;
}
cudaError_t knn_layout(int P, KnnLayout* L) {
    size_t cub_bytes = 0;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const uint32_t*)nullptr,
                                                    (uint32_t*)nullptr, P > 0 ? P : 1, 0, 63);
    if (e != cudaSuccess) return e;
    const size_t n = (size_t)(P > 0 ? P : 1);
    L->T = knn_table_levels(P);
    size_t off = 0;
    L->hdr = off;    off += g4r_align(sizeof(KnnHeader));
    L->code_a = off; off += g4r_align(n * 8);
    L->code_b = off; off += g4r_align(n * 8);
    L->idx_a = off;  off += g4r_align(n * 4);
    L->idx_b = off;  off += g4r_align(n * 4);
    L->sorted = off; off += g4r_align(n * 16);
    L->table = off;  off += g4r_align((((size_t)1 << (3 * L->T)) + 1) * 4);
    L->cub = off;    off += g4r_align(cub_bytes);
    L->cub_bytes = cub_bytes;
    L->total = off;
    return cudaSuccess;
}
}  // namespace

extern "C" size_t g4r_knn_scratch_bytes(int32_t P) {
    KnnLayout L;
    if (P < 0 || knn_layout(P, &L) != cudaSuccess) { g4r_set_error(G4R_EINVAL, "g4r_knn_scratch_bytes(%d): the sort's size query failed", P); return 0; }
    return L.total;
}

extern "C" int g4r_knn_mean_dist2(int32_t P, const float* points, float* mean_dist2, void* scratch, size_t scratch_bytes, void* stream) {
    if (P < 0) return g4r_set_error(G4R_EINVAL, "P = %d is negative", P);
    if (P == 0) return G4R_OK;
    if (!points || !mean_dist2 || !scratch) return g4r_set_error(G4R_EINVAL, "NULL argument");
    KnnLayout L;
    if (knn_layout(P, &L) != cudaSuccess) return g4r_set_error(G4R_ECUDA, "the sort's size query failed");
    if (scratch_bytes < L.total) return g4r_set_error(G4R_EINVAL, "scratch holds %zu bytes, g4r_knn_scratch_bytes(%d) = %zu", scratch_bytes, P, L.total);
    cudaStream_t s = (cudaStream_t)stream;
    char* base = (char*)scratch;
    KnnHeader* hdr = (KnnHeader*)(base + L.hdr);
    uint64_t *code_a = (uint64_t*)(base + L.code_a), *code_b = (uint64_t*)(base + L.code_b);
    uint32_t *idx_a = (uint32_t*)(base + L.idx_a), *idx_b = (uint32_t*)(base + L.idx_b), *table = (uint32_t*)(base + L.table);
    KnnPoint* sorted = (KnnPoint*)(base + L.sorted);
    const int blocks = (P + G4R_BLOCK - 1) / G4R_BLOCK;

    G4R_CUDA_OK(cudaMemsetAsync(hdr, 0, sizeof(KnnHeader), s));
    knn_bbox_kernel<<<blocks < 148 * 8 ? blocks : 148 * 8, G4R_BLOCK, 0, s>>>(P, points, hdr);
    G4R_LAUNCH_OK("knn_bbox_kernel");
    knn_code_kernel<<<blocks, G4R_BLOCK, 0, s>>>(P, points, hdr, code_a, idx_a);
    G4R_LAUNCH_OK("knn_code_kernel");
    size_t cub_bytes = L.cub_bytes;
    G4R_CUDA_OK(cub::DeviceRadixSort::SortPairs(base + L.cub, cub_bytes, (const uint64_t*)code_a, code_b, (const uint32_t*)idx_a, idx_b, P, 0, 63, s));
    knn_gather_kernel<<<blocks, G4R_BLOCK, 0, s>>>(P, points, idx_b, sorted);
    G4R_LAUNCH_OK("knn_gather_kernel");
    knn_table_kernel<<<(int)(((1u << (3 * L.T)) + 1 + G4R_BLOCK - 1) / G4R_BLOCK), G4R_BLOCK, 0, s>>>(P, L.T, code_b, table);
    G4R_LAUNCH_OK("knn_table_kernel");
    KnnIndex ix;
    ix.code = code_b; ix.pts = sorted; ix.table = table; ix.P = P; ix.T = L.T;
    knn_query_kernel<<<blocks, G4R_BLOCK, 0, s>>>(ix, hdr, mean_dist2);
    G4R_LAUNCH_OK("knn_query_kernel");
    return G4R_OK;
}
