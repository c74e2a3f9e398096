// binning.cu -- tile binning: per-tile offsets, instance scatter, per-tile (depth,id) sort.
//
// Replaces the reference's global binning pipeline (DGR/cuda_rasterizer/rasterizer_impl.cu):
//   cub::DeviceScan::InclusiveSum over P Gaussians            (:280)
//   blocking D2H copy of num_rendered                          (:284)
//   duplicateWithKeys                                           (:70-111, :292)
//   cub::DeviceRadixSort::SortPairs on 64-bit (tile|depth) keys (:306-311)  <- 6 HBM passes
//   cudaMemset(ranges) + identifyTileRanges                     (:313-321)
// with a two-level scheme that never sorts across tiles:
//   1. project_kernel already histogrammed instances per TILE (project.cu);
//   2. tile_scan_kernel: exclusive scan over the (few thousand) tiles -> ranges[tile] and N;
//   3. scatter_kernel: each visible Gaussian drops (depth bits, id) into its tiles' segments;
//   4. tile_sort_kernel: one CTA per tile sorts its segment by the 64-bit key (depth<<32 | id)
//      entirely in shared memory and emits the sorted ids.
// Result equivalence: the reference sorts keys (tile<<32 | depth bits) with a STABLE radix sort
// over values emitted in increasing Gaussian id, so inside a tile instances are ordered by
// (depth bits, id).  Sorting each tile segment by (depth bits, id) gives the identical
// point_list; ranges[t] = [start,end) is identical by construction, (0,0) for empty tiles
// exactly like the reference's memset + identifyTileRanges.
#include "g4r_common.cuh"

// ---------------------------------------------------------------------------------------------
// 2. exclusive scan over tiles (single CTA; tiles <= a few 10^4)
// ---------------------------------------------------------------------------------------------
#define SCAN_THREADS 1024
#define ORDER_BUCKETS 256
// Also emits `order`: the tiles grouped into ORDER_BUCKETS classes of decreasing instance count (a counting sort on
// count/max).  The per-tile kernels map blockIdx.x through it, so the hardware block scheduler starts the heaviest
// tiles first and the light ones fill the tail (longest-processing-time-first); results do not depend on it.
// `mirror` (may be NULL): a word of mapped pinned HOST memory that also receives N, so the host can read it after the event
// behind this kernel without a D2H copy in the stream (a 4-byte cudaMemcpyAsync between two kernels costs the stream a copy-engine
// round trip: ~10-18 us of idle device per frame in the profiler trace of the eager loop, tools/e2e_gaps.py).
__global__ void __launch_bounds__(SCAN_THREADS) tile_scan_kernel(uint32_t* __restrict__ counts, uint2* __restrict__ ranges,
                                                                 uint32_t* __restrict__ header, uint32_t* __restrict__ order, int tiles,
                                                                 uint32_t* __restrict__ mirror) {
    __shared__ uint32_t s_warp[SCAN_THREADS / 32];
    __shared__ uint32_t s_wmax[SCAN_THREADS / 32];
    __shared__ uint32_t s_bucket[ORDER_BUCKETS];
    const int per = (tiles + SCAN_THREADS - 1) / SCAN_THREADS;
    const int t0 = threadIdx.x * per, t1 = min(tiles, t0 + per);
    if (threadIdx.x < ORDER_BUCKETS) s_bucket[threadIdx.x] = 0;
    uint32_t sum = 0, cmax = 0;
    for (int t = t0; t < t1; ++t) {
        const uint32_t c = counts[(size_t)t * G4R_COUNT_STRIDE];
        sum += c;
        cmax = max(cmax, c);
    }
    // block-wide exclusive scan of `sum`, block-wide max of the counts
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    cmax = __reduce_max_sync(0xffffffffu, cmax);
    if (lane == 31) s_warp[warp] = incl;
    if (lane == 0) s_wmax[warp] = cmax;
    __syncthreads();
    cmax = __reduce_max_sync(0xffffffffu, s_wmax[lane]);
    const float to_bucket = cmax ? (float)(ORDER_BUCKETS - 1) / (float)cmax : 0.0f;
    if (warp == 0) {
        uint32_t w = s_warp[lane];
        uint32_t wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += o;
        }
        s_warp[lane] = wi - w;                  // exclusive prefix of the warp totals
        if (lane == 31) {
            header[0] = wi;                     // N = total number of (tile, Gaussian) instances
            if (mirror) { *mirror = wi; __threadfence_system(); }
        }
    }
    __syncthreads();
    uint32_t run = s_warp[warp] + incl - sum;
    for (int t = t0; t < t1; ++t) {
        const uint32_t c = counts[(size_t)t * G4R_COUNT_STRIDE];
        counts[(size_t)t * G4R_COUNT_STRIDE] = 0;     // the same word becomes the scatter cursor of this tile
        ranges[t] = c ? make_uint2(run, run + c) : make_uint2(0u, 0u);
        run += c;
        const int b = ORDER_BUCKETS - 1 - min(ORDER_BUCKETS - 1, (int)((float)c * to_bucket));   // heaviest -> bucket 0
        atomicAdd(&s_bucket[b], 1u);
    }
    __syncthreads();
    if (warp == 0) {                            // exclusive scan of the 256 bucket sizes (8 per lane)
        uint32_t v[ORDER_BUCKETS / 32], tot = 0;
#pragma unroll
        for (int i = 0; i < ORDER_BUCKETS / 32; ++i) { v[i] = s_bucket[lane * (ORDER_BUCKETS / 32) + i]; tot += v[i]; }
        uint32_t inc = tot;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += o;
        }
        uint32_t start = inc - tot;
#pragma unroll
        for (int i = 0; i < ORDER_BUCKETS / 32; ++i) { s_bucket[lane * (ORDER_BUCKETS / 32) + i] = start; start += v[i]; }
    }
    __syncthreads();
    for (int t = t0; t < t1; ++t) {
        const uint2 r = ranges[t];
        const int b = ORDER_BUCKETS - 1 - min(ORDER_BUCKETS - 1, (int)((float)(r.y - r.x) * to_bucket));
        order[atomicAdd(&s_bucket[b], 1u)] = (uint32_t)t;
    }
}

int launch_tile_scan(const G4RFrame& f, void* img, cudaStream_t s, uint32_t* mirror) {
    const ImageLayout il(f.width, f.height);
    char* b = (char*)img;
    g4r_stage_begin(ST_TILE_SCAN, s);
    tile_scan_kernel<<<1, SCAN_THREADS, 0, s>>>((uint32_t*)(b + il.counts), (uint2*)(b + il.ranges), (uint32_t*)(b + il.header),
                                                (uint32_t*)(b + il.order), il.tiles, mirror);
    g4r_stage_end(ST_TILE_SCAN, s);
    G4R_LAUNCH_OK("tile_scan_kernel");
    return G4R_OK;
}

// ---------------------------------------------------------------------------------------------
// 3. scatter (depth bits, id) into tile segments
// ---------------------------------------------------------------------------------------------
// Largest N that did not fit its capacity in a forward that ran without host read-back (ctx == NULL: CUDA-graph capture
// and replay).  One word per device (a __device__ variable exists once per device); read and cleared by
// g4r_overflow_status().  Eager forwards re-run phase 2 on overflow and do not touch it.
__device__ unsigned int g_overflow_max = 0;

int g4r_overflow_read(int reset, unsigned int* out) {
    G4R_CUDA_OK(cudaDeviceSynchronize());
    G4R_CUDA_OK(cudaMemcpyFromSymbol(out, g_overflow_max, sizeof(unsigned int)));
    if (reset && *out) {
        const unsigned int zero = 0;
        G4R_CUDA_OK(cudaMemcpyToSymbol(g_overflow_max, &zero, sizeof(unsigned int)));
    }
    return G4R_OK;
}

__global__ void __launch_bounds__(G4R_BLOCK) scatter_kernel(int P, const int32_t* __restrict__ radii, const float4* __restrict__ rec,
                                                            const uint2* __restrict__ ranges, uint32_t* __restrict__ cursors,
                                                            uint2* __restrict__ pairs, uint32_t* __restrict__ header,
                                                            uint32_t capacity, uint32_t gx, uint32_t gy, TileOwner own, bool record_overflow) {
    // header[1] = capacity of this phase-2 run: the backward compares it with N (header[0]) so that a CUDA-graph replay
    // whose instance count outgrew the captured capacity never walks an unwritten point_list.
    if (blockIdx.x == 0 && threadIdx.x == 0) header[1] = capacity;
    if (header[0] > capacity) {                 // uniform: an eager caller re-runs phase 2 with a larger buffer
        if (record_overflow && blockIdx.x == 0 && threadIdx.x == 0) atomicMax(&g_overflow_max, header[0]);
        return;
    }
    const int i = blockIdx.x * G4R_BLOCK + threadIdx.x;
    if (i >= P) return;
    const int radius = radii[i];
    if (radius <= 0) return;
    const float4 a = ldg4(rec + (size_t)i * 3);
    const float4 b = ldg4(rec + (size_t)i * 3 + 1);
    const TileRect r = tile_rect(a.x, a.y, radius, gx, gy);
    const uint32_t key = __float_as_uint(b.z);  // depth bits, as duplicateWithKeys packs them (:104)
    for (uint32_t ty = r.y0; ty < r.y1; ++ty)
        for (uint32_t tx = r.x0; tx < r.x1; ++tx) {
            const uint32_t t = ty * gx + tx;
            if (!own.owns(t, gx)) continue;                                      // not this rank's tile
            const uint32_t slot = __ldg(&ranges[t].x) + atomicAdd(cursors + (size_t)t * G4R_COUNT_STRIDE, 1u);   // cursors start at 0
            pairs[slot] = make_uint2(key, (uint32_t)i);
        }
}

// (Measured and removed again in round 2: a chunked variant in which a CTA histograms 2048 Gaussians per tile in shared memory
// and touches each global counter once per (CTA, non-empty tile) -- 4.5x fewer global atomics.  Scatter 31.6 vs 32.3 us at C3,
// 132 vs 87 us at C4 (profiles/r02_v4_tune_binning_chunked.json): same-address atomics are not what bounds this kernel.)

// Sharded render: per-owned-tile histogram over the all-gathered records (project_kernel's fused histogram only sees
// the local shard).
__global__ void __launch_bounds__(G4R_BLOCK) count_tiles_kernel(int P, const int32_t* __restrict__ radii, const float4* __restrict__ rec,
                                                                uint32_t* __restrict__ counts, uint32_t gx, uint32_t gy, TileOwner own) {
    const int i = blockIdx.x * G4R_BLOCK + threadIdx.x;
    if (i >= P) return;
    const int radius = radii[i];
    if (radius <= 0) return;
    const float4 a = ldg4(rec + (size_t)i * 3);
    const TileRect r = tile_rect(a.x, a.y, radius, gx, gy);
    for (uint32_t ty = r.y0; ty < r.y1; ++ty)
        for (uint32_t tx = r.x0; tx < r.x1; ++tx) {
            const uint32_t t = ty * gx + tx;
            if (own.owns(t, gx)) atomicAdd(counts + (size_t)t * G4R_COUNT_STRIDE, 1u);
        }
}

int launch_count_tiles(const G4RFrame& f, int P, const int32_t* radii, const void* geom, void* img, cudaStream_t s) {
    const GeomLayout gl(P);
    const ImageLayout il(f.width, f.height);
    count_tiles_kernel<<<(P + G4R_BLOCK - 1) / G4R_BLOCK, G4R_BLOCK, 0, s>>>(
        P, radii, (const float4*)((const char*)geom + gl.rec), (uint32_t*)((char*)img + il.counts), (uint32_t)il.tiles_x,
        (uint32_t)il.tiles_y, g4r_owner(f));
    G4R_LAUNCH_OK("count_tiles_kernel");
    return G4R_OK;
}

// First / last tile row of every Gaussian's rectangle: destinations of the all-to-all exchange of the sharded render.
__global__ void __launch_bounds__(G4R_BLOCK) tile_rows_kernel(int P, const int32_t* __restrict__ radii, const float4* __restrict__ rec,
                                                              int32_t* __restrict__ rows, uint32_t gx, uint32_t gy) {
    const int i = blockIdx.x * G4R_BLOCK + threadIdx.x;
    if (i >= P) return;
    const int radius = radii[i];
    int2 r = make_int2(1, 0);
    if (radius > 0) {
        const float4 a = ldg4(rec + (size_t)i * 3);
        const TileRect t = tile_rect(a.x, a.y, radius, gx, gy);
        if (t.x1 > t.x0 && t.y1 > t.y0) r = make_int2((int)t.y0, (int)t.y1 - 1);
    }
    reinterpret_cast<int2*>(rows)[i] = r;
}

int launch_tile_rows(const G4RFrame& f, int P, const int32_t* radii, const void* geom, int32_t* rows, cudaStream_t s) {
    const GeomLayout gl(P);
    const ImageLayout il(f.width, f.height);
    tile_rows_kernel<<<(P + G4R_BLOCK - 1) / G4R_BLOCK, G4R_BLOCK, 0, s>>>(P, radii, (const float4*)((const char*)geom + gl.rec), rows,
                                                                          (uint32_t)il.tiles_x, (uint32_t)il.tiles_y);
    G4R_LAUNCH_OK("tile_rows_kernel");
    return G4R_OK;
}

// ---------------------------------------------------------------------------------------------
// 4. per-tile sort by (depth bits, id)
// ---------------------------------------------------------------------------------------------
#define SORT_SMEM_MAX 4096                      // bitonic path: 32 KB of 64-bit keys
#define BUCKET_MAX_L 2048                       // bucket path: 2 x 16 KB of keys + 8 KB of bucket counters
#define BUCKET_MAX_FILL 24                      // a fuller bucket sends the tile to the bitonic path

// Stable LSD radix pass over one tile segment living in global memory (rare, oversized tiles).
// Sorted entry e of a tile: the id into point_list and, in stream mode, the Gaussian's 48-byte record (id in the spare slot)
// into the sorted splat stream.
struct SortOut {
    uint32_t* point_list;
    float4* stream;            // NULL: no stream
    const float4* rec;
    __device__ __forceinline__ void emit(uint32_t pos, uint32_t id) const {
        point_list[pos] = id;
        if (stream) {
            const float4* r = rec + (size_t)id * 3;
            const float4 a = ldg4(r), b = ldg4(r + 1);
            float4 c = ldg4(r + 2);
            c.w = __uint_as_float(id);
            float4* d = stream + (size_t)pos * 3;
            d[0] = a; d[1] = b; d[2] = c;
        }
    }
};

static __device__ void big_tile_sort(uint2* a, uint2* b, uint32_t L, const SortOut& out, uint32_t pos0, uint32_t* s_hist, uint32_t* s_base) {
    __shared__ int s_skip;
    __shared__ uint32_t s_w[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint2* src = a;
    uint2* dst = b;
    for (int pass = 0; pass < 8; ++pass) {
        const int shift = (pass & 3) * 8;
        const bool hi = pass >= 4;              // passes 0-3: id (secondary key); 4-7: depth bits (primary)
        s_hist[tid] = 0;
        if (tid == 0) s_skip = 0;
        __syncthreads();
        for (uint32_t e = tid; e < L; e += G4R_BLOCK) {
            const uint2 kv = src[e];
            atomicAdd(&s_hist[((hi ? kv.x : kv.y) >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (s_hist[tid] == L) s_skip = 1;       // every key shares this digit: pass is the identity
        __syncthreads();
        if (s_skip) { __syncthreads(); continue; }
        // exclusive scan of the 256 bins
        {
            const uint32_t c = s_hist[tid];
            uint32_t incl = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += o;
            }
            if (lane == 31) s_w[warp] = incl;
            __syncthreads();
            uint32_t off = 0;
            for (int w = 0; w < warp; ++w) off += s_w[w];
            s_base[tid] = off + incl - c;
        }
        __syncthreads();
        // ordered scatter: chunks of 256 in segment order, warps take turns inside a chunk
        for (uint32_t c0 = 0; c0 < L; c0 += G4R_BLOCK) {
            const uint32_t e = c0 + tid;
            const bool valid = e < L;
            uint2 kv = make_uint2(0u, 0u);
            uint32_t digit = 0;
            if (valid) { kv = src[e]; digit = ((hi ? kv.x : kv.y) >> shift) & 255u; }
            const uint32_t act = __ballot_sync(0xffffffffu, valid);
            uint32_t peers = 0, rank = 0;
            if (valid) {
                peers = __match_any_sync(act, digit);
                rank = __popc(peers & ((1u << lane) - 1u));
            }
            for (int w = 0; w < G4R_BLOCK / 32; ++w) {
                if (warp == w && valid) {
                    uint32_t off = 0;
                    const int leader = __ffs(peers) - 1;
                    if (lane == leader) { off = s_base[digit]; s_base[digit] = off + __popc(peers); }
                    off = __shfl_sync(peers, off, leader);
                    dst[off + rank] = kv;
                }
                __syncthreads();
            }
        }
        uint2* t = src; src = dst; dst = t;
        __syncthreads();
    }
    for (uint32_t e = tid; e < L; e += G4R_BLOCK) out.emit(pos0 + e, src[e].y);
}

// Bitonic network over n = 2^m >= L keys in shared memory (keys beyond L are +inf padding).
static __device__ void bitonic_sort(unsigned long long* s_key, uint32_t n) {
    const int tid = threadIdx.x;
    for (uint32_t k = 2; k <= n; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t t = tid; t < (n >> 1); t += G4R_BLOCK) {
                const uint32_t lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const uint32_t hi = lo | j;
                const unsigned long long x = s_key[lo], y = s_key[hi];
                const bool up = (lo & k) == 0;
                if ((x > y) == up) { s_key[lo] = y; s_key[hi] = x; }
            }
            __syncthreads();
        }
    }
}

// One CTA per tile.  Fast path (L <= 2048): one-pass bucket sort -- the depth-bit range [min,max] of the tile is cut
// into nb >= L equal integer intervals (monotone in the key), a counting placement groups the keys by bucket, and each
// (tiny) bucket is insertion-sorted on the full 64-bit key.  ~60 instructions per key instead of ~660 for a bitonic
// network over the next power of two.  Skewed tiles (a bucket fuller than BUCKET_MAX_FILL) and 2048 < L <= 4096 use the
// bitonic network; larger tiles the CTA-local radix sort in global memory.  All three produce the same total order.
__global__ void __launch_bounds__(G4R_BLOCK) tile_sort_kernel(const uint2* __restrict__ ranges, uint2* __restrict__ pairs,
                                                              uint2* __restrict__ pairs_alt, const SortOut out,
                                                              const uint32_t* __restrict__ header, uint32_t capacity,
                                                              const uint32_t* __restrict__ order) {
    if (header[0] > capacity) return;
    __shared__ __align__(16) unsigned long long s_key[SORT_SMEM_MAX];     // bucket path: [0,2048) input, [2048,4096) output
    __shared__ uint32_t s_cnt[BUCKET_MAX_L];                               // bucket counters / cursors
    __shared__ uint32_t s_red[2 * (G4R_BLOCK / 32)];
    const uint2 range = ranges[order ? order[blockIdx.x] : blockIdx.x];
    const uint32_t L = range.y - range.x;
    if (L == 0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (L == 1) { if (tid == 0) out.emit(range.x, pairs[range.x].y); return; }
    if (L > SORT_SMEM_MAX) { big_tile_sort(pairs + range.x, pairs_alt + range.x, L, out, range.x, s_cnt, s_cnt + 256); return; }

    bool use_bitonic = L > BUCKET_MAX_L;
    if (!use_bitonic) {
        // ---- load + depth-bit range --------------------------------------------------------------------------
        uint32_t kmin = 0xffffffffu, kmax = 0u;
        for (uint32_t e = tid; e < L; e += G4R_BLOCK) {
            const uint2 kv = pairs[range.x + e];
            s_key[e] = ((unsigned long long)kv.x << 32) | kv.y;
            kmin = min(kmin, kv.x);
            kmax = max(kmax, kv.x);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, d));
            kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, d));
        }
        if (lane == 0) { s_red[warp] = kmin; s_red[8 + warp] = kmax; }
        uint32_t nb = 64;
        while (nb < L) nb <<= 1;                                         // nb in [64, 2048], nb >= L
        for (uint32_t b = tid; b < nb; b += G4R_BLOCK) s_cnt[b] = 0;
        __syncthreads();
#pragma unroll
        for (int w = 0; w < G4R_BLOCK / 32; ++w) { kmin = min(kmin, s_red[w]); kmax = max(kmax, s_red[8 + w]); }
        // bucket(k) = min(nb-1, trunc(float(k - kmin) * nb / (kmax - kmin + 1))): every step (int->float rz, multiply by a
        // positive constant rz, truncate, clamp) is monotone non-decreasing in k, which is all the sort needs.
        const float bscale = __fdiv_rz((float)nb, __uint2float_rz(kmax - kmin) + 1.0f);
#define BUCKET_OF(k) min(nb - 1u, __float2uint_rz(__fmul_rz(__uint2float_rz((k) - kmin), bscale)))
        // ---- histogram ---------------------------------------------------------------------------------------
        for (uint32_t e = tid; e < L; e += G4R_BLOCK) {
            const uint32_t k = (uint32_t)(s_key[e] >> 32);
            atomicAdd(&s_cnt[BUCKET_OF(k)], 1u);
        }
        __syncthreads();
        // ---- exclusive scan of the nb counters (nb / 256 consecutive buckets per thread) + fill check -----------
        const uint32_t per = nb / G4R_BLOCK;                             // 0 (nb < 256), 1, 2, 4 or 8
        uint32_t local[8];
        uint32_t sum = 0, worst = 0;
        if (per == 0) {
            local[0] = 0;
            if (tid < (int)nb) { local[0] = s_cnt[tid]; sum = local[0]; worst = local[0]; }
        } else {
            for (uint32_t q = 0; q < per; ++q) { local[q] = s_cnt[tid * per + q]; sum += local[q]; worst = max(worst, local[q]); }
        }
        uint32_t incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        __syncthreads();                                                 // everyone is done reading s_red (min/max)
        if (lane == 31) s_red[warp] = incl;
        use_bitonic = __syncthreads_or(worst > BUCKET_MAX_FILL) != 0;
        if (!use_bitonic) {
            uint32_t off = incl - sum;
            for (int w = 0; w < warp; ++w) off += s_red[w];
            if (per == 0) {
                if (tid < (int)nb) s_cnt[tid] = off;
            } else {
                for (uint32_t q = 0; q < per; ++q) { s_cnt[tid * per + q] = off; off += local[q]; }
            }
            __syncthreads();
            // ---- counting placement into the second half of s_key (bucket cursors advance to the bucket ends) ------
            unsigned long long* s_out = s_key + BUCKET_MAX_L;
            for (uint32_t e = tid; e < L; e += G4R_BLOCK) {
                const unsigned long long key = s_key[e];
                const uint32_t k = (uint32_t)(key >> 32);
                const uint32_t pos = atomicAdd(&s_cnt[BUCKET_OF(k)], 1u);
                s_out[pos] = key;
            }
            __syncthreads();
            // ---- insertion sort inside each bucket: bucket b = [end(b-1), end(b)) --------------------------------------
            for (uint32_t b = tid; b < nb; b += G4R_BLOCK) {
                const uint32_t lo = b ? s_cnt[b - 1] : 0u, hi = s_cnt[b];
                for (uint32_t i = lo + 1; i < hi; ++i) {
                    const unsigned long long key = s_out[i];
                    uint32_t j = i;
                    while (j > lo && s_out[j - 1] > key) { s_out[j] = s_out[j - 1]; --j; }
                    s_out[j] = key;
                }
            }
            __syncthreads();
            for (uint32_t e = tid; e < L; e += G4R_BLOCK) out.emit(range.x + e, (uint32_t)s_out[e]);
            return;
        }
#undef BUCKET_OF
    }

    // ---- bitonic path ------------------------------------------------------------------------------------------------
    uint32_t n = 32;
    while (n < L) n <<= 1;
    __syncthreads();
    for (uint32_t e = tid; e < n; e += G4R_BLOCK) {
        unsigned long long k = ~0ull;            // padding sorts to the end
        if (e < L) { const uint2 kv = pairs[range.x + e]; k = ((unsigned long long)kv.x << 32) | kv.y; }
        s_key[e] = k;
    }
    __syncthreads();
    bitonic_sort(s_key, n);
    for (uint32_t e = tid; e < L; e += G4R_BLOCK) out.emit(range.x + e, (uint32_t)s_key[e]);
}

int launch_scatter_sort(const G4RFrame& f, int P, const int32_t* radii, const void* geom, void* img, void* binning,
                        void* sort_scratch, int64_t capacity, bool record_overflow, cudaStream_t s) {
    const GeomLayout gl(P);
    const ImageLayout il(f.width, f.height);
    const BinLayout bl(capacity, g4r_stream_mode());
    const SortLayout sl(capacity);
    char* ib = (char*)img;
    char* bb = (char*)binning;
    char* sb = (char*)sort_scratch;
    const uint32_t cap = (uint32_t)(capacity > 0xffffffffll ? 0xffffffffll : capacity);
    const float4* rec = (const float4*)((const char*)geom + gl.rec);
    g4r_stage_begin(ST_SCATTER, s);
    scatter_kernel<<<(P + G4R_BLOCK - 1) / G4R_BLOCK, G4R_BLOCK, 0, s>>>(P, radii, rec, (const uint2*)(ib + il.ranges),
                                                                         (uint32_t*)(ib + il.counts), (uint2*)(sb + sl.pairs),
                                                                         (uint32_t*)(ib + il.header), cap, (uint32_t)il.tiles_x,
                                                                         (uint32_t)il.tiles_y, g4r_owner(f), record_overflow);
    g4r_stage_end(ST_SCATTER, s);
    G4R_LAUNCH_OK("scatter_kernel");
    static const bool lpt = g4r_tunable("LPT", 1) != 0;
    SortOut sort_out;
    sort_out.point_list = (uint32_t*)(bb + bl.point_list);
    sort_out.stream = g4r_stream_mode() ? (float4*)(bb + bl.stream) : nullptr;
    sort_out.rec = rec;
    g4r_stage_begin(ST_TILE_SORT, s);
    tile_sort_kernel<<<il.tiles, G4R_BLOCK, 0, s>>>((const uint2*)(ib + il.ranges), (uint2*)(sb + sl.pairs), (uint2*)(sb + sl.pairs_alt),
                                                    sort_out, (const uint32_t*)(ib + il.header), cap,
                                                    lpt ? (const uint32_t*)(ib + il.order) : nullptr);
    g4r_stage_end(ST_TILE_SORT, s);
    G4R_LAUNCH_OK("tile_sort_kernel");
    return G4R_OK;
}
