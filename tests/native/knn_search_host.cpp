// Host build of the product's search code (4dgs-slam_b200/csrc/knn_search.cuh) over an index built with std::sort: lets the CPU
// test suite check the search logic (level choice, pruning, termination bound) against the brute-force oracle without a GPU.
// Test harness only -- the product path is knn.cu on the device.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <vector>

#include "knn_search.cuh"

extern "C" int knn_host_mean_dist2(int32_t P, const float* pts, float* out, uint64_t* stats4, uint32_t* evals_per_query) {
    if (P <= 0) return 0;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = 0; i < P; ++i)
        for (int a = 0; a < 3; ++a) { mn[a] = std::min(mn[a], pts[3 * (size_t)i + a]); mx[a] = std::max(mx[a], pts[3 * (size_t)i + a]); }
    const KnnGrid g = knn_make_grid(mn[0], mn[1], mn[2], mx[0], mx[1], mx[2]);
    std::vector<uint64_t> code(P);
    for (int i = 0; i < P; ++i) {
        float u[3]; uint32_t c[3];
        knn_cell(g, pts[3 * (size_t)i], pts[3 * (size_t)i + 1], pts[3 * (size_t)i + 2], u, c);
        code[i] = knn_morton(c[0], c[1], c[2]);
    }
    std::vector<uint32_t> idx(P);
    std::iota(idx.begin(), idx.end(), 0u);
    std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return code[a] < code[b]; });
    std::vector<uint64_t> scode(P);
    std::vector<KnnPoint> sorted(P);
    for (int j = 0; j < P; ++j) {
        const uint32_t i = idx[j];
        scode[j] = code[i];
        sorted[j] = KnnPoint{pts[3 * (size_t)i], pts[3 * (size_t)i + 1], pts[3 * (size_t)i + 2], i};
    }
    int T = 1;
    while (T < 8 && (1ll << (3 * T)) < (long long)P) ++T;
    std::vector<uint32_t> table(((size_t)1 << (3 * T)) + 1);
    const int sh = 63 - 3 * T;
    for (uint32_t t = 0; t <= (1u << (3 * T)); ++t)     // the same table as knn_table_kernel
        table[t] = knn_lower_bound(scode.data(), 0u, (uint32_t)P, (uint64_t)t << sh);
    KnnIndex ix{scode.data(), sorted.data(), table.data(), P, T};
    KnnStats st{0, 0, 0, 0};
    uint64_t tot[4] = {0, 0, 0, 0};
    for (int j = 0; j < P; ++j) {
        st = KnnStats{0, 0, 0, 0};
        out[sorted[j].idx] = knn_query(ix, g, j, &st);
        if (evals_per_query) { evals_per_query[2 * (size_t)sorted[j].idx] = st.evals; evals_per_query[2 * (size_t)sorted[j].idx + 1] = st.nodes; }
        tot[0] += st.evals; tot[1] += st.cells; tot[2] += st.rounds; tot[3] += st.nodes;
    }
    if (stats4) std::memcpy(stats4, tot, sizeof(tot));
    return 0;
}
