/* knn_oracle.c -- CPU restatement of simple-knn's distCUDA2 (TEST INFRASTRUCTURE ONLY: tests/, __graft_entry__.smoke(), bench.py's
 * cpu_baseline may use it; the product never does).
 *
 * Reference: submodules/simple-knn/simple_knn.cu:130-183 (updateKBest<3>, boxMeanDist) and :185-220 (SimpleKNN::knn).  The
 * reference's Morton sort and 1024-point boxes only prune; what it returns is the exact value below, so the restatement is the
 * brute force.  Arithmetic contract, read off the SASS of the reference build (cuobjdump, boxMeanDist: FADD, FADD, FMUL, FADD,
 * FFMA, FFMA, then FADD, FADD and the IEEE division sequence):
 *     d    = other - query                                   (per component, simple_knn.cu:133)
 *     dist = fma(d.z, d.z, fma(d.x, d.x, d.y * d.y))         (:134 after nvcc's default contraction: the FMUL is on the
 *                                                            y component -- registers loaded from offset +4 -- then x, then z)
 *     out  = ((best0 + best1) + best2) / 3.0f                (:182), best* = the three smallest dist over all other points,
 *            FLT_MAX where fewer than three exist (P = 3 gives FLT_MAX / 3, P < 3 overflows to +inf, like the reference)
 * Parity pin: tests/golden/knn_digests_ref.json holds sha256 digests of the UNMODIFIED reference build's output on the
 * bit-reproducible clouds of tools/scenes.py (generated on a B200 by tools/knn_digests.py); tests/test_knn.py checks this
 * restatement and the CUDA path against them. */
#include <float.h>
#include <math.h>
#include <stdint.h>

int knn_oracle_mean_dist2(int32_t P, const float* pts, float* out) {
#pragma omp parallel for schedule(static)
    for (int32_t q = 0; q < P; ++q) {
        float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
        const float qx = pts[3 * (int64_t)q], qy = pts[3 * (int64_t)q + 1], qz = pts[3 * (int64_t)q + 2];
        for (int32_t i = 0; i < P; ++i) {
            if (i == q) continue;
            const float dx = pts[3 * (int64_t)i] - qx, dy = pts[3 * (int64_t)i + 1] - qy, dz = pts[3 * (int64_t)i + 2] - qz;
            float d = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
            for (int j = 0; j < 3; ++j) {
                if (best[j] > d) { const float t = best[j]; best[j] = d; d = t; }
            }
        }
        out[q] = ((best[0] + best[1]) + best[2]) / 3.0f;
    }
    return 0;
}
