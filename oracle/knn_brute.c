/* Oracle (TEST INFRASTRUCTURE ONLY): exhaustive bounded k-NN in plain C.
 *
 * Restates the arithmetic contract of pykdtree's query as called from
 * pyresample/kd_tree.py:545-548 (see oracle/knn_ref.py for the full citation):
 *   d2 = ((dx*dx) + dy*dy) + dz*dz   in the tree dtype, no FMA (-ffp-contract=off)
 *   accept iff d2 < (dtype)(radius*radius)   (strict)
 *   order by (d2, index) ascending; missing -> index n, d2 = inf
 * O(N*M); used to validate the cKDTree-backed oracle and the CUDA index on
 * small and medium inputs.
 */
#include <math.h>
#include <stdint.h>

#define DEFINE_BRUTE(NAME, T)                                                                      \
    void NAME(const T *data, long n, const T *query, long m, int k, double r2, const uint8_t *mask, \
              uint32_t *idx_out, T *d2_out)                                                        \
    {                                                                                              \
        const T dub = (T)r2;                                                                       \
        _Pragma("omp parallel for schedule(dynamic, 64)")                                          \
        for (long j = 0; j < m; ++j) {                                                             \
            uint32_t *bi = idx_out + (long)j * k;                                                  \
            T *bd = d2_out + (long)j * k;                                                          \
            for (int t = 0; t < k; ++t) { bi[t] = (uint32_t)n; bd[t] = (T)INFINITY; }              \
            const T qx = query[3 * j], qy = query[3 * j + 1], qz = query[3 * j + 2];               \
            for (long i = 0; i < n; ++i) {                                                         \
                if (mask && mask[i]) continue;                                                     \
                T dx = qx - data[3 * i], dy = qy - data[3 * i + 1], dz = qz - data[3 * i + 2];     \
                T d2 = (dx * dx + dy * dy) + dz * dz;                                              \
                if (!(d2 < dub)) continue;                                                         \
                /* i ascends, so on equal d2 the earlier (lower) index stays in front */           \
                if (!(d2 < bd[k - 1])) continue;                                                   \
                int t = k - 1;                                                                     \
                while (t > 0 && d2 < bd[t - 1]) { bd[t] = bd[t - 1]; bi[t] = bi[t - 1]; --t; }     \
                bd[t] = d2; bi[t] = (uint32_t)i;                                                   \
            }                                                                                      \
        }                                                                                          \
    }

DEFINE_BRUTE(knn_brute_f32, float)
DEFINE_BRUTE(knn_brute_f64, double)
