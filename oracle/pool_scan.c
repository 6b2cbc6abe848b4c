/* CPU oracle helper (TEST INFRASTRUCTURE ONLY; see oracle/gcn_oracle.py for the rules about who may use oracle/).
 *
 * Plain-C restatement of the reference's neighbour max-pool and of its autograd, used by gcn_oracle.py when the
 * batch is too large for the numpy j-loop (benchmark-scale parity cases). Same arithmetic, same scan order:
 *
 *   forward   openchem/modules/encoders/gcn_encoder.py:44-50   x.repeat -> view(B,N,N,D) -> * adj.expand -> max(dim=2)
 *             cand[j] = adj[b,i,j] * T[b,j,c]; the maximum over j, FIRST index on ties (torch's CPU max keeps the
 *             first index of equal maxima: strict '>' scan from j = 0)
 *   backward  autograd of the above: MaxBackward scatters dP to the arg-max slot, MulBackward multiplies by adj,
 *             the repeat's backward sums over i:  dT[b,j,c] = sum_{i : arg[b,i,c] == j} adj[b,i,j] * dP[b,i,c]
 *             (ascending i, like the numpy restatement's sum over axis 1)
 *
 * Pinned by tests/test_oracle.py: bit-identical to gcn_oracle.neighbour_max_pool / neighbour_max_pool_backward (which
 * are themselves pinned to the unmodified reference by oracle/make_golden.py) on random, tied and signed-zero inputs.
 *
 * Build: gcc -O2 -fopenmp -shared -fPIC oracle/pool_scan.c -o oracle/_build/libpool_scan.so   (__graft_entry__.build()).
 * -ffast-math must NOT be used: the comparisons rely on IEEE semantics.
 */
#include <stdint.h>

#define POOL_FWD(NAME, T_)                                                                              \
  void NAME(const T_* T, const T_* adj, long B, long N, long D, T_* P, int64_t* arg) {                   \
    _Pragma("omp parallel for schedule(static)") for (long b = 0; b < B; ++b) {                          \
      const T_* Tb = T + b * N * D;                                                                      \
      const T_* Ab = adj + b * N * N;                                                                    \
      for (long i = 0; i < N; ++i) {                                                                     \
        T_* Pi = P + (b * N + i) * D;                                                                    \
        int64_t* Gi = arg + (b * N + i) * D;                                                             \
        const T_ a0 = Ab[i * N];                                                                         \
        for (long c = 0; c < D; ++c) {                                                                   \
          Pi[c] = a0 * Tb[c];                                                                            \
          Gi[c] = 0;                                                                                     \
        }                                                                                                \
        for (long j = 1; j < N; ++j) {                                                                   \
          const T_ a = Ab[i * N + j];                                                                    \
          const T_* Tj = Tb + j * D;                                                                     \
          for (long c = 0; c < D; ++c) {                                                                 \
            const T_ cand = a * Tj[c];                                                                   \
            if (cand > Pi[c]) {                                                                          \
              Pi[c] = cand;                                                                              \
              Gi[c] = j;                                                                                 \
            }                                                                                            \
          }                                                                                              \
        }                                                                                                \
      }                                                                                                  \
    }                                                                                                    \
  }

#define POOL_BWD(NAME, T_)                                                                              \
  void NAME(const T_* dP, const T_* adj, const int64_t* arg, long B, long N, long D, T_* dT) {           \
    _Pragma("omp parallel for schedule(static)") for (long b = 0; b < B; ++b) {                          \
      T_* dTb = dT + b * N * D;                                                                          \
      const T_* Ab = adj + b * N * N;                                                                    \
      for (long k = 0; k < N * D; ++k) dTb[k] = (T_)0;                                                   \
      for (long i = 0; i < N; ++i) {                                                                     \
        const T_* Gp = dP + (b * N + i) * D;                                                             \
        const int64_t* Gi = arg + (b * N + i) * D;                                                       \
        for (long c = 0; c < D; ++c) {                                                                   \
          const long j = (long)Gi[c];                                                                    \
          dTb[j * D + c] += Ab[i * N + j] * Gp[c];                                                       \
        }                                                                                                \
      }                                                                                                  \
    }                                                                                                    \
  }

POOL_FWD(pool_fwd_f64, double)
POOL_FWD(pool_fwd_f32, float)
POOL_BWD(pool_bwd_f64, double)
POOL_BWD(pool_bwd_f32, float)

int pool_scan_abi(void) { return 1; }
