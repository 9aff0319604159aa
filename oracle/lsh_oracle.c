/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product.
 *
 * Plain-C restatement of SMQTK-Indexing's LSH query arithmetic, used (a) by the
 * tests to cross-check the numpy oracle at sizes numpy is slow at and (b) by
 * bench.py as the CPU baseline ("port") timed on the GPU box's host cores.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.
 *
 * Parity status: PINNED through oracle/np_oracle.py -- tests/test_oracle.py checks
 * every function here against the numpy oracle, which is itself pinned to golden
 * vectors produced by the unmodified reference (tests/golden, oracle/gen_golden.py).
 *
 * Citations are relative to the reference root (/root/reference):
 *   hamming distance   smqtk_indexing/utils/metrics.py:140-155   bin(i ^ j).count('1')
 *   linear scan top-n  smqtk_indexing/impls/hash_index/linear.py:232-240  heapq.nsmallest
 *   ITQ hash           smqtk_indexing/impls/lsh_functor/itq.py:404-408    (x - mean) . R >= 0
 *   distances          smqtk_indexing/utils/metrics.py:49-137
 *
 * The reference is single-threaded Python; this port is deliberately a STRONG
 * baseline (compiled, hardware popcount, OpenMP over queries) so that the GPU/CPU
 * ratio is not flattered by interpreter overhead.
 *
 * Build: gcc -O3 -fopenmp -mpopcnt -shared -fPIC oracle/lsh_oracle.c -o oracle/_build/liblsh_oracle.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* popcount(h xor e) over W 32-bit words  (metrics.py:155) */
static inline int hamming_words(const uint32_t* a, const uint32_t* b, int W) {
  int d = 0;
  int i = 0;
  for (; i + 2 <= W; i += 2) {
    uint64_t x, y;
    memcpy(&x, a + i, 8);
    memcpy(&y, b + i, 8);
    d += __builtin_popcountll(x ^ y);
  }
  for (; i < W; ++i) d += __builtin_popcount(a[i] ^ b[i]);
  return d;
}

/*
 * For each query the k rows with the smallest distance, canonical order
 * ascending (distance, row)  -- linear.py:235-240 with the tie order fixed as in
 * np_oracle.hamming_topk.  out_dist / out_idx: [Q][k], padded with -1.
 */
void oracle_hamming_topk(const uint32_t* db, int64_t U, int W, const uint32_t* q, int Q, int k, int64_t idx_base,
                         int32_t* out_dist, int64_t* out_idx, int nthreads) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
  for (int qi = 0; qi < Q; ++qi) {
    const uint32_t* qw = q + (size_t)qi * W;
    int32_t* bd = out_dist + (size_t)qi * k;
    int64_t* bi = out_idx + (size_t)qi * k;
    int n = 0;
    for (int64_t r = 0; r < U; ++r) {
      const int d = hamming_words(db + (size_t)r * W, qw, W);
      /* rows arrive in ascending order: a tie with the current k-th never displaces it */
      if (n == k && d >= bd[k - 1]) continue;
      int pos = (n < k) ? n : k - 1;
      while (pos > 0 && bd[pos - 1] > d) {
        bd[pos] = bd[pos - 1];
        bi[pos] = bi[pos - 1];
        --pos;
      }
      bd[pos] = d;
      bi[pos] = idx_base + r;
      if (n < k) ++n;
    }
    for (int j = n; j < k; ++j) {
      bd[j] = -1;
      bi[j] = -1;
    }
  }
}

/* z = (x / norm(x) - mean) . R in float64, bits = z >= 0  (itq.py:404-408; norm_kind 0 none, 1 = L2, 2 = L1) */
void oracle_itq_hash(const float* X, int64_t n, int D, const double* mean, const double* R, int b, int norm_kind,
                     uint8_t* bits_out, double* z_out, int nthreads) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
  {
    double* a = (double*)malloc(sizeof(double) * (size_t)D);
    double* z = (double*)malloc(sizeof(double) * (size_t)b);
#pragma omp for schedule(static)
    for (int64_t r = 0; r < n; ++r) {
      const float* x = X + (size_t)r * D;
      double nv = 1.0;
      if (norm_kind) {
        double s = 0.0;
        for (int d = 0; d < D; ++d) s += (norm_kind == 1) ? (double)x[d] * x[d] : fabs((double)x[d]);
        nv = (norm_kind == 1) ? sqrt(s) : s;
        if (nv == 0.0) nv = 1.0;
      }
      for (int d = 0; d < D; ++d) a[d] = (norm_kind ? (double)x[d] / nv : (double)x[d]) - mean[d];
      for (int j = 0; j < b; ++j) z[j] = 0.0;
      for (int d = 0; d < D; ++d) {
        const double ad = a[d];
        const double* rr = R + (size_t)d * b;
        for (int j = 0; j < b; ++j) z[j] += ad * rr[j];
      }
      for (int j = 0; j < b; ++j) {
        bits_out[(size_t)r * b + j] = z[j] >= 0.0;
        if (z_out) z_out[(size_t)r * b + j] = z[j];
      }
    }
    free(a);
    free(z);
  }
}

/* metric 0 euclidean (metrics.py:83-86), 1 cosine (:103-137), 2 hik (:70); one query vs m rows, float64 */
void oracle_distances(const float* q, const float* rows, int64_t m, int D, int metric, double* out) {
  for (int64_t r = 0; r < m; ++r) {
    const float* c = rows + (size_t)r * D;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int d = 0; d < D; ++d) {
      const double a = q[d], b = c[d];
      if (metric == 0) {
        s0 += (a - b) * (a - b);
      } else if (metric == 1) {
        s0 += a * b; s1 += a * a; s2 += b * b;
      } else {
        s0 += a + b - fabs(a - b);
      }
    }
    if (metric == 0) out[r] = sqrt(s0);
    else if (metric == 1) {
      double sim = s0 / (sqrt(s1) * sqrt(s2));
      if (sim == sim) { if (sim > 1.0) sim = 1.0; if (sim < -1.0) sim = -1.0; }
      out[r] = 2.0 * acos(sim) / 3.14159265358979323846;
    } else out[r] = 1.0 - 0.5 * s0;
  }
}
