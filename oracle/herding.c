/* herding.c — CPU oracle for iCaRL barycenter herding.  TEST INFRASTRUCTURE ONLY.
 *
 * Restates continuum.rehearsal.herd_closest_to_barycenter (continuum==1.2.2, requirements.txt:1 of the
 * reference; call site template.py:300-302).  continuum is not vendored under /root/reference and is not
 * installed, so this follows its published algorithm: PARITY UNPINNED (no reference test/vector exists).
 *
 * What this file pins down that NumPy leaves open is the float32 evaluation ORDER, so that the CUDA kernel can
 * be compared bit-exactly: every sum runs sequentially (k = 0..d-1 over features, i = 0..n-1 over samples),
 * multiply and add are separate roundings (compile with -ffp-contract=off), argmax keeps the first maximum.
 *
 * Build: make -C oracle   ->  oracle/_build/libherd_oracle.so
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* features [n][d] float32 row-major, y [n] int64.  Writes, for each class in ascending id order, the indices
 * (into the n samples) of its exemplars in herding-rank order; returns the number written, or -1 on error. */
int64_t herd_barycenter_c(const float* features, const int64_t* y, int64_t n, int d, int nb_per_class,
                          int64_t* out_idx) {
  if (n <= 0 || d <= 0 || nb_per_class < 0) return -1;
  int64_t ymin = y[0], ymax = y[0];
  for (int64_t i = 1; i < n; ++i) {
    if (y[i] < ymin) ymin = y[i];
    if (y[i] > ymax) ymax = y[i];
  }
  int64_t written = 0;
  int64_t* ci = (int64_t*)malloc(sizeof(int64_t) * n);
  float* D = (float*)malloc(sizeof(float) * n * d);
  float* mu = (float*)malloc(sizeof(float) * d);
  float* w = (float*)malloc(sizeof(float) * d);
  int64_t* rank = (int64_t*)malloc(sizeof(int64_t) * n);
  for (int64_t c = ymin; c <= ymax; ++c) {
    int64_t m = 0;
    for (int64_t i = 0; i < n; ++i)
      if (y[i] == c) ci[m++] = i;
    if (m == 0) continue;
    for (int64_t i = 0; i < m; ++i) {
      const float* f = features + ci[i] * d;
      float ss = 0.f;
      for (int k = 0; k < d; ++k) {
        float sq = f[k] * f[k];
        ss = ss + sq;
      }
      float nrm = sqrtf(ss) + 1e-8f;
      for (int k = 0; k < d; ++k) D[i * d + k] = f[k] / nrm;
    }
    for (int k = 0; k < d; ++k) {
      float s = 0.f;
      for (int64_t i = 0; i < m; ++i) s = s + D[i * d + k];
      mu[k] = s / (float)m;
      w[k] = mu[k];
    }
    memset(rank, 0, sizeof(int64_t) * m);
    int64_t quota = nb_per_class < m ? nb_per_class : m;
    int64_t it = 0, eff = 0;
    while (it != quota && eff < 1000) {
      float best = 0.f;
      int64_t bi = -1;
      for (int64_t i = 0; i < m; ++i) {
        float s = 0.f;
        for (int k = 0; k < d; ++k) {
          float pr = w[k] * D[i * d + k];
          s = s + pr;
        }
        if (bi < 0 || s > best) {
          best = s;
          bi = i;
        }
      }
      ++eff;
      if (rank[bi] == 0) {
        rank[bi] = 1 + it;
        ++it;
      }
      for (int k = 0; k < d; ++k) {
        float a = w[k] + mu[k];
        w[k] = a - D[bi * d + k];
      }
    }
    /* stable order: picked samples by rank, then unpicked ones in index order, truncated to nb_per_class */
    int64_t take = nb_per_class < m ? nb_per_class : m;
    int64_t got = 0;
    for (int64_t r = 1; r <= it && got < take; ++r)
      for (int64_t i = 0; i < m; ++i)
        if (rank[i] == r) {
          out_idx[written + got++] = ci[i];
          break;
        }
    for (int64_t i = 0; i < m && got < take; ++i)
      if (rank[i] == 0) out_idx[written + got++] = ci[i];
    written += got;
  }
  free(ci); free(D); free(mu); free(w); free(rank);
  return written;
}

/* herd_cluster_c — continuum.rehearsal.herd_closest_to_cluster (--herding_method cluster, template.py:34,214):
 * per class, the nb_per_class samples whose feature is closest (L2) to the class mean, nearest first.
 * Same conventions as above (PARITY UNPINNED against continuum; fixed float32 order: the mean sums sequentially over
 * the samples, the squared distance sequentially over the feature dimension, separate multiply and add; ties keep
 * the lower index = a stable ascending sort). */
int64_t herd_cluster_c(const float* features, const int64_t* y, int64_t n, int d, int nb_per_class, int64_t* out_idx) {
  if (n <= 0 || d <= 0 || nb_per_class < 0) return -1;
  int64_t ymin = y[0], ymax = y[0];
  for (int64_t i = 1; i < n; ++i) {
    if (y[i] < ymin) ymin = y[i];
    if (y[i] > ymax) ymax = y[i];
  }
  int64_t written = 0;
  int64_t* ci = (int64_t*)malloc(sizeof(int64_t) * n);
  float* mean = (float*)malloc(sizeof(float) * d);
  float* dist = (float*)malloc(sizeof(float) * n);
  for (int64_t c = ymin; c <= ymax; ++c) {
    int64_t m = 0;
    for (int64_t i = 0; i < n; ++i)
      if (y[i] == c) ci[m++] = i;
    if (m == 0) continue;
    for (int k = 0; k < d; ++k) {
      float s = 0.f;
      for (int64_t i = 0; i < m; ++i) s = s + features[ci[i] * d + k];
      mean[k] = s / (float)m;
    }
    for (int64_t i = 0; i < m; ++i) {
      const float* f = features + ci[i] * d;
      float ss = 0.f;
      for (int k = 0; k < d; ++k) {
        float df = mean[k] - f[k];
        float sq = df * df;
        ss = ss + sq;
      }
      dist[i] = sqrtf(ss);
    }
    int64_t take = nb_per_class < m ? nb_per_class : m;
    for (int64_t i = 0; i < m; ++i) {
      int64_t r = 0;
      for (int64_t j = 0; j < m; ++j)
        if (dist[j] < dist[i] || (dist[j] == dist[i] && j < i)) ++r;
      if (r < take) out_idx[written + r] = ci[i];
    }
    written += take;
  }
  free(ci); free(mean); free(dist);
  return written;
}
