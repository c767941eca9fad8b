/*
 * oracle_ops.c -- TEST INFRASTRUCTURE ONLY (the parity checker), never the product path.
 *
 * CPU restatement, in plain C, of the two CUDA-only point ops the reference's
 * GSFusion module calls (the reference ships no CPU implementation of either):
 *
 *   oracle_fps        follows mmdetection3d/mmdet3d/ops/furthest_point_sample/src/
 *                     furthest_point_sample_cuda.cu:11-23 (block size rule, __update)
 *                     and :26-141 (kernel), with the host-side init of
 *                     furthest_point_sample.py:28-33 (temp = 1e10, start index 0).
 *   oracle_ball_query follows mmdetection3d/mmdet3d/ops/ball_query/src/
 *                     ball_query_cuda.cu:11-54, output zero-initialised as in
 *                     ball_query.py:35.
 *
 * Parity pinned: both reproduce the reference's own known-answer tests
 * (mmdetection3d/tests/test_models/test_common_modules/test_pointnet_ops.py:10-74),
 * see tests/test_oracle_golden.py.
 *
 * The CUDA kernel is a SIMT program; what is restated here is its *result*:
 * "thread" t of a block of `bs` threads scans points t, t+bs, ... keeping the
 * first strict maximum, and the shared-memory tree keeps the lower slot on ties.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

static int fps_block_size(int n) {
  /* opt_n_threads(): pow_2 = log(n)/log(2) truncated, clamp to [1, 1024] */
  const int pow_2 = (int)(log((double)n) / log(2.0));
  int bs = 1 << pow_2;
  if (bs > 1024) bs = 1024;
  if (bs < 1) bs = 1;
  return bs;
}

/* xyz: (b, n, 3) float32; idx out: (b, m) int32 */
int oracle_fps(int b, int n, int m, const float *xyz, int *idx) {
  if (m <= 0) return 0;
  const int bs = fps_block_size(n);
  float *temp = (float *)malloc(sizeof(float) * (size_t)n);
  float *dists = (float *)malloc(sizeof(float) * (size_t)bs);
  int *dists_i = (int *)malloc(sizeof(int) * (size_t)bs);
  if (!temp || !dists || !dists_i) return -1;
  for (int bi = 0; bi < b; ++bi) {
    const float *p = xyz + (size_t)bi * n * 3;
    int *out = idx + (size_t)bi * m;
    for (int k = 0; k < n; ++k) temp[k] = 1e10f;
    int old = 0;
    out[0] = old;
    for (int j = 1; j < m; ++j) {
      const float x1 = p[old * 3 + 0], y1 = p[old * 3 + 1], z1 = p[old * 3 + 2];
      for (int t = 0; t < bs; ++t) {
        int besti = 0;
        float best = -1.0f;
        for (int k = t; k < n; k += bs) {
          const float x2 = p[k * 3 + 0], y2 = p[k * 3 + 1], z2 = p[k * 3 + 2];
          const float d = (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1) +
                          (z2 - z1) * (z2 - z1);
          const float d2 = d < temp[k] ? d : temp[k];
          temp[k] = d2;
          besti = d2 > best ? k : besti;
          best = d2 > best ? d2 : best;
        }
        dists[t] = best;
        dists_i[t] = besti;
      }
      for (int s = bs / 2; s >= 1; s >>= 1) {
        for (int t = 0; t < s; ++t) {
          const float v1 = dists[t], v2 = dists[t + s];
          const int i1 = dists_i[t], i2 = dists_i[t + s];
          dists[t] = v1 > v2 ? v1 : v2;
          dists_i[t] = v2 > v1 ? i2 : i1;
        }
      }
      old = dists_i[0];
      out[j] = old;
    }
  }
  free(temp);
  free(dists);
  free(dists_i);
  return 0;
}

/* new_xyz: (b, m, 3) centres; xyz: (b, n, 3); idx out: (b, m, nsample) int32 */
int oracle_ball_query(int b, int n, int m, float min_radius, float max_radius,
                      int nsample, const float *new_xyz, const float *xyz,
                      int *idx) {
  memset(idx, 0, sizeof(int) * (size_t)b * m * nsample);
  const float max_radius2 = max_radius * max_radius;
  const float min_radius2 = min_radius * min_radius;
  for (int bi = 0; bi < b; ++bi) {
    for (int pi = 0; pi < m; ++pi) {
      const float *c = new_xyz + ((size_t)bi * m + pi) * 3;
      const float *p = xyz + (size_t)bi * n * 3;
      int *o = idx + ((size_t)bi * m + pi) * nsample;
      const float new_x = c[0], new_y = c[1], new_z = c[2];
      int cnt = 0;
      for (int k = 0; k < n; ++k) {
        const float x = p[k * 3 + 0], y = p[k * 3 + 1], z = p[k * 3 + 2];
        const float d2 = (new_x - x) * (new_x - x) + (new_y - y) * (new_y - y) +
                         (new_z - z) * (new_z - z);
        if (d2 == 0 || (d2 >= min_radius2 && d2 < max_radius2)) {
          if (cnt == 0)
            for (int l = 0; l < nsample; ++l) o[l] = k;
          o[cnt] = k;
          ++cnt;
          if (cnt >= nsample) break;
        }
      }
    }
  }
  return 0;
}
