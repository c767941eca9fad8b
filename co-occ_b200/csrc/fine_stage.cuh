// fine_stage.cuh -- per-work-item bodies of the OccHead fine / cascade stage kernels:
//   trilinear sampling of out_voxel_feats at fine voxel centres   occ_head.py:212-221 (F.grid_sample 5-D,
//                                                                   mode='bilinear', zeros padding, align_corners=False)
//   projection of fine voxels onto the cameras                    P/utils/coordinate_transform.py:29-70
//   bilinear sampling of the image features + masked camera sum   occ_head.py:231-233 (F.grid_sample 4-D, align_corners=True)
//   GroupNorm(16) on point rows and on image feature maps         occ_head.py:58-78 (img_mlp_0 / img_mlp / fine_mlp)
//
// Every kernel is "one work item = one call of a body function"; the bodies are __host__ __device__ so that
// the same source is (a) launched as CUDA kernels by fine_stage.cu and (b) looped over on the CPU by the test
// harness tests/emul/fine_emul.cpp, which checks the arithmetic against the oracle without a GPU.  The harness is
// test infrastructure: the product only ever runs the __global__ wrappers.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define COOCC_HD __host__ __device__ __forceinline__
#else
#define COOCC_HD inline
#endif

namespace coocc {
namespace fine {

COOCC_HD void atomic_addf(float* p, float v) {
#if defined(__CUDA_ARCH__)
  atomicAdd(p, v);
#else
  *p += v;          // the CPU harness runs the items one after another
#endif
}
// four consecutive floats at a 16-byte aligned address: one vector reduction instead of four scalar atomics
COOCC_HD void atomic_add4f(float* p, float a, float b, float c, float d) {
#if defined(__CUDA_ARCH__)
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
#else
  p[0] += a; p[1] += b; p[2] += c; p[3] += d;
#endif
}

// ------------------------------------------------------------------------------------------
// 3-D sampling: out[m, c] = trilinear(feats, fine_coord[:, m]),  feats = [X*Y*Z][ld] NDHWC rows
// ------------------------------------------------------------------------------------------
struct Sample3dP {
  const float* feats; long long ld;
  const unsigned short* feats_bf16;     // non-null: the grid is stored in bf16 (feats unused), row stride ld elements
  int X, Y, Z, C;
  const int* coords;     // [3][M] fine voxel indices (x, y, z)
  int M;
  int SX, SY, SZ;        // final_occ_size
  float* out; long long ldo;           // fwd: [M][ldo]
  const float* gout; long long ldg;    // bwd: [M][ldg]
  float* dfeats; long long ldd;        // bwd: [X*Y*Z][ldd], accumulated with atomics (zeroed by the caller)
};

struct Corner3 { int v[8]; float w[8]; };

// bf16 storage bits -> fp32 (host + device)
COOCC_HD float bf16_bits(unsigned short b) {
  const unsigned int u = (unsigned int)b << 16;
  float f;
  memcpy(&f, &u, sizeof(f));
  return f;
}

// grid = (c / (S - 1) - 0.5) * 2 (occ_head.py:214-216), unnormalised with align_corners=False:
// ((g + 1) * size - 1) / 2; corners in ATen's order tnw, tne, tsw, tse, bnw, bne, bsw, bse
COOCC_HD Corner3 corners3(const Sample3dP& p, int m) {
  const float gx = ((float)p.coords[m] / (float)(p.SX - 1) - 0.5f) * 2.f;
  const float gy = ((float)p.coords[p.M + m] / (float)(p.SY - 1) - 0.5f) * 2.f;
  const float gz = ((float)p.coords[2 * p.M + m] / (float)(p.SZ - 1) - 0.5f) * 2.f;
  const float ix = ((gx + 1.f) * (float)p.X - 1.f) / 2.f;
  const float iy = ((gy + 1.f) * (float)p.Y - 1.f) / 2.f;
  const float iz = ((gz + 1.f) * (float)p.Z - 1.f) / 2.f;
  const float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
  const int x0 = (int)fx, y0 = (int)fy, z0 = (int)fz;
  const float ax = ix - fx, ay = iy - fy, az = iz - fz;          // distance to the lower corner
  const float bx = (fx + 1.f) - ix, by = (fy + 1.f) - iy, bz = (fz + 1.f) - iz;
  Corner3 c;
  // (dz, dy, dx) in {0,1}^3, x fastest: tnw(0,0,0) tne(0,0,1) tsw(0,1,0) tse(0,1,1) bnw(1,0,0) ...
  for (int k = 0; k < 8; ++k) {
    const int dx = k & 1, dy = (k >> 1) & 1, dz = k >> 2;
    const int x = x0 + dx, y = y0 + dy, z = z0 + dz;
    const bool in = x >= 0 && x < p.X && y >= 0 && y < p.Y && z >= 0 && z < p.Z;
    c.v[k] = in ? (x * p.Y + y) * p.Z + z : -1;
    c.w[k] = (dx ? ax : bx) * (dy ? ay : by) * (dz ? az : bz);
  }
  return c;
}

// item = m * (C/4) + c4
COOCC_HD void sample3d_fwd_item(const Sample3dP& p, long long id) {
  const int c4n = p.C >> 2;
  const int m = (int)(id / c4n), c = (int)(id % c4n) * 4;
  const Corner3 k = corners3(p, m);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  for (int j = 0; j < 8; ++j) {
    if (k.v[j] < 0) continue;
    if (p.feats_bf16 != nullptr) {
      const unsigned short* s = p.feats_bf16 + (long long)k.v[j] * p.ld + c;
      a0 += bf16_bits(s[0]) * k.w[j]; a1 += bf16_bits(s[1]) * k.w[j];
      a2 += bf16_bits(s[2]) * k.w[j]; a3 += bf16_bits(s[3]) * k.w[j];
      continue;
    }
    const float* s = p.feats + (long long)k.v[j] * p.ld + c;
    a0 += s[0] * k.w[j]; a1 += s[1] * k.w[j]; a2 += s[2] * k.w[j]; a3 += s[3] * k.w[j];
  }
  float* o = p.out + (long long)m * p.ldo + c;
  o[0] = a0; o[1] = a1; o[2] = a2; o[3] = a3;
}

COOCC_HD void sample3d_bwd_item(const Sample3dP& p, long long id) {
  const int c4n = p.C >> 2;
  const int m = (int)(id / c4n), c = (int)(id % c4n) * 4;
  const Corner3 k = corners3(p, m);
  const float* g = p.gout + (long long)m * p.ldg + c;
  for (int j = 0; j < 8; ++j) {
    if (k.v[j] < 0) continue;
    float* d = p.dfeats + (long long)k.v[j] * p.ldd + c;
    atomic_add4f(d, g[0] * k.w[j], g[1] * k.w[j], g[2] * k.w[j], g[3] * k.w[j]);
  }
}

// ------------------------------------------------------------------------------------------
// projection onto the cameras: uv [ncam][M][2] (grid_sample convention), mask [M][ncam]
// ------------------------------------------------------------------------------------------
struct ProjectP {
  const int* coords; int M;                 // [3][M]
  int ncam;
  float vs[3], lo[3];                       // voxel size, lower corner of point_cloud_range (fp32, computed like :33)
  const float* inv_bda;                     // [9]
  const float* cam;                         // [ncam][9 inv_rots | 3 trans | 9 intrins | 4 post_rots[:2,:2] | 2 post_trans[:2]] = 27 floats
  float W_img, H_img;
  float* uv; unsigned char* mask;
};

COOCC_HD void mat3(const float* A, const float* x, float* y) {
  y[0] = A[0] * x[0] + A[1] * x[1] + A[2] * x[2];
  y[1] = A[3] * x[0] + A[4] * x[1] + A[5] * x[2];
  y[2] = A[6] * x[0] + A[7] * x[1] + A[8] * x[2];
}

// item = m * ncam + cam
COOCC_HD void project_item(const ProjectP& p, long long id) {
  const int m = (int)(id / p.ncam), cam = (int)(id % p.ncam);
  float q[3], r[3];
  for (int a = 0; a < 3; ++a) q[a] = (float)p.coords[a * p.M + m] * p.vs[a] + p.lo[a];      // :34
  mat3(p.inv_bda, q, r);                                                                  // :38-39
  const float* c = p.cam + cam * 27;
  for (int a = 0; a < 3; ++a) r[a] -= c[9 + a];                                           // :47
  mat3(c, r, q);                                                                          // :48-49
  mat3(c + 12, q, r);                                                                     // :53
  const float d = r[2];
  const float u0 = r[0] / (d + 1e-5f), v0 = r[1] / (d + 1e-5f);                           // :58-59
  float u = c[21] * u0 + c[22] * v0 + c[25];                                              // :62-63
  float v = c[23] * u0 + c[24] * v0 + c[26];
  u = (u / (p.W_img - 1.f) - 0.5f) * 2.f;                                                 // :65-66
  v = (v / (p.H_img - 1.f) - 0.5f) * 2.f;
  float* o = p.uv + ((long long)cam * p.M + m) * 2;
  o[0] = u; o[1] = v;
  p.mask[(long long)m * p.ncam + cam] = (d > 1e-5f && u > -1.f && u < 1.f && v > -1.f && v < 1.f) ? 1 : 0;   // :68-70
}

// ------------------------------------------------------------------------------------------
// 2-D sampling: out[m, c] = sum_cam mask[m, cam] * bilinear(img[cam], uv[cam, m]),  img = [ncam*H*W][ld] NHWC rows
// ------------------------------------------------------------------------------------------
struct Sample2dP {
  const float* img; long long ld;
  int ncam, H, W, C;
  const float* uv; const unsigned char* mask; int M;
  float* out; long long ldo;
  const float* gout; long long ldg;
  float* dimg; long long ldd;
};

struct Corner2 { int v[4]; float w[4]; };

// align_corners=True: ix = (u + 1) / 2 * (W - 1); corners nw, ne, sw, se (ATen grid_sampler_2d)
COOCC_HD Corner2 corners2(const Sample2dP& p, int cam, int m) {
  const float* q = p.uv + ((long long)cam * p.M + m) * 2;
  const float ix = ((q[0] + 1.f) / 2.f) * (float)(p.W - 1);
  const float iy = ((q[1] + 1.f) / 2.f) * (float)(p.H - 1);
  const float fx = floorf(ix), fy = floorf(iy);
  const int x0 = (int)fx, y0 = (int)fy;
  const float ax = ix - fx, ay = iy - fy, bx = (fx + 1.f) - ix, by = (fy + 1.f) - iy;
  Corner2 c;
  for (int k = 0; k < 4; ++k) {
    const int dx = k & 1, dy = k >> 1;
    const int x = x0 + dx, y = y0 + dy;
    const bool in = x >= 0 && x < p.W && y >= 0 && y < p.H;
    c.v[k] = in ? (cam * p.H + y) * p.W + x : -1;
    c.w[k] = (dx ? ax : bx) * (dy ? ay : by);
  }
  return c;
}

// item = m * (C/4) + c4
COOCC_HD void sample2d_fwd_item(const Sample2dP& p, long long id) {
  const int c4n = p.C >> 2;
  const int m = (int)(id / c4n), c = (int)(id % c4n) * 4;
  float a[4] = {0.f, 0.f, 0.f, 0.f};
  for (int cam = 0; cam < p.ncam; ++cam) {
    if (!p.mask[(long long)m * p.ncam + cam]) continue;           // the reference multiplies by the mask (:233)
    const Corner2 k = corners2(p, cam, m);
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j < 4; ++j) {
      if (k.v[j] < 0) continue;
      const float* src = p.img + (long long)k.v[j] * p.ld + c;
      for (int u = 0; u < 4; ++u) s[u] += src[u] * k.w[j];
    }
    for (int u = 0; u < 4; ++u) a[u] += s[u];
  }
  float* o = p.out + (long long)m * p.ldo + c;
  for (int u = 0; u < 4; ++u) o[u] = a[u];
}

COOCC_HD void sample2d_bwd_item(const Sample2dP& p, long long id) {
  const int c4n = p.C >> 2;
  const int m = (int)(id / c4n), c = (int)(id % c4n) * 4;
  const float* g = p.gout + (long long)m * p.ldg + c;
  for (int cam = 0; cam < p.ncam; ++cam) {
    if (!p.mask[(long long)m * p.ncam + cam]) continue;
    const Corner2 k = corners2(p, cam, m);
    for (int j = 0; j < 4; ++j) {
      if (k.v[j] < 0) continue;
      float* d = p.dimg + (long long)k.v[j] * p.ldd + c;
      atomic_add4f(d, g[0] * k.w[j], g[1] * k.w[j], g[2] * k.w[j], g[3] * k.w[j]);
    }
  }
}

// ------------------------------------------------------------------------------------------
// GroupNorm + ReLU.  x = [N rows][ld] with C channels in G groups; `span` consecutive rows form one sample
// (span = 1: nn.GroupNorm on [M, C] point rows; span = H*W: nn.GroupNorm on an NHWC image feature map).
// Statistics per (sample, group) over span * C/G values, biased variance, eps inside the sqrt.
// ------------------------------------------------------------------------------------------
struct GroupNormP {
  const float* x; long long ldx;
  long long rows; int C, G, span;
  const float* gamma; const float* beta; float eps; int relu;
  float* stats;            // [rows/span][G][2] mean, rstd
  float* y; long long ldy;
  // backward
  const float* dy; long long lddy;
  float* sums;             // [rows/span][G][2] sum(dy*gamma), sum(dy*gamma*xhat)   (dy already masked by the ReLU)
  float* dx; long long lddx;
  float* dgamma; float* dbeta;       // [C], accumulated with atomics (zeroed by the caller)
};

// item = sample * G + g
COOCC_HD void gn_stats_item(const GroupNormP& p, long long id) {
  const long long s = id / p.G;
  const int g = (int)(id % p.G), cpg = p.C / p.G;
  double sum = 0.0, sq = 0.0;
  for (long long r = s * p.span; r < (s + 1) * p.span; ++r) {
    const float* row = p.x + r * p.ldx + g * cpg;
    for (int c = 0; c < cpg; ++c) { sum += row[c]; sq += (double)row[c] * row[c]; }
  }
  const double n = (double)p.span * cpg;
  const double mean = sum / n;
  double var = sq / n - mean * mean;
  if (var < 0.0) var = 0.0;
  p.stats[id * 2 + 0] = (float)mean;
  p.stats[id * 2 + 1] = (float)(1.0 / sqrt(var + (double)p.eps));
}

// item = row * G + g
COOCC_HD void gn_apply_item(const GroupNormP& p, long long id) {
  const long long r = id / p.G;
  const int g = (int)(id % p.G), cpg = p.C / p.G;
  const float* st = p.stats + ((r / p.span) * p.G + g) * 2;
  const float* row = p.x + r * p.ldx + g * cpg;
  float* out = p.y + r * p.ldy + g * cpg;
  for (int c = 0; c < cpg; ++c) {
    float v = (row[c] - st[0]) * st[1] * p.gamma[g * cpg + c] + p.beta[g * cpg + c];
    if (p.relu && v < 0.f) v = 0.f;
    out[c] = v;
  }
}

// backward phase 1, item = sample * G + g: group sums; also dgamma / dbeta contributions
COOCC_HD void gn_bwd_sums_item(const GroupNormP& p, long long id) {
  const long long s = id / p.G;
  const int g = (int)(id % p.G), cpg = p.C / p.G;
  const float* st = p.stats + id * 2;
  double s1 = 0.0, s2 = 0.0;
  for (int c = 0; c < cpg; ++c) {
    const int ch = g * cpg + c;
    double dg = 0.0, db = 0.0;
    for (long long r = s * p.span; r < (s + 1) * p.span; ++r) {
      const float xh = (p.x[r * p.ldx + ch] - st[0]) * st[1];
      float d = p.dy[r * p.lddy + ch];
      if (p.relu && !(xh * p.gamma[ch] + p.beta[ch] > 0.f)) d = 0.f;
      dg += (double)d * xh;
      db += d;
    }
    s1 += db * p.gamma[ch];
    s2 += dg * p.gamma[ch];
    atomic_addf(p.dgamma + ch, (float)dg);
    atomic_addf(p.dbeta + ch, (float)db);
  }
  p.sums[id * 2 + 0] = (float)s1;
  p.sums[id * 2 + 1] = (float)s2;
}

// backward phase 2, item = row * G + g:  dx = rstd * (dy*gamma - mean_g(dy*gamma) - xhat * mean_g(dy*gamma*xhat))
COOCC_HD void gn_bwd_apply_item(const GroupNormP& p, long long id) {
  const long long r = id / p.G;
  const int g = (int)(id % p.G), cpg = p.C / p.G;
  const long long sg = (r / p.span) * p.G + g;
  const float* st = p.stats + sg * 2;
  const float inv_n = 1.f / ((float)p.span * (float)cpg);
  const float m1 = p.sums[sg * 2 + 0] * inv_n, m2 = p.sums[sg * 2 + 1] * inv_n;
  for (int c = 0; c < cpg; ++c) {
    const int ch = g * cpg + c;
    const float xh = (p.x[r * p.ldx + ch] - st[0]) * st[1];
    float d = p.dy[r * p.lddy + ch];
    if (p.relu && !(xh * p.gamma[ch] + p.beta[ch] > 0.f)) d = 0.f;
    p.dx[r * p.lddx + ch] = st[1] * (d * p.gamma[ch] - m1 - xh * m2);
  }
}

}  // namespace fine
}  // namespace coocc
