// fine_stage.cu -- CUDA launchers of the OccHead fine / cascade stage (bodies in fine_stage.cuh).
// STATUS: written at the end of round 1 after the GPU budget was spent -- the arithmetic of every body is
// verified on the CPU (tests/test_fine_emul.py runs the same source through tests/emul/fine_emul.cpp against
// torch's grid_sample / group_norm and the pinned oracle), the launches themselves have not run on a B200 yet.
#include <cuda_runtime.h>

#include "../../include/coocc_b200.h"
#include "fine_stage.cuh"

namespace coocc {
namespace fine {

template <typename P, void (*Body)(const P&, long long)>
__global__ void __launch_bounds__(256) items_kernel(const P p, long long n) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL) Body(p, i);
}

template <typename P, void (*Body)(const P&, long long)>
static int run(const P& p, long long n, cudaStream_t st) {
  if (n <= 0) return 0;
  long long b = (n + 255) / 256;
  if (b > 148LL * 32) b = 148LL * 32;
  items_kernel<P, Body><<<(unsigned)b, 256, 0, st>>>(p, n);
  return cudaGetLastError() == cudaSuccess ? 0 : COOCC_ERR_CUDA;
}

}  // namespace fine
}  // namespace coocc

using namespace coocc::fine;

extern "C" int coocc_fine_sample3d_fwd(const float* feats, long long ld, int X, int Y, int Z, int C, const int* coords,
                                       int M, int SX, int SY, int SZ, float* out, long long ldo, void* stream) {
  if (!feats || !coords || !out || C < 4 || (C & 3) || (ld & 3) || (ldo & 3) || M < 0 || SX < 2 || SY < 2 || SZ < 2)
    return COOCC_ERR_ARG;
  Sample3dP p{};
  p.feats = feats; p.ld = ld; p.X = X; p.Y = Y; p.Z = Z; p.C = C; p.coords = coords; p.M = M;
  p.SX = SX; p.SY = SY; p.SZ = SZ; p.out = out; p.ldo = ldo;
  return run<Sample3dP, sample3d_fwd_item>(p, (long long)M * (C >> 2), (cudaStream_t)stream);
}

extern "C" int coocc_fine_sample3d_bwd(const float* gout, long long ldg, int X, int Y, int Z, int C, const int* coords,
                                       int M, int SX, int SY, int SZ, float* dfeats, long long ldd, void* stream) {
  if (!gout || !coords || !dfeats || C < 4 || (C & 3) || (ldg & 3) || M < 0 || SX < 2 || SY < 2 || SZ < 2)
    return COOCC_ERR_ARG;
  Sample3dP p{};
  p.X = X; p.Y = Y; p.Z = Z; p.C = C; p.coords = coords; p.M = M; p.SX = SX; p.SY = SY; p.SZ = SZ;
  p.gout = gout; p.ldg = ldg; p.dfeats = dfeats; p.ldd = ldd;
  return run<Sample3dP, sample3d_bwd_item>(p, (long long)M * (C >> 2), (cudaStream_t)stream);
}

extern "C" int coocc_fine_project(const int* coords, int M, int ncam, const float* vs3, const float* lo3,
                                  const float* inv_bda, const float* cam27, float W_img, float H_img, float* uv,
                                  unsigned char* mask, void* stream) {
  if (!coords || !vs3 || !lo3 || !inv_bda || !cam27 || !uv || !mask || M < 0 || ncam < 1) return COOCC_ERR_ARG;
  ProjectP p{};
  p.coords = coords; p.M = M; p.ncam = ncam;
  for (int a = 0; a < 3; ++a) { p.vs[a] = vs3[a]; p.lo[a] = lo3[a]; }      // host float[3]
  p.inv_bda = inv_bda; p.cam = cam27; p.W_img = W_img; p.H_img = H_img; p.uv = uv; p.mask = mask;
  return run<ProjectP, project_item>(p, (long long)M * ncam, (cudaStream_t)stream);
}

extern "C" int coocc_fine_sample2d_fwd(const float* img, long long ld, int ncam, int H, int W, int C, const float* uv,
                                       const unsigned char* mask, int M, float* out, long long ldo, void* stream) {
  if (!img || !uv || !mask || !out || C < 4 || (C & 3) || (ld & 3) || (ldo & 3) || M < 0) return COOCC_ERR_ARG;
  Sample2dP p{};
  p.img = img; p.ld = ld; p.ncam = ncam; p.H = H; p.W = W; p.C = C; p.uv = uv; p.mask = mask; p.M = M;
  p.out = out; p.ldo = ldo;
  return run<Sample2dP, sample2d_fwd_item>(p, (long long)M * (C >> 2), (cudaStream_t)stream);
}

extern "C" int coocc_fine_sample2d_bwd(const float* gout, long long ldg, int ncam, int H, int W, int C, const float* uv,
                                       const unsigned char* mask, int M, float* dimg, long long ldd, void* stream) {
  if (!gout || !uv || !mask || !dimg || C < 4 || (C & 3) || (ldg & 3) || M < 0) return COOCC_ERR_ARG;
  Sample2dP p{};
  p.ncam = ncam; p.H = H; p.W = W; p.C = C; p.uv = uv; p.mask = mask; p.M = M;
  p.gout = gout; p.ldg = ldg; p.dimg = dimg; p.ldd = ldd;
  return run<Sample2dP, sample2d_bwd_item>(p, (long long)M * (C >> 2), (cudaStream_t)stream);
}

extern "C" int coocc_groupnorm_fwd(const float* x, long long ldx, long long rows, int C, int G, int span,
                                   const float* gamma, const float* beta, float eps, int relu, float* stats, float* y,
                                   long long ldy, void* stream) {
  if (!x || !gamma || !beta || !stats || !y || rows < 0 || C < 1 || G < 1 || C % G || span < 1 || rows % span)
    return COOCC_ERR_ARG;
  GroupNormP p{};
  p.x = x; p.ldx = ldx; p.rows = rows; p.C = C; p.G = G; p.span = span; p.gamma = gamma; p.beta = beta; p.eps = eps;
  p.relu = relu; p.stats = stats; p.y = y; p.ldy = ldy;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = run<GroupNormP, gn_stats_item>(p, rows / span * G, st);
  if (!rc) rc = run<GroupNormP, gn_apply_item>(p, rows * G, st);
  return rc;
}

extern "C" int coocc_groupnorm_bwd(const float* x, long long ldx, long long rows, int C, int G, int span,
                                   const float* gamma, const float* beta, int relu, const float* stats, const float* dy,
                                   long long lddy, float* sums, float* dx, long long lddx, float* dgamma, float* dbeta,
                                   void* stream) {
  if (!x || !gamma || !beta || !stats || !dy || !sums || !dx || !dgamma || !dbeta || rows < 0 || C < 1 || G < 1 || C % G ||
      span < 1 || rows % span)
    return COOCC_ERR_ARG;
  GroupNormP p{};
  p.x = x; p.ldx = ldx; p.rows = rows; p.C = C; p.G = G; p.span = span; p.gamma = gamma; p.beta = beta; p.relu = relu;
  p.stats = const_cast<float*>(stats); p.dy = dy; p.lddy = lddy; p.sums = sums; p.dx = dx; p.lddx = lddx;
  p.dgamma = dgamma; p.dbeta = dbeta;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = run<GroupNormP, gn_bwd_sums_item>(p, rows / span * G, st);
  if (!rc) rc = run<GroupNormP, gn_bwd_apply_item>(p, rows * G, st);
  return rc;
}
