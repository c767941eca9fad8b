// fine_emul.cpp -- TEST INFRASTRUCTURE ONLY.  Runs the work-item bodies of co-occ_b200/csrc/fine_stage.cuh on the
// CPU, one item after another, with the same entry-point signatures as the CUDA launchers in fine_stage.cu (minus
// the stream), so that tests/test_fine_emul.py can check the kernels' arithmetic against torch / the oracle
// without a GPU.  Never linked into the product library.
#include "../../co-occ_b200/csrc/fine_stage.cuh"

using namespace coocc::fine;

template <typename P, void (*Body)(const P&, long long)>
static void loop(const P& p, long long n) {
  for (long long i = 0; i < n; ++i) Body(p, i);
}

extern "C" {
int emul_sample3d_fwd(const float* feats, long long ld, int X, int Y, int Z, int C, const int* coords, int M, int SX,
                      int SY, int SZ, float* out, long long ldo) {
  Sample3dP p{};
  p.feats = feats; p.ld = ld; p.X = X; p.Y = Y; p.Z = Z; p.C = C; p.coords = coords; p.M = M;
  p.SX = SX; p.SY = SY; p.SZ = SZ; p.out = out; p.ldo = ldo;
  loop<Sample3dP, sample3d_fwd_item>(p, (long long)M * (C >> 2));
  return 0;
}
int emul_sample3d_bwd(const float* gout, long long ldg, int X, int Y, int Z, int C, const int* coords, int M, int SX,
                      int SY, int SZ, float* dfeats, long long ldd) {
  Sample3dP p{};
  p.X = X; p.Y = Y; p.Z = Z; p.C = C; p.coords = coords; p.M = M; p.SX = SX; p.SY = SY; p.SZ = SZ;
  p.gout = gout; p.ldg = ldg; p.dfeats = dfeats; p.ldd = ldd;
  loop<Sample3dP, sample3d_bwd_item>(p, (long long)M * (C >> 2));
  return 0;
}
int emul_project(const int* coords, int M, int ncam, const float* vs3, const float* lo3, const float* inv_bda,
                 const float* cam27, float W_img, float H_img, float* uv, unsigned char* mask) {
  ProjectP p{};
  p.coords = coords; p.M = M; p.ncam = ncam;
  for (int a = 0; a < 3; ++a) { p.vs[a] = vs3[a]; p.lo[a] = lo3[a]; }
  p.inv_bda = inv_bda; p.cam = cam27; p.W_img = W_img; p.H_img = H_img; p.uv = uv; p.mask = mask;
  loop<ProjectP, project_item>(p, (long long)M * ncam);
  return 0;
}
int emul_sample2d_fwd(const float* img, long long ld, int ncam, int H, int W, int C, const float* uv,
                      const unsigned char* mask, int M, float* out, long long ldo) {
  Sample2dP p{};
  p.img = img; p.ld = ld; p.ncam = ncam; p.H = H; p.W = W; p.C = C; p.uv = uv; p.mask = mask; p.M = M;
  p.out = out; p.ldo = ldo;
  loop<Sample2dP, sample2d_fwd_item>(p, (long long)M * (C >> 2));
  return 0;
}
int emul_sample2d_bwd(const float* gout, long long ldg, int ncam, int H, int W, int C, const float* uv,
                      const unsigned char* mask, int M, float* dimg, long long ldd) {
  Sample2dP p{};
  p.ncam = ncam; p.H = H; p.W = W; p.C = C; p.uv = uv; p.mask = mask; p.M = M;
  p.gout = gout; p.ldg = ldg; p.dimg = dimg; p.ldd = ldd;
  loop<Sample2dP, sample2d_bwd_item>(p, (long long)M * (C >> 2));
  return 0;
}
int emul_groupnorm_fwd(const float* x, long long ldx, long long rows, int C, int G, int span, const float* gamma,
                       const float* beta, float eps, int relu, float* stats, float* y, long long ldy) {
  GroupNormP p{};
  p.x = x; p.ldx = ldx; p.rows = rows; p.C = C; p.G = G; p.span = span; p.gamma = gamma; p.beta = beta; p.eps = eps;
  p.relu = relu; p.stats = stats; p.y = y; p.ldy = ldy;
  loop<GroupNormP, gn_stats_item>(p, rows / span * G);
  loop<GroupNormP, gn_apply_item>(p, rows * G);
  return 0;
}
int emul_groupnorm_bwd(const float* x, long long ldx, long long rows, int C, int G, int span, const float* gamma,
                       const float* beta, int relu, const float* stats, const float* dy, long long lddy, float* sums,
                       float* dx, long long lddx, float* dgamma, float* dbeta) {
  GroupNormP p{};
  p.x = x; p.ldx = ldx; p.rows = rows; p.C = C; p.G = G; p.span = span; p.gamma = gamma; p.beta = beta; p.relu = relu;
  p.stats = const_cast<float*>(stats); p.dy = dy; p.lddy = lddy; p.sums = sums; p.dx = dx; p.lddx = lddx;
  p.dgamma = dgamma; p.dbeta = dbeta;
  loop<GroupNormP, gn_bwd_sums_item>(p, rows / span * G);
  loop<GroupNormP, gn_bwd_apply_item>(p, rows * G);
  return 0;
}
}
