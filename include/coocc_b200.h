/*
 * coocc_b200.h -- C ABI of libcoocc_b200.so, the sm_100a kernels of the Co-Occ fused-voxel hot path.
 *
 * Conventions (they replace the reference's pybind11 op convention, see
 * mmdetection3d/mmdet3d/ops/ball_query/src/ball_query.cpp:32-45 and
 * mmdetection3d/mmdet3d/ops/furthest_point_sample/src/furthest_point_sample.cpp:35-46:
 * `xxx_wrapper(int b, int n, ..., at::Tensor out) -> int`, caller-allocated outputs, work on the
 * current CUDA stream, `exit(-1)` on a launch error):
 *   - plain device pointers + sizes, no torch types; outputs and workspaces are caller-allocated;
 *   - every entry point is stream-ordered on `stream` (a cudaStream_t passed as void*), re-entrant,
 *     keeps no hidden device state, never synchronises the device unless stated, never exits;
 *   - return value: 0 on success, a negative COOCC_ERR_* code otherwise.
 *
 * Layouts: feature grids are NDHWC, i.e. a row-major [V = X*Y*Z, C] matrix whose voxel index is
 * v = (x*Y + y)*Z + z (this is torch's channels_last_3d for the reference's [1,C,X,Y,Z] tensors).
 * Convolution weights are [Cout][kx][ky][kz][Cin] (channels_last_3d of [Cout,Cin,k,k,k]).
 */
#ifndef COOCC_B200_H_
#define COOCC_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define COOCC_ERR_ARG (-1)        /* invalid argument */
#define COOCC_ERR_ALIGN (-2)      /* pointer / stride not 16-byte aligned as required by TMA */
#define COOCC_ERR_DRIVER (-3)     /* CUDA driver entry point missing */
#define COOCC_ERR_TENSORMAP (-4)  /* cuTensorMapEncode* rejected the descriptor */
#define COOCC_ERR_CUDA (-5)       /* kernel launch / runtime error (see cudaGetLastError) */
#define COOCC_ERR_CAPACITY (-6)   /* problem exceeds a kernel's static capacity */

#define COOCC_DTYPE_TF32 0 /* fp32 storage, tf32 tensor-core math (TMA rounds to nearest), fp32 accumulate */
#define COOCC_DTYPE_BF16 1 /* bf16 storage and math, fp32 accumulate */
#define COOCC_DTYPE_TF32X3 2 /* fp32 storage, fp32-accurate: 3-pass hi/lo tf32 split (parity mode) */

int coocc_version(void);

/* ------------------------------------------------------------------------------------------
 * Dense 3D convolution on tcgen05 tensor cores (csrc/conv_tc.cu).
 * Replaces nn.Conv3d (cuDNN) at P/coocc/fuser/bifuser_n.py:23-30, P/coocc/backbones/resnet3d.py:16-31,
 * P/coocc/necks/fpn3d.py:48-67, P/coocc/dense_heads/occ_head.py:102-132 and nn.Linear (cuBLAS) at
 * P/utils/nerf_mlp.py:92-105 / bifuser_n.py:32-36 (a Linear is the ksize=1 case with X=rows, Y=Z=1).
 * padding = ksize/2, dilation 1, batch 1 (the reference asserts B == 1, coocc_ray.py:365).
 * ------------------------------------------------------------------------------------------ */
typedef struct coocc_conv_desc {
  int X, Y, Z;     /* input spatial extent */
  int Cin, Cout;   /* channels */
  int ksize;       /* 1 or 3 */
  int stride;      /* 1 or 2 */
  int dtype;       /* COOCC_DTYPE_* : element type of x / w / dy */
  long long ldx;   /* row stride of the input activation matrix, in elements (>= Cin) */
  long long ldy;   /* row stride of the output-gradient matrix dy, in elements (>= Cout) */
  int out_bf16;    /* 1: y (fwd) / dx (dgrad) are written as bf16 (row stride in bf16 elements, % 8 == 0
                      for vector stores); 0: fp32.  Not available with COOCC_DTYPE_TF32X3. */
} coocc_conv_desc;

/* y[v_out, co] = sum x[...] w[...] (+ bias[co]) (relu).  y is fp32 (or bf16, d->out_bf16) with row stride ldo.
 * stats (optional, may be NULL): float[2*Cout], must be zeroed by the caller; receives the
 * per-channel sum and sum of squares of the raw conv output (before bias/relu) -- the batch
 * statistics BatchNorm3d needs, produced in the conv epilogue instead of a second pass. */
int coocc_conv3d_fwd(const coocc_conv_desc* d, const void* x, const void* w, void* y, long long ldo,
                     const float* bias, int relu, float* stats, void* stream);

/* dx[v, ci] (fp32 or bf16 per d->out_bf16, row stride ldo) over the convolution's INPUT grid d->X/Y/Z; dy has the
 * conv's output extent.  stride 1: one launch.  stride 2: one launch per parity class of the input lattice (8 for
 * 3x3x3, 1 + a zero fill for 1x1x1), each a small convolution of dy on its own grid with the taps of matching parity
 * whose output rows are scattered to the class's voxels -- the forward's FLOPs (zero insertion with coocc_dilate2 + a
 * stride-1 call, the round-1 route, multiplies eight times as many zeros). */
int coocc_conv3d_dgrad(const coocc_conv_desc* d, const void* dy, const void* w, void* dx, long long ldo,
                       void* stream);
/* dx = dgrad(dy, w) + addend: the data gradient plus the gradient arriving through a skip connection of x (a residual
 * branch, a second consumer), added in the convolution's epilogue instead of in a separate pass.  bf16 descriptors with
 * bf16 output and stride 1 only; addend = [V][Cin] bf16 rows with row stride ld_add. */
int coocc_conv3d_dgrad_add(const coocc_conv_desc* d, const void* dy, const void* w, void* dx, long long ldo,
                           const void* addend, long long ld_add, void* stream);
/* SMs the persistent convolution grids may occupy from now on (0 = all): lowered by the step runner while it captures
 * the launches that overlap the FPS side branch. */
int coocc_conv_set_sm_budget(int n);
/* Dynamic tile scheduling for the convolution launches issued from now on (0 = static round-robin, the default): tiles
 * are handed out by an atomic counter, so CTAs that start late because another kernel holds their SM take what is
 * left instead of running a static share as a second wave. */
int coocc_conv_set_dynamic(int mode);   /* 0 static, 1 every launch, 2 only launches with >= 2 x SMs tiles of >= 27 k-blocks */

/* dw[co][tap][ci] += ... (fp32; caller zero-fills dw; split-K partial sums are added atomically). */
int coocc_conv3d_wgrad(const coocc_conv_desc* d, const void* x, const void* dy, float* dw, void* stream);
/* benchmark hook: ky = 0/1 disables/enables the ky-fused 3x3x3 stride-1 path (one y-extended
 * activation tile feeds three filter taps); ky_mt = 128-pixel blocks per CTA tile for
 * Cout <= 128 layers (1, 2, 4).  Negative values keep the current setting. */
int coocc_conv_tune(int ky, int ky_mt);

/* ------------------------------------------------------------------------------------------
 * GSFusion index pipeline (csrc/gsf_index.cu), integer-exact.
 * Replaces P/coocc/fuser/bifuser_n.py:38-125 (fps_NN_fast), :129-135 (nonzero + channels-last
 * gather) and the CUDA ops furthest_point_sample / ball_query it calls
 * (mmdet3d/ops/furthest_point_sample/src/furthest_point_sample.cpp:35-46,
 *  mmdet3d/ops/ball_query/src/ball_query.cpp:32-45).
 * "list" = occupied voxel ids in ascending order (= torch.nonzero row order), "rank" = inverse
 * table voxel -> position in list or -1; counts live on the device.
 * ------------------------------------------------------------------------------------------ */
/* strided [C,X,Y,Z] fp32 (element strides sC,sX,sY,sZ) -> dst[v*ldo + c]; flags[v] = (sum_c != 0) */
int coocc_gsf_pack(const float* src, long long sC, long long sX, long long sY, long long sZ, int C, int X,
                   int Y, int Z, float* dst, long long ldo, unsigned char* flags, void* stream);
long long coocc_gsf_compact_workspace(int V);
int coocc_gsf_compact(const unsigned char* flags, int V, int* list, int* rank, int* count, void* workspace,
                      void* stream);
/* furthest point sampling of `m` points, start index 0, the reference kernel's tie order
 * (max distance, then bit-reversed (k mod 1024), then k).  One or two problems per launch
 * (list1 may be NULL); n_max bounds both counts.  out[j] = position in list. */
int coocc_gsf_fps(const int* list0, const int* count0, int* out0, const int* list1, const int* count1,
                  int* out1, int n_max, int m, int Y, int Z, void* stream);
/* benchmark hook: force the FPS cluster size (0 = automatic) and exchange-variant flags */
int coocc_gsf_fps_tune(int cluster_size, int flags);
/* Ordering of a side-stream FPS launch against a persistent kernel on another stream (index pipelining of the step
 * runner): FPS launches issued while the signal is on report their resident clusters; the gate is a one-thread kernel
 * that returns once `nclusters` have started (and clears the count). */
int coocc_gsf_fps_signal(int on);
int coocc_gsf_fps_gate(int nclusters, void* stream);
/* K nearest keys (d2 <= 176 <=> dist < 13.3) of each representative, order (d2 asc, key asc);
 * out_idx = position in the key list or -1, out_d2 = squared distance or -1. */
int coocc_gsf_rep_topk(const int* rep_idx, int nrep, const int* qlist, const int* key_rank, int X, int Y,
                       int Z, int K, int* out_idx, int* out_d2, void* stream);
/* ball query (first `nsample` queries in index order with d2 < radius^2) + assignment:
 * winner[k*nq_stride + q] = max representative position r with topk_idx[r][k] valid and q in
 * ball(r) (caller initialises winner to -1).  group_out (optional, [nrep][nsample]) receives
 * the ball_query result itself. */
int coocc_gsf_ball_assign(const int* rep_idx, int nrep, const int* qlist, const int* q_rank,
                          const int* topk_idx, int X, int Y, int Z, int K, int radius, int nsample,
                          int nq_stride, int* winner, int* group_out, void* stream);
/* K == 1, N_q <= 2048 branch (bifuser_n.py:55-60): nearest key of every query or -1 */
int coocc_gsf_direct_nn(const int* qlist, const int* qcount, int nq_max, const int* key_rank, int X, int Y,
                        int Z, int* nn, void* stream);
int coocc_gsf_direct_winner(const int* nn, const int* qcount, int* winner, int nq_max, void* stream);
int coocc_iota(int* p, int n, void* stream);

/* ------------------------------------------------------------------------------------------
 * GSFusion feature path (csrc/gsf_feat.cu): bifuser_n.py:138-172 forward and its backward.
 * rows / P / dP / dF are [K][nrep+1][C]; row nrep is the "index -1 -> last key" row (Q3).
 * ------------------------------------------------------------------------------------------ */
int coocc_gsf_gather_rows(const float* grid, long long ld, const int* lookup, const int* lookup_count,
                          const int* topk_idx, int nrep, int K, int C, float* rows, int* err, void* stream);
/* C[i][j] (+)= sum_k A[i*sAi + k*sAk] * B[k*sBk + j*sBj]   (small fp32 SIMT GEMM) */
int coocc_sgemm(int M, int N, int Kd, const float* A, long long sAi, long long sAk, const float* B,
                long long sBk, long long sBj, float* C, long long ldc, int accumulate, void* stream);
int coocc_gsf_modulate_fwd(const float* P, const float* bias, const int* winner, int nq_stride,
                           const int* qlist, const int* qcount, int nq_max, int nrep, int K, int C,
                           const float* own, long long ld_own, float* dst, long long ld_dst, void* stream);
int coocc_gsf_modulate_bwd(const float* P, const float* bias, const int* winner, int nq_stride,
                           const int* qlist, const int* qcount, int nq_max, int nrep, int K, int C,
                           const float* own, long long ld_own, const float* g, long long ld_g, float* d_own,
                           long long ld_down, float* dP, float* dbias, void* stream);
int coocc_gsf_scatter_rows(const float* dF, const int* lookup, const int* lookup_count, const int* topk_idx,
                           int nrep, int K, int C, float* d_grid, long long ld, void* stream);

/* ------------------------------------------------------------------------------------------
 * Volume-render regulariser (csrc/render.cu): coocc_ray.py:358-433.
 * tab[T][4] = (rgb_raw[3], relu(sigma)) per voxel of the render box (the reference's hard-coded
 * 100x100x8 grid clipped to the feature grid), t = (x*by + y)*bz + z.
 * ------------------------------------------------------------------------------------------ */
int coocc_render_box(int X, int Y, int Z, int* bx, int* by, int* bz);
int coocc_render_box_gather(const float* grid, long long ld, int X, int Y, int Z, int C, float* rows,
                            void* stream);
int coocc_render_box_scatter_add(const float* rows, int C, int X, int Y, int Z, float* grid, long long ld,
                                 void* stream);
/* bf16 grids (C and ld multiples of 8): rows gathered in bf16; scatter writes the gradient of the WHOLE grid (the box
 * rows inside the box, zeros elsewhere) in one pass. */
int coocc_render_box_gather_bf16(const void* grid, long long ld, int X, int Y, int Z, int C, void* rows,
                                 void* stream);
int coocc_render_box_scatter_bf16(const void* rows, int C, int X, int Y, int Z, void* grid, long long ld,
                                  void* stream);
/* geom [ncam][D][H][W][3] fp32 ego metres -> rgb_map [ncam][H][W][3], depth_map [ncam][H][W];
 * err[0] = 1 if an in-box sample falls outside the feature grid (the reference raises). */
int coocc_render_composite_fwd(const float* geom, int ncam, int D, int H, int W, const float* tab, int X,
                               int Y, int Z, float* rgb_map, float* depth_map, int* err, void* stream);
int coocc_render_composite_bwd(const float* geom, int ncam, int D, int H, int W, const float* tab, int X,
                               int Y, int Z, const float* g_rgb_map, const float* g_depth_map, float* d_tab,
                               void* stream);
/* x16 bilinear upsample + losses; acc3 = scratch float[3]; losses2 = (loss_depth_render, loss_rgb) */
int coocc_render_upsample_loss_fwd(const float* rgb_map, const float* depth_map, int ncam, int H, int W,
                                   int D, const float* gt_img, const float* gt_depth, float* rgbs,
                                   float* depths, float* acc3, float* losses2, void* stream);
int coocc_render_upsample_loss_bwd(const float* rgbs, const float* depths, int ncam, int H, int W, int D,
                                   const float* gt_img, const float* gt_depth, const float* acc3,
                                   const float* g_losses2, float* g_rgb_map, float* g_depth_map,
                                   void* stream);

/* ------------------------------------------------------------------------------------------
 * Occupancy-head voxel losses (csrc/occ_loss.cu) -- SURVEY §8f rank 1, coarse level.
 * Replaces OccHead.loss_voxel, P/coocc/dense_heads/occ_head.py:267-293: the torch.mode label
 * downsample (:269-280), CE_ssc_loss / sem_scal_loss / geo_scal_loss (P/utils/semkitti.py:139-149,
 * 92-136, 62-89) and lovasz_softmax (P/coocc/dense_heads/lovasz_softmax.py:20-34, 156-203).
 * ------------------------------------------------------------------------------------------ */
/* labels: [X*ratio][Y*ratio][Z*ratio] contiguous, label_bytes = 1 (uint8), 4 (int32) or 8 (int64);
 * out int32[X*Y*Z]: the most frequent label of each ratio^3 cell, where inside a non-empty cell every 0
 * counts as a value of its own (so an empty sub-voxel never out-votes a pair of equal labels and a
 * cell of singletons becomes 255), ties -> smallest value (torch.mode). */
int coocc_occ_label_mode(const void* labels, int label_bytes, int X, int Y, int Z, int ratio, int empty_idx,
                         int* out, void* stream);
/* bytes of the workspace shared by coocc_occ_loss_fwd / _bwd for V voxels and C classes (C <= 32) */
long long coocc_occ_loss_workspace(int V, int C);
/* logits [V][ld] fp32 (ld >= C), labels int32[V] (`ignore` = void label), class_w float[C] or NULL.
 * losses4 (device float[4]) = (CE, sem_scal, geo_scal, lovasz_softmax), unweighted.
 * The workspace keeps what the backward needs (coefficients, Lovasz gradients). */
int coocc_occ_loss_fwd(const float* logits, long long ld, const int* labels, int V, int C, const float* class_w,
                       int ignore, int empty_idx, void* workspace, float* losses4, void* stream);
/* dlogits [V][ldd] = sum_k g_losses4[k] * d losses4[k] / d logits  (g_losses4: device float[4]) */
int coocc_occ_loss_bwd(const float* logits, long long ld, const int* labels, int V, int C, const float* class_w,
                       int ignore, int empty_idx, const void* workspace, const float* g_losses4, float* dlogits,
                       long long ldd, void* stream);

/* ------------------------------------------------------------------------------------------
 * Test-time metric (csrc/eval_hist.cu) -- SURVEY §8f rank 4.
 * Replaces COOCC_Ray.evaluation_semantic + fast_hist, P/coocc/detectors/coocc_ray.py:659-684, 726-730:
 * trilinear up-sampling (align_corners=False) of the logits [X*Y*Z][ld] to the label grid GX x GY x GZ,
 * argmax, and the confusion matrices hist[label][prediction] over the voxels whose label != ignore:
 *   hist_ssc          long long[C*C]  ('SSC', max_label = C)
 *   hist_ssc_visible  long long[C*C]  same, restricted to visible[g] != 0 (NULL when visible is NULL)
 *   hist_sc           long long[4]    ('SC': empty vs. non-empty on both sides)
 * gt: labels of gt_bytes = 1, 4 or 8 bytes.  The three outputs are zeroed by the call.
 * ------------------------------------------------------------------------------------------ */
int coocc_eval_confusion(const float* logits, long long ld, int X, int Y, int Z, int C, const void* gt,
                         int gt_bytes, int GX, int GY, int GZ, const unsigned char* visible, int empty_idx,
                         int ignore, long long* hist_ssc, long long* hist_ssc_visible, long long* hist_sc,
                         void* stream);

/* ------------------------------------------------------------------------------------------
 * Lift-Splat voxel pooling + frustum geometry (csrc/lss_pool.cu) -- SURVEY §8f rank 2, the producer of
 * img_voxel_feats and geom.  Replaces voxel_pooling (P/coocc/image2bev/ViewTransformerLSSVoxel.py:100-123),
 * bev_pool (M/ops/bev_pool/bev_pool.py:80-97, src/bev_pool_cuda.cu:20-98) and get_geometry
 * (P/coocc/image2bev/ViewTransformerLSSBEVDepth.py:117-150).  Batch 1.
 * ------------------------------------------------------------------------------------------ */
/* frustum [D][H][W][3]; cam_mats [ncam][24] = post_trans(3) | inverse(post_rots)(9) | rots @ inverse(intrins)(9) |
 * trans(3), row-major; bda [9]; geom [ncam][D][H][W][3] (ego-frame metres). */
int coocc_lss_geometry(const float* frustum, int ncam, int D, int H, int W, const float* cam_mats,
                       const float* bda, float* geom, void* stream);
long long coocc_lss_workspace(long long npts, long long V);
/* geom [npts][3]; lo3 / dx3: HOST float[3] = (bx - dx/2) and dx of the voxel grid, X,Y,Z = nx.
 * key(point) = voxel id (x*Y + y)*Z + z of trunc((geom - lo) / dx), or X*Y*Z for a dropped point.
 * Sorts (key, point id) stably by key inside `workspace` (coocc_lss_workspace(npts, X*Y*Z) bytes); the four
 * outputs receive device pointers into the workspace: sorted keys, sorted point ids, per-point keys and
 * segments int[V+2] (segments[v] = first sorted position with key >= v). */
int coocc_lss_sort(const float* geom, long long npts, const float* lo3, const float* dx3, int X, int Y, int Z,
                   void* workspace, const unsigned int** sorted_keys, const unsigned int** sorted_vals,
                   const unsigned int** point_keys, const int** segments, void* stream);
/* out[v][0..C) = sum over the points of voxel v of w * feat[row]   (out: [V][ldo], fully written).
 * depth == NULL: row = point id, w = 1 (feat = the reference's flattened volume [npts][ldf]).
 * depth != NULL (fused lift + splat): point id = ((n*D + d)*HW + hw), row = n*HW + hw, w = depth[point id]
 * (feat = image features as [ncam*HW][ldf], depth = depth_prob [ncam][D][HW]). */
int coocc_lss_pool_fwd(const unsigned int* sorted_keys, const unsigned int* sorted_vals, const int* segments,
                       long long npts, int V, int C, const float* feat, long long ldf, const float* depth, int D,
                       int HW, float* out, long long ldo, void* stream);
/* depth == NULL: dfeat [npts][lddf] = gout[key]; else dfeat [ncam*HW][lddf] and ddepth [npts]. */
int coocc_lss_pool_bwd(const unsigned int* point_keys, long long npts, int V, int C, const float* gout,
                       long long ldg, const float* feat, long long ldf, const float* depth, int ncam, int D, int HW,
                       float* dfeat, long long lddf, float* ddepth, void* stream);

/* ------------------------------------------------------------------------------------------
 * OccHead fine / cascade stage (csrc/fine_stage.cu, bodies in csrc/fine_stage.cuh) -- SURVEY §8f rank 1, second half.
 * Replaces, in P/coocc/dense_heads/occ_head.py:182-237, the 5-D F.grid_sample of out_voxel_feats (:219), the
 * camera projection project_points_on_img (P/utils/coordinate_transform.py:29-70), the 4-D F.grid_sample of the
 * image features with the masked camera sum (:231-233) and nn.GroupNorm(16) + ReLU of img_mlp_0 / img_mlp /
 * fine_mlp (:58-78).  All fp32.  Verified on the CPU (tests/emul/fine_emul.cpp) and on a B200 against the reference
 * fixture (tests/test_gpu_fine.py).
 * ------------------------------------------------------------------------------------------ */
/* feats [X*Y*Z][ld] NDHWC rows; coords int32 [3][M] fine voxel indices; S* = final_occ_size; out [M][ldo]:
 * trilinear sample at grid = (c/(S-1) - 0.5)*2, zeros padding, align_corners=False. */
int coocc_fine_sample3d_fwd(const float* feats, long long ld, int X, int Y, int Z, int C, const int* coords, int M,
                            int SX, int SY, int SZ, float* out, long long ldo, void* stream);
/* dfeats [X*Y*Z][ldd] += scatter of gout [M][ldg] (atomics; zeroed by the caller) */
/* same, the feature grid stored in bf16 (samples are still fp32) */
int coocc_fine_sample3d_fwd_bf16(const void* feats, long long ld, int X, int Y, int Z, int C, const int* coords, int M,
                                 int SX, int SY, int SZ, float* out, long long ldo, void* stream);
int coocc_fine_sample3d_bwd(const float* gout, long long ldg, int X, int Y, int Z, int C, const int* coords, int M,
                            int SX, int SY, int SZ, float* dfeats, long long ldd, void* stream);
/* vs3 / lo3: HOST float[3] voxel size and lower corner; inv_bda float[9]; cam27 [ncam][27] = inverse(rots)(9) |
 * trans(3) | intrins(9) | post_rots[:2,:2](4) | post_trans[:2](2); uv [ncam][M][2], mask uint8 [M][ncam]. */
int coocc_fine_project(const int* coords, int M, int ncam, const float* vs3, const float* lo3, const float* inv_bda,
                       const float* cam27, float W_img, float H_img, float* uv, unsigned char* mask, void* stream);
/* img [ncam*H*W][ld] NHWC rows; out [M][ldo] = sum over cameras with mask of the bilinear sample (align_corners=True) */
int coocc_fine_sample2d_fwd(const float* img, long long ld, int ncam, int H, int W, int C, const float* uv,
                            const unsigned char* mask, int M, float* out, long long ldo, void* stream);
int coocc_fine_sample2d_bwd(const float* gout, long long ldg, int ncam, int H, int W, int C, const float* uv,
                            const unsigned char* mask, int M, float* dimg, long long ldd, void* stream);
/* GroupNorm(G) (+ ReLU) over rows [rows][ldx]; `span` consecutive rows form one sample (1: point rows, H*W: an
 * NHWC feature map).  stats float[rows/span][G][2] (mean, rstd) is written by fwd and read by bwd. */
int coocc_groupnorm_fwd(const float* x, long long ldx, long long rows, int C, int G, int span, const float* gamma,
                        const float* beta, float eps, int relu, float* stats, float* y, long long ldy, void* stream);
/* sums: scratch float[rows/span][G][2]; dgamma / dbeta float[C] accumulated with atomics (zeroed by the caller) */
int coocc_groupnorm_bwd(const float* x, long long ldx, long long rows, int C, int G, int span, const float* gamma,
                        const float* beta, int relu, const float* stats, const float* dy, long long lddy, float* sums,
                        float* dx, long long lddx, float* dgamma, float* dbeta, void* stream);

/* Device-side point selection (csrc/fine_select.cu), replacing `argmax != empty` + torch.nonzero + the host
 * torch.randperm subset of coarse_to_fine_coordinates (occ_head.py:183-205, P/utils/coordinate_transform.py:3-21)
 * without a host round trip (CUDA-graph capturable).  logits [X*Y*Z][ld] fp32 coarse prediction; state = device
 * uint64[2] (seed, draw counter; the call increments the counter); coords int32 [3][ratio^3 * topk], child slot
 * o * topk + j of parent slot j; nsel = device int32[2] <- (N occupied, P = min(N, topk)); slots j >= P are padding
 * (coordinates 0).  workspace: coocc_fine_select_workspace(X*Y*Z) bytes. */
long long coocc_fine_select_workspace(int V);
int coocc_fine_select(const float* logits, long long ld, int X, int Y, int Z, int C, int empty_idx, int ratio,
                      int topk, unsigned long long* state, int* coords, int* nsel, void* workspace, void* stream);
/* labels[i] = gt[coords[:, i]] (gt = [GX][GY][GZ] integers of gt_bytes = 1, 4 or 8 bytes: occ_head.py:298), `ignore`
 * for the padding slots (parent slot i % topk >= nsel[1]); nsel == NULL: every slot valid. */
int coocc_fine_gather_labels(const int* coords, long long M, int topk, const int* nsel, const void* gt, int gt_bytes,
                             int GX, int GY, int GZ, int ignore, int* labels, void* stream);

/* ------------------------------------------------------------------------------------------
 * Sparse LiDAR encoder, index side (csrc/sparse_conv.cu) -- SURVEY §8f rank 3.  Replaces, for
 * SparseLiDAREnc8x (P/coocc/voxel_encoder/sparse_lidar_enc.py:125-177), spconv 2.3.6's rulebook construction and
 * gather / scatter (SubMConv3d, SparseConv3d(3, stride 2, padding 1); third-party, not vendored in the reference).
 * A sparse convolution = sp_gather_cols (explicit sparse im2col, [N_out, 27*Cin]) + coocc_conv3d_fwd as a 1x1x1
 * convolution over N_out rows (weight [Cout, kz, ky, kx, Cin] read in place as [Cout, 27*Cin]).
 * coords: int32 [n][4] = (batch, z, y, x), batch 0.  Offset k = (kz*3 + ky)*3 + kx.
 * ------------------------------------------------------------------------------------------ */
/* flags uint8 [oD*oH*oW] (zeroed by the caller) <- 1 at every output site reached by an active input */
int coocc_sp_flag_outputs(const int* coords, int n, int stride, int pad, int oD, int oH, int oW, unsigned char* flags,
                          void* stream);
/* nbr int32 [n_out][27] <- row id (from grid int32 [D*H*W] of the input level, -1 = inactive) of out*stride - pad + k */
int coocc_sp_neighbors(const int* out_coords, int n_out, int stride, int pad, const int* grid, int D, int H, int W,
                       int* nbr, void* stream);
/* cols [n_out][ldc] fp32, cols[o][k*C + c] = feats[nbr[o][k]][c] or 0 */
int coocc_sp_gather_cols(const float* feats, long long ldf, int C, const int* nbr, int n_out, float* cols, long long ldc,
                         void* stream);
/* dfeats [n_in][ldf] += transpose of the gather (zeroed by the caller; atomics) */
int coocc_sp_scatter_cols(const float* dcols, long long ldc, int C, const int* nbr, int n_out, float* dfeats,
                          long long ldf, void* stream);

/* ------------------------------------------------------------------------------------------
 * Small all-reduce over NVLink peer memory (csrc/peer_reduce.cu): the SyncBN statistics exchange of the data-parallel
 * path.  Replaces the per-BatchNorm collectives of torch.nn.SyncBatchNorm, which the reference turns every BatchNorm
 * into (tools/train.py:222-223; all_gather in forward, all_reduce in backward) -- 72 NCCL calls of <= 8 KB per step.
 * A rank's exchange buffer is coocc_peer_buffer_bytes() long, zero-initialised, and mapped into every process (the
 * caller allocates and exchanges it, e.g. torch.distributed._symmetric_memory).  peer_bufs: HOST array of `world`
 * device pointers (entry r = rank r's buffer as seen from this device).  data: n <= slot_floats floats, summed over
 * the ranks in rank order, in place.  slot = position of the call in the step's sequence of calls modulo nslots
 * (>= 2), identical on all ranks; epoch: device unsigned[nslots] of the calling rank, zero-initialised.
 * ------------------------------------------------------------------------------------------ */
long long coocc_peer_buffer_bytes(int world, int nslots, int slot_floats);
int coocc_peer_allreduce(float* data, int n, void* const* peer_bufs, int rank, int world, int slot, int nslots,
                         int slot_floats, unsigned* epoch, void* stream);

/* ------------------------------------------------------------------------------------------
 * Multi-tensor AdamW step that also writes the bf16 operand copy of every parameter (csrc/adamw.cu, body in
 * csrc/adamw.cuh).  Replaces torch.optim.AdamW (the reference's optimizer, coocc_multi_r50_256x704.py:283-290) plus
 * the per-step fp32 -> bf16 weight conversions of the bf16 mode.  Verified against torch.optim.AdamW on the CPU
 * (tests/test_adamw_emul.py) and on a B200 (tests/test_gpu_adamw.py).
 * tensors: device array of ntensors 56-byte entries {float* p; float* g; float* m; float* v; uint16_t* bf16_shadow
 * (or NULL); long long n; float wd_mult; float lr_mult} (the multipliers are mmcv's paramwise_cfg, e.g.
 * norm_decay_mult = 0).  chunk_tensor / chunk_index: device int[nchunks], the tensor id and the chunk number of
 * every chunk of chunk_elems (multiple of 4) elements.  step: device float, incremented by the call (t = 1 for the
 * first update).  zero_grad != 0 clears the gradients after they have been consumed.  dyn: device float[2]
 * {learning-rate multiplier (lr schedule), gradient scale (clip_grad_norm coefficient)} read at run time, or NULL.
 * ------------------------------------------------------------------------------------------ */
int coocc_adamw_step(const void* tensors, int ntensors, const int* chunk_tensor, const int* chunk_index, int nchunks,
                     int chunk_elems, float lr, float beta1, float beta2, float eps, float weight_decay, float* step,
                     int zero_grad, const float* dyn, void* stream);

/* ------------------------------------------------------------------------------------------
 * BatchNorm3d (training mode) + ReLU + residual, and x2 zero insertion (csrc/elementwise.cu).
 * Replaces torch.nn.BatchNorm3d / SyncBatchNorm + nn.ReLU (+ `out += residual`) at
 * P/coocc/fuser/bifuser_n.py:25-29, P/coocc/backbones/resnet3d.py:46-62, P/coocc/necks/fpn3d.py:48-67,
 * P/coocc/dense_heads/occ_head.py:102-132.  Tensors are [V, C] row-major, C % 4 == 0.
 * ------------------------------------------------------------------------------------------ */
/* stats = float[2*C] (sum, sum of squares) from coocc_conv3d_fwd; writes mean_invstd = float[2*C] and
 * updates running_mean / running_var (may be NULL) with `momentum` (unbiased variance). */
int coocc_bn_finalize(const float* stats, int C, long long count, float eps, float momentum,
                      float* running_mean, float* running_var, float* mean_invstd, void* stream);
/* Storage type of the [V, C] activation tensors of the calls below: act_bf16 = 0 -> fp32,
 * 1 -> bf16 (every activation / gradient pointer of the call; statistics, gamma, beta, per-voxel
 * weights and all arithmetic stay fp32).  bf16 mode keeps activations between convolutions in the
 * operand type the next convolution consumes. */
/* out = relu?((x - mean) * invstd * gamma + beta (+ residual)) */
int coocc_bn_act_fwd(const void* x, long long ldx, long long V, int C, const float* mean_invstd,
                     const float* gamma, const float* beta, const void* residual, long long ldr, int relu,
                     void* out, long long ldo, int act_bf16, void* stream);
/* backward of the above, in two stream-ordered halves (a SyncBatchNorm all-reduce of `sums` fits in
 * between).  reduce: sums (float[2*C], zeroed by the caller) += (sum dz, sum dz*xhat), dz = dout*[y>0];
 * afterwards sums[0:C] = dbeta, sums[C:2C] = dgamma.  apply: dx = gradient w.r.t. x with the batch
 * terms divided by `count` (rows the statistics were taken over); dres (optional) = dz.
 * ReLU mask: read from `out` (the forward's output) when it is given -- required when a residual was added --
 * and recomputed from x, gamma, beta with the forward's own expression when out == NULL (one tensor less
 * to read; gamma / beta may be NULL otherwise in the reduce call). */
int coocc_bn_act_bwd_reduce(const void* dout, long long ldd, const void* out, long long ldo, const void* x,
                            long long ldx, long long V, int C, const float* mean_invstd, const float* gamma,
                            const float* beta, int relu, float* sums, int act_bf16, void* stream);
int coocc_bn_act_bwd_apply(const void* dout, long long ldd, const void* out, long long ldo, const void* x,
                           long long ldx, long long V, int C, const float* mean_invstd, const float* gamma,
                           const float* beta, int relu, const float* sums, long long count, void* dx, long long lddx,
                           int act_bf16, void* dres, long long lddr, void* stream);
/* Backward glue of a Linear / convolution with bias and ReLU epilogue (nerf_mlp.py:92-105 and the fine-stage MLPs):
 * out[r][c] = dy[r][c] * [y[r][c] > 0] (y NULL: no mask) in fp32 or bf16, db[c] += sum_r out[r][c] (db NULL: skipped;
 * must be zero-initialised).  dy / y / out are [M][C] rows with their own strides and storage types. */
int coocc_relu_bias_bwd(const void* dy, long long ld_dy, int dy_bf16, const void* y, long long ld_y, int y_bf16,
                        int M, int C, void* out, long long ldo, int out_bf16, float* db, void* stream);
/* dst[(2x,2y,2z)] = src[(x,y,z)], zero elsewhere (dst extent X,Y,Z; src extent oX,oY,oZ) */
int coocc_dilate2(const void* src, long long lds, int oX, int oY, int oZ, int C, void* dst, long long ldd,
                  int X, int Y, int Z, int is_bf16, void* stream);

/* ------------------------------------------------------------------------------------------
 * Trilinear resize (align_corners=False, explicit size) on NDHWC rows, fused with the add /
 * per-voxel weight that follows it (csrc/trilinear.cu).  Replaces F.interpolate at
 * P/coocc/necks/fpn3d.py:91-94 and P/coocc/dense_heads/occ_head.py:161-165.
 *   out[v,:] = base[v,:] (optional) + wts[v*ldw] (optional) * interp(src)[v,:]
 * ------------------------------------------------------------------------------------------ */
int coocc_trilinear_fwd(const void* src, long long lds, int sX, int sY, int sZ, int C, const void* base,
                        long long ldb, const float* wts, long long ldw, void* out, long long ldo, int oX,
                        int oY, int oZ, int act_bf16, void* stream);
/* dsrc = transpose(interp) applied to (wts * dout) */
int coocc_trilinear_bwd(const void* dout, long long ldd, int oX, int oY, int oZ, int C, const float* wts,
                        long long ldw, void* dsrc, long long lds, int sX, int sY, int sZ, int act_bf16,
                        void* stream);
/* benchmark hook: 0 = direct transposed gather in coocc_trilinear_bwd, 1 = three separable 1-D passes (default) */
int coocc_trilinear_tune(int separable);
/* dw[v*lddw] = sum_c dout[v,c] * interp(src)[v,c]   (dw fp32) */
int coocc_trilinear_wgrad(const void* dout, long long ldd, const void* src, long long lds, int sX, int sY,
                          int sZ, int oX, int oY, int oZ, int C, float* dw, long long lddw, int act_bf16,
                          void* stream);

/* Multi-level mix, the whole OccHead level fusion (occ_head.py:161-165) in one pass:
 *   out[v,:] = base[v,:] (optional) + sum_{l<nlev} w_l(v) * interp_l(src_l)[v,:],  w_l(v) = wts[v*ldw + l] or 1
 * src / lds: nlev source matrices and row strides; sdims = {sX,sY,sZ} per level (3*nlev ints); nlev <= 4.
 * The dsrc rows of a same-size level use that level's lds. */
int coocc_trilinear_mix_fwd(int nlev, const void* const* src, const long long* lds, const int* sdims, int C,
                            const void* base, long long ldb, const float* wts, long long ldw, void* out,
                            long long ldo, int oX, int oY, int oZ, int act_bf16, void* stream);
/* one pass over dout: dw[v*lddw + l] = <dout[v,:], interp_l(src_l)[v,:]> (dw may be NULL) and, for the
 * levels with dsrc_same[l] != NULL (their size must equal the output size), dsrc_l[v,:] = w_l(v)*dout[v,:].
 * Coarser levels' gradients: coocc_trilinear_bwd per level. */
int coocc_trilinear_mix_bwd(int nlev, const void* const* src, const long long* lds, const int* sdims,
                            void* const* dsrc_same, int C, const void* dout, long long ldd, const float* wts,
                            long long ldw, float* dw, long long lddw, int oX, int oY, int oZ, int act_bf16,
                            void* stream);

#ifdef __cplusplus
}
#endif
#endif /* COOCC_B200_H_ */
