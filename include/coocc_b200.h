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
} coocc_conv_desc;

/* y[v_out, co] = sum x[...] w[...] (+ bias[co]) (relu).  y is fp32 with row stride ldo.
 * stats (optional, may be NULL): float[2*Cout], must be zeroed by the caller; receives the
 * per-channel sum and sum of squares of the raw conv output (before bias/relu) -- the batch
 * statistics BatchNorm3d needs, produced in the conv epilogue instead of a second pass. */
int coocc_conv3d_fwd(const coocc_conv_desc* d, const void* x, const void* w, float* y, long long ldo,
                     const float* bias, int relu, float* stats, void* stream);

/* dx[v, ci] (fp32, row stride ldo) for a stride-1 convolution; dy has the conv's output extent
 * (= input extent).  Strided convolutions: scatter dy onto the input lattice with
 * coocc_dilate2 first and call this with stride = 1. */
int coocc_conv3d_dgrad(const coocc_conv_desc* d, const void* dy, const void* w, float* dx, long long ldo,
                       void* stream);

/* dw[co][tap][ci] += ... (fp32; caller zero-fills dw; split-K partial sums are added atomically). */
int coocc_conv3d_wgrad(const coocc_conv_desc* d, const void* x, const void* dy, float* dw, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* COOCC_B200_H_ */
