"""torch.autograd plumbing around the C ABI (include/coocc_b200.h).

Tensors cross the boundary as raw device pointers + sizes; PyTorch only owns the memory, the
stream and the autograd graph.  Activations are "2-D channels-last": a [V, C] fp32 view of the
reference's [1,C,X,Y,Z] tensor (torch.channels_last_3d), row stride a multiple of 4 elements.
"""
import ctypes

import torch

from . import _lib

FPS_NUM = 2048        # P/coocc/fuser/bifuser_n.py:137
BALL_RADIUS = 6
BALL_SAMPLES = 200

DT_TF32, DT_BF16, DT_TF32X3 = 0, 1, 2

# Arithmetic of the tensor-core convolutions / linears:
#   "fp32" : COOCC_DTYPE_TF32X3, 3-pass hi/lo split, fp32-accurate (parity runs)
#   "tf32" : single-pass TF32 (what the reference's cuDNN/cuBLAS do by default on Ampere+)
#   "bf16" : bf16 operands, fp32 accumulation
_PRECISION = {"mode": "tf32"}
_DT = {"fp32": DT_TF32X3, "tf32": DT_TF32, "bf16": DT_BF16}


def set_precision(mode):
    if mode not in _DT:
        raise ValueError("precision must be one of %s" % sorted(_DT))
    _PRECISION["mode"] = mode


def get_precision():
    return _PRECISION["mode"]


# bf16 mode only: activations / gradients between the convolutions are *stored* in bf16 (the conv
# epilogue, the BatchNorm/ReLU/residual kernels and the trilinear kernels read and write bf16), which
# is the operand type the next convolution consumes anyway.  Off: fp32 storage + explicit conversions.
ACT_BF16 = {"enabled": True}


def act_bf16():
    return _PRECISION["mode"] == "bf16" and ACT_BF16["enabled"]


PROFILE = None     # bench.py sets this to a list to time every conv launch with CUDA events
PROFILE_AHEAD_MS = 0   # > 0: whenever a bracket finds the stream idle, a spin kernel of this length is queued first

# Device-side error flags cannot be read while a CUDA graph is being captured: graph.GraphedStep
# sets this to a list, the (flag, exception) pairs land there and GraphedStep.check() reads them
# after a replay.
DEFERRED_ERRORS = None


def _raise_if_set(flag, exc):
    if torch.cuda.is_current_stream_capturing():
        if DEFERRED_ERRORS is None:
            raise RuntimeError("device error flag read during stream capture outside GraphedStep")
        DEFERRED_ERRORS.append((flag, exc))
        return
    if int(flag.item()) != 0:
        raise exc


def _timed(tag, flops, fn, info=""):
    """Run one C-ABI launch; when bench.py has set PROFILE, bracket it with CUDA events.  `flops` is the launch's
    algorithmic work: FLOPs for the tensor-core kernels, BYTES for tags starting with "hbm:"."""
    if PROFILE is None:
        return fn()
    if PROFILE_AHEAD_MS and torch.cuda.current_stream().query():
        # The stream has drained: the host is the bottleneck and the event pair below would time the host's launch
        # overhead (~50 us per call), not the kernel.  Park the GPU for a while so the host gets ahead again.
        torch.cuda._sleep(int(PROFILE_AHEAD_MS * 1.9e6))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = fn()
    e1.record()
    PROFILE.append((e0, e1, flops, tag + info))
    return rc


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _pb(t, byte_off):
    return ctypes.c_void_p(t.data_ptr() + byte_off)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("coocc_b200 has no CPU path: tensors must live on a CUDA device")


# ----------------------------------------------------------------------------------------
# layout helpers
# ----------------------------------------------------------------------------------------
def to_cl2d(x5d):
    """[1,C,X,Y,Z] -> ([V,C] view, (X,Y,Z)).  Zero-copy when x5d is channels_last_3d."""
    B, C, X, Y, Z = x5d.shape
    assert B == 1, "the hot path is batch-1 like the reference (coocc_ray.py:365)"
    x = x5d.permute(0, 2, 3, 4, 1)
    try:
        # also zero-copy for rows with a padded stride (e.g. the 17-class logits, row stride 20)
        return x.view(X * Y * Z, C), (X, Y, Z)
    except RuntimeError:
        return x.contiguous().reshape(X * Y * Z, C), (X, Y, Z)


def to_5d(x2d, dims):
    X, Y, Z = dims
    return x2d.reshape(1, X, Y, Z, x2d.shape[1]).permute(0, 4, 1, 2, 3)


def _rows_ok(t):
    q = 8 if t.dtype == torch.bfloat16 else 4
    return t.dim() == 2 and t.stride(1) == 1 and t.stride(0) % q == 0 and t.data_ptr() % 16 == 0 \
        and t.dtype in (torch.float32, torch.bfloat16)


def _as_rows(t, dtype=None):
    """[R,C] fp32 (or bf16) matrix whose rows are 16-byte aligned (pads the row stride if needed);
    `dtype` forces the storage type."""
    if dtype is None:
        dtype = t.dtype if t.dtype == torch.bfloat16 else torch.float32
    if t.dtype == dtype and _rows_ok(t):
        return t
    R, C = t.shape
    q = 8 if dtype == torch.bfloat16 else 4
    ld = (C + q - 1) // q * q
    if ld == C:
        return t.to(dtype).contiguous()
    buf = torch.zeros(R, ld, device=t.device, dtype=dtype)
    buf[:, :C] = t
    return buf[:, :C]


def _is_bf16(t):
    return 1 if t.dtype == torch.bfloat16 else 0


def weight_rows(w5d):
    """[Cout,Cin,k,k,k] parameter -> [Cout, k^3*Cin] matrix in (kx,ky,kz,Cin) order (zero-copy
    when the parameter is stored channels_last_3d)."""
    w = w5d.permute(0, 2, 3, 4, 1)
    if not w.is_contiguous():
        w = w.contiguous()
    return w.reshape(w5d.shape[0], -1)


def weight_operand(w5d, dtype):
    """weight_rows(w5d) in the storage type of `dtype`.  In bf16 mode a parameter that carries an up-to-date bf16
    shadow (written by optim.FusedAdamW in the same pass as the update) is used as is instead of being converted."""
    if dtype == DT_BF16:
        sh = getattr(w5d, "_coocc_bf16", None)
        if sh is not None and getattr(w5d, "_coocc_bf16_version", None) == w5d._version \
                and getattr(w5d, "_coocc_bf16_layout", None) == (w5d.data_ptr(), tuple(w5d.stride())) \
                and w5d.permute(0, 2, 3, 4, 1).is_contiguous() and (sh.numel() // w5d.shape[0]) % 8 == 0:
            return sh.view(w5d.shape[0], -1)
    return _operand(weight_rows(w5d), dtype)


def out_dim(n, k, s):
    return (n + 2 * (k // 2) - k) // s + 1


# ----------------------------------------------------------------------------------------
# convolution / linear on tcgen05
# ----------------------------------------------------------------------------------------
def _conv_desc(dims, cin, cout, k, s, ldx, ldy, dtype=None):
    X, Y, Z = dims
    if dtype is None:
        dtype = _DT[_PRECISION["mode"]]
    return _lib.ConvDesc(X, Y, Z, cin, cout, k, s, dtype, ldx, ldy)


def _operand(t, dtype):
    """operand in the storage type of `dtype` (bf16 mode converts; fp32 modes pass through)."""
    if dtype != DT_BF16 or t.dtype == torch.bfloat16:
        return t
    R, C = t.shape
    ld = (C + 7) // 8 * 8
    if ld == C:
        return t.to(torch.bfloat16)
    buf = torch.zeros(R, ld, device=t.device, dtype=torch.bfloat16)
    buf[:, :C] = t
    return buf[:, :C]


def conv_fwd_raw(x2d, w2d, dims, cin, cout, k, s, bias=None, relu=False, stats=None, out=None, dtype=None,
                 out_bf16=False):
    L = _lib.lib()
    if dtype is None:
        dtype = _DT[_PRECISION["mode"]]
    odims = tuple(out_dim(n, k, s) for n in dims)
    vo = odims[0] * odims[1] * odims[2]
    if out is None:
        if out_bf16:
            out = torch.empty(vo, (cout + 7) // 8 * 8, device=x2d.device, dtype=torch.bfloat16)
        else:
            out = torch.empty(vo, (cout + 3) // 4 * 4, device=x2d.device, dtype=torch.float32)
    x2d, w2d = _operand(x2d, dtype), _operand(w2d, dtype)      # no-ops when already converted
    d = _conv_desc(dims, cin, cout, k, s, x2d.stride(0), (cout + 7) // 8 * 8, dtype)
    d.out_bf16 = _is_bf16(out)
    flops = 2.0 * vo * (k ** 3) * cin * cout
    rc = _timed("fwd", flops, lambda: L.coocc_conv3d_fwd(ctypes.byref(d), _p(x2d), _p(w2d), _p(out), out.stride(0),
                                                         _p(bias), 1 if relu else 0, _p(stats), _stream()),
                " %s %d->%d k%d s%d" % (dims, cin, cout, k, s))
    _lib.check(rc, "conv3d_fwd")
    return out[:, :cout], odims


class _Conv3dFn(torch.autograd.Function):
    """y[Vout,Cout] = conv3d(x[V,Cin], w[Cout,Cin,k,k,k]) (+bias)(relu), padding k//2."""

    @staticmethod
    def forward(ctx, x2d, w5d, bias, dims, k, s, relu, want_stats=False, out_bf16=False, skip=False):
        _require_cuda(x2d, w5d)
        ctx.set_materialize_grads(False)
        ctx.flags = (bool(want_stats), bool(skip))
        global PRE_TAIL_MARK
        ctx.pre_tail, PRE_TAIL_MARK = PRE_TAIL_MARK, False
        cout, cin = w5d.shape[0], w5d.shape[1]
        dtype = _DT[_PRECISION["mode"]]
        out_bf16 = bool(out_bf16) and dtype == DT_BF16
        # operands in their storage type (bf16 mode: converted once here and kept for backward)
        xo = _operand(_as_rows(x2d), dtype)
        wo = weight_operand(w5d, dtype)
        stats = zeros_small((2, cout), x2d.device) if want_stats else None
        y, odims = conv_fwd_raw(xo, wo, dims, cin, cout, k, s, bias, relu, stats=stats, dtype=dtype,
                                out_bf16=out_bf16)
        ctx.save_for_backward(xo, wo, y if relu else None)
        ctx.meta = (dims, odims, k, s, relu, bias is not None, dtype, cin, cout)
        ctx.wparam = w5d if getattr(w5d, "_coocc_grad", None) is not None else None
        ctx.x_bf16 = x2d.dtype == torch.bfloat16
        outs = [y]
        if want_stats:
            ctx.mark_non_differentiable(stats)
            outs.append(stats)
        if skip:
            # x handed on as an OUTPUT of this node: a second consumer of x (residual branch, another head) that takes
            # this alias sends its gradient into backward() below, where the data-gradient kernel adds it in its
            # epilogue -- instead of autograd summing two [V, C] tensors in a separate pass
            outs.append(x2d.view_as(x2d))
        return outs[0] if len(outs) == 1 else tuple(outs)

    @staticmethod
    def backward(ctx, dy, *more):
        xo, wo, y = ctx.saved_tensors
        dims, odims, k, s, relu, has_bias, dtype, cin, cout = ctx.meta
        want_stats, skip = ctx.flags
        dskip = more[1 if want_stats else 0] if skip else None
        L = _lib.lib()
        if dy is None:                                      # only the alias was used downstream
            return dskip, None, None, None, None, None, None, None, None, None
        dy = _as_rows(dy)
        want_db = has_bias and ctx.needs_input_grad[2]
        db = None
        if (relu or want_db) and dy.dtype in (torch.float32, torch.bfloat16) and dy.stride(1) == 1 and cout <= 1024:
            # ReLU mask, bias gradient and the conversion to the operand type in one pass (csrc/elementwise.cu)
            want_bf16 = dtype == DT_BF16
            if relu or (want_bf16 and dy.dtype != torch.bfloat16) or (not want_bf16 and dy.dtype != torch.float32):
                ld = (cout + 7) // 8 * 8
                mk = torch.empty if ld == cout else torch.zeros
                dyo_full = mk(dy.shape[0], ld, device=dy.device, dtype=torch.bfloat16 if want_bf16 else torch.float32)
                dyo = dyo_full[:, :cout]
            else:
                dyo_full, dyo = None, dy                   # already in operand form: only the bias sums are needed
            if want_db:
                db = zeros_small((cout,), dy.device)
            ym = _as_rows(y) if relu else None
            _lib.check(L.coocc_relu_bias_bwd(_p(dy), dy.stride(0), 1 if dy.dtype == torch.bfloat16 else 0,
                                             _p(ym), ym.stride(0) if relu else 0,
                                             1 if (relu and ym.dtype == torch.bfloat16) else 0, dy.shape[0], cout,
                                             _p(dyo_full), dyo_full.stride(0) if dyo_full is not None else 0,
                                             1 if want_bf16 else 0, _p(db), _stream()), "relu_bias_bwd")
        else:
            if relu:
                dy = dy * (y > 0)
            db = dy.float().sum(0) if want_db else None
            dyo = _operand(dy, dtype)
        dx = dw = None
        if ctx.needs_input_grad[1]:
            # the kernel accumulates (split-K red.add).  With a gradient arena (ddp.GradArena) it adds straight into
            # the parameter's slice of the arena -- no zero-filled scratch, no AccumulateGrad pass -- and the
            # gradient is reported as None to autograd; otherwise into a fresh zero buffer that is returned.
            sink = ctx.wparam._coocc_grad if ctx.wparam is not None else None
            if sink is not None:
                dw2d = sink.permute(0, 2, 3, 4, 1).reshape(cout, k ** 3 * cin)
                assert dw2d.data_ptr() == sink.data_ptr()
            else:
                dw2d = torch.zeros(cout, k ** 3 * cin, device=dy.device, dtype=torch.float32)
            d = _conv_desc(dims, cin, cout, k, s, xo.stride(0), dyo.stride(0), dtype)
            flops = 2.0 * odims[0] * odims[1] * odims[2] * (k ** 3) * cin * cout
            _lib.check(_timed("wgrad", flops, lambda: L.coocc_conv3d_wgrad(ctypes.byref(d), _p(xo), _p(dyo), _p(dw2d),
                                                                          _stream()),
                              " %s %d->%d k%d s%d" % (dims, cin, cout, k, s)), "conv3d_wgrad")
            if sink is not None:
                cb = getattr(ctx.wparam, "_coocc_on_grad", None)      # ddp.GradReducer: this gradient is final
                if cb is not None:
                    cb(ctx.wparam)
            else:
                # gradient in the parameter's own (channels_last_3d) layout: a view, no copy
                dw = dw2d.reshape(cout, k, k, k, cin).permute(0, 4, 1, 2, 3)
        if ctx.pre_tail and PRE_TAIL_HOOK is not None:
            PRE_TAIL_HOOK("dgrad")  # graph.GraphedStep: fork the next step's index branch before the LAST convolution
        if ctx.needs_input_grad[0]:
            # the data gradient is written in the storage type of x (bf16 activations stay bf16)
            if ctx.x_bf16 and dtype == DT_BF16:
                ldo = (cin + 7) // 8 * 8
                dxb = torch.empty(dims[0] * dims[1] * dims[2], ldo, device=dy.device, dtype=torch.bfloat16)
            else:
                ldo = (cin + 3) // 4 * 4
                dxb = torch.empty(dims[0] * dims[1] * dims[2], ldo, device=dy.device, dtype=torch.float32)
            # stride 2: the kernel runs one launch per parity class of the input lattice (csrc/conv_tc.cu dgrad_impl,
            # `cls`) -- the forward's FLOPs; round 1 zero-inserted dy (coocc_dilate2) and paid eight times as many
            d = _conv_desc(dims, cin, cout, k, s, ldo, dyo.stride(0), dtype)
            d.out_bf16 = _is_bf16(dxb)
            flops = 2.0 * odims[0] * odims[1] * odims[2] * (k ** 3) * cin * cout
            fuse = (dskip is not None and d.out_bf16 and s == 1 and dtype == DT_BF16 and dskip.dtype == torch.bfloat16
                    and dskip.dim() == 2 and dskip.stride(1) == 1 and dskip.stride(0) % 8 == 0)
            if fuse:
                _lib.check(_timed("dgrad", flops, lambda: L.coocc_conv3d_dgrad_add(
                    ctypes.byref(d), _p(dyo), _p(wo), _p(dxb), dxb.stride(0), _p(dskip), dskip.stride(0), _stream()),
                                  " %s %d->%d k%d s%d" % (dims, cin, cout, k, s)), "conv3d_dgrad_add")
                dskip = None
            else:
                _lib.check(_timed("dgrad", flops, lambda: L.coocc_conv3d_dgrad(ctypes.byref(d), _p(dyo), _p(wo), _p(dxb),
                                                                              dxb.stride(0), _stream()),
                                  " %s %d->%d k%d s%d" % (dims, cin, cout, k, s)), "conv3d_dgrad")
            dx = dxb[:, :cin]
            if ctx.x_bf16 and dx.dtype != torch.bfloat16:
                dx = dx.to(torch.bfloat16)
        if dskip is not None:
            dx = dskip if dx is None else dx + dskip.to(dx.dtype)
        return dx, dw, db, None, None, None, None, None, None, None


def conv3d(x2d, w5d, dims, k, s=1, bias=None, relu=False, want_stats=False, out_bf16=False, skip=False):
    """skip=True appends an alias of x2d to the outputs: route x's OTHER consumer through it and its gradient is added
    in this convolution's data-gradient epilogue (see _Conv3dFn.forward)."""
    return _Conv3dFn.apply(x2d, w5d, bias, dims, k, s, relu, want_stats, out_bf16, skip)


def _sync_group():
    """process group for SyncBatchNorm statistics, or None (single process / disabled)."""
    import torch.distributed as dist
    if SYNC_BN["enabled"] and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist
    return None


# The reference converts every BatchNorm of the model to SyncBatchNorm when training distributed
# (tools/train.py:222-223, sync_bn=True in the configs): batch statistics and their gradients are
# reduced over all ranks.  Here that is one small all-reduce of the [2,C] epilogue sums in forward
# and one of the [2,C] backward sums.  Single process: plain BatchNorm3d.
SYNC_BN = {"enabled": True}

# peer.PeerExchange or None: when set (bench.py / the training harness, N > 1 on NVLink), the [2,C] statistics go
# through the peer-memory all-reduce kernel (csrc/peer_reduce.cu) instead of an NCCL call each
PEER = None


def begin_step():
    """start of a training step (eager or captured): resets the per-step call counter of the peer exchange"""
    if PEER is not None:
        PEER.begin_step()


# One pre-zeroed pool per step for the many small accumulators the kernels add into (BatchNorm statistics [2,C] of every
# conv epilogue, the [2,C] backward sums): ~80 torch.zeros() fill launches per step become one.  The step runner
# (graph.GraphedStep, bench.py) brackets a step with zero_pool_begin() -- a device memset, so it must sit INSIDE a
# captured graph -- and zero_pool_end(); outside such a bracket zeros_small() is plain torch.zeros().
_ZERO_POOL = {"buf": None, "cursor": 0, "armed": False}
ZERO_POOL_FLOATS = 1 << 19          # 2 MB


def zero_pool_begin(device):
    zp = _ZERO_POOL
    if zp["buf"] is None or zp["buf"].device != torch.device(device):
        zp["buf"] = torch.zeros(ZERO_POOL_FLOATS, device=device, dtype=torch.float32)
    else:
        zp["buf"].zero_()
    zp["cursor"], zp["armed"] = 0, True
    zp["step"] = zp.get("step", 0) + 1
    _PENDING_COUNTERS.clear()


def zero_pool_end():
    _ZERO_POOL["armed"] = False
    if _PENDING_COUNTERS:
        torch._foreach_add_(list(_PENDING_COUNTERS), 1)      # one launch for all num_batches_tracked counters
        _PENDING_COUNTERS.clear()


_PENDING_COUNTERS = []


def count_batch(counter):
    """num_batches_tracked += 1 (nn.BatchNorm semantics).  Inside a bracketed step the 36 one-element increments are
    deferred to zero_pool_end() and issued as one multi-tensor launch."""
    if _ZERO_POOL["armed"] and counter.is_cuda:
        _PENDING_COUNTERS.append(counter)
    else:
        counter.add_(1)


def _bn_grad_sink(gamma, beta, C):
    """[2, C] view of the gradient arena covering (d beta, d gamma) when ddp.GradArena laid the two gradients out back
    to back, the step is bracketed (arena cleared before the step, one backward per step) and this BatchNorm has not
    been used yet in this step; else None.  The backward-reduce kernel then accumulates its sums straight into the
    gradients: no zero-filled scratch, no two AccumulateGrad launches per BatchNorm."""
    zp = _ZERO_POOL
    if not zp["armed"]:
        return None
    gs, bs = getattr(gamma, "_coocc_grad", None), getattr(beta, "_coocc_grad", None)
    if gs is None or bs is None or gamma.grad is not gs or beta.grad is not bs:
        return None
    if not (gs.is_contiguous() and bs.is_contiguous() and gs.numel() == C and bs.numel() == C
            and gs.data_ptr() == bs.data_ptr() + 4 * C and bs.data_ptr() % 16 == 0):
        return None
    if getattr(gamma, "_coocc_sink_step", None) == zp["step"]:
        return None                       # second use in one step: the sums would mix
    gamma._coocc_sink_step = zp["step"]
    return bs.as_strided((2, C), (C, 1))


def zeros_small(shape, device):
    """zero-initialised fp32 tensor for a kernel to accumulate into (16-byte aligned)"""
    zp = _ZERO_POOL
    n = 1
    for v in shape:
        n *= int(v)
    n4 = (n + 3) // 4 * 4
    if zp["armed"] and zp["buf"].device == torch.device(device) and zp["cursor"] + n4 <= ZERO_POOL_FLOATS:
        t = zp["buf"][zp["cursor"]:zp["cursor"] + n].view(*shape)
        zp["cursor"] += n4
        return t
    return torch.zeros(*shape, device=device, dtype=torch.float32)


def _stats_all_reduce(dist, t):
    if PEER is not None and PEER.fits(t):
        PEER.all_reduce(t)
    else:
        dist.all_reduce(t)


class _BNActFn(torch.autograd.Function):
    """out = relu?(batchnorm_train(x) (+ residual)) from the conv-epilogue statistics."""

    @staticmethod
    def forward(ctx, x, stats, gamma, beta, residual, relu, eps, momentum, running_mean, running_var,
                var_count=False):
        L = _lib.lib()
        V, C = x.shape
        dev = x.device
        count = V
        dist = _sync_group()
        if dist is not None:
            _stats_all_reduce(dist, stats)
            count = V * dist.get_world_size()
            if var_count:       # rows differ per rank (sparse LiDAR features): the statistics span sum_r V_r rows
                cnt = torch.tensor([float(V)], device=dev)
                dist.all_reduce(cnt)
                count = int(cnt.item())
        mi = torch.empty(2, C, device=dev, dtype=torch.float32)
        _lib.check(L.coocc_bn_finalize(_p(stats), C, count, float(eps), float(momentum), _p(running_mean),
                                       _p(running_var), _p(mi), _stream()), "bn_finalize")
        x = _as_rows(x)
        out = torch.empty(V, C, device=dev, dtype=x.dtype)
        res_dtype = None
        if residual is not None:
            res_dtype = residual.dtype
            residual = _as_rows(residual, x.dtype)
        e = x.element_size()
        _lib.check(_timed("hbm:bn_act_fwd", float(e * V * C * (3 if residual is not None else 2)),
                          lambda: L.coocc_bn_act_fwd(_p(x), x.stride(0), V, C, _p(mi), _p(gamma), _p(beta), _p(residual),
                                                     residual.stride(0) if residual is not None else 0,
                                                     1 if relu else 0, _p(out), out.stride(0), _is_bf16(x), _stream())),
                   "bn_act_fwd")
        # the backward needs the ReLU mask: with a residual it reads it from `out`, otherwise it recomputes
        # y > 0 from x, gamma, beta (the forward's own expression) and `out` is not read again
        ctx.save_for_backward(x, out if (relu and residual is not None) else None, mi, gamma, beta)
        ctx.meta = (relu, residual is not None, count, res_dtype)
        ctx.bn_params = (gamma, beta) if getattr(gamma, "_coocc_grad", None) is not None else (None, None)
        return out

    @staticmethod
    def backward(ctx, dout):
        L = _lib.lib()
        x, out, mi, gamma, beta = ctx.saved_tensors
        relu, has_res, count, res_dtype = ctx.meta
        V, C = x.shape
        dout = _as_rows(dout, x.dtype)
        bf = _is_bf16(x)
        gparam, bparam = ctx.bn_params
        sink = _bn_grad_sink(gparam, bparam, C) if (gparam is not None and ctx.needs_input_grad[2]
                                                    and ctx.needs_input_grad[3]) else None
        sums = sink if sink is not None else zeros_small((2, C), x.device)
        ldo = out.stride(0) if out is not None else 0
        e = x.element_size()
        _lib.check(_timed("hbm:bn_act_bwd_reduce", float(e * V * C * (3 if out is not None else 2)),
                          lambda: L.coocc_bn_act_bwd_reduce(_p(dout), dout.stride(0), _p(out), ldo, _p(x), x.stride(0), V,
                                                            C, _p(mi), _p(gamma), _p(beta), 1 if relu else 0, _p(sums),
                                                            bf, _stream())), "bn_act_bwd_reduce")
        local = sums
        if count != V:                      # SyncBN: batch terms use the sums over all ranks
            dist = _sync_group()
            if sink is not None:
                sums = sums.clone()         # the arena keeps the per-rank sums (DDP averages them)
            else:
                local = sums.clone()        # dgamma / dbeta stay per-rank (DDP averages them)
            _stats_all_reduce(dist, sums)
        dx = torch.empty(V, C, device=x.device, dtype=x.dtype)
        dres = torch.empty(V, C, device=x.device, dtype=x.dtype) if has_res else None
        _lib.check(_timed("hbm:bn_act_bwd_apply", float(e * V * C * (5 if has_res else 3)),
                          lambda: L.coocc_bn_act_bwd_apply(_p(dout), dout.stride(0), _p(out), ldo, _p(x), x.stride(0), V, C,
                                                           _p(mi), _p(gamma), _p(beta), 1 if relu else 0, _p(sums), count,
                                                           _p(dx), dx.stride(0), bf, _p(dres), C if has_res else 0,
                                                           _stream())), "bn_act_bwd_apply")
        if has_res and dres.dtype != res_dtype:
            dres = dres.to(res_dtype)
        if sink is not None:
            for prm in (bparam, gparam):
                cb = getattr(prm, "_coocc_on_grad", None)       # ddp.GradReducer: this gradient is final
                if cb is not None:
                    cb(prm)
            return dx, None, None, None, dres, None, None, None, None, None, None
        return dx, None, local[1], local[0], dres, None, None, None, None, None, None


def bn_act_eval(x, running_mean, running_var, gamma, beta, eps, residual=None, relu=True):
    """Inference-mode BatchNorm (+ residual, ReLU) from the running statistics; no autograd."""
    L = _lib.lib()
    x = _as_rows(x)
    V, C = x.shape
    mi = torch.stack([running_mean.float(), torch.rsqrt(running_var.float() + eps)]).contiguous()
    out = torch.empty(V, C, device=x.device, dtype=x.dtype)
    if residual is not None:
        residual = _as_rows(residual, x.dtype)
    _lib.check(L.coocc_bn_act_fwd(_p(x), x.stride(0), V, C, _p(mi), _p(gamma), _p(beta), _p(residual),
                                  residual.stride(0) if residual is not None else 0, 1 if relu else 0,
                                  _p(out), out.stride(0), _is_bf16(x), _stream()), "bn_act_fwd")
    return out


def bn_act(x, stats, gamma, beta, residual=None, relu=True, eps=1e-5, momentum=0.1, running_mean=None,
           running_var=None, var_count=False):
    return _BNActFn.apply(x, stats, gamma, beta, residual, relu, eps, momentum, running_mean, running_var, var_count)


def linear(x2d, weight, bias=None, relu=False, out_bf16=False):
    """nn.Linear on the tensor cores: a 1x1x1 convolution over `rows` voxels.  out_bf16: in bf16 mode the result is
    written in bf16 (for hidden layers whose only consumer is the next Linear, which rounds to bf16 anyway)."""
    w5d = weight.reshape(weight.shape[0], weight.shape[1], 1, 1, 1)
    return _Conv3dFn.apply(x2d, w5d, bias, (x2d.shape[0], 1, 1), 1, 1, relu, False, out_bf16)


class _ResizeAddFn(torch.autograd.Function):
    """out = base (opt.) + wts[:, col] (opt.) * trilinear_resize(src)   on [V,C] rows."""

    @staticmethod
    def forward(ctx, src, sdims, odims, base, wts, col):
        L = _lib.lib()
        src = _as_rows(src)
        C = src.shape[1]
        Vo = odims[0] * odims[1] * odims[2]
        out = torch.empty(Vo, C, device=src.device, dtype=src.dtype)
        base_dtype = None
        if base is not None:
            base_dtype = base.dtype
            base = _as_rows(base, src.dtype)
        wptr, ldw = None, 0
        if wts is not None:
            wts = wts.contiguous()
            wptr, ldw = _pb(wts, 4 * col), wts.stride(0)
        _lib.check(L.coocc_trilinear_fwd(_p(src), src.stride(0), sdims[0], sdims[1], sdims[2], C, _p(base),
                                         base.stride(0) if base is not None else 0, wptr, ldw, _p(out), C,
                                         odims[0], odims[1], odims[2], _is_bf16(src), _stream()), "trilinear_fwd")
        ctx.save_for_backward(src, wts)
        ctx.meta = (sdims, odims, col, base_dtype)
        return out

    @staticmethod
    def backward(ctx, dout):
        L = _lib.lib()
        src, wts = ctx.saved_tensors
        sdims, odims, col, base_dtype = ctx.meta
        dout = _as_rows(dout, src.dtype)
        bf = _is_bf16(src)
        C = src.shape[1]
        dsrc = dwts = None
        wptr, ldw = (None, 0) if wts is None else (_pb(wts, 4 * col), wts.stride(0))
        if ctx.needs_input_grad[0]:
            dsrc = torch.empty(src.shape[0], C, device=src.device, dtype=src.dtype)
            _lib.check(L.coocc_trilinear_bwd(_p(dout), dout.stride(0), odims[0], odims[1], odims[2], C, wptr, ldw,
                                             _p(dsrc), C, sdims[0], sdims[1], sdims[2], bf, _stream()), "trilinear_bwd")
        if wts is not None and ctx.needs_input_grad[4]:
            dwts = torch.zeros_like(wts)
            _lib.check(L.coocc_trilinear_wgrad(_p(dout), dout.stride(0), _p(src), src.stride(0), sdims[0], sdims[1],
                                               sdims[2], odims[0], odims[1], odims[2], C, _pb(dwts, 4 * col),
                                               dwts.stride(0), bf, _stream()), "trilinear_wgrad")
        dbase = None
        if base_dtype is not None:
            dbase = dout if dout.dtype == base_dtype else dout.to(base_dtype)
        return dsrc, None, None, dbase, dwts, None


def resize_add(src, sdims, odims, base=None, wts=None, col=0):
    return _ResizeAddFn.apply(src, tuple(sdims), tuple(odims), base, wts, col)


class _ResizeMixFn(torch.autograd.Function):
    """out = base (opt.) + sum_l wts[:, l] (opt.) * trilinear_resize(src_l)  -- all levels in one pass."""

    @staticmethod
    def forward(ctx, sdims_list, odims, base, wts, *srcs):
        L = _lib.lib()
        n = len(srcs)
        dt = torch.bfloat16 if srcs[0].dtype == torch.bfloat16 else torch.float32
        srcs = [_as_rows(s_, dt) for s_ in srcs]
        C = srcs[0].shape[1]
        Vo = odims[0] * odims[1] * odims[2]
        out = torch.empty(Vo, C, device=srcs[0].device, dtype=dt)
        base_dtype = None
        if base is not None:
            base_dtype = base.dtype
            base = _as_rows(base, dt)
        wptr, ldw = None, 0
        if wts is not None:
            wts = wts.contiguous()
            assert wts.dtype == torch.float32 and wts.shape == (Vo, n)
            wptr, ldw = _p(wts), wts.stride(0)
        ptrs = (ctypes.c_void_p * n)(*[s_.data_ptr() for s_ in srcs])
        lds = (ctypes.c_longlong * n)(*[s_.stride(0) for s_ in srcs])
        sd = (ctypes.c_int * (3 * n))(*[int(v) for d_ in sdims_list for v in d_])
        _lib.check(L.coocc_trilinear_mix_fwd(n, ptrs, lds, sd, C, _p(base), base.stride(0) if base is not None else 0,
                                             wptr, ldw, _p(out), out.stride(0), odims[0], odims[1], odims[2],
                                             1 if dt == torch.bfloat16 else 0, _stream()), "trilinear_mix_fwd")
        ctx.save_for_backward(wts, *srcs)
        ctx.meta = (tuple(tuple(d_) for d_ in sdims_list), tuple(odims), base_dtype)
        return out

    @staticmethod
    def backward(ctx, dout):
        L = _lib.lib()
        wts, *srcs = ctx.saved_tensors
        sdims_list, odims, base_dtype = ctx.meta
        n = len(srcs)
        dt = srcs[0].dtype
        bf = 1 if dt == torch.bfloat16 else 0
        dout = _as_rows(dout, dt)
        C = srcs[0].shape[1]
        Vo = odims[0] * odims[1] * odims[2]
        need_dw = wts is not None and ctx.needs_input_grad[3]
        dsrcs = [None] * n
        same_ptrs = [None] * n
        for l in range(n):
            if ctx.needs_input_grad[4 + l] and sdims_list[l] == odims:
                dsrcs[l] = torch.empty(Vo, srcs[l].stride(0), device=dout.device, dtype=dt)[:, :C]
                same_ptrs[l] = dsrcs[l].data_ptr()
        dw = torch.empty(Vo, n, device=dout.device, dtype=torch.float32) if need_dw else None
        if need_dw or any(p_ is not None for p_ in same_ptrs):
            ptrs = (ctypes.c_void_p * n)(*[s_.data_ptr() for s_ in srcs])
            lds = (ctypes.c_longlong * n)(*[s_.stride(0) for s_ in srcs])
            sd = (ctypes.c_int * (3 * n))(*[int(v) for d_ in sdims_list for v in d_])
            dp = (ctypes.c_void_p * n)(*same_ptrs)
            _lib.check(L.coocc_trilinear_mix_bwd(n, ptrs, lds, sd, dp, C, _p(dout), dout.stride(0), _p(wts),
                                                 wts.stride(0) if wts is not None else 0, _p(dw), n if need_dw else 0,
                                                 odims[0], odims[1], odims[2], bf, _stream()), "trilinear_mix_bwd")
        for l in range(n):
            if ctx.needs_input_grad[4 + l] and dsrcs[l] is None:
                sdm = sdims_list[l]
                dsrcs[l] = torch.empty(srcs[l].shape[0], C, device=dout.device, dtype=dt)
                wptr, ldw = (None, 0) if wts is None else (_pb(wts, 4 * l), wts.stride(0))
                _lib.check(L.coocc_trilinear_bwd(_p(dout), dout.stride(0), odims[0], odims[1], odims[2], C, wptr, ldw,
                                                 _p(dsrcs[l]), C, sdm[0], sdm[1], sdm[2], bf, _stream()), "trilinear_bwd")
        dbase = None
        if base_dtype is not None:
            dbase = dout if dout.dtype == base_dtype else dout.to(base_dtype)
        return (None, None, dbase, dw, *dsrcs)


def resize_mix(srcs, sdims_list, odims, base=None, wts=None):
    """base + sum_l wts[:, l] * resize(srcs[l]) on [V,C] rows (one fused pass, see csrc/trilinear.cu)."""
    return _ResizeMixFn.apply([tuple(d) for d in sdims_list], tuple(odims), base, wts, *srcs)


# ----------------------------------------------------------------------------------------
# GSFusion
# ----------------------------------------------------------------------------------------
class ReferenceQuirk(RuntimeError):
    pass


def _gsf_direction(L, st, qlist, qrank, qcnt, nq, klist, krank, dims, K, dev, want_group=False):
    """Index pipeline for one direction: returns dict(rep_idx, nrep, topk_idx, winner[K,nq])."""
    X, Y, Z = dims
    V = X * Y * Z
    out = {}
    if nq <= FPS_NUM:
        if K != 1:
            # bifuser_n.py:88-93 applies a 2-D mask to a 1-D row
            raise IndexError("too many indices for tensor of dimension 1")
        nn = torch.empty(max(nq, 1), device=dev, dtype=torch.int32)
        _lib.check(L.coocc_gsf_direct_nn(_p(qlist), _p(qcnt), max(nq, 1), _p(krank), X, Y, Z, _p(nn), st), "direct_nn")
        winner = torch.empty(1, V, device=dev, dtype=torch.int32)
        _lib.check(L.coocc_gsf_direct_winner(_p(nn), _p(qcnt), _p(winner), V, st), "direct_winner")
        rep_idx = torch.empty(max(nq, 1), device=dev, dtype=torch.int32)
        _lib.check(L.coocc_iota(_p(rep_idx), max(nq, 1), st), "iota")
        out.update(rep_idx=rep_idx, nrep=nq, topk_idx=nn.reshape(-1, 1), winner=winner, topk_d2=None, group=None)
        return out
    return None


def gsf_prologue(img5d, pts5d, out=None):
    """pack + compact of both grids (bifuser_n.py:129-135), no host synchronisation.
    Returns dict(cat [V,4C] with the img/pts slices filled and the fused slices zero, flags, lists,
    ranks, counts [2] int32 on the device, ws).  `out` = a previous result whose buffers are reused
    (graph.GraphedStep keeps them at fixed addresses)."""
    L = _lib.lib()
    st = _stream()
    _require_cuda(img5d, pts5d)
    B, C, X, Y, Z = img5d.shape
    assert B == 1 and pts5d.shape == img5d.shape
    assert img5d.dtype == torch.float32 and pts5d.dtype == torch.float32
    dev = img5d.device
    V = X * Y * Z
    if out is None:
        out = dict(cat=torch.zeros(V, 4 * C, device=dev, dtype=torch.float32),
                   flags=torch.empty(2, V, device=dev, dtype=torch.uint8),
                   lists=torch.empty(2, V, device=dev, dtype=torch.int32),
                   ranks=torch.empty(2, V, device=dev, dtype=torch.int32),
                   counts=torch.zeros(2, device=dev, dtype=torch.int32),
                   ws=torch.empty(max(int(L.coocc_gsf_compact_workspace(V)), 4), device=dev, dtype=torch.uint8),
                   dims=(X, Y, Z), C=C)
    else:
        assert out["dims"] == (X, Y, Z) and out["C"] == C
        out["cat"][:, 2 * C:].zero_()
        out["counts"].zero_()
    cat, flags, lists, ranks, counts, ws = (out[k] for k in ("cat", "flags", "lists", "ranks", "counts", "ws"))
    for i, t in enumerate((img5d, pts5d)):
        sB, sC, sX, sY, sZ = t.stride()
        _lib.check(_timed("hbm:gsf_pack", float(8 * V * C + V),
                          lambda t=t, i=i, sC=sC, sX=sX, sY=sY, sZ=sZ: L.coocc_gsf_pack(
                              _p(t), sC, sX, sY, sZ, C, X, Y, Z, _pb(cat, i * C * 4), 4 * C, _p(flags[i]), st)),
                   "gsf_pack")
        _lib.check(L.coocc_gsf_compact(_p(flags[i]), V, _p(lists[i]), _p(ranks[i]), _pb(counts, 4 * i),
                                       _p(ws), st), "gsf_compact")
    return out


# graph.GraphedStep sets this while it captures / replays a step: the prologue has already run
# (eagerly, into fixed buffers) and the occupied-voxel counts are known on the host, so the rest of
# the fuser needs no synchronisation.  dict(prologue=..., n_img, n_pts, nb_img, nb_pts): n_* exact
# counts (they select the reference's branches), nb_* >= n_* launch bounds baked into the graph
# (every kernel reads the exact count from `counts` on the device).
GSF_OVERRIDE = None


def gsf_index(img5d, pts5d, K, fix_k1_fps=False, want_parts=False):
    """Runs pack + compact + (FPS, top-K, ball-assign | direct NN) for both directions.
    Returns the [V,4C] concat buffer with img/pts slices filled and the index state."""
    L = _lib.lib()
    st = _stream()
    ov = GSF_OVERRIDE
    if ov is not None:
        pro = ov["prologue"]
        n_img, n_pts, nb_img, nb_pts = ov["n_img"], ov["n_pts"], ov["nb_img"], ov["nb_pts"]
        assert tuple(img5d.shape[2:]) == pro["dims"] and img5d.shape[1] == pro["C"]
    else:
        pro = gsf_prologue(img5d, pts5d)
        n_img, n_pts = (int(v) for v in pro["counts"].tolist())          # the one host sync of the fuser
        nb_img, nb_pts = n_img, n_pts
    cat, lists, ranks, counts = pro["cat"], pro["lists"], pro["ranks"], pro["counts"]
    dims, C = pro["dims"], pro["C"]
    X, Y, Z = dims
    V = X * Y * Z
    dev = cat.device
    state = dict(dims=dims, C=C, K=K, n_img=n_img, n_pts=n_pts, nb_img=nb_img, nb_pts=nb_pts, lists=lists,
                 ranks=ranks, counts=counts)

    # direction A: queries = LiDAR voxels, keys = image voxels (bifuser_n.py:137)
    # direction B: queries = image voxels, keys = LiDAR voxels (bifuser_n.py:151)
    specs = [("A", 1, 0, n_pts, nb_pts), ("B", 0, 1, n_img, nb_img)]
    fps_jobs = []
    for name, qi, ki, nq, nb in specs:
        if nq <= FPS_NUM:
            if ov is not None:
                raise RuntimeError("GSF_OVERRIDE covers the FPS branch only (N_query > 2048)")
            state[name] = _gsf_direction(L, st, lists[qi], ranks[qi], counts[qi:qi + 1], nq, lists[ki], ranks[ki],
                                         dims, K, dev)
        else:
            if K == 1 and not fix_k1_fps:
                # bifuser_n.py:62-85: the K==1 FPS branch never returns (Q11) -> None ->
                # the caller's gather broadcasts [1,4,C] against [N,C] and fails
                raise ReferenceQuirk("reference BiFuser_N(knum=1) fails for N_query > 2048 (fps_NN_fast returns "
                                     "None, bifuser_n.py:62-85); pass fix_k1_fps=True for the intended result")
            fps_jobs.append((name, qi, ki, nb))
    if fps_jobs and ov is not None and ov.get("tables") is not None:
        # index tables of this input computed ahead of time (graph.GraphedStep: on a side branch of the PREVIOUS step's
        # graph, under its GSFusion backward and optimizer)
        for name, qi, ki, nb in fps_jobs:
            state[name] = ov["tables"][name]
    elif fps_jobs:
        state.update(gsf_index_tables(pro, fps_jobs, K, want_parts))
    return cat, state


def gsf_index_tables(pro, fps_jobs, K, want_parts=False, out=None):
    """FPS + rep->key top-K + ball assignment for the directions in `fps_jobs` [(name, query grid, key grid, launch
    bound)] on the compacted lists of `pro` (gsf_prologue).  No host synchronisation; `out` = a previous result whose
    tensors are reused (fixed addresses for CUDA graphs).  Weight-independent: depends on the inputs' occupancy only."""
    L = _lib.lib()
    st = _stream()
    lists, ranks, counts = pro["lists"], pro["ranks"], pro["counts"]
    X, Y, Z = pro["dims"]
    V = X * Y * Z
    dev = lists.device
    res = {}
    mk = lambda name, key, *shape: (out[name][key] if out is not None else
                                    torch.empty(*shape, device=dev, dtype=torch.int32))
    reps = [mk(j[0], "rep_idx", FPS_NUM) for j in fps_jobs]
    j0 = fps_jobs[0]
    j1 = fps_jobs[1] if len(fps_jobs) > 1 else None
    _lib.check(L.coocc_gsf_fps(_p(lists[j0[1]]), _pb(counts, 4 * j0[1]), _p(reps[0]),
                               _p(lists[j1[1]]) if j1 else None, _pb(counts, 4 * j1[1]) if j1 else None,
                               _p(reps[1]) if j1 else None, max(j[3] for j in fps_jobs), FPS_NUM, Y, Z, st),
               "gsf_fps")
    for (name, qi, ki, nq), rep in zip(fps_jobs, reps):
        topk_idx = mk(name, "topk_idx", FPS_NUM, K)
        topk_d2 = mk(name, "topk_d2", FPS_NUM, K)
        _lib.check(L.coocc_gsf_rep_topk(_p(rep), FPS_NUM, _p(lists[qi]), _p(ranks[ki]), X, Y, Z, K,
                                        _p(topk_idx), _p(topk_d2), st), "gsf_rep_topk")
        if out is not None:
            winner = out[name]["winner"]
            winner.fill_(-1)
        else:
            winner = torch.full((K, V), -1, device=dev, dtype=torch.int32)
        group = torch.empty(FPS_NUM, BALL_SAMPLES, device=dev, dtype=torch.int32) if want_parts else None
        _lib.check(L.coocc_gsf_ball_assign(_p(rep), FPS_NUM, _p(lists[qi]), _p(ranks[qi]), _p(topk_idx), X, Y, Z,
                                           K, BALL_RADIUS, BALL_SAMPLES, V, _p(winner), _p(group), st),
                   "gsf_ball_assign")
        res[name] = dict(rep_idx=rep, nrep=FPS_NUM, topk_idx=topk_idx, topk_d2=topk_d2, winner=winner, group=group)
    return res


# set by BiFuser_N.forward right before the convolution that consumes the concat: its data gradient is the last
# convolution of a step's backward.  PRE_TAIL_HOOK (graph.GraphedStep) runs just before that launch.
PRE_TAIL_MARK = False
PRE_TAIL_HOOK = None

# graph.GraphedStep: called once at the start of the fuser's backward -- the point of a step behind which only
# HBM-bound work is left (GSFusion backward, gradient reduction tail, optimizer): the next step's index tables are
# computed on a side stream from there on
TAIL_HOOK = None


def gsf_nn_indices(state, name):
    """Reference-style result of fps_NN_fast for direction `name`: int64 [K, N_q] (-1 = unassigned)."""
    d = state[name]
    qi = 1 if name == "A" else 0
    nq = state["n_pts"] if name == "A" else state["n_img"]
    w = d["winner"][:, :nq].long()
    tk = d["topk_idx"].long()                      # [nrep, K]
    K = w.shape[0]
    out = torch.full_like(w, -1)
    for k in range(K):
        ok = w[k] >= 0
        out[k][ok] = tk[w[k][ok], k]
    return out


class _GSFusionFn(torch.autograd.Function):
    """cat[V,4C] = [img | pts | fused_img | fused_pts]  (bifuser_n.py:129-172)."""

    @staticmethod
    def forward(ctx, img5d, pts5d, knn_w, knn_b, K, fix_k1_fps):
        L = _lib.lib()
        st = _stream()
        cat, state = gsf_index(img5d, pts5d, K, fix_k1_fps)
        X, Y, Z = state["dims"]
        C = state["C"]
        dev = cat.device
        lists, counts = state["lists"], state["counts"]
        err = torch.zeros(1, device=dev, dtype=torch.int32)
        knn_w = knn_w.contiguous()
        knn_b = knn_b.contiguous()
        quirk_q2 = K > 1        # bifuser_n.py:158 indexes the image table with LiDAR-table positions
        # (direction, query list idx, own column, key grid column, lookup list idx, dst column)
        plan = [("A", 1, 1, 0, 0, 2), ("B", 0, 0, 1, 0 if quirk_q2 else 1, 3)]
        saved = {}
        for name, qi, own_col, key_col, look, dst_col in plan:
            d = state[name]
            nrep, nq = d["nrep"], (state["nb_pts"] if name == "A" else state["nb_img"])
            rows = torch.empty(K, nrep + 1, C, device=dev, dtype=torch.float32)
            _lib.check(L.coocc_gsf_gather_rows(_pb(cat, key_col * C * 4), 4 * C, _p(lists[look]),
                                               _pb(counts, 4 * look), _p(d["topk_idx"]), nrep, K, C, _p(rows),
                                               _p(err), st), "gsf_gather_rows")
            P = torch.empty(K, nrep + 1, C, device=dev, dtype=torch.float32)
            for k in range(K):
                # P[k] = rows[k] @ W_k^T,  W_k = knn_w[:, k*C:(k+1)*C]
                _lib.check(L.coocc_sgemm(nrep + 1, C, C, _p(rows[k]), C, 1, _pb(knn_w, k * C * 4), 1, K * C,
                                         _p(P[k]), C, 0, st), "sgemm")
            if nq > 0:
                _lib.check(L.coocc_gsf_modulate_fwd(_p(P), _p(knn_b), _p(d["winner"]), d["winner"].stride(0),
                                                    _p(lists[qi]), _pb(counts, 4 * qi), nq, nrep, K, C,
                                                    _pb(cat, own_col * C * 4), 4 * C, _pb(cat, dst_col * C * 4), 4 * C,
                                                    st), "gsf_modulate_fwd")
            saved[name] = (rows, P)
        if quirk_q2 and state["n_pts"] > state["n_img"] and int(err.item()) != 0:
            raise IndexError("index out of range: LiDAR-table position used on the image table "
                             "(reference bifuser_n.py:158, SURVEY Q2)")
        ctx.state, ctx.saved, ctx.plan = state, saved, plan
        ctx.save_for_backward(cat, knn_w, knn_b)
        return cat

    @staticmethod
    def backward(ctx, dcat):
        if TAIL_HOOK is not None:
            TAIL_HOOK()
        L = _lib.lib()
        st = _stream()
        cat, knn_w, knn_b = ctx.saved_tensors
        state, saved, plan = ctx.state, ctx.saved, ctx.plan
        K, C = state["K"], state["C"]
        X, Y, Z = state["dims"]
        dev = cat.device
        lists, counts = state["lists"], state["counts"]
        dcat = dcat.contiguous()
        if dcat.dtype != torch.float32:
            dcat = dcat.float()
        # d_img / d_pts are accumulated in place in the first two column blocks of dcat (row stride 4C): the
        # pass-through gradient of the concat is already there, the kernels add the gather / modulate terms
        dgrid = [dcat[:, 0:C], dcat[:, C:2 * C]]
        ldg = dcat.stride(0)
        dW = torch.zeros_like(knn_w)
        db = torch.zeros_like(knn_b)
        for name, qi, own_col, key_col, look, dst_col in plan:
            d = state[name]
            rows, P = saved[name]
            nrep, nq = d["nrep"], (state["nb_pts"] if name == "A" else state["nb_img"])
            if nq == 0:
                continue
            dP = torch.zeros_like(P)
            _lib.check(L.coocc_gsf_modulate_bwd(_p(P), _p(knn_b), _p(d["winner"]), d["winner"].stride(0),
                                                _p(lists[qi]), _pb(counts, 4 * qi), nq, nrep, K, C,
                                                _pb(cat, own_col * C * 4), 4 * C, _pb(dcat, dst_col * C * 4), 4 * C,
                                                _p(dgrid[own_col]), ldg, _p(dP), _p(db), st), "gsf_modulate_bwd")
            dF = torch.empty_like(P)
            for k in range(K):
                # dW[:, kC:(k+1)C] += dP[k]^T @ rows[k]
                _lib.check(L.coocc_sgemm(C, C, nrep + 1, _p(dP[k]), 1, C, _p(rows[k]), C, 1,
                                         _pb(dW, k * C * 4), K * C, 1, st), "sgemm")
                # dF[k] = dP[k] @ W_k
                _lib.check(L.coocc_sgemm(nrep + 1, C, C, _p(dP[k]), C, 1, _pb(knn_w, k * C * 4), K * C, 1,
                                         _p(dF[k]), C, 0, st), "sgemm")
            _lib.check(L.coocc_gsf_scatter_rows(_p(dF), _p(lists[look]), _pb(counts, 4 * look), _p(d["topk_idx"]),
                                                nrep, K, C, _p(dgrid[key_col]), ldg, st), "gsf_scatter_rows")
        d_img = to_5d(dgrid[0], (X, Y, Z))
        d_pts = to_5d(dgrid[1], (X, Y, Z))
        return d_img, d_pts, dW, db, None, None


def gsfusion_concat(img5d, pts5d, knn_w, knn_b, K, fix_k1_fps=False):
    return _GSFusionFn.apply(img5d, pts5d, knn_w, knn_b, K, fix_k1_fps)


# ----------------------------------------------------------------------------------------
# Volume rendering
# ----------------------------------------------------------------------------------------
def render_box(dims):
    L = _lib.lib()
    bx, by, bz = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    L.coocc_render_box(dims[0], dims[1], dims[2], ctypes.byref(bx), ctypes.byref(by), ctypes.byref(bz))
    return bx.value, by.value, bz.value


class _BoxRowsFn(torch.autograd.Function):
    """feature rows of the render box, [T,C], from the [V,C] grid."""

    @staticmethod
    def forward(ctx, x2d, dims):
        L = _lib.lib()
        x2d = _as_rows(x2d)
        box = render_box(dims)
        ctx.dims = dims
        ctx.shape = x2d.shape
        if box == tuple(dims):
            ctx.full = True
            return x2d.view_as(x2d)
        ctx.full = False
        T = box[0] * box[1] * box[2]
        ctx.bf16 = x2d.dtype == torch.bfloat16 and x2d.shape[1] % 8 == 0 and x2d.stride(0) % 8 == 0
        if ctx.bf16:
            # bf16 activations: rows gathered as stored (the heads' first Linear takes bf16 operands anyway)
            rows = torch.empty(T, x2d.shape[1], device=x2d.device, dtype=torch.bfloat16)
            _lib.check(L.coocc_render_box_gather_bf16(_p(x2d), x2d.stride(0), dims[0], dims[1], dims[2], x2d.shape[1],
                                                      _p(rows), _stream()), "render_box_gather_bf16")
            return rows
        if x2d.dtype != torch.float32:
            x2d = x2d.float()
        rows = torch.empty(T, x2d.shape[1], device=x2d.device, dtype=torch.float32)
        _lib.check(L.coocc_render_box_gather(_p(x2d), x2d.stride(0), dims[0], dims[1], dims[2], x2d.shape[1],
                                             _p(rows), _stream()), "render_box_gather")
        return rows

    @staticmethod
    def backward(ctx, g):
        if ctx.full:
            return g, None
        L = _lib.lib()
        g = g.contiguous()
        if ctx.bf16:
            # one pass writes the whole grid's gradient (zeros outside the box): no zero fill, no fp32 round trip
            g = g if g.dtype == torch.bfloat16 else g.to(torch.bfloat16)
            dx = torch.empty(ctx.shape, device=g.device, dtype=torch.bfloat16)
            _lib.check(L.coocc_render_box_scatter_bf16(_p(g), g.shape[1], ctx.dims[0], ctx.dims[1], ctx.dims[2],
                                                       _p(dx), dx.stride(0), _stream()), "render_box_scatter_bf16")
            return dx, None
        dx = torch.zeros(ctx.shape, device=g.device, dtype=torch.float32)
        _lib.check(L.coocc_render_box_scatter_add(_p(g), g.shape[1], ctx.dims[0], ctx.dims[1], ctx.dims[2],
                                                  _p(dx), dx.stride(0), _stream()), "render_box_scatter_add")
        return dx, None


class _CompositeFn(torch.autograd.Function):
    """tab[T,4] (rgb_raw, relu(sigma)) + geom -> (rgb_map [N,H,W,3], depth_map [N,H,W])."""

    @staticmethod
    def forward(ctx, tab, geom, dims):
        L = _lib.lib()
        N, D, H, W, _ = geom.shape
        tab = tab.contiguous()
        geom = geom.contiguous()
        rgb_map = torch.empty(N, H, W, 3, device=tab.device, dtype=torch.float32)
        depth_map = torch.empty(N, H, W, device=tab.device, dtype=torch.float32)
        err = torch.zeros(1, device=tab.device, dtype=torch.int32)
        _lib.check(L.coocc_render_composite_fwd(_p(geom), N, D, H, W, _p(tab), dims[0], dims[1], dims[2],
                                                _p(rgb_map), _p(depth_map), _p(err), _stream()),
                   "render_composite_fwd")
        box = render_box(dims)
        if box != (100, 100, 8):
            _raise_if_set(err, IndexError("render sample inside the 100x100x8 box but outside the feature grid "
                                          "(reference coocc_ray.py:385 raises, SURVEY Q6)"))
        ctx.save_for_backward(tab, geom)
        ctx.dims = dims
        return rgb_map, depth_map

    @staticmethod
    def backward(ctx, g_rgb, g_depth):
        L = _lib.lib()
        tab, geom = ctx.saved_tensors
        N, D, H, W, _ = geom.shape
        d_tab = torch.zeros_like(tab)
        _lib.check(L.coocc_render_composite_bwd(_p(geom), N, D, H, W, _p(tab), ctx.dims[0], ctx.dims[1],
                                                ctx.dims[2], _p(g_rgb.contiguous()), _p(g_depth.contiguous()),
                                                _p(d_tab), _stream()), "render_composite_bwd")
        return d_tab, None, None


class _UpsampleLossFn(torch.autograd.Function):
    """x16 bilinear + (loss_depth_render, loss_rgb)  (coocc_ray.py:412-433)."""

    @staticmethod
    def forward(ctx, rgb_map, depth_map, gt_img, gt_depth, D):
        L = _lib.lib()
        N, H, W, _ = rgb_map.shape
        dev = rgb_map.device
        rgbs = torch.empty(N, 16 * H, 16 * W, 3, device=dev, dtype=torch.float32)
        depths = torch.empty(N, 16 * H, 16 * W, device=dev, dtype=torch.float32)
        acc = torch.empty(3, device=dev, dtype=torch.float32)
        losses = torch.empty(2, device=dev, dtype=torch.float32)
        gt_img = gt_img.contiguous()
        gt_depth = gt_depth.contiguous()
        _lib.check(L.coocc_render_upsample_loss_fwd(_p(rgb_map.contiguous()), _p(depth_map.contiguous()), N, H, W, D,
                                                    _p(gt_img), _p(gt_depth), _p(rgbs), _p(depths), _p(acc),
                                                    _p(losses), _stream()), "render_upsample_loss_fwd")
        ctx.save_for_backward(rgbs, depths, gt_img, gt_depth, acc)
        ctx.meta = (N, H, W, D)
        ctx.mark_non_differentiable(rgbs, depths)
        return losses, rgbs, depths

    @staticmethod
    def backward(ctx, g_losses, _g1, _g2):
        L = _lib.lib()
        rgbs, depths, gt_img, gt_depth, acc = ctx.saved_tensors
        N, H, W, D = ctx.meta
        g_rgb = torch.empty(N, H, W, 3, device=rgbs.device, dtype=torch.float32)
        g_depth = torch.empty(N, H, W, device=rgbs.device, dtype=torch.float32)
        _lib.check(L.coocc_render_upsample_loss_bwd(_p(rgbs), _p(depths), N, H, W, D, _p(gt_img), _p(gt_depth),
                                                    _p(acc), _p(g_losses.contiguous()), _p(g_rgb), _p(g_depth),
                                                    _stream()), "render_upsample_loss_bwd")
        return g_rgb, g_depth, None, None, None


def box_rows(x2d, dims):
    return _BoxRowsFn.apply(x2d, tuple(dims))


def composite(tab, geom, dims):
    return _CompositeFn.apply(tab, geom, tuple(dims))


def upsample_losses(rgb_map, depth_map, gt_img, gt_depth, D):
    return _UpsampleLossFn.apply(rgb_map, depth_map, gt_img, gt_depth, D)


# ----------------------------------------------------------------------------------------
# Occupancy-head voxel losses (OccHead.loss_voxel, occ_head.py:267-293)
# ----------------------------------------------------------------------------------------
def downsample_labels(target_voxels, dims, empty_idx=0):
    """[1, X*r, Y*r, Z*r] integer labels -> int32 [X*Y*Z]: the reference's torch.mode vote
    (occ_head.py:269-280); r = 1 just converts."""
    L = _lib.lib()
    _require_cuda(target_voxels)
    assert target_voxels.dim() == 4 and target_voxels.shape[0] == 1, "batch 1 like the rest of the path"
    X, Y, Z = dims
    r = target_voxels.shape[2] // X      # occ_head.py:271 takes the ratio from the second spatial extent
    assert tuple(target_voxels.shape[1:]) == (X * r, Y * r, Z * r), "label grid must be an integer multiple of the output"
    if target_voxels.dtype not in (torch.uint8, torch.int32, torch.int64):
        target_voxels = target_voxels.long()
    t = target_voxels.contiguous()
    if r == 1:
        return t.reshape(-1).to(torch.int32)
    out = torch.empty(X * Y * Z, device=t.device, dtype=torch.int32)
    _lib.check(L.coocc_occ_label_mode(_p(t), t.element_size(), X, Y, Z, r, int(empty_idx), _p(out), _stream()),
               "occ_label_mode")
    return out


class _OccLossFn(torch.autograd.Function):
    """logits [V,C] fp32 rows, labels int32 [V] -> float[4] = (CE, sem_scal, geo_scal, lovasz)."""

    @staticmethod
    def forward(ctx, logits, labels, class_w, ignore, empty_idx):
        L = _lib.lib()
        _require_cuda(logits, labels)
        logits = _as_rows(logits.float() if logits.dtype != torch.float32 else logits)
        V, C = logits.shape
        assert labels.dtype == torch.int32 and labels.numel() == V
        nbytes = int(L.coocc_occ_loss_workspace(V, C))
        if nbytes < 0:
            raise RuntimeError("coocc_occ_loss: unsupported size V=%d C=%d" % (V, C))
        ws = torch.empty(nbytes, device=logits.device, dtype=torch.uint8)
        losses = torch.empty(4, device=logits.device, dtype=torch.float32)
        cw = class_w.to(device=logits.device, dtype=torch.float32).contiguous() if class_w is not None else None
        _lib.check(L.coocc_occ_loss_fwd(_p(logits), logits.stride(0), _p(labels), V, C, _p(cw), int(ignore),
                                        int(empty_idx), _p(ws), _p(losses), _stream()), "occ_loss_fwd")
        ctx.save_for_backward(logits, labels, cw, ws)
        ctx.meta = (int(ignore), int(empty_idx))
        return losses

    @staticmethod
    def backward(ctx, g):
        L = _lib.lib()
        logits, labels, cw, ws = ctx.saved_tensors
        ignore, empty_idx = ctx.meta
        V, C = logits.shape
        g = g.contiguous().float()
        ld = (C + 3) // 4 * 4
        d = torch.empty(V, ld, device=logits.device, dtype=torch.float32)
        _lib.check(L.coocc_occ_loss_bwd(_p(logits), logits.stride(0), _p(labels), V, C, _p(cw), ignore, empty_idx,
                                        _p(ws), _p(g), _p(d), ld, _stream()), "occ_loss_bwd")
        return d[:, :C], None, None, None, None


def occ_voxel_losses(logits2d, labels, class_w=None, ignore=255, empty_idx=0):
    return _OccLossFn.apply(logits2d, labels, class_w, ignore, empty_idx)


# ----------------------------------------------------------------------------------------
# Test-time metric (COOCC_Ray.evaluation_semantic, coocc_ray.py:659-684)
# ----------------------------------------------------------------------------------------
def eval_confusion(logits2d, dims, gt, visible_mask=None, empty_idx=0, ignore=255):
    """logits [V,C] fp32 rows of the [1,C,X,Y,Z] prediction, gt [1,GX,GY,GZ] integer labels ->
    (hist_ssc [C,C], hist_ssc_visible [C,C] or None, hist_sc [2,2]) int64 on the device: trilinear
    up-sampling to the label grid, argmax and fast_hist in one kernel."""
    L = _lib.lib()
    _require_cuda(logits2d, gt)
    x = _as_rows(logits2d.float() if logits2d.dtype != torch.float32 else logits2d)
    V, C = x.shape
    X, Y, Z = dims
    assert gt.dim() == 4 and gt.shape[0] == 1
    if gt.dtype not in (torch.uint8, torch.int32, torch.int64):
        gt = gt.long()
    g = gt.contiguous()
    GX, GY, GZ = g.shape[1:]
    vis = None
    if visible_mask is not None:
        vis = (visible_mask[0] != 0).to(torch.uint8).contiguous()
        assert tuple(vis.shape) == (GX, GY, GZ)
    h_ssc = torch.empty(C, C, device=x.device, dtype=torch.int64)
    h_vis = torch.empty(C, C, device=x.device, dtype=torch.int64) if vis is not None else None
    h_sc = torch.empty(2, 2, device=x.device, dtype=torch.int64)
    _lib.check(L.coocc_eval_confusion(_p(x), x.stride(0), X, Y, Z, C, _p(g), g.element_size(), GX, GY, GZ, _p(vis),
                                      int(empty_idx), int(ignore), _p(h_ssc), _p(h_vis), _p(h_sc), _stream()),
               "eval_confusion")
    return h_ssc, h_vis, h_sc


# ----------------------------------------------------------------------------------------
# OccHead fine / cascade stage (occ_head.py:182-237) -- csrc/fine_stage.cu, csrc/fine_select.cu.
# ----------------------------------------------------------------------------------------
def fine_select(logits2d, dims, empty_idx, ratio, topk, state):
    """Device-side replacement of `argmax != empty` -> nonzero -> random subset of `topk` parents -> ratio^3 children
    (occ_head.py:183-205, coordinate_transform.py:3-21) with no host synchronisation (csrc/fine_select.cu).
    logits2d [V,C] fp32 rows of the coarse prediction; state int64[2] on the device (seed, draw counter).
    Returns (coords int32 [3, ratio^3*topk], nsel int32 [2] = (N occupied, P = min(N, topk))); child slot o*topk + j
    belongs to parent slot j, slots j >= P are padding (coordinates 0)."""
    L = _lib.lib()
    _require_cuda(logits2d, state)
    x = _as_rows(logits2d.float() if logits2d.dtype != torch.float32 else logits2d)
    V, C = x.shape
    X, Y, Z = dims
    assert V == X * Y * Z and state.dtype == torch.int64 and state.numel() == 2
    M = ratio ** 3 * int(topk)
    coords = torch.empty(3, M, device=x.device, dtype=torch.int32)
    nsel = torch.empty(2, device=x.device, dtype=torch.int32)
    ws = torch.empty(int(L.coocc_fine_select_workspace(V)), device=x.device, dtype=torch.uint8)
    _lib.check(L.coocc_fine_select(_p(x), x.stride(0), X, Y, Z, C, int(empty_idx), int(ratio), int(topk), _p(state),
                                   _p(coords), _p(nsel), _p(ws), _stream()), "fine_select")
    return coords, nsel


def fine_gather_labels(coords, topk, nsel, target_voxels, ignore=255):
    """target_voxels[:, c0, c1, c2] (occ_head.py:298) as int32 [M]; padding slots of fine_select get `ignore`."""
    L = _lib.lib()
    _require_cuda(coords, target_voxels)
    assert coords.dtype == torch.int32 and coords.is_contiguous() and target_voxels.shape[0] == 1
    if target_voxels.dtype not in (torch.uint8, torch.int32, torch.int64):
        target_voxels = target_voxels.long()
    g = target_voxels.contiguous()
    M = coords.shape[1]
    labels = torch.empty(M, device=coords.device, dtype=torch.int32)
    _lib.check(L.coocc_fine_gather_labels(_p(coords), M, int(topk), _p(nsel), _p(g), g.element_size(), g.shape[1],
                                          g.shape[2], g.shape[3], int(ignore), _p(labels), _stream()),
               "fine_gather_labels")
    return labels


def _inv3x3(m):
    """closed-form inverse of [...,3,3] matrices from elementwise ops (torch.inverse goes through a solver library
    that may synchronise; this form can be captured in a CUDA graph)."""
    a, b, c = m[..., 0, :], m[..., 1, :], m[..., 2, :]
    r0, r1, r2 = torch.linalg.cross(b, c), torch.linalg.cross(c, a), torch.linalg.cross(a, b)
    det = (a * r0).sum(-1)
    return torch.stack([r0, r1, r2], -1) / det[..., None, None]
class _Sample3dFn(torch.autograd.Function):
    """feats [V,C] fp32 NDHWC rows, coords int32 [3,M] fine voxel indices -> [M,C] trilinear samples (:212-221)."""

    @staticmethod
    def forward(ctx, feats, dims, coords, final_size):
        L = _lib.lib()
        _require_cuda(feats, coords)
        ctx.in_dtype = feats.dtype
        if feats.dtype not in (torch.float32, torch.bfloat16):
            feats = feats.float()
        feats = _as_rows(feats)                  # bf16 grids are sampled as stored (no fp32 copy of the grid)
        C = feats.shape[1]
        coords = coords.to(torch.int32).contiguous()
        M = coords.shape[1]
        out = torch.empty(M, C, device=feats.device, dtype=torch.float32)
        fn = L.coocc_fine_sample3d_fwd_bf16 if feats.dtype == torch.bfloat16 else L.coocc_fine_sample3d_fwd
        _lib.check(fn(_p(feats), feats.stride(0), dims[0], dims[1], dims[2], C, _p(coords), M,
                      int(final_size[0]), int(final_size[1]), int(final_size[2]), _p(out), C, _stream()),
                   "fine_sample3d_fwd")
        ctx.save_for_backward(coords)
        ctx.meta = (tuple(dims), C, M, tuple(int(v) for v in final_size), feats.shape[0])
        return out

    @staticmethod
    def backward(ctx, g):
        L = _lib.lib()
        (coords,) = ctx.saved_tensors
        dims, C, M, fs, V = ctx.meta
        g = _as_rows(g.float())
        d = torch.zeros(V, C, device=g.device, dtype=torch.float32)
        _lib.check(L.coocc_fine_sample3d_bwd(_p(g), g.stride(0), dims[0], dims[1], dims[2], C, _p(coords), M, fs[0],
                                             fs[1], fs[2], _p(d), C, _stream()), "fine_sample3d_bwd")
        return d.to(ctx.in_dtype), None, None, None


def fine_sample_voxels(feats2d, dims, coords, final_size):
    return _Sample3dFn.apply(feats2d, tuple(dims), coords, final_size)


def fine_project(coords, rots, trans, intrins, post_rots, post_trans, bda, pts_range, W_img, H_img, grid_fine):
    """project_points_on_img (coordinate_transform.py:29-70, nuScenes): coords int [3,M]; rots/intrins/post_rots
    [n,3,3], trans/post_trans [n,3], bda [3,3] -> (uv [n,M,2] fp32, mask uint8 [M,n]).  No gradient (the
    reference computes it under no_grad)."""
    L = _lib.lib()
    _require_cuda(coords, rots)
    with torch.no_grad():
        n = rots.shape[0]
        pr = pts_range.detach().float().cpu()
        vs = (pr[3:] - pr[:3]) / torch.tensor([grid_fine[0] - 1, grid_fine[1] - 1, grid_fine[2] - 1])      # :33
        inv = _inv3x3 if torch.cuda.is_current_stream_capturing() else torch.inverse
        cam = torch.cat([inv(rots.float()).reshape(n, 9), trans.reshape(n, 3), intrins.reshape(n, 9),
                         post_rots[:, :2, :2].reshape(n, 4), post_trans[:, :2].reshape(n, 2)], 1).float().contiguous()
        inv_bda = inv(bda.float()).contiguous()
        coords = coords.to(torch.int32).contiguous()
        M = coords.shape[1]
        uv = torch.empty(n, M, 2, device=coords.device, dtype=torch.float32)
        mask = torch.empty(M, n, device=coords.device, dtype=torch.uint8)
        f3 = lambda t: (ctypes.c_float * 3)(*[float(v) for v in t.tolist()])
        _lib.check(L.coocc_fine_project(_p(coords), M, n, f3(vs), f3(pr[:3]), _p(inv_bda), _p(cam), float(W_img),
                                        float(H_img), _p(uv), _p(mask), _stream()), "fine_project")
    return uv, mask


class _Sample2dFn(torch.autograd.Function):
    """img [n*H*W, C] NHWC rows, uv [n,M,2], mask [M,n] -> [M,C]: masked sum over cameras of bilinear samples (:231-233)."""

    @staticmethod
    def forward(ctx, img, n, H, W, uv, mask):
        L = _lib.lib()
        ctx.in_dtype = img.dtype
        img = _as_rows(img.float() if img.dtype != torch.float32 else img)
        C = img.shape[1]
        M = uv.shape[1]
        out = torch.empty(M, C, device=img.device, dtype=torch.float32)
        _lib.check(L.coocc_fine_sample2d_fwd(_p(img), img.stride(0), n, H, W, C, _p(uv), _p(mask), M, _p(out), C,
                                             _stream()), "fine_sample2d_fwd")
        ctx.save_for_backward(uv, mask)
        ctx.meta = (n, H, W, C, M)
        return out

    @staticmethod
    def backward(ctx, g):
        L = _lib.lib()
        uv, mask = ctx.saved_tensors
        n, H, W, C, M = ctx.meta
        g = _as_rows(g.float())
        d = torch.zeros(n * H * W, C, device=g.device, dtype=torch.float32)
        _lib.check(L.coocc_fine_sample2d_bwd(_p(g), g.stride(0), n, H, W, C, _p(uv), _p(mask), M, _p(d), C, _stream()),
                   "fine_sample2d_bwd")
        return d.to(ctx.in_dtype), None, None, None, None, None


def fine_sample_images(img_rows, n, H, W, uv, mask):
    return _Sample2dFn.apply(img_rows, n, H, W, uv, mask)


class _GroupNormFn(torch.autograd.Function):
    """nn.GroupNorm(G) (+ ReLU) on rows [R, C]; `span` consecutive rows are one sample (1 = point rows, H*W = a map)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, G, span, eps, relu):
        L = _lib.lib()
        _require_cuda(x)
        ctx.in_dtype = x.dtype
        x = _as_rows(x.float() if x.dtype != torch.float32 else x)
        R, C = x.shape
        gamma, beta = gamma.float().contiguous(), beta.float().contiguous()
        stats = torch.empty(max(R // span, 1) * G * 2, device=x.device, dtype=torch.float32)
        y = torch.empty(R, C, device=x.device, dtype=torch.float32)
        _lib.check(L.coocc_groupnorm_fwd(_p(x), x.stride(0), R, C, G, span, _p(gamma), _p(beta), float(eps),
                                         1 if relu else 0, _p(stats), _p(y), C, _stream()), "groupnorm_fwd")
        ctx.save_for_backward(x, gamma, beta, stats)
        ctx.meta = (G, span, relu)
        return y

    @staticmethod
    def backward(ctx, dy):
        L = _lib.lib()
        x, gamma, beta, stats = ctx.saved_tensors
        G, span, relu = ctx.meta
        R, C = x.shape
        dy = _as_rows(dy.float())
        sums = torch.empty_like(stats)
        dx = torch.empty(R, C, device=x.device, dtype=torch.float32)
        dg = torch.zeros(C, device=x.device, dtype=torch.float32)
        db = torch.zeros(C, device=x.device, dtype=torch.float32)
        _lib.check(L.coocc_groupnorm_bwd(_p(x), x.stride(0), R, C, G, span, _p(gamma), _p(beta), 1 if relu else 0,
                                         _p(stats), _p(dy), dy.stride(0), _p(sums), _p(dx), C, _p(dg), _p(db), _stream()),
                   "groupnorm_bwd")
        return dx.to(ctx.in_dtype), dg, db, None, None, None, None


def group_norm_rows(x2d, gn, span=1, relu=True):
    """x2d [R,C] rows through the nn.GroupNorm parameter container `gn` (+ ReLU)."""
    return _GroupNormFn.apply(x2d, gn.weight, gn.bias, gn.num_groups, span, gn.eps, relu)
