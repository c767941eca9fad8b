"""ctypes binding of csrc/libcoocc_b200.so (the C ABI declared in include/coocc_b200.h).

There is no fallback: if the shared library is missing or fails to load, `lib()` raises.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
SO_PATH = os.path.join(CSRC, "libcoocc_b200.so")
if os.environ.get("COOCC_SO"):          # A/B runs against another build of the library (tools/)
    SO_PATH = os.environ["COOCC_SO"]

_lib = None

c_int, c_ll, c_void_p, c_float = ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_float

ERRORS = {-1: "invalid argument", -2: "pointer/stride not 16-byte aligned", -3: "CUDA driver entry point missing",
          -4: "tensor map encode failed", -5: "CUDA launch/runtime error", -6: "capacity exceeded"}


class ConvDesc(ctypes.Structure):
    _fields_ = [("X", c_int), ("Y", c_int), ("Z", c_int), ("Cin", c_int), ("Cout", c_int),
                ("ksize", c_int), ("stride", c_int), ("dtype", c_int), ("ldx", c_ll), ("ldy", c_ll),
                ("out_bf16", c_int)]


def build(verbose=False):
    """Compile every CUDA source for sm_100a into csrc/libcoocc_b200.so (nvcc cross-compiles
    without a GPU)."""
    out = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout[-4000:])
        print(out.stderr[-4000:])
    if out.returncode != 0:
        raise RuntimeError("building libcoocc_b200.so failed")
    return SO_PATH


_SIGS = {
    "coocc_version": (c_int, []),
    "coocc_conv3d_fwd": (c_int, [ctypes.POINTER(ConvDesc), c_void_p, c_void_p, c_void_p, c_ll, c_void_p, c_int, c_void_p, c_void_p]),
    "coocc_conv3d_dgrad": (c_int, [ctypes.POINTER(ConvDesc), c_void_p, c_void_p, c_void_p, c_ll, c_void_p]),
    "coocc_conv3d_dgrad_add": (c_int, [ctypes.POINTER(ConvDesc), c_void_p, c_void_p, c_void_p, c_ll, c_void_p, c_ll, c_void_p]),
    "coocc_conv3d_wgrad": (c_int, [ctypes.POINTER(ConvDesc), c_void_p, c_void_p, c_void_p, c_void_p]),
    "coocc_conv_tune": (c_int, [c_int, c_int]),
    "coocc_gsf_pack": (c_int, [c_void_p, c_ll, c_ll, c_ll, c_ll, c_int, c_int, c_int, c_int, c_void_p, c_ll, c_void_p, c_void_p]),
    "coocc_gsf_compact_workspace": (c_ll, [c_int]),
    "coocc_gsf_compact": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "coocc_gsf_fps": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "coocc_gsf_fps_tune": (c_int, [c_int, c_int]),
    "coocc_gsf_fps_signal": (c_int, [c_int]),
    "coocc_gsf_fps_gate": (c_int, [c_int, c_void_p]),
    "coocc_gsf_rep_topk": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "coocc_gsf_ball_assign": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "coocc_gsf_direct_nn": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "coocc_gsf_direct_winner": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "coocc_iota": (c_int, [c_void_p, c_int, c_void_p]),
    "coocc_gsf_gather_rows": (c_int, [c_void_p, c_ll, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "coocc_sgemm": (c_int, [c_int, c_int, c_int, c_void_p, c_ll, c_ll, c_void_p, c_ll, c_ll, c_void_p, c_ll, c_int, c_void_p]),
    "coocc_gsf_modulate_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_ll, c_void_p, c_ll, c_void_p]),
    "coocc_gsf_modulate_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_void_p, c_void_p]),
    "coocc_gsf_scatter_rows": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_ll, c_void_p]),
    "coocc_bn_finalize": (c_int, [c_void_p, c_int, c_ll, c_float, c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    "coocc_bn_act_fwd": (c_int, [c_void_p, c_ll, c_ll, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_int, c_void_p, c_ll, c_int, c_void_p]),
    "coocc_bn_act_bwd_reduce": (c_int, [c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_ll, c_ll, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p]),
    "coocc_bn_act_bwd_apply": (c_int, [c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_ll, c_ll, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_ll, c_void_p, c_ll, c_int, c_void_p, c_ll, c_void_p]),
    "coocc_dilate2": (c_int, [c_void_p, c_ll, c_int, c_int, c_int, c_int, c_void_p, c_ll, c_int, c_int, c_int, c_int, c_void_p]),
    "coocc_trilinear_fwd": (c_int, [c_void_p, c_ll, c_int, c_int, c_int, c_int, c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_ll, c_int, c_int, c_int, c_int, c_void_p]),
    "coocc_trilinear_bwd": (c_int, [c_void_p, c_ll, c_int, c_int, c_int, c_int, c_void_p, c_ll, c_void_p, c_ll, c_int, c_int, c_int, c_int, c_void_p]),
    "coocc_trilinear_mix_fwd": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_ll, c_int, c_int, c_int, c_int, c_void_p]),
    "coocc_trilinear_mix_bwd": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_ll, c_int, c_int, c_int, c_int, c_void_p]),
    "coocc_trilinear_tune": (c_int, [c_int]),
    "coocc_trilinear_wgrad": (c_int, [c_void_p, c_ll, c_void_p, c_ll, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_ll, c_int, c_void_p]),
    "coocc_render_box": (c_int, [c_int, c_int, c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_int), ctypes.POINTER(c_int)]),
    "coocc_render_box_gather": (c_int, [c_void_p, c_ll, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "coocc_render_box_scatter_add": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_ll, c_void_p]),
    "coocc_conv_set_sm_budget": (c_int, [c_int]),
    "coocc_conv_set_dynamic": (c_int, [c_int]),
    "coocc_relu_bias_bwd": (c_int, [c_void_p, c_ll, c_int, c_void_p, c_ll, c_int, c_int, c_int, c_void_p, c_ll, c_int, c_void_p, c_void_p]),
    "coocc_render_box_gather_bf16": (c_int, [c_void_p, c_ll, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "coocc_render_box_scatter_bf16": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_ll, c_void_p]),
    "coocc_render_composite_fwd": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "coocc_render_composite_bwd": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "coocc_render_upsample_loss_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "coocc_render_upsample_loss_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "coocc_occ_label_mode": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "coocc_occ_loss_workspace": (c_ll, [c_int, c_int]),
    "coocc_occ_loss_fwd": (c_int, [c_void_p, c_ll, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "coocc_occ_loss_bwd": (c_int, [c_void_p, c_ll, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_ll, c_void_p]),
    "coocc_eval_confusion": (c_int, [c_void_p, c_ll, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "coocc_lss_geometry": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "coocc_lss_workspace": (c_ll, [c_ll, c_ll]),
    "coocc_lss_sort": (c_int, [c_void_p, c_ll, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "coocc_lss_pool_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_ll, c_int, c_int, c_void_p, c_ll, c_void_p, c_int, c_int, c_void_p, c_ll, c_void_p]),
    "coocc_lss_pool_bwd": (c_int, [c_void_p, c_ll, c_int, c_int, c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_int, c_int, c_int, c_void_p, c_ll, c_void_p, c_void_p]),
    "coocc_fine_sample3d_fwd": (c_int, [c_void_p, c_ll, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_ll, c_void_p]),
    "coocc_fine_sample3d_fwd_bf16": (c_int, [c_void_p, c_ll, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_ll, c_void_p]),
    "coocc_fine_sample3d_bwd": (c_int, [c_void_p, c_ll, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_ll, c_void_p]),
    "coocc_fine_project": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float, c_void_p, c_void_p, c_void_p]),
    "coocc_fine_sample2d_fwd": (c_int, [c_void_p, c_ll, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_ll, c_void_p]),
    "coocc_fine_sample2d_bwd": (c_int, [c_void_p, c_ll, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_ll, c_void_p]),
    "coocc_groupnorm_fwd": (c_int, [c_void_p, c_ll, c_ll, c_int, c_int, c_int, c_void_p, c_void_p, c_float, c_int, c_void_p, c_void_p, c_ll, c_void_p]),
    "coocc_groupnorm_bwd": (c_int, [c_void_p, c_ll, c_ll, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_ll, c_void_p, c_void_p, c_ll, c_void_p, c_void_p, c_void_p]),
    "coocc_fine_select_workspace": (c_ll, [c_int]),
    "coocc_fine_select": (c_int, [c_void_p, c_ll, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p]),
    "coocc_fine_gather_labels": (c_int, [c_void_p, c_ll, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                         c_void_p, c_void_p]),
    "coocc_sp_flag_outputs": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "coocc_sp_neighbors": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "coocc_sp_gather_cols": (c_int, [c_void_p, c_ll, c_int, c_void_p, c_int, c_void_p, c_ll, c_void_p]),
    "coocc_sp_scatter_cols": (c_int, [c_void_p, c_ll, c_int, c_void_p, c_int, c_void_p, c_ll, c_void_p]),
    "coocc_peer_buffer_bytes": (c_ll, [c_int, c_int, c_int]),
    "coocc_peer_allreduce": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "coocc_adamw_step": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_float, c_float, c_float, c_float, c_float, c_void_p, c_int, c_void_p, c_void_p]),
}


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(SO_PATH):
            raise RuntimeError(
                "coocc_b200: %s not found -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU / PyTorch fallback for the hot path)" % SO_PATH)
        _lib = ctypes.CDLL(SO_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(_lib, name)      # AttributeError here == header/library mismatch
            fn.restype = res
            fn.argtypes = args
    return _lib


def exported_symbols():
    return list(_SIGS.keys())


CALLS = {"n": 0}      # number of kernel-launching C-ABI calls made (bench.py reports it)


def check(rc, what):
    CALLS["n"] += 1
    if rc != 0:
        raise RuntimeError("coocc_b200.%s failed: %s (%d)" % (what, ERRORS.get(rc, "unknown"), rc))
