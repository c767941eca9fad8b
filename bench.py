#!/usr/bin/env python
"""bench.py -- forward+backward voxels/sec of the Co-Occ fused-voxel hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload northstar|r50|r101|openocc] [--precision tf32|bf16|fp32]

One "step" = one training pass of the hot path over one synthetic scene per GPU:
GSFusion (BiFuser_N) -> CustomResNet3D-18 -> FPN3D -> OccHead coarse logits + OccHead.loss (label
vote to the working grid, CE + sem_scal + geo_scal + Lovasz-softmax),
the volume-render regulariser with its two losses, backward through all of it, the data-parallel
gradient all-reduce (N > 1) and the AdamW update.  `value` = X*Y*Z voxels of the working grid x
scenes per step / time, summed over ranks (weak scaling, one scene per GPU like samples_per_gpu=1).

--impl reference times the reference's own algorithm on the host CPUs (the oracle restatement,
oracle/oracle.py, which is bit-identical to the unmodified reference Python -- see
tests/test_oracle_vs_reference.py) on a bounded sample of the same workload.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "voxels/sec fwd+bwd (fused-voxel hot path: GSFusion + 3D conv decoder/head + volume render)"
UNIT = "voxels/s"
CPU_SAMPLE_GRID = (40, 40, 8)
# dram__bytes_read.sum + dram__bytes_write.sum of one tc_conv_kernel launch from the committed
# `ncu --set full` capture (tools/profile_round.sh -> profiles/r01b_ncu_full_tc_conv_bf16.summary.txt)
NCU_TRAFFIC = dict(bytes=286.06e6,
                   note="ncu --set full, fwd 3x3x3 128->128 on the 200x200x16 grid, bf16 in/out: DRAM read 164.85 MB + write "
                        "121.21 MB per launch vs 327.7 MB algorithmic (x read once, y written once, tail of y still in L2); "
                        "tensor pipe 79% active; dgrad 285.9 MB; wgrad 872 MB vs 327.7 MB algorithmic (the three ky taps "
                        "re-read X through L2 misses) at 89% tensor-pipe activity "
                        "(profiles/r01b_ncu_full_tc_conv_bf16.summary.txt)")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="northstar")
    ap.add_argument("--precision", default=os.environ.get("COOCC_PRECISION", "bf16"), choices=["tf32", "bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--optimizer", default=os.environ.get("COOCC_OPTIMIZER", "torch"), choices=["torch", "coocc"],
                    help="torch: torch.optim.AdamW(fused=True); coocc: co-occ_b200/optim.py FusedAdamW (one multi-tensor "
                         "kernel that also writes the bf16 weight operands) -- CPU-verified, not yet validated on a B200")
    ap.add_argument("--parity-mode", action="store_true",
                    help="also time the step in the fp32-accurate (3xTF32) arithmetic of the parity tests and report it "
                         "as `parity_mode` next to the headline")
    ap.add_argument("--launch", default=os.environ.get("COOCC_LAUNCH", "graph"), choices=["graph", "eager"],
                    help="graph: the step is replayed as one CUDA graph (co-occ_b200/graph.py); eager: one launch per kernel")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------
# CPU legs (oracle = the checker, used here only as the timed CPU baseline / reference arm)
# ------------------------------------------------------------------------------------------
def cpu_sample_inputs(cfg, seed=0):
    from coocc_b200 import synthetic as S
    grid = CPU_SAMPLE_GRID
    img, pts = S.make_voxel_feats(grid, cfg["C"], 0.6, 0.25, seed)
    inp = dict(img_voxel_feats=img, pts_voxel_feats=pts,
               geom=S.make_geom(grid, cfg["cams"], cfg["fH"], cfg["fW"], cfg["D"], seed))
    inp["gt_img"], inp["gt_depth"] = S.make_render_targets(cfg["cams"], cfg["fH"], cfg["fW"], seed)
    inp["gt_occ"] = S.make_gt_occ(grid, 2, seed)
    return inp, grid


def cpu_step_fn(cfg, seed=0):
    """fwd+bwd of the reference algorithm (oracle restatement) on the bounded sample."""
    from coocc_b200 import synthetic as S
    from oracle import oracle as O
    from oracle import losses as OL
    inp, grid = cpu_sample_inputs(cfg, seed)
    C, K = cfg["C"], cfg["K"]
    planes = [C, 2 * C, 4 * C, 8 * C]
    P = dict(occ_fuser=S.fuser_params(C, K), semantic_encoder=S.resnet3d_params(C, planes),
             semantic_neck=S.fpn3d_params(planes, 2 * C), pts_bbox_head=S.occhead_params([2 * C] * 4),
             render=S.render_params(C))
    P = {m: {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in p.items()} for m, p in P.items()}
    leaves = [v for p in P.values() for v in p.values() if v.requires_grad]
    torch.set_num_threads(os.cpu_count() or 1)

    def step():
        for v in leaves:
            v.grad = None
        img = inp["img_voxel_feats"].clone().requires_grad_(True)
        pts = inp["pts_voxel_feats"].clone().requires_grad_(True)
        out = O.hot_path_forward(P, dict(inp, img_voxel_feats=img, pts_voxel_feats=pts), K, tie="canonical")
        loss = sum(OL.loss_voxel(out["occ"], inp["gt_occ"]).values()) + out["loss_depth_render"] + out["loss_rgb"]
        loss.backward()
        return loss.item()

    return step, grid[0] * grid[1] * grid[2]


def cpu_stage_seconds(cfg, seed=0):
    """Forward-pass seconds of each stage of the CPU reference on the bounded sample."""
    from coocc_b200 import synthetic as S
    from oracle import oracle as O
    from oracle import losses as OL
    inp, grid = cpu_sample_inputs(cfg, seed)
    C, K = cfg["C"], cfg["K"]
    planes = [C, 2 * C, 4 * C, 8 * C]
    P = dict(occ_fuser=S.fuser_params(C, K), semantic_encoder=S.resnet3d_params(C, planes),
             semantic_neck=S.fpn3d_params(planes, 2 * C), pts_bbox_head=S.occhead_params([2 * C] * 4),
             render=S.render_params(C))
    out = {}
    with torch.no_grad():
        t = time.perf_counter()
        fused = O.bifuser_forward(P["occ_fuser"], inp["img_voxel_feats"], inp["pts_voxel_feats"], K, tie="canonical")
        out["occ_fuser"] = time.perf_counter() - t
        t = time.perf_counter()
        mid = O.resnet3d_forward(P["semantic_encoder"], fused)
        out["semantic_encoder"] = time.perf_counter() - t
        t = time.perf_counter()
        neck = O.fpn3d_forward(P["semantic_neck"], mid)
        out["semantic_neck"] = time.perf_counter() - t
        t = time.perf_counter()
        _, occ = O.occhead_coarse_forward(P["pts_bbox_head"], neck)
        OL.loss_voxel(occ, inp["gt_occ"])
        out["occ_head+loss"] = time.perf_counter() - t
        t = time.perf_counter()
        O.render_forward(P["render"], fused, inp["geom"], inp["gt_depth"], inp["gt_img"])
        out["render"] = time.perf_counter() - t
    return out


def cpu_baseline(cfg, warm=1, steps=2):
    step, nvox = cpu_step_fn(cfg)
    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return dict(value=nvox / dt, unit=UNIT, cores=os.cpu_count() or 1, kind="port",
                sample="oracle restatement of the reference (PyTorch CPU fp32, %d threads), fwd+bwd on a %dx%dx%d x C=%d grid, "
                       "K=%d, %d cams x %d rays x %d samples, %d timed steps of %.2f s"
                       % (os.cpu_count() or 1, *CPU_SAMPLE_GRID, cfg["C"], cfg["K"], cfg["cams"], cfg["fH"] * cfg["fW"],
                          cfg["D"], steps, dt),
                seconds_per_step=dt, forward_stage_seconds=cpu_stage_seconds(cfg))


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    step, nvox = cpu_step_fn(cfg)
    for _ in range(max(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    val = nvox / dt
    line = dict(metric=METRIC, value=val, unit=UNIT, impl="reference", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=dt * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32", data="synthetic",
                config=workload_config(args, cfg),
                cpu_baseline=dict(value=val, unit=UNIT, cores=os.cpu_count() or 1, kind="port",
                                  sample="each step = fwd+bwd of the reference algorithm on a %dx%dx%d x C=%d sample grid "
                                         "of the workload (%.2f s/step)" % (*CPU_SAMPLE_GRID, cfg["C"], dt)),
                e2e=dict(value=val, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def workload_config(args, cfg):
    X, Y, Z = cfg["grid"]
    return dict(workload="%s: %dx%dx%d working grid, C=%d, K=%d, %d cams x %d rays x %d samples, 1 scene/GPU"
                         % (args.workload, X, Y, Z, cfg["C"], cfg["K"], cfg["cams"], cfg["fH"] * cfg["fW"], cfg["D"]),
                step="GSFusion + ResNet3D-18 + FPN3D + OccHead(coarse) + OccHead.loss (label vote, CE, sem_scal, geo_scal, "
                     "Lovasz) + render losses, backward, grad all-reduce, AdamW",
                l2="inputs (%.0f MB/step) exceed the 126 MB L2" % (2 * X * Y * Z * cfg["C"] * 4 / 1e6),
                parallelism="dp%d (replicas only)" % args.gpus, precision=args.precision,
                optimizer="torch.optim.AdamW(fused)" if args.optimizer == "torch" else "coocc FusedAdamW + bf16 shadow",
                launch="cuda_graph (whole step, co-occ_b200/graph.py)" if args.launch == "graph" else "eager")


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------------------
# ours
# ------------------------------------------------------------------------------------------
def run_ours(args, cfg):
    import torch.distributed as dist
    import coocc_b200
    from coocc_b200 import functional as CF
    from coocc_b200 import synthetic as S
    from coocc_b200 import _lib
    from coocc_b200.ddp import GradReducer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if args.launch == "graph":
            # whole-step capture includes NCCL calls: the watchdog must not poll a capturing stream
            os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")
        dist.init_process_group("nccl", device_id=dev)
    coocc_b200.set_precision(args.precision)
    C, K = cfg["C"], cfg["K"]
    X, Y, Z = cfg["grid"]
    nvox = X * Y * Z

    torch.manual_seed(0)                      # reference initialisers, same weights on every rank
    model = coocc_b200.HotPath(coocc_b200.model_cfg(C, K), C).to(dev)
    model.train()
    params = [p for p in model.parameters() if p.requires_grad]
    reducer = GradReducer(params)
    if args.optimizer == "coocc":
        from coocc_b200.optim import FusedAdamW
        opt = FusedAdamW(params, lr=1e-4, weight_decay=0.01, shadow=(args.precision == "bf16"))
    else:
        opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=0.01, fused=True, capturable=(args.launch == "graph"))
    LOSS_KEYS = ["loss_voxel_ce_c_0", "loss_voxel_sem_scal_c_0", "loss_voxel_geo_scal_c_0", "loss_voxel_lovasz_c_0",
                 "loss_depth_render", "loss_rgb"]
    gstep = coocc_b200.GraphedStep(model, opt, reducer, LOSS_KEYS, enabled=(args.launch == "graph"))

    # ---- synthetic scene of this rank in pinned host memory (upstream memory layouts) ------
    seed = rank
    img_v, pts_v = S.make_voxel_feats(cfg["grid"], C, cfg["p_img"], cfg["p_pts"], seed)
    host = dict(img=img_v.permute(0, 1, 4, 2, 3).contiguous().pin_memory(),     # stored [1,C,Z,X,Y]
                pts=pts_v.permute(0, 1, 4, 3, 2).contiguous().pin_memory(),     # stored [1,C,Z,Y,X]
                geom=S.make_geom(cfg["grid"], cfg["cams"], cfg["fH"], cfg["fW"], cfg["D"], seed).pin_memory())
    gi, gd = S.make_render_targets(cfg["cams"], cfg["fH"], cfg["fW"], seed)
    host["gt_img"], host["gt_depth"] = gi.pin_memory(), gd.pin_memory()
    host["gt_occ"] = S.make_gt_occ(cfg["grid"], 2, seed).pin_memory()      # int64 labels at twice the working grid
    h2d_bytes = sum(t.numel() * t.element_size() for t in host.values())

    def to_device():
        d = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        d["img"] = d["img"].permute(0, 1, 3, 4, 2)      # logical [1,C,X,Y,Z], upstream strides
        d["pts"] = d["pts"].permute(0, 1, 4, 3, 2)
        return d

    def step(d):
        # forward + backward + gradient all-reduce + AdamW; replayed as one CUDA graph after the first
        # calls (--launch graph), launched kernel by kernel otherwise
        return gstep(d["img"], d["pts"], d["geom"], d["gt_depth"], d["gt_img"], d["gt_occ"])

    def eager_step(d):
        opt.zero_grad(set_to_none=True)
        losses, _, _ = model.forward_train(d["img"], d["pts"], d["geom"], d["gt_depth"], d["gt_img"], d["gt_occ"])
        loss = sum(losses[k] for k in LOSS_KEYS)
        loss.backward()
        reducer.finish()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    resident = to_device()
    for _ in range(max(args.warmup, 3)):
        step(resident)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    _lib.CALLS["n"] = 0
    r0 = gstep.stats["replays"]
    ms = timed(lambda: step(resident), args.steps)
    # kernel-launching C-ABI calls executed in the timed region: the eager ones + those replayed inside graphs
    launches = _lib.CALLS["n"] + (gstep.stats["replays"] - r0) * gstep.launches_per_replay
    # end-to-end: host (pinned) inputs in, loss out, every step.  The copy of step i+1's inputs is
    # issued on a side stream while step i computes (what a prefetching loader does); every step's
    # H2D copy and D2H loss read happen inside the timed region.
    last = {}
    copy_stream = torch.cuda.Stream()
    pending = {}

    def prefetch():
        with torch.cuda.stream(copy_stream):
            pending["d"] = to_device()
            pending["ev"] = torch.cuda.Event()
            pending["ev"].record(copy_stream)

    def e2e_step():
        torch.cuda.current_stream().wait_event(pending["ev"])
        d = pending["d"]
        for t in d.values():
            t.record_stream(torch.cuda.current_stream())
        prefetch()                                    # next step's inputs, overlapped with this step
        last["loss"] = float(step(d).item())

    prefetch()
    e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    clk = clocks.stop() if rank == 0 else None

    # ---- roofline of the dominant kernel (tc_conv_kernel), measured live with CUDA events ---
    CF.PROFILE = []
    eager_step(resident)
    torch.cuda.synchronize()
    prof = CF.PROFILE
    CF.PROFILE = None
    t_conv = sum(a.elapsed_time(b) for a, b, _, _ in prof) / 1e3
    f_conv = sum(f for _, _, f, _ in prof)
    n_conv = len(prof)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    peak = peaks.get("bf16_tflops_sustained")
    peak_src = "measured bf16 sustained (MEASURED_PEAKS.json)"
    if peak is None:
        peak, peak_src = 1400.0, "fallback sustained bf16 (B200_PROFILING.md)"
    achieved = f_conv / t_conv / 1e12 if t_conv > 0 else 0.0

    def shutdown():
        if world == 1:
            return
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        if args.launch == "graph":
            # measured on 2 x B200: destroy_process_group() does not return while CUDA graphs holding
            # captured NCCL kernels are alive; all ranks are past the final barrier, so just exit
            os._exit(0)
        dist.destroy_process_group()

    if rank != 0:
        shutdown()
        return
    value = nvox * world * args.steps / (ms / 1e3)
    e2e_val = nvox * world * args.steps / (ms_e2e / 1e3)
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
                ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype={"tf32": "tf32", "bf16": "bf16", "fp32": "f32"}[args.precision], data="synthetic",
                config=workload_config(args, cfg), clocks=clk,
                e2e=dict(value=e2e_val, unit=UNIT, ms_per_step=ms_e2e / args.steps, h2d_bytes_per_step=h2d_bytes,
                         d2h_bytes_per_step=4, loss=last.get("loss")),
                gpu_launches=launches,
                roofline=dict(bound="tensor", kernel="tc_conv_kernel (all conv/linear launches of one step)",
                              achieved=achieved, peak=peak, unit="TFLOP/s", frac=achieved / peak,
                              traffic=NCU_TRAFFIC["bytes"] if args.precision == "bf16" else None,
                              launches_per_step=n_conv, algorithmic_flops_per_step=f_conv,
                              kernel_ms_per_step=t_conv * 1e3, share_of_step=t_conv * 1e3 / (ms / args.steps),
                              peak_source=peak_src,
                              traffic_note=NCU_TRAFFIC["note"],
                              note="tf32 math has half the nominal bf16 rate" if args.precision == "tf32" else ""))
    if world == 1:      # (with more ranks SyncBN's all-reduces would need every rank to take part)
        try:
            line["forward_stage_ms"] = gpu_stage_ms(model, resident)
        except Exception as e:  # noqa: BLE001 -- informational only
            line["forward_stage_ms"] = dict(error=(str(e).splitlines() or [type(e).__name__])[0][:200])
    if gstep.capture_error is not None:
        line["config"]["launch"] = "eager (CUDA-graph capture failed: %s)" % gstep.capture_error
    if world == 1 and args.precision != "fp32" and args.parity_mode:
        # the same step in the parity arithmetic (3xTF32 split = fp32-accurate convolutions, fp32 storage), the mode
        # tests/test_gpu_parity.py holds to the 1e-3 bound; reported next to the bf16 headline, eager launches
        try:
            coocc_b200.set_precision("fp32")
            for _ in range(2):
                eager_step(resident)
            ms32 = timed(lambda: eager_step(resident), 3) / 3
            line["parity_mode"] = dict(precision="fp32 (3xTF32 split, fp32 activations)", ms_per_step=ms32,
                                       value=nvox / (ms32 / 1e3), unit=UNIT)
        except Exception as e:  # noqa: BLE001
            line["parity_mode"] = dict(error=(str(e).splitlines() or [type(e).__name__])[0][:200])
        finally:
            coocc_b200.set_precision(args.precision)
    line["parity"] = dict(mode=args.precision,
                          note="fp32 mode (3xTF32 split) matches the oracle to 3e-5 on logits; tf32 ~1e-2; bf16 ~5e-2 "
                               "(tests/test_gpu_parity.py, profiles/r01_gpu_parity_*.log); indices bit-exact in every mode")
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(cfg)
    print(json.dumps(line), flush=True)
    shutdown()


def gpu_stage_ms(model, d):
    """Forward-pass milliseconds of each stage (eager launches, CUDA events on the current stream), the GPU
    counterpart of cpu_baseline.forward_stage_seconds."""
    from coocc_b200.modules import render_fn
    marks = []

    def mark(name):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        marks.append((name, e))

    with torch.no_grad():
        for _ in range(2):                      # second pass is the timed one (first warms allocator / caches)
            marks.clear()
            mark("start")
            fused = model.occ_fuser(d["img"], d["pts"])
            mark("occ_fuser")
            mid = model.semantic_encoder(fused)
            mark("semantic_encoder")
            neck = model.semantic_neck(mid)
            mark("semantic_neck")
            outs = model.pts_bbox_head(voxel_feats=neck)
            model.pts_bbox_head.loss(output_voxels=outs["output_voxels"], target_voxels=d["gt_occ"])
            mark("occ_head+loss")
            render_fn(fused, d["geom"], model.sigma_head, model.rgb_head, d["gt_depth"], d["gt_img"])
            mark("render")
        torch.cuda.synchronize()
    return {marks[i][0]: marks[i - 1][1].elapsed_time(marks[i][1]) for i in range(1, len(marks))}


def main():
    args = parse()
    from coocc_b200 import synthetic as S
    cfg = S.CONFIGS[args.workload]
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
