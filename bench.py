#!/usr/bin/env python
"""bench.py -- forward+backward voxels/sec of the Co-Occ fused-voxel hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload northstar|r50|r101|openocc|stress] [--precision bf16|tf32|fp32]

One "step" = one training pass of the hot path over one synthetic scene per GPU, built from the reference config's
own model dicts (coocc_multi_r50_256x704.py:136-178): GSFusion (BiFuser_N, knum=2) -> CustomResNet3D-18 -> FPN3D ->
OccHead (coarse logits, the fine / cascade stage with cascade_ratio=2, fine_topk=15000, sample_from_voxel/img) +
OccHead.loss (label vote, CE + sem_scal + geo_scal + Lovasz on the coarse grid and on the fine points), the
volume-render regulariser with its two losses, backward through all of it, the data-parallel gradient all-reduce
(N > 1) and the AdamW update.  `value` = X*Y*Z voxels of the working grid x scenes per step / time, summed over ranks
(weak scaling, one scene per GPU like samples_per_gpu=1).

The N=1 line also carries: `precision_runs` (the same step in tf32 and in the fp32-accurate 3xTF32 mode the parity
tests hold to 1e-3), `roofline` of the dominant kernel plus `hbm_kernels` (the bandwidth-bound kernels, CUDA-event
timed), `cpu_baseline` (the reference's algorithm on the host CPUs) and `replicas_identical` for N > 1.

--impl reference times the reference's own algorithm on the host CPUs (the oracle restatement, oracle/oracle.py +
oracle/finestage.py + oracle/losses.py, bit-identical to the unmodified reference Python -- see
tests/test_oracle_vs_reference.py) on the reference's own working grid, 100x100x8 (config.workload says so).
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
# The CPU legs run the reference's own working grid (coocc_multi_r50_256x704.py:23-31: occ_size / lss_downsample):
# a full north-star step (8x the voxels) takes about a minute per step on 16 host cores.
CPU_GRID = (100, 100, 8)
COARSE_KEYS = ["loss_voxel_ce_c_0", "loss_voxel_sem_scal_c_0", "loss_voxel_geo_scal_c_0", "loss_voxel_lovasz_c_0"]
FINE_KEYS = ["loss_voxel_ce_fine", "loss_voxel_sem_scal_fine", "loss_voxel_geo_scal_fine", "loss_voxel_lovasz_fine"]
RENDER_KEYS = ["loss_depth_render", "loss_rgb"]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="northstar")
    ap.add_argument("--precision", default=os.environ.get("COOCC_PRECISION", "bf16"), choices=["tf32", "bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-precision-runs", action="store_true",
                    help="skip the tf32 / fp32 (3xTF32) timings that the N=1 line carries next to the headline")
    ap.add_argument("--optimizer", default=os.environ.get("COOCC_OPTIMIZER", "coocc"), choices=["torch", "coocc"],
                    help="coocc: co-occ_b200/optim.py FusedAdamW (one multi-tensor kernel that also writes the bf16 weight "
                         "operands); torch: torch.optim.AdamW(fused=True)")
    ap.add_argument("--no-peer", action="store_true", help="SyncBN statistics through NCCL instead of the peer-memory kernel")
    ap.add_argument("--no-pipeline", action="store_true", help="no index pipelining across steps (GraphedStep next_inputs)")
    ap.add_argument("--no-fine", action="store_true", help="coarse head only (cascade_ratio=1), the round-1 step")
    ap.add_argument("--launch", default=os.environ.get("COOCC_LAUNCH", "graph"), choices=["graph", "eager"],
                    help="graph: the step is replayed as one CUDA graph (co-occ_b200/graph.py); eager: one launch per kernel")
    return ap.parse_args()


def traffic_record():
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from this round's `ncu --set full` capture
    (tools/profile_round.sh -> tools/ncu_summary.py writes profiles/traffic.json); None if no capture is committed."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:  # noqa: BLE001
        return None


# ------------------------------------------------------------------------------------------
# CPU legs (oracle = the checker, used here only as the timed CPU baseline / reference arm)
# ------------------------------------------------------------------------------------------
def cpu_sample_inputs(cfg, seed=0):
    from coocc_b200 import synthetic as S
    grid = CPU_GRID
    img, pts = S.make_voxel_feats(grid, cfg["C"], cfg["p_img"], cfg["p_pts"], seed)
    inp = dict(img_voxel_feats=img, pts_voxel_feats=pts,
               geom=S.make_geom(grid, cfg["cams"], cfg["fH"], cfg["fW"], cfg["D"], seed))
    inp["gt_img"], inp["gt_depth"] = S.make_render_targets(cfg["cams"], cfg["fH"], cfg["fW"], seed)
    inp["gt_occ"] = S.make_gt_occ(grid, 2, seed)
    inp["img_feats"] = S.make_img_feats(cfg["cams"], cfg["fH"], cfg["fW"], seed)
    inp["transform"] = S.make_transform(cfg["cams"], cfg["fH"], cfg["fW"], seed)
    return inp, grid


def cpu_params(cfg, fine):
    from coocc_b200 import synthetic as S
    C, K = cfg["C"], cfg["K"]
    planes = [C, 2 * C, 4 * C, 8 * C]
    head = S.occhead_params([2 * C] * 4)
    if fine:
        head.update(S.fine_head_params())
    P = dict(occ_fuser=S.fuser_params(C, K), semantic_encoder=S.resnet3d_params(C, planes),
             semantic_neck=S.fpn3d_params(planes, 2 * C), pts_bbox_head=head, render=S.render_params(C))
    return {m: {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in p.items()} for m, p in P.items()}


def cpu_step_fn(cfg, fine, seed=0):
    """fwd+bwd of the reference algorithm (oracle restatement) on the reference's working grid."""
    from oracle import finestage as OF
    from oracle import losses as OL
    from oracle import oracle as O
    inp, grid = cpu_sample_inputs(cfg, seed)
    K = cfg["K"]
    P = cpu_params(cfg, fine)
    leaves = [v for p in P.values() for v in p.values() if v.requires_grad]
    torch.set_num_threads(os.cpu_count() or 1)
    occ_size = [2 * g for g in grid]
    pcr = torch.tensor([-50.0, -50.0, -5.0, 50.0, 50.0, 3.0])

    def step():
        for v in leaves:
            v.grad = None
        img = inp["img_voxel_feats"].clone().requires_grad_(True)
        pts = inp["pts_voxel_feats"].clone().requires_grad_(True)
        out = O.hot_path_forward(P, dict(inp, img_voxel_feats=img, pts_voxel_feats=pts), K, tie="canonical")
        loss = sum(OL.loss_voxel(out["occ"], inp["gt_occ"]).values()) + out["loss_depth_render"] + out["loss_rgb"]
        if fine:                                                     # occ_head.py:182-237, 295-337
            fc, fo = OF.fine_forward(P["pts_bbox_head"], out["out_voxel_feats"], out["occ"], inp["img_feats"],
                                     inp["transform"], occ_size, pcr, 2, 15000)
            loss = loss + sum(OF.loss_point(fc, fo, inp["gt_occ"]).values())
        loss.backward()
        return loss.item()

    return step, grid[0] * grid[1] * grid[2]


def cpu_stage_seconds(cfg, seed=0):
    """Forward-pass seconds of each stage of the CPU reference on its working grid."""
    from oracle import losses as OL
    from oracle import oracle as O
    inp, grid = cpu_sample_inputs(cfg, seed)
    K = cfg["K"]
    P = cpu_params(cfg, False)
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


def cpu_sample_text(cfg, fine, dt, steps):
    return ("oracle restatement of the reference (PyTorch CPU fp32, %d threads): fwd+bwd of the whole step%s on the "
            "reference's own working grid %dx%dx%d (= 1/%d of the workload's voxels per step), C=%d, K=%d, %d cams x %d rays "
            "x %d samples; %d timed step(s) of %.2f s; value = voxels of THAT grid / s"
            % (os.cpu_count() or 1, " incl. fine stage" if fine else "", *CPU_GRID,
               max(1, cfg["grid"][0] * cfg["grid"][1] * cfg["grid"][2] // (CPU_GRID[0] * CPU_GRID[1] * CPU_GRID[2])),
               cfg["C"], cfg["K"], cfg["cams"], cfg["fH"] * cfg["fW"], cfg["D"], steps, dt))


def cpu_baseline(cfg, fine, warm=1, steps=1):
    step, nvox = cpu_step_fn(cfg, fine)
    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return dict(value=nvox / dt, unit=UNIT, cores=os.cpu_count() or 1, kind="port",
                sample=cpu_sample_text(cfg, fine, dt, steps), grid=list(CPU_GRID),
                seconds_per_step=dt, forward_stage_seconds=cpu_stage_seconds(cfg))


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fine = not args.no_fine and cfg["C"] == 128
    step, nvox = cpu_step_fn(cfg, fine)
    # a step on the reference's working grid costs 10-30 s of CPU: the requested W + K steps run as long as they fit a
    # wall-clock budget (COOCC_CPU_BUDGET_S, default 300 s); `steps` / `warmup` below are the counts actually run
    budget = float(os.environ.get("COOCC_CPU_BUDGET_S", "300"))
    t_start = time.perf_counter()
    n_warm = 0
    while n_warm < args.warmup and (n_warm == 0 or time.perf_counter() - t_start < 0.15 * budget):
        step()
        n_warm += 1
    t0 = time.perf_counter()
    n_steps = 0
    while n_steps < args.steps and (n_steps == 0 or (time.perf_counter() - t_start) * (1 + 1.0 / (n_steps + n_warm)) < budget):
        step()
        n_steps += 1
    dt = (time.perf_counter() - t0) / n_steps
    val = nvox / dt
    requested = dict(steps=args.steps, warmup=args.warmup)
    args.steps, args.warmup = n_steps, n_warm
    conf = workload_config(args, cfg, fine)
    # what actually ran: the CPU arm's grid, not the GPU arm's
    conf["workload"] = ("CPU arm of %s: %dx%dx%d working grid (the reference's own, 1/%d of the GPU arm's voxels per step), "
                        "C=%d, K=%d, %d cams x %d rays x %d samples, one process on the host cores"
                        % (args.workload, *CPU_GRID,
                           max(1, cfg["grid"][0] * cfg["grid"][1] * cfg["grid"][2] // (CPU_GRID[0] * CPU_GRID[1] * CPU_GRID[2])),
                           cfg["C"], cfg["K"], cfg["cams"], cfg["fH"] * cfg["fW"], cfg["D"]))
    conf["gpu_arm_workload"] = workload_config(args, cfg, fine)["workload"]
    conf["precision"], conf["optimizer"], conf["launch"] = "f32", "none (fwd+bwd only)", "torch CPU"
    conf["parallelism"] = "1 CPU process (rank 0 only), %d threads" % (os.cpu_count() or 1)
    line = dict(metric=METRIC, value=val, unit=UNIT, impl="reference", n_gpus=args.gpus, gpus_used=0, steps=args.steps,
                warmup=args.warmup, ms_per_step=dt * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32", data="synthetic", config=conf, requested=requested,
                cpu_baseline=dict(value=val, unit=UNIT, cores=os.cpu_count() or 1, kind="port", grid=list(CPU_GRID),
                                  sample=cpu_sample_text(cfg, fine, dt, args.steps)),
                e2e=dict(value=val, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def workload_config(args, cfg, fine):
    X, Y, Z = cfg["grid"]
    return dict(workload="%s: %dx%dx%d working grid, C=%d, K=%d, %d cams x %d rays x %d samples, 1 scene/GPU"
                         % (args.workload, X, Y, Z, cfg["C"], cfg["K"], cfg["cams"], cfg["fH"] * cfg["fW"], cfg["D"]),
                step="GSFusion + ResNet3D-18 + FPN3D + OccHead (coarse%s) + OccHead.loss (label vote; CE, sem_scal, geo_scal, "
                     "Lovasz on the coarse grid%s) + render losses, backward, grad all-reduce, AdamW"
                     % ((" + fine/cascade stage: cascade_ratio=2, fine_topk=15000, voxel + 6-camera image sampling", " and on the "
                         "fine points") if fine else ("", "")),
                l2="inputs (%.0f MB/step) exceed the 126 MB L2" % (2 * X * Y * Z * cfg["C"] * 4 / 1e6),
                parallelism="dp%d (replicas only)" % args.gpus, precision=args.precision,
                optimizer="torch.optim.AdamW(fused)" if args.optimizer == "torch" else
                "coocc FusedAdamW (norm_decay_mult 0, grad_clip 5, bf16 shadow, gradient arena)",
                launch=("cuda_graph (whole step, co-occ_b200/graph.py%s)" % (
                    "" if args.no_pipeline else "; the next batch's neighbour-search tables are computed on a side branch "
                    "of the graph, under the GSFusion backward and the optimizer")) if args.launch == "graph" else "eager",
                setup_steps="3 untimed calls before the warm-up (lazy init, one graph capture per buffer parity)")


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
    syncbn = "single process"
    if world > 1:
        syncbn = "NCCL all-reduce per BatchNorm"
        if not args.no_peer:
            try:
                from coocc_b200.peer import PeerExchange
                CF.PEER = PeerExchange()
                syncbn = "peer-memory all-reduce kernel over NVLink (csrc/peer_reduce.cu)"
            except Exception as e:  # noqa: BLE001 -- no peer access between the devices: NCCL carries the statistics
                syncbn += " (peer exchange unavailable: %s)" % (str(e).splitlines() or [type(e).__name__])[0][:120]
    C, K = cfg["C"], cfg["K"]
    X, Y, Z = cfg["grid"]
    nvox = X * Y * Z
    fine = not args.no_fine and C == 128

    torch.manual_seed(0)                      # reference initialisers, same weights on every rank
    model = coocc_b200.HotPath(coocc_b200.model_cfg(C, K, fine=fine, grid=cfg["grid"]), C).to(dev)
    model.train()
    params = [p for p in model.parameters() if p.requires_grad]
    # gradients live in one flat arena: the weight-gradient kernels accumulate into it, the all-reduce buckets are
    # slices of it, AdamW reads and clears it (co-occ_b200/ddp.py)
    arena = None
    if args.optimizer == "coocc":
        from coocc_b200.ddp import GradArena
        arena = GradArena(params)
    reducer = GradReducer(params, arena=arena, bucket_bytes=int(os.environ.get("COOCC_BUCKET_MB", "16")) << 20)

    def make_opt():
        # the reference's optimizer recipe (coocc_multi_r50_256x704.py:263-279): AdamW lr 1e-4, weight_decay 0.01,
        # norm_decay_mult 0, grad_clip max_norm 5
        if args.optimizer == "coocc":
            from coocc_b200.optim import FusedAdamW, norm_decay_mults
            return FusedAdamW(params, lr=1e-4, weight_decay=0.01, shadow=True, arena=arena,
                              param_mults=norm_decay_mults(model, 0.0), max_norm=5.0)
        return torch.optim.AdamW(params, lr=1e-4, weight_decay=0.01, fused=True, capturable=(args.launch == "graph"))

    opt = make_opt()
    LOSS_KEYS = COARSE_KEYS + (FINE_KEYS if fine else []) + RENDER_KEYS
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
    tr_host = None
    if fine:
        host["img_feats"] = S.make_img_feats(cfg["cams"], cfg["fH"], cfg["fW"], seed).pin_memory()
        tr_host = S.make_transform(cfg["cams"], cfg["fH"], cfg["fW"], seed)
        for i in range(6):                                                 # the calibration matrices, img_inputs[1:7]
            host["tr%d" % i] = tr_host[i].contiguous().pin_memory()
    h2d_bytes = sum(t.numel() * t.element_size() for t in host.values())

    def to_device():
        d = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        d["img"] = d["img"].permute(0, 1, 3, 4, 2)      # logical [1,C,X,Y,Z], upstream strides
        d["pts"] = d["pts"].permute(0, 1, 4, 3, 2)
        d["transform"] = None
        if fine:
            d["transform"] = tuple(d["tr%d" % i] for i in range(6)) + tuple(tr_host[6:])
        return d

    def call(fn, d):
        return fn(d["img"], d["pts"], d["geom"], d["gt_depth"], d["gt_img"], d["gt_occ"], d.get("img_feats"), d["transform"])

    def step(d, nxt=None):
        # forward + backward + gradient all-reduce + AdamW; replayed as one CUDA graph after the first
        # calls (--launch graph), launched kernel by kernel otherwise.  nxt = the following step's inputs (already on
        # the device, as a prefetching loader has them): its neighbour-search tables are computed inside this step's
        # graph, under the GSFusion backward and the optimizer (graph.GraphedStep, index pipelining)
        if nxt is None or args.no_pipeline:
            return call(gstep, d)
        return gstep(d["img"], d["pts"], d["geom"], d["gt_depth"], d["gt_img"], d["gt_occ"], d.get("img_feats"),
                     d["transform"], next_inputs=(nxt["img"], nxt["pts"]))

    def eager_step(d, optimizer=None):
        o = optimizer or opt
        o.zero_grad(set_to_none=True)
        reducer.begin()
        CF.begin_step()
        CF.zero_pool_begin(dev)
        losses, _, _ = call(model.forward_train, d)
        loss = sum(losses[k] for k in LOSS_KEYS)
        loss.backward()
        CF.zero_pool_end()
        reducer.finish()
        o.step()
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
    for _ in range(3):                       # setup: lazy init (eager), graph captures (two buffer parities)
        step(resident, resident)
    for _ in range(args.warmup):
        step(resident, resident)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    _lib.CALLS["n"] = 0
    r0 = gstep.stats["replays"]
    ms = timed(lambda: step(resident, resident), args.steps)
    # kernel-launching C-ABI calls executed in the timed region: the eager ones + those replayed inside graphs
    launches = _lib.CALLS["n"] + (gstep.stats["replays"] - r0) * gstep.launches_per_replay
    # end-to-end: host (pinned) inputs in, loss out, every step.  The copy of step i+1's inputs is
    # issued on a side stream while step i computes (what a prefetching loader does); every step's
    # H2D copy and D2H loss read happen inside the timed region.
    last = {}
    copy_stream = torch.cuda.Stream()
    queue = []                 # [(inputs on the device, copy-done event)]: this step's and the next step's

    def prefetch():
        with torch.cuda.stream(copy_stream):
            d = to_device()
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        queue.append((d, ev))

    def e2e_step():
        # a loader two batches ahead: step i runs on batch i and prepares the index tables of batch i+1 (both on the
        # device by now); the copy of batch i+2 is issued here and overlaps this step
        cur = torch.cuda.current_stream()
        for d_, ev_ in queue[:2]:
            cur.wait_event(ev_)
            for t in d_.values():
                if torch.is_tensor(t):
                    t.record_stream(cur)
        d, nxt = queue[0][0], queue[1][0]
        queue.pop(0)
        prefetch()
        last["loss"] = float(step(d, nxt).item())

    prefetch()
    prefetch()
    e2e_step()
    e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    clk = clocks.stop() if rank == 0 else None

    # ---- multi-GPU correctness: replicas must hold bit-identical parameters after the run ----
    replicas_identical = None
    if world > 1:
        cs = torch.zeros(2, device=dev, dtype=torch.float64)
        for p in params:
            f = p.detach().double()
            cs[0] += f.sum()
            cs[1] += (f * f).sum()
        lo, hi = cs.clone(), cs.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        replicas_identical = bool(torch.equal(lo, hi))

    # ---- rooflines, measured live with CUDA events: one eager step with every conv / BN / pack launch bracketed ---
    def bracketed_step(opt_=None):
        """one eager step with every conv / BN / pack launch bracketed by CUDA events -> [(ms, work, tag)].  Run twice,
        per-launch minimum: a host hiccup longer than the GPU's head start would otherwise land inside a bracket."""
        runs = []
        for _ in range(2):
            CF.PROFILE = []
            CF.PROFILE_AHEAD_MS = 40    # brackets must time kernels, not the host's launch overhead (see CF._timed)
            if opt_ is None:
                eager_step(resident)
            else:
                eager_step(resident, opt_)
            torch.cuda.synchronize()
            prof, CF.PROFILE = CF.PROFILE, None
            runs.append([(a.elapsed_time(b), w, t) for a, b, w, t in prof])
        if len(runs[0]) != len(runs[1]) or any(x[2] != y[2] for x, y in zip(*runs)):
            return runs[1]
        return [(min(x[0], y[0]), x[1], x[2]) for x, y in zip(*runs)]

    def profile_step():
        prof = bracketed_step()
        conv = [(ms_, w) for ms_, w, t in prof if not t.startswith("hbm:")]
        hbm = {}
        for ms_, w, t in prof:
            if t.startswith("hbm:"):
                r = hbm.setdefault(t[4:], [0.0, 0.0, 0])
                r[0] += ms_
                r[1] += w
                r[2] += 1
        return conv, hbm

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    peak = peaks.get("bf16_tflops_sustained")
    peak_src = "measured bf16 sustained (MEASURED_PEAKS.json)"
    if peak is None:
        peak, peak_src = 1400.0, "fallback sustained bf16 (B200_PROFILING.md)"
    hbm_peak = peaks.get("hbm_gbs") or 6550.0

    def conv_roofline(conv):
        t = sum(c[0] for c in conv) / 1e3
        f = sum(c[1] for c in conv)
        return dict(achieved=f / t / 1e12 if t > 0 else 0.0, kernel_ms_per_step=t * 1e3, launches_per_step=len(conv),
                    algorithmic_flops_per_step=f)

    conv, hbm = profile_step()
    cr = conv_roofline(conv)

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
    tr = traffic_record()
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype={"tf32": "tf32", "bf16": "bf16", "fp32": "f32"}[args.precision], data="synthetic",
                config=dict(workload_config(args, cfg, fine), sync_bn=syncbn), clocks=clk,
                e2e=dict(value=e2e_val, unit=UNIT, ms_per_step=ms_e2e / args.steps, h2d_bytes_per_step=h2d_bytes,
                         d2h_bytes_per_step=4, loss=last.get("loss")),
                gpu_launches=launches,
                roofline=dict(bound="tensor", kernel="tc_conv_kernel (all conv/linear launches of one step)",
                              achieved=cr["achieved"], peak=peak, unit="TFLOP/s", frac=cr["achieved"] / peak,
                              traffic=(tr or {}).get("bytes") if args.precision == "bf16" else None,
                              launches_per_step=cr["launches_per_step"],
                              algorithmic_flops_per_step=cr["algorithmic_flops_per_step"],
                              kernel_ms_per_step=cr["kernel_ms_per_step"],
                              share_of_step=cr["kernel_ms_per_step"] / (ms / args.steps), peak_source=peak_src,
                              traffic_note=(tr or {}).get("note", "no ncu --set full capture committed for this round"),
                              note="tf32 math has half the nominal bf16 rate" if args.precision == "tf32" else ""))
    # HBM-bound kernels of the step (algorithmic bytes / CUDA-event time, against the measured copy bandwidth)
    line["hbm_kernels"] = {k: dict(ms_per_step=v[0], launches=v[2], algorithmic_bytes=v[1],
                                   achieved_gbs=v[1] / (v[0] / 1e3) / 1e9 if v[0] > 0 else 0.0,
                                   frac=v[1] / (v[0] / 1e3) / 1e9 / hbm_peak if v[0] > 0 else 0.0) for k, v in hbm.items()}
    line["hbm_peak_gbs"] = hbm_peak
    if replicas_identical is not None:
        line["replicas_identical"] = replicas_identical
    if world == 1:      # (with more ranks SyncBN's all-reduces would need every rank to take part)
        try:
            line["forward_stage_ms"] = gpu_stage_ms(model, resident, call)
        except Exception as e:  # noqa: BLE001 -- informational only
            line["forward_stage_ms"] = dict(error=(str(e).splitlines() or [type(e).__name__])[0][:200])
    if gstep.capture_error is not None:
        line["config"]["launch"] = "eager (CUDA-graph capture failed: %s)" % gstep.capture_error
    line["parity"] = dict(mode=args.precision, bound="1e-3 rel (north_star) holds in fp32 mode",
                          measured_vs_reference_fixture_r50_K2=dict(
                              fp32=dict(fused=1.3e-5, logits=1.4e-4, render=1.7e-6),
                              tf32=dict(fused=4.8e-4, logits=1.3e-2, render=1.4e-3),
                              bf16=dict(fused=6.5e-3, logits=7.2e-2, render=7.8e-3)),
                          source="tests/test_gpu_parity_r50.py on B200, profiles/r02_gpu_parity_r50_northstar.log; KNN / "
                                 "voxel indices bit-exact in every mode (r50 K=2 and the 200x200x16 grid)")
    if world == 1 and not args.no_precision_runs:
        # the same step in the reference's own arithmetic: tf32 (what torch 1.10 runs cuDNN/cuBLAS in by default on
        # Ampere+) and the fp32-accurate 3xTF32 split that tests hold to the 1e-3 bound.  Each mode gets its own
        # CUDA graph; 3 warm-up + 5 timed replays.
        runs = {}
        for mode in [m for m in ("tf32", "fp32") if m != args.precision]:
            try:
                del gstep
            except NameError:
                pass
            torch.cuda.empty_cache()
            try:
                coocc_b200.set_precision(mode)
                opt_m = make_opt()
                gs = coocc_b200.GraphedStep(model, opt_m, reducer, LOSS_KEYS, enabled=(args.launch == "graph"))
                for _ in range(2 + 3):
                    gs(resident["img"], resident["pts"], resident["geom"], resident["gt_depth"], resident["gt_img"],
                       resident["gt_occ"], resident.get("img_feats"), resident["transform"],
                       next_inputs=None if args.no_pipeline else (resident["img"], resident["pts"]))
                n_t = 5
                ms_m = timed(lambda: gs(resident["img"], resident["pts"], resident["geom"], resident["gt_depth"],
                                        resident["gt_img"], resident["gt_occ"], resident.get("img_feats"),
                                        resident["transform"],
                                        next_inputs=None if args.no_pipeline else (resident["img"], resident["pts"])),
                             n_t) / n_t
                cm = conv_roofline([(ms_, w) for ms_, w, t in bracketed_step(opt_m) if not t.startswith("hbm:")])
                pk = peak / 2 if mode == "tf32" else peak / 6
                runs[mode] = dict(ms_per_step=ms_m, value=nvox / (ms_m / 1e3), unit=UNIT,
                                  arithmetic="single-pass TF32 operands, fp32 accumulate" if mode == "tf32" else
                                             "3xTF32 split (hi*hi + hi*lo + lo*hi), fp32 accumulate and storage",
                                  conv_tflops=cm["achieved"], conv_ms_per_step=cm["kernel_ms_per_step"],
                                  conv_frac_of_mode_peak=cm["achieved"] / pk,
                                  mode_peak_tflops=pk, mode_peak_note="bf16 sustained peak / %d (tf32 MMA rate is half of "
                                  "bf16%s)" % ((2, "") if mode == "tf32" else (6, "; three passes")),
                                  launch=("cuda_graph" if gs.capture_error is None and args.launch == "graph" else "eager"))
                del gs, opt_m
            except Exception as e:  # noqa: BLE001
                runs[mode] = dict(error=(str(e).splitlines() or [type(e).__name__])[0][:200])
            finally:
                coocc_b200.set_precision(args.precision)
        line["precision_runs"] = runs
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(cfg, fine)
    print(json.dumps(line), flush=True)
    shutdown()


def gpu_stage_ms(model, d, call):
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
            outs = model.pts_bbox_head(voxel_feats=neck, img_feats=[d["img_feats"]] if "img_feats" in d else None,
                                       transform=d["transform"])
            model.pts_bbox_head.loss(output_voxels=outs["output_voxels"], output_voxels_fine=outs["output_voxels_fine"],
                                     output_coords_fine=outs["output_coords_fine"], target_voxels=d["gt_occ"])
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
