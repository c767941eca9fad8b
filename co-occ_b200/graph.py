"""Whole-step CUDA graphs for the hot path.

One training step of the path is ~800 kernel launches of 10-1000 us each; issued one by one from
Python the host cannot keep ahead of a B200 (measured: 57 ms of device work took 63-75 ms of wall
time).  GraphedStep captures forward + backward + gradient all-reduce + optimizer update into one
CUDA graph and replays it, so the step costs one launch.

What stays outside the graph is the only data-dependent host decision of the path: after
`pack + compact` (functional.gsf_prologue, run eagerly into fixed buffers) the occupied-voxel counts
N_img / N_pts are read back, because they select the reference's branches (bifuser_n.py:55 vs :62,
SURVEY Q1/Q2/Q11) and size the FPS cluster.  Graphs are keyed by the counts rounded up to
`bucket`; every kernel reads the exact counts from device memory, so one graph serves all scenes of
its bucket.  Scenes that take a reference branch the graphs do not cover (N <= 2048, N_pts > N_img
with K > 1) run the ordinary eager path.

Data parallel: the gradient all-reduce (ddp.GradReducer) and the SyncBN statistics all-reduces are
NCCL calls on the capturing stream and become graph nodes.  NCCL's watchdog thread must not poll
events of a capturing stream: set TORCH_NCCL_ASYNC_ERROR_HANDLING=0 before init_process_group
(bench.py does), as PyTorch's CUDA-graph notes require for whole-iteration capture with DDP.
"""
import os
import sys

import torch

from . import functional as CF


def _dbg(msg):
    if os.environ.get("COOCC_DEBUG"):
        print("[coocc_b200.graph rank %s] %s" % (os.environ.get("RANK", "0"), msg), file=sys.stderr, flush=True)


class GraphedStep:
    """step(img, pts, geom, gt_depth, gt_img, gt_occ) -> loss (a 0-dim tensor that the next
    replay overwrites).  Gradients live in the buffers of the graph that produced them: with an optimizer inside the
    graph that is invisible; without one (optimizer=None) `p.grad` refers to the most recently CAPTURED graph, so read
    gradients only while a single graph exists.

    model      HotPath (train mode)
    optimizer  torch optimizer constructed with capturable=True (or None: forward+backward only)
    reducer    ddp.GradReducer or None
    loss_keys  which entries of forward_train's loss dict are summed (None = all)
    """

    def __init__(self, model, optimizer=None, reducer=None, loss_keys=None, bucket=8192, max_graphs=4,
                 enabled=True, pipeline_index=None):
        self.model, self.opt, self.reducer = model, optimizer, reducer
        self.loss_keys = loss_keys
        self.bucket, self.max_graphs = int(bucket), int(max_graphs)
        self.enabled = enabled
        self.graphs = {}            # key -> dict(graph, loss, static inputs, launches)
        self.prologue = None        # fixed-address pack/compact buffers
        self.static = None          # fixed-address copies of the small inputs
        self.stats = dict(replays=0, eager=0, captures=0)
        self.launches_per_replay = 0
        self.capture_error = None   # set if a capture failed and the runner fell back to eager launches
        self._stream = None         # capture stream; the warm-up step runs on it too, so that autograd's
        self._warm = False
        self._last = None           # AccumulateGrad nodes live on the stream the capture uses
        # Index pipelining (step(..., next_inputs=...)): the neighbour search of BiFuser_N (FPS + top-K + ball assignment,
        # 3.2 ms, 2047 dependent FPS rounds on 32 SMs) depends on the inputs' occupancy only, so the tables of step i+1
        # are computed on a side branch of step i's graph, forked at the start of the fuser's backward -- the point
        # behind which only HBM-bound kernels are left (GSFusion backward, gradient-reduction tail, optimizer).  Two
        # sets of pack / table buffers alternate; a step whose tables were not prepared computes them up front.
        if pipeline_index is None:
            pipeline_index = os.environ.get("COOCC_PIPELINE_INDEX", "1") != "0"
        self.pipeline_index = bool(pipeline_index)
        self._pro = [None, None]
        self._tables = [None, None]
        self._par = 0
        self._ahead = None          # dict(img, pts, n_img, n_pts, par): what the last step prepared
        self._side = None
        # fork before the last convolution of the backward (its grid shrunk to the SMs the index branch leaves free)
        # instead of behind it: hides the whole 3.2 ms instead of the 1.8 ms of the non-conv tail
        self.early_fork = os.environ.get("COOCC_PIPELINE_EARLY", "1") != "0"
        self.dynamic_tiles = os.environ.get("COOCC_PIPELINE_DYNAMIC", "1") != "0"
        # where the branch forks: "dgrad" = before the last convolution of the backward (big grids: that convolution
        # alone is as long as the FPS), "fuser" = at the start of the fuser's backward (the reference's own grids, where
        # the FPS is as long as the fuser's whole backward), "auto" = by grid size
        self.fork_at = os.environ.get("COOCC_PIPELINE_FORK", "auto")
        self._num_sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count \
            if torch.cuda.is_available() else 148
        self._tr_fixed, self._tr_len = None, 0

    # ------------------------------------------------------------------------------------
    def _loss(self, losses):
        keys = self.loss_keys if self.loss_keys is not None else list(losses)
        return sum(losses[k] for k in keys)

    def _eager(self, img, pts, geom, gt_depth, gt_img, gt_occ, img_feats=None, transform=None):
        if self.opt is not None:
            self.opt.zero_grad(set_to_none=True)
        else:
            self.model.zero_grad(set_to_none=True)
        if self.reducer is not None:
            self.reducer.begin()
        CF.begin_step()
        CF.zero_pool_begin(img.device)
        try:
            losses, _, _ = self.model.forward_train(img, pts, geom, gt_depth, gt_img, gt_occ, img_feats, transform)
            loss = self._loss(losses)
            loss.backward()
        finally:
            CF.zero_pool_end()
        if self.reducer is not None:
            self.reducer.finish()
        if self.opt is not None:
            self.opt.step()
        return loss.detach()

    def _round(self, n, cap):
        return min(cap, (n + self.bucket - 1) // self.bucket * self.bucket)

    def _bind_static(self, geom, gt_depth, gt_img, gt_occ, img_feats=None, transform=None):
        new = dict(geom=geom, gt_depth=gt_depth, gt_img=gt_img, gt_occ=gt_occ, img_feats=img_feats)
        # transform = img_inputs[1:]: the calibration tensors get fixed-address copies, everything else (None
        # placeholders, the (H, W) image size -- read on the host) is passed through as python values
        tr_fixed = None
        if transform is not None:
            for i, t in enumerate(transform):
                if torch.is_tensor(t) and t.is_cuda:
                    new["tr%d" % i] = t
            size = transform[-1]
            tr_fixed = tuple(float(v[0]) if hasattr(v, "__getitem__") else float(v) for v in size)
        if self.static is not None and self._tr_fixed != tr_fixed:
            raise RuntimeError("GraphedStep: image size changed; build a new GraphedStep")
        self._tr_fixed = tr_fixed
        self._tr_len = len(transform) if transform is not None else 0
        if self.static is None:
            self.static = {k: (v.clone() if v is not None else None) for k, v in new.items()}
            return
        for k, v in new.items():
            s = self.static[k]
            if (s is None) != (v is None) or (v is not None and (s.shape != v.shape or s.dtype != v.dtype)):
                raise RuntimeError("GraphedStep: input %r changed shape/dtype; build a new GraphedStep" % k)
            if v is not None:
                s.copy_(v, non_blocking=True)

    # ------------------------------------------------------------------------------------
    def _static_transform(self):
        if not self._tr_len:
            return None
        tr = [self.static.get("tr%d" % i) for i in range(self._tr_len)]
        tr[-1] = self._tr_fixed
        return tuple(tr)

    def __call__(self, img, pts, geom, gt_depth, gt_img, gt_occ=None, img_feats=None, transform=None,
                 next_inputs=None):
        """next_inputs = (img_voxel_feats, pts_voxel_feats) of the FOLLOWING step (device tensors, e.g. from a
        prefetching loader): enables index pipelining; pass the very same tensor objects as img / pts next time."""
        args = (img, pts, geom, gt_depth, gt_img, gt_occ, img_feats, transform)
        if not self.enabled:
            self.stats["eager"] += 1
            return self._eager(*args)
        if self.pipeline_index and next_inputs is not None and self._warm and \
                not (self.opt is not None and len(self.opt.state) == 0):
            return self._call_pipelined(args, next_inputs)
        self._ahead = None
        K = self.model.occ_fuser.knum
        # optimizer state must exist before a capture (lazy state init allocates and syncs)
        if self._stream is None:
            self._stream = torch.cuda.Stream()
        if not self._warm or (self.opt is not None and len(self.opt.state) == 0):
            # first call: one ordinary step on the capture stream (lazy CUDA / optimizer-state
            # initialisation allocates and synchronises, which a capture cannot)
            self._warm = True
            self.stats["eager"] += 1
            _dbg("warm-up step (eager)")
            cur = torch.cuda.current_stream()
            self._stream.wait_stream(cur)
            with torch.cuda.stream(self._stream):
                loss = self._eager(*args)
            cur.wait_stream(self._stream)
            return loss
        self.prologue = CF.gsf_prologue(img, pts, out=self.prologue)
        n_img, n_pts = (int(v) for v in self.prologue["counts"].tolist())     # the step's one host sync
        X, Y, Z = self.prologue["dims"]
        V = X * Y * Z
        covered = (n_img > CF.FPS_NUM and n_pts > CF.FPS_NUM and not (K > 1 and n_pts > n_img)
                   and (K > 1 or self.model.occ_fuser.fix_k1_fps))
        if not covered:
            self.stats["eager"] += 1
            return self._eager(*args)
        nb_img, nb_pts = self._round(n_img, V), self._round(n_pts, V)
        key = (nb_img, nb_pts)
        self._bind_static(geom, gt_depth, gt_img, gt_occ, img_feats, transform)
        ov = dict(prologue=self.prologue, n_img=n_img, n_pts=n_pts, nb_img=nb_img, nb_pts=nb_pts)
        entry = self.graphs.get(key)
        if entry is None:
            if len(self.graphs) >= self.max_graphs:
                self.stats["eager"] += 1
                return self._eager(*args)
            try:
                entry = self._capture(key, ov, img, pts)
            except Exception as e:  # noqa: BLE001 -- a failed capture must not take the training run down
                # (the eager path computes the same step; the failure is kept for the caller to report)
                self.enabled = False
                self.capture_error = "%s: %s" % (type(e).__name__, (str(e).splitlines() or [""])[0][:200])
                _dbg("capture failed, falling back to eager: " + self.capture_error)
                torch.cuda.synchronize()
                self.stats["eager"] += 1
                return self._eager(*args)
        entry["graph"].replay()
        self.stats["replays"] += 1
        if self.stats["replays"] <= 2 and os.environ.get("COOCC_DEBUG"):
            torch.cuda.synchronize()
            _dbg("replay %d done, loss %.6f" % (self.stats["replays"], float(entry["loss"])))
        self._last = entry
        return entry["loss"]

    # ------------------------------------------------------------------------------------
    def _covered(self, n_img, n_pts, K):
        return (n_img > CF.FPS_NUM and n_pts > CF.FPS_NUM and not (K > 1 and n_pts > n_img)
                and (K > 1 or self.model.occ_fuser.fix_k1_fps))

    def _alloc_tables(self, K, V, dev):
        mk = lambda *shape: torch.empty(*shape, device=dev, dtype=torch.int32)
        return {name: dict(rep_idx=mk(CF.FPS_NUM), nrep=CF.FPS_NUM, topk_idx=mk(CF.FPS_NUM, K), topk_d2=mk(CF.FPS_NUM, K),
                           winner=mk(K, V), group=None) for name in ("A", "B")}

    @staticmethod
    def _jobs(nb_img, nb_pts):
        # functional.gsf_index: direction A = LiDAR queries / image keys, B = image queries / LiDAR keys
        return [("A", 1, 0, nb_pts), ("B", 0, 1, nb_img)]

    def _call_pipelined(self, args, next_inputs):
        img, pts, geom, gt_depth, gt_img, gt_occ, img_feats, transform = args
        K = self.model.occ_fuser.knum
        p = self._par
        ah, self._ahead = self._ahead, None
        prepared = ah is not None and ah["img"] is img and ah["pts"] is pts and ah["par"] == p
        if prepared:
            n_img, n_pts = ah["n_img"], ah["n_pts"]
        else:
            self._pro[p] = CF.gsf_prologue(img, pts, out=self._pro[p])
            n_img, n_pts = (int(v) for v in self._pro[p]["counts"].tolist())
        if not self._covered(n_img, n_pts, K):
            self.stats["eager"] += 1
            return self._eager(*args)
        X, Y, Z = self._pro[p]["dims"]
        V = X * Y * Z
        nb_img, nb_pts = self._round(n_img, V), self._round(n_pts, V)
        if self._tables[p] is None:
            self._tables[p] = self._alloc_tables(K, V, img.device)
        if self._tables[1 - p] is None:
            self._tables[1 - p] = self._alloc_tables(K, V, img.device)
        if not prepared:
            CF.gsf_index_tables(self._pro[p], self._jobs(nb_img, nb_pts), K, out=self._tables[p])
        # the following step's inputs: pack + compact now (as every step does), tables inside this step's graph
        nimg, npts = next_inputs
        self._pro[1 - p] = CF.gsf_prologue(nimg, npts, out=self._pro[1 - p])
        m_img, m_pts = (int(v) for v in self._pro[1 - p]["counts"].tolist())      # the step's one host sync
        nxt = None
        if self._covered(m_img, m_pts, K):
            mb_img, mb_pts = self._round(m_img, V), self._round(m_pts, V)
            nxt = dict(prologue=self._pro[1 - p], jobs=self._jobs(mb_img, mb_pts), out=self._tables[1 - p], K=K)
        key = (nb_img, nb_pts, (mb_img, mb_pts) if nxt is not None else None, p)
        self._bind_static(geom, gt_depth, gt_img, gt_occ, img_feats, transform)
        ov = dict(prologue=self._pro[p], n_img=n_img, n_pts=n_pts, nb_img=nb_img, nb_pts=nb_pts, tables=self._tables[p])
        entry = self.graphs.get(key)
        if entry is None:
            if len(self.graphs) >= self.max_graphs:
                self.stats["eager"] += 1
                return self._eager(*args)
            try:
                entry = self._capture(key, ov, img, pts, nxt)
            except Exception as e:  # noqa: BLE001 -- same policy as the plain path
                self.enabled = False
                self.capture_error = "%s: %s" % (type(e).__name__, (str(e).splitlines() or [""])[0][:200])
                _dbg("capture failed, falling back to eager: " + self.capture_error)
                torch.cuda.synchronize()
                self.stats["eager"] += 1
                return self._eager(*args)
        entry["graph"].replay()
        self.stats["replays"] += 1
        self.stats["pipelined"] = self.stats.get("pipelined", 0) + (1 if prepared else 0)
        self._last = entry
        if nxt is not None:
            self._ahead = dict(img=nimg, pts=npts, n_img=m_img, n_pts=m_pts, par=1 - p)
        self._par = 1 - p
        return entry["loss"]

    def check(self):
        """Raise what the eager path would have raised from device-side error flags of the last
        replayed step (synchronises; the reference's IndexError cases)."""
        entry = self._last
        if entry is None:
            return
        for flag, exc in entry["deferred"]:
            if int(flag.item()) != 0:
                raise exc

    def _capture(self, key, ov, img, pts, nxt=None):
        from . import _lib
        s = self.static
        if self.opt is not None:
            self.opt.zero_grad(set_to_none=True)
        else:
            self.model.zero_grad(set_to_none=True)
        if self.reducer is not None:
            self.reducer.begin()
        CF.begin_step()
        g = torch.cuda.CUDAGraph()
        n0 = _lib.CALLS["n"]
        CF.GSF_OVERRIDE = ov
        CF.DEFERRED_ERRORS = deferred = []
        fired = [False]
        lib = _lib.lib()
        # data parallel: NCCL's all-reduce CTAs hold SMs while the backward's convolutions launch; with the static
        # round-robin a persistent grid then runs the CTAs of those SMs as a second wave.  The long-tile launches of
        # the captured step use the dynamic tile scheduler instead (mode 2; COOCC_DDP_DYNAMIC=0 disables).
        base_dyn = 2 if (self.reducer is not None and getattr(self.reducer, "world", 1) > 1
                         and os.environ.get("COOCC_DDP_DYNAMIC", "1") != "0") else 0
        lib.coocc_conv_set_dynamic(base_dyn)
        if nxt is not None:
            if self._side is None:
                self._side = torch.cuda.Stream(priority=-1)     # its clusters are placed before the next conv's CTAs
            side = self._side

            def fork(gate=False):
                if fired[0]:
                    return
                fired[0] = True
                side.wait_stream(torch.cuda.current_stream())
                lib.coocc_gsf_fps_signal(1 if gate else 0)
                try:
                    with torch.cuda.stream(side):
                        CF.gsf_index_tables(nxt["prologue"], nxt["jobs"], nxt["K"], out=nxt["out"])
                finally:
                    lib.coocc_gsf_fps_signal(0)
                if gate:
                    # hold the next kernel of this stream back until both FPS clusters are resident (csrc/gsf_index.cu)
                    _lib.check(lib.coocc_gsf_fps_gate(len(nxt["jobs"]), CF._stream()), "gsf_fps_gate")

            def share_sms():
                # The index branch takes 32 SMs (two 16-CTA FPS clusters).  The convolutions that run next to it are
                # captured with the dynamic tile scheduler (CTAs that get their SM late -- when the clusters are done
                # -- take the tiles that are left), or, COOCC_PIPELINE_DYNAMIC=0, with a static grid on the other SMs;
                # a plain 148-CTA grid would run its last 32 CTAs as a second wave.
                if self.dynamic_tiles:
                    lib.coocc_conv_set_dynamic(1)
                else:
                    lib.coocc_conv_set_sm_budget(self._num_sms - 32)

            def pre_tail_hook(where):
                # (autograd thread.)  where = "dgrad": in front of the data gradient of the fuser's first convolution,
                # the last convolution of the backward; "fuser": at the start of the fuser's con_enc backward
                if self.early_fork and where == self._fork_at and not fired[0]:
                    fork(gate=True)
                    share_sms()

            def tail_hook():
                # (start of the fuser's GSFusion backward: only HBM-bound kernels from here on)
                lib.coocc_conv_set_sm_budget(0)
                lib.coocc_conv_set_dynamic(base_dyn)
                fork()
            V_ = 1
            for n_ in ov["prologue"]["dims"]:
                V_ *= n_
            self._fork_at = self.fork_at if self.fork_at != "auto" else ("dgrad" if V_ >= 320000 else "fuser")
            CF.TAIL_HOOK = tail_hook
            CF.PRE_TAIL_HOOK = pre_tail_hook
        try:
            _dbg("capture begin key=%s" % (key,))
            # thread_local: NCCL's watchdog / heartbeat threads query events while this thread captures
            with torch.cuda.graph(g, stream=self._stream, capture_error_mode="thread_local"):
                CF.zero_pool_begin(img.device)          # (a memset node: every replay starts from a clean pool)
                losses, _, _ = self.model.forward_train(img, pts, s["geom"], s["gt_depth"], s["gt_img"], s["gt_occ"],
                                                        s["img_feats"], self._static_transform())
                loss = self._loss(losses)
                loss.backward()
                CF.zero_pool_end()
                if self.reducer is not None:
                    self.reducer.finish()
                if self.opt is not None:
                    self.opt.step()
                loss = loss.detach()
                if nxt is not None:
                    if fired[0]:
                        torch.cuda.current_stream().wait_stream(self._side)     # join the side branch
                    else:           # the fuser's backward did not run (inputs without gradient): no overlap, same result
                        CF.gsf_index_tables(nxt["prologue"], nxt["jobs"], nxt["K"], out=nxt["out"])
        finally:
            CF.TAIL_HOOK = None
            CF.PRE_TAIL_HOOK = None
            lib.coocc_conv_set_sm_budget(0)
            lib.coocc_conv_set_dynamic(0)
            CF.GSF_OVERRIDE = None
            CF.DEFERRED_ERRORS = None
            CF.zero_pool_end()
        _dbg("capture end")
        entry = dict(graph=g, loss=loss, launches=_lib.CALLS["n"] - n0, deferred=deferred)
        self.launches_per_replay = entry["launches"]
        self.graphs[key] = entry
        self.stats["captures"] += 1
        return entry
