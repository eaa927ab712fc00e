"""Whole-step sessions on top of the engine: the hot loops of the reference trainer with every
launch of a step captured in one CUDA graph.

  * ``InferenceSession``  — ``val_epoch_`` core (trainer/trainer.py:209-220): forward in eval mode +
    heat-map arg-max keypoints (tools/misc/heatmap.py:173-178), host tensors in, keypoints out.
  * ``FusedTrainStep``    — ``train_epoch_`` core (trainer/trainer.py:115-136): zero_grad, forward,
    mean-of-three-heads FocalLoss_BCE_2d (the trainer's heatmap_criterion, trainer.py:426) or nn.MSELoss
    (BASELINE.json's configurations) with its gradient fused into the head backward, backward, [data-parallel all-reduce], the reference's AdamW (tools/optimizers/
    adamw.py:38-100) over one flat parameter buffer.  Where the reference makes three D2H copies, a
    CPU loss and one H2D gradient copy per step (trainer.py:127-135), nothing leaves the device here.

Data parallelism: one process per GPU, full replica each, local batch =
global / world; the only exchange is ONE ``all_reduce(SUM)`` of the flat 553 260-element fp32
gradient buffer per step (NCCL over NVLink; gloo in the CPU tests of the host logic), followed by
the same AdamW update on every rank with ``grad_scale = 1/world``.  BatchNorm statistics stay
per-rank like ``nn.DataParallel`` (trainer.py:338) keeps them per replica.
"""
from __future__ import annotations

from typing import Tuple

import torch

from . import ops, parallel
from .training import TrainState, backward_train, flat_layout, forward_train, pack_train


def _device(device) -> torch.device:
    d = torch.device(device)
    return torch.device("cuda", torch.cuda.current_device()) if d.type == "cuda" and d.index is None else d


def _engine(model, device):
    return model._engine(_device(device))


class InferenceSession:
    """Fixed-shape inference + keypoint extraction.  ``run(x_host)`` copies the batch to the device,
    replays the captured forward + arg-max and returns ``(xy int32 [B,C,2], peak fp32 [B,C])`` on the
    device; ``heat`` holds the heat maps of the selected head.  ``run_many`` streams a sequence of
    host batches through two input buffers so that the H2D copy of batch k+1 overlaps the kernels
    of batch k (the activation arena is shared: compute is serial on one stream).

    ``head``: one head index (results as above), or a tuple such as ``(0, 1, 2)`` — everything ``UNet_Nested.forward`` returns
    (models/unet.py:300) and the validation loop consumes (trainer/trainer.py:212-221): results are then
    ``xy [len(head),B,C,2]`` / ``peak [len(head),B,C]`` / ``heat [len(head),B,C,H,W]``, one arg-max launch for all.
    ``input``: "float32" (the reference's [B,C,H,W] fp32 tensors), or "uint8_nchw" / "uint8_nhwc": 8-bit images copied as bytes
    (4x less host-to-device traffic) and scaled by 1/255 on the device like torchvision's ToTensor, which the reference applies on
    the CPU before its H2D copy (datasets/datasets_base.py:71-72, trainer/trainer.py:109)."""

    def __init__(self, model, B: int, H: int, W: int, head=2, device="cuda", use_graph: bool = True, input: str = "float32"):
        if input not in ("float32", "uint8_nchw", "uint8_nhwc"):
            raise ValueError("input must be 'float32', 'uint8_nchw' or 'uint8_nhwc'")
        self.model, self.head, self.input = model, head, input
        self.heads = tuple(head) if isinstance(head, (tuple, list)) else (int(head),)
        if not self.heads or any(k not in (0, 1, 2) for k in self.heads) or list(self.heads) != sorted(set(self.heads)):
            raise ValueError("head must be 0, 1, 2 or an increasing tuple of them")
        self.dev = _device(device)
        self.eng = _engine(model, self.dev)
        if model.training:
            raise RuntimeError("InferenceSession needs model.eval()")
        self.B = B
        self.use_graph = use_graph
        shape = (B, H, W, model.in_channels) if input == "uint8_nhwc" else (B, model.in_channels, H, W)
        self.xs = [torch.zeros(shape, dtype=torch.float32 if input == "float32" else torch.uint8, device=self.dev) for _ in range(2)]
        self.H, self.W = H, W
        self.out = [None, None]
        self.graphs = [None, None]
        self._capture()
        self.copy_stream = torch.cuda.Stream(device=self.dev)

    def _capture(self):
        """(Re)build the captured passes.  The graphs hold raw pointers into the engine's activation arena and packed eval-mode
        weights: the session keeps both alive itself (the engine evicts arenas of other shapes and replaces the packed weights
        whenever a parameter changes) and remembers the parameter key the weights were folded from."""
        self.graphs = [None, None]
        with torch.no_grad(), torch.cuda.device(self.dev):
            n0 = ops.launch_count
            self._body(0)  # warm-up: packs weights, allocates the activation arena
            self.launches = ops.launch_count - n0 - self._pack_launches
            torch.cuda.synchronize()
            if self.use_graph:
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    self._body(0)
                torch.cuda.current_stream().wait_stream(s)
                for i in range(2):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        self._body(i)
                    self.graphs[i] = g
            else:
                self._body(1)
            self._key = self.eng._param_key()
            self._P = self.eng.packed_eval()
            self._arena = self.eng.arena(self.B, self.H, self.W, "eval4" if self._P["conv00.c1"].get("c4") else "eval")

    # kernels per pass: nchw->nhwc, 8 encoder convs, 3 pools, 6 x (deconv + 2 convs), arg-max = 31
    def _body(self, i: int):
        n0 = ops.launch_count
        P = self.eng.packed_eval()
        self._pack_launches = ops.launch_count - n0
        self.eng.forward_eval(self.xs[i], heads=self.heads, channels_last=self.input == "uint8_nhwc")
        heat = self.eng.last_heats  # [len(heads), B, C, H, W]
        xy, val = ops.argmax_peaks(heat.view(-1, *heat.shape[2:]))
        if isinstance(self.head, (tuple, list)):
            self.out[i] = (xy.view(len(self.heads), self.B, -1, 2), val.view(len(self.heads), self.B, -1), heat)
        else:
            self.out[i] = (xy, val, heat[0])

    @property
    def x(self):
        return self.xs[0]

    @property
    def graph(self):
        return self.graphs[0]

    @property
    def heat(self):
        return self.out[0][2]

    @property
    def xy(self):
        return self.out[0][0]

    @property
    def val(self):
        return self.out[0][1]

    def run_device(self, i: int = 0):
        """Inputs already in ``self.xs[i]`` (device): one pass of the hot path."""
        if self.eng._packed_key is None or self._key != self.eng._param_key():
            self._capture()  # a parameter or BatchNorm buffer changed since the weights were folded: re-pack and re-capture
        if self.graphs[i] is not None:
            self.graphs[i].replay()
        else:
            with torch.no_grad():
                self._body(i)
        return self.out[i][0], self.out[i][1]

    def run(self, x_host: torch.Tensor):
        self.xs[0].copy_(x_host, non_blocking=True)
        return self.run_device(0)

    def run_many(self, host_batches, xy_host=None, val_host=None):
        """Pipelined: for every (pinned) host batch, H2D -> forward + arg-max -> D2H of the keypoints.
        ``xy_host`` / ``val_host``: optional lists of pinned host tensors receiving the results.
        Returns after everything is enqueued; the caller synchronises the current stream."""
        cur = torch.cuda.current_stream(self.dev)
        cs = self.copy_stream
        cs.wait_stream(cur)
        h2d = [torch.cuda.Event(), torch.cuda.Event()]
        done = [None, None]
        n = len(host_batches)

        def prefetch(k):
            i = k & 1
            with torch.cuda.stream(cs):
                if done[i] is not None:
                    cs.wait_event(done[i])  # the graph that last read this buffer has finished
                self.xs[i].copy_(host_batches[k], non_blocking=True)
                h2d[i].record(cs)

        if n:
            prefetch(0)
        for k in range(n):
            i = k & 1
            if k + 1 < n:
                prefetch(k + 1)
            cur.wait_event(h2d[i])
            xy, val = self.run_device(i)
            if xy_host is not None:
                xy_host[k].copy_(xy, non_blocking=True)
            if val_host is not None:
                val_host[k].copy_(val, non_blocking=True)
            done[i] = torch.cuda.Event()
            done[i].record(cur)
        return n


class FusedTrainStep:
    def __init__(self, model, B: int, H: int, W: int, lr: float = 3e-6, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-4, device="cuda", use_graph: bool = True, process_group=None, seed: int = 0, loss: str = "focal",
                 focal_gamma: float = 3.0, optimizer: str = "adamw", momentum: float = 0.9, dampening: float = 0.0, final_lr: float = 0.1,
                 bound_gamma: float = 1e-3, input: str = "float32"):
        """Defaults follow the reference trainer: ``--optimizer adamw --lr 3e-6 --weight-decay 1e-4`` (train.py:38-45) and
        ``loss="focal"``, the criterion its training loop optimises — ``heatmap_criterion = FocalLoss_BCE_2d(gamma=3,
        size_average=False)`` (trainer.py:426, used at 125-135), mean over the three heads.  ``loss="mse"`` is the
        ``nn.MSELoss`` heat-map loss BASELINE.json's configurations name (the reference itself only uses MSE for the
        landmark criterion of validation, trainer.py:427,221).  In data-parallel runs every rank draws its own dropout
        masks (``seed + rank``), like nn.DataParallel's replicas draw from their devices' generators.
        ``optimizer``: any of the trainer's choices (trainer.py:344-376) — "adamw" (tools/optimizers/adamw.py), "adam" / "sgd"
        (torch.optim with the trainer's arguments; ``momentum`` = train.py:39), "sgdw" (tools/optimizers/sgdw.py as shipped, built by the
        trainer without momentum: pass ``momentum=0`` for that), "adabound" (tools/optimizers/adabound.py).  The learning rate lives
        on the device: ``set_lr`` is what an lr scheduler (MultiStepLR / ExponentialLR, trainer.py:383-388) calls between steps."""
        if input not in ("float32", "uint8_nchw", "uint8_nhwc"):
            raise ValueError("input must be 'float32', 'uint8_nchw' or 'uint8_nhwc'")
        self.model, self.input = model, input
        self.dev = _device(device)
        if not model.training:
            raise RuntimeError("FusedTrainStep needs model.train()")
        self.eng = _engine(model, self.dev)
        if optimizer not in ops.OPTIMIZER_KINDS:
            raise ValueError("optimizer must be one of %s" % sorted(ops.OPTIMIZER_KINDS))
        self.optimizer = optimizer
        if optimizer in ("sgd", "sgdw"):
            self.hyper = dict(lr=lr, beta1=momentum, beta2=dampening if optimizer == "sgdw" else 0.0, eps=eps, weight_decay=weight_decay)
        else:
            self.hyper = dict(lr=lr, beta1=betas[0], beta2=betas[1], eps=eps, weight_decay=weight_decay, final_lr=final_lr, gamma=bound_gamma, base_lr=lr)
        self.pg = process_group
        self.world, self.rank = 1, 0
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world, self.rank = torch.distributed.get_world_size(process_group), torch.distributed.get_rank(process_group)
        self.ts = TrainState(self.eng, B, H, W)
        ts = self.ts
        p_drop = float(model.drop_out.p)
        ts.use_masks, ts.drop_scale = p_drop > 0.0, (1.0 / (1.0 - p_drop) if p_drop > 0 else 1.0)
        self.p_drop, self.seed = p_drop, seed + self.rank  # every replica draws its own dropout masks (nn.DataParallel: one RNG stream per device)
        # flat fp32 parameter storage: every nn.Parameter becomes a view into one buffer (state_dict unchanged)
        lay, n = flat_layout(model)
        self.flat_p = torch.empty(n, dtype=torch.float32, device=self.dev)
        with torch.no_grad():
            for name, p in model.named_parameters():
                off, cnt = lay[name]
                self.flat_p[off:off + cnt].copy_(p.detach().reshape(-1))
                p.data = self.flat_p[off:off + cnt].view_as(p)
        self._versioned = list(model.parameters()) + [b for b in model.buffers() if b.dtype.is_floating_point]
        self.flat_g = torch.zeros(n, dtype=torch.float32, device=self.dev)
        self.flat_m = torch.zeros(n, dtype=torch.float32, device=self.dev)
        self.flat_v = torch.zeros(n, dtype=torch.float32, device=self.dev)
        self.step_counter = torch.zeros(1, dtype=torch.int64, device=self.dev)
        self.step_size = torch.zeros(4, dtype=torch.float32, device=self.dev)  # per-step optimizer scalars (unpp_optim_step)
        self.lr_dev = torch.full((1,), float(lr), dtype=torch.float32, device=self.dev)
        self.loss = torch.zeros(1, dtype=torch.float32, device=self.dev)
        # ``input``: "float32" (the reference's [B,C,H,W] tensors, trainer.py:109) or 8-bit images copied as bytes and scaled by 1/255 on
        # the device like torchvision's ToTensor (datasets/datasets_base.py:71-72): "uint8_nchw" / "uint8_nhwc" ([B,H,W,C], what PIL holds)
        xshape = (B, H, W, model.in_channels) if input == "uint8_nhwc" else (B, model.in_channels, H, W)
        self.x = torch.zeros(xshape, dtype=torch.float32 if input == "float32" else torch.uint8, device=self.dev)
        self.target = torch.zeros(B, model.n_classes, H, W, dtype=torch.float32, device=self.dev)
        self.numel_head = B * model.n_classes * H * W
        if loss == "mse":      # (1/3) sum_k mean((p_k - T)^2): nn.MSELoss per head (trainer.py:427), mean of the heads (trainer.py:125-134)
            self.loss_kind, self.coef, self.loss_scale = 0, 2.0 / (3.0 * self.numel_head), 1.0 / (3.0 * self.numel_head)
        elif loss == "focal":  # (1/3) sum_k FocalLoss_BCE_2d(gamma)(p_k, T): sum over pixels / (B*C) (focal_loss.py:282,301; trainer.py:426)
            self.loss_kind = 1
            self.coef = self.loss_scale = 1.0 / (3.0 * B * model.n_classes)
        else:
            raise ValueError("loss must be 'mse' or 'focal'")
        self.focal_gamma = focal_gamma
        self.graph_a = self.graph_b = None
        self._side_stream = torch.cuda.Stream(self.dev)
        self.steps_done = 0
        with torch.cuda.device(self.dev):
            snapshot = (self.flat_p.clone(), [b.clone() for b in model.buffers()])
            self._fwd_bwd()  # warm-up (allocates scratch); undone below
            self._update()
            torch.cuda.synchronize()
            if use_graph:
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    self._fwd_bwd()
                    self._update()
                torch.cuda.current_stream().wait_stream(s)
                ga = torch.cuda.CUDAGraph()
                with torch.cuda.graph(ga):
                    self._fwd_bwd()
                    if self.world == 1:
                        self._update()
                self.graph_a = ga
                if self.world > 1:
                    gb = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(gb):
                        self._update()
                    self.graph_b = gb
            # restore the state the warm-up / capture passes touched
            self.flat_p.copy_(snapshot[0])
            for b, b0 in zip(model.buffers(), snapshot[1]):
                b.copy_(b0)
            self.flat_m.zero_()
            self.flat_v.zero_()
            self.step_counter.zero_()
            torch.cuda.synchronize()

    # one step = these two pieces; every launch in them is a libunpp.so kernel except the int64
    # num_batches_tracked increments of the eight BatchNorm layers
    def _fwd_bwd(self):
        ts = self.ts
        # The head dropout masks are not needed before the first decoder head: they are generated on a side stream (a parallel
        # branch of the captured graph) next to the weight re-packing and the encoder; the CUDA-core mask kernel co-resides with
        # the one-CTA-per-SM tensor-core convolutions (no shared memory, few registers) instead of serialising 3 launches.
        cur = torch.cuda.current_stream(self.dev)
        side = self._side_stream
        side.wait_stream(cur)
        packed = torch.cuda.Event()
        with torch.cuda.stream(side):  # the weight re-packing runs next to the input layout conversion, the masks next to the encoder
            pack_train(ts)
            packed.record(side)
            if ts.use_masks:
                for k in range(3):
                    ops.dropout_mask(ts.t[f"mask{k}"].view(-1), self.p_drop, self.seed * 7919 + k, self.step_counter)
        forward_train(ts, self.x, after_input=lambda: cur.wait_event(packed), before_decoder=lambda: cur.wait_stream(side),
                      channels_last=self.input == "uint8_nhwc")
        backward_train(ts, self.flat_g, target=self.target, coef=self.coef, loss_kind=self.loss_kind, gamma=self.focal_gamma)
        nacc, ncls = ts.head_nacc, self.model.n_classes
        ops.reduce_partials(ts.t["head_red"], 3, nacc, 1, self.loss, scale=self.loss_scale, partial_offset=ncls * 17)

    def _update(self):
        ops.optim_step(self.optimizer, self.flat_p, self.flat_g, self.flat_m, self.flat_v, grad_scale=1.0 / self.world, step_counter=self.step_counter,
                       lr_dev=self.lr_dev, scalars=self.step_size, **self.hyper)

    def set_lr(self, lr: float) -> None:
        """New learning rate for the following steps (the captured graph reads it from device memory)."""
        self.hyper["lr"] = float(lr)
        self.lr_dev.fill_(float(lr))

    def step_device(self) -> torch.Tensor:
        """Inputs already in ``self.x`` / ``self.target``.  Returns the (device) loss of this rank's batch."""
        if self.graph_a is not None:
            self.graph_a.replay()
        else:
            self._fwd_bwd()
            if self.world == 1:
                self._update()
        if self.world > 1:
            ev = getattr(self, "allreduce_events", None)  # bench.py: a list that receives one (before, after) CUDA-event pair per step
            if ev is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            parallel.allreduce_gradients(self.flat_g, self.pg)  # the step's only collective: 2.2 MB flat fp32 buffer
            if ev is not None:
                e1.record()
                ev.append((e0, e1))
            if self.graph_b is not None:
                self.graph_b.replay()
            else:
                self._update()
        self.steps_done += 1
        # parameters (and BatchNorm running statistics) changed through raw pointers: tell every version-keyed cache
        torch._C._increment_version(self._versioned)
        self.eng._packed_key = None
        return self.loss

    def step(self, x_host: torch.Tensor, target_host: torch.Tensor) -> torch.Tensor:
        self.x.copy_(x_host, non_blocking=True)
        self.target.copy_(target_host, non_blocking=True)
        return self.step_device()

    def step_many(self, x_hosts, target_hosts, loss_hosts=None) -> int:
        """Pipelined training over a list of pinned host batches: the H2D copy of batch k+1 (into staging buffers, on a
        copy stream) overlaps the step of batch k; a device-to-device copy (~0.1 ms) moves the staged batch into the
        fixed-address inputs the captured step reads.  ``loss_hosts``: optional pinned [1] tensors receiving each loss.
        ``target_hosts``: heat-map targets [B,C,H,W] fp32 — or KEY POINTS [B,points,2] fp32, what the trainer's data loader
        actually delivers as labels (trainer.py:109): the targets of helper.create_heatmap (trainer.py:122-123, numpy on the CPU
        every iteration in the reference) are then synthesised on the device, and a step moves 56 bytes of labels per image.
        Returns after everything is enqueued; the caller synchronises the current stream."""
        cur = torch.cuda.current_stream(self.dev)
        kp_mode = n_kp = None
        if len(target_hosts):
            kp_mode = target_hosts[0].dim() == 3
            n_kp = target_hosts[0].shape[1] if kp_mode else None
        if getattr(self, "_stage_kind", None) != (kp_mode, n_kp):
            # staged images, staged targets and (key-point mode) the staged key points the targets are synthesised from
            self._stage = [(torch.empty_like(self.x), torch.empty_like(self.target),
                            torch.empty(self.x.shape[0], n_kp, 2, dtype=torch.float32, device=self.dev) if kp_mode else None) for _ in range(2)]
            self._stage_kind = (kp_mode, n_kp)
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream(self.dev)
        cs = self._copy_stream
        cs.wait_stream(cur)
        h2d = [torch.cuda.Event(), torch.cuda.Event()]
        taken = [None, None]
        n = len(x_hosts)

        def prefetch(k):
            i = k & 1
            with torch.cuda.stream(cs):
                if taken[i] is not None:
                    cs.wait_event(taken[i])  # the staged batch has been moved into the step's inputs
                self._stage[i][0].copy_(x_hosts[k], non_blocking=True)
                if kp_mode:  # the target synthesis of batch k+1 runs on the copy stream too, next to the step of batch k
                    self._stage[i][2].copy_(target_hosts[k], non_blocking=True)
                    ops.create_heatmap(self._stage[i][2], self.target.shape[2], self.target.shape[3], out=self._stage[i][1])
                else:
                    self._stage[i][1].copy_(target_hosts[k], non_blocking=True)
                h2d[i].record(cs)

        if n:
            prefetch(0)
        for k in range(n):
            i = k & 1
            cur.wait_event(h2d[i])
            self.x.copy_(self._stage[i][0], non_blocking=True)
            self.target.copy_(self._stage[i][1], non_blocking=True)
            taken[i] = torch.cuda.Event()
            taken[i].record(cur)
            if k + 1 < n:
                prefetch(k + 1)
            loss = self.step_device()
            if loss_hosts is not None:
                loss_hosts[k].copy_(loss, non_blocking=True)
        return n

    def step_keypoints(self, x_host: torch.Tensor, keypoints_host: torch.Tensor) -> torch.Tensor:
        """The trainer's step with its target synthesis on the device: where trainer/trainer.py:122-123 builds the
        heat-map targets on the CPU with numpy every iteration (helper.create_heatmap), the key points [B,7,2] go
        H2D (56 bytes per image) and ``unpp_create_heatmap`` writes the targets straight into the step's buffer."""
        self.x.copy_(x_host, non_blocking=True)
        kp = keypoints_host.to(self.dev, non_blocking=True).float()
        ops.create_heatmap(kp, self.target.shape[2], self.target.shape[3], out=self.target)
        return self.step_device()

    @property
    def heats(self):
        return self.ts.heats

    # ------------------------------------------------------------------------------ checkpoint / resume (trainer/trainer.py:402-413)
    def optimizer_state_dict(self) -> dict:
        """The flat optimizer state in the layout ``torch.optim.Optimizer.state_dict()`` gives the reference's optimizers
        (tools/optimizers/adamw.py:60-68: per parameter ``step`` / ``exp_avg`` / ``exp_avg_sq``; SGD-family: ``momentum_buffer``), so a
        ``.tar`` written here resumes under the reference trainer (``optimizer.load_state_dict(checkpoint['optimizer_state_dict'])``,
        trainer.py:411) and a ``.tar`` written by it resumes here."""
        step = int(self.step_counter.item())
        adam = self.optimizer in ("adamw", "adam", "adabound")
        state = {}
        for i, (name, p) in enumerate(self.model.named_parameters()):
            off, n = self.ts.lay[name]
            if adam:
                state[i] = {"step": step, "exp_avg": self.flat_m[off:off + n].view_as(p).clone(), "exp_avg_sq": self.flat_v[off:off + n].view_as(p).clone()}
            elif self.hyper["beta1"] != 0 and step > 0:
                state[i] = {"momentum_buffer": self.flat_m[off:off + n].view_as(p).clone()}
        h = self.hyper
        if adam:
            group = {"lr": h["lr"], "betas": (h["beta1"], h["beta2"]), "eps": h["eps"], "weight_decay": h["weight_decay"], "amsgrad": False}
            if self.optimizer == "adabound":
                group.update(final_lr=h["final_lr"], gamma=h["gamma"], amsbound=False)
                del group["amsgrad"]
        else:
            group = {"lr": h["lr"], "momentum": h["beta1"], "dampening": h["beta2"], "weight_decay": h["weight_decay"], "nesterov": False}
        group["params"] = list(range(len(self.ts.lay)))
        # (torch's Optimizer.load_state_dict ignores extra top-level keys: the step count also seeds the dropout stream, and SGD's state has none)
        return {"state": state, "param_groups": [group], "unpp_step": step}

    def load_optimizer_state_dict(self, sd: dict) -> None:
        """Inverse of ``optimizer_state_dict`` (also takes what the reference's / torch's optimizers saved for this model)."""
        names = [k for k, _ in self.model.named_parameters()]
        state = sd["state"]
        step = 0
        with torch.no_grad():
            self.flat_m.zero_()
            self.flat_v.zero_()
            for i, name in enumerate(names):
                st = state.get(i, state.get(str(i)))
                if not st:
                    continue
                off, n = self.ts.lay[name]
                if "exp_avg" in st:
                    self.flat_m[off:off + n].copy_(st["exp_avg"].reshape(-1))
                    self.flat_v[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
                if "momentum_buffer" in st and st["momentum_buffer"] is not None:
                    self.flat_m[off:off + n].copy_(st["momentum_buffer"].reshape(-1))
                    step = max(step, 1)  # the buffer exists: not the first step any more
                step = max(step, int(st.get("step", 0)))
            self.step_counter.fill_(int(sd.get("unpp_step", step)))
        if sd.get("param_groups"):
            self.set_lr(float(sd["param_groups"][0]["lr"]))

    def save_checkpoint(self, path: str, epoch: int) -> None:
        """A ``.tar`` in the format the reference trainer resumes from (trainer.py:402-413): ``model_state_dict`` (the unchanged
        98-key layout), ``optimizer_state_dict`` and ``epoch``."""
        torch.save({"epoch": int(epoch), "model_state_dict": self.model.state_dict(), "optimizer_state_dict": self.optimizer_state_dict()}, path)

    def load_checkpoint(self, path: str, resume_opt: bool = True) -> int:
        """Resume like trainer.py:402-413: weights always, optimizer state and epoch with ``resume_opt``.  Returns the epoch to start from.
        The parameters stay views of the flat buffer (``load_state_dict`` copies in place)."""
        ck = torch.load(path, map_location=self.dev)
        self.model.load_state_dict(ck["model_state_dict"])
        torch._C._increment_version(self._versioned)
        self.eng._packed_key = None
        if resume_opt:
            self.load_optimizer_state_dict(ck["optimizer_state_dict"])
            return int(ck["epoch"]) + 1
        return 0


broadcast_parameters = parallel.broadcast_parameters
