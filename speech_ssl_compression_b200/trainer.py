"""Training-step driver: flat parameter storage, the fused Adam step and an optional
CUDA-graph capture of one whole optimizer step (forward + backward + gradient all-reduce +
clip + Adam), which removes Python / launch overhead from the hot loop.

Mirrors what reference ``runner.py:357-427`` does per optimizer step: forward -> loss / accum
-> backward -> grads /= n -> clip_grad_norm_ -> Adam.step -> zero_grad.
"""
import os

import numpy as np
import torch

from . import kernels as K
from . import ops
from .parallel import DataParallelB200, FlatBuffers


class FusedAdam:
    """Adam over the flat fp32 buffers in one HBM pass (``mh_adam_step``): grad scaling,
    global-norm clipping (norm from ``mh_sumsq``), moment updates, parameter update and
    gradient zeroing.  Numerically the same update rule as ``torch.optim.Adam``."""

    def __init__(self, flat: FlatBuffers, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.flat = flat
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), tuple(betas), float(eps), float(weight_decay)
        self.exp_avg = torch.zeros_like(flat.flat_param)
        self.exp_avg_sq = torch.zeros_like(flat.flat_param)
        self.shadow = flat.enable_shadow() if flat.flat_param.is_cuda else None
        self.step_count = torch.zeros(1, device=flat.flat_param.device, dtype=torch.int64)
        self.sumsq = torch.zeros(1, device=flat.flat_param.device, dtype=torch.float32)
        self.param_order = None  # all parameters the reference would hand to Adam, in its order (set by TrainStep)

    def step(self, grad_scale=1.0, max_norm=0.0):
        K.counter_add(self.step_count, 1)
        self.sumsq.zero_()
        mask = self.flat.flat_mask  # weight-pruning mode: pruned elements get no gradient, operands are param * mask
        K.sumsq_add(self.flat.flat_grad, self.sumsq, mask=mask)
        K.adam_step(self.flat.flat_param, self.flat.flat_grad, self.exp_avg, self.exp_avg_sq, self.step_count,
                    lr=self.lr, beta1=self.betas[0], beta2=self.betas[1], eps=self.eps, weight_decay=self.weight_decay,
                    grad_scale=grad_scale, max_norm=max_norm, sumsq=self.sumsq, zero_grad=True, shadow=self.shadow,
                    mask=mask, effective=self.flat.flat_eff if mask is not None else None)
        ops.bump_weight_epoch()  # operands that are NOT views of the shadow (pruned / masked weights) are rebuilt
        self.flat._shadow_epoch = ops._EPOCH[0]  # ... the shadow itself is current

    def grad_norm(self, grad_scale=1.0):
        return float(self.sumsq.sqrt().item()) * grad_scale

    def zero_grad(self):
        self.flat.flat_grad.zero_()

    def state_dict(self):
        """``torch.optim.Adam.state_dict()`` layout (what the reference stores under ``"Optimizer"``, runner.py:154-172,
        mh_utils.py:17): ``state[i] = {step, exp_avg, exp_avg_sq}`` with i the parameter's position in
        ``param_order`` (the expert's ``parameters()`` order -- frozen parameters hold an index but no state, exactly
        like gradient-less parameters in torch), one param group."""
        order = self.param_order if self.param_order is not None else self.flat.params
        slot = {id(p): (o, p) for p, o in zip(self.flat.params, self.flat.offsets)}
        state = {}
        step = self.step_count.to(torch.float32).reshape(()).cpu()
        if float(step) > 0:
            for i, p in enumerate(order):
                if id(p) in slot:
                    o = slot[id(p)][0]
                    state[i] = {"step": step.clone(), "exp_avg": self.exp_avg[o:o + p.numel()].view(p.shape).clone(),
                                "exp_avg_sq": self.exp_avg_sq[o:o + p.numel()].view(p.shape).clone()}
        group = {"lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay, "amsgrad": False,
                 "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "decoupled_weight_decay": False, "params": list(range(len(order)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        """Accepts the torch layout above (also a checkpoint written by the reference's ``torch.optim.Adam``) and the
        flat layout of this class's first version.  Shapes must match the current parameters."""
        if "param_groups" not in sd:
            self.exp_avg.copy_(sd["exp_avg"])
            self.exp_avg_sq.copy_(sd["exp_avg_sq"])
            self.step_count.copy_(sd["step"])
            return
        order = self.param_order if self.param_order is not None else self.flat.params
        if len(sd["param_groups"][0]["params"]) != len(order):
            raise ValueError(f"optimizer state holds {len(sd['param_groups'][0]['params'])} parameters, the model has {len(order)}")
        slot = {id(p): o for p, o in zip(self.flat.params, self.flat.offsets)}
        step = 0
        for i, st in sd["state"].items():
            p = order[int(i)]
            if id(p) not in slot:
                continue
            if tuple(st["exp_avg"].shape) != tuple(p.shape):
                raise ValueError(f"optimizer state {i}: shape {tuple(st['exp_avg'].shape)} != parameter {tuple(p.shape)}")
            o = slot[id(p)]
            self.exp_avg[o:o + p.numel()].view(p.shape).copy_(st["exp_avg"])
            self.exp_avg_sq[o:o + p.numel()].view(p.shape).copy_(st["exp_avg_sq"])
            step = max(step, int(float(st["step"])))
        self.step_count.fill_(step)
        g = sd["param_groups"][0]
        self.lr, self.betas, self.eps = float(g["lr"]), tuple(g["betas"]), float(g["eps"])
        self.weight_decay = float(g["weight_decay"])


def trainable_params(expert):
    """Layer-contiguous parameter order (so each encoder layer is one all-reduce bucket).  Inside an
    attention module the q, k, v weights (then their biases) are placed back to back in the order of the
    fused QKV GEMM's columns, so their three weight-gradient GEMMs / bias column sums collapse into one
    call writing a [3E, C] block of the flat gradient buffer (ops._wgrad_qkv)."""
    from .fairseq_code import MultiheadAttention

    order = {}
    for m in expert.modules():
        if isinstance(m, MultiheadAttention):
            group = []
            for name in ("weight", "bias"):
                for proj in (m.q_proj, m.k_proj, m.v_proj):
                    t = proj._parameters.get(name + "_orig", proj._parameters.get(name))
                    if t is not None:
                        group.append(t)
            first = min((id(t) for t in group), default=None)
            for t in group:
                order[id(t)] = group
    seen, out = set(), []
    for p in expert.parameters():
        if not p.requires_grad or id(p) in seen:
            continue
        for t in order.get(id(p), [p]):  # the first q/k/v tensor met pulls its whole group in, in GEMM order
            if t.requires_grad and id(t) not in seen:
                seen.add(id(t))
                out.append(t)
    return out


class TrainStep:
    """One optimizer step of an expert (pre-training / pruning fine-tune / distillation) on a
    statically shaped batch, eager or CUDA-graph captured.

    Static device buffers: feat (B,T,D) f32, label (B,T) i64, pad_mask (B,T) f32, span mask
    (B,T) bool.  ``load_batch`` does the host->device copies (from pinned memory) and draws the
    span mask on the host exactly like the reference (NumPy global stream, including the
    per-layer layer-drop draws that shift the stream)."""

    def __init__(self, expert, B, T, D, *, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_norm=10.0,
                 use_graph=True, device="cuda", accum=1):
        self.expert, self.B, self.T, self.D = expert, B, T, D
        # gradient accumulation (runner.py:370-371, :396-399): ``accum`` micro-batches per optimizer step.  The wgrad
        # kernels add into the flat gradient buffer, so a micro-batch is one more forward + backward; the 1/accum of
        # ``loss / gradient_accumulate_steps`` is applied once, as the fused Adam step's grad_scale.
        self.accum = max(int(accum), 1)
        self._micro = 0
        self.device = torch.device(device)
        self.max_norm = max_norm
        self.model = expert.model
        self.model.static_rows = True
        self.teacher = getattr(expert, "teacher_model", None)
        if self.teacher is not None:
            self.teacher.static_rows = True
        self.flat = FlatBuffers(trainable_params(expert))
        self.flat.bind_modules(expert)
        self.opt = FusedAdam(self.flat, lr, betas, eps, weight_decay)
        self.opt.param_order = list(expert.parameters())  # runner.py:314 Adam(expert.parameters())
        self.dp = getattr(expert, "dp", None)
        if self.dp is not None:
            self.dp.attach(self.flat)
            self.flat.sync_shadow()  # attach() broadcast rank 0's parameters
            if self.dp.enabled and use_graph and os.environ.get("MH_DP_GRAPH", "1") == "0":
                # The data-parallel step (NCCL bucket all-reduces on the side stream included) is captured into one
                # CUDA graph like the single-GPU one (measured on 2 and 8 B200: 22.7 / 22.8 ms against 23.1 / 23.3 ms
                # eager); MH_DP_GRAPH=0 runs it eagerly.
                use_graph = False
        dev = self.device
        self.feat = torch.zeros(B, T, D, device=dev)
        self.label = torch.zeros(B, T, device=dev, dtype=torch.int64)
        self.pad = torch.ones(B, T, device=dev)
        self.mask = torch.zeros(B, T, device=dev, dtype=torch.bool)
        self.loss = torch.zeros(1, device=dev)
        self.loss_acc = torch.zeros(1, device=dev)  # sum of the micro-batch losses of the step in flight
        self.h_loss = torch.zeros(1).pin_memory()
        self.rng_counter = torch.zeros(1, device=dev, dtype=torch.int64)
        K.set_dropout_offset(self.rng_counter)
        self.use_graph = use_graph
        self.graphs = None  # {False: forward + backward of a non-final micro-batch, True: final micro-batch + optimizer}
        self.n_layerdrop_draws = self.model.model_config.encoder_layers + (
            self.teacher.model_config.encoder_layers if self.teacher is not None else 0)
        self.h2d_bytes = self.d2h_bytes = 0

    # ------------------------------------------------------------------------------------------
    def draw_mask(self, lens):
        cfg = (self.teacher or self.model).model_config
        from .fairseq_code import compute_mask_indices

        m = compute_mask_indices((self.B, self.T), None, cfg.mask_prob, cfg.mask_length, cfg.mask_selection,
                                 cfg.mask_other, min_masks=2, no_overlap=cfg.no_mask_overlap,
                                 min_space=cfg.mask_min_space, require_same_masks=False, valid_lens=lens)
        return m

    def load_batch(self, feat, label, pad, lens, masking=True):
        """feat / label / pad: pinned host tensors.  ``stage_batch`` + ``commit_staged`` in one call: the copies are
        asynchronous, and the pinned span-mask buffers are only rewritten once their previous copy has run -- a
        training loop that never synchronises may be several steps ahead of the GPU."""
        self.stage_batch(feat, label, pad, lens, masking)
        self.commit_staged()

    # ---- input pipelining (no host work or sync between steps): the next batch is drawn / copied
    #      while the current step runs, the loss is read back one step late
    def stage_batch(self, feat, label, pad, lens, masking=True):
        """Asynchronous ``load_batch``: the host->device copies (pinned memory) go to STAGING buffers on a copy
        stream, so they and the host-side span-mask draw overlap the step in flight; ``commit_staged`` moves
        them into the step's static input buffers.  Call it AFTER ``run`` of the previous step (the NumPy
        stream must see mask draw, layer-drop draws, mask draw, ... in the reference's order)."""
        if not hasattr(self, "_cs"):
            dev = self.device
            self._cs = torch.cuda.Stream(device=dev)
            self._s_feat, self._s_label = torch.empty_like(self.feat), torch.empty_like(self.label)
            self._s_pad, self._s_mask = torch.empty_like(self.pad), torch.empty_like(self.mask)
            self._h_masks = [torch.zeros(self.B, self.T, dtype=torch.bool).pin_memory() for _ in range(2)]
            self._h_mask_ev = [torch.cuda.Event(), torch.cuda.Event()]
            self._staged_ev, self._commit_ev = torch.cuda.Event(), torch.cuda.Event()
            self._commit_ev.record()
            self._stage_i = 0
        k = self._stage_i & 1
        hm = self._h_masks[k]
        self._stage_i += 1
        if masking:
            m = self.draw_mask(lens)
            self._h_mask_ev[k].synchronize()  # the copy that last read this pinned buffer (two stages ago) has run
            hm.copy_(torch.from_numpy(m))
        self._cs.wait_event(self._commit_ev)  # the previous commit has read the staging buffers
        with torch.cuda.stream(self._cs):
            self._s_feat.copy_(feat, non_blocking=True)
            self._s_label.copy_(label, non_blocking=True)
            self._s_pad.copy_(pad, non_blocking=True)
            if masking:
                self._s_mask.copy_(hm, non_blocking=True)
                self._h_mask_ev[k].record()
            self._staged_ev.record()
        self._staged = (feat.numel() * 4 + label.numel() * 8 + pad.numel() * 4 + (hm.numel() if masking else 0),
                        list(lens), masking)

    def commit_staged(self):
        """Device-to-device copy of the staged batch into the step's input buffers (current stream)."""
        nbytes, lens, masking = self._staged
        cur = torch.cuda.current_stream()
        cur.wait_event(self._staged_ev)
        self.feat.copy_(self._s_feat, non_blocking=True)
        self.label.copy_(self._s_label, non_blocking=True)
        self.pad.copy_(self._s_pad, non_blocking=True)
        if masking:
            self.mask.copy_(self._s_mask, non_blocking=True)
        self._commit_ev.record(cur)
        self.h2d_bytes, self._lens = nbytes, lens

    def read_loss_async(self):
        """Queue the device->host copy of this step's loss; returns a handle for ``collect_loss``."""
        if not hasattr(self, "_h_losses"):
            self._h_losses = [(torch.zeros(1).pin_memory(), torch.cuda.Event()) for _ in range(4)]
            self._loss_i = 0
        buf, ev = self._h_losses[self._loss_i & 3]
        self._loss_i += 1
        buf.copy_(self.loss, non_blocking=True)
        ev.record()
        self.d2h_bytes = 4
        return buf, ev

    @staticmethod
    def collect_loss(handle):
        buf, ev = handle
        ev.synchronize()
        return float(buf[0])

    def _body(self, last=True):
        """One micro-batch: forward + backward into the flat gradient buffer.  ``last``: also the gradient
        all-reduce, the loss mean over the step's micro-batches and clip + Adam (+ zero_grad)."""
        K.counter_add(self.rng_counter, 1)
        data = (self.feat, self.label, self.pad, None)
        if self.dp is not None:
            self.dp.sync = last  # bucket all-reduces only behind the final micro-batch's backward
        self._patch_mask(True)
        try:
            loss, _ = self.expert(data)
        finally:
            self._patch_mask(False)
        loss.backward()
        loss = loss.detach().reshape(1)
        if not last:
            self.loss_acc.add_(loss)
            return
        if self.dp is not None:
            self.dp.finish()
        if self.accum > 1:
            self.loss.copy_((self.loss_acc + loss) / self.accum)
            self.loss_acc.zero_()
        else:
            self.loss.copy_(loss)
        self.opt.step(grad_scale=1.0 / self.accum, max_norm=self.max_norm)

    def _patch_mask(self, on):
        """Feed the pre-drawn device mask through the model's ``teacher_mask_indices`` door so no
        host RNG / H2D copy happens inside the (capturable) step body."""
        first = self.teacher or self.model
        if on:
            mask = self.mask
            self._orig_draw = first._draw_mask
            first._draw_mask = lambda B, T, lens, pm, tmi, dev: (tmi if tmi is not None else mask)
        else:
            first._draw_mask = self._orig_draw

    def _snapshot(self):
        o, f = self.opt, self.flat
        t = [f.flat_param, f.flat_grad, o.exp_avg, o.exp_avg_sq, o.step_count, o.sumsq, self.rng_counter, self.loss,
             self.loss_acc]
        if f.flat_bf16 is not None:
            t.append(f.flat_bf16)
        if f.flat_eff is not None:
            t.append(f.flat_eff)
        return [(x, x.clone()) for x in t], np.random.get_state()

    @staticmethod
    def _restore(snap):
        for x, c in snap[0]:
            x.copy_(c)
        np.random.set_state(snap[1])

    def capture(self, warmup=2):
        """Warm up eagerly on a side stream, then capture the step into CUDA graphs (one for a non-final micro-batch when
        accumulating, one for the final micro-batch + optimizer).  The warm-up runs real optimizer steps on whatever
        batch is loaded, so parameters, Adam state, step / dropout counters and the NumPy stream are put back afterwards:
        capturing changes nothing the training run can see."""
        if getattr(self.flat, "_shadow_epoch", None) != ops._EPOCH[0]:
            # parameters / prune masks changed outside the optimizer since the operand shadow was written: rebuild it
            # NOW, so that the snapshot below holds (and later restores) a current shadow
            self.flat.sync_shadow()
            self.flat._shadow_epoch = ops._EPOCH[0]
        snap = self._snapshot()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                if self.accum > 1:
                    self._body(False)
                self._body(True)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self._restore(snap)
        ops.bump_weight_epoch()  # operand prep of masked / non-flat weights must be part of the captured step
        self.flat._shadow_epoch = ops._EPOCH[0]  # (the bf16 shadow is current: restored with the parameters)
        self.graphs, self.launches = {}, {}
        pool = torch.cuda.graph_pool_handle()
        for last in ((False, True) if self.accum > 1 else (True,)):
            g = torch.cuda.CUDAGraph()
            n0 = K.L.launch_count()
            with torch.cuda.graph(g, pool=pool):  # capturing executes nothing on the device ...
                self._body(last)
            self.graphs[last], self.launches[last] = g, K.L.launch_count() - n0
            ops.bump_weight_epoch()  # each graph carries its own operand prep of non-shadow weights
            self.flat._shadow_epoch = ops._EPOCH[0]
        np.random.set_state(snap[1])  # ... but the per-layer layer-drop draws of the captured forwards advanced the host stream
        self.launches_per_step = self.launches[True] + (self.accum - 1) * self.launches.get(False, 0)
        return self

    def run(self):
        """Execute one micro-batch on the currently loaded batch; every ``accum``-th call is the final one of an
        optimizer step (returns True then).  The loss of a finished step (mean over its micro-batches) is in
        ``self.loss`` on the device; ``read_loss`` / ``read_loss_async`` fetch it."""
        last = (self._micro + 1) % self.accum == 0
        self._micro += 1
        if self.use_graph:
            if self.graphs is None:
                self.capture()
            for _ in range(self.n_layerdrop_draws):
                np.random.random()  # a replay skips the per-layer host draws of the eager forward (module.py:243)
            self.graphs[last].replay()
        else:
            n0 = K.L.launch_count()
            self._body(last)
            n = K.L.launch_count() - n0
            self.launches_per_step = n if self.accum == 1 else (n + getattr(self, "_micro_launches", 0) if last else 0)
            self._micro_launches = 0 if last else getattr(self, "_micro_launches", 0) + n
        return last

    def read_loss(self):
        self.h_loss.copy_(self.loss, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        self.d2h_bytes = 4
        return float(self.h_loss[0])
