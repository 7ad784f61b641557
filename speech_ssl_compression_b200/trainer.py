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

    def step(self, grad_scale=1.0, max_norm=0.0):
        K.counter_add(self.step_count, 1)
        self.sumsq.zero_()
        K.sumsq_add(self.flat.flat_grad, self.sumsq)
        K.adam_step(self.flat.flat_param, self.flat.flat_grad, self.exp_avg, self.exp_avg_sq, self.step_count,
                    lr=self.lr, beta1=self.betas[0], beta2=self.betas[1], eps=self.eps, weight_decay=self.weight_decay,
                    grad_scale=grad_scale, max_norm=max_norm, sumsq=self.sumsq, zero_grad=True, shadow=self.shadow)
        ops.bump_weight_epoch()  # operands that are NOT views of the shadow (pruned / masked weights) are rebuilt
        self.flat._shadow_epoch = ops._EPOCH[0]  # ... the shadow itself is current

    def grad_norm(self, grad_scale=1.0):
        return float(self.sumsq.sqrt().item()) * grad_scale

    def zero_grad(self):
        self.flat.flat_grad.zero_()

    def state_dict(self):
        return {"exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq, "step": self.step_count,
                "lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay}

    def load_state_dict(self, sd):
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.step_count.copy_(sd["step"])


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
                 use_graph=True, device="cuda"):
        self.expert, self.B, self.T, self.D = expert, B, T, D
        self.device = torch.device(device)
        self.max_norm = max_norm
        self.model = expert.model
        self.model.static_rows = True
        self.teacher = getattr(expert, "teacher_model", None)
        if self.teacher is not None:
            self.teacher.static_rows = True
        self.flat = FlatBuffers(trainable_params(expert))
        self.opt = FusedAdam(self.flat, lr, betas, eps, weight_decay)
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
        self.h_loss = torch.zeros(1).pin_memory()
        self.rng_counter = torch.zeros(1, device=dev, dtype=torch.int64)
        K.set_dropout_offset(self.rng_counter)
        self.use_graph = use_graph
        self.graph = None
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

    def _body(self):
        K.counter_add(self.rng_counter, 1)
        data = (self.feat, self.label, self.pad, None)
        self._patch_mask(True)
        try:
            loss, _ = self.expert(data)
        finally:
            self._patch_mask(False)
        loss.backward()
        if self.dp is not None:
            self.dp.finish()
        self.loss.copy_(loss.detach().reshape(1))
        self.opt.step(grad_scale=1.0, max_norm=self.max_norm)

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

    def capture(self, warmup=2):
        """Warm up eagerly on a side stream, then capture one step into a CUDA graph."""
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._body()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        ops.bump_weight_epoch()  # operand prep of masked / non-flat weights must be part of the captured step
        self.flat._shadow_epoch = ops._EPOCH[0]  # (the bf16 shadow is current: written by the warm-up Adam steps)
        self.graph = torch.cuda.CUDAGraph()
        n0 = K.L.launch_count()
        rng_state = np.random.get_state()  # the per-layer layer-drop draws of the captured forward are replayed
        with torch.cuda.graph(self.graph):  # by run(); capturing must not advance the host stream
            self._body()
        np.random.set_state(rng_state)
        self.launches_per_step = K.L.launch_count() - n0
        return self

    def run(self):
        """Execute one optimizer step on the currently loaded batch; returns nothing (loss is in
        ``self.loss`` on the device; ``read_loss`` fetches it)."""
        if self.use_graph:
            if self.graph is None:
                self.capture()
            for _ in range(self.n_layerdrop_draws):
                np.random.random()  # a replay skips the per-layer host draws of the eager forward (module.py:243)
            self.graph.replay()
        else:
            n0 = K.L.launch_count()
            self._body()
            self.launches_per_step = K.L.launch_count() - n0

    def read_loss(self):
        self.h_loss.copy_(self.loss, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        self.d2h_bytes = 4
        return float(self.h_loss[0])
