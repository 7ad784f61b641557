"""Data parallelism: one process per GPU, NCCL over NVLink (replaces the reference's
single-process ``nn.DataParallel``, ``upstream/melhubert/pretrain_expert.py:28-30``).

Gradients live in ONE flat fp32 buffer (``param.grad`` are views into it, so the wgrad GEMMs
accumulate straight into it).  Each encoder layer owns a contiguous slice; when a layer's
backward returns, its slice is all-reduced on a side stream while the remaining layers keep
computing -- the last bucket (everything outside the layers) is reduced in ``finish()``.
The cross-entropy normaliser is made global with a 2-float all-reduce (``all_reduce_sum``),
which reproduces DataParallel's gather-then-mean loss semantics.
"""
import os

import torch
import torch.distributed as dist


def init_distributed(backend=None):
    """Initialise torch.distributed from the torchrun environment (idempotent)."""
    if dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1:
        return 0, 1
    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world


class FlatBuffers:
    """Flat fp32 parameter / gradient storage: every ``Parameter.data`` and ``.grad`` becomes a
    view into one buffer (128-byte aligned slots), layer-contiguous."""

    def __init__(self, params, align=32):
        self.params = [p for p in params]
        dev = self.params[0].device
        offs, total = [], 0
        for p in self.params:
            offs.append(total)
            total += (p.numel() + align - 1) // align * align
        self.offsets, self.total = offs, total
        self.flat_param = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_grad = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_bf16 = None  # bf16 shadow of flat_param (GEMM operands), kept current by the fused optimizer
        for p, o in zip(self.params, offs):
            view = self.flat_param[o:o + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
            p.grad = self.flat_grad[o:o + p.numel()].view(p.shape)
            p._mh_flat = (self, o)  # lets ops.packed_operands find the parameter's slot in the shadow

    def enable_shadow(self):
        """Allocate and fill the bf16 shadow (CUDA only)."""
        from . import kernels as K

        from . import ops

        self.flat_bf16 = torch.empty(self.total, device=self.flat_param.device, dtype=torch.bfloat16)
        self.sync_shadow()
        self._shadow_epoch = ops._EPOCH[0]
        return self.flat_bf16

    def sync_shadow(self):
        from . import kernels as K

        if self.flat_bf16 is not None:
            K.to_bf16(self.flat_param, self.flat_bf16)

    def span(self, params):
        """(start, end) of the contiguous slice covering ``params`` (must be adjacent)."""
        idx = [i for i, p in enumerate(self.params) if any(p is q for q in params)]
        if not idx:
            return None
        lo, hi = min(idx), max(idx)
        end = self.offsets[hi + 1] if hi + 1 < len(self.offsets) else self.total
        return self.offsets[lo], end


class DataParallelB200:
    def __init__(self, model, overlap=True):
        self.model = model
        self.rank, self.world_size = init_distributed()
        self.overlap = overlap
        self.enabled = self.world_size > 1
        self.flat = None
        self._pending = []
        self._comm_stream = None
        self.sync = True  # False while a non-final micro-batch of a gradient-accumulation step runs (no all-reduce)

    # -- loss normaliser ---------------------------------------------------------------------
    def all_reduce_sum(self, t):
        if self.enabled:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t

    # -- gradient buckets --------------------------------------------------------------------
    def attach(self, flat: FlatBuffers):
        """Register the flat gradient buffer and hook every encoder layer's backward."""
        self.flat = flat
        if self.enabled and float(getattr(self.model.encoder, "layerdrop", 0.0) or 0.0) > 0.0:
            # a layer skipped by its host-side draw (module.py:243) issues no bucket all-reduce: ranks whose draws
            # differ would post mismatched collectives and hang
            raise RuntimeError("encoder_layerdrop > 0 is not supported with data parallelism")
        if self.enabled:  # identical start on every rank: one broadcast of the whole flat parameter buffer
            dist.broadcast(flat.flat_param, src=0)
        self._layer_spans = []
        covered = []
        for layer in self.model.encoder.layers:
            span = flat.span(list(layer.parameters()))
            self._layer_spans.append(span)
            covered.append(span)
            layer._mh_grad_ready_hook = self._make_hook(span)
        # what is left (pre_extract_proj, pos_conv, encoder LN, final_proj): contiguous head / tail pieces
        rest, cur = [], 0
        for s, e in sorted(covered):
            if s > cur:
                rest.append((cur, s))
            cur = max(cur, e)
        if cur < flat.total:
            rest.append((cur, flat.total))
        self._rest_spans = rest
        if self.enabled and torch.cuda.is_available():
            self._comm_stream = torch.cuda.Stream()

    def _reduce_span(self, span):
        if not self.enabled or span is None or not self.sync:
            return
        buf = self.flat.flat_grad[span[0]:span[1]]
        if self._comm_stream is not None and self.overlap:
            self._comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._comm_stream):
                dist.all_reduce(buf, op=dist.ReduceOp.SUM)
        else:
            dist.all_reduce(buf, op=dist.ReduceOp.SUM)

    def _make_hook(self, span):
        def hook(layer):
            self._reduce_span(span)
        return hook

    def finish(self):
        """All-reduce the non-layer remainder and join the communication stream."""
        if not self.sync:
            return
        for span in self._rest_spans:
            self._reduce_span(span)
        if self._comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self._comm_stream)
