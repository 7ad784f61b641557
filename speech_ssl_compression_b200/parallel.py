"""Data parallelism: one process per GPU over NVLink / NVSwitch (replaces the reference's
single-process ``nn.DataParallel``, ``upstream/melhubert/pretrain_expert.py:28-30``).

Gradients live in ONE flat fp32 buffer (``param.grad`` are views into it, so the wgrad GEMMs
accumulate straight into it).  Each encoder layer owns a contiguous slice; when a layer's
backward returns, its slice is summed over the ranks on a side stream while the remaining layers keep
computing -- the last bucket (everything outside the layers) is reduced in ``finish()``.
The cross-entropy normaliser is made global with a 2-float all-reduce (``all_reduce_sum``),
which reproduces DataParallel's gather-then-mean loss semantics.

Two transports for the buckets (``MH_DP_TRANSPORT``):
  * ``peer`` (default on CUDA): own reduce-scatter / all-gather kernels that pull over peer-mapped gradient buffers
    (``csrc/peer.cu``; CUDA IPC handles exchanged once through the process group).  Their CTAs are small enough to
    share SMs with the persistent GEMM / attention kernels, which ``ncclAllReduce`` CTAs cannot -- NCCL on the side
    stream cost 1.4-1.7 ms per step (SCALE_r01) by knocking persistent GEMM clusters into a second wave.
  * ``nccl``: ``torch.distributed.all_reduce`` per bucket (also what the CPU / gloo tests exercise).
torch.distributed (NCCL / gloo) stays the control plane: rendezvous, the initial parameter broadcast, handle exchange.
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
        # weight-pruning mode: one mask byte per element (0 = pruned) and the fp32 effective parameters param * mask
        # (bias operands); built by refresh_masks() from the modules' <name>_mask buffers, consumed by the masked
        # optimizer kernels -- replaces the reference's 144 per-forward masked_fill launches (prune.py:24-38)
        self.flat_mask = self.flat_eff = None
        self.owners = {}       # id(param) -> (module, parameter name), filled by bind_modules()
        self._mask_sig = {}    # id(param) -> (mask data_ptr, mask version) of what flat_mask currently holds
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

    def bind_modules(self, root):
        """Remember which module owns each flat parameter (needed to find its prune mask)."""
        mine = {id(p) for p in self.params}
        for mod in root.modules():
            for name, p in mod._parameters.items():
                if p is not None and id(p) in mine:
                    self.owners[id(p)] = (mod, name)

    def mask_of(self, p):
        own = self.owners.get(id(p))
        if own is None or not own[1].endswith("_orig"):
            return None
        return own[0]._buffers.get(own[1][:-5] + "_mask")

    def refresh_masks(self):
        """flat_mask <- the modules' current prune masks (all ones where a parameter has none)."""
        masks = [(p, o, self.mask_of(p)) for p, o in zip(self.params, self.offsets)]
        if not any(m is not None for _, _, m in masks):
            self.flat_mask = self.flat_eff = None
            self._mask_sig = {}
            return
        if self.flat_mask is None:
            self.flat_mask = torch.ones(self.total, device=self.flat_param.device, dtype=torch.uint8)
            self.flat_eff = torch.empty_like(self.flat_param)
        self._mask_sig = {}
        for p, o, m in masks:
            if m is None:
                self.flat_mask[o:o + p.numel()].fill_(1)
            else:
                self.flat_mask[o:o + p.numel()].copy_(m.reshape(-1))
                self._mask_sig[id(p)] = (m.data_ptr(), m._version)

    def masks_current(self, p, m):
        return self._mask_sig.get(id(p)) == (m.data_ptr(), m._version)

    def sync_shadow(self):
        from . import kernels as K

        if self.flat_bf16 is not None:
            if self.owners:
                self.refresh_masks()
            K.flat_effective(self.flat_param, self.flat_mask, self.flat_bf16, self.flat_eff)

    def span(self, params):
        """(start, end) of the contiguous slice covering ``params`` (must be adjacent)."""
        idx = [i for i, p in enumerate(self.params) if any(p is q for q in params)]
        if not idx:
            return None
        lo, hi = min(idx), max(idx)
        end = self.offsets[hi + 1] if hi + 1 < len(self.offsets) else self.total
        return self.offsets[lo], end


class PeerGradExchange:
    """Peer-memory transport: every rank maps every other rank's flat gradient buffer, flag array and mailbox
    (CUDA IPC), then sums buckets with ``mh_peer_reduce_scatter`` / ``mh_peer_all_gather`` (csrc/peer.cu)."""

    def __init__(self, flat_grad, rank, world):
        from torch.multiprocessing.reductions import reduce_tensor

        self.rank, self.world = rank, world
        self.use_sm = os.environ.get("MH_DP_TRANSPORT") == "peer-sm"
        self.staging = None  # copy-engine transport: (world - 1) staged peer shards of the largest bucket
        dev = flat_grad.device
        # own allocations (cudaMalloc blocks of their own: an IPC handle exports the whole block)
        self.flags = torch.zeros(8, device=dev, dtype=torch.int64)
        self.mail = torch.zeros(2 * 8 * 16, device=dev, dtype=torch.float32)
        self.state = torch.zeros(2, device=dev, dtype=torch.int64)
        torch.cuda.synchronize()
        mine = []
        for t in (flat_grad, self.flags, self.mail):
            fn, args = reduce_tensor(t)
            mine.append((fn, list(args)))
        everyone = [None] * world
        dist.all_gather_object(everyone, mine)
        self._keep = []  # peer-mapped tensors must stay alive as long as their pointers are used
        self.grad_ptrs, self.flag_ptrs, self.mail_ptrs = [], [], []
        for p in range(world):
            if p == rank:
                ts = (flat_grad, self.flags, self.mail)
            else:
                ts = []
                for fn, args in everyone[p]:
                    # rebuild_cuda_tensor(cls, size, stride, offset, storage_cls, dtype, storage_device, handle, ...):
                    # open the handle on THIS device -- the mapping is then reachable from our kernels through peer
                    # access (cudaIpcMemLazyEnablePeerAccess) without creating a context on the peer device
                    args = list(args)
                    args[6] = dev.index
                    ts.append(fn(*args))
                self._keep.append(ts)
            self.grad_ptrs.append(ts[0].data_ptr())
            self.flag_ptrs.append(ts[1].data_ptr())
            self.mail_ptrs.append(ts[2].data_ptr())
        dist.barrier()

    def reduce_span(self, start, end):
        from . import kernels as K

        n = end - start
        if self.use_sm:  # SM-driven pulls (A/B only)
            K.peer_reduce_scatter(self.grad_ptrs, self.flag_ptrs, self.state, start, n, self.rank, self.world)
            K.peer_all_gather(self.grad_ptrs, self.flag_ptrs, self.state, start, n, self.rank, self.world)
            return
        shard = ((n + self.world - 1) // self.world + 3) // 4 * 4
        need = (self.world - 1) * ((shard + 31) // 32 * 32)
        if self.staging is None or self.staging.numel() < need:
            self.staging = torch.empty(need, device=self.flags.device, dtype=torch.float32)
        K.peer_reduce_scatter_ce(self.grad_ptrs, self.flag_ptrs, self.state, self.staging, start, n, self.rank, self.world)
        K.peer_all_gather_ce(self.grad_ptrs, self.flag_ptrs, self.state, start, n, self.rank, self.world)

    def barrier(self, vals=None):
        """Cross-rank barrier on the current stream; ``vals`` (<= 16 fp32, optional) is summed over the ranks in place."""
        from . import kernels as K

        K.peer_barrier_sum(self.grad_ptrs, self.flag_ptrs, self.mail_ptrs, self.state, vals, self.rank, self.world)


class DataParallelB200:
    def __init__(self, model, overlap=True):
        self.model = model
        self.rank, self.world_size = init_distributed()
        self.overlap = overlap and os.environ.get("MH_DP_OVERLAP", "1") != "0"
        self.enabled = self.world_size > 1
        self.flat = None
        self._pending = []
        self._comm_stream = None
        self.sync = True  # False while a non-final micro-batch of a gradient-accumulation step runs (no all-reduce)
        self.peer = None  # PeerGradExchange once attached (CUDA, MH_DP_TRANSPORT != nccl)

    # -- loss normaliser ---------------------------------------------------------------------
    def all_reduce_sum(self, t):
        if self.enabled:
            if self.peer is not None and t.dtype == torch.float32 and t.numel() <= 16 and t.is_contiguous():
                self._peer_on_main(lambda: self.peer.barrier(t))  # (no NCCL kernel in the step at all)
            else:
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t

    def _peer_on_main(self, fn):
        """Peer kernels share ONE epoch counter per rank, so they must execute in issue order.  Calls from the main
        stream need no explicit edge: ``finish()`` joined the communication stream at the end of the previous step
        (and a bucket reduction makes the communication stream wait for the main stream first, ``_reduce_span``)."""
        fn()

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
            self._comm_stream = torch.cuda.Stream(priority=-1)
            if os.environ.get("MH_DP_TRANSPORT", "peer") not in ("nccl", "none") and flat.flat_grad.is_cuda:
                # the mapping can fail for reasons outside this code (no peer access between the devices, IPC disabled
                # in the container): all ranks then agree to fall back to NCCL for the buckets -- loudly
                try:
                    peer, err = PeerGradExchange(flat.flat_grad, self.rank, self.world_size), None
                except Exception as e:  # noqa: BLE001
                    peer, err = None, e
                ok = torch.tensor([1.0 if peer is not None else 0.0], device=flat.flat_grad.device)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)
                if float(ok) == 1.0:
                    self.peer = peer
                else:
                    print(f"[DataParallelB200] rank {self.rank}: peer-memory gradient exchange unavailable "
                          f"({err if err is not None else 'another rank failed'}); using NCCL all-reduce for the buckets", flush=True)

    def _reduce_span(self, span):
        if not self.enabled or span is None or not self.sync:
            return
        if os.environ.get("MH_DP_TRANSPORT") == "none":  # timing experiments only: no gradient exchange at all
            return
        buf = self.flat.flat_grad[span[0]:span[1]]
        reduce = (lambda: self.peer.reduce_span(span[0], span[1])) if self.peer is not None else \
            (lambda: dist.all_reduce(buf, op=dist.ReduceOp.SUM))
        if self._comm_stream is not None and self.overlap:
            self._comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._comm_stream):
                reduce()
            self._comm_used = True
        else:
            reduce()

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
        if self._comm_stream is not None and getattr(self, "_comm_used", False):
            torch.cuda.current_stream().wait_stream(self._comm_stream)
            self._comm_used = False
        if self.peer is not None:
            self.peer.barrier()  # every peer has pulled what it needs: the optimizer may zero the buffer
