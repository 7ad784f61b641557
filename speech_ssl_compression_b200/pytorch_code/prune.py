"""Mask-based pruning re-parametrisation, B200 edition.

Same observable behaviour as the reference's ``pytorch_code/prune.py`` (itself a fork of
``torch.nn.utils.prune`` with bool masks) for the calls the MelHuBERT path makes
(reference prune.py:24-38, 64-85, 190-215, 266-296, 553-573, 834-846, 1049-1171):

* ``module.<name>_orig``  trainable Parameter (the SAME Parameter object as before pruning,
  so optimizer state survives),
* ``module.<name>_mask``  persistent bool buffer (state_dict key ``..._mask``),
* ``module.<name>``       non-persistent buffer = ``orig.masked_fill(~mask, 0)``, refreshed by a
  forward pre-hook whose object carries ``_tensor_name``.

What is different: the *hot path never runs these hooks*.  The fused encoder reads
``<name>_orig`` + ``<name>_mask`` directly and folds the mask into the bf16 operand
preparation kernel (``mh_weight_prep``), and masks the weight gradient inside the wgrad GEMM
epilogue.  The global magnitude selection runs on the GPU as an exact radix select
(``mh_abs_kth_smallest``) with the tie rule "lowest flat index wins" (DESIGN.md, H3); on CPU
tensors (tests / tools) it falls back to ``torch.topk`` exactly as the reference does.
"""
import numbers

import torch


class BasePruningMethod:
    PRUNING_TYPE = "unstructured"
    _tensor_name = None

    def __call__(self, module, inputs):
        module._buffers[self._tensor_name] = self.apply_mask(module)

    def compute_mask(self, t, default_mask):
        raise NotImplementedError

    def apply_mask(self, module):
        if self._tensor_name is None:
            raise AssertionError(f"Module {module} has to be pruned")
        mask = getattr(module, self._tensor_name + "_mask")
        orig = getattr(module, self._tensor_name + "_orig")
        return orig.masked_fill(~mask.bool(), 0)

    @classmethod
    def apply(cls, module, name, *args, importance_scores=None, **kwargs):
        method = cls(*args, **kwargs)
        method._tensor_name = name
        previous = [k for k, h in module._forward_pre_hooks.items()
                    if isinstance(h, BasePruningMethod) and h._tensor_name == name]
        if len(previous) > 1:
            raise AssertionError(f"multiple pruning hooks on {name}")
        first_time = not previous
        container = None
        if previous:
            old = module._forward_pre_hooks.pop(previous[0])
            container = old if isinstance(old, PruningContainer) else PruningContainer(old)
            container.add_pruning_method(method)
        orig = getattr(module, name)
        scores = orig if importance_scores is None else importance_scores
        if scores.shape != orig.shape:
            raise AssertionError("importance_scores must have the parameter's shape")
        if first_time:
            module.register_parameter(name + "_orig", orig)
            del module._parameters[name]
            default_mask = torch.ones_like(orig, dtype=torch.bool)
        else:
            default_mask = getattr(module, name + "_mask").detach().clone(memory_format=torch.contiguous_format)
        hook = container if container is not None else method
        try:
            mask = hook.compute_mask(scores, default_mask=default_mask)
            module.register_buffer(name + "_mask", mask)
            module.register_buffer(name, hook.apply_mask(module), persistent=False)
            module.register_forward_pre_hook(hook)
        except Exception:
            if first_time:
                module.register_parameter(name, getattr(module, name + "_orig"))
                del module._parameters[name + "_orig"]
            raise
        return hook

    def remove(self, module):
        """Bake the mask in; ``name`` becomes a plain Parameter again (same object as ``name_orig``)."""
        name = self._tensor_name
        weight = self.apply_mask(module)
        module._buffers.pop(name, None)
        orig = module._parameters.pop(name + "_orig")
        # in place: the Parameter's storage is a view into the flat parameter buffer the fused optimizer
        # updates (parallel.FlatBuffers); re-pointing ``orig.data`` (what torch.nn.utils.prune does) would
        # orphan that slot and silently freeze the tensor
        with torch.no_grad():
            orig.data.copy_(weight.detach())
        del module._buffers[name + "_mask"]
        module.register_parameter(name, orig)


class PruningContainer(BasePruningMethod):
    def __init__(self, *methods):
        self._pruning_methods = tuple()
        for m in methods:
            if self._tensor_name is None:
                self._tensor_name = m._tensor_name
            self.add_pruning_method(m)

    def add_pruning_method(self, method):
        if method is None:
            return
        if not isinstance(method, BasePruningMethod):
            raise TypeError(f"{type(method)} is not a BasePruningMethod subclass")
        if self._tensor_name != method._tensor_name:
            raise ValueError(f"container acts on '{self._tensor_name}', method on '{method._tensor_name}'")
        self._pruning_methods += (method,)

    def __len__(self):
        return len(self._pruning_methods)

    def __iter__(self):
        return iter(self._pruning_methods)

    def __getitem__(self, idx):
        return self._pruning_methods[idx]

    def compute_mask(self, t, default_mask):
        method = self._pruning_methods[-1]
        mask = default_mask.bool()
        if method.PRUNING_TYPE == "unstructured":
            live = mask.clone()
            part = method.compute_mask(t[live], default_mask=mask[live])
            mask[live] = part.bool()
        elif method.PRUNING_TYPE == "global":
            mask = method.compute_mask(t, default_mask=mask).bool()
        else:
            raise ValueError(f"unsupported PRUNING_TYPE {method.PRUNING_TYPE}")
        return mask


class Identity(BasePruningMethod):
    def compute_mask(self, t, default_mask):
        return default_mask

    @classmethod
    def apply(cls, module, name):
        return super().apply(module, name)


def _n_to_prune(amount, size):
    if isinstance(amount, numbers.Integral):
        n = int(amount)
    else:
        if not 0.0 <= float(amount) <= 1.0:
            raise ValueError(f"amount={amount} should be a float in [0, 1] or a non-negative int")
        n = int(round(float(amount) * size))
    if n < 0 or n > size:
        raise ValueError(f"amount={amount} out of range for a tensor of {size} elements")
    return n


class L1Unstructured(BasePruningMethod):
    """Zero the ``amount`` entries of smallest magnitude (reference prune.py:553-573)."""

    def __init__(self, amount):
        self.amount = amount

    def compute_mask(self, t, default_mask):
        k = _n_to_prune(self.amount, t.nelement())
        mask = default_mask.clone(memory_format=torch.contiguous_format)
        if k:
            idx = t.abs().view(-1).topk(k=k, largest=False).indices
            mask.view(-1)[idx] = False
        return mask

    @classmethod
    def apply(cls, module, name, amount, importance_scores=None):
        return super().apply(module, name, amount=amount, importance_scores=importance_scores)


class CustomFromMask(BasePruningMethod):
    PRUNING_TYPE = "global"

    def __init__(self, mask):
        self.mask = mask

    def compute_mask(self, t, default_mask):
        if default_mask.shape != self.mask.shape:
            raise AssertionError("mask shape mismatch")
        return default_mask.bool() & self.mask.bool()

    @classmethod
    def apply(cls, module, name, mask):
        return super().apply(module, name, mask=mask)


def identity(module, name):
    Identity.apply(module, name)
    return module


def l1_unstructured(module, name, amount, importance_scores=None):
    L1Unstructured.apply(module, name, amount=amount, importance_scores=importance_scores)
    return module


def custom_from_mask(module, name, mask):
    CustomFromMask.apply(module, name, mask)
    return module


def _global_l1_masks_gpu(tensors, old_masks, k):
    """Exact global k-smallest-|w| selection on the device (radix select + ordered ties)."""
    from .. import kernels as K

    masks = [m.to(torch.bool).contiguous().clone() for m in old_masks]
    if k == 0:
        return masks
    flat = [t.detach().contiguous().float() for t in tensors]
    res = K.abs_kth_smallest(flat, k)
    K.apply_threshold_masks(flat, masks, res, k)
    return masks


def global_unstructured(parameters, pruning_method, importance_scores=None, **kwargs):
    """Reference prune.py:1049-1171: one magnitude ranking over the concatenation of all
    ``(module, name)`` tensors, new mask AND-ed with the existing one, installed per tensor
    through ``custom_from_mask``."""
    parameters = list(parameters)
    importance_scores = importance_scores or {}
    if not isinstance(importance_scores, dict):
        raise TypeError("global_unstructured(): importance_scores must be of type dict")
    method = pruning_method(**kwargs)
    if method.PRUNING_TYPE != "unstructured":
        raise TypeError(f'Only "unstructured" PRUNING_TYPE supported, found {method.PRUNING_TYPE}')
    scores = [importance_scores.get((m, n), getattr(m, n)) for m, n in parameters]
    old = [getattr(m, n + "_mask", None) for m, n in parameters]
    old = [torch.ones_like(s, dtype=torch.bool) if o is None else o for o, s in zip(old, scores)]
    if isinstance(method, Identity):
        new = old
    elif isinstance(method, L1Unstructured) and scores[0].is_cuda:
        total = sum(s.numel() for s in scores)
        new = _global_l1_masks_gpu(scores, old, _n_to_prune(method.amount, total))
    else:
        # generic path (CPU tensors / other strategies): same calls as the reference
        vec = torch.cat([s.detach().reshape(-1) for s in scores])
        default = torch.cat([o.reshape(-1) for o in old])
        box = PruningContainer()
        box._tensor_name = method._tensor_name = "temp"
        box.add_pruning_method(method)
        final = box.compute_mask(vec, default)
        new, ptr = [], 0
        for s in scores:
            new.append(final[ptr:ptr + s.numel()].view_as(s))
            ptr += s.numel()
    for (module, name), mask in zip(parameters, new):
        custom_from_mask(module, name, mask=mask)


def remove(module, name):
    for k, hook in module._forward_pre_hooks.items():
        if isinstance(hook, BasePruningMethod) and hook._tensor_name == name:
            hook.remove(module)
            del module._forward_pre_hooks[k]
            return module
    raise ValueError(f"Parameter '{name}' of module {module} has to be pruned before pruning can be removed")


def is_pruned(module):
    for _, sub in module.named_modules():
        for hook in sub._forward_pre_hooks.values():
            if isinstance(hook, BasePruningMethod):
                return True
    return False
