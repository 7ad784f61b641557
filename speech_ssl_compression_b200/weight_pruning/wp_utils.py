"""Iterative global magnitude pruning -- drop-in for reference
``weight_pruning/wp_utils.py:12-183`` (same schedule, convergence gate, checkpoint keys and
file names).  The global k-smallest-|w| selection (reference: one ``torch.topk`` over 85 M
values, 11.5 s on 8 CPU cores) is an exact radix select on the GPU; the per-forward mask
application the reference does through 144 pre-hooks is folded into ``mh_weight_prep``.
"""
import os
import random
import re

import numpy as np
import torch
from tqdm import tqdm

from .. import ckpt
from .. import ops
from ..pytorch_code import prune

_PRUNABLE = ("q_proj", "k_proj", "v_proj", "out_proj")


def get_params_to_prune(upstream, bias=True):
    """Per layer: q, k, v, out, fc1, fc2 weights, then (optionally) the same six biases --
    the order defines the flat index used by the global ranking (reference :13-48)."""
    model = getattr(upstream, "module", upstream)
    params = []
    for layer in model.encoder.layers:
        mods = [getattr(layer.self_attn, n) for n in _PRUNABLE] + [layer.fc1, layer.fc2]
        params += [(m, "weight") for m in mods]
        if bias:
            params += [(m, "bias") for m in mods]
    pattern = re.compile(r".*encoder\.layers\.[0-9]+\.((self_attn\.([qkv]|out)_proj)|fc[12])\.weight")
    return tuple(params), (lambda name: pattern.fullmatch(name))


def _resume_random_state(state):
    if state:
        random.setstate(state["random"])
        np.random.set_state(state["numpy"])
        torch.set_rng_state(state["torch"])
        if torch.cuda.is_available() and state.get("torch.cuda") is not None:
            torch.cuda.set_rng_state(state["torch.cuda"])


class WeightPruningTools:
    def __init__(self, args, runner_config, upstream_config, upstream, initial_weight):
        self.args, self.runner_config, self.upstream_config = args, runner_config, upstream_config
        self.upstream, self.initial_weight = upstream, initial_weight
        pc = runner_config["prune"]
        self.prune_condition, self.prune_strategy = pc["pruning_condition"], pc["strategy"]
        self.n_iters = pc.get("n_iters", 38)
        self.warnup, self.period = pc.get("warnup", 25000), pc.get("period", 25000)
        assert self.warnup > 0 and self.period > 0, "Do not set warnup and period to 0."
        self.avg_len = pc.get("average_length", 15000)
        self.con_tol = pc.get("converge_loss_tolerance", 0.001)
        sp = pc["sparsity"]
        if isinstance(sp, float):
            self.sparsity = [sp * (n + 1) / self.n_iters for n in range(self.n_iters)]
        elif isinstance(sp, list):
            self.sparsity = sp
        else:
            raise NotImplementedError
        self.prune_steps = list(self.warnup + (np.arange(self.n_iters) * self.period))
        self.smooth_loss, self.tgt_smooth_loss = None, -float("inf")
        self.smooth_factor = pc.get("smooth_factor", 0.999)
        self.buffer_loss, self.pruning_times = [], 0
        params, _ = get_params_to_prune(self.upstream.model)
        if not prune.is_pruned(self.upstream.model):
            prune.global_unstructured(params, pruning_method=prune.Identity)
        if self.initial_weight:
            states = torch.load(self.initial_weight, map_location="cpu", weights_only=False)
            if "Pruning" in states:
                self.smooth_loss = states["Pruning"]["smooth_loss"]
                self.tgt_smooth_loss = states["Pruning"]["tgt_smooth_loss"]
                self.pruning_times = states["Pruning"]["pruning_times"]
            if "RandomState" in states:
                _resume_random_state(states["RandomState"])
        print("=" * 40 + "\n[Weight Pruning] - Pruning-related hyperparameters:")
        print(f"Pruning iterations: {self.n_iters}\nWarnup steps: {self.warnup}\nPruning steps: {self.prune_steps}")
        print("=" * 40)

    def update_smooth_loss(self, batch_loss):
        if self.smooth_loss is not None:
            self.smooth_loss = self.smooth_loss * self.smooth_factor + batch_loss * (1 - self.smooth_factor)
        elif len(self.buffer_loss) == 3:
            self.smooth_loss = sum(self.buffer_loss) / 3
            self.buffer_loss = []
        else:
            self.buffer_loss.append(batch_loss)

    def update_target_smooth_loss(self, global_step):
        if (self.prune_condition == "converge" and global_step > self.warnup
                and (global_step - self.warnup + self.avg_len) in self.prune_steps):
            self.tgt_smooth_loss = self.smooth_loss

    def prune_api(self, optimizer, global_step, total_step):
        if (self.prune_condition == "converge" and self.smooth_loss is not None
                and self.tgt_smooth_loss - self.con_tol > self.smooth_loss):
            tqdm.write("[Weight Pruning] - Not converge, keep training")
            return "not-converge"
        prefix = "mask-" if prune.is_pruned(self.upstream.model) else ""
        cur = 0 if self.pruning_times == 0 else self.sparsity[self.pruning_times - 1]
        self._save(optimizer, global_step, total_step, f"{prefix}before-pruning-states-{global_step}-sparsity-{cur}.ckpt")
        params, _ = get_params_to_prune(self.upstream.model)
        amount = self.sparsity[self.pruning_times]
        for module, name in params:  # bake zeros in (same Parameter objects: optimizer state survives)
            prune.remove(module, name)
        prune.global_unstructured(params, pruning_method=getattr(prune, self.prune_strategy), amount=amount)
        ops.bump_weight_epoch()
        tqdm.write(f"[Weight Pruning] - {self.pruning_times + 1} iters of pruning at {global_step} steps")
        self.pruning_times += 1
        self.smooth_loss = None
        return "pruned"

    def _save(self, optimizer, global_step, total_step, filename):
        states = {
            "Optimizer": optimizer.state_dict(), "Step": global_step, "TotalStep": total_step, "Args": self.args,
            "Runner": self.runner_config,
            "Pruning": {"smooth_loss": self.smooth_loss, "tgt_smooth_loss": self.tgt_smooth_loss,
                        "pruning_times": self.pruning_times},
            "RandomState": {"random": random.getstate(), "numpy": np.random.get_state(), "torch": torch.get_rng_state(),
                            "torch.cuda": torch.cuda.get_rng_state() if torch.cuda.is_available() else None},
        }
        states = self.upstream.add_state_to_save(states)
        path = os.path.join(self.args.expdir, filename)
        tqdm.write(f"[Weight Pruning] - Save the checkpoint to: {path}")
        ckpt.save(states, path)
