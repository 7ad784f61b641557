"""L1 attention-head pruning -- drop-in for reference ``head_pruning/hp_utils.py:9-369``
(``metric: l1``; the data-driven scorer is dead code in the reference).

Scores: sum |W[h]| + sum |b[h]| over the q, k and v projections.  The reference computes them
with 3 * heads tiny ``torch.sum(...).tolist()`` device syncs per layer; here one fp64 reduction
kernel per projection (``mh_row_abs_sums``) produces per-row sums that are folded per head on
the host.  Selection (stable sort, ``by_layer`` / ``by_whole``), the ``Pruned_heads`` record
and the physical slicing follow the reference exactly.
"""
import os

import torch
from tqdm import tqdm

from .. import ckpt
from .. import kernels as K
from ..fairseq_code import MultiheadAttention
from ..surgery import drop_heads


def set_prune_interval(prune_interval, warm_up_steps, total_prune_steps):
    if isinstance(prune_interval, int):
        return [warm_up_steps + prune_interval * i for i in range(total_prune_steps)]
    if isinstance(prune_interval, list):
        return [warm_up_steps + p for p in prune_interval]
    raise NotImplementedError


def _row_l1(t):
    """per-row L1 norms as a Python list of doubles (GPU kernel for CUDA tensors)."""
    t = t.detach()
    if t.is_cuda:
        t2 = t.float().contiguous().view(t.shape[0], -1)
        return K.row_abs_sums(t2).cpu().tolist()
    return t.double().abs().view(t.shape[0], -1).sum(1).tolist()


class HeadPruningTools:
    def __init__(self, args, runner_config, upstream_config, upstream):
        self.args, self.runner_config, self.upstream_config, self.upstream = args, runner_config, upstream_config, upstream
        self.num_layers = len(upstream.model.encoder.layers)
        metric = runner_config["prune"]["metric"]
        if metric == "l1":
            self.num_heads_each_step = self.num_layers
        elif metric == "data-driven":
            raise NotImplementedError("data-driven head scoring is disabled in the reference itself (hp_utils.py:58-59)")
        else:
            raise NotImplementedError
        self.total_heads = sum(l.self_attn.num_heads for l in upstream.model.encoder.layers)
        self.total_prune_step = runner_config["prune"]["total_steps"]
        assert self.num_heads_each_step * self.total_prune_step <= self.total_heads
        self.pruned_heads = []

    def prune_api(self):
        self.prune()
        self.total_heads -= self.num_heads_each_step
        cur = sum(l.self_attn.num_heads for l in self.upstream.model.encoder.layers)
        assert cur == self.total_heads
        tqdm.write(f"[Head Pruning] {self.total_heads} heads are remained")

    def get_layer_heads_norm(self, mha, layer):
        assert isinstance(mha, MultiheadAttention)
        hd, n = mha.head_dim, mha.num_heads
        per_proj = []
        for proj in (mha.k_proj, mha.q_proj, mha.v_proj):  # reference adds k + q + v in this order
            wr = _row_l1(proj.weight)
            br = proj.bias.detach().double().abs().cpu().tolist()
            per_proj.append([sum(wr[h * hd:(h + 1) * hd]) + sum(br[h * hd:(h + 1) * hd]) for h in range(n)])
        return [((layer, h), per_proj[0][h] + per_proj[1][h] + per_proj[2][h]) for h in range(n)]

    def get_heads_norm(self, encoder):
        out = []
        for layer in range(self.num_layers):
            out.extend(self.get_layer_heads_norm(encoder.layers[layer].self_attn, layer))
        return out

    def prune(self):
        n_to_prune = self.num_heads_each_step
        heads_and_score = self.get_heads_norm(self.upstream.model.encoder)
        ckpt.save(heads_and_score, os.path.join(self.args.expdir, f"heads_and_score_{self.total_heads}.ckpt"))
        ranked = [hs[0] for hs in sorted(heads_and_score, key=lambda x: x[1])]  # stable, ascending score
        target = self.runner_config["prune"]["target"]
        if target == "by_whole":
            # protect the best head of every layer, then take the n lowest of the rest
            guard = {l: 1 for l in range(self.num_layers)}
            rest = []
            for layer, head in reversed(ranked):
                if layer in guard:
                    if guard[layer] > 0:
                        guard[layer] -= 1
                        continue
                    guard.pop(layer)
                rest.insert(0, (layer, head))
            assert len(rest) >= n_to_prune
            to_prune = rest[:n_to_prune]
        elif target == "by_layer":
            assert len(ranked) >= n_to_prune
            want = set(range(n_to_prune))
            to_prune = []
            for layer, head in ranked:
                if not want:
                    break
                if layer in want:
                    to_prune.append((layer, head))
                    want.remove(layer)
        else:
            raise NotImplementedError(target)
        group = {}
        for layer, head in to_prune:
            group[layer] = group.get(layer, []) + [head]
        tqdm.write(f"[Head Pruning] - These heads are pruned:{group}")
        self.pruned_heads.append(group)
        self.upstream.pruned_heads = self.pruned_heads
        for idx, layer in enumerate(self.upstream.model.encoder.layers):
            if idx in group:
                self.prune_layer_heads(layer.self_attn, group[idx])

    def prune_layer_heads(self, mha, heads):
        drop_heads(mha, heads)

    def save_model(self, optimizer, global_step):
        states = {"Optimizer": optimizer.state_dict(), "Step": global_step, "Args": self.args,
                  "Runner": self.runner_config, "Pruned_heads": self.pruned_heads}
        states = self.upstream.add_state_to_save(states)
        path = os.path.join(self.args.expdir, f"states_prune_{self.total_heads}.ckpt")
        tqdm.write(f"[Head Pruning] - Save the checkpoint to: {path}")
        tqdm.write("[Head Pruning] - Number of parameters saved: " + str(sum(p.numel() for p in states["model"].values())))
        ckpt.save(states, path)
