"""FFN row pruning -- drop-in for reference ``row_pruning/rp_utils.py:8-129``.

score(i) = sum |fc1.weight[i, :]| + |fc1.bias[i]| + sum |fc2.weight[:, i]|; the lowest
``num_rows_each_step`` rows of every layer are sliced out of fc1 (rows, bias) and fc2
(columns).  Two fp64 reduction kernels per layer replace the reference's 2 * ffn_dim
``.tolist()`` device syncs.
"""
import os

import torch
from tqdm import tqdm

from .. import ckpt
from .. import kernels as K
from ..head_pruning.hp_utils import set_prune_interval  # noqa: F401  (same helper in the reference)
from ..surgery import drop_ffn_rows


class RowPruningTools:
    def __init__(self, args, runner_config, upstream_config, upstream):
        self.args, self.runner_config, self.upstream_config, self.upstream = args, runner_config, upstream_config, upstream
        self.num_layers = len(upstream.model.encoder.layers)
        self.num_rows_each_step = runner_config["prune"]["num_rows_each_step"]
        self.total_ffn_dim = upstream.model.encoder.layers[0].fc1.weight.shape[0]
        self.total_prune_step = runner_config["prune"]["total_steps"]
        assert self.num_rows_each_step * self.total_prune_step <= upstream.model.encoder.ffn_embedding_dim

    def prune_api(self):
        self.prune(self.upstream.model.encoder)
        self.total_ffn_dim -= self.num_rows_each_step
        self.upstream.model.encoder.ffn_embedding_dim = self.total_ffn_dim
        self.upstream.upstream_config["melhubert"]["encoder_ffn_embed_dim"] = self.total_ffn_dim
        tqdm.write(f"[Row Pruning] {self.total_ffn_dim} hidden dimension are remained in fead forward network")

    def get_layer_rows_norm(self, fc1, fc2, layer):
        w1, b1, w2 = fc1.weight.detach(), fc1.bias.detach(), fc2.weight.detach()
        if w1.is_cuda:
            r1 = K.row_abs_sums(w1.float().contiguous()).cpu().tolist()
            c2 = K.col_abs_sums(w2.float().contiguous()).cpu().tolist()
        else:
            r1 = w1.double().abs().sum(1).tolist()
            c2 = w2.double().abs().sum(0).tolist()
        bb = b1.double().abs().cpu().tolist()
        return [(i, (r1[i] + bb[i]) + c2[i]) for i in range(len(r1))]

    def prune(self, encoder):
        for layer in range(self.num_layers):
            scored = sorted(self.get_layer_rows_norm(encoder.layers[layer].fc1, encoder.layers[layer].fc2, layer),
                            key=lambda x: x[1])
            self.prune_layer_ffn(encoder.layers[layer], [i for i, _ in scored[:self.num_rows_each_step]])

    def prune_layer_ffn(self, layer, to_prune):
        drop_ffn_rows(layer, to_prune)

    def save_model(self, optimizer, global_step):
        states = {"Optimizer": optimizer.state_dict(), "Step": global_step, "Args": self.args, "Runner": self.runner_config}
        states = self.upstream.add_state_to_save(states)
        path = os.path.join(self.args.expdir, f"states_prune_{self.total_ffn_dim}.ckpt")
        tqdm.write(f"[Row Pruning] - Save the checkpoint to: {path}")
        tqdm.write("[Row Pruning] - Number of parameters saved: " + str(sum(p.numel() for p in states["model"].values())))
        ckpt.save(states, path)
