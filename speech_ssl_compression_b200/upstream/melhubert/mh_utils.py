"""Checkpoint writer for plain pre-training / distillation (reference
``upstream/melhubert/mh_utils.py:6-29``): same dict keys and file names."""
import os

import torch
from tqdm import tqdm

from ... import ckpt


class MelHuBERTTools:
    def __init__(self, args, runner_config, upstream_config, upstream):
        self.args, self.runner_config = args, runner_config
        self.upstream_config, self.upstream = upstream_config, upstream
        self.save_every_x_epochs = runner_config["runner"].get("save_every_x_epochs")
        assert self.save_every_x_epochs, "Must specify an integer for save_every_x_epochs to save model"

    def save_model(self, optimizer, global_step, num_epoch=-1, name=None):
        if global_step == 0:
            return
        states = {"Optimizer": optimizer.state_dict(), "Step": global_step, "Args": self.args,
                  "Runner": self.runner_config}
        states = self.upstream.add_state_to_save(states)
        path = os.path.join(self.args.expdir, name or f"checkpoint-epoch-{num_epoch}.ckpt")
        tqdm.write(f"[MelHuBERT] - Save the checkpoint to: {path}")
        ckpt.save(states, path)
