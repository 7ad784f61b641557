#!/usr/bin/env python
"""Training entry point -- same flags as the reference's ``train.py:19-46`` (plus ``-m melhubert``, which the
reference's Runner implements but its argparse forgets, SURVEY Q1, and ``--synthetic`` for machines without
LibriSpeech).  Multi-GPU: ``torchrun --nproc-per-node N train.py ... --multi_gpu`` (one process per GPU)."""
import argparse
import os
import random
import sys
from shutil import copyfile

import numpy as np
import torch
import yaml

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def get_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("-c", "--runner_config", help="yaml of the experiment (everything but the upstream model)")
    ap.add_argument("-g", "--upstream_config", help="yaml of the upstream model")
    ap.add_argument("-n", "--expdir", help="save the experiment at this path")
    ap.add_argument("-m", "--mode", choices=["melhubert", "weight-pruning", "head-pruning", "row-pruning", "distillation"])
    ap.add_argument("-f", "--frame_period", default=20, choices=[10, 20], type=int)
    ap.add_argument("-u", "--upstream", default="melhubert", choices=["melhubert"], type=str)
    ap.add_argument("-i", "--initial_weight", help="initial / teacher weights")
    ap.add_argument("--init_optimizer_from_initial_weight", action="store_true")
    ap.add_argument("--seed", default=1337, type=int)
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--multi_gpu", action="store_true")
    ap.add_argument("--synthetic", action="store_true", help="synthetic log-mel batches instead of the csv/npy dataset")
    ap.add_argument("--max_steps", type=int, default=None, help="override runner.total_steps")
    ap.add_argument("--no_graph", action="store_true", help="run the step eagerly instead of replaying its CUDA graph")
    args = ap.parse_args(argv)
    os.makedirs(args.expdir, exist_ok=True)
    assert args.runner_config and args.upstream_config, "Please specify .yaml config files."
    with open(args.runner_config) as f:
        runner_config = yaml.load(f, Loader=yaml.FullLoader)
    if int(os.environ.get("RANK", "0")) == 0:
        copyfile(args.runner_config, f"{args.expdir}/config_runner.yaml")
        copyfile(args.upstream_config, f"{args.expdir}/config_model.yaml")
    return args, runner_config


def main(argv=None):
    args, runner_config = get_args(argv)
    random.seed(args.seed)
    np.random.seed(args.seed)
    torch.manual_seed(args.seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(args.seed)
    from runner import Runner

    Runner(args, runner_config).train()


if __name__ == "__main__":
    main()
