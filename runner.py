"""Training-loop host for the five MelHuBERT modes -- the caller of the hot path (reference ``runner.py:36-461``).

Keeps the reference's CLI / yaml schema, prune scheduling and checkpoint names; the step itself (forward, backward,
gradient all-reduce, /n, clip, Adam) is ``trainer.TrainStep`` on the sm_100a kernels.  Intended deviations, all listed in
SURVEY appendix A: ``datarc`` / ``optimizer`` are read where the shipped yamls put them (Q3, Q4); gradients are divided
by the number of accumulated micro-batches, not by a counter that only resets at ``log_step`` (Q5); scalars go to a CSV
(tensorboardX is not installed); ``--synthetic`` feeds synthetic log-mel buckets of the dataset's tuple layout
(LibriSpeech is not available offline)."""
import csv
import math
import os

import numpy as np
import torch
import yaml
from tqdm import tqdm


class SyntheticBuckets:
    """Batches with the layout of ``MelFeatDataset.__getitem__`` (melhubert_dataset.py:120):
    (feat f32 (B,T,D), label i64 (B,T) with -100 at pads, pad_mask f32 (B,T), lens)."""

    def __init__(self, B, T, D, n=64, seed=1337):
        self.B, self.T, self.D, self.n, self.seed = B, T, D, n, seed

    def __len__(self):
        return self.n

    def __iter__(self):
        from bench import synth_host_batch

        for i in range(self.n):
            yield synth_host_batch(self.B, self.T, self.D, seed=self.seed + i)


class CsvNpyBuckets:
    """Length-sorted buckets over the reference's preprocessed csv (file_path,label_path,length) + .npy files
    (datasets/melhubert_dataset.py:17-120): 20 ms frame stacking, random crop to ``sequence_length``, suffix padding."""

    def __init__(self, datarc, task, frame_period, B):
        import pandas as pd

        root = datarc["sets"] if isinstance(datarc["sets"], list) else [datarc["sets"]]
        table = pd.concat([pd.read_csv(s) for s in root], ignore_index=True).sort_values("length", ascending=False)
        self.files = list(zip(table["file_path"], table["label_path"]))
        self.B, self.fp, self.seq = B, frame_period, int(task["sequence_length"])
        self.buckets = [self.files[i:i + B] for i in range(0, len(self.files), B)]

    def __len__(self):
        return len(self.buckets)

    def __iter__(self):
        import random

        order = torch.randperm(len(self.buckets)).tolist()
        for bi in order:
            feats, labels = [], []
            for fpath, lpath in self.buckets[bi]:
                x, y = np.load(fpath), np.load(lpath)
                if self.fp == 20:
                    x = x[: (len(x) // 2) * 2]
                    x = np.concatenate([x[0::2], x[1::2]], axis=1)
                    y = y[: len(x)]
                if len(x) > self.seq:
                    s = random.randint(0, len(x) - self.seq)
                    x, y = x[s:s + self.seq], y[s:s + self.seq]
                feats.append(torch.from_numpy(x).float())
                labels.append(torch.from_numpy(y).long())
            lens = [len(f) for f in feats]
            T = max(lens)
            feat = torch.zeros(len(feats), T, feats[0].shape[1])
            label = torch.full((len(feats), T), -100, dtype=torch.long)
            pad = torch.zeros(len(feats), T)
            for i, (f, y) in enumerate(zip(feats, labels)):
                feat[i, :lens[i]], label[i, :lens[i]], pad[i, :lens[i]] = f, y, 1
            yield feat.pin_memory(), label.pin_memory(), pad.pin_memory(), lens


class Runner:
    def __init__(self, args, runner_config):
        self.args, self.runner_config = args, runner_config
        self.upstream_config = yaml.load(open(args.upstream_config), Loader=yaml.FullLoader)
        fdim = self.upstream_config["melhubert"]["feat_emb_dim"]
        assert fdim == (80 if args.frame_period == 20 else 40), \
            f"Feature embedding dimension should be {80 if args.frame_period == 20 else 40} when the frame period is {args.frame_period}"
        from speech_ssl_compression_b200.upstream.melhubert.mh_utils import MelHuBERTTools
        from speech_ssl_compression_b200.upstream.melhubert.pretrain_expert import MelHuBERTPretrainer

        mode = args.mode
        if mode == "distillation":
            from speech_ssl_compression_b200.distillation.pretrain_expert import MelHuBERTDistiller

            self.expert = MelHuBERTDistiller(self.upstream_config, args.initial_weight, args.device, args.multi_gpu).to(args.device)
        else:
            self.expert = MelHuBERTPretrainer(self.upstream_config, args.initial_weight, args.device, args.multi_gpu).to(args.device)
        for need in ("forward", "load_model", "add_state_to_save"):
            assert hasattr(self.expert, need)
        self.tools, self.prune_steps = None, []
        pc = runner_config.get("prune", {})
        if mode in ("melhubert", "distillation"):
            self.tools = MelHuBERTTools(args, runner_config, self.upstream_config, self.expert)
        elif mode == "weight-pruning":
            from speech_ssl_compression_b200.weight_pruning.wp_utils import WeightPruningTools

            self.tools = WeightPruningTools(args, runner_config, self.upstream_config, self.expert, args.initial_weight)
            self.prune_steps = [int(s) for s in self.tools.prune_steps]
        elif mode in ("head-pruning", "row-pruning"):
            from speech_ssl_compression_b200.head_pruning.hp_utils import HeadPruningTools, set_prune_interval
            from speech_ssl_compression_b200.row_pruning.rp_utils import RowPruningTools

            cls = HeadPruningTools if mode == "head-pruning" else RowPruningTools
            self.tools = cls(args, runner_config, self.upstream_config, self.expert)
            self.prune_steps = set_prune_interval(pc["interval"], pc["warm_up"], pc["total_steps"])
            assert len(self.prune_steps) == pc["total_steps"]
        else:
            raise SystemExit("We do not support this mode currently.")
        self.rank = int(os.environ.get("RANK", "0"))

    # ------------------------------------------------------------------------------------------
    def _batches(self):
        rc = self.runner_config
        datarc = rc.get("datarc") or rc.get("pretrain_expert", {}).get("datarc", {})
        B = int(datarc.get("train_batch_size", 4))
        task = self.upstream_config.get("task", {"sequence_length": 750 if self.args.frame_period == 20 else 1500})
        D = self.upstream_config["melhubert"]["feat_emb_dim"]
        if getattr(self.args, "synthetic", False) or not datarc.get("sets"):
            return SyntheticBuckets(B, int(task["sequence_length"]), D, seed=self.args.seed + 97 * self.rank), B, int(task["sequence_length"]), D
        return CsvNpyBuckets(datarc, task, self.args.frame_period, B), B, int(task["sequence_length"]), D

    def _new_step(self, B, T, D):
        from speech_ssl_compression_b200.trainer import TrainStep

        oc = self.runner_config.get("optimizer", {})
        rc = self.runner_config.get("runner", {})
        return TrainStep(self.expert, B, T, D, lr=float(oc.get("lr", 1e-3)), betas=tuple(oc.get("betas", (0.9, 0.999))),
                         eps=float(oc.get("eps", 1e-8)), weight_decay=float(oc.get("weight_decay", 0.0)),
                         max_norm=float(rc.get("gradient_clipping", 10.0)), use_graph=False)

    def train(self):
        rc = self.runner_config.get("runner", {})
        total = int(getattr(self.args, "max_steps", None) or rc.get("total_steps", 1000))
        log_step = int(rc.get("log_step", 100))
        data, B, T, D = self._batches()
        self.expert.train()
        step = self._new_step(B, T, D)
        log = None
        if self.rank == 0:
            log = csv.writer(open(os.path.join(self.args.expdir, "train_log.csv"), "a", newline=""))
        gs, bar, mode = 0, tqdm(total=total, dynamic_ncols=True, desc="overall", disable=self.rank != 0), self.args.mode
        prune_idx = 0
        while gs < total:
            for feat, label, pad, lens in data:
                if gs >= total:
                    break
                if prune_idx < len(self.prune_steps) and gs == self.prune_steps[prune_idx]:
                    if mode == "weight-pruning":
                        state = self.tools.prune_api(step.opt, gs, total)
                        if state == "not-converge":
                            self.prune_steps = self.prune_steps[:prune_idx] + [s + self.tools.period for s in self.prune_steps[prune_idx:]]
                            prune_idx -= 1
                    else:
                        self.tools.save_model(step.opt, gs)
                        self.tools.prune_api()
                        step = self._new_step(B, T, D)  # shapes changed: new flat buffers / Adam state (runner.py:348,355)
                    prune_idx += 1
                if feat.shape[1] != step.T or feat.shape[0] != step.B:  # ragged last bucket / shorter bucket
                    f2 = torch.zeros(step.B, step.T, D).pin_memory(); l2 = torch.full((step.B, step.T), -100).pin_memory()
                    p2 = torch.zeros(step.B, step.T).pin_memory()
                    b, t = min(feat.shape[0], step.B), min(feat.shape[1], step.T)
                    f2[:b, :t], l2[:b, :t], p2[:b, :t] = feat[:b, :t], label[:b, :t], pad[:b, :t]
                    lens = [min(int(x), t) for x in lens[:b]] + [1] * (step.B - b)
                    p2[b:, :1] = 1
                    feat, label, pad = f2, l2, p2
                step.load_batch(feat, label, pad, lens)
                step.run()
                gs += 1
                bar.update(1)
                if gs % log_step == 0 or gs == total:
                    loss = step.read_loss()
                    if mode == "weight-pruning":
                        self.tools.update_smooth_loss(loss)
                        self.tools.update_target_smooth_loss(gs)
                    if math.isnan(loss):
                        tqdm.write(f"[Runner] - loss is NaN at step {gs}")
                    if log is not None:
                        log.writerow([gs, f"{mode}/train-loss", loss, f"{mode}/train-gradient norm", step.opt.grad_norm()])
        bar.close()
        if self.rank == 0 and mode == "weight-pruning":
            self.tools._save(step.opt, gs, total, "last-step.ckpt")  # + Pruning / RandomState / TotalStep (wp_utils.py:162-180)
        elif self.rank == 0:
            states = {"Optimizer": step.opt.state_dict(), "Step": gs, "Args": self.args, "Runner": self.runner_config}
            states = self.expert.add_state_to_save(states)
            torch.save(states, os.path.join(self.args.expdir, "last-step.ckpt"))
            tqdm.write(f"[Runner] - saved {os.path.join(self.args.expdir, 'last-step.ckpt')}")
