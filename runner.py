"""Training-loop host for the five MelHuBERT modes -- the caller of the hot path (reference ``runner.py:36-461``).

Keeps the reference's CLI / yaml schema, prune scheduling and checkpoint names; the step itself (forward, backward,
gradient all-reduce, /n, clip, Adam) is ``trainer.TrainStep`` on the sm_100a kernels.  Intended deviations, all listed in
SURVEY appendix A: ``datarc`` / ``optimizer`` are read where the shipped yamls put them (Q3, Q4); the accumulated
gradient is divided by ``gradient_accumulate_steps`` (``loss / accum``, runner.py:370-371) and NOT additionally by the
reference's ``all_sample_size`` counter, which only resets at ``log_step`` and so grows from 1 to ``log_step`` (Q5);
scalars go to a CSV (tensorboardX is not installed); ``--synthetic`` feeds synthetic log-mel buckets of the dataset's
tuple layout (LibriSpeech is not available offline).

The loop runs the benchmarked path: the step is CUDA-graph captured (re-captured after every prune event), the next
micro-batch is staged on a copy stream while the current one runs, and the loss is read back one optimizer step late
 -- except right before a weight-pruning event, whose convergence gate needs the freshest smoothed loss.
Multi-process data parallelism (``--multi_gpu`` under torchrun): every rank walks its own shard of the bucket
permutation with its own crop / span-mask streams; rank 0 alone writes checkpoints."""
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

    def __init__(self, datarc, task, frame_period, B, rank=0, world=1):
        import pandas as pd

        self.rank, self.world = rank, world

        root = datarc["sets"] if isinstance(datarc["sets"], list) else [datarc["sets"]]
        table = pd.concat([pd.read_csv(s) for s in root], ignore_index=True).sort_values("length", ascending=False)
        self.files = list(zip(table["file_path"], table["label_path"]))
        self.B, self.fp, self.seq = B, frame_period, int(task["sequence_length"])
        self.buckets = [self.files[i:i + B] for i in range(0, len(self.files), B)]

    def __len__(self):
        return len(self.buckets) // self.world

    def __iter__(self):
        import random

        # one permutation shared by all ranks (same torch seed), dealt round-robin: equal count per rank, no overlap
        order = torch.randperm(len(self.buckets)).tolist()
        order = order[: len(order) // self.world * self.world][self.rank::self.world]
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
            if torch.cuda.is_available():
                feat, label, pad = feat.pin_memory(), label.pin_memory(), pad.pin_memory()
            yield feat, label, pad, lens


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
        self.rank, self.world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
        if not args.multi_gpu:
            self.world = 1

    # ------------------------------------------------------------------------------------------
    def _batches(self):
        rc = self.runner_config
        datarc = rc.get("datarc") or rc.get("pretrain_expert", {}).get("datarc", {})
        B = int(datarc.get("train_batch_size", 4))
        task = self.upstream_config.get("task", {"sequence_length": 750 if self.args.frame_period == 20 else 1500})
        D = self.upstream_config["melhubert"]["feat_emb_dim"]
        if getattr(self.args, "synthetic", False) or not datarc.get("sets"):
            return SyntheticBuckets(B, int(task["sequence_length"]), D, seed=self.args.seed + 97 * self.rank), B, int(task["sequence_length"]), D
        return CsvNpyBuckets(datarc, task, self.args.frame_period, B, self.rank, self.world), B, int(task["sequence_length"]), D

    def _new_step(self, B, T, D, accum, carry=None):
        """Flat buffers + fused Adam + (graph-captured) step for the model's CURRENT shapes (runner.py:312-314;
        rebuilt after head / row pruning like the reference's ``optimizer = self._get_optimizer(...)``, :348, :355)."""
        from speech_ssl_compression_b200.trainer import TrainStep

        oc = self.runner_config.get("optimizer", {})
        rc = self.runner_config.get("runner", {})
        step = TrainStep(self.expert, B, T, D, lr=float(oc.get("lr", 1e-3)), betas=tuple(oc.get("betas", (0.9, 0.999))),
                         eps=float(oc.get("eps", 1e-8)), weight_decay=float(oc.get("weight_decay", 0.0)),
                         max_norm=float(rc.get("gradient_clipping", 10.0)), accum=accum,
                         use_graph=not getattr(self.args, "no_graph", False))
        if carry is None and getattr(self.args, "init_optimizer_from_initial_weight", False):
            states = torch.load(self.args.initial_weight, map_location="cpu", weights_only=False)
            try:
                step.opt.load_state_dict(states["Optimizer"])
                print(f"[Runner] Load initilization optimizer weight from {self.args.initial_weight}")
            except Exception as e:  # (runner.py:163-170)
                raise NotImplementedError(f"Could not load the initilization weight of optimizer: {e}")
        return step

    @staticmethod
    def _fit(step, feat, label, pad, lens, D):
        """A bucket shorter / smaller than the step's static shape is padded up (padded utterances keep one valid
        frame so every row has a key to attend to; their labels are -100 and never reach the loss)."""
        if feat.shape[1] == step.T and feat.shape[0] == step.B:
            return feat, label, pad, lens
        f2 = torch.zeros(step.B, step.T, D).pin_memory()
        l2 = torch.full((step.B, step.T), -100).pin_memory()
        p2 = torch.zeros(step.B, step.T).pin_memory()
        b, t = min(feat.shape[0], step.B), min(feat.shape[1], step.T)
        f2[:b, :t], l2[:b, :t], p2[:b, :t] = feat[:b, :t], label[:b, :t], pad[:b, :t]
        lens = [min(int(x), t) for x in lens[:b]] + [1] * (step.B - b)
        p2[b:, :1] = 1
        return f2, l2, p2, lens

    def train(self):
        import random

        rc = self.runner_config.get("runner", {})
        accum = max(int(rc.get("gradient_accumulate_steps", 1)), 1)
        log_step = int(rc.get("log_step", 100))
        data, B, T, D = self._batches()
        n_epochs = int(rc.get("n_epochs", -1) or -1)
        total = int(rc.get("total_steps", 1000))
        if n_epochs > 0:  # runner.py:291-294
            total = int(n_epochs * len(data) / accum)
        total = int(getattr(self.args, "max_steps", None) or total)
        step_per_epoch = max(len(data) // accum, 1)
        mode = self.args.mode
        if self.rank == 0:
            print(f"[Runner] - Accumulated batch size: {B * accum}" + (f" x {self.world} ranks" if self.world > 1 else ""))
        self.expert.train()
        step = self._new_step(B, T, D, accum)
        if self.world > 1:  # per-rank crop / span-mask streams (model init above used the common seed)
            random.seed(self.args.seed + 7919 * self.rank)
            np.random.seed(self.args.seed + 7919 * self.rank)
        log = None
        if self.rank == 0:
            log = csv.writer(open(os.path.join(self.args.expdir, "train_log.csv"), "a", newline=""))
        bar = tqdm(total=total, dynamic_ncols=True, desc="overall", disable=self.rank != 0)
        save_every = getattr(self.tools, "save_every_x_epochs", None)
        self.loss_history = []
        pending = []  # [(optimizer step number, loss handle)]: read back one step late

        def drain(keep=0):
            while len(pending) > keep:
                gs_l, handle = pending.pop(0)
                loss = step.collect_loss(handle)
                self.loss_history.append(loss)
                if mode == "weight-pruning":  # every optimizer step (runner.py:402-406)
                    self.tools.update_smooth_loss(loss)
                    self.tools.update_target_smooth_loss(gs_l)
                if math.isnan(loss):
                    tqdm.write(f"[Runner] - loss is NaN at step {gs_l}")
                if log is not None and (gs_l % log_step == 0 or gs_l == total):
                    log.writerow([gs_l, f"{mode}/train-loss", loss, f"{mode}/train-gradient norm", self._last_norm])

        def events(gs):
            """Scheduled work in front of optimizer step gs + 1 (runner.py:326-356).  Returns the (possibly new) step."""
            nonlocal total
            st = step
            if mode in ("melhubert", "distillation"):
                if save_every and gs % max(int(save_every * step_per_epoch), 1) == 0:
                    self.tools.save_model(st.opt, gs, gs // step_per_epoch)
            elif gs in self.prune_steps:
                drain()
                if mode == "weight-pruning":
                    state = self.tools.prune_api(st.opt, gs, total)
                    if state == "not-converge":  # runner.py:338-340
                        total += self.tools.period
                        bar.total = total
                        self.prune_steps.append(max(self.prune_steps) + self.tools.period)
                    else:
                        st.graphs = None  # new mask buffers: the operand prep is re-captured (Adam state is kept)
                else:
                    self.tools.save_model(st.opt, gs)
                    self.tools.prune_api()
                    st = self._new_step(B, T, D, accum, carry=st)  # shapes changed: new flat buffers / Adam state
            return st

        self._last_norm = float("nan")
        gs, micro = 0, 0
        it = iter(data)

        def next_batch():
            nonlocal it
            try:
                return next(it)
            except StopIteration:
                it = iter(data)
                return next(it)

        step = events(0)
        step.stage_batch(*self._fit(step, *next_batch(), D))
        while gs < total:
            step.commit_staged()
            try:
                last = step.run()
            except RuntimeError as e:
                if "CUDA out of memory" not in str(e):
                    raise
                # runner.py:379-386: drop the micro-batch, free the cache, start the accumulation over
                tqdm.write(f"[Runner] - CUDA out of memory at step {gs + 1}")
                torch.cuda.empty_cache()
                step.opt.zero_grad()
                step.loss_acc.zero_()
                step._micro, micro, last = 0, 0, False
            micro += 1
            if last:
                gs += 1
                pending.append((gs, step.read_loss_async()))
                if gs % log_step == 0 or gs == total:
                    drain()
                    self._last_norm = step.opt.grad_norm(1.0 / accum)
                    if math.isnan(self._last_norm):  # the fused Adam kernel left parameters / moments untouched
                        tqdm.write(f"[Runner] - Error : grad norm is NaN at global step {gs}")
                else:
                    drain(keep=1)
                bar.update(1)
                if gs >= total:
                    break
                step = events(gs)
            step.stage_batch(*self._fit(step, *next_batch(), D))
        drain()
        bar.close()
        self.global_step = gs
        if mode == "weight-pruning":
            self.tools._save(step.opt, gs, total, "last-step.ckpt")  # + Pruning / RandomState / TotalStep (wp_utils.py:162-180)
        elif mode in ("melhubert", "distillation"):
            self.tools.save_model(step.opt, gs, gs // step_per_epoch, name="last-step.ckpt")
        else:
            self.tools.save_model(step.opt, gs)  # states_prune_<n>.ckpt, runner.py:454-457
            from speech_ssl_compression_b200 import ckpt

            states = {"Optimizer": step.opt.state_dict(), "Step": gs, "Args": self.args, "Runner": self.runner_config}
            if mode == "head-pruning":
                states["Pruned_heads"] = self.tools.pruned_heads
            ckpt.save(self.expert.add_state_to_save(states), os.path.join(self.args.expdir, "last-step.ckpt"))
        return self.loss_history
