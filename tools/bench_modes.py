"""Builds the expert (model + criterion) for each bench.py mode from the reference's config
shapes, with random-init weights (no checkpoints are available offline)."""
import os
import tempfile

import torch

from bench import fwd_flops_per_frame
from bench import model_cfg as _bench_cfg


def build_expert(mode, multi_gpu, T, device="cuda", frame=20, heads=6):
    from speech_ssl_compression_b200.upstream.melhubert.pretrain_expert import MelHuBERTPretrainer

    rho = 0.48  # fraction of frames that go through final_proj (masked & valid), SURVEY §8(d)
    D = 80 if frame == 20 else 40

    def model_cfg(**kw):  # every mode on the requested frame period (10 ms: D_in 40, mask spans of 10)
        return _bench_cfg(frame, **kw)

    if mode == "pretrain":
        cfg = model_cfg()
        ex = MelHuBERTPretrainer({"melhubert": cfg}, None, device, multi_gpu).to(device)
        return ex, 3 * fwd_flops_per_frame(T, D, 12, 12, 3072, rho) / 1e6
    if mode == "distillation":
        from speech_ssl_compression_b200.distillation.pretrain_expert import MelHuBERTDistiller
        from speech_ssl_compression_b200.model import MelHuBERTConfig, MelHuBERTModel

        tcfg = model_cfg(skip_masked=False, skip_nomask=False)
        scfg = model_cfg(layers=2, skip_masked=False, skip_nomask=False, initial_from_teacher=False)
        torch.manual_seed(1337)
        teacher = MelHuBERTModel(MelHuBERTConfig(tcfg))
        ck = os.path.join(tempfile.mkdtemp(), f"teacher-{os.getpid()}.ckpt")
        torch.save({"model": teacher.state_dict()}, ck)
        del teacher
        ucfg = {"melhubert": scfg, "teacher": tcfg, "loss_param": {"T": 1, "alpha": 1, "type": "nomasked"}}
        ex = MelHuBERTDistiller(ucfg, ck, device, multi_gpu).to(device)
        os.remove(ck)
        fl = fwd_flops_per_frame(T, D, 12, 12, 3072, 1.0) + 3 * fwd_flops_per_frame(T, D, 2, 12, 3072, 1.0)
        return ex, fl / 1e6
    if mode == "weight-pruning":
        from speech_ssl_compression_b200.pytorch_code import prune
        from speech_ssl_compression_b200.weight_pruning.wp_utils import get_params_to_prune

        cfg = model_cfg()
        ex = MelHuBERTPretrainer({"melhubert": cfg}, None, device, multi_gpu).to(device)
        params, _ = get_params_to_prune(ex.model)
        prune.global_unstructured(params, pruning_method=prune.L1Unstructured, amount=0.5)
        return ex, 3 * fwd_flops_per_frame(T, D, 12, 12, 3072, rho) / 1e6
    if mode in ("row-pruning", "row+weight"):
        # cfg4: 12 row-prune steps of 128 rows (f = 1536), then -- "row+weight" -- global_unstructured(L1, 0.5) on the
        # shrunken model (row-prune first, weight-prune the result)
        cfg = model_cfg(ffn=1536)
        ex = MelHuBERTPretrainer({"melhubert": cfg}, None, device, multi_gpu).to(device)
        if mode == "row+weight":
            from speech_ssl_compression_b200.pytorch_code import prune
            from speech_ssl_compression_b200.weight_pruning.wp_utils import get_params_to_prune

            params, _ = get_params_to_prune(ex.model)
            prune.global_unstructured(params, pruning_method=prune.L1Unstructured, amount=0.5)
        return ex, 3 * fwd_flops_per_frame(T, D, 12, 12, 1536, rho) / 1e6
    if mode == "head-pruning":
        from speech_ssl_compression_b200.surgery import drop_heads

        cfg = model_cfg()
        ex = MelHuBERTPretrainer({"melhubert": cfg}, None, device, multi_gpu).to(device)
        drop = [0, 3, 5, 7, 9, 11, 1, 4, 8, 10, 2][: 12 - heads]
        for layer in ex.model.encoder.layers:
            if drop:
                drop_heads(layer.self_attn, sorted(drop))
        return ex, 3 * fwd_flops_per_frame(T, D, 12, heads, 3072, rho) / 1e6
    raise ValueError(mode)
