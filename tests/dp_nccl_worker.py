"""Worker of tests/test_gpu_multi.py, one process per GPU under torchrun (NCCL).  Checks SURVEY 4-5 on real NCCL:
(1) gradients after the bucket all-reduces == single-GPU gradients on the concatenated batch (the global-mean
cross-entropy of nn.DataParallel's gather-then-mean, reference upstream/melhubert/pretrain_expert.py:28-30), also with
gradient accumulation (all-reduce only behind the final micro-batch); (2) parameters stay identical on all ranks
through graph-captured optimizer steps; (3) prune events select the same heads / rows / weight masks on every rank."""
import hashlib
import os
import sys
import tempfile
from argparse import Namespace

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import melhubert_oracle as O  # noqa: E402  (test infrastructure: synthetic weights / batches only)
from speech_ssl_compression_b200 import kernels as K  # noqa: E402
from speech_ssl_compression_b200.parallel import init_distributed  # noqa: E402
from speech_ssl_compression_b200.trainer import TrainStep  # noqa: E402
from speech_ssl_compression_b200.upstream.melhubert.pretrain_expert import MelHuBERTPretrainer  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def fwd_bwd(ts, last=True):
    """TrainStep._body without the optimizer: gradients stay in the flat buffer."""
    K.counter_add(ts.rng_counter, 1)
    if ts.dp is not None:
        ts.dp.sync = last
    ts._patch_mask(True)
    try:
        loss, _ = ts.expert((ts.feat, ts.label, ts.pad, None))
    finally:
        ts._patch_mask(False)
    loss.backward()
    if ts.dp is not None and last:
        ts.dp.finish()
    torch.cuda.synchronize()
    return float(loss.detach())


def same_everywhere(t, what):
    h = hashlib.sha256(t.detach().cpu().contiguous().numpy().tobytes()).hexdigest()
    got = [None] * dist.get_world_size()
    dist.all_gather_object(got, h)
    assert len(set(got)) == 1, f"{what} differs between ranks"


def main():
    rank, world = init_distributed("nccl")
    dev = f"cuda:{torch.cuda.current_device()}"
    cfg = dict(feat_emb_dim=80, encoder_layers=2, encoder_embed_dim=768, encoder_ffn_embed_dim=3072,
               encoder_attention_heads=12, num_cluster=512, mask_prob=0.7, mask_length=5, skip_masked=False,
               skip_nomask=True, dropout=0.0, attention_dropout=0.0, activation_dropout=0.0)
    Bl, T, D = 2, 256, 80
    Bg = Bl * world
    lens = [256, 200, 256, 131, 90, 256, 177, 256][:Bg]
    sd = O.synth_state_dict(cfg, seed=7)

    def expert(multi):
        ex = MelHuBERTPretrainer({"melhubert": dict(cfg)}, None, dev, multi)
        ex.model.load_state_dict(sd)
        return ex.to(dev).train()

    batches, masks = [], []
    for s in (8, 9):
        batches.append(O.synth_batch(Bg, T, D, lens, seed=s))
        np.random.seed(100 + s)
        masks.append(torch.from_numpy(O.span_mask(Bg, T, lens, 0.7, 5)))
    sl = slice(rank * Bl, (rank + 1) * Bl)

    def load(ts, i, part):
        f, l, p = (x[part].contiguous().pin_memory() for x in batches[i])
        ts.load_batch(f, l, p, lens[part], masking=False)
        ts.mask.copy_(masks[i][part])

    single = TrainStep(expert(False), Bg, T, D, use_graph=False)
    multi = TrainStep(expert(True), Bl, T, D, use_graph=False)
    assert multi.dp.enabled and multi.dp.world_size == world
    assert (multi.dp.peer is not None) == (os.environ.get("MH_DP_TRANSPORT", "peer") != "nccl")
    assert multi.dp.peer is None or multi.dp.peer.use_sm == (os.environ.get("MH_DP_TRANSPORT") == "peer-sm")
    # (1) one micro-batch, then two accumulated ones
    for n_micro in (1, 2):
        single.flat.flat_grad.zero_()
        multi.flat.flat_grad.zero_()
        for i in range(n_micro):
            load(single, i, slice(0, Bg))
            load(multi, i, sl)
            l1 = fwd_bwd(single)
            l2 = fwd_bwd(multi, last=(i == n_micro - 1))
            assert abs(l1 - l2) < 2e-3 * abs(l1), (l1, l2)  # global-mean loss on every rank
        g1, g2 = single.flat.flat_grad, multi.flat.flat_grad
        assert rel(g2, g1) < 2e-2, rel(g2, g1)
        for span in multi.dp._layer_spans + multi.dp._rest_spans:
            assert rel(g2[span[0]:span[1]], g1[span[0]:span[1]]) < 3e-2, span
        same_everywhere(g2, f"all-reduced gradient ({n_micro} micro-batches)")
    single.flat.flat_grad.zero_()
    multi.flat.flat_grad.zero_()
    del single
    # (2) graph-captured data-parallel steps with accumulation: replicas stay bit-identical
    ts = TrainStep(multi.expert, Bl, T, D, lr=1e-3, use_graph=True, accum=2)
    for i in range(6):
        load(ts, i % 2, sl)
        ts.run()
    torch.cuda.synchronize()
    assert int(ts.opt.step_count) == 3
    same_everywhere(ts.flat.flat_param, "parameters after 3 data-parallel optimizer steps")
    # (3) prune events
    from speech_ssl_compression_b200.head_pruning.hp_utils import HeadPruningTools
    from speech_ssl_compression_b200.pytorch_code import prune
    from speech_ssl_compression_b200.row_pruning.rp_utils import RowPruningTools
    from speech_ssl_compression_b200.weight_pruning.wp_utils import get_params_to_prune

    ex = ts.expert
    args = Namespace(expdir=tempfile.mkdtemp(), device=dev)
    hp = HeadPruningTools(args, {"prune": {"metric": "l1", "target": "by_layer", "total_steps": 2}}, {"melhubert": cfg}, ex)
    hp.prune_api()
    rec = [None] * world
    dist.all_gather_object(rec, repr(hp.pruned_heads))
    assert len(set(rec)) == 1 and [l.self_attn.num_heads for l in ex.model.encoder.layers] == [11, 11]
    rp = RowPruningTools(args, {"prune": {"num_rows_each_step": 128, "total_steps": 2}}, {"melhubert": cfg}, ex)
    rp.prune_api()
    for layer in ex.model.encoder.layers:
        assert layer.fc1.weight.shape[0] == 3072 - 128
        same_everywhere(layer.fc1.bias, "row selection")
    params, _ = get_params_to_prune(ex.model)
    prune.global_unstructured(params, pruning_method=prune.L1Unstructured, amount=0.5)
    for name, buf in ex.model.named_buffers():
        if name.endswith("_mask"):
            same_everywhere(buf, name)
    if rank == 0:
        assert not [f for f in os.listdir(args.expdir) if f.startswith("states_")]  # (save_model was not called)
    dist.barrier()
    print(f"[dp_nccl_worker] rank {rank}/{world}: OK", flush=True)
    os._exit(0)  # captured graphs still reference the communicator: let process exit release it


if __name__ == "__main__":
    main()
