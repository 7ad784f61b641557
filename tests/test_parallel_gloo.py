"""Data-parallel host logic on CPU with the gloo backend, world_size = 2 (the N > 1 path of SURVEY §8e):
flat gradient buffer, per-layer bucket all-reduce fired from backward hooks, remainder bucket in finish(),
and the 2-scalar all-reduce that makes the cross-entropy a mean over the GLOBAL masked-frame set
(nn.DataParallel's gather-then-mean semantics, pretrain_expert.py:28-30)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.manual_seed(100 + rank)  # different init per rank: the wrapper must broadcast rank 0's weights
    from speech_ssl_compression_b200.model import MelHuBERTConfig, MelHuBERTModel
    from speech_ssl_compression_b200.parallel import DataParallelB200, FlatBuffers, init_distributed

    r, w = init_distributed("gloo")
    assert (r, w) == (rank, world)
    model = MelHuBERTModel(MelHuBERTConfig(dict(feat_emb_dim=80, encoder_layers=2, encoder_ffn_embed_dim=256)))
    dp = DataParallelB200(model, overlap=True)
    flat = FlatBuffers([p for p in model.parameters() if p.requires_grad])
    dp.attach(flat)
    # every parameter is a view into the flat buffer, gradients too
    p0 = next(model.parameters())
    assert p0.data_ptr() == flat.flat_param.data_ptr() and p0.grad.data_ptr() == flat.flat_grad.data_ptr()
    # broadcast happened: checksum identical on both ranks
    chk = flat.flat_param.double().sum().reshape(1)
    both = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(both, chk)
    assert torch.equal(both[0], both[1])
    # "backward": every rank fills its gradients with rank-dependent values, layers fire their hooks
    g = torch.Generator().manual_seed(7 + rank)
    local = torch.randn(flat.total, generator=g)
    flat.flat_grad.copy_(local)
    for layer in reversed(model.encoder.layers):
        layer._mh_grad_ready_hook(layer)
    dp.finish()
    g0 = torch.randn(flat.total, generator=torch.Generator().manual_seed(7))
    g1 = torch.randn(flat.total, generator=torch.Generator().manual_seed(8))
    assert torch.allclose(flat.flat_grad, g0 + g1, atol=1e-6)               # every span reduced exactly once
    # loss normaliser: (sum of row losses, count) -> global mean
    acc = torch.tensor([10.0 * (rank + 1), 4.0 + rank])
    dp.all_reduce_sum(acc)
    assert acc.tolist() == [30.0, 9.0]
    # layer spans tile the buffer with the remainder, no overlap
    spans = sorted(dp._layer_spans + dp._rest_spans)
    assert spans[0][0] == 0 and spans[-1][1] == flat.total
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    out.put((rank, float(flat.flat_grad.sum())))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_buffer_bucket_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0, p.exitcode
    res = dict(q.get() for _ in range(2))
    assert abs(res[0] - res[1]) < 1e-3
