"""GPU diagnostic: full MelHuBERT model (CUDA kernels) vs the CPU oracle on the same inputs."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import melhubert_oracle as O
from speech_ssl_compression_b200.model import MelHuBERTModel, MelHuBERTConfig

LENS = [750, 712, 655, 601]
dev = "cuda"


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


layers = int(sys.argv[1]) if len(sys.argv) > 1 else 12
cfg = dict(feat_emb_dim=80, encoder_layers=layers, mask_prob=0.7, mask_length=5, dropout=0.0, attention_dropout=0.0,
           activation_dropout=0.0)
sd = O.synth_state_dict(cfg, seed=7)
feat, label, pad = O.synth_batch(4, 750, 80, LENS)
model = MelHuBERTModel(MelHuBERTConfig(cfg))
model.load_state_dict(sd)
model.to(dev)

# ---- eval forward
model.eval()
t0 = time.time()
with torch.no_grad():
    out = model(feat.to(dev), pad.to(dev), get_hidden=True, no_pred=True)
torch.cuda.synchronize()
print("eval forward ok in %.2fs" % (time.time() - t0), flush=True)
with torch.no_grad():
    ref = O.model_forward(sd, cfg, feat, pad, no_pred=True)
print("pre_feat rel", rel(out[6], ref["pre_feat"]))
for i, (a, b) in enumerate(zip(out[5], ref["layer_hiddens"])):
    print(f"layer {i:2d} hidden rel-L2 {rel(a, b):.3e}", flush=True)
print("hidden rel", rel(out[0], ref["hidden"]))

# ---- train forward/backward (dropout 0)
model.train()
np.random.seed(1337)
o = model(feat.to(dev), pad.to(dev), label.to(dev), mask=True, valid_lens=LENS)
hidden, logit_m, _, label_m, _, _, _, mask_idx = o
from speech_ssl_compression_b200 import ops
loss = ops.cross_entropy(logit_m, label_m)
loss.backward()
torch.cuda.synchronize()
np.random.seed(1337)
mask = torch.from_numpy(O.span_mask(4, 750, LENS, 0.7, 5))
print("mask equal:", torch.equal(mask, mask_idx.cpu()))
sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
r = O.model_forward(sdg, cfg, feat, pad, label, mask_indices=mask)
print("label_m equal:", torch.equal(r["label_m"], label_m.cpu()), " N_m", label_m.numel())
print("logit_m rel", rel(logit_m, r["logit_m"]))
rl = O.ce_mean(r["logit_m"], r["label_m"])
print("loss", float(loss), "oracle", float(rl))
rl.backward()
worst = []
for n, p in model.named_parameters():
    if p.grad is None:
        print("NO GRAD", n); continue
    worst.append((rel(p.grad, sdg[n].grad), n))
worst.sort(reverse=True)
for e, n in worst[:12]:
    print(f"grad rel {e:.3e}  {n}")
print("median grad rel", sorted(w[0] for w in worst)[len(worst) // 2])
