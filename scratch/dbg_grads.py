import sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from helpers import *
from mobilequant_b200.quantization import algorithm as A
from mobilequant_b200.model.hf_model import causal_mask_4d
tag = sys.argv[1] if len(sys.argv) > 1 else "llama_w8_e2e"
g = load_golden(f"model_{tag}.pt"); cuda = torch.device("cuda:0")
m = sim_qmodel(g, cuda); args = calib_args(g, "/tmp")
layers = m.model.layers; T = g["samples"][0].shape[1]
emb = m.model.embed_tokens(g["samples"][0].to(cuda))
if m.config.normalize_embed: emb = emb * (m.config.hidden_size ** 0.5)
mask = causal_mask_4d(1, T, torch.float32, cuda); pos = torch.arange(T, device=cuda).unsqueeze(0)
backbone = A.LayerList(layers)
A.disable_quant(m)
with torch.no_grad(): fp_t = backbone(emb, mask, pos)[0]
A.enable_quant(args, m)
for i, l in enumerate(layers):
    for k, v in g["let0"][i].items(): l.register_parameter(k, torch.nn.Parameter(v.to(cuda)))
    A.smooth_lm_temporary(l, m.config, True, False)
out = backbone(emb, mask, pos)[0]
loss = torch.nn.functional.mse_loss(fp_t, out); loss.backward()
print("loss", loss.item(), g["loss0"])
for i, l in enumerate(layers):
    got = {k: p.grad for k, p in l.named_parameters() if p.grad is not None and "smooth_shift" not in k}
    for k, ref in g["grads0"][i].items():
        gg = got[k].cpu()
        print(i, k, "ref|max| %.3e got|max| %.3e maxdiff %.3e" % (ref.abs().max(), gg.abs().max(), (gg-ref).abs().max()),
              ("ref %.4e got %.4e" % (ref.item(), gg.item())) if ref.numel()==1 else "")
