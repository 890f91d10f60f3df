"""Are the gradients of one training step bit-identical for different numbers of weight-pass streams?"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, "tests")
from helpers import load_golden, sim_qmodel, calib_args
from mobilequant_b200.quantization import algorithm as A
from mobilequant_b200.model.hf_model import causal_mask_4d
cuda = torch.device("cuda")
g = load_golden("model_llama_w8_e2e.pt")
T = g["samples"][0].shape[1]
res = {}
for n in (0, 1, 2, 3, 4, 2, 1):
    torch.manual_seed(0)
    m = sim_qmodel(g, cuda)
    args = calib_args(g, "/tmp")
    layers = m.model.layers
    emb = m.model.embed_tokens(g["samples"][0].to(cuda))
    mask = causal_mask_4d(1, T, torch.float32, cuda); pos = torch.arange(T, device=cuda).unsqueeze(0)
    backbone = A.LayerList(layers)
    A.disable_quant(m)
    with torch.no_grad():
        fp_t = backbone(emb, mask, pos)[0]
    A.enable_quant(args, m)
    for i, l in enumerate(layers):
        for k, v in g["let0"][i].items():
            l.register_parameter(k, torch.nn.Parameter(v.to(cuda)))
    side = [torch.cuda.Stream() for _ in range(n)] if n else None
    for rep in range(2):
        for p in m.parameters():
            p.grad = None
        for l in layers:
            A.smooth_lm_temporary(l, m.config, True, False)
        A._prefetch_weights(layers, side)
        out = backbone(emb, mask, pos)[0]
        loss = torch.nn.functional.mse_loss(fp_t, out)
        loss.backward()
        for s in side or ():
            torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
    grads = {f"{i}.{k}": p.grad.detach().clone() for i, l in enumerate(layers) for k, p in l.named_parameters() if p.grad is not None}
    key = n
    if 0 in res:
        bad = [k for k in grads if not torch.equal(grads[k], res[0][k])]
        print(f"streams={n}: loss {loss.item():.9f}  params differing from the single-stream run: {len(bad)} of {len(grads)}", bad[:6])
        for k in bad[:3]:
            d = (grads[k] - res[0][k]).abs().max().item(); print("    ", k, "max abs diff", d, "ref max", res[0][k].abs().max().item())
    else:
        res[0] = grads
        print(f"streams=0 (reference): loss {loss.item():.9f}")
