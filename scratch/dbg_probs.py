import math, torch, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from test_quant_kernels_gpu import _probs_chain
from mobilequant_b200.quantization.functional import AttnProbsFn, StaticFakeQuantFn
from mobilequant_b200.quantization.qmodule import compute_scale_offset_from_min_max
cuda = torch.device('cuda')
for (T, hd, bits1) in [(256, 80, 8), (256, 64, 8), (256, 80, 16), (64, 64, 8)]:
    torch.manual_seed(T + hd)
    B, nh = 2, 3
    S0 = torch.randn(B, nh, T, T, device=cuda) * 6.0
    W = torch.randn(B, nh, T, T, device=cuda)
    def qp(mn, mx, bits):
        s, o, _, _, lo, hi = compute_scale_offset_from_min_max(mn, mx, bits, False)
        return [torch.nn.Parameter(s.to(cuda)), torch.nn.Parameter(o.to(cuda)), lo, hi]
    res = []
    for fused in (False, True):
        S = S0.clone().requires_grad_(True)
        q1, q2 = qp(-20.0, 18.0, bits1), qp(0.0, 1.0, 16)
        if fused:
            mul = (torch.ones(()) / torch.tensor(math.sqrt(hd))).item()
            out = AttnProbsFn.apply(S, mul, *q1, *q2)
        else:
            sq = StaticFakeQuantFn.apply(S, *q1); sq.retain_grad()
            from mobilequant_b200.model.hf_model import causal_mask_4d
            a = sq / math.sqrt(hd) + causal_mask_4d(B, T, torch.float32, cuda)
            p = torch.softmax(a, -1, dtype=torch.float32)
            out = StaticFakeQuantFn.apply(p, *q2)
        (out * W).sum().backward()
        res.append((q1[0].grad.item(), q1[1].grad.item(), q2[0].grad.item(), q2[1].grad.item()))
        if not fused:
            g = sq.grad.double(); s = q1[0].detach().double(); o = q1[1].detach().double()
            u = S0.double() / s.float().double()
            u32 = (S0 / q1[0].detach())
            t3 = torch.round(u32) + q1[1].detach()
            m = (t3 >= 0) & (t3 <= q1[3])
            t5 = (t3.clamp(0, q1[3]) - q1[1].detach()).double()
            gs = (g * t5 - (g * s * m) * (u32.double() / s)).sum().item()
            go = ((g * s * m) - g * s).sum().item()
            print("double-precision check of fq1 grads from the chain's dsq:", gs, go, "clipped frac", 1 - m.float().mean().item(), "n nonzero g", (g != 0).sum().item())
    print(T, hd, bits1, "chain", res[0], "fused", res[1])
