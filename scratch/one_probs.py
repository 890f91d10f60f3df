"""One forward + backward of the fused calibration attention core at TinyLlama shapes (for ncu)."""
import math, sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mobilequant_b200.quantization.functional import AttnProbsFn, StaticFakeQuantFn
from mobilequant_b200.quantization.qmodule import compute_scale_offset_from_min_max
dev = torch.device("cuda")
T = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
torch.manual_seed(0)
S = (torch.randn(1, 32, T, T, device=dev) * 6).requires_grad_(True)
W = torch.randn(1, 32, T, T, device=dev)
def qp(mn, mx, bits):
    s, o, _, _, lo, hi = compute_scale_offset_from_min_max(mn, mx, bits, False)
    return [torch.nn.Parameter(s.to(dev)), torch.nn.Parameter(o.to(dev)), lo, hi]
q1, q2 = qp(-20.0, 18.0, 16), qp(0.0, 1.0, 16)
for it in range(3):
    out = AttnProbsFn.apply(S, 0.125, *q1, *q2)
    out.backward(W)
x = torch.randn(1024, 5632, device=dev, requires_grad=True)
q3 = qp(-4.0, 4.0, 8)
for it in range(2):
    y = StaticFakeQuantFn.apply(x, *q3)
    y.backward(W[0, 0, :, :].repeat(1, 6)[:, :5632].contiguous())
torch.cuda.synchronize()
from mobilequant_b200 import kernels as K
Sc = S.detach(); f = lambda q: (q[0].detach(), q[1].detach(), q[2], q[3])
P, stats = K.attn_probs_fwd(Sc, T, True, 0.125, f(q1), f(q2))
xd, yd = x.detach(), W[0, 0, :, :].repeat(1, 6)[:, :5632].contiguous()
def timed(fn, n=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n
print("attn_probs fwd %.1f us  bwd %.1f us | fq_fwd %.1f us  fq_bwd %.1f us (1024 x 5632)" % (
    timed(lambda: K.attn_probs_fwd(Sc, T, True, 0.125, f(q1), f(q2))),
    timed(lambda: K.attn_probs_bwd(Sc, stats, W, T, True, 0.125, f(q1), f(q2))),
    timed(lambda: K.fq_fwd(xd, q3[0].detach(), q3[1].detach(), q3[2], q3[3])),
    timed(lambda: K.fq_bwd(xd, yd, q3[0].detach(), q3[1].detach(), q3[2], q3[3]))))
