"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.pt by running the UNMODIFIED reference (oracle/ref_shim.py)
in the build container, and checks oracle/fakequant_ref.py against it on the way (bit-exact where stated).

    python oracle/make_golden.py            # writes tests/golden/*.pt, prints the oracle-vs-reference report

Seeds follow ptq/mobilequant.py:87-90 (1337).  Fixtures are kept small (a few hundred KB) so they can be committed.
"""
import os, sys, io, json, copy, types, logging
import torch, torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_shim import load_reference, ref_config, ref_update_quant_cfg
from oracle import fakequant_ref as fr

GOLD = os.path.join(ROOT, "tests", "golden")
os.makedirs(GOLD, exist_ok=True)
hm, qm, alg = load_reference()


def gen(seed):
    g = torch.Generator(); g.manual_seed(seed); return g


def golden_quantizer():
    """Quantizer.forward (qm:251-295) static / dynamic / LWC, with autograd grads."""
    cases = []
    g = gen(1337)
    for bits, sym, per_ch, lwc, shape in [
        (8, False, False, False, (4, 37, 64)),     # per-tensor asym activation (default A8)
        (16, False, False, False, (3, 50, 32)),    # 16-bit activation
        (8, True, False, False, (130, 96)),        # symmetric
        (8, False, False, True, (48, 96)),         # W8 per-tensor + LWC
        (8, False, True, True, (48, 96)),          # W8 per-channel + LWC (w2)
        (4, True, True, True, (40, 128)),          # W4 per-channel symmetric + LWC
        (4, False, True, True, (40, 128)),         # W4 per-channel asym + LWC
        (16, False, False, True, (1, 128)),        # norm weight 16-bit per-tensor + LWC
    ]:
        x = torch.randn(shape, generator=g) * 0.7 + 0.1
        if x.dim() == 2 and lwc:
            x = x * 0.02
        qcfg = qm.QuantConfig(bitwidth=bits, is_symmetric=sym, is_per_channel=per_ch)
        q = qm.Quantizer(qcfg)
        xin = x.clone().requires_grad_(True)
        case = dict(bits=bits, sym=sym, per_channel=per_ch, lwc=lwc, x=x)
        if lwc:
            q.enable_lwc(x)
            with torch.no_grad():
                q.upbound_factor.add_(torch.randn(q.upbound_factor.shape, generator=g) * 0.5)
                q.lowbound_factor.add_(torch.randn(q.lowbound_factor.shape, generator=g) * 0.5)
            y = q(xin)
            case.update(up=q.upbound_factor.detach().clone(), low=q.lowbound_factor.detach().clone(),
                        scale=q.scale.detach().clone(), offset=q.offset.detach().clone())
        else:
            lo, hi = x.min().item() * 0.9, x.max().item() * 0.8      # force some clamping
            q.set_scale_offset_from_minmax(lo, hi, "parameter")
            y = q(xin)
            case.update(minmax=[lo, hi], scale=q.scale.detach().clone(), offset=q.offset.detach().clone())
        gy = torch.randn(y.shape, generator=g)
        y.backward(gy)
        case.update(y=y.detach().clone(), gy=gy, gx=xin.grad.clone(), qmin=q.qmin, qmax=q.qmax)
        if lwc:
            case.update(g_up=q.upbound_factor.grad.clone(), g_low=q.lowbound_factor.grad.clone())
        else:
            case.update(g_scale=q.scale.grad.clone(), g_offset=q.offset.grad.clone())
        # --- oracle check (bit-exact forward, grads to fp32 round-off)
        if lwc:
            su, sl = torch.sigmoid(case["up"]), torch.sigmoid(case["low"])
            yo, so, oo, qmin, qmax = fr.dynamic_fake_quant(x, bits, sym, per_ch, su, sl, return_params=True)
        else:
            so, oo, qmin, qmax = fr.scale_offset_from_minmax(lo, hi, bits, sym)
            yo = fr.fake_quant(x, so, oo, qmin, qmax)
        assert torch.equal(yo, case["y"]), ("oracle != reference", bits, sym, per_ch, lwc)
        assert torch.equal(so.reshape(-1), case["scale"].reshape(-1)) and qmin == q.qmin and qmax == q.qmax
        cases.append(case)
    torch.save(cases, os.path.join(GOLD, "quantizer.pt"))
    print(f"quantizer.pt: {len(cases)} cases, oracle forward bit-exact vs reference Quantizer")


if __name__ == "__main__":
    golden_quantizer()
