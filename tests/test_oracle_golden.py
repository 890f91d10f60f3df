"""The CPU oracle (oracle/fakequant_ref.py) against the golden vectors produced by the UNMODIFIED reference
(oracle/make_golden.py, run in the build container)."""
import os
import torch
import pytest
from oracle import fakequant_ref as fr


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


def test_quantizer_forward_bit_exact_and_grads(golden_dir):
    for c in _load(golden_dir, "quantizer.pt"):
        x = c["x"].clone().requires_grad_(True)
        if c["lwc"]:
            up = c["up"].clone().requires_grad_(True); low = c["low"].clone().requires_grad_(True)
            y, s, o, qmin, qmax = fr.dynamic_fake_quant(x, c["bits"], c["sym"], c["per_channel"], torch.sigmoid(up),
                                                        torch.sigmoid(low), return_params=True)
        else:
            s0, o0, qmin, qmax = fr.scale_offset_from_minmax(c["minmax"][0], c["minmax"][1], c["bits"], c["sym"])
            s = s0.clone().requires_grad_(True); o = o0.clone().requires_grad_(True)
            y = fr.fake_quant(x, s, o, qmin, qmax)
        assert torch.equal(y.detach(), c["y"])                       # bit-exact
        assert (qmin, qmax) == (c["qmin"], c["qmax"])
        y.backward(c["gy"])
        assert torch.equal(x.grad, c["gx"])
        if c["lwc"]:
            assert torch.equal(up.grad, c["g_up"]) and torch.equal(low.grad, c["g_low"])
        else:
            assert torch.equal(s.grad, c["g_scale"]) and torch.equal(o.grad, c["g_offset"])


def test_scale_offset_edge_cases():
    # scale clamp (qm:58) and symmetric ranges (qm:45-49)
    s, o, qmin, qmax = fr.scale_offset_from_minmax(0.0, 0.0, 8, False)
    assert s.item() == pytest.approx(1e-5) and o.item() == 0 and (qmin, qmax) == (0, 255)
    s, o, qmin, qmax = fr.scale_offset_from_minmax(-1.0, 0.5, 4, True)
    assert (qmin, qmax) == (-8, 7) and s.item() == pytest.approx(1.0 / 7) and o.item() == 0
    s, o, _, _ = fr.scale_offset_from_minmax(-0.3, 0.9, 8, False)
    mn, mx = fr.minmax_from_scale_offset(s, o, 8, False)
    assert abs(mn.item() + 0.3) < s.item() and abs(mx.item() - 0.9) < s.item()
