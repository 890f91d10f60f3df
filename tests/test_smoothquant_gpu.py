"""SmoothQuant initialiser: per-channel activation scales / shifts and the closed-form weight migration
against golden vectors written by the reference's own function bodies (oracle/make_golden_smooth.py)."""
import pytest
import torch
from helpers import load_golden

pytestmark = pytest.mark.gpu


def _model(g, dev):
    from mobilequant_b200.model import HFConfig, HFForCausalLM
    m = HFForCausalLM(HFConfig(**g["cfg"], use_cache=False, use_matmul_as_module=True, l2norm_as_rmsnorm=True))
    missing, unexpected = m.load_state_dict(g["state_dict"], strict=False)
    assert not unexpected and not missing, (missing, unexpected)
    return m.float().eval().to(dev)


@pytest.mark.parametrize("tag", ["llama", "stablelm"])
def test_act_scales_and_shifts_bit_exact(cuda, tag):
    from mobilequant_b200.ptq.generate_act_scale_shift import get_act_scales, get_act_shifts
    g = load_golden(f"smooth_{tag}.pt")
    model = _model(g, cuda)
    scales = get_act_scales(model, g["samples"])
    shifts = get_act_shifts(model, g["samples"])
    assert scales.keys() == g["act_scales"].keys() and shifts.keys() == g["act_shifts"].keys()
    # the statistics are order-free reductions of the same fp32 activations; the activations themselves come from a GPU
    # forward (TF32 off) vs the reference's CPU forward, hence a tolerance instead of equality
    for k in scales:
        assert torch.allclose(scales[k], g["act_scales"][k], rtol=1e-4, atol=1e-5), k
        assert torch.allclose(shifts[k], g["act_shifts"][k], rtol=1e-4, atol=1e-5), k


@pytest.mark.parametrize("tag", ["llama", "stablelm"])
@pytest.mark.parametrize("variant", ["default", "alpha075_orig_omni"])
def test_smooth_lm_matches_reference(cuda, tag, variant):
    """Given the reference's act_scales, the smoothed weights are bit-identical (same elementwise fp32 operations)."""
    from mobilequant_b200.ptq.smoothquant import smooth_lm
    g = load_golden(f"smooth_{tag}.pt")
    model = _model(g, cuda)
    gold = g["smoothed"][variant]
    smooth_lm(model, g["act_scales"], **gold["kw"])
    sd = model.state_dict()
    changed = {k for k, v in sd.items() if k in g["state_dict"] and not torch.equal(v.cpu(), g["state_dict"][k])}
    assert changed == set(gold["changed"].keys())
    for k, ref in gold["changed"].items():
        got = sd[k].cpu()
        # pow() is the only non-IEEE-exact operation (CPU libm vs CUDA powf): a few ulp on the scale vector
        assert torch.allclose(got, ref, rtol=1e-6, atol=0), (k, (got - ref).abs().max().item())
