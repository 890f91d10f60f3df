"""Sim-layout export against the reference's own folding statements (oracle/make_golden_sim.py)."""
import pytest
import torch
from helpers import load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag", ["llama_w8_slinear", "llama_w4", "gemma_w8_slinear"])
def test_convert_state_dict_matches_reference(cuda, tag):
    from mobilequant_b200.device.convert_sim import convert_state_dict
    g = load_golden("sim_export.pt")
    case = g[tag]
    hf = {k: v.to(cuda) for k, v in g["hf_state"].items()}
    out = convert_state_dict(hf, n_embd=64, head_dim=16, n_layer=2, impl_sym_pch_as_slinear=case["impl_sym_pch_as_slinear"],
                             is_gemma="gemma" in case["model_name"], device=cuda)
    ref = case["out_states"]
    assert set(out.keys()) == set(ref.keys())
    for k, v in ref.items():
        assert torch.equal(out[k], v), k            # max / divide / multiply by a constant: same fp32 operations, bit-identical
