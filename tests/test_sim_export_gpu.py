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


@pytest.mark.parametrize("tag", ["llama_w8_e2e", "llama_w4_omni"])
def test_export_quantized_weights_and_encodings(cuda, tag, tmp_path):
    """N3: packed integer weight export + self-contained `.encodings` from the calibrated artefacts.  The exported codes
    de-quantise bit-exactly to the fake-quantised weights of the calibration path (the reference's Quantizer.forward on the
    fused weight), 4-bit codes are packed two per byte, and the activation encodings carry the act_dict ranges."""
    import json, os
    import numpy as np
    from helpers import load_golden, product_model
    from mobilequant_b200 import kernels as K
    from mobilequant_b200.device import encodings as E
    g = load_golden(f"model_{tag}.pt")
    model = product_model(g, cuda)
    enc, weights = E.export_all(model, g["act_dict"], g["qcfg"], str(tmp_path), name="tiny")
    assert os.path.exists(tmp_path / "tiny.encodings") and os.path.exists(tmp_path / "tiny_kv_cache.encodings") and os.path.exists(tmp_path / "tiny_qweights.pth")
    w_cfg = g["w_cfg"]
    n_lin = 0
    for name, mod in model.named_modules():
        if not isinstance(mod, torch.nn.Linear) or name not in g["qcfg"]:
            continue
        n_lin += 1
        e = weights[name]
        c = g["qcfg"][name]["weight"]
        bits, sym, pc = int(c["bitwidth"]), c["is_symmetric"] == "True", c["is_per_channel"] == "True"
        ref = K.wprep_fwd(mod.weight.detach().float().contiguous(), bits, sym, pc, want_fq=True)["w_fq"].cpu()
        codes = e["codes"]
        if e["packed"]:
            assert codes.dtype == torch.uint8 and codes.shape == (mod.weight.shape[0], mod.weight.shape[1] // 2)
            lo, hi = (codes & 0xF).to(torch.int16), (codes >> 4).to(torch.int16)
            if sym:                                                   # 4-bit two's complement
                lo, hi = torch.where(lo > 7, lo - 16, lo), torch.where(hi > 7, hi - 16, hi)
            codes = torch.stack([lo, hi], dim=-1).reshape(mod.weight.shape)
        deq = (codes.float() - e["offset"].view(-1, 1)) * e["scale"].view(-1, 1)
        assert torch.equal(deq, ref), name
        rows = enc["param_encodings"][name + ".weight"]
        assert len(rows) == (mod.weight.shape[0] if pc else 1) and rows[0]["bitwidth"] == bits
    assert n_lin == 7 * g["cfg"]["num_hidden_layers"]
    assert (w_cfg["bits"] == 4) == any(e["packed"] for e in weights.values())
    a = enc["activation_encodings"]
    mn, mx = g["act_dict"]["model.layers.1.mlp.w2"]["output"]
    e = a["module_add_9"]["input"]["1"]                                # second residual add of block 1 <- mlp.w2.output
    assert e["bitwidth"] == 16 and e["min"] == mn and e["max"] == mx and e["scale"] == (mx - mn) / 65535
    kv = json.load(open(tmp_path / "tiny_kv_cache.encodings"))
    assert kv["k_cache"]["min"] == min(g["act_dict"][f"model.layers.{i}.self_attn.qk_bmm"]["input2"][0] for i in range(2))
