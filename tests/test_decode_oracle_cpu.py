"""CPU check of the decode oracle: the cache-based one-token step (oracle/int_ref.py:IntModel.decode, a restatement of
SimModel.forward with k_cache / v_cache, mobilellm/model/sim_model.py:105-132,271-330) reproduces, bit for bit, row `pos` of the
full-sequence integer forward -- the equivalence the GPU decode tests rely on."""
import numpy as np
import pytest
import torch
from oracle import int_ref as ir
from oracle import model_ref as mr
from helpers import load_golden


@pytest.mark.parametrize("tag", ["llama_w8_e2e", "stablelm_w8_omni", "gemma_w8_e2e"])
def test_cache_decode_equals_full_forward_rows(tag):
    g = load_golden(f"model_{tag}.pt")
    im = ir.IntModel(g["state_dict"], g["cfg"], mr.recipe_from_qcfg_json(g["qcfg"]), g["act_dict"])
    ids = torch.cat(g["samples"][:2], dim=0).numpy()[:, :24]
    B, T = ids.shape
    T0 = T - 5
    cos, sin = ir.rope_tables(T, im.rot, g["cfg"].get("rope_theta", 10000.0))
    h_full, _ = im.backbone(im.embed(ids), B, T, cos, sin)
    h_full = h_full.reshape(B, T, -1)
    cache = {}
    h_pre, _ = im.backbone(im.embed(ids[:, :T0]), B, T0, cos[:T0], sin[:T0], cache=cache)
    assert np.array_equal(h_pre.reshape(B, T0, -1), h_full[:, :T0])          # causal: the prefix does not see the future
    for t in range(T0, T):
        h = im.decode(im.embed(ids[:, t]), t, cache, cos, sin)
        assert np.array_equal(h, h_full[:, t]), f"position {t}"
    assert cache[0][0].shape[2] == T
