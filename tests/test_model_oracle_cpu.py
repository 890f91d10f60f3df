"""CPU: the functional oracle (oracle/model_ref.py) reproduces the goldens written by the unmodified reference, and
the product's host logic (module rewriting, recipes, JSON artefacts, LR schedule) matches them."""
import json, math, os
import pytest
import torch
from oracle import model_ref as mr
from helpers import load_golden, MODEL_GOLDENS, product_model


@pytest.mark.parametrize("tag", MODEL_GOLDENS)
def test_oracle_forward_and_ranges(tag):
    g = load_golden(f"model_{tag}.pt")
    act = mr.act_range(g["state_dict"], g["cfg"], g["samples"])
    assert act == g["act_dict"]                                   # bit-exact min/max of every hooked tensor
    recipe = mr.recipe_from_qcfg_json(g["qcfg"])
    qs = mr.QState(recipe, g["act_dict"])
    with torch.no_grad():
        logits, hid = mr.model_forward(g["state_dict"], g["cfg"], g["samples"][0], qs, quant=True)
    assert torch.equal(logits, g["logits_fq"])


def test_oracle_calibration_llama_e2e():
    g = load_golden("model_llama_w8_e2e.pt")
    recipe = mr.recipe_from_qcfg_json(g["qcfg"])
    emb = torch.stack([mr.embed(g["state_dict"], g["cfg"], s)[0] for s in g["samples"]])
    res = mr.calibrate(g["state_dict"], g["cfg"], recipe, g["act_dict"], emb, mode="e2e", epochs=g["epochs"], **g["hp"])
    for i in g["learned"]:
        for k, v in g["learned"][i].items():
            assert torch.allclose(res["params"][i][k], v.float(), rtol=0, atol=1e-6), (i, k)
    assert res["act_dict"].keys() == g["act_after"].keys()


@pytest.mark.parametrize("tag", MODEL_GOLDENS)
def test_product_rewrite_matches_reference_qcfg(tag):
    """create_sim_qmodel + update_quant_cfg give the reference's default_qcfg.json and module paths."""
    from mobilequant_b200.quantization import qmodule as Q
    g = load_golden(f"model_{tag}.pt")
    m = product_model(g)
    w = g["w_cfg"]
    Q.create_sim_qmodel(m, Q.QuantConfig(bitwidth=w["bits"], is_symmetric=w["sym"], is_per_channel=w["per_channel"]),
                        Q.QuantConfig(bitwidth=8))
    Q.update_quant_cfg(m)
    assert Q.export_qcfg(m) == g["qcfg"]
    Q.set_scale_and_offset(m, g["act_dict"], "parameter")
    exported = Q.export_act_range(m)
    assert exported.keys() == g["act_after"].keys()
    for n in exported:
        assert exported[n].keys() == g["act_after"][n].keys()
    # a second rewrite round-trips through the JSON schema
    Q.update_qcfg(m, json.loads(json.dumps(g["qcfg"])))
    assert Q.export_qcfg(m) == g["qcfg"]
    sd_names = set(Q.create_fp_model(m).state_dict().keys())
    assert set(g["state_dict"].keys()) - {"lm_head.weight"} <= sd_names | {"lm_head.weight"}


def test_quantconfig_string_schema():
    from mobilequant_b200.quantization.qmodule import QuantConfig
    c = QuantConfig(bitwidth=4, is_symmetric=True, is_per_channel=True)
    d = c.to_dict()
    assert d == {"bitwidth": "4", "group_size": "-1", "is_symmetric": "True", "is_per_channel": "True", "is_dynamic": "False"}
    assert QuantConfig.from_dict(d) == c


def test_get_lr_schedule():
    from mobilequant_b200.quantization.algorithm import get_lr
    assert get_lr(1e-3, 1e-4, 0, 10, 100) == 0.0
    assert get_lr(1e-3, 1e-4, 5, 10, 100) == pytest.approx(5e-4)
    assert get_lr(1e-3, 1e-4, 10, 10, 100) == pytest.approx(1e-3)
    assert get_lr(1e-3, 1e-4, 100, 10, 100) == pytest.approx(1e-4)
    assert get_lr(1e-3, 1e-4, 101, 10, 100) == 1e-4
    for it in range(0, 60):
        assert get_lr(1e-3, 1e-4, it, 0, 50) == pytest.approx(mr.get_lr(1e-3, 1e-4, it, 0, 50))


def test_json_artifact_format(tmp_path):
    from mobilequant_b200.utils.io import json_save, json_load
    p = tmp_path / "act_dict.json"
    json_save(str(p), {"b": {"output": [0.0, 1.0]}, "a": {"input": [-1.5, 2.0]}})
    txt = p.read_text()
    assert txt.index('"a"') < txt.index('"b"') and "\n    " in txt       # sort_keys + indent=4 (io.py:34-36)
    assert json_load(str(p))["a"]["input"] == [-1.5, 2.0]


def test_generate_qcfg_matches_reference_script():
    """ptq/generate_qcfg.py: same flags and the same default_qcfg.json as the reference's script for the flag sets its
    experiment scripts use (golden written by oracle/make_golden_qcfg.py with the reference's own Q* classes)."""
    import os
    from helpers import GOLDEN
    from mobilequant_b200.model import HFConfig, HFForCausalLM
    from mobilequant_b200.quantization.qmodule import QuantConfig
    from mobilequant_b200.ptq import generate_qcfg as G
    gold = json.load(open(os.path.join(GOLDEN, "generate_qcfg.json")))
    assert len(gold) >= 5
    for name, case in gold.items():
        model = HFForCausalLM(HFConfig(**case["cfg"], use_cache=False, use_matmul_as_module=True, l2norm_as_rmsnorm=True))
        wb, wg, wpc, wsym = case["weight"]
        ab, asym, adyn = case["act"]
        got = G.generate_qcfg(model, QuantConfig(bitwidth=wb, group_size=wg, is_per_channel=wpc, is_symmetric=wsym),
                              QuantConfig(bitwidth=ab, is_symmetric=asym, is_dynamic=adyn), *case["switches"])
        assert got == case["default_qcfg"], name
    # the reference's command lines parse unchanged (experiments/*: --use_16bit_softmax_input --use_16bit_softmax_output)
    a = G.build_parser().parse_args(["--hf_path", "x", "--use_16bit_softmax_input", "--use_16bit_softmax_output", "--use_16bit_output_for_mlp"])
    assert a.weight_bitwidth == 8 and a.act_bitwidth == 8 and a.use_16bit_output_for_mlp
    with pytest.raises(SystemExit):
        G.build_parser().parse_args(["--hf_path", "x", "--use_8bit_softmax_input"])      # that flag belongs to ptq/mobilequant.py


def test_calibration_set_is_never_silently_random(tmp_path):
    """ptq/mobilequant.py: a cached dataloader is used when present; without it only --calib_dataset random runs."""
    import types, logging
    from mobilequant_b200.ptq import mobilequant as M
    cfg = types.SimpleNamespace(vocab_size=100, bos_token_id=1)
    log = logging.getLogger("t")
    args = M.build_parser().parse_args(["--hf_path", "/x/llama-1.1b-chat", "--nsamples", "3", "--seqlen", "8", "--cache_dir", str(tmp_path)])
    assert args.calib_dataset == "pile"
    with pytest.raises(FileNotFoundError, match="calib_dataset random"):
        M.load_calibration_set(args, cfg, log)
    saved = [(torch.full((1, 8), i), None) for i in range(3)]
    torch.save(saved, tmp_path / "dataloader_llama_pile_3.cache")                      # ptq/mobilequant.py:221
    got = M.load_calibration_set(args, cfg, log)
    assert len(got) == 3 and torch.equal(got[2][0], saved[2][0])
    args.calib_dataset = "random"
    args.nsamples = 4
    got = M.load_calibration_set(args, cfg, log)
    assert len(got) == 4 and got[0][0].shape == (1, 8) and int(got[0][0].min()) >= 2
    with pytest.raises(NotImplementedError, match="lm-eval"):
        M.main(["--hf_path", "/x", "--tasks", "wikitext"])


def test_encodings_writer_matches_reference_functions():
    """device/encodings.py against the UNMODIFIED text of the reference's update_encodings & co (device/utils.py:278-560, executed by
    oracle/make_golden_encodings.py on synthetic AIMET-style files) and the kv-cache block of device/calibrate.py:275-285."""
    import copy, os
    from helpers import GOLDEN
    from mobilequant_b200.device import encodings as E
    gold = json.load(open(os.path.join(GOLDEN, "encodings.json")))
    assert len(gold) >= 3
    for tag, c in gold.items():
        got, untouched = E.update_encodings(copy.deepcopy(c["ori_encodings"]), c["act_dict"], c["num_blocks"], c["q_proj_factor"],
                                            c["impl_sym_pch_as_slinear"])
        assert got == c["updated"], tag
        assert untouched == ["module_add_mask", "module_embedding"], tag
        assert E.kv_cache_encodings(c["act_dict"], c["num_blocks"], 8) == c["kv_cache"], tag
    # self-contained writer: same numbers under canonical node names, bitwidths taken from default_qcfg.json
    c = gold["llama_plain"]
    mods = {k: {f: {"bitwidth": "16" if ("norm" in k and f == "input") or k.endswith("o_proj") or k.endswith("w2") and f == "output" else "8"}
                for f in ("input", "input2", "output")} for k in c["act_dict"]}
    enc = E.encodings_from_act_dict(c["act_dict"], mods, c["num_blocks"], head_dim=64)
    e = enc["activation_encodings"]["module_matmul_1"]["input"]["1"]
    mn, mx = c["act_dict"]["model.layers.0.self_attn.pv_bmm"]["input2"]
    assert e["min"] == mn and e["max"] == mx and e["scale"] == (mx - mn) / 255 and e["offset"] == int(mn * 255 / (mx - mn))
    q = enc["activation_encodings"]["layers.0.self_attn.q_proj"]["output"]["0"]
    assert q["max"] == c["act_dict"]["model.layers.0.self_attn.q_proj"]["output"][1] * 0.125       # attention scaling folded into q_proj
    assert enc["activation_encodings"]["layers.0.input_layernorm.module_normalize"]["input"]["0"]["bitwidth"] == 16
