"""TEST INFRASTRUCTURE ONLY -- imports the UNMODIFIED reference from /root/reference.

Only usable in the build container (the GPU box has no /root/reference).  Used by
oracle/make_golden.py to (a) validate oracle/fakequant_ref.py + oracle/int_ref.py against the
reference's own code and (b) write the committed fixtures under tests/golden/.

The shim applies the pre-import patches listed in SURVEY.md 8c (transformers 5.5 vs the 4.41 the
reference pins).  Nothing here is product code and nothing under mobilequant_b200/ may import it.
"""
import sys, types, importlib
import torch, torch.nn as nn

REF_ROOT = "/root/reference"


def load_reference():
    if "mobilellm" in sys.modules and getattr(sys.modules["mobilellm"], "__mq_shimmed__", False):
        m = sys.modules
        return m["mobilellm.model.hf_model"], m["mobilellm.quantization.qmodule"], m["mobilellm.quantization.algorithm"]
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import transformers.cache_utils as cu, transformers.utils.import_utils as iu
    import transformers.utils as tu
    from transformers.activations import ACT2FN
    if not hasattr(cu, "SinkCache"):
        cu.SinkCache = type("SinkCache", (cu.Cache,), {})           # hm:12
    if not hasattr(iu, "is_torch_fx_available"):
        iu.is_torch_fx_available = lambda: False                    # hm:35
    if not hasattr(tu, "is_torch_fx_available"):
        tu.is_torch_fx_available = lambda: False
    ACT2FN["silu"] = nn.SiLU                                        # 4.41 behaviour, qm:853
    # termcolor / lm_eval are only needed by utils.io / utils.bench
    if "termcolor" not in sys.modules:
        tc = types.ModuleType("termcolor"); tc.colored = lambda s, *a, **k: s
        sys.modules["termcolor"] = tc
    hm = importlib.import_module("mobilellm.model.hf_model")
    qm = importlib.import_module("mobilellm.quantization.qmodule")
    alg = importlib.import_module("mobilellm.quantization.algorithm")
    alg.map_layers_to_multi_gpus = lambda layers: [setattr(l, "device", torch.device("cpu")) for l in layers]
    sys.modules["mobilellm"].__mq_shimmed__ = True
    return hm, qm, alg


def ref_config(hm, **kw):
    base = dict(vocab_size=512, hidden_size=128, intermediate_size=352, num_hidden_layers=2,
                num_attention_heads=4, num_key_value_heads=2, hidden_act="silu", layer_norm_eps=1e-5,
                use_cache=False, use_matmul_as_module=True, l2norm_as_rmsnorm=True,
                max_position_embeddings=2048)
    base.update(kw)
    cfg = hm.HFConfig(**base)
    cfg._attn_implementation = "eager"
    return cfg


def ref_update_quant_cfg(qm, model, use_8bit_softmax_input=False, use_8bit_softmax_output=False):
    """Restatement of the script-local closure ptq/mobilequant.py:175-201 (it cannot be imported:
    the script parses argv and needs lm_eval at import time)."""
    for name, module in reversed(model._modules.items()):
        if isinstance(module, qm.QLinear):
            if any(k in name for k in ("q_proj", "k_proj", "v_proj", "o_proj", "w1", "w3")):
                module.input_quantizer = None
            if "w2" in name:
                module.weight_quantizer.qcfg.is_per_channel = True
                module.output_quantizer.qcfg.bitwidth = 16
            elif "o_proj" in name:
                module.output_quantizer.qcfg.bitwidth = 16
        elif isinstance(module, (qm.QRMSNorm, qm.QLayerNorm)):
            module.input_quantizer.qcfg.bitwidth = 16
            module.weight_quantizer.qcfg.bitwidth = 16
            module.weight_quantizer.qcfg.is_symmetric = False
            module.weight_quantizer.qcfg.is_per_channel = False
        elif isinstance(module, qm.QMatMul):
            if "qk_bmm" in name and not use_8bit_softmax_input:
                module.output_quantizer.qcfg.bitwidth = 16
            if "pv_bmm" in name and not use_8bit_softmax_output:
                module.input_quantizer.qcfg.bitwidth = 16
        elif isinstance(module, (qm.QSiLU, qm.QGELU)):
            module.input_quantizer = None
        elif len(list(module.children())) > 1:
            ref_update_quant_cfg(qm, module, use_8bit_softmax_input, use_8bit_softmax_output)
    return model


def ref_create_mixed_precision_model(qm, model, use_16bit_output_for_mlp=False, use_16bit_softmax_input=False, use_16bit_softmax_output=False):
    """Restatement of the script-local closure ptq/generate_qcfg.py:85-113 (the script parses argv and loads a tokenizer at
    import time, so it cannot be imported); the Q* classes it touches are the reference's own."""
    for name, module in reversed(model._modules.items()):
        if isinstance(module, qm.QLinear):
            if "w2" in name:
                model._modules[name].weight_quantizer.qcfg.is_per_channel = True
                model._modules[name].output_quantizer.qcfg.bitwidth = 16
            elif "o_proj" in name:
                model._modules[name].output_quantizer.qcfg.bitwidth = 16
            if use_16bit_output_for_mlp and ("w1" in name or "w3" in name):
                model._modules[name].output_quantizer.qcfg.bitwidth = 16
        elif isinstance(module, (qm.QRMSNorm, qm.QLayerNorm)):
            model._modules[name].input_quantizer.qcfg.bitwidth = 16
            model._modules[name].weight_quantizer.qcfg.bitwidth = 16
        elif isinstance(module, qm.QMatMul):
            if "qk_bmm" in name and use_16bit_softmax_input:
                model._modules[name].output_quantizer.qcfg.bitwidth = 16
            if "pv_bmm" in name and use_16bit_softmax_output:
                model._modules[name].input_quantizer.qcfg.bitwidth = 16
        elif isinstance(module, (qm.QSiLU, qm.QGELU)):
            if model._modules[name].input_quantizer is not None:
                model._modules[name].input_quantizer.enable = False
        elif len(list(module.children())) > 1:
            ref_create_mixed_precision_model(qm, module, use_16bit_output_for_mlp, use_16bit_softmax_input, use_16bit_softmax_output)
    return model
