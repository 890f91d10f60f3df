"""default_qcfg.json writer (mirror of ptq/generate_qcfg.py:16-118: same flags, same mixed-precision rules, same JSON).

Two recipes exist in the reference and they are NOT the same:
  * ptq/generate_qcfg.py:85-113 `create_mixed_precision_model` -- this script: w2 per-channel + 16-bit output, o_proj 16-bit
    output, norm input / weight 16 bit, softmax input / output 16 bit only with --use_16bit_softmax_input / _output, optional
    --use_16bit_output_for_mlp (gemma);
  * ptq/mobilequant.py:175-201 `update_quant_cfg` -- applied by the PTQ entry point before calibration: 16-bit softmax by
    default (--use_8bit_softmax_input / _output switch it off), norm weights forced per-tensor asymmetric
    (quantization/qmodule.py:update_quant_cfg here).  `default_qcfg()` below returns what that entry point exports.
"""
import argparse, os
from ..model.hf_model import HFForCausalLM
from ..quantization.qmodule import (QuantConfig, QLinear, QRMSNorm, QLayerNorm, QMatMul, QSiLU, QGELU, create_sim_qmodel,
                                    update_quant_cfg, export_qcfg)
from ..utils.io import json_save


def create_mixed_precision_model(model, use_16bit_output_for_mlp=False, use_16bit_softmax_input=False, use_16bit_softmax_output=False):
    """ptq/generate_qcfg.py:85-113."""
    for name, module in reversed(list(model._modules.items())):
        if isinstance(module, QLinear):
            if "w2" in name:
                module.weight_quantizer.qcfg.is_per_channel = True
                module.output_quantizer.qcfg.bitwidth = 16
            elif "o_proj" in name:
                module.output_quantizer.qcfg.bitwidth = 16
            if use_16bit_output_for_mlp and ("w1" in name or "w3" in name):
                module.output_quantizer.qcfg.bitwidth = 16
        elif isinstance(module, (QRMSNorm, QLayerNorm)):
            module.input_quantizer.qcfg.bitwidth = 16
            module.weight_quantizer.qcfg.bitwidth = 16
        elif isinstance(module, QMatMul):
            if "qk_bmm" in name and use_16bit_softmax_input:
                module.output_quantizer.qcfg.bitwidth = 16
            if "pv_bmm" in name and use_16bit_softmax_output:
                module.input_quantizer.qcfg.bitwidth = 16
        elif isinstance(module, (QSiLU, QGELU)):
            if module.input_quantizer is not None:           # already quantised by w1 (create_sim_qmodel drops it, qm:853-858)
                module.input_quantizer.enable = False
        elif len(list(module.children())) > 1:
            create_mixed_precision_model(module, use_16bit_output_for_mlp, use_16bit_softmax_input, use_16bit_softmax_output)
    return model


def generate_qcfg(model, weight_qcfg, act_qcfg, use_16bit_output_for_mlp=False, use_16bit_softmax_input=False, use_16bit_softmax_output=False):
    """The default_qcfg.json payload of ptq/generate_qcfg.py for `model` (converted in place to its Q* form)."""
    model = create_sim_qmodel(model, weight_qcfg, act_qcfg)
    create_mixed_precision_model(model, use_16bit_output_for_mlp, use_16bit_softmax_input, use_16bit_softmax_output)
    return export_qcfg(model)


def default_qcfg(config, weight_qcfg, act_qcfg, use_8bit_softmax_input=False, use_8bit_softmax_output=False):
    """default_qcfg.json as ptq/mobilequant.py exports it (update_quant_cfg recipe, :175-201, :243-244) for an architecture,
    without materialising its weights (meta device)."""
    import torch
    with torch.device("meta"):
        shell = HFForCausalLM(config)
    shell = create_sim_qmodel(shell, weight_qcfg, act_qcfg)
    update_quant_cfg(shell, use_8bit_softmax_input, use_8bit_softmax_output)
    return export_qcfg(shell)


def add_quant_args(p):
    """Quantisation flags of ptq/mobilequant.py:40-50,73-74 (its defaults: W4 / A16)."""
    p.add_argument("--weight_bitwidth", type=int, default=4)
    p.add_argument("--weight_group_size", type=int, default=-1)
    p.add_argument("--weight_is_per_channel", default=False, action="store_true")
    p.add_argument("--weight_is_symmetric", default=False, action="store_true")
    p.add_argument("--weight_is_dynamic", default=False, action="store_true")
    p.add_argument("--act_bitwidth", type=int, default=16)
    p.add_argument("--act_group_size", type=int, default=-1)
    p.add_argument("--act_is_per_channel", default=False, action="store_true")
    p.add_argument("--act_is_symmetric", default=False, action="store_true")
    p.add_argument("--act_is_dynamic", default=False, action="store_true")
    p.add_argument("--use_8bit_softmax_input", default=False, action="store_true")
    p.add_argument("--use_8bit_softmax_output", default=False, action="store_true")


def qcfgs_from_args(args):
    w = QuantConfig(args.weight_bitwidth, args.weight_group_size, args.weight_is_symmetric, args.weight_is_per_channel,
                    getattr(args, "weight_is_dynamic", False))
    a = QuantConfig(args.act_bitwidth, args.act_group_size, args.act_is_symmetric, args.act_is_per_channel, args.act_is_dynamic)
    return w, a


def build_parser():
    """ptq/generate_qcfg.py:24-39 (defaults W8 / A8; the 16-bit switches are opt-in here, unlike ptq/mobilequant.py)."""
    p = argparse.ArgumentParser()
    p.add_argument("--hf_path", type=str, help="path of the hf model")
    p.add_argument("--weight_bitwidth", type=int, default=8)
    p.add_argument("--weight_group_size", type=int, default=-1)
    p.add_argument("--weight_is_per_channel", default=False, action="store_true")
    p.add_argument("--weight_is_symmetric", default=False, action="store_true")
    p.add_argument("--act_bitwidth", type=int, default=8)
    p.add_argument("--act_group_size", type=int, default=-1)
    p.add_argument("--act_is_per_channel", default=False, action="store_true")
    p.add_argument("--act_is_symmetric", default=False, action="store_true")
    p.add_argument("--act_is_dynamic", default=False, action="store_true")
    p.add_argument("--use_16bit_output_for_mlp", default=False, action="store_true", help="for gemma")
    p.add_argument("--use_16bit_softmax_input", default=False, action="store_true", help="for phi")
    p.add_argument("--use_16bit_softmax_output", default=False, action="store_true", help="for phi")
    p.add_argument("--output_dir", default=None, type=str)
    return p


def main(argv=None):
    args = build_parser().parse_args(argv)
    if args.weight_group_size != -1:
        assert args.weight_is_per_channel, ("weight_is_per_channel should be activated if we'd like to use per-group quantization "
                                            "(i.e. weight_group_size != -1)")                       # ptq/generate_qcfg.py:45-46
    model = HFForCausalLM.from_pretrained(args.hf_path, use_matmul_as_module=True, l2norm_as_rmsnorm=True)
    w, a = qcfgs_from_args(args)
    out = args.output_dir or args.hf_path
    os.makedirs(out, exist_ok=True)
    json_save(os.path.join(out, "default_qcfg.json"),
              generate_qcfg(model, w, a, args.use_16bit_output_for_mlp, args.use_16bit_softmax_input, args.use_16bit_softmax_output))


if __name__ == "__main__":
    main()
