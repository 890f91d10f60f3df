"""default_qcfg.json writer (mirror of ptq/generate_qcfg.py:16-118)."""
import argparse, os
from ..model.hf_model import HFForCausalLM
from ..quantization.qmodule import QuantConfig, create_sim_qmodel, update_quant_cfg, export_qcfg
from ..utils.io import json_save


def generate_qcfg(model, weight_qcfg, act_qcfg, use_8bit_softmax_input=False, use_8bit_softmax_output=False):
    model = create_sim_qmodel(model, weight_qcfg, act_qcfg)
    update_quant_cfg(model, use_8bit_softmax_input, use_8bit_softmax_output)
    return export_qcfg(model)


def default_qcfg(config, weight_qcfg, act_qcfg, **kw):
    """default_qcfg.json content for an architecture without materialising its weights (meta device)."""
    import torch
    with torch.device("meta"):
        shell = HFForCausalLM(config)
    return generate_qcfg(shell, weight_qcfg, act_qcfg, **kw)


def add_quant_args(p):
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
    w = QuantConfig(args.weight_bitwidth, args.weight_group_size, args.weight_is_symmetric, args.weight_is_per_channel, args.weight_is_dynamic)
    a = QuantConfig(args.act_bitwidth, args.act_group_size, args.act_is_symmetric, args.act_is_per_channel, args.act_is_dynamic)
    return w, a


def main(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--hf_path", type=str, required=True)
    p.add_argument("--output_dir", default=None, type=str)
    add_quant_args(p)
    args = p.parse_args(argv)
    model = HFForCausalLM.from_pretrained(args.hf_path, use_matmul_as_module=True, l2norm_as_rmsnorm=True)
    w, a = qcfgs_from_args(args)
    out = args.output_dir or args.hf_path
    os.makedirs(out, exist_ok=True)
    json_save(os.path.join(out, "default_qcfg.json"), generate_qcfg(model, w, a, args.use_8bit_softmax_input, args.use_8bit_softmax_output))


if __name__ == "__main__":
    main()
