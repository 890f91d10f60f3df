"""MobileQuant PTQ entry point (mirror of ptq/mobilequant.py: same flags, same artefacts).

    python -m mobilequant_b200.ptq.mobilequant --hf_path <dir> --mode e2e --lwc --let --lrl \\
        --weight_bitwidth 8 --act_bitwidth 8 --nsamples 512 --seqlen 1024 --epochs 1 --output_dir out

Calibration data: the reference's cached loader `{cache_dir}/dataloader_{family}_{calib_dataset}_{nsamples}.cache`
(ptq/mobilequant.py:221-227, a torch-saved list of (input_ids [1, seqlen], target)) is used when it exists.  There is
no tokenizer / dataset access here, so without that cache only `--calib_dataset random` (random token ids of the
reference's own convention, generate_act_range.py:106-108 / device/export.py:116) can run -- any other choice raises
instead of silently calibrating on random tokens.  Evaluation (--tasks) needs the lm-eval fork and is rejected.
Under torchrun the calibration set is sharded over ranks (sample-parallel) instead of the reference's layer-sharding.
"""
import argparse, os, random, time
import torch
import torch.distributed as dist
from ..model.hf_model import HFForCausalLM
from ..quantization.algorithm import omniquant, e2equant
from ..quantization.qmodule import (create_fp_model, export_act_range, create_sim_qmodel, set_scale_and_offset,
                                    export_qcfg, update_quant_cfg)
from ..utils.io import json_load, json_save, create_logger
from .generate_qcfg import add_quant_args, qcfgs_from_args
from .generate_act_range import random_samples

STR_TO_DTYPE = {"float16": torch.float16, "float32": torch.float32, "bfloat16": torch.bfloat16}


def build_parser():
    p = argparse.ArgumentParser()
    p.add_argument("--hf_path", type=str, default=None)
    p.add_argument("--dtype", type=str, default=None)
    p.add_argument("--output_dir", default="results/quant", type=str)
    p.add_argument("--cache_dir", default="./cache", type=str)
    p.add_argument("--resume", type=str, default=None)
    p.add_argument("--calib_dataset", type=str, default="pile", choices=["wikitext2", "pile", "random"])
    p.add_argument("--nsamples", type=int, default=128)
    p.add_argument("--seqlen", type=int, default=2048)
    p.add_argument("--act_dict_path", type=str, default=None)
    p.add_argument("--override_qcfg_path", type=str, default=None)
    add_quant_args(p)
    p.add_argument("--let", default=False, action="store_true")
    p.add_argument("--lwc", default=False, action="store_true")
    p.add_argument("--lrl", default=False, action="store_true")
    p.add_argument("--let_lr", type=float, default=1e-3)
    p.add_argument("--lwc_lr", type=float, default=1e-2)
    p.add_argument("--lrl_lr", type=float, default=1e-6)
    p.add_argument("--let_min_lr", type=float, default=1e-3)
    p.add_argument("--lwc_min_lr", type=float, default=1e-2)
    p.add_argument("--lrl_min_lr", type=float, default=1e-6)
    p.add_argument("--wd", type=float, default=0)
    p.add_argument("--epochs", type=int, default=10)
    p.add_argument("--warmup_epochs", type=int, default=0)
    p.add_argument("--use_shift", default=False, action="store_true")
    p.add_argument("--aug_loss", default=False, action="store_true")
    p.add_argument("--deactive_amp", action="store_true")
    p.add_argument("--batch_size", type=int, default=1)
    p.add_argument("--num_fewshot", type=int, default=0)
    p.add_argument("--tasks", default="", type=str, help="not supported (lm-eval); must stay empty")
    p.add_argument("--mode", default="omniquant", type=str, choices=["e2e", "omniquant"])
    p.add_argument("--original_omniquant", default=False, action="store_true")
    p.add_argument("--cache_in_gpu", default=False, action="store_true")
    return p


def quantize(args, model, dataloader, logger, act_dict):
    """ptq/mobilequant.py:153-246 from an already-loaded float model to the exported artefacts."""
    weight_qcfg, act_qcfg = qcfgs_from_args(args)
    model = create_sim_qmodel(model, weight_qcfg, act_qcfg)
    for p in model.parameters():
        p.requires_grad = False
    update_quant_cfg(model, args.use_8bit_softmax_input, args.use_8bit_softmax_output)
    if not args.act_is_dynamic:
        set_scale_and_offset(model, act_dict, "parameter" if args.lrl else None)
    if args.weight_bitwidth < 16 or args.act_bitwidth < 16:
        logger.info("=== start quantization ===")
        tick = time.time()
        if args.mode.lower() == "e2e":
            e2equant(args, model, dataloader, logger)
        else:
            omniquant(args, model, dataloader, logger, device=torch.device("cuda", torch.cuda.current_device()))
        torch.cuda.synchronize()
        logger.info(time.time() - tick)
    rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    act_out = export_act_range(model)
    qcfg_out = export_qcfg(model)
    if rank == 0 and args.output_dir:
        os.makedirs(args.output_dir, exist_ok=True)
        json_save(os.path.join(args.output_dir, "act_dict.json"), act_out)
        json_save(os.path.join(args.output_dir, "default_qcfg.json"), qcfg_out)
    model = create_fp_model(model)
    if rank == 0 and args.output_dir:
        model.save_pretrained(args.output_dir, safe_serialization=False)
    return model, act_out, qcfg_out


def load_calibration_set(args, config, logger, seed=1337):
    """ptq/mobilequant.py:221-227: the cached dataloader when present; random ids only when asked for by name."""
    family = os.path.basename(os.path.normpath(args.hf_path or "model")).split("-")[0]                 # ptq/mobilequant.py:78
    cache = os.path.join(args.cache_dir or ".", f"dataloader_{family}_{args.calib_dataset}_{args.nsamples}.cache")
    if os.path.exists(cache):
        logger.info(f"load calibration set from {cache}")
        return torch.load(cache, weights_only=False)
    if args.calib_dataset != "random":
        raise FileNotFoundError(
            f"calibration set {cache!r} not found and --calib_dataset {args.calib_dataset} cannot be built here (no tokenizer / "
            "dataset access): create the cache with the reference's get_loaders (ptq/mobilequant.py:221-227) and point --cache_dir "
            "at it, or pass --calib_dataset random to calibrate on uniformly random token ids")
    logger.info("calibrating on uniformly random token ids (--calib_dataset random)")
    samples = random_samples(args.nsamples, args.seqlen, config.vocab_size, config.bos_token_id or 1, seed)
    return [(s, None) for s in samples]


def main(argv=None):
    args = build_parser().parse_args(argv)
    if args.tasks:
        raise NotImplementedError("--tasks needs the lm-eval fork of the reference (eval/harness_eval.py); run it on the exported "
                                  "artefacts (act_dict.json, default_qcfg.json, fp checkpoint)")
    if args.epochs > 0:
        assert args.lwc or args.let or args.lrl
    if (8 <= args.weight_bitwidth < 16) or (8 <= args.act_bitwidth < 16):
        args.deactive_amp = True                                       # ptq/mobilequant.py:122-123
    if "LOCAL_RANK" in os.environ and not dist.is_initialized():
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
        dist.init_process_group("nccl")
    seed = 1337
    random.seed(seed); torch.manual_seed(seed); torch.cuda.manual_seed(seed)
    torch.backends.cuda.matmul.allow_tf32 = True                       # ptq/mobilequant.py:91-92
    torch.backends.cudnn.allow_tf32 = True
    logger = create_logger(args.output_dir)
    logger.info(args)
    model = HFForCausalLM.from_pretrained(args.hf_path, use_matmul_as_module=True, l2norm_as_rmsnorm=True).float()
    args.dtype = torch.float32 if args.dtype is None else STR_TO_DTYPE[args.dtype]
    act_path = args.act_dict_path or os.path.join(args.hf_path, "act_dict.json")
    act_dict = json_load(act_path) if act_path.endswith(".json") else torch.load(act_path)
    dataloader = load_calibration_set(args, model.config, logger, seed)
    quantize(args, model, dataloader, logger, act_dict)


if __name__ == "__main__":
    main()
