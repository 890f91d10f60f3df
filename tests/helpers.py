import os
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODEL_GOLDENS = ["llama_w8_e2e", "llama_w4_omni", "stablelm_w8_omni", "gemma_w8_e2e"]


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


def product_model(gold, device=None):
    """The product HFForCausalLM carrying the golden's reference weights."""
    from mobilequant_b200.model import HFConfig, HFForCausalLM
    cfg = dict(gold["cfg"])
    m = HFForCausalLM(HFConfig(**cfg, use_cache=False, use_matmul_as_module=True, l2norm_as_rmsnorm=True))
    missing, unexpected = m.load_state_dict(gold["state_dict"], strict=False)
    assert not unexpected, unexpected
    assert all("lm_head" in k for k in missing) or not missing, missing
    m = m.float().eval()
    return m.to(device) if device is not None else m


def sim_qmodel(gold, device, lrl=True):
    from mobilequant_b200.quantization import qmodule as Q
    m = product_model(gold)
    w = gold["w_cfg"]
    Q.create_sim_qmodel(m, Q.QuantConfig(bitwidth=w["bits"], is_symmetric=w["sym"], is_per_channel=w["per_channel"]),
                        Q.QuantConfig(bitwidth=8))
    for p in m.parameters():
        p.requires_grad = False
    Q.update_quant_cfg(m)
    Q.set_scale_and_offset(m, gold["act_dict"], "parameter" if lrl else None)
    return m.to(device)


def calib_args(gold, out_dir, **kw):
    import types
    hp = gold["hp"]
    a = types.SimpleNamespace(nsamples=len(gold["samples"]), seqlen=gold["samples"][0].shape[1], batch_size=1,
                              epochs=gold["epochs"], warmup_epochs=0, deactive_amp=True, let=True, lwc=True, lrl=True,
                              use_shift=False, aug_loss=False, wd=0.0, resume=None, cache_in_gpu=True,
                              original_omniquant=False, dtype=torch.float32, output_dir=str(out_dir), **hp)
    for k, v in kw.items():
        setattr(a, k, v)
    return a
