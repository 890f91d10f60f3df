"""TEST INFRASTRUCTURE ONLY -- golden vectors for the sim-layout export.

Executes the UNMODIFIED statements of the reference's device/convert_sim.py (the state-dict folding inside main(), from
`hf_state  = model_hf.state_dict()` to the `out_states` loop) on a tiny reference HFForCausalLM / SimModel pair on the CPU.

    python oracle/make_golden_sim.py        # writes tests/golden/sim_export.pt
"""
import os, sys, math, types, textwrap, importlib
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_shim import load_reference, ref_config, REF_ROOT

hm, qm, alg = load_reference()
sm = importlib.import_module("mobilellm.model.sim_model")
src = open(os.path.join(REF_ROOT, "device", "convert_sim.py")).read().splitlines()
i0 = next(i for i, l in enumerate(src) if l.strip().startswith("hf_state  = model_hf.state_dict()"))
i1 = next(i for i, l in enumerate(src) if l.strip().startswith("out_states[k] = v.cpu()"))
block = textwrap.dedent("\n".join(l for l in src[i0:i1 + 1] if "print(msg)" not in l))

cases = {}
for tag, model_name, impl in (("llama_w8_slinear", "llama-tiny-w8a8", True), ("llama_w4", "llama-tiny-w4a8", False), ("gemma_w8_slinear", "gemma-tiny-w8a8", True)):
    torch.manual_seed(1337)
    cfg = ref_config(hm, vocab_size=128, hidden_size=64, intermediate_size=96, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2)
    model_hf = hm.HFForCausalLM(cfg).float().eval()
    sim_config = sm.SimConfig(n_layer=2, n_head=4, n_kv_head=2, head_dim=16, n_embd=64, intermediate_size=96, vocab_size=128, block_size=32,
                              impl_sym_pch_as_slinear=impl)
    model_sim = sm.SimModel(sim_config).to(torch.float32)
    ns = dict(torch=torch, math=math, model_hf=model_hf, model_sim=model_sim, sim_config=sim_config, args=types.SimpleNamespace(model_name=model_name))
    exec(block, ns)
    cases["hf_state"] = {k: v.detach().clone() for k, v in model_hf.state_dict().items()}      # same seed: identical for every case
    cases[tag] = dict(model_name=model_name, impl_sym_pch_as_slinear=impl, out_states=ns["out_states"])
    print(tag, len(ns["out_states"]))
torch.save(cases, os.path.join(ROOT, "tests", "golden", "sim_export.pt"))
print(os.path.getsize(os.path.join(ROOT, "tests", "golden", "sim_export.pt")))
