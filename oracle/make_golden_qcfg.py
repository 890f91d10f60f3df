"""TEST INFRASTRUCTURE ONLY -- default_qcfg.json as the reference's ptq/generate_qcfg.py writes it (its Q* classes, its
create_sim_qmodel / export_qcfg, the mixed-precision closure restated in oracle/ref_shim.py), for the flag sets the
reference's experiment scripts use.  Writes tests/golden/generate_qcfg.json.

    python oracle/make_golden_qcfg.py
"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle.ref_shim import load_reference, ref_config, ref_create_mixed_precision_model
from oracle.make_golden import TINY, TINY_MHA, TINY_GELU

hm, qm, alg = load_reference()

CASES = {
    # name: (model config, weight (bits, group, per_channel, symmetric), act (bits, sym, dynamic), 16-bit switches (mlp, sm_in, sm_out))
    "llama_w8a8_default": (TINY, (8, -1, False, False), (8, False, False), (False, False, False)),
    "llama_w8a8_softmax16": (TINY, (8, -1, False, False), (8, False, False), (False, True, True)),      # experiments/*/: --use_16bit_softmax_input --use_16bit_softmax_output
    "llama_w4a8_pc_sym": (TINY, (4, -1, True, True), (8, False, False), (False, True, True)),
    "stablelm_w8a8_softmax16": (TINY_MHA, (8, -1, False, False), (8, False, False), (False, True, True)),
    "gemma_w8a8_mlp16": (TINY_GELU, (8, -1, False, False), (8, False, False), (True, True, True)),
}


def main():
    out = {}
    for name, (cfgd, w, a, sw) in CASES.items():
        torch.manual_seed(1337)
        model = hm.HFForCausalLM(ref_config(hm, **cfgd))
        wq = qm.QuantConfig(); wq.bitwidth, wq.group_size, wq.is_per_channel, wq.is_symmetric = w
        aq = qm.QuantConfig(); aq.bitwidth, aq.is_symmetric, aq.is_dynamic = a
        model = qm.create_sim_qmodel(model, wq, aq)
        ref_create_mixed_precision_model(qm, model, *sw)
        out[name] = dict(cfg=cfgd, weight=list(w), act=list(a), switches=list(sw), default_qcfg=qm.export_qcfg(model))
    with open(os.path.join(ROOT, "tests", "golden", "generate_qcfg.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("generate_qcfg.json:", {k: len(v["default_qcfg"]) for k, v in out.items()})


if __name__ == "__main__":
    main()
