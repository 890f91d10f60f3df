"""TEST INFRASTRUCTURE ONLY -- golden vectors for the SmoothQuant initialiser.

Runs the UNMODIFIED function bodies of the reference's ptq/generate_act_scale_shift.py (get_act_scales, get_act_shifts) and
ptq/smoothquant.py (smooth_ln_fcs, smooth_fc_fcs, smooth_lm) on the CPU.  Those scripts parse argv and import datasets /
lm_eval at module level, so the function definitions are lifted out of the source files with `ast` and executed in a
namespace holding the names they use; the model is the reference's own HFForCausalLM (oracle/ref_shim.py).

    python oracle/make_golden_smooth.py        # writes tests/golden/smooth_{llama,stablelm}.pt
"""
import os, sys, ast, types, copy
from functools import partial
import torch, torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_shim import load_reference, ref_config, REF_ROOT

GOLD = os.path.join(ROOT, "tests", "golden")
hm, qm, alg = load_reference()


def lift(path, names, ns):
    src = open(path).read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    return ns


class _Tok:
    def __init__(self, samples):
        self.samples = samples
        self.bos_token_id, self.vocab_size = 1, 256

    def __call__(self, line, **kw):
        return types.SimpleNamespace(input_ids=self.samples[int(line)])


def run(tag, **cfg_over):
    torch.manual_seed(1337)
    cfg = ref_config(hm, vocab_size=256, hidden_size=64, intermediate_size=176, num_hidden_layers=2, num_attention_heads=4,
                     **cfg_over)
    model = hm.HFForCausalLM(cfg).float().eval()
    with torch.no_grad():                                   # make norm biases / qkv biases non-trivial where they exist
        g = torch.Generator().manual_seed(3)
        for n, p in model.named_parameters():
            if n.endswith("bias"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.05)
    sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(11)
    samples = [torch.randint(3, 255, (1, 24), generator=g) for _ in range(3)]
    dataset = [{"text": str(i)} for i in range(len(samples))]
    tok = _Tok(samples)
    ns = dict(torch=torch, nn=nn, partial=partial, tqdm=lambda x, *a, **k: x, HFRMSNorm=hm.HFRMSNorm, HFDecoderLayer=hm.HFDecoderLayer,
              args=types.SimpleNamespace(use_rand_samples=False, seq_len=24), print=lambda *a, **k: None)
    lift(os.path.join(REF_ROOT, "ptq", "generate_act_scale_shift.py"), {"get_act_scales", "get_act_shifts"}, ns)
    lift(os.path.join(REF_ROOT, "ptq", "smoothquant.py"), {"smooth_ln_fcs", "smooth_fc_fcs", "smooth_lm"}, ns)
    act_scales = ns["get_act_scales"](model, tok, dataset, len(samples), 24)
    act_shifts = ns["get_act_shifts"](model, tok, dataset, len(samples), 24)
    out = {}
    for key, kw in (("default", {}), ("alpha075_orig_omni", dict(alpha=0.75, original_omniquant=True))):
        m2 = copy.deepcopy(model)
        ns["smooth_lm"](m2, act_scales, **kw)
        out[key] = dict(kw=kw, changed={k: v.detach().clone() for k, v in m2.state_dict().items() if not torch.equal(v, sd0[k])})
    cd = {k: getattr(cfg, k) for k in ("vocab_size", "hidden_size", "intermediate_size", "num_hidden_layers", "num_attention_heads",
                                       "num_key_value_heads", "hidden_act", "layer_norm_eps", "max_position_embeddings")}
    cd.update(cfg_over)
    torch.save(dict(cfg=cd, state_dict=sd0, samples=samples, act_scales=act_scales, act_shifts=act_shifts, smoothed=out),
               os.path.join(GOLD, f"smooth_{tag}.pt"))
    print(tag, "scales", len(act_scales), "shifts", len(act_shifts), "bytes", os.path.getsize(os.path.join(GOLD, f"smooth_{tag}.pt")))


if __name__ == "__main__":
    run("llama", num_key_value_heads=2)
    run("stablelm", num_key_value_heads=4, norm_class="layernorm", attention_bias=True, use_qkv_bias_only=True, partial_rotary_factor=0.25)
