"""steady-state e2equant step time: difference of two runs with different sample counts"""
import sys, os, time, types, tempfile, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mobilequant_b200.model.hf_config import named_config
from mobilequant_b200.model import HFForCausalLM
from mobilequant_b200.quantization import qmodule as Q, algorithm as A
from mobilequant_b200.ptq.generate_act_range import get_act_range
from bench import synth_ids
dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = True
mode = sys.argv[1] if len(sys.argv) > 1 else "e2e"
T = 1024
res = {}
for nsamples in (4, 16, 144):
    cfg = named_config("tinyllama-1.1b")
    torch.manual_seed(1337)
    with torch.device(dev):
        model = HFForCausalLM(cfg).float()
    model.eval()
    act = get_act_range(model, [synth_ids(1, T, cfg.vocab_size, 7 + i) for i in range(2)])
    Q.create_sim_qmodel(model, Q.QuantConfig(bitwidth=8), Q.QuantConfig(bitwidth=8))
    for p in model.parameters():
        p.requires_grad = False
    Q.update_quant_cfg(model)
    Q.set_scale_and_offset(model, act, "parameter")
    out = tempfile.mkdtemp()
    args = types.SimpleNamespace(nsamples=nsamples, seqlen=T, batch_size=1, epochs=1, warmup_epochs=0, deactive_amp=True, let=True,
                                 lwc=True, lrl=True, use_shift=False, aug_loss=False, let_lr=1e-3, lwc_lr=1e-2, lrl_lr=1e-6,
                                 let_min_lr=1e-4, lwc_min_lr=1e-3, lrl_min_lr=1e-7, wd=0.0, resume=None, cache_in_gpu=True,
                                 original_omniquant=False, dtype=torch.float32, output_dir=out)
    loader = [(synth_ids(1, T, cfg.vocab_size, 50 + i), None) for i in range(nsamples)]
    class L:
        def info(self, *a, **k): pass
    torch.cuda.synchronize(); t0 = time.perf_counter()
    (A.e2equant if mode == "e2e" else A.omniquant)(args, model, loader, L(), device=dev)
    torch.cuda.synchronize(); res[nsamples] = time.perf_counter() - t0
    print(mode, nsamples, "samples:", round(res[nsamples], 2), "s", flush=True)
    del model
    torch.cuda.empty_cache()
per = (res[144] - res[16]) / 128
print(f"{mode}: steady state {per*1e3:.1f} ms/sample ({1/per:.2f} samples/s); fixed cost {res[16]-16*per:.2f} s; projected 512 samples: {res[16]-16*per+512*per:.1f} s")
