"""BASELINE config 2 for real: MobileQuant e2e (LET + LWC + LRL) calibration of TinyLlama-1.1B shapes on 512 synthetic samples,
seq 1024, 1 epoch, on one B200.  Wall clock from the first FP-target forward to parameters.pth on disk."""
import sys, os, time, types, tempfile, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mobilequant_b200.model.hf_config import named_config
from mobilequant_b200.model import HFForCausalLM
from mobilequant_b200.quantization import qmodule as Q, algorithm as A
from mobilequant_b200.ptq.generate_act_range import get_act_range
from bench import synth_ids
dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = True
nsamples = int(sys.argv[1]) if len(sys.argv) > 1 else 512
T = 1024
cfg = named_config("tinyllama-1.1b")
torch.manual_seed(1337)
with torch.device(dev):
    model = HFForCausalLM(cfg).float()
model.eval()
t_a = time.perf_counter()
act = get_act_range(model, [synth_ids(1, T, cfg.vocab_size, 7 + i) for i in range(16)])
torch.cuda.synchronize(); t_act = time.perf_counter() - t_a
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
msgs = []
class L:
    def info(self, *a, **k): msgs.append(" ".join(str(x) for x in a))
torch.cuda.synchronize(); t0 = time.perf_counter()
A.e2equant(args, model, loader, L(), device=dev)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
sd = torch.load(os.path.join(out, "parameters.pth"), weights_only=False)
finite = all(torch.isfinite(v).all().item() for d in sd.values() for v in d.values())
res = {"config": "TinyLlama-1.1B shapes, W8A8, e2equant LET+LWC+LRL, bs 1, seq 1024, 1 epoch", "samples": nsamples, "seconds": dt,
       "samples_per_s": nsamples / dt, "act_range_16_samples_s": t_act, "parameters_pth_layers": len(sd), "all_finite": finite,
       "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30, "last_log": msgs[-3:]}
print(json.dumps(res))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/r2_calib512.json", "w"), indent=1)
