"""Where does one e2equant step go?  (TinyLlama shapes, T=1024, bs 1)"""
import sys, os, time, types, tempfile, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mobilequant_b200.model.hf_config import named_config
from mobilequant_b200.model import HFForCausalLM
from mobilequant_b200.quantization import qmodule as Q, algorithm as A
from mobilequant_b200.ptq.generate_act_range import get_act_range
from mobilequant_b200.ptq.generate_qcfg import default_qcfg
from bench import synth_ids

layers = int(sys.argv[1]) if len(sys.argv) > 1 else 22
nsamples = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = True
cfg = named_config("tinyllama-1.1b", num_hidden_layers=layers)
T = 1024
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
    def info(self, *a, **k):
        print(*a)


from torch.profiler import profile, ProfilerActivity
torch.cuda.synchronize(); t0 = time.perf_counter()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    A.e2equant(args, model, loader, L(), device=dev)
    torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"e2equant {nsamples} samples {layers} layers: {dt:.2f} s -> {nsamples/dt:.2f} samples/s (under profiler)")
import collections
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        a = agg[e.name[:90]]; a[0] += 1; a[1] += e.device_time / 1e3
tot = sum(v[1] for v in agg.values())
print(f"GPU kernel time total {tot:.1f} ms over {nsamples} samples = {tot/nsamples:.1f} ms/sample (FP pass + training step + fuse)")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{t/nsamples:8.3f} ms/sample {c/nsamples:8.1f} launches/sample  {k}")

print("---- by aten op and input shape (self device time)")
rows = sorted(prof.key_averages(group_by_input_shape=True), key=lambda e: -e.self_device_time_total)[:60]
for e in rows:
    print(f"{e.self_device_time_total/1e3/nsamples:8.3f} ms/sample {e.count/nsamples:8.1f} calls/sample  {e.key[:40]:40s} {str(e.input_shapes)[:110]}")
