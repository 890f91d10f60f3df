"""Per-kernel-class device time of one eager decode step (CUDA events around each launch)."""
import sys, json, torch
sys.path.insert(0, "/root/repo")
import bench
from mobilequant_b200 import kernels as K
from mobilequant_b200.model import HFForCausalLM
from mobilequant_b200.engine import IntEngine
from mobilequant_b200.quantization import qmodule as Q
from mobilequant_b200.ptq.generate_act_range import get_act_range
from mobilequant_b200.ptq.generate_qcfg import default_qcfg
import types
name, T, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
dev = torch.device("cuda:0")
args = types.SimpleNamespace(model=name, layers=None)
cfg = bench.model_cfg(args)
torch.manual_seed(1337)
with torch.device(dev):
    model = HFForCausalLM(cfg).float()
model.eval()
act = get_act_range(model, [bench.synth_ids(1, 256, cfg.vocab_size, 7)])
qcfg = default_qcfg(cfg, Q.QuantConfig(bitwidth=8), Q.QuantConfig(bitwidth=8))
eng = IntEngine(model, qcfg, act, dev)
ids = bench.synth_ids(B, T, cfg.vocab_size, 1).to(dev)
cache = eng.new_cache(B, T)
logits = eng.prefill(ids[:, :T - 8].contiguous(), cache)
tok = logits.argmax(-1)
for _ in range(2):
    logits = eng.decode_step(tok, cache); tok = logits.argmax(-1)
torch.cuda.synchronize()
K.enable_event_timing(True)
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
e0.record()
h = eng._embed(tok).contiguous()
eng.decode_hidden(h, cache)
e1.record()
lg = eng._head(h); tok = lg.argmax(-1)
e2.record()
torch.cuda.synchronize()
per = K.collect_event_timing()
K.enable_event_timing(False)
print(name, "B", B, "T", T, {k: (round(v["ms"], 3), v["n"]) for k, v in per.items()}, "backbone(eager) ms", round(e0.elapsed_time(e1), 3), "head ms", round(e1.elapsed_time(e2), 3))
