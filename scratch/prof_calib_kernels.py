"""One eager e2equant step on a 2-layer TinyLlama-shape model (seq 1024): the launches ncu captures for the calibration kernels."""
import sys, os, types, tempfile, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["MQ_CUDA_GRAPH"] = "0"
os.environ["MQ_WPREP_STREAM"] = "0"
from mobilequant_b200.model.hf_config import named_config
from mobilequant_b200.model import HFForCausalLM
from mobilequant_b200.quantization import qmodule as Q, algorithm as A
from mobilequant_b200.ptq.generate_act_range import get_act_range
from bench import synth_ids
dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = True
cfg = named_config("tinyllama-1.1b", num_hidden_layers=2)
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
args = types.SimpleNamespace(nsamples=2, seqlen=T, batch_size=1, epochs=1, warmup_epochs=0, deactive_amp=True, let=True,
                             lwc=True, lrl=True, use_shift=False, aug_loss=False, let_lr=1e-3, lwc_lr=1e-2, lrl_lr=1e-6,
                             let_min_lr=1e-4, lwc_min_lr=1e-3, lrl_min_lr=1e-7, wd=0.0, resume=None, cache_in_gpu=True,
                             original_omniquant=False, dtype=torch.float32, output_dir=tempfile.mkdtemp())


class L:
    def info(self, *a, **k):
        pass


A.e2equant(args, model, [(synth_ids(1, T, cfg.vocab_size, 50 + i), None) for i in range(2)], L(), device=dev)
torch.cuda.synchronize()
