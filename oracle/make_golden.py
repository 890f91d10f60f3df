"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.pt by running the UNMODIFIED reference (oracle/ref_shim.py)
in the build container, and checks oracle/fakequant_ref.py against it on the way (bit-exact where stated).

    python oracle/make_golden.py            # writes tests/golden/*.pt, prints the oracle-vs-reference report

Seeds follow ptq/mobilequant.py:87-90 (1337).  Fixtures are kept small (a few hundred KB) so they can be committed.
"""
import os, sys, io, json, copy, types, logging
from collections import OrderedDict
import torch, torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_shim import load_reference, ref_config, ref_update_quant_cfg
from oracle import fakequant_ref as fr

GOLD = os.path.join(ROOT, "tests", "golden")
os.makedirs(GOLD, exist_ok=True)
hm, qm, alg = load_reference()


def gen(seed):
    g = torch.Generator(); g.manual_seed(seed); return g


def golden_quantizer():
    """Quantizer.forward (qm:251-295) static / dynamic / LWC, with autograd grads."""
    cases = []
    g = gen(1337)
    for bits, sym, per_ch, lwc, shape in [
        (8, False, False, False, (4, 37, 64)),     # per-tensor asym activation (default A8)
        (16, False, False, False, (3, 50, 32)),    # 16-bit activation
        (8, True, False, False, (130, 96)),        # symmetric
        (8, False, False, True, (48, 96)),         # W8 per-tensor + LWC
        (8, False, True, True, (48, 96)),          # W8 per-channel + LWC (w2)
        (4, True, True, True, (40, 128)),          # W4 per-channel symmetric + LWC
        (4, False, True, True, (40, 128)),         # W4 per-channel asym + LWC
        (16, False, False, True, (1, 128)),        # norm weight 16-bit per-tensor + LWC
    ]:
        x = torch.randn(shape, generator=g) * 0.7 + 0.1
        if x.dim() == 2 and lwc:
            x = x * 0.02
        qcfg = qm.QuantConfig(bitwidth=bits, is_symmetric=sym, is_per_channel=per_ch)
        q = qm.Quantizer(qcfg)
        xin = x.clone().requires_grad_(True)
        case = dict(bits=bits, sym=sym, per_channel=per_ch, lwc=lwc, x=x)
        if lwc:
            q.enable_lwc(x)
            with torch.no_grad():
                q.upbound_factor.add_(torch.randn(q.upbound_factor.shape, generator=g) * 0.5)
                q.lowbound_factor.add_(torch.randn(q.lowbound_factor.shape, generator=g) * 0.5)
            y = q(xin)
            case.update(up=q.upbound_factor.detach().clone(), low=q.lowbound_factor.detach().clone(),
                        scale=q.scale.detach().clone(), offset=q.offset.detach().clone())
        else:
            lo, hi = x.min().item() * 0.9, x.max().item() * 0.8      # force some clamping
            q.set_scale_offset_from_minmax(lo, hi, "parameter")
            y = q(xin)
            case.update(minmax=[lo, hi], scale=q.scale.detach().clone(), offset=q.offset.detach().clone())
        gy = torch.randn(y.shape, generator=g)
        y.backward(gy)
        case.update(y=y.detach().clone(), gy=gy, gx=xin.grad.clone(), qmin=q.qmin, qmax=q.qmax)
        if lwc:
            case.update(g_up=q.upbound_factor.grad.clone(), g_low=q.lowbound_factor.grad.clone())
        else:
            case.update(g_scale=q.scale.grad.clone(), g_offset=q.offset.grad.clone())
        # --- oracle check (bit-exact forward, grads to fp32 round-off)
        if lwc:
            su, sl = torch.sigmoid(case["up"]), torch.sigmoid(case["low"])
            yo, so, oo, qmin, qmax = fr.dynamic_fake_quant(x, bits, sym, per_ch, su, sl, return_params=True)
        else:
            so, oo, qmin, qmax = fr.scale_offset_from_minmax(lo, hi, bits, sym)
            yo = fr.fake_quant(x, so, oo, qmin, qmax)
        assert torch.equal(yo, case["y"]), ("oracle != reference", bits, sym, per_ch, lwc)
        assert torch.equal(so.reshape(-1), case["scale"].reshape(-1)) and qmin == q.qmin and qmax == q.qmax
        cases.append(case)
    torch.save(cases, os.path.join(GOLD, "quantizer.pt"))
    print(f"quantizer.pt: {len(cases)} cases, oracle forward bit-exact vs reference Quantizer")


if __name__ == "__main__":
    golden_quantizer()


# =================================================================================================================
# model-level goldens
# =================================================================================================================
from oracle import model_ref as mr

TINY = dict(vocab_size=512, hidden_size=128, intermediate_size=352, num_hidden_layers=2, num_attention_heads=4,
            num_key_value_heads=2, hidden_act="silu", layer_norm_eps=1e-5, max_position_embeddings=2048)
TINY_MHA = dict(TINY, num_key_value_heads=4, norm_class="layernorm", attention_bias=True, use_qkv_bias_only=True,
                partial_rotary_factor=0.25)                                  # StableLM-like switches
TINY_GELU = dict(TINY, hidden_act="gelu", num_key_value_heads=1, head_dim=64, normalize_embed=True)   # Gemma-like


def build_ref_model(cfgd, seed=1337):
    torch.manual_seed(seed)
    cfg = ref_config(hm, **cfgd)
    model = hm.HFForCausalLM(cfg).float().eval()
    return model, cfg


def cfg_to_dict(cfgd):
    d = dict(num_linears_per_mlp=3, norm_class="rmsnorm", rope_theta=10000.0, partial_rotary_factor=1.0,
             shared_attention_norm=False, parallel_residual=False, normalize_embed=False, head_dim=None)
    d.update(cfgd)
    return d


def ref_act_range(model, samples):
    """The hook logic of ptq/generate_act_range.py:55-95 attached to the reference model (the script itself parses
    argv and loads datasets at import, so its closure is restated; the model and forward are the reference's)."""
    from functools import partial
    act = {}

    def upd(name, field, t):
        mn, mx = t.min().item(), t.max().item()
        e = act.setdefault(name, {})
        e[field] = [mn, mx] if field not in e else [min(e[field][0], mn), max(e[field][1], mx)]

    def hook(m, xx, yy, name):
        x = xx[0] if isinstance(xx, tuple) else xx
        upd(name, "input", x.detach())
        y = yy[0] if isinstance(yy, tuple) else yy
        upd(name, "output", y.detach())
        if isinstance(m, hm.FMatMul):
            upd(name, "input2", xx[1].detach())

    from transformers.activations import GELUActivation
    hooks = [m.register_forward_hook(partial(hook, name=n)) for n, m in model.named_modules()
             if isinstance(m, (nn.Linear, nn.SiLU, nn.GELU, GELUActivation, nn.LayerNorm, hm.HFRMSNorm, hm.FMatMul))]
    with torch.no_grad():
        for s in samples:
            model(s)
    for h in hooks:
        h.remove()
    return act


def ref_code_trace(qmodel_fwd, sample, layers):
    """Integer codes of the reference's own fake-quant forward (the unmodified Quantizer.forward, qm:251-295), captured by
    forward hooks on the Quantizer submodules of the given decoder layers: code = rne(y / scale) + offset (exact: y is
    (code - offset) * scale).  Also the fp32 residual stream entering the post-attention norm and leaving the block.
    These pin the LSB-flip rate of the integer engine against the reference itself (tests/test_engine_gpu.py)."""
    trace = {i: {} for i in layers}
    hooks = []

    def qhook(i, key):
        def fn(mod, inp, out):
            s, o = mod.scale.detach(), mod.offset.detach()
            c = torch.round(out.detach() / s) + o
            trace[i][key] = c.to(torch.int32).clone()
        return fn

    want = {"input_layernorm.output_quantizer": "x1", "self_attn.q_proj.output_quantizer": "q_proj", "self_attn.k_proj.output_quantizer": "k_proj",
            "self_attn.v_proj.output_quantizer": "v_proj", "self_attn.qk_bmm.input_quantizer": "q", "self_attn.qk_bmm.input2_quantizer": "kT",
            "self_attn.pv_bmm.input2_quantizer": "v", "self_attn.pv_bmm.output_quantizer": "attn",
            "post_attention_layernorm.output_quantizer": "x2", "mlp.w2.input_quantizer": "act"}
    for i in layers:
        layer = qmodel_fwd.model.layers[i]
        for name, key in want.items():
            try:
                m = layer.get_submodule(name)
            except AttributeError:                     # shared_attention_norm: no post_attention_layernorm
                continue
            hooks.append(m.register_forward_hook(qhook(i, key)))
        # residual stream after attention == second operand-free input of resid_add_2 (hm:1257,1270)
        hooks.append(layer.resid_add_2.register_forward_hook(
            lambda mod, inp, out, i=i: trace[i].__setitem__("h_mid", inp[0].detach().clone())))
        hooks.append(layer.register_forward_hook(lambda mod, inp, out, i=i: trace[i].__setitem__("h_out", out[0].detach().clone())))
    with torch.no_grad():
        qmodel_fwd(sample)
    for h in hooks:
        h.remove()
    return trace


def make_args(**kw):
    a = types.SimpleNamespace(nsamples=4, seqlen=32, batch_size=1, epochs=2, warmup_epochs=0, deactive_amp=True,
                              let=True, lwc=True, lrl=True, use_shift=False, aug_loss=False, let_lr=1e-3, lwc_lr=1e-2,
                              lrl_lr=1e-6, let_min_lr=1e-4, lwc_min_lr=1e-3, lrl_min_lr=1e-7, wd=0.0, resume=None,
                              cache_in_gpu=False, original_omniquant=False, dtype=torch.float32, output_dir="/tmp/mq_golden")
    for k, v in kw.items():
        setattr(a, k, v)
    os.makedirs(a.output_dir, exist_ok=True)
    return a


def golden_model(tag, cfgd, w_bits, w_sym, w_pc, mode, T=32, nsamples=2, epochs=1):
    model, cfg = build_ref_model(cfgd)
    g = gen(1337)
    samples = [torch.randint(3, cfgd["vocab_size"], (1, T), generator=g) for _ in range(nsamples)]
    sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    cd = cfg_to_dict(cfgd)

    # ---- act ranges: reference hooks vs oracle
    act = ref_act_range(model, samples)
    act_o = mr.act_range(sd0, cd, samples)
    assert set(act) == set(act_o), (set(act) ^ set(act_o))
    worst = max(abs(act[n][f][i] - act_o[n][f][i]) for n in act for f in act[n] for i in (0, 1))
    print(f"[{tag}] act_range oracle-vs-reference max |diff| = {worst:.3e}")
    assert worst == 0.0

    # ---- fake-quant forward: reference QModel vs oracle
    wq = qm.QuantConfig(bitwidth=w_bits, is_symmetric=w_sym, is_per_channel=w_pc)
    aq = qm.QuantConfig(bitwidth=8)
    qmodel = qm.create_sim_qmodel(model, wq, aq)
    for p in qmodel.parameters():
        p.requires_grad = False
    ref_update_quant_cfg(qm, qmodel)
    qm.set_scale_and_offset(qmodel, act, "parameter")
    qcfg = qm.export_qcfg(qmodel)
    recipe = mr.recipe_from_qcfg_json(qcfg)
    recipe_d = mr.default_recipe(cd, w_bits, w_sym, w_pc, 8)
    assert recipe == recipe_d, "default_recipe != reference export_qcfg"
    with torch.no_grad():
        # weight quantizers are static here: first forward caches min/max (qm:262-277); run on a deepcopy so the
        # original keeps the mobilequant.py ordering (no forward before enable_lwc, SURVEY 8c "ordering trap")
        qm_fwd = copy.deepcopy(qmodel)
        out_ref = qm_fwd(samples[0])
        logits_ref, = (out_ref.logits,)
        ref_trace = ref_code_trace(qm_fwd, samples[0], [0, cd["num_hidden_layers"] - 1])
    qs = mr.QState(recipe, act)
    with torch.no_grad():
        logits_o, hid_o = mr.model_forward(sd0, cd, samples[0], qs, quant=True)
    d = (logits_o - logits_ref).abs().max().item()
    print(f"[{tag}] fake-quant forward oracle-vs-reference max |dlogits| = {d:.3e} (|logits| max {logits_ref.abs().max():.3f})")
    assert d == 0.0, "oracle forward is not bit-exact vs the reference"

    # ---- gradients of every learnable at step 0 (e2e graph over all layers), reference modules driven by hand:
    #      enable_quant (alg:690) -> register LET (alg:692-706) -> smooth_lm_temporary (alg:742-743) -> MSE (alg:745)
    args = make_args(nsamples=nsamples, seqlen=T, epochs=epochs)
    embeds = torch.stack([mr.embed(sd0, cd, s)[0] for s in samples])
    gm = copy.deepcopy(qmodel)
    layers = gm.model.layers
    mask = mr.causal_mask(1, T); pos = torch.arange(T).unsqueeze(0)
    backbone = alg.LayerList(layers)
    alg.disable_quant(gm)
    with torch.no_grad():
        fp_t = backbone(embeds[0:1], attention_mask=mask, position_ids=pos)[0]
    alg.enable_quant(args, gm)
    pairs = {"q_proj": "qkv", "w1": "fc1"}
    if layers[0].self_attn.v_proj.weight.shape[0] == layers[0].self_attn.o_proj.weight.shape[1]:
        pairs["o_proj"] = "out"
    pairs["w2"] = "fc2"
    gg = gen(7)
    for l in layers:
        if l.self_attn.q_proj.weight.shape[0] == l.self_attn.k_proj.weight.shape[0]:
            l.register_parameter("qkt_smooth_scale", nn.Parameter(1 + 0.05 * torch.randn(l.self_attn.q_proj.out_features, generator=gg)))
        for name, module in l.named_modules():
            if isinstance(module, qm.QLinear):
                for key in pairs:
                    if key in name:
                        l.register_parameter(f"{pairs[key]}_smooth_shift", nn.Parameter(torch.zeros(module.in_features)))
                        l.register_parameter(f"{pairs[key]}_smooth_scale", nn.Parameter(1 + 0.05 * torch.randn(module.in_features, generator=gg)))
    let0 = {i: {k: v.detach().clone() for k, v in l.named_parameters() if "smooth" in k} for i, l in enumerate(layers)}
    for l in layers:
        alg.smooth_lm_temporary(l, gm.config, True, False)
    out = backbone(embeds[0:1], attention_mask=mask, position_ids=pos)[0]
    loss0 = torch.nn.functional.mse_loss(fp_t, out)
    loss0.backward()
    grads0 = {i: OrderedDict((k, p.grad.detach().clone()) for k, p in l.named_parameters()
                             if p.grad is not None and "smooth_shift" not in k) for i, l in enumerate(layers)}
    # oracle gradients for the same point
    blocks = [mr.Block(sd0, i, cd) for i in range(cd["num_hidden_layers"])]
    qs_g = mr.QState(recipe, act, learnable=True)
    h = embeds[0:1]
    with torch.no_grad():
        hf = h
        for b in blocks:
            b.sim = True
            hf = b.forward(hf, qs_g, False, mask, pos)
    assert torch.equal(hf, fp_t)
    for b in blocks:
        b.let = OrderedDict((k, v.clone().requires_grad_(True)) for k, v in let0[b.i].items())
        b.lwc = mr.init_lwc(b, recipe)
    for b in blocks:
        h = b.forward(h, qs_g, True, mask, pos)
    lo = torch.nn.functional.mse_loss(h, hf)
    lo.backward()
    worst = 0.0
    for i, b in enumerate(blocks):
        og = {}
        og.update({k: v.grad for k, v in b.let.items() if "smooth_scale" in k})
        og.update({k: v.grad for k, v in b.lwc.items()})
        og.update({k[len(b.p):]: v.grad for k, v in qs_g.p.items() if k.startswith(b.p)})
        assert set(og) == set(grads0[i]), (set(og) ^ set(grads0[i]))
        for k in og:
            den = grads0[i][k].abs().max().item() + 1e-12
            worst = max(worst, (og[k] - grads0[i][k]).abs().max().item() / den)
    print(f"[{tag}] step-0 loss ref {loss0.item():.6e} oracle {lo.item():.6e}; grads oracle-vs-reference max rel diff = {worst:.3e}")
    assert abs(loss0.item() - lo.item()) <= 1e-6 * abs(loss0.item()) and worst < 1e-4
    del gm

    # ---- calibration loop: reference omniquant / e2equant vs oracle (Adam normalises every gradient, so 1-ulp
    #      summation-order differences grow by up to lr per step on parameters whose gradient is ~0: compare after few
    #      steps, LRL tightly, LET/LWC to a few lr)
    loader = [(s, None) for s in samples]
    buf = io.StringIO()
    logger = logging.getLogger(f"gold_{tag}"); logger.setLevel(logging.INFO); logger.handlers = [logging.StreamHandler(buf)]
    if mode == "e2e":
        alg.e2equant(args, qmodel, loader, logger)
        learned = torch.load(os.path.join(args.output_dir, "parameters.pth"), weights_only=False)
    else:
        alg.omniquant(args, qmodel, loader, logger, device="cpu")
        learned = torch.load(os.path.join(args.output_dir, "quant_parameters.pth"), weights_only=False)
    act_after = qm.export_act_range(qmodel)
    # the fused model the reference hands to create_fp_model / save_pretrained (alg:147-184, ptq/mobilequant.py:240-246):
    # LET folded into the weights, every weight clamped to its learned LWC range, zero-shift bias buffers registered
    fused_sd = {k: v.detach().clone() for k, v in qmodel.state_dict().items() if "quantizer" not in k and "smooth" not in k}
    res = mr.calibrate(sd0, cd, recipe, act, embeds, mode=mode, epochs=epochs, let_lr=args.let_lr, lwc_lr=args.lwc_lr,
                       lrl_lr=args.lrl_lr, let_min_lr=args.let_min_lr, lwc_min_lr=args.lwc_min_lr, lrl_min_lr=args.lrl_min_lr)
    worst = dict(let=0.0, lwc=0.0, lrl=0.0)
    for i in learned:
        assert set(learned[i].keys()) == set(res["params"][i].keys()), (sorted(learned[i].keys()), sorted(res["params"][i].keys()))
        for k in learned[i]:
            kind = "let" if "smooth" in k else ("lwc" if "bound_factor" in k else "lrl")
            worst[kind] = max(worst[kind], (learned[i][k].float() - res["params"][i][k]).abs().max().item())
    print(f"[{tag}] {mode} learned params oracle-vs-reference max |diff|: {worst}")
    assert worst["lrl"] < 2e-6 and worst["let"] < 5e-3 and worst["lwc"] < 5e-2
    wa = max(abs(act_after[n][f][j] - res["act_dict"][n][f][j]) for n in act_after for f in act_after[n] for j in (0, 1))
    print(f"[{tag}] {mode} exported act ranges max |diff| = {wa:.3e}")
    assert wa < 1e-3
    torch.save(dict(cfg=cd, state_dict=sd0, samples=samples, act_dict=act, qcfg=qcfg, logits_fq=logits_ref, hidden_fq=hid_o,
                    mode=mode, epochs=epochs, hp=dict(let_lr=args.let_lr, lwc_lr=args.lwc_lr, lrl_lr=args.lrl_lr,
                                                      let_min_lr=args.let_min_lr, lwc_min_lr=args.lwc_min_lr, lrl_min_lr=args.lrl_min_lr),
                    learned=learned, act_after=act_after, fused_state_dict=fused_sd, ref_trace=ref_trace, losses=res["losses"], let0=let0, grads0=grads0, loss0=loss0.item(),
                    w_cfg=dict(bits=w_bits, sym=w_sym, per_channel=w_pc)),
               os.path.join(GOLD, f"model_{tag}.pt"))


def golden_trace(tag, cfgd, T):
    """Forward-only fixture at a longer sequence (several attention tiles, head_dim 64): weights, ranges, qcfg and the
    reference's fake-quant codes of the first and last block."""
    model, cfg = build_ref_model(cfgd)
    g = gen(4242)
    samples = [torch.randint(3, cfgd["vocab_size"], (1, T), generator=g) for _ in range(2)]
    sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    cd = cfg_to_dict(cfgd)
    act = ref_act_range(model, samples)
    qmodel = qm.create_sim_qmodel(model, qm.QuantConfig(bitwidth=8), qm.QuantConfig(bitwidth=8))
    for p in qmodel.parameters():
        p.requires_grad = False
    ref_update_quant_cfg(qm, qmodel)
    qm.set_scale_and_offset(qmodel, act, "parameter")
    qcfg = qm.export_qcfg(qmodel)
    with torch.no_grad():
        logits_ref = qmodel(samples[0]).logits
    ref_trace = ref_code_trace(qmodel, samples[0], [0, cd["num_hidden_layers"] - 1])
    qs = mr.QState(mr.recipe_from_qcfg_json(qcfg), act)
    with torch.no_grad():
        logits_o, _ = mr.model_forward(sd0, cd, samples[0], qs, quant=True)
    d = (logits_o - logits_ref).abs().max().item()
    print(f"[{tag}] fake-quant forward oracle-vs-reference max |dlogits| = {d:.3e}")
    assert d == 0.0
    torch.save(dict(cfg=cd, state_dict=sd0, samples=samples, act_dict=act, qcfg=qcfg, logits_fq=logits_ref, ref_trace=ref_trace,
                    w_cfg=dict(bits=8, sym=False, per_channel=False)), os.path.join(GOLD, f"trace_{tag}.pt"))


TINY_HD64 = dict(TINY, num_attention_heads=2, num_key_value_heads=1)      # head_dim 64: the tcgen05 attention kernel's shape class
# phi-like (scripts/convert_ckpt.py:28): parallel residual, one shared LayerNorm, two-linear GELU MLP, biases, partial rotary
TINY_PHI = dict(TINY, num_key_value_heads=4, norm_class="layernorm", attention_bias=True, mlp_bias=True, hidden_act="gelu",
                num_linears_per_mlp=2, shared_attention_norm=True, parallel_residual=True, partial_rotary_factor=0.5)

if __name__ == "__main__":
    golden_trace("llama_hd64_t256", TINY_HD64, 256)
    golden_trace("phi_t64", TINY_PHI, 64)
    golden_model("llama_w8_e2e", TINY, 8, False, False, "e2e")
    golden_model("llama_w4_omni", TINY, 4, True, True, "omniquant")
    golden_model("stablelm_w8_omni", TINY_MHA, 8, False, False, "omniquant")
    golden_model("gemma_w8_e2e", TINY_GELU, 8, False, False, "e2e")
