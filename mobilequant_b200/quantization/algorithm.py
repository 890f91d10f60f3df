"""MobileQuant calibration loops, GPU-resident (drop-in mirror of mobilellm/quantization/algorithm.py: same public
names, argument meaning, parameter names and checkpoint layout).

Differences in *execution*, not in semantics:
  * LET is not materialised with ATen ops: smooth_lm_temporary attaches a `let` descriptor to each module and the
    weight quantizer's fused kernel applies  W * s[k] (/ or *) r[n]  + min/max + LWC + fake-quant in one pass
    (mq_wprep_fwd) and reduces dL/dW^ straight into the per-channel LET / LWC gradients (mq_wprep_bwd).
  * calibration activations (inps / fp_inps / quant_inps, alg:445-449) stay resident in HBM (180 GB on a B200 holds
    512 x 1024 x 2048 fp32 x 3 = 13 GB); the reference shuttles them CPU<->GPU every step (alg:477,532,573).
  * the reference's layer-sharding over GPUs (utils/parallel_utils.py) is replaced by sample-sharded data parallelism:
    one process per GPU, gradients of the 0.3-1.0 M learnable scalars all-reduced over NCCL each step
    (== the reference run with --batch_size world_size, alg:532-533).
"""
import copy, gc, math, os
from collections import OrderedDict
import torch
import torch.nn as nn
import torch.distributed as dist
from .qmodule import QLinear, QLayerNorm, QRMSNorm, QMatMul, QSiLU, QGELU, MODE_DIV, MODE_MUL
from ..utils.optim import NativeScalerWithGradNormCount, FlatAdamW

CLIPMIN, CLIPMAX = 1e-5, 1e6


class TruncateFunction(torch.autograd.Function):          # alg:27-42
    @staticmethod
    def forward(ctx, input, threshold):
        # same values as the reference's boolean-mask assignment (exact zeros stay zero, alg:31) without the
        # nonzero() host synchronisation, so that a whole training step can be captured in a CUDA graph
        return torch.where(input.abs() < threshold, input.sign() * threshold, input)

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output.clone(), None


def truncate_number(number, threshold=1e-2):
    return TruncateFunction.apply(number, threshold)


# ---------------------------------------------------------------------------------------------------------------
# LET: temporary (per training step) and in-place (fuse) variants, alg:47-144
# ---------------------------------------------------------------------------------------------------------------
def _let(module):
    if getattr(module, "let", None) is None:
        module.let = {}
    return module.let


def _has_bias(m):
    return getattr(m, "bias", None) is not None


def _plain(t):
    """A non-Parameter alias (assigning an nn.Parameter to a module attribute would register it as a parameter)."""
    return t.view_as(t) if isinstance(t, nn.Parameter) else t


def smooth_ln_fcs_temporary(ln, fcs, scales, shifts, use_shift=True):
    """alg:47-68: ln.w / s, ln.b -> (b - shift)/s ; fc.W * s, fc.b + W @ shift."""
    fcs = fcs if isinstance(fcs, list) else [fcs]
    ln.use_temporary_parameter = True
    ln.let = dict(col_fac=scales, col_mode=MODE_DIV)
    if _has_bias(ln):
        ln.temp_bias = (ln.bias - shifts) / scales
    else:
        ln.temp_bias = (-1 * shifts) / scales if use_shift else None
    for fc in fcs:
        fc.use_temporary_parameter = True
        fc.let = dict(col_fac=scales, col_mode=MODE_MUL)
        if use_shift:
            fc.temp_bias = fc.bias + fc.weight @ shifts if _has_bias(fc) else fc.weight @ shifts
        else:                                   # shift == 0: W @ 0 adds an exact zero, skip the GEMV
            fc.temp_bias = _plain(fc.bias) if _has_bias(fc) else None


def smooth_fc_fc_temporary(fc1, fc2, scales, shifts, use_shift=True):
    """alg:71-87: fc1.W' / s[n], fc1.b' -> (b' - shift)/s ; fc2.W * s[k], fc2.b + W @ shift."""
    fc1.use_temporary_parameter = True
    fc2.use_temporary_parameter = True
    l1 = _let(fc1)
    l1.update(row_fac=scales, row_mode=MODE_DIV)
    b1 = getattr(fc1, "temp_bias", None) if hasattr(fc1, "temp_bias") else fc1.bias
    if b1 is not None:
        fc1.temp_bias = (b1 - shifts) / scales.view(-1)
    else:
        fc1.temp_bias = (-1 * shifts) / scales.view(-1) if use_shift else None
    fc2.let = dict(col_fac=scales, col_mode=MODE_MUL)
    if use_shift:
        fc2.temp_bias = fc2.bias + fc2.weight @ shifts if _has_bias(fc2) else fc2.weight @ shifts
    else:
        fc2.temp_bias = _plain(fc2.bias) if _has_bias(fc2) else None


def smooth_q_k_temporary(q_proj, k_proj, scales):
    """alg:90-96: q / s[n], k * s[n] (weights and biases)."""
    q_proj.use_temporary_parameter = True
    k_proj.use_temporary_parameter = True
    _let(q_proj).update(row_fac=scales, row_mode=MODE_DIV)
    _let(k_proj).update(row_fac=scales, row_mode=MODE_MUL)
    if getattr(q_proj, "temp_bias", None) is not None:
        q_proj.temp_bias = q_proj.temp_bias / scales.view(-1)
    if getattr(k_proj, "temp_bias", None) is not None:
        k_proj.temp_bias = k_proj.temp_bias * scales.view(-1)


def _set_bias(m, value):
    if _has_bias(m):
        m.bias.copy_(value)
    else:
        if hasattr(m, "bias"):
            del m.bias
        m.register_buffer("bias", value)


def smooth_ln_fcs_inplace(ln, fcs, scales, shifts):       # alg:99-120
    fcs = fcs if isinstance(fcs, list) else [fcs]
    ln.use_temporary_parameter = False
    _set_bias(ln, ((ln.bias - shifts) / scales) if _has_bias(ln) else (-1 * shifts) / scales)
    ln.weight.div_(scales)
    for fc in fcs:
        fc.use_temporary_parameter = False
        _set_bias(fc, (fc.bias + fc.weight @ shifts) if _has_bias(fc) else fc.weight @ shifts)
        fc.weight.mul_(scales.view(1, -1))


def smooth_fc_fc_inplace(fc1, fc2, scales, shifts):       # alg:123-135
    fc1.use_temporary_parameter = False
    fc2.use_temporary_parameter = False
    fc1.bias.sub_(shifts)
    fc1.bias.div_(scales.view(-1))
    fc1.weight.div_(scales.view(-1, 1))
    _set_bias(fc2, (fc2.bias + fc2.weight @ shifts) if _has_bias(fc2) else fc2.weight @ shifts)
    fc2.weight.mul_(scales.view(1, -1))


def smooth_q_k_inplace(q_proj, k_proj, scales):           # alg:138-144
    q_proj.use_temporary_parameter = False
    k_proj.use_temporary_parameter = False
    q_proj.weight.div_(scales.view(-1, 1))
    q_proj.bias.div_(scales.view(-1))
    k_proj.weight.mul_(scales.view(-1, 1))
    k_proj.bias.mul_(scales.view(-1))


def _truncate_let(model, use_shift):
    template = "smooth" if use_shift else "smooth_scale"
    with torch.no_grad():
        for name, p in model.named_parameters():
            if template in name:
                p.copy_(truncate_number(p))               # alg:190-193 (in place: graph replays keep the storage)


def _let_plan(model, config, original_omniquant):
    """Which (ln, fcs) / (fc1, fc2) / (q, k) pairs exist for this block -- alg:195-220."""
    at, mlp = model.self_attn, model.mlp
    plan = dict(ln=[], fcfc=[], qk=None)
    three = config.num_linears_per_mlp == 3
    if config.shared_attention_norm:
        plan["ln"].append((model.input_layernorm, [at.q_proj, at.k_proj, at.v_proj, mlp.w1] + ([mlp.w3] if three else []), "qkv"))
    else:
        plan["ln"].append((model.input_layernorm, [at.q_proj, at.k_proj, at.v_proj], "qkv"))
        plan["ln"].append((model.post_attention_layernorm, [mlp.w1] + ([mlp.w3] if three else []), "fc1"))
    if at.v_proj.weight.shape[0] == at.o_proj.weight.shape[1]:
        plan["fcfc"].append((at.v_proj, at.o_proj, "out"))
    if three and not original_omniquant:
        plan["fcfc"].append((mlp.w3, mlp.w2, "fc2"))
    if at.q_proj.weight.shape[0] == at.k_proj.weight.shape[0]:
        plan["qk"] = (at.q_proj, at.k_proj)
    return plan


def smooth_lm_temporary(model, config, use_let, use_shift=False, original_omniquant=False):
    """alg:187-233 for one decoder block."""
    for m in model.modules():
        if isinstance(m, (QLinear, QRMSNorm, QLayerNorm)):
            m.let = None
            if hasattr(m, "temp_bias"):
                del m.temp_bias
    if use_let:
        _truncate_let(model, use_shift)
        plan = _let_plan(model, config, original_omniquant)
        for ln, fcs, key in plan["ln"]:
            smooth_ln_fcs_temporary(ln, fcs, getattr(model, f"{key}_smooth_scale"), getattr(model, f"{key}_smooth_shift"), use_shift)
        for fc1, fc2, key in plan["fcfc"]:
            smooth_fc_fc_temporary(fc1, fc2, getattr(model, f"{key}_smooth_scale"), getattr(model, f"{key}_smooth_shift"), use_shift)
        if plan["qk"] is not None:
            smooth_q_k_temporary(plan["qk"][0], plan["qk"][1], model.qkt_smooth_scale)
    for m in model.modules():
        if isinstance(m, QLinear):
            m.use_temporary_parameter = True
            if not hasattr(m, "temp_bias"):
                m.temp_bias = _plain(m.bias)


@torch.no_grad()
def smooth_lm_inplace(model, config, use_let, use_shift=False, original_omniquant=False):
    """alg:147-184: fuse LET into the weights and clamp every weight to its learned LWC range."""
    if use_let:
        _truncate_let(model, use_shift)
        plan = _let_plan(model, config, original_omniquant)
        for ln, fcs, key in plan["ln"]:
            smooth_ln_fcs_inplace(ln, fcs, getattr(model, f"{key}_smooth_scale"), getattr(model, f"{key}_smooth_shift"))
        for fc1, fc2, key in plan["fcfc"]:
            smooth_fc_fc_inplace(fc1, fc2, getattr(model, f"{key}_smooth_scale"), getattr(model, f"{key}_smooth_shift"))
        if plan["qk"] is not None:
            smooth_q_k_inplace(plan["qk"][0], plan["qk"][1], model.qkt_smooth_scale)
    for m in model.modules():
        if isinstance(m, (QLinear, QRMSNorm, QLayerNorm)):
            m.weight.data = m.weight_quantizer.run_lwc(m.weight)
            m.use_temporary_parameter = False
            m.let = None


# ---------------------------------------------------------------------------------------------------------------
# parameter groups -- alg:239-292
# ---------------------------------------------------------------------------------------------------------------
def let_parameters(model, use_shift=False):
    template = "smooth" if use_shift else "smooth_scale"
    return iter([p for n, p in model.named_parameters() if n.find(template) > -1])


def lwc_parameters(model):
    return iter([p for n, p in model.named_parameters() if n.find("bound_factor") > -1])


def lrl_parameters(model):
    out = []
    for n, p in model.named_parameters():
        if n.find("quantizer.offset") > -1:
            out.append(p)
        if n.find("quantizer.scale") > -1:
            out.append(p)
    return iter(out)


def _is_quant_param(n, template):
    return n.find("bound_factor") > -1 or n.find(template) > -1 or n.find("quantizer.offset") > -1 or n.find("quantizer.scale") > -1


def get_parameters(model, use_shift=False):
    template = "smooth" if use_shift else "smooth_scale"
    return iter([p for n, p in model.named_parameters() if _is_quant_param(n, template)])


def quant_state_dict(model, destination=None, prefix="", keep_vars=False, use_shift=False):
    destination = OrderedDict() if destination is None else destination
    template = "smooth" if use_shift else "smooth_scale"
    for n, p in model.named_parameters():
        if _is_quant_param(n, template):
            # (learnables are views into the optimiser's flat buffer: clone so that a saved checkpoint holds small tensors)
            destination[prefix + n] = p if keep_vars else p.detach().clone()
    return destination


def clear_temp_variable(model):
    for m in model.modules():
        if isinstance(m, (QLinear, QLayerNorm, QRMSNorm)):
            for a in ("temp_weight", "temp_bias"):
                if hasattr(m, a):
                    delattr(m, a)
            m.let = None


def _release_weight_buffers(model):
    """Drop the calibration-time scratch the Q* modules hold on to: the shared buffers sibling projections write their
    fake-quantised weights into (qmodule.py:_grouped_weight; one fp32 copy of every grouped weight) and prefetched weights."""
    for m in model.modules():
        for a in ("_wout", "_wgroup", "_prepared_weight", "_w_side"):
            if hasattr(m, a):
                delattr(m, a)


def get_lr(max_lr, min_lr, it, warmup_iters, max_iters):
    """alg:296-307: linear warm-up then cosine decay."""
    if it < warmup_iters:
        return max_lr * it / warmup_iters
    if it > max_iters:
        return min_lr
    decay_ratio = (it - warmup_iters) / (max_iters - warmup_iters)
    assert 0 <= decay_ratio <= 1
    coeff = 0.5 * (1.0 + math.cos(math.pi * decay_ratio))
    return min_lr + coeff * (max_lr - min_lr)


class LayerList(nn.Module):                               # alg:313-322
    def __init__(self, layers):
        super().__init__()
        self.layers = layers

    def forward(self, hidden_states, attention_mask=None, position_ids=None):
        for i in range(len(self.layers)):
            out = self.layers[i](hidden_states, attention_mask, position_ids)
            hidden_states = out[0]
        return out


def _each_q(module):
    for slot in ("input", "input2", "weight", "output"):
        q = getattr(module, f"{slot}_quantizer", None)
        if q is not None:
            yield slot, q


def enable_quant(args, model):
    """alg:325-351."""
    for m in model.modules():
        if isinstance(m, (QLinear, QRMSNorm, QLayerNorm, QMatMul, QSiLU, QGELU)):
            for slot, q in _each_q(m):
                q.enable = True
                if slot == "weight" and args.lwc:
                    q.enable_lwc(m.weight)
    return model


def disable_quant(model):
    """alg:354-378."""
    for m in model.modules():
        if isinstance(m, (QLinear, QRMSNorm, QLayerNorm, QMatMul, QSiLU, QGELU)):
            for slot, q in _each_q(m):
                q.enable = False
                if slot == "weight":
                    q.disable_lwc()
    return model


# ---------------------------------------------------------------------------------------------------------------
# shared pieces of the two loops
# ---------------------------------------------------------------------------------------------------------------
def _dp():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def _catch_first_layer_inputs(model, layers, dataloader, nsamples, seqlen, device, dtype):
    """alg:409-441 / 617-645: run the embedding, stop at layer 0, keep its inputs + mask + position ids."""
    inps = torch.zeros((nsamples, seqlen, model.config.hidden_size), dtype=dtype, device=device)
    cache = {"i": 0}

    class Catcher(nn.Module):
        def __init__(self, module):
            super().__init__()
            self.module = module

        def forward(self, inp, **kwargs):
            inps[cache["i"]] = inp
            cache["i"] += 1
            cache["attention_mask"] = kwargs["attention_mask"]
            cache["position_ids"] = kwargs["position_ids"]
            raise ValueError

    layers[0] = Catcher(layers[0])
    with torch.no_grad():
        for batch in dataloader:
            if cache["i"] >= nsamples:
                break
            try:
                model(batch[0].to(device))
            except ValueError:
                pass
    layers[0] = layers[0].module
    return inps, cache.get("attention_mask"), cache.get("position_ids")


def _register_let(layer, pairs, device, dtype):
    """alg:484-496 / 692-706."""
    if layer.self_attn.q_proj.weight.shape[0] == layer.self_attn.k_proj.weight.shape[0]:
        layer.register_parameter("qkt_smooth_scale", nn.Parameter(torch.ones(layer.self_attn.q_proj.out_features, device=device, dtype=dtype)))
    for name, module in layer.named_modules():
        if isinstance(module, QLinear):
            for key in pairs:
                if key in name:
                    layer.register_parameter(f"{pairs[key]}_smooth_shift", nn.Parameter(torch.zeros(module.in_features, device=device, dtype=dtype)))
                    layer.register_parameter(f"{pairs[key]}_smooth_scale", nn.Parameter(torch.ones(module.in_features, device=device, dtype=dtype)))


def _drop_learned(layer):
    for name in [n for n, _ in layer.named_parameters() if n.find("bound_factor") > -1 or n.find("smooth") > -1]:
        obj = layer
        parts = name.split(".")
        for p in parts[:-1]:
            obj = getattr(obj, p)
        delattr(obj, parts[-1])


def _load_learned(layer, state, use_shift):
    """--resume (alg:497-498, 707-709).  The checkpoints hold only the learnable quantisation parameters (quant_state_dict),
    so the load cannot be strict over the whole block; instead every key of the checkpoint must exist in the block and
    every learnable the block registered must be in the checkpoint -- a silent partial resume is an error."""
    res = layer.load_state_dict(state, strict=False)
    if res.unexpected_keys:
        raise KeyError(f"--resume: unexpected keys {res.unexpected_keys[:4]}...")
    want = set(quant_state_dict(layer, use_shift=use_shift).keys())
    missing = sorted(want - set(state.keys()))
    if missing:
        raise KeyError(f"--resume: the checkpoint lacks learnable parameters {missing[:4]}...")


def _allreduce_grads(params, world):
    """Data-parallel exchange: SUM of the learnable-scalar gradients over NCCL, then the batch mean (alg:459)."""
    from ..utils.dist import allreduce_grads
    allreduce_grads(params, world)


def _train_step(args, loss, optimizer, loss_scaler, params_fn, world):
    optimizer.zero_grad()
    if world > 1:
        loss.backward()
        _allreduce_grads(list(params_fn()), world)
        return loss_scaler.step_only(optimizer, parameters=params_fn())
    return loss_scaler(loss, optimizer, parameters=params_fn())


def _use_graphs(device):
    return device.type == "cuda" and os.environ.get("MQ_CUDA_GRAPH", "1") != "0"


def _weight_stream(device):
    """Side stream(s) of the weight pass (MQ_WPREP_STREAM=0 keeps everything on one stream).  MQ_WPREP_STREAMS=n deals the weights
    of a block round-robin onto n streams; measured on B200 (256 samples: 11.4 s with 1, 11.5-11.9 s with 2-4) the machine is
    already full with one, so one is the default."""
    if device.type != "cuda" or os.environ.get("MQ_WPREP_STREAM", "1") == "0":
        return None
    n = max(1, int(os.environ.get("MQ_WPREP_STREAMS", "1")))
    return [torch.cuda.Stream(device=device) for _ in range(n)]


def _prefetch_weights(layers, side):
    """Run the LET + LWC + fake-quant pass of every weight of `layers` (mq_wprep_fwd: min/max, clip, quantise -- 8 bytes of
    HBM traffic per weight, independent of the activations) on `side`, ahead of the forward pass that consumes them on the
    current stream; each module picks its tensor up after waiting for its own event (qmodule.py:_fq_weight), so layer k's
    GEMMs overlap the weight pass of the layers behind it.  autograd runs each backward on the stream of its forward, so
    the gradient reductions of the weight pass (mq_wprep_bwd) overlap the activation backward the same way."""
    if side is None:
        return
    main = torch.cuda.current_stream()
    for s in side:
        s.wait_stream(main)                         # LET / LWC parameters of this step are final
    k = 0
    for layer in layers:
        for m in layer.modules():
            if isinstance(m, (QLinear, QRMSNorm, QLayerNorm)):
                s = side[k % len(side)]
                k += 1
                with torch.cuda.stream(s):
                    w = m._fq_weight_now()
                    ev = torch.cuda.Event()
                    ev.record(s)
                m._prepared_weight = (w, ev)


def _make_optimizer(groups, wd, device):
    """AdamW of the reference (alg:513,716-722).  On the GPU: utils/optim.py:FlatAdamW -- every learnable becomes a view of
    one flat buffer and grad norm + skip-on-non-finite (GradScaler semantics, optim.py:37-38) + AdamW are one fused kernel
    pair (mq_adamw_step) with the learning rates and the step counter on the device, which is also what lets a step be
    replayed as a CUDA graph.  (The library optimiser spent 38 ms of a 94 ms TinyLlama step walking ~1500 tiny tensors.)"""
    groups = [{"params": list(g["params"]), "lr": g["lr"]} for g in groups]
    if device.type == "cuda":
        return FlatAdamW(groups, weight_decay=wd, device=device)
    return torch.optim.AdamW(groups, weight_decay=wd)


def _set_lr(optimizer, idx, value):
    optimizer.param_groups[idx]["lr"] = value


class _Replay:
    """fn(*static_inputs) -> tuple of tensors, executed eagerly the first time (warm-up: cuBLAS workspaces, autograd,
    optimizer state) and captured into a CUDA graph on the second call; later calls copy the inputs into the static
    buffers and replay.  The reference issues ~11 k kernel launches per e2e step from Python; a replay
    is one launch, which is what makes the GPU -- not the interpreter -- the bound of the calibration loop."""

    def __init__(self, fn, example_inputs, enabled, before_capture=None):
        self.fn, self.enabled, self.before_capture = fn, enabled, before_capture
        self.static_in = [torch.empty_like(t) for t in example_inputs] if enabled else None
        self.graph, self.static_out, self.calls = None, None, 0
        self.stream = torch.cuda.Stream() if enabled else None

    def __call__(self, *inputs):
        if not self.enabled:
            return self.fn(*inputs)
        for dst, src in zip(self.static_in, inputs):
            dst.copy_(src)
        self.calls += 1
        if self.calls == 1:                         # eager warm-up on the side stream the capture will use
            self.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                out = self.fn(*self.static_in)
            torch.cuda.current_stream().wait_stream(self.stream)
            return out
        if self.graph is None:
            if self.before_capture is not None:
                self.before_capture()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=self.stream):
                self.static_out = self.fn(*self.static_in)
        self.graph.replay()
        return self.static_out


_FP_PASS_SAMPLES = 8          # samples per forward of the e2e FP-target pass


def _eager_no_grad(fn, x):
    with torch.no_grad():
        return fn(x)


def _no_grad_replay(fn, example, device):
    """Forward-only replay unit (FP-target and quant-input refresh passes, alg:471-479,567-573,674-688)."""
    def body(x):
        with torch.no_grad():
            return fn(x)
    return _Replay(body, [example], _use_graphs(device))


def _make_step(args, forward_fn, loss_func, optimizer, loss_scaler, params_fn, world, device, example_inputs, side=None):
    """One optimiser step as a replayable unit: forward, MSE against the FP target(s), backward, (all-reduce,) global
    grad norm, AdamW.  Returns step(x, y[, y2]) -> (loss, norm), both detached 0-d tensors."""
    flat = isinstance(optimizer, FlatAdamW)
    graphed = flat and _use_graphs(device)

    def body(x, *targets):
        if flat:
            optimizer.zero_grad()                    # one memset of the flat gradient buffer
        out = forward_fn(x)
        loss = loss_func(targets[0], out)
        for t in targets[1:]:
            loss = loss + loss_func(t, out)
        if not flat:
            return loss.detach(), _train_step(args, loss, optimizer, loss_scaler, params_fn, world).detach()
        loss.backward()
        for s in side or ():
            torch.cuda.current_stream().wait_stream(s)        # the weight pass's gradient reductions ran on the side streams
        optimizer.allreduce_grads(world)             # one SUM all-reduce of the flat buffer (no-op on one rank)
        norm = optimizer.step()                      # grad norm, skip-on-non-finite and AdamW on the device
        return loss.detach(), norm

    runner = _Replay(body, example_inputs, graphed)

    def step(*inputs):
        if flat:
            optimizer.sync_lr()                      # this step's learning rates -> device (outside the graph)
        loss, norm = runner(*inputs)
        return loss.clone(), norm.clone()
    return step


# ---------------------------------------------------------------------------------------------------------------
def omniquant(args, model, dataloader, logger, device=None):
    """Block-wise calibration, alg:381-584."""
    logger.info("Starting ...")
    use_cache = model.config.use_cache
    model.config.use_cache = False
    if device is None:
        device = next(model.parameters()).device
    device = torch.device(device)
    rank, world = _dp()
    layers = model.model.layers
    model.model.embed_tokens = model.model.embed_tokens.to(device)
    model.model.norm = model.model.norm.to(device)
    pairs = {"q_proj": "qkv", "w1": "fc1"}
    if layers[0].self_attn.v_proj.weight.shape[0] == layers[0].self_attn.o_proj.weight.shape[1]:
        pairs["o_proj"] = "out"
    if model.config.num_linears_per_mlp == 3 and not args.original_omniquant:
        pairs["w2"] = "fc2"
    layers[0] = layers[0].to(device)
    if args.epochs > 0 and not args.deactive_amp:
        raise NotImplementedError("the B200 path calibrates in fp32 (the reference's W8A8/W4A8 recipes all set "
                                  "--deactive_amp, ptq/mobilequant.py:122-123)")
    dtype = torch.float32
    inps, attention_mask, position_ids = _catch_first_layer_inputs(model, layers, dataloader, args.nsamples, args.seqlen, device, dtype)
    quant_inps = inps
    fp_inps = inps.clone()
    fp_inps_2 = inps.clone() if args.aug_loss else None
    attention_mask_batch = attention_mask.repeat(args.batch_size, 1, 1, 1) if attention_mask is not None else None
    loss_func = torch.nn.MSELoss()
    omni_parameters = torch.load(args.resume) if args.resume else {}
    steps_per_epoch = args.nsamples // args.batch_size
    gsteps = steps_per_epoch // world                       # global optimiser steps per epoch (world micro-batches each)
    my_batches = [g * world + rank for g in range(gsteps)]  # micro-batches owned by this rank
    my_samples = [k for j in my_batches for k in range(j * args.batch_size, (j + 1) * args.batch_size)] if world > 1 \
        else list(range(args.nsamples))

    for i in range(len(layers)):
        logger.info(f"=== Start quantize layer {i} ===")
        qlayer = layers[i].to(device)
        disable_quant(qlayer)
        if args.epochs > 0:
            fp_pass = _no_grad_replay(lambda x: qlayer(x, attention_mask=attention_mask, position_ids=position_ids)[0], fp_inps[:1], device)
            for j in my_samples:
                fp_inps[j] = fp_pass(fp_inps[j].unsqueeze(0))[0]
                if args.aug_loss:
                    fp_inps_2[j] = fp_pass(quant_inps[j].unsqueeze(0))[0]
            del fp_pass
        enable_quant(args, qlayer)
        if args.let:
            _register_let(qlayer, pairs, device, dtype)
        if args.resume:
            _load_learned(qlayer, omni_parameters[i], args.use_shift)
        if args.epochs > 0:
            groups = [{"params": let_parameters(qlayer, args.use_shift), "lr": args.let_lr},
                      {"params": lwc_parameters(qlayer), "lr": args.lwc_lr}]
            if args.lrl:
                groups.append({"params": lrl_parameters(qlayer), "lr": args.lrl_lr})
            optimizer = _make_optimizer(groups, args.wd, device)
            loss_scaler = NativeScalerWithGradNormCount()
            max_iters = args.epochs * gsteps
            warmup_iters = args.warmup_epochs * gsteps

            side = _weight_stream(device)

            def forward_fn(x, qlayer=qlayer, side=side):
                smooth_lm_temporary(qlayer, model.config, args.let, args.use_shift, args.original_omniquant)
                _prefetch_weights([qlayer], side)
                return qlayer(x, attention_mask=attention_mask_batch, position_ids=position_ids)[0]

            bs = args.batch_size
            example = [quant_inps[:bs], fp_inps[:bs]] + ([fp_inps_2[:bs]] if args.aug_loss else [])
            step = _make_step(args, forward_fn, loss_func, optimizer, loss_scaler, lambda qlayer=qlayer: get_parameters(qlayer, args.use_shift),
                              world, device, example, side)
            for epochs in range(args.epochs):
                loss_list, norm_list = [], []
                for g, j in enumerate(my_batches):
                    index = j * args.batch_size
                    it = epochs * gsteps + g
                    _set_lr(optimizer, 0, get_lr(args.let_lr, args.let_min_lr, it, warmup_iters, max_iters))
                    _set_lr(optimizer, 1, get_lr(args.lwc_lr, args.lwc_min_lr, it, warmup_iters, max_iters))
                    if args.lrl:
                        _set_lr(optimizer, 2, get_lr(args.lrl_lr, args.lrl_min_lr, it, warmup_iters, max_iters))
                    batch = [quant_inps[index:index + bs], fp_inps[index:index + bs]] + ([fp_inps_2[index:index + bs]] if args.aug_loss else [])
                    loss, norm = step(*batch)
                    loss_list.append(loss)
                    norm_list.append(norm)
                loss_mean = torch.stack(loss_list).mean().item()
                if not math.isfinite(loss_mean):
                    raise FloatingPointError(f"layer {i} epoch {epochs}: loss is not finite")   # reference: pdb (alg:536-538)
                norm_mean = torch.stack(norm_list).mean().item()
                logger.info(f"layer {i} iter {epochs} loss:{loss_mean} norm:{norm_mean} max memory_allocated {torch.cuda.max_memory_allocated(device) / 1024**2} ")
            del step
            clear_temp_variable(qlayer)
            _release_weight_buffers(qlayer)
            del optimizer
        if args.epochs > 0:
            omni_parameters[i] = quant_state_dict(qlayer)
            if rank == 0:
                torch.save(omni_parameters, os.path.join(args.output_dir, "quant_parameters.pth"))
        smooth_lm_inplace(qlayer, model.config, args.let, args.use_shift, args.original_omniquant)
        _drop_learned(qlayer)
        if args.epochs > 0:
            refresh = _no_grad_replay(lambda x: qlayer(x, attention_mask=attention_mask, position_ids=position_ids)[0], quant_inps[:1], device)
            for j in my_samples:
                quant_inps[j] = refresh(quant_inps[j].unsqueeze(0))[0]
            del refresh
        layers[i] = qlayer
    del inps, quant_inps, fp_inps, fp_inps_2
    gc.collect()
    model.config.use_cache = use_cache
    return model


def e2equant(args, model, dataloader, logger, device=None):
    """End-to-end calibration, alg:587-787."""
    logger.info("Starting ...")
    use_cache = model.config.use_cache
    model.config.use_cache = False
    rank, world = _dp()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    model.to(device)                                       # replaces map_layers_to_multi_gpus (alg:593)
    layers = model.model.layers
    pairs = {"q_proj": "qkv", "w1": "fc1"}
    if layers[0].self_attn.v_proj.weight.shape[0] == layers[0].self_attn.o_proj.weight.shape[1]:
        pairs["o_proj"] = "out"
    if model.config.num_linears_per_mlp == 3:
        pairs["w2"] = "fc2"
    if args.epochs > 0 and not args.deactive_amp:
        raise NotImplementedError("the B200 path calibrates in fp32 (--deactive_amp)")
    dtype = torch.float32
    inps, attention_mask, position_ids = _catch_first_layer_inputs(model, layers, dataloader, args.nsamples, args.seqlen, device, dtype)
    quant_inps = inps
    fp_inps = inps.clone()
    fp_inps_2 = inps.clone() if args.aug_loss else None
    attention_mask_batch = attention_mask.repeat(args.batch_size, 1, 1, 1) if attention_mask is not None else None
    loss_func = torch.nn.MSELoss()
    e2e_parameters = torch.load(args.resume) if args.resume else {}
    batch_size = args.batch_size
    steps_per_epoch = args.nsamples // batch_size
    gsteps = steps_per_epoch // world
    my_batches = [g * world + rank for g in range(gsteps)]
    disable_quant(model)
    backbone = LayerList(layers)

    if args.epochs > 0:
        # FP targets of this rank's shard only.  The float model is the same function of every sample, so the targets are
        # computed several samples per forward (GEMMs at M = 8 * seqlen instead of seqlen); a ragged tail runs eagerly.
        mine = [j * batch_size + t for j in my_batches for t in range(batch_size)]
        chunk = max(batch_size, min(_FP_PASS_SAMPLES, len(mine)))
        mask_chunk = attention_mask.repeat(chunk, 1, 1, 1) if attention_mask is not None else None
        run = lambda x, m=mask_chunk: backbone(x, attention_mask=m, position_ids=position_ids)[0]
        fp_pass = _no_grad_replay(run, fp_inps[:chunk], device)
        for c0 in range(0, len(mine), chunk):
            idx = torch.tensor(mine[c0:c0 + chunk], device=device)
            if len(idx) == chunk:
                fn = fp_pass
            else:
                tail_mask = attention_mask.repeat(len(idx), 1, 1, 1) if attention_mask is not None else None
                fn = lambda x, m=tail_mask: _eager_no_grad(lambda z: run(z, m), x)
            fp_inps[idx] = fn(fp_inps[idx])                  # (a replay returns its static output buffer: store before the next call)
            if args.aug_loss:
                fp_inps_2[idx] = fn(quant_inps[idx])
        del fp_pass
    enable_quant(args, model)
    if args.let:
        for i in range(len(layers)):
            _register_let(layers[i], pairs, device, dtype)
            if args.resume:
                _load_learned(layers[i], e2e_parameters[i], args.use_shift)

    optimizer = None
    if args.epochs > 0:
        optimizer = _make_optimizer([{"params": list(let_parameters(model, args.use_shift)), "lr": args.let_lr},
                                     {"params": list(lwc_parameters(model)), "lr": args.lwc_lr},
                                     {"params": list(lrl_parameters(model)), "lr": args.lrl_lr}], args.wd, device)
        loss_scaler = NativeScalerWithGradNormCount()
        max_iters = args.epochs * gsteps
        warmup_iters = args.warmup_epochs * gsteps

        side = _weight_stream(device)

        def forward_fn(x):
            for k in range(len(layers)):
                smooth_lm_temporary(layers[k], model.config, args.let, args.use_shift)
            _prefetch_weights(layers, side)
            return backbone(x, attention_mask=attention_mask_batch, position_ids=position_ids)[0]

        example = [quant_inps[:batch_size], fp_inps[:batch_size]] + ([fp_inps_2[:batch_size]] if args.aug_loss else [])
        step = _make_step(args, forward_fn, loss_func, optimizer, loss_scaler, lambda: get_parameters(model, args.use_shift), world, device, example,
                          side)
        for epochs in range(args.epochs):
            loss_list, norm_list = [], []
            for g, j in enumerate(my_batches):
                index = j * batch_size
                it = epochs * gsteps + g
                _set_lr(optimizer, 0, get_lr(args.let_lr, args.let_min_lr, it, warmup_iters, max_iters))
                _set_lr(optimizer, 1, get_lr(args.lwc_lr, args.lwc_min_lr, it, warmup_iters, max_iters))
                _set_lr(optimizer, 2, get_lr(args.lrl_lr, args.lrl_min_lr, it, warmup_iters, max_iters))
                batch = [quant_inps[index:index + batch_size], fp_inps[index:index + batch_size]] + \
                        ([fp_inps_2[index:index + batch_size]] if args.aug_loss else [])
                loss, norm = step(*batch)
                loss_list.append(loss)
                norm_list.append(norm)
            loss_mean = torch.stack(loss_list).mean().item()
            if not math.isfinite(loss_mean):
                raise FloatingPointError(f"epoch {epochs}: loss is not finite")
            norm_mean = torch.stack(norm_list).mean().item()
            logger.info(f"Epoch {epochs} loss:{loss_mean} norm:{norm_mean} max memory_allocated {torch.cuda.max_memory_allocated(device) / 1024**2} ")
            for k in range(len(layers)):
                e2e_parameters[k] = quant_state_dict(layers[k])
            if rank == 0:
                torch.save(e2e_parameters, os.path.join(args.output_dir, "parameters.pth"))

    for i in range(len(layers)):
        e2e_parameters[i] = OrderedDict((k, v.clone()) for k, v in quant_state_dict(layers[i]).items())
        smooth_lm_inplace(layers[i], model.config, args.let, args.use_shift)
        _drop_learned(layers[i])
        _release_weight_buffers(layers[i])
    if rank == 0:
        torch.save(e2e_parameters, os.path.join(args.output_dir, "parameters.pth"))
    step = None
    del optimizer, inps, quant_inps, fp_inps, fp_inps_2
    gc.collect()
    model.config.use_cache = use_cache
    return model
