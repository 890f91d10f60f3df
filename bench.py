#!/usr/bin/env python
"""Headline benchmark of the MobileQuant hot path on B200 (contract: task statement / BASELINE.json).

Workload (config.workload): statically-quantised W8A8 integer forward of TinyLlama-1.1B shapes, batch x seq 1024
synthetic prompts, random-init weights (no checkpoints offline): embeddings -> 22 integer decoder blocks (IntEngine)
-> final norm -> lm_head over all positions -> arg-max of the last position.  metric = int8 tokens/s.

  value  : device-timed (CUDA events), token ids already resident in HBM
  e2e    : same through the public API with pinned-host token ids H2D and the next-token ids D2H every step
  calib  : MobileQuant e2e calibration (LET+LWC+LRL, module path) samples/s on the same model (second half of the metric)
  roofline / cpu_baseline : see DESIGN.md "Measurement"

`--impl reference` times the reference's own CPU implementation of this forward (the fp32 fake-quant simulation,
eval/harness_eval.py --mode custom recipe) as restated in oracle/model_ref.py, on the host cores.
"""
import argparse, json, os, subprocess, sys, threading, time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODEL = "tinyllama-1.1b"
SEQ = 1024


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--batch", type=int, default=8, help="sequences per GPU per step")
    p.add_argument("--seqlen", type=int, default=SEQ)
    p.add_argument("--model", default=MODEL, help="tinyllama-1.1b | stablelm-2-1.6b | gemma-2b (synthetic weights of that shape)")
    p.add_argument("--wbits", type=int, default=8, choices=[4, 8],
                   help="8: W8A8 per-tensor asymmetric weights (headline); 4: W4A8 per-channel symmetric (BASELINE config 3)")
    p.add_argument("--config", type=int, default=None, choices=[2, 3, 4, 5],
                   help="preset of BASELINE.json:configs -- 2: the default headline run; 3: TinyLlama W4A8 per-channel, batch 32; "
                        "4: stablelm-2-1.6b (calibration leg sharded over the ranks); 5: gemma-2b, seq 2048")
    p.add_argument("--calib-samples", type=int, default=512,
                   help="calibration samples of the calib leg (BASELINE config 2: 512), sharded over the ranks")
    p.add_argument("--layers", type=int, default=None, help="debug: override num_hidden_layers")
    p.add_argument("--no-calib", action="store_true")
    p.add_argument("--no-decode", action="store_true")
    p.add_argument("--decode-steps", type=int, default=64, help="decode tokens per sequence in the decode leg")
    p.add_argument("--decode-batch", type=int, default=None, help="sequences in the decode leg (default: --batch)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--cpu-seq", type=int, default=1, help="sequences in the bounded CPU sample")
    p.add_argument("--profile-step", action="store_true",
                   help="ncu helper: run the warm-up, then ONE step between cudaProfilerStart/Stop and exit (use with "
                        "`ncu --profile-from-start off`); prints no bench line")
    a = p.parse_args()
    if a.config == 3:
        a.wbits, a.batch = 4, 32
    elif a.config == 4:
        a.model = "stablelm-2-1.6b"
    elif a.config == 5:
        a.model, a.seqlen = "gemma-2b", 2048
    return a


# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None, "reasons": reasons,
                "samples": len(sm)}


def model_cfg(args):
    from mobilequant_b200.model.hf_config import named_config
    over = {}
    if args.layers:
        over["num_hidden_layers"] = args.layers
    return named_config(args.model, **over)


def synth_ids(n, T, vocab, seed):
    import torch
    g = torch.Generator().manual_seed(seed)
    return torch.randint(3, vocab, (n, T), generator=g)       # reference's random-id convention, device/export.py:116


# ---------------------------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's CPU path for this forward (fp32 fake-quant simulation), oracle port, all host threads."""
    import torch
    from oracle import model_ref as mr
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    cfg = model_cfg(args)
    cd = {k: getattr(cfg, k) for k in ("vocab_size", "hidden_size", "intermediate_size", "num_hidden_layers", "num_attention_heads",
                                       "num_key_value_heads", "hidden_act", "head_dim", "norm_class", "num_linears_per_mlp",
                                       "partial_rotary_factor", "rope_theta", "normalize_embed", "layer_norm_eps")}
    torch.manual_seed(1337)
    from mobilequant_b200.model import HFForCausalLM
    sd = {k: v.detach() for k, v in HFForCausalLM(cfg).float().state_dict().items()}
    T = args.seqlen
    ids = synth_ids(args.cpu_seq, T, cfg.vocab_size, 1337)
    act = mr.act_range(sd, cd, [ids[:1]])
    qs = mr.QState(mr.default_recipe(cd, 8, False, False, 8), act)
    times = []
    with torch.no_grad():
        warm = min(args.warmup, 1)              # a CPU step is ~7 s: one warm-up pass (page-in, thread pool) is enough
        for i in range(warm + args.steps):
            t0 = time.perf_counter()
            logits, _ = mr.model_forward(sd, cd, ids, qs, quant=True)
            nxt = logits[:, -1].argmax(-1)
            dt = time.perf_counter() - t0
            if i >= warm:
                times.append(dt)
            elif args.steps * dt > 150.0 and ids.shape[1] > 128:
                # bounded sample: keep the whole K-step run within a few minutes by shortening the sequence of a step
                # (shorter sequences cost the CPU path less attention per token: the reported tok/s errs in its favour)
                T = max(128, int(T * 150.0 / (args.steps * dt)) // 64 * 64)
                ids = ids[:, :T].contiguous()
    tot = sum(times)
    v = args.steps * ids.numel() / tot
    line = {"impl": "reference", "metric": "int8_tok_per_s", "value": v, "unit": "tok/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": warm, "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.model} W8A8 static fake-quant forward (reference CPU path), {args.cpu_seq} x seq{T} per step",
                       "timing": "host wall clock, inputs larger than LLC (4.4 GB fp32 weights)"},
            "cpu_baseline": {"value": v, "unit": "tok/s", "cores": cores, "kind": "port",
                             "sample": f"{args.cpu_seq} sequence(s) x {T} tokens per step, {args.steps} steps"},
            "e2e": {"value": v, "unit": "tok/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
def _trace(msg, _t0=[None]):
    """MQB200_BENCH_TRACE=1: wall-clock phase marks on stderr (start-up diagnosis; nothing of this is inside a timed region)."""
    if os.environ.get("MQB200_BENCH_TRACE") == "1":
        now = time.perf_counter()
        if _t0[0] is None:
            _t0[0] = now
        print(f"[bench rank {os.environ.get('RANK', '0')}] +{now - _t0[0]:7.1f} s  {msg}", file=sys.stderr, flush=True)


def run_ours(args):
    _trace("start")
    import torch
    import torch.distributed as dist
    from mobilequant_b200 import kernels as K
    from mobilequant_b200.model import HFForCausalLM
    from mobilequant_b200.engine import IntEngine
    from mobilequant_b200.quantization import qmodule as Q
    from mobilequant_b200.ptq.generate_act_range import get_act_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _trace("process group up")
    torch.backends.cuda.matmul.allow_tf32 = True          # as ptq/mobilequant.py:91 (only the fp lm_head / teacher use it)
    cfg = model_cfg(args)
    T, B = args.seqlen, args.batch
    torch.manual_seed(1337)
    with torch.device(dev):
        model = HFForCausalLM(cfg).float()
    model.eval()
    _trace("model built")
    # activation ranges from a short calibration pass on synthetic ids (config 1 of BASELINE.json, on the GPU; with
    # several ranks the two samples are sharded and the packed ranges all-reduced, identical on every replica)
    # (at least one sample per rank: the pass shards samples i % world == rank)
    act = get_act_range(model, [synth_ids(1, T, cfg.vocab_size, 7 + i) for i in range(max(2, world))])
    from mobilequant_b200.ptq.generate_qcfg import default_qcfg
    wq = Q.QuantConfig(bitwidth=8) if args.wbits == 8 else Q.QuantConfig(bitwidth=4, is_symmetric=True, is_per_channel=True)
    qcfg = default_qcfg(cfg, wq, Q.QuantConfig(bitwidth=8))
    _trace("act ranges done")
    eng = IntEngine(model, qcfg, act, dev)
    _trace("engine built")
    wtag = "W8A8" if args.wbits == 8 else "W4A8 per-channel symmetric (weights packed 2/byte in HBM, %.2f GB, expanded per layer into L2)" % (eng.weight_bytes() / 1e9)
    ids_host = synth_ids(B, T, cfg.vocab_size, 1000 + rank).pin_memory()
    ids_dev = ids_host.to(dev)
    out_host = torch.empty(B, dtype=torch.int64).pin_memory()

    def step_resident():
        logits = eng(ids_dev)
        return logits[:, -1].argmax(-1)

    def step_e2e():
        d = ids_host.to(dev, non_blocking=True)
        nxt = eng(d)[:, -1].argmax(-1)
        out_host.copy_(nxt, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return nxt

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    for _ in range(max(3, args.warmup)):
        step_resident()
    if args.profile_step:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step_resident()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    _trace("warm-up done")
    K.reset_launch_count()
    with ClockSampler(local) as clk:
        ms = timed(step_resident, args.steps)
    launches = K.launch_count()
    for _ in range(2):
        step_e2e()
    ms_e2e_dev = timed(step_e2e, args.steps)
    _trace("timed regions done")
    tokens_step = B * T * world
    value = tokens_step * args.steps / (ms / 1e3)
    e2e = tokens_step * args.steps / (ms_e2e_dev / 1e3)

    line = {"metric": "int8_tok_per_s", "value": value, "unit": "tok/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": f"{args.model} {wtag} static-quant integer forward, batch {B} x seq {T} per GPU, random-init weights",
                       "global_batch": B * world, "seq_len": T, "parallelism": f"replicas x{world} (batch-sharded, no collective)",
                       "l2": "inputs larger than L2: 1.1 GB int8 weights + 1 GB logits streamed per step"},
            "clocks": clk.summary(), "gpu_launches": launches,
            "e2e": {"value": e2e, "unit": "tok/s", "h2d_bytes_per_step": ids_host.numel() * 8 * world,
                    "d2h_bytes_per_step": out_host.numel() * 8 * world}}

    if rank == 0:
        line["roofline"], line["roofline_gemm"], line["kernel_shares"] = roofline(eng, ids_dev, B, T, line["clocks"])
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(model, cfg, act, T, args.cpu_seq)
        if not args.no_decode and world == 1:
            Bd = args.decode_batch or B
            ids_dec = ids_dev if Bd == B else synth_ids(Bd, T, cfg.vocab_size, 2000 + rank).to(dev)
            line["decode"] = decode_throughput(eng, ids_dec, Bd, T, args.decode_steps)
    if not args.no_calib:
        # the calibration loops are sample-sharded over the ranks (one optimiser step = `world` micro-batches, gradients of the
        # learnables all-reduced over NCCL inside the captured step): every rank runs the leg, rank 0 reports the aggregate
        del eng
        torch.cuda.empty_cache()
        nsamp = (args.calib_samples + world - 1) // world * world
        calib = calib_throughput(model, wq, act, cfg, T, dev, nsamp, world, rank)
        if rank == 0:
            line["calib"] = calib
    if rank == 0:
        print(json.dumps(line), flush=True)
    _trace("line printed")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    _trace("done")


def roofline(eng, ids, B, T, clocks=None):
    """Per-kernel-class device time of one forward (CUDA events around every launch) and two roofline entries: the kernel
    with the largest share of the step (`roofline`) and the tcgen05 int8 GEMM the north star's >= 70 % target is stated on
    (`roofline_gemm`); both against 2 x the measured bf16 peak (int8 issues at twice the bf16 rate on sm_100): the BURST
    figure unless the clock record of this very run shows the board power cap (then the sustained one)."""
    import torch
    from mobilequant_b200 import kernels as K
    peaks = {}
    pth = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pth):
        peaks = json.load(open(pth))
    K.enable_event_timing(True)
    eng(ids)
    torch.cuda.synchronize()
    per = K.collect_event_timing()
    K.enable_event_timing(False)
    tot = sum(v["ms"] for v in per.values())
    shares = {k: {"ms": round(v["ms"], 4), "launches": v["n"], "share": round(v["ms"] / tot, 4)} for k, v in per.items()}
    capped = bool(clocks) and "sw_power_cap" in (clocks.get("reasons") or [])
    burst, sustained = peaks.get("bf16_tflops"), peaks.get("bf16_tflops_sustained")
    bf16 = (sustained if capped else burst) or burst or sustained
    peak = 2.0 * bf16 if bf16 else 2 * 1590.0
    peak_source = ("2 x measured %s bf16 (MEASURED_PEAKS.json %s; %s); int8 issues at twice the bf16 rate on sm_100" % (
        ("sustained", "bf16_tflops_sustained", "sw_power_cap seen in this run's clock record") if capped else
        ("burst", "bf16_tflops", "no power cap in this run's clock record"))) if bf16 else "2 x fallback bf16 1.59 PF"
    M = B * T
    cfg = eng.cfg
    L = cfg.num_hidden_layers

    def traffic(name):
        tpath = os.path.join(ROOT, "profiles", name)                 # ncu --set full: dram read + write per launch
        return json.load(open(tpath)).get("avg_bytes_per_launch") if os.path.exists(tpath) else None

    # library INT8 proxy measured in the same run
    a = torch.randint(-128, 127, (8192, 8192), dtype=torch.int8, device=ids.device)
    b = torch.randint(-128, 127, (8192, 8192), dtype=torch.int8, device=ids.device)
    for _ in range(3):
        torch._int_mm(a, b.t())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        torch._int_mm(a, b.t())
    e1.record(); torch.cuda.synchronize()
    lib = 10 * 2 * 8192 ** 3 / (e0.elapsed_time(e1) / 1e3) / 1e12

    g = per["qgemm"]
    ops = 2.0 * M * (eng.H * (eng.nh + 2 * eng.nkv) * eng.hd + eng.nh * eng.hd * eng.H + 2 * eng.Ipad * eng.H + eng.Ipad * eng.H) * L
    achieved = ops / (g["ms"] / 1e3) / 1e12
    rl_gemm = {"kernel": "qgemm_kernel (tcgen05 kind::i8, TMA ring, TMEM double buffer)", "bound": "tensor", "achieved": achieved, "peak": peak,
               "unit": "TOP/s", "frac": achieved / peak, "traffic": traffic("r1d_qgemm_traffic.json"), "peak_source": peak_source,
               "frac_of_burst_peak": (achieved / (2.0 * burst)) if burst else None,
               "avg_launch_ms": g["ms"] / g["n"], "launches_per_step": g["n"], "ops_per_step": ops,
               "int8_library_proxy_tops": lib, "frac_of_library_proxy": achieved / lib, "frac_of_spec_4500": achieved / 4500.0}
    rl = rl_gemm
    at = per.get("qattn")
    if at:
        # algorithmic int8 work of the exact quantised attention: Q.K^T (2 hd ops per score) + the two byte-plane P.V
        # contractions (2 x 2 hd), over the causally visible scores only
        scores = B * eng.nh * T * (T + 1) / 2.0 * L
        aops = scores * 6.0 * eng.hd
        ach = aops / (at["ms"] / 1e3) / 1e12
        tc = eng.hd in (64, 128) and T % 16 == 0 and os.environ.get("MQB200_QATTN", "tc")[:2] == "tc"
        shares["qattn"].update({"scores_per_s": scores / (at["ms"] / 1e3), "scores_per_step": scores})
        rl_attn = {"kernel": ("qattn_tc_kernel (tcgen05 kind::i8 S = Q.K^T and hi/lo P.V into TMEM, TMA-fed K/V, thread-per-row exact softmax)"
                              if tc else "qattn4_kernel / qattn_kernel (mma.sync)"),
                   "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TOP/s", "frac": ach / peak,
                   "traffic": traffic("r2_qattn_tc_traffic.json") if tc else None, "peak_source": peak_source,
                   "avg_launch_ms": at["ms"] / at["n"], "launches_per_step": at["n"], "ops_per_step": aops,
                   "note": "the contraction is < 5 % of this kernel: it is bound by the ALU pipe (exact 16-bit quantised softmax, ~70 "
                           "integer/fp32 instructions per score, profiles/r2_ncu_summary.md), not by the tensor pipe or HBM; "
                           "scores/s is the meaningful rate (kernel_shares.qattn)"}
        if at["ms"] > g["ms"]:
            rl = rl_attn
        else:
            shares["qattn"]["roofline"] = rl_attn
    return rl, rl_gemm, shares


def decode_throughput(eng, ids, B, T, nsteps):
    """Greedy decode against the int8 KV cache: prefill T - nsteps tokens, then nsteps one-token steps replayed as one
    CUDA graph each (token -> logits -> argmax -> token and the position stay on the device).  HBM roofline: every step
    streams the int8 weights, the fp32 lm_head and the K/V codes of the tokens seen so far exactly once."""
    import torch
    peaks = {}
    pth = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pth):
        peaks = json.load(open(pth))
    T0 = T - nsteps
    cache = eng.new_cache(B, T)
    logits = eng.prefill(ids[:, :T0].contiguous(), cache)
    graph, tokens, _ = eng.capture_decode(cache)
    tokens.copy_(logits.argmax(-1))
    for _ in range(3):                                   # warm replays, then rewind the position
        graph.replay()
    graph.rewind(T0)
    tokens.copy_(logits.argmax(-1))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(nsteps):
        graph.replay()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / nsteps
    wbytes = sum(L[k]["N"] * L[k]["K"] for L in eng.layers for k in ("qkv", "o", "w13", "w2"))   # one byte per code reaches the GEMV
    head = eng.lm_head.numel() * 4
    avg_keys = T0 + (nsteps + 1) / 2.0
    kv = 2.0 * len(eng.layers) * B * eng.nkv * avg_keys * eng.hd
    bytes_step = wbytes + head + kv
    peak = peaks.get("hbm_gbs", 6500.0)
    ach = bytes_step / (ms / 1e3) / 1e9
    return {"value": B * 1e3 / ms, "unit": "tok/s", "batch": B, "context": T0, "steps": nsteps, "ms_per_step": ms,
            "what": "greedy decode, one CUDA-graph replay per token, kernels chained by programmatic dependent launch: qnorm, tcgen05 "
                    "skinny GEMM (split-K, integer red.add, weight tiles requested ahead of the dependency) + requant epilogue x4, fused "
                    "RoPE/append/attention on the int8 KV cache per layer; fp32 lm_head + argmax",
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "bytes_per_step": bytes_step, "weights_int8": wbytes, "lm_head_fp32": head, "kv_cache_read": kv,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6500 GB/s"}}


def calib_throughput(model, wq, act, cfg, T, dev, nsamples=96, world=1, rank=0):
    """MobileQuant e2e calibration (LET + LWC + LRL, experiments/w8a8/main/e2e_llama-s1024-ep60.sh learning rates):
    FP-target pass + one optimiser step per `world` samples (CUDA-graph replays after the first two), fuse, parameters.pth.
    world > 1: samples are sharded over the ranks, one SUM all-reduce of the flat gradient buffer per step (== the reference
    with --batch_size world, alg:532-533); the reported rate is global samples over the slowest rank's wall time."""
    import types, tempfile, torch
    import torch.distributed as dist
    from mobilequant_b200.quantization import qmodule as Q, algorithm as A

    class _L:
        def info(self, *a, **k):
            pass
    Q.create_sim_qmodel(model, wq, Q.QuantConfig(bitwidth=8))
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
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    A.e2equant(args, model, loader, _L(), device=dev)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    nlearn = sum(v.numel() for d in torch.load(os.path.join(out, "parameters.pth"), weights_only=False).values() for v in d.values()) if rank == 0 else 0
    ar_ms = None
    if world > 1:
        t = torch.tensor([dt], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = t.item()
        # cost of the step's one collective, measured on its own: SUM all-reduce of a buffer of the learnables' size
        n = torch.tensor([nlearn], device=dev, dtype=torch.int64)
        dist.broadcast(n, 0)
        buf = torch.zeros(int(n.item()), device=dev)
        for _ in range(5):
            dist.all_reduce(buf)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            dist.all_reduce(buf)
        e1.record(); torch.cuda.synchronize()
        ar_ms = e0.elapsed_time(e1) / 20
    full = None
    fpath = os.path.join(ROOT, "profiles", "r2_calib512.json")         # the same recipe run on all 512 samples (scratch/calib512.py)
    if os.path.exists(fpath) and world == 1:
        fr = json.load(open(fpath))
        full = {"samples": fr["samples"], "seconds": fr["seconds"], "samples_per_s": fr["samples_per_s"], "source": "profiles/r2_calib512.json"}
    steps = nsamples // world
    res = {"value": nsamples / dt, "unit": "samples/s", "samples": nsamples, "seconds": dt, "optimizer_steps": steps,
           "learnable_scalars": nlearn, "full_512_sample_run": full,
           "what": "e2equant LET+LWC+LRL, micro-batch 1 per rank, seq %d, 1 epoch, includes FP-target pass, fuse and parameters.pth save" % T,
           "path": "per decoder block: fused QRMSNorm (both quantizers), ONE q|k|v GEMM + fused quantise/RoPE/quantise, fused score-quantise/"
                   "scale/mask/softmax/quantise, ONE w1|w3 GEMM + fused gated-SiLU core (5 quantizers), fused weight pass (LET+LWC+fake-quant) on a "
                   "side stream, flat AdamW -- all libmqb200 kernels, forward and backward; the GEMMs themselves are library TF32 (cuBLAS); whole "
                   "step replayed as one CUDA graph; 512 samples at this rate: %.1f s (fixed costs included pro rata)" % (512 * dt / nsamples)}
    if world > 1:
        res["parallelism"] = "dp%d: samples sharded i %% world == rank, one NCCL SUM all-reduce of the %d learnable-scalar gradients per step" % (world, nlearn)
        res["allreduce_ms_per_step"] = ar_ms
        res["allreduce_share_of_step"] = ar_ms / (1e3 * dt / steps) if steps else None
    return res


def cpu_baseline(model, cfg, act, T, nseq):
    """The oracle port of the reference's fp32 fake-quant forward on the host cores, bounded sample."""
    import torch
    from oracle import model_ref as mr
    torch.set_num_threads(os.cpu_count() or 1)
    cd = {k: getattr(cfg, k) for k in ("vocab_size", "hidden_size", "intermediate_size", "num_hidden_layers", "num_attention_heads",
                                       "num_key_value_heads", "hidden_act", "head_dim", "norm_class", "num_linears_per_mlp",
                                       "partial_rotary_factor", "rope_theta", "normalize_embed", "layer_norm_eps")}
    from mobilequant_b200.quantization import qmodule as Q
    sd = {k: v.detach().float().cpu() for k, v in model.state_dict().items() if "quantizer" not in k and "smooth" not in k}
    qs = mr.QState(mr.default_recipe(cd, 8, False, False, 8), act)
    ids = synth_ids(nseq, T, cfg.vocab_size, 1337)
    with torch.no_grad():
        t0 = time.perf_counter()
        logits, _ = mr.model_forward(sd, cd, ids, qs, quant=True)
        logits[:, -1].argmax(-1)
        dt = time.perf_counter() - t0
    return {"value": ids.numel() / dt, "unit": "tok/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{nseq} sequence(s) x {T} tokens, one fp32 fake-quant forward (oracle/model_ref.py), {dt:.1f} s"}


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
