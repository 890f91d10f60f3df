"""TEST INFRASTRUCTURE ONLY -- golden for the `.encodings` writer.

device/utils.py imports aimet_torch / onnx at module level and cannot be imported, so the UNMODIFIED source text of its four
functions (update_encodings_from_min_max, prefix_match_linear, override_encoding, update_encodings) is cut out of the file
and executed here on a synthetic AIMET-style encodings file (node names as AIMET assigns them: prefixes + a numeric suffix
for the module-named nodes, counters for the functional ones) and a synthetic act_dict.  Also the kv-cache statement block
of device/calibrate.py:275-285.  Writes tests/golden/encodings.json.

    python oracle/make_golden_encodings.py
"""
import os, sys, json, ast, copy, types, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/device/utils.py"

src = open(REF).read()
tree = ast.parse(src)
want = {"update_encodings_from_min_max", "prefix_match_linear", "override_encoding", "update_encodings"}
ns = {}
for node in tree.body:
    if isinstance(node, ast.FunctionDef) and node.name in want:
        exec(compile(ast.Module(body=[node], type_ignores=[]), REF, "exec"), ns)
assert want <= set(ns), want - set(ns)


def synthetic_inputs(num_blocks, slinear, silu, seed):
    rnd = random.Random(seed)
    mods = ["input_layernorm", "post_attention_layernorm", "self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj", "self_attn.o_proj",
            "self_attn.qk_bmm", "self_attn.pv_bmm", "mlp.w1", "mlp.w3", "mlp.w2", "mlp.act_fn"]
    act = {}
    for i in range(num_blocks):
        for m in mods:
            e = {}
            for f in ("input", "input2", "output"):
                lo = -rnd.uniform(0.1, 9.0)
                e[f] = [lo, rnd.uniform(0.1, 9.0)]
            act[f"model.layers.{i}.{m}"] = e
    blank = lambda b: {"bitwidth": b, "dtype": "int", "is_symmetric": "False", "max": 1.0, "min": -1.0, "offset": -128, "scale": 0.01}
    io = lambda b: {"input": {"0": blank(b), "1": blank(b)}, "output": {"0": blank(b)}}
    nodes = {"module_embedding": io(16), "module_add_mask": io(16)}           # two nodes the reference leaves alone
    for i in range(num_blocks):
        L, A, M = f"layers.{i}.", f"layers.{i}.self_attn.", f"layers.{i}.mlp."
        suf = lambda: f"_{rnd.randint(1, 99)}"
        for n in ("input_layernorm", "post_attention_layernorm"):
            nodes[L + n + ".module_normalize" + suf()] = io(16)
            nodes[L + n + ".module_mul" + suf()] = io(16)
        for n in ("q_proj", "k_proj", "v_proj", "o_proj"):
            nodes[A + n] = io(8 if n != "o_proj" else 16)
        nodes[M + "w1"] = io(8); nodes[M + "w3"] = io(8)
        if silu:
            nodes[M + "act.sigmoid"] = io(8); nodes[M + "act.mul"] = io(8)
        nodes[M + ("w2.linear" if slinear else "w2")] = io(16)
        nodes["module_matmul" if i == 0 else f"module_matmul_{2 * i}"] = io(16)
        nodes[f"module_matmul_{2 * i + 1}"] = io(16)
        nodes[A + "softmax"] = io(16)
        nodes[f"module_add_{5 * i + 3}"] = io(16); nodes[f"module_add_{5 * i + 4}"] = io(16)
        nodes[A + "module_reshape" if i == 0 else A + f"module_reshape_{6 * i}"] = io(8)
        for j in range(1, 6):
            nodes[A + f"module_reshape_{6 * i + j}"] = io(8)
        nodes[A + "module_transpose" if i == 0 else A + f"module_transpose_{5 * i}"] = io(8)
        for j in range(1, 5):
            nodes[A + f"module_transpose_{5 * i + j}"] = io(8)
        nodes["module_cat" if i == 0 else f"module_cat_{2 * i}"] = io(8)
        nodes[f"module_cat_{2 * i + 1}"] = io(8)
        m = 8 if slinear else 7
        for j in (1, 2, 3, 4, 6) + ((7,) if slinear else ()):
            nodes[f"module_mul_{m * i + j}"] = io(8)
    return act, {"activation_encodings": nodes, "param_encodings": {"lm_head.weight": [blank(8)]}}


def kv_cache(act, num_blocks, bitwidth):
    """device/calibrate.py:275-285 on act_dict ranges (the script reads them back from the ctx encodings it just wrote)."""
    k = [act[f"model.layers.{i}.self_attn.qk_bmm"]["input2"] for i in range(num_blocks)]
    v = [act[f"model.layers.{i}.self_attn.pv_bmm"]["input2"] for i in range(num_blocks)]
    k_min, k_max, v_min, v_max = min(r[0] for r in k), max(r[1] for r in k), min(r[0] for r in v), max(r[1] for r in v)
    qmax = 2 ** bitwidth - 1
    k_scale, v_scale = (k_max - k_min) / qmax, (v_max - v_min) / qmax
    k_cache_enc = {"bitwidth": bitwidth, "dtype": "int", "is_symmetric": "False", "max": k_max, "min": k_min, "offset": int(k_min / k_scale), "scale": k_scale}
    v_cache_enc = {"bitwidth": bitwidth, "dtype": "int", "is_symmetric": "False", "max": v_max, "min": v_min, "offset": int(v_min / v_scale), "scale": v_scale}
    return {"k_cache": k_cache_enc, "v_cache": v_cache_enc}


out = {}
# (GELU models have no act.sigmoid node: the reference's prefix_match_linear asserts exactly one match (utils.py:292) before
#  its `if len(tgt_name) > 0` can skip the block, so it cannot process them; the product skips the sigmoid rows instead)
for tag, nb, slinear, silu, qf in (("llama_slinear", 3, True, True, 0.125), ("llama_plain", 2, False, True, 0.125), ("llama_12_blocks", 12, True, True, 0.0625)):
    act, ori = synthetic_inputs(nb, slinear, silu, 1337 + nb)
    cfg = types.SimpleNamespace(impl_sym_pch_as_slinear=slinear)
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        res = ns["update_encodings"](copy.deepcopy(ori), copy.deepcopy(act), nb, qf, cfg)
    out[tag] = dict(num_blocks=nb, impl_sym_pch_as_slinear=slinear, q_proj_factor=qf, act_dict=act, ori_encodings=ori, updated=res,
                    kv_cache=kv_cache(act, nb, 8))
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "encodings.json"), "w"), indent=1, sort_keys=True)
print({k: len(v["updated"]["activation_encodings"]) for k, v in out.items()})
