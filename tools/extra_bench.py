#!/usr/bin/env python
"""Secondary measurements for BASELINE.md section 4: the other BASELINE.json configs on one B200.

  cfg1  1 spatial layer, 20+36+50 tokens, B=4           fwd+bwd samples/s
  cfg2  shipped c3 yml, B=128                            greedy 12-step decode samples/s (cached vs D-pass loop)
  cfg3  c5 context, 100 obj + 100 OCR, B=256             fwd+bwd samples/s
  cfg5  joint tokens 106/256/512/1024, B=64              fwd+bwd samples/s (attention GB/s: tools/attn_bench.py --sweep)
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sam_textvqa_b200 import dp, ops, spatial_utils, synth
from sam_textvqa_b200.config import c3_config
from sam_textvqa_b200.registry import registry
from sam_textvqa_b200.sa_m4c import SAM4C, BertConfig

dev = torch.device("cuda:0")
V = 5000
registry.answer_vocab = ["w%d" % i for i in range(V)]
registry.BOS_IDX = 1
graph_fn = lambda boxes: spatial_utils.build_graph_batch(boxes.astype("float32"), 0.5)[0]


def build(mmt_over):
    mmt, tb = c3_config(**mmt_over)
    torch.manual_seed(0)
    return SAM4C(BertConfig.from_dict(mmt), BertConfig.from_dict(tb)).to(dev)


def timeit(fn, steps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / steps


def train_bench(name, mmt_over, B, O, R, ctx, steps=5):
    model = build(mmt_over).train()
    grads = dp.FlatGradBuffer(model.parameters())
    batch = synth.make_batch(B, O=O, R=R, V=V, seed=0, contexts=(ctx,), graph_fn=graph_fn)
    adj = {str(ctx): batch.pop("spatial_adj_matrices")[str(ctx)].to(dev)}
    if ctx != 1:
        adj["1"] = adj[str(ctx)]
    res = {k: v.to(dev) for k, v in batch.items() if torch.is_tensor(v)}

    def step():
        grads.zero()
        bd = dict(res); bd["spatial_adj_matrices"] = adj
        loss = ops.bce_with_mask_loss(model(bd)["textvqa_scores"], res["targets"], res["train_loss_mask"])
        loss.backward()
    ms = timeit(step, steps)
    out = {"config": name, "B": B, "L": 20 + O + R + 12, "ms_per_step": ms, "samples_per_s": B / ms * 1e3,
           "peak_mem_GB": torch.cuda.max_memory_allocated() / 2**30}
    print(json.dumps(out), flush=True)
    del model, grads, res, adj
    torch.cuda.empty_cache(); torch.cuda.reset_peak_memory_stats()


def eval_bench(B=128):
    model = build({}).eval()
    batch = synth.make_batch(B, V=V, seed=0, contexts=(3,), graph_fn=graph_fn)
    adj = {"3": batch.pop("spatial_adj_matrices")["3"].to(dev)}
    res = {k: v.to(dev) for k, v in batch.items() if torch.is_tensor(v)}
    for mode in ("cached", "reference"):
        os.environ["SAMK_GREEDY"] = mode

        def step():
            bd = dict(res); bd["spatial_adj_matrices"] = adj
            with torch.no_grad():
                model(bd)
        ms = timeit(step, 3, 1)
        print(json.dumps({"config": "cfg2 greedy 12-step decode (%s)" % mode, "B": B, "ms_per_batch": ms,
                          "samples_per_s": B / ms * 1e3}), flush=True)


which = sys.argv[1:] or ["cfg1", "eval", "cfg3", "cfg5"]
if "cfg1" in which:
    train_bench("cfg1 1xs layer 20+36+50, B=4", dict(layer_type_list=["s"], mix_list=["share3"]), 4, 36, 50, 3, steps=20)
if "eval" in which:
    eval_bench()
if "cfg3" in which:
    train_bench("cfg3 c5, 100 obj + 100 OCR, B=256",
                dict(mix_list=["none", "none", "share5", "share5", "share5", "share5"], ocr_feature_size=3002), 256, 100, 100, 5, steps=3)
if "cfg5" in which:
    for O in (36, 186, 442, 954):
        train_bench("cfg5 joint=%d tokens, B=64" % (20 + O + 50), {}, 64, O, 50, 3, steps=3)
