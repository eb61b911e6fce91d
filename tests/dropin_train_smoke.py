#!/usr/bin/env python
"""Drop-in check at the boundary the reference's trainer uses, in a fresh interpreter:

    sys.modules["sam.sa_m4c"] = sam_textvqa_b200.sa_m4c        # `from sam.sa_m4c import SAM4C, BertConfig`, train.py:15
    import sam.task_utils                                       # UNMODIFIED (oracle/_ref or /root/reference)

then the body of train.py's loop (:127-144) -- forward_model (task_utils.py:99-135: batch to the GPU, model(batch_dict),
the reference's own loss and TextVQAAccuracy metric), loss.backward(), clip_gradients, get_optim_scheduler's Adam +
LambdaLR, model.zero_grad() -- drives the samk module through fake loaders for a few iterations.  The same loop also
drives the UNMODIFIED reference model (same seeded weights, on the GPU in eager PyTorch) and the two loss curves,
accuracies and predictions are printed as one JSON line.  TEST / EVIDENCE TOOL: not part of the product path.
"""
import argparse
import json
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--precision", default="f16")
    ap.add_argument("--layers", default="n,s")
    args = ap.parse_args()
    os.environ["SAMK_PRECISION"] = args.precision
    import numpy as np
    import torch
    from oracle import ref_loader
    if not ref_loader.reference_available():
        print(json.dumps({"skipped": "no reference tree (oracle/_ref) on this machine"}))
        return
    for p in (ref_loader.REF_ROOT, ref_loader.SHIM_DIR):
        if p not in sys.path:
            sys.path.insert(0, p)
    M, S, registry = ref_loader.load_reference(500)            # the unmodified modules, before the alias
    import sam_textvqa_b200.sa_m4c as samk_model                # binds to tools.registry (it is importable here)
    import sam_textvqa_b200.spatial_utils as samk_spatial
    sys.modules["sam.sa_m4c"] = samk_model
    sys.modules["sam.spatial_utils"] = samk_spatial
    import sam.task_utils as TU                                 # unmodified; its `sam.sa_m4c` users now get samk
    from tools.objects_to_byte_tensor import enc_obj2bytes
    from sam_textvqa_b200 import synth
    from sam_textvqa_b200.config import c3_config
    assert samk_model.registry is registry

    V = 500

    class Vocab(list):                                          # what TextVQAAccuracy needs of registry.answer_vocab
        def idx2word(self, i):
            return self[i]
    registry.answer_vocab = Vocab("w%d" % i for i in range(V))
    registry.BOS_IDX, registry.EOS_IDX, registry.PAD_IDX = 1, 2, 0

    kinds = args.layers.split(",")
    mmt, tb = c3_config(layer_type_list=kinds, mix_list=["share3" if k == "s" else "none" for k in kinds],
                        hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, obj_drop=0.0, ocr_drop=0.0)
    tb = dict(tb, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    task_cfg = {"loss": "textvqa", "metric": "textvqa", "warmup_iters": 2, "warmup_factor": 0.2,
                "lr_decay_iters": [14000, 19000], "lr_decay": 0.1}
    device = torch.device("cuda:0")
    graph_fn = lambda boxes: samk_spatial.build_graph_batch(boxes, 0.5)[0]
    B = 4

    def loader(n):
        out = []
        for i in range(n):
            b = synth.make_batch(B, O=36, V=V, seed=20 + i, contexts=(1, 3), graph_fn=graph_fn)
            b.pop("boxes"); b.pop("spatial_types")
            b["question_id"] = torch.arange(B) + 100 * i
            b["ocr_tokens"] = torch.stack([enc_obj2bytes(["tok%d" % j for j in range(50)]) for _ in range(B)])
            b["answers"] = torch.stack([enc_obj2bytes(["w5 w7", "w5"] * 5) for _ in range(B)])
            out.append(b)
        return out

    def run(model):
        """train.py:127-144, restated call for call"""
        model = model.to(device).train()
        base_lr = 1e-4
        optimizer, warmup_scheduler = TU.get_optim_scheduler(task_cfg, model.get_optimizer_parameters(base_lr), base_lr)
        loaders = {"train": loader(args.iters)}
        losses, accs, preds = [], [], []
        for _ in range(args.iters):
            loss, acc, bs, predictions = TU.forward_model(task_cfg, device, model, loaders, "train")
            loss.backward()
            TU.clip_gradients(model, 0.25)
            optimizer.step()
            warmup_scheduler.step()
            model.zero_grad()
            losses.append(float(loss))
            accs.append(float(acc))
            preds.append([p["pred_answer"] for p in predictions])
            assert bs == B
        return losses, accs, preds

    def build(mod, seed=0):
        model = mod.SAM4C(mod.BertConfig.from_dict(mmt), mod.BertConfig.from_dict(tb))
        sd = model.state_dict()
        model.load_state_dict(synth.seeded_state([(k, v.shape) for k, v in sd.items()], seed), strict=True)
        return model

    ours = run(build(samk_model))
    theirs = run(build(M))
    from sam_textvqa_b200 import ops
    print(json.dumps({"precision": ops.get_precision(), "iters": args.iters, "layers": kinds,
                      "loss_samk": ours[0], "loss_reference": theirs[0], "acc_samk": ours[1], "acc_reference": theirs[1],
                      "predictions_equal": ours[2] == theirs[2], "kernels_launched": ops.launch_count,
                      "reference_tree": ref_loader.REF_ROOT}))


if __name__ == "__main__":
    main()
