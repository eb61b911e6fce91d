#!/usr/bin/env python
"""Data-parallel correctness on real GPUs (launched by torchrun, one process per GPU, NCCL):

    the global batch sharded over N ranks + dp.global_loss_scale + ONE summed all-reduce of the flat gradient buffer
 == the same global batch on one GPU (loss and every gradient),

which is what replaces nn.DataParallel (train.py:111-112: scatter, replicate, gather the scores, global loss
normaliser sam/task_utils.py:28-29).  Rank 0 computes the single-GPU result itself and prints one JSON line.
TEST / EVIDENCE TOOL.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from sam_textvqa_b200 import dp, ops, spatial_utils, synth
    from sam_textvqa_b200.config import c3_config
    from sam_textvqa_b200.registry import registry
    from sam_textvqa_b200.sa_m4c import SAM4C, BertConfig
    from tests._util import rel_err, sam4c_state_shapes
    precision = os.environ.get("SAMK_PRECISION", "f16")
    V, B = 500, 4 * world
    registry.answer_vocab = ["w%d" % i for i in range(V)]
    registry.BOS_IDX = 1
    mmt, tb = c3_config(layer_type_list=["n", "s", "s"], mix_list=["none", "share3", "share3"], hidden_dropout_prob=0.0,
                        attention_probs_dropout_prob=0.0, obj_drop=0.0, ocr_drop=0.0)
    tb = dict(tb, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    model = SAM4C(BertConfig.from_dict(mmt), BertConfig.from_dict(tb))
    model.load_state_dict(synth.seeded_state(sam4c_state_shapes(mmt, tb, V), 3), strict=True)
    model = model.to(dev).train()
    grads = dp.FlatGradBuffer(model.parameters())
    graph_fn = lambda boxes: spatial_utils.build_graph_batch(boxes, 0.5)[0]
    full = synth.make_batch(B, O=36, V=V, seed=11, contexts=(1, 3), graph_fn=graph_fn)
    full.pop("boxes"); full.pop("spatial_types")
    full["train_loss_mask"][1, 5:] = 0          # unequal numbers of valid steps per rank: the normaliser must be global
    full["train_loss_mask"][B - 1, 2:] = 0

    def run(batch, scale):
        grads.zero()
        bd = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
        loss = ops.bce_with_mask_loss(model(bd)["textvqa_scores"], bd["targets"], bd["train_loss_mask"]) * scale
        loss.backward()
        return loss.detach()

    shard = dp.shard_batch(full, rank, world)
    scale = dp.global_loss_scale(shard["train_loss_mask"].to(dev))
    transport = os.environ.get("SAMK_DP_TRANSPORT", "nccl")
    buckets = 0
    if transport == "peer":
        # overlapped, bucketed exchange through samk_exchange_sum: step 1 learns the report counts (plain all-reduce),
        # step 2 sends finished buckets under the backward pass
        grads.enable_overlap(average=False, bucket_bytes=4 << 20, transport="peer")
        for _ in range(2):
            grads.zero()
            grads.begin_step()
            bd = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in shard.items()}
            loss = ops.bce_with_mask_loss(model(bd)["textvqa_scores"], bd["targets"], bd["train_loss_mask"]) * scale
            loss.backward()
            n0 = ops.launch_count
            grads.finish_step()
        loss = loss.detach()
        torch.cuda.synchronize()
        grads._ov["peer"].check()
        buckets = len(getattr(grads._ov["peer"], "log", []))
        ops.grad_ready_hook = None
    else:
        loss = run(shard, scale)
        dp.GradExchange(grads, world).all_reduce()
    dist.all_reduce(loss)                      # the global loss is the sum of the scaled local ones
    got = grads.flat.clone()
    out = None
    if rank == 0:
        ref_loss = run(full, 1.0)
        want = grads.flat
        out = {"world": world, "precision": precision, "loss_dp": float(loss), "loss_single": float(ref_loss),
               "grad_rel_err": rel_err(got, want), "grad_norm": float(want.norm()), "transport": transport, "buckets": buckets,
               "nccl": ".".join(str(x) for x in torch.cuda.nccl.version())}
    dist.barrier()
    if out is not None:
        print(json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
