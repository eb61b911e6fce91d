"""GPU probe: unit ops + full SAM4C (cfg1) against the goldens / oracle."""
import sys, os, time, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sam_textvqa_b200 import ops, synth
from sam_textvqa_b200.registry import registry
from tests._util import cfg1, golden_batch, load_golden, rel_err, sam4c_state_shapes

dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False

def section(name):
    print("\n=== " + name, flush=True)

def guarded(fn):
    try:
        fn()
    except Exception:
        traceback.print_exc()
        sys.stdout.flush()

def t_layernorm():
    section("layernorm fwd/bwd")
    x = torch.randn(1000, 768, device=dev, requires_grad=True)
    g = (1 + 0.1 * torch.randn(768, device=dev)).requires_grad_(True)
    b = (0.1 * torch.randn(768, device=dev)).requires_grad_(True)
    y = ops.layer_norm(x, g, b, 1e-12)
    ref = torch.nn.functional.layer_norm(x, (768,), g, b, 1e-12)
    print("fwd", rel_err(y, ref))
    w = torch.randn_like(y)
    gx, gg, gb = torch.autograd.grad((y * w).sum(), (x, g, b))
    rx, rg, rb = torch.autograd.grad((ref * w).sum(), (x, g, b))
    print("dx", rel_err(gx, rx), "dg", rel_err(gg, rg), "db", rel_err(gb, rb))

def t_linear():
    section("linear fwd/bwd (both precisions)")
    for prec in ("f16", "bf16x3"):
        ops.set_precision(prec)
        for (M, N, K) in ((472, 768, 768), (144, 768, 4), (200, 768, 2952)):
            x = torch.randn(M, K + (1 if K == 4 else 0), device=dev)[:, :K]
            x.requires_grad_(True)
            W = (0.05 * torch.randn(N, K if K != 2952 else 3002, device=dev)).requires_grad_(True)
            b = torch.randn(N, device=dev, requires_grad=True)
            y = ops.linear(x, W, b, K)
            ref = x @ W[:, :K].t() + b
            w = torch.randn_like(ref)
            gx, gW, gb = torch.autograd.grad((y * w).sum(), (x, W, b))
            rx, rW, rb = torch.autograd.grad((ref * w).sum(), (x, W, b))
            print(prec, (M, N, K), "y %.2e dx %.2e dW %.2e db %.2e" % (rel_err(y, ref), rel_err(gx, rx), rel_err(gW, rW), rel_err(gb, rb)))
    ops.set_precision("f16")

def t_attention():
    section("attention vs golden attn_unit (reference SpatialBertSelfAttention)")
    g = load_golden("attn_unit.npz")
    T, A, D = int(g["T"]), int(g["A"]), int(g["D"])
    hidden = torch.from_numpy(g["hidden"]).to(dev)
    B, L, d = hidden.shape
    names = [(n, s) for n in ("query", "key", "value") for s in ("weight", "bias")]
    sd = synth.seeded_state([("%s.%s" % (n, s), (768, 768) if s == "weight" else (768,)) for n, s in names], 3)
    Wqkv = torch.cat([sd["query.weight"], sd["key.weight"], sd["value.weight"]]).to(dev)
    bqkv = torch.cat([sd["query.bias"], sd["key.bias"], sd["value.bias"]]).to(dev)
    qkv = (hidden.view(B * L, d).double() @ Wqkv.double().t() + bqkv.double()).float()
    from sam_textvqa_b200.sa_m4c import pack_relation_bits
    bits = pack_relation_bits(torch.from_numpy(g["adj"]), dev)
    valid = torch.from_numpy(g["valid"]).to(dev).to(torch.uint8).contiguous()
    dims = (B, L, 12, T, A, D)
    for dt in (torch.float32, torch.bfloat16):
        ctx, lse = ops.attention_fwd(qkv.to(dt).contiguous(), valid, bits, dims, True, 0b11, 0.0, (0, 0))
        print(dt, "ctx rel err", rel_err(ctx.float().view(B, L, d), g["ctx"]), "text rows max", ctx.view(B, L, d)[:, :T].abs().max().item())
    # backward vs torch autograd of an explicit masked softmax
    q3 = qkv.view(B, L, 3, 12, 64).clone().requires_grad_(True)
    adj = torch.from_numpy(g["adj"]).to(dev).float()
    m = torch.ones(B, L, L, 12, device=dev)
    m[:, T:T + A, T:T + A] = adj
    m[:, :T, :T] = 0; m[:, :T, T:T + A] = 0
    ext = valid.float()[:, None, :].repeat(1, L, 1)
    ext[:, -D:, -D:] = torch.tril(torch.ones(D, D, device=dev))
    allow = (m.permute(0, 3, 1, 2) > 0) & (ext[:, None] > 0)
    qq, kk, vv = q3[:, :, 0].permute(0, 2, 1, 3), q3[:, :, 1].permute(0, 2, 1, 3), q3[:, :, 2].permute(0, 2, 1, 3)
    s = (qq @ kk.transpose(-1, -2)) / 8.0
    s = s.masked_fill(~allow, float("-inf"))
    p = torch.softmax(s, -1)
    p = torch.where(allow.any(-1, keepdim=True), p, torch.zeros_like(p))
    ref = (p @ vv).permute(0, 2, 1, 3).reshape(B, L, d)
    w = torch.randn_like(ref)
    (gref,) = torch.autograd.grad((ref * w).sum(), q3)
    ctx, lse = ops.attention_fwd(qkv, valid, bits, dims, True, 0b11, 0.0, (0, 0))
    print("fwd vs torch", rel_err(ctx.view(B, L, d), ref))
    dqkv = ops.attention_bwd(w.view(B * L, d).contiguous(), qkv, ctx, lse, valid, bits, dims, True, 0b11, 0.0, (0, 0))
    gr = gref.reshape(B * L, 3, 768)
    dq = dqkv.view(B * L, 3, 768)
    print("dq %.2e dk %.2e dv %.2e" % (rel_err(dq[:, 0], gr[:, 0]), rel_err(dq[:, 1], gr[:, 1]), rel_err(dq[:, 2], gr[:, 2])))

def build_model(V=500):
    from sam_textvqa_b200.sa_m4c import SAM4C, BertConfig
    registry.answer_vocab = ["w%d" % i for i in range(V)]
    registry.BOS_IDX = 1
    mmt, tb = cfg1()
    model = SAM4C(BertConfig.from_dict(mmt), BertConfig.from_dict(tb))
    model.load_state_dict(synth.seeded_state(sam4c_state_shapes(mmt, tb, V), 0), strict=True)
    return model.to(dev)

def t_model():
    g = load_golden("sam4c_cfg1.npz")
    model = build_model()
    for prec in ("bf16x3", "f16"):
        section("SAM4C cfg1 teacher-forced, precision " + prec)
        ops.set_precision(prec)
        ops.clear_weight_cache()
        model.train()
        model.zero_grad()
        batch = golden_batch(g)
        t0 = time.time()
        scores = model(batch)["textvqa_scores"]
        ref = torch.from_numpy(g["tf/scores"])
        live = ref > -5000
        print("obj_mmt_in", rel_err(batch["obj_mmt_in"].cpu(), g["tf/obj_mmt_in"]), "ocr_mmt_in", rel_err(batch["ocr_mmt_in"].cpu(), g["tf/ocr_mmt_in"]),
              "text_bert", rel_err(batch["text_bert_emb"].cpu(), g["tf/text_bert_emb"]), "seq", rel_err(batch["mmt_seq_output"].cpu(), g["tf/mmt_seq_output"]))
        print("scores rel err (max|d|/max|ref|, live logits):", rel_err(scores.cpu(), ref, live), " argmax equal:", bool((scores.argmax(-1).cpu() == ref.argmax(-1)).all()))
        print("masked slots ok:", bool((scores.cpu()[~live] < -9000).all()))
        loss = ops.bce_with_mask_loss(scores, batch["targets"].to(dev), batch["train_loss_mask"].to(dev))
        print("loss", loss.item(), "ref", float(g["tf/loss"]))
        loss.backward()
        torch.cuda.synchronize()
        print("fwd+bwd wall %.3fs" % (time.time() - t0))
        grads = dict((n, p.grad) for n, p in model.named_parameters())
        worst = 0
        for k in g.files:
            if k.startswith("grad/"):
                got = grads[k[5:]].detach().cpu()
                if got.numel() > 70000:
                    got = got.flatten()[:: max(1, got.numel() // 4096)]
                e = rel_err(got, g[k]); worst = max(worst, e)
                print("   %-70s %.2e" % (k, e))
        tot = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in model.parameters() if p.grad is not None)).item()
        print("grad norm", tot, "ref", float(g["grad_norm_total"]), "worst grad rel err", worst)
    section("greedy decode (bf16x3)")
    ops.set_precision("bf16x3"); ops.clear_weight_cache()
    model.eval()
    batch = golden_batch(g)
    with torch.no_grad():
        scores = model(batch)["textvqa_scores"]
    print("tokens equal:", np.array_equal(batch["train_prev_inds"].cpu().numpy(), g["greedy/prev_inds"]),
          "scores", rel_err(scores.cpu(), g["greedy/scores"], torch.from_numpy(g["greedy/scores"]) > -5000))
    print(batch["train_prev_inds"].cpu().numpy()[:2])

for f in (t_layernorm, t_linear, t_attention, t_model):
    guarded(f)
print("launches", ops.launch_count)
