"""Parity at BASELINE.json's REAL widths, on the code path bench.py runs ("What's weak" #1).

The golden fixtures pin the boundary at toy widths (hidden 256, T <= 367), where ``LlamaCore`` takes the unfused
single-CTA kernels.  The tests here run Vicuna-7B widths (hidden 4096 / 32 heads / F 11008, 36 x 1408 views, panorama
hidden 1024 / 16 heads, vocab 32006) on 2 decoder layers with T >= 1024 packed rows, so the CTA-pair GEMM and its fused
epilogues (``nv_gemm_rope / swiglu / dswiglu / attnd_bf16``), the 128-wide attention kernels and -- for generate -- the
swap-AB skinny GEMMs at N = 4096 / 11008 / 32006 execute inside the stack, and compare with the CPU oracle (which
finishes 2 full-width layers in seconds):

* ``truth`` = oracle in fp32 on the same (bf16-valued) weights, ``ref`` = oracle in the reference's amp_bf16;
  stack rule:     max|cuda - truth| <= 2 max|ref - truth| + 1e-3 max|truth|          (as tests/test_llama_gpu.py)
  boundary rule:  max|cuda - ref|   <= 3 max|ref - truth| + 2e-2 max|ref|            (as tests/test_navmodel_gpu.py)
* greedy token ids: per ROW, bit-exact up to the first step whose oracle top-1/top-2 margin is inside the bf16 noise
  floor (3 bf16 ulp of the top logit); a mismatch at a decided step fails; rows cut short by a near-tie are counted and printed.
"""
import math
import sys
import types
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
bf16 = torch.bfloat16

HID, HEADS, INTER, LAYERS = 4096, 32, 11008, 2


@pytest.fixture(scope="module", autouse=True)
def _cpu_threads():
    from oracle.hostcpu import pick_cpu_threads
    pick_cpu_threads()


class _Spy:
    """Counts calls of the fused-epilogue entry points so the test proves WHICH path ran."""

    NAMES = ("gemm_rope", "gemm_swiglu", "gemm_dswiglu", "gemm_attnd", "gemm_skinny", "gemm_skinny_swiglu")

    def __init__(self, ops):
        self.ops, self.n, self.orig = ops, {k: 0 for k in self.NAMES}, {}

    def __enter__(self):
        for k in self.NAMES:
            f = getattr(self.ops, k)
            self.orig[k] = f

            def wrap(*a, _f=f, _k=k, **kw):
                self.n[_k] += 1
                return _f(*a, **kw)
            setattr(self.ops, k, wrap)
        return self

    def __exit__(self, *a):
        for k, f in self.orig.items():
            setattr(self.ops, k, f)


def _stack_check(name, mine, truth, ref):
    e_ref = (ref - truth).abs().max().item()
    e_mine = (mine - truth).abs().max().item()
    lim = 2 * e_ref + 1e-3 * truth.abs().max().item()
    assert e_mine <= lim, f"{name}: cuda err {e_mine:.4g} > 2*ref err {e_ref:.4g} + 1e-3*{truth.abs().max().item():.3g}"
    return e_mine, e_ref


# ---------------------------------------------------------------------------------------------------------------------
# (i) LlamaCore forward + backward at hidden 4096 / 32 heads / F 11008
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("lens", [[700, 333, 150], [2048]], ids=["c2_ragged_T1183", "c5_dense_S2048"])
def test_llama_stack_fullwidth_fused_path_vs_oracle(cuda_dev, lens):
    from oracle import navillm_oracle as O
    from navillm_b200 import llama, ops
    from tests.test_llama_gpu import pack
    cfg = O.OracleConfig(hidden=HID, n_layers=LAYERS, n_heads=HEADS, inter=INTER, vocab=64, pano_hidden=64, pano_heads=2,
                         pano_inter=64, image_feat_size=32, obj_feat_size=32)
    sd = {k: v for k, v in O.init_state_dict(cfg, seed=3).items() if k.startswith("lang_model.model.")}
    g = torch.Generator().manual_seed(5)
    for k in sd:
        if "layernorm" in k or k.endswith("model.norm.weight"):
            sd[k] = (1 + 0.1 * torch.randn(sd[k].shape, generator=g)).to(bf16)
    dims = llama.LlamaDims(hidden=HID, n_layers=LAYERS, n_heads=HEADS, inter=INTER, vocab=64)
    model = llama.LlamaModelParams(dims)
    model.load_state_dict({k[len("lang_model.model."):]: v for k, v in sd.items()})
    flat = llama.FlatParams(model.flat_order(), cuda_dev)
    core = llama.LlamaCore(dims, model, flat)

    B, S, D = len(lens), max(lens), HID
    emb = (torch.randn(B, S, D, generator=g) * 0.5).to(bf16)
    mask = torch.zeros(B, S, dtype=torch.long)
    for b, L in enumerate(lens):
        mask[b, S - L:] = 1                                     # left padding, like the reference tokenizer
    G = torch.randn(B, S, D, generator=g).to(bf16) * mask[..., None]

    def run_oracle(dtype):
        sdd = {k: v.to(dtype).requires_grad_(True) for k, v in sd.items()}
        e = emb.to(dtype).requires_grad_(True)
        cfg2 = O.OracleConfig(**{**cfg.__dict__, "precision": "fp32" if dtype == torch.float32 else "amp_bf16"})
        h = O.llama_model(sdd, cfg2, e, mask)
        (h.float() * G.float()).sum().backward()
        return h.detach().float(), e.grad.float(), {k: v.grad.float() for k, v in sdd.items() if v.grad is not None}

    h_truth, de_truth, gw_truth = run_oracle(torch.float32)
    h_ref, de_ref, gw_ref = run_oracle(bf16)

    rows, pos, cu, seqlens = pack(emb, mask)
    assert rows.numel() >= 1024, "the fused-epilogue path needs T >= 1024"
    x = emb.view(B * S, D)[rows].to(cuda_dev).contiguous()
    with _Spy(ops) as spy:
        hid, tape = core.forward(x, pos.to(cuda_dev), cu.to(cuda_dev), seqlens)
        hn, rstd = ops.rmsnorm_fwd(hid, model.norm.weight.data, dims.rms_eps)
        dy = G.view(B * S, D)[rows].to(cuda_dev).contiguous()
        dhid = ops.rmsnorm_bwd(hid, model.norm.weight.data, rstd, dy, dw=model.norm.weight.grad)
        dx = core.backward(dhid, tape)
        torch.cuda.synchronize()
    # the bench's code path ran: RoPE / SwiGLU / dSwiGLU / attention-D epilogues in every layer
    assert spy.n["gemm_rope"] == LAYERS and spy.n["gemm_swiglu"] == LAYERS, spy.n
    assert spy.n["gemm_dswiglu"] == LAYERS and spy.n["gemm_attnd"] == LAYERS, spy.n

    m = mask.bool().flatten()
    _stack_check("hidden", hn.float().cpu(), h_truth.view(B * S, D)[m], h_ref.view(B * S, D)[m])
    _stack_check("d_embeds", dx.float().cpu(), de_truth.view(B * S, D)[m], de_ref.view(B * S, D)[m])
    named = dict(model.named_parameters())
    for l in range(LAYERS):
        for nm in ("self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj", "self_attn.o_proj", "mlp.gate_proj", "mlp.up_proj",
                   "mlp.down_proj", "input_layernorm", "post_attention_layernorm"):
            k = f"layers.{l}.{nm}.weight"
            _stack_check(k, named[k].grad.float().cpu(), gw_truth["lang_model.model." + k], gw_ref["lang_model.model." + k])
    _stack_check("norm.weight", named["norm.weight"].grad.float().cpu(), gw_truth["lang_model.model.norm.weight"],
                 gw_ref["lang_model.model.norm.weight"])


# ---------------------------------------------------------------------------------------------------------------------
# (ii) NavModel('panorama') + ('navigation') + action CE + backward at 36 x 1408 views, pano 1024/16, LM 4096 x 2 layers
# ---------------------------------------------------------------------------------------------------------------------
def _full_navmodel(dev, base_vocab, seed=0):
    from navillm_b200.nav_model import NavModel
    from navillm_b200.tokenizer import SyntheticTokenizer
    tok = SyntheticTokenizer(base_vocab=base_vocab)
    args = types.SimpleNamespace(precision="amp_bf16", pretrained_model_name_or_path="vicuna-7b-synthetic", image_feat_size=1408,
                                 angle_feat_size=4, obj_feat_size=768, enable_og=True, fuse_obj=False, feat_dropout=0.4,
                                 resume_from_checkpoint=None, from_scratch=True, device=str(dev), seed=seed)
    mc = types.SimpleNamespace(num_pano_layers=2, tokenizer=tok, llama_config=dict(num_hidden_layers=LAYERS, vocab_size=base_vocab))
    torch.manual_seed(seed)
    model = NavModel(args, None, mc).eval()
    return model, tok


def _oracle_cfg(tok, precision):
    from oracle import navillm_oracle as O
    return O.OracleConfig(hidden=HID, n_layers=LAYERS, n_heads=HEADS, inter=INTER, vocab=len(tok), image_feat_size=1408,
                          obj_feat_size=768, pano_hidden=1024, pano_heads=16, pano_inter=4096, cand_id=tok.special["<cand>"],
                          hist_id=tok.special["<hist>"], obj_id=tok.special["<obj>"],
                          cls_ids=(tok.special["<cls_1>"], tok.special["<cls_2>"]), precision=precision)


def _boundary_check(name, mine, ref, truth, k=3.0, floor=2e-2):
    mine, ref, truth = mine.float().cpu(), ref.float(), truth.float()
    fin = torch.isfinite(ref)
    assert torch.equal(torch.isfinite(mine), fin), f"{name}: -inf pattern differs"
    scale = ref[fin].abs().max().item() + 1e-12
    e_ref = (ref[fin] - truth[fin]).abs().max().item()
    e = (mine[fin] - ref[fin]).abs().max().item()
    assert e <= k * e_ref + floor * scale, f"{name}: |cuda-ref|={e:.4g} > {k}*{e_ref:.4g} + {floor}*{scale:.3g}"


GRAD_KEYS = ["out_head.0.weight", "out_head.0.bias", "lang_model.model.norm.weight",
             "lang_model.model.layers.0.self_attn.q_proj.weight", "lang_model.model.layers.0.self_attn.v_proj.weight",
             "lang_model.model.layers.1.self_attn.o_proj.weight", "lang_model.model.layers.0.mlp.gate_proj.weight",
             "lang_model.model.layers.1.mlp.up_proj.weight", "lang_model.model.layers.1.mlp.down_proj.weight",
             "lang_model.model.layers.1.input_layernorm.weight", "lang_model.model.embed_tokens.weight",
             "token_type_embeddings.weight", "gmap_pos_embeddings.0.weight", "vp_pos_embeddings.1.weight",
             "gmap_step_embeddings.weight", "img_embeddings.mapper.weight", "img_embeddings.img_linear.weight",
             "img_embeddings.pano_encoder.layers.0.self_attn.in_proj_weight", "img_embeddings.pano_encoder.layers.1.linear2.weight"]


def test_navmodel_navigation_fullwidth_vs_oracle(cuda_dev):
    """B = 4 R2R-shaped rows of bench.py's own C2 generator (36 x 1408 views, hist = 8, 24 graph nodes, 15 candidates,
    prompt lengths U{256..1024} => T >= 1024): panorama -> navigation -> CE -> backward, eval mode, default tf32 panorama."""
    import bench
    from oracle import navillm_oracle as O
    from navillm_b200 import ops
    B = 4
    model, tok = _full_navmodel(cuda_dev, base_vocab=4096)
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    host, meta = bench.make_workload(77, B=B)
    assert int(meta["lens"].sum()) >= 1024

    def batches(to):
        pano_in = {"view_img_fts": to(host["view_img_fts"]), "view_lens": meta["view_lens"], "loc_fts": to(host["loc_fts"]),
                   "nav_types": to(host["nav_types"])}
        nav = {"data_type": ["r2r"] * B, "pano_masks": torch.ones((B, bench.N_VIEWS + 1), dtype=torch.bool),
               "vp_pos_fts": to(host["vp_pos_fts"]), "vp_cand_vpids": meta["vp_cand_vpids"],
               "gmap_img_embeds": to(host["gmap_img_embeds"]), "gmap_step_ids": meta["gmap_step_ids"],
               "gmap_pos_fts": to(host["gmap_pos_fts"]), "gmap_masks": meta["gmap_masks"], "gmap_pair_dists": None,
               "gmap_visited_masks": meta["gmap_visited_masks"], "gmap_vpids": meta["gmap_vpids"],
               "hist_vis": [list(to(host["hist_vis"])[b].unbind(0)) for b in range(B)], "prompts": meta["prompts"]}
        return pano_in, nav

    def run_oracle(precision):
        cfg = _oracle_cfg(tok, precision)
        dt = cfg.lm_dtype
        sdd = {k: (v.to(dt) if v.dtype == bf16 else v.clone()).requires_grad_(v.is_floating_point()) for k, v in sd.items()}
        pano_in, nav = batches(lambda v: v)
        pano = O.forward_panorama(sdd, cfg, **pano_in)
        nav["vp_img_embeds"] = torch.cat([torch.zeros_like(pano["pano_embeds"][:, :1]), pano["pano_embeds"]], 1)
        torch.manual_seed(7)
        out = O.forward_navigation(sdd, cfg, nav, tok)
        loss = F.cross_entropy(out["fuse_logits"].float(), meta["target_cols"], reduction="sum") / B
        loss.backward()
        return {"pano_embeds": pano["pano_embeds"].detach(), "fuse_embeds": out["fuse_embeds"], "fuse_logits": out["fuse_logits"].detach(),
                "loss": loss.detach(), "grads": {k: sdd[k].grad.float() for k in GRAD_KEYS}}

    truth, ref = run_oracle("fp32"), run_oracle("amp_bf16")

    pano_in, nav = batches(lambda v: v.to(cuda_dev))
    with _Spy(ops) as spy:
        pano = model("panorama", pano_in)
        nav["vp_img_embeds"] = torch.cat([torch.zeros_like(pano["pano_embeds"][:, :1]), pano["pano_embeds"]], 1)
        torch.manual_seed(7)
        out = model("navigation", nav)
        loss = F.cross_entropy(out["fuse_logits"].float(), meta["target_cols"].to(cuda_dev), reduction="sum") / B
        loss.backward()
        torch.cuda.synchronize()
    # layer 0 is full (fused path); the last layer is pruned to the <cls_1> rows and takes the skinny kernels
    assert spy.n["gemm_rope"] == LAYERS and spy.n["gemm_swiglu"] == LAYERS - 1 and spy.n["gemm_dswiglu"] >= LAYERS - 1, spy.n

    # fp32 parts: tf32 panorama GEMMs (10-bit-mantissa operands) => 4e-3 of max, as in tests/test_navmodel_gpu.py
    for k, mine in (("pano_embeds", pano["pano_embeds"].detach()), ("fuse_embeds", out["fuse_embeds"])):
        r = ref[k]
        assert (mine.cpu() - r).abs().max().item() <= 4e-3 * r.abs().max().item(), k
    _boundary_check("fuse_logits", out["fuse_logits"].detach(), ref["fuse_logits"], truth["fuse_logits"])
    _boundary_check("loss", loss.detach(), ref["loss"], truth["loss"])
    named = dict(model.named_parameters())
    for k in GRAD_KEYS:
        assert named[k].grad is not None, k
        _boundary_check("grad " + k, named[k].grad, ref["grads"][k], truth["grads"][k], k=3.0, floor=5e-2)


def test_navmodel_3dqa_lm_loss_fullwidth_vs_oracle(cuda_dev):
    """The LM-loss path at full width: model('3dqa', training=True) -- panorama encoder over 36 x 1408 views, prompt with 36
    <cand> tokens + question, answer labels, lm_head on the label rows at the real vocabulary (32006 columns: the odd-width
    GEMM, the CE kernel over 32006 classes, the [32006, 4096] lm_head wgrad) -- loss and gradients against the oracle."""
    from oracle import navillm_oracle as O
    B = 4
    model, tok = _full_navmodel(cuda_dev, base_vocab=32000)
    model.train(False)                                   # dropout off; gradients still flow (training=True below)
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    rng = np.random.RandomState(11)
    g = torch.Generator().manual_seed(11)
    feats = [torch.randn(36, 1408, generator=g) for _ in range(B)]
    prompts = ["Scene " + " ".join(["<cand>"] * 36) + " Question " + " ".join(f"w{i}" for i in rng.randint(0, 5000, size=int(rng.randint(200, 330))))
               + " Answer" for _ in range(B)]
    answers = [[" ".join(f"w{i}" for i in rng.randint(0, 5000, size=int(rng.randint(4, 17))))] for _ in range(B)]
    batch = {"question": [""] * B, "prompts": prompts, "answers": answers, "data_type": ["scanqa"] * B}
    keys = ["lang_model.lm_head.weight", "lang_model.model.norm.weight", "lang_model.model.embed_tokens.weight",
            "lang_model.model.layers.0.self_attn.q_proj.weight", "lang_model.model.layers.1.mlp.down_proj.weight",
            "lang_model.model.layers.1.mlp.gate_proj.weight", "img_embeddings.mapper.weight", "img_embeddings.img_linear.weight",
            "vp_pos_embeddings.0.bias", "token_type_embeddings.weight"]

    def run_oracle(precision):
        cfg = _oracle_cfg(tok, precision)
        dt = cfg.lm_dtype
        sdd = {k: (v.to(dt) if v.dtype == bf16 else v.clone()).requires_grad_(v.is_floating_point()) for k, v in sd.items()}
        out = O.forward_3dqa(sdd, cfg, dict(batch, features=feats), tok, tok.eos_token, training=True)
        out["loss"].float().backward()
        return {"loss": out["loss"].detach().float(), "grads": {k: sdd[k].grad.float() for k in keys}}

    truth, ref = run_oracle("fp32"), run_oracle("amp_bf16")
    out = model("3dqa", dict(batch, features=[f.to(cuda_dev) for f in feats]), training=True)
    out.loss.float().backward()
    torch.cuda.synchronize()
    _boundary_check("3dqa loss", out.loss.detach(), ref["loss"], truth["loss"])
    named = dict(model.named_parameters())
    for k in keys:
        assert named[k].grad is not None, k
        _boundary_check("grad " + k, named[k].grad, ref["grads"][k], truth["grads"][k], k=3.0, floor=5e-2)


# ---------------------------------------------------------------------------------------------------------------------
# (iii) greedy generate at C3's shape: B = 8, 256 <cand> tokens, >= 32 new tokens, vocab 32006
# ---------------------------------------------------------------------------------------------------------------------
def compare_greedy_rows(ids, ref_ids, ref_logits, S0, n_new, tag=""):
    """Per-row prefix comparison.  Returns (matched tokens per row, rows cut short by a near-tie)."""
    matched, cut = [], []
    for b in range(ids.shape[0]):
        n = 0
        for t in range(n_new):
            if ids[b, S0 + t] == ref_ids[b, S0 + t]:
                n += 1
                continue
            top2 = torch.topk(ref_logits[t][b], 2).values
            # bf16 noise floor: logits are bf16 numbers (7 mantissa bits); each of the two competing logits may sit one ulp
            # off in either implementation, plus upstream hidden-state noise: 3 ulp of the top logit
            ulp = 2.0 ** (math.floor(math.log2(max(top2[0].abs().item(), 1e-30))) - 7)
            margin, noise = (top2[0] - top2[1]).item(), 3 * ulp
            assert margin <= noise, (f"{tag} row {b} step {t}: token {int(ids[b, S0 + t])} != oracle {int(ref_ids[b, S0 + t])} "
                                     f"with margin {margin:.4g} > bf16 noise {noise:.4g} (3 ulp)")
            cut.append((b, t))
            break                                             # legitimate near-tie: this row diverges from here
        matched.append(n)
    return matched, cut


def test_generate_c3_shape_vs_oracle(cuda_dev):
    from oracle import navillm_oracle as O
    from navillm_b200 import ops
    B, N_CAND, N_NEW = 8, 256, 32
    model, tok = _full_navmodel(cuda_dev, base_vocab=32000)
    assert len(tok) == 32006
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items() if k.startswith("lang_model.")}
    rng = np.random.RandomState(5)
    prompts = []
    for b in range(B):
        words = " ".join(f"w{i}" for i in rng.randint(0, 5000, size=int(rng.randint(30, 61))))
        prompts.append("Scene " + " ".join(["<cand>"] * N_CAND) + " Question " + words + " Answer")
    text = tok(prompts)
    g = torch.Generator().manual_seed(9)
    cand = torch.randn(B * N_CAND, HID, generator=g)
    cfg = _oracle_cfg(tok, "amp_bf16")
    ref_ids, ref_logits = O.greedy_generate(sd, cfg, text["input_ids"], text["attention_mask"], cand_vis=cand,
                                            max_new_tokens=N_NEW, stop_on_eos=False, return_logits=True)
    S0 = text["input_ids"].shape[1]
    for graph in (False, True):
        with _Spy(ops) as spy:
            ids = model.lang_model.generate(input_ids=text["input_ids"], attention_mask=text["attention_mask"],
                                            cand_vis=cand.to(cuda_dev), max_new_tokens=N_NEW, stop_on_eos=False,
                                            use_cuda_graph=graph).cpu()
        assert ids.shape == ref_ids.shape and torch.equal(ids[:, :S0], text["input_ids"])
        assert spy.n["gemm_rope"] == LAYERS, spy.n                         # prefill took the fused CTA-pair path (T >= 1024)
        assert spy.n["gemm_skinny"] > 0 and spy.n["gemm_skinny_swiglu"] > 0, spy.n   # decode took the swap-AB kernels
        matched, cut = compare_greedy_rows(ids, ref_ids, ref_logits, S0, N_NEW, tag=f"graph={graph}")
        print(f"\n[c3 generate, graph={graph}] matched tokens per row {matched} of {N_NEW}; "
              f"{len(cut)} of {B} rows cut short by a bf16 near-tie at (row, step) {cut}")
        assert sum(matched) >= B * N_NEW // 4, f"too few bit-exact tokens before near-ties: {matched}"
