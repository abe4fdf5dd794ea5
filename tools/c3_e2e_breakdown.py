"""Developer tool: where the e2e time of the C3 generate path goes (model('3dqa', training=False) vs lang_model.generate)."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    layers = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    if layers != 32:
        import navillm_b200.nav_model as nm
        nm.VICUNA_7B["num_hidden_layers"] = layers
    model = bench.build_model(dev).eval()
    model._ensure()
    lm = model.lang_model
    B, NV, NEW = 8, 256, 128
    rng = np.random.RandomState(0)
    prompts = ["Scene " + " ".join(["<cand>"] * NV) + " Question " + " ".join(f"w{i}" for i in rng.randint(0, 5000, size=59)) + " Answer"
               for _ in range(B)]
    g = torch.Generator().manual_seed(0)
    feats = [torch.randn(NV, bench.IMG_FEAT, generator=g).pin_memory() for _ in range(B)]

    def t(fn, n=3):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            r = fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n * 1e3, r

    ms_tok, text = t(lambda: lm.tokenize(prompts))
    print(f"tokenize            {ms_tok:8.2f} ms")
    view = torch.stack([f.to(dev) for f in feats], 0)
    with torch.no_grad():
        ms, pano = t(lambda: model.img_embeddings.forward_panorama_per_step(view_img_fts=view, view_lens=torch.full((B,), NV, device=dev)))
        print(f"panorama encoder    {ms:8.2f} ms  (B={B}, N={NV})")
        ms, cand = t(lambda: model._masked_rows_plus_const(pano["pano_embeds"].reshape(B * NV, -1), np.ones((B, NV), dtype=bool)))
        print(f"rows + const        {ms:8.2f} ms")
    for graph in (True, False):
        ms, out = t(lambda: lm.generate(input_ids=text["input_ids"], attention_mask=text["attention_mask"], cand_vis=cand,
                                        max_new_tokens=NEW, stop_on_eos=False, use_cuda_graph=graph))
        print(f"generate graph={graph!s:5} {ms:8.2f} ms")
    ids = out.tolist()
    ms, _ = t(lambda: lm.tokenizer.batch_decode([s[text["input_ids"].shape[1]:] for s in ids], skip_special_tokens=True))
    print(f"batch_decode        {ms:8.2f} ms")
    b = {"question": [""] * B, "prompts": prompts, "features": [f.to(dev) for f in feats]}
    ms, _ = t(lambda: model("3dqa", dict(b), training=False, max_new_tokens=NEW, do_sample=False, stop_on_eos=False))
    print(f"model('3dqa') total {ms:8.2f} ms")


if __name__ == "__main__":
    main()
