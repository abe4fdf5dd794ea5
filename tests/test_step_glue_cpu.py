"""Host-side step glue (navillm_b200/step_glue.py) against golden outputs of the unmodified reference
functions (tests/golden/make_glue_golden.py -> step_glue.pt): GraphMap / FloydGraph and the three per-step packing
functions of MP3DAgent, replayed over a 3-episode, 5-step synthetic rollout.  Values must be identical (same float32
formulas); only device placement and the dead gmap_pair_dists differ by design."""
from pathlib import Path

import numpy as np
import torch

GOLD = Path(__file__).parent / "golden" / "step_glue.pt"


def _eq(a, b, name):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    assert a.shape == b.shape, (name, a.shape, b.shape)
    if a.dtype.is_floating_point:
        assert torch.allclose(a.float(), b.float(), rtol=0, atol=1e-6), (name, (a.float() - b.float()).abs().max().item())
    else:
        assert torch.equal(a.to(b.dtype), b), name


def _replay(use_slab):
    from navillm_b200.step_glue import EmbedSlab, GraphMap, StepPacker
    gold = torch.load(GOLD, weights_only=False)
    B = len(gold["start"])
    slab = EmbedSlab(B, gold["D"], "cpu", cap=32) if use_slab else None
    gmaps = [GraphMap(v, slab, i) for i, v in enumerate(gold["start"])]
    packer = StepPacker(gold["F"], "cpu", enc_full_graph=True)
    for t, st in enumerate(gold["steps"]):
        obs = st["obs"]
        for i, gm in enumerate(gmaps):
            gm.update_graph(obs[i])
            gm.node_step_ids[obs[i]["viewpoint"]] = t + 1
        pano = packer.panorama_feature_variable_object(obs)
        ref = st["pano_in"]
        for k in ("view_img_fts", "loc_fts", "nav_types", "view_lens"):
            _eq(pano[k], ref[k], f"step {t} {k}")
        assert pano["cand_vpids"] == ref["cand_vpids"]
        if "obj_img_fts" in ref:
            for k in ("obj_img_fts", "obj_loc_fts", "obj_lens"):
                _eq(pano[k], ref[k], f"step {t} {k}")
            assert pano["obj_ids"] == ref["obj_ids"]
        pe, pm = st["pano_embeds"], st["pano_masks"]
        avg = (pe * pm.unsqueeze(2)).sum(1) / pm.sum(1, keepdim=True)
        for i, gm in enumerate(gmaps):
            gm.update_node_embed(obs[i]["viewpoint"], avg[i].clone(), rewrite=True)
            for j, cv in enumerate(pano["cand_vpids"][i]):
                if not gm.graph.visited(cv):
                    gm.update_node_embed(cv, pe[i, j].clone())
        nav = packer.nav_variables(obs, gmaps, pe, pm, pano["cand_vpids"], pano["nav_types"])
        rn = st["nav"]
        assert nav["gmap_vpids"] == rn["gmap_vpids"] and nav["vp_cand_vpids"] == rn["vp_cand_vpids"]
        assert nav["no_vp_left"] == rn["no_vp_left"]
        for k in ("gmap_img_embeds", "gmap_step_ids", "gmap_pos_fts", "gmap_visited_masks", "gmap_masks", "vp_img_embeds",
                  "pano_masks", "vp_pos_fts", "vp_nav_masks"):
            _eq(nav[k], rn[k], f"step {t} {k}")
        assert nav["gmap_pair_dists"] is None            # dead work in the reference (models/nav_model.py:141-143)
    return packer, gmaps, gold


def test_step_glue_matches_reference_per_node_embeds():
    _replay(use_slab=False)


def test_step_glue_matches_reference_with_shared_slab_and_single_upload():
    packer, gmaps, gold = _replay(use_slab=True)
    # one staging upload per packing call: 2 per step (panorama + navigation)
    assert packer.h2d_copies == 2 * len(gold["steps"])


def test_floyd_graph_matches_reference_semantics_on_random_graphs():
    """Independent of the golden file: the vectorised relaxation equals the reference's nested-dict sweep (restated
    here in 12 lines) on random edge insertions, including distances, hop counts and the strict '<' tie rule."""
    from collections import defaultdict
    from navillm_b200.step_glue import FloydGraph
    rng = np.random.RandomState(0)
    for trial in range(20):
        n = int(rng.randint(4, 12))
        names = [f"n{i}" for i in range(n)]
        dis = defaultdict(lambda: defaultdict(lambda: 95959595))
        point = defaultdict(lambda: defaultdict(lambda: ""))
        fg = FloydGraph(capacity=2)

        def path(x, y):
            if x == y:
                return []
            if point[x][y] == "":
                return [y]
            k = point[x][y]
            return path(x, k) + path(k, y)
        for _ in range(int(rng.randint(n, 3 * n))):
            x, y = rng.choice(n, 2, replace=False)
            w = float(rng.choice([1.0, 2.0, 2.5, 3.0]))          # repeated weights exercise ties
            x, y = names[x], names[y]
            if w < dis[x][y]:
                dis[x][y] = dis[y][x] = w; point[x][y] = point[y][x] = ""
            fg.add_edge(x, y, w)
            if rng.rand() < 0.5:
                k = x
                for a in list(dis):
                    for b in list(dis):
                        if a != b and dis[a][k] + dis[k][b] < dis[a][b]:
                            dis[a][b] = dis[b][a] = dis[a][k] + dis[k][b]
                            point[a][b] = point[b][a] = k
                fg.update(k)
        for a in names:
            for b in names:
                if a in fg._ids and b in fg._ids and a != b:
                    assert fg.distance(a, b) == dis[a][b], (trial, a, b)
                    assert fg.path_len(a, b) == len(path(a, b)), (trial, a, b)
