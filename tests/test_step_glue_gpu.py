"""Step glue on the device: shared embedding slab + pinned single-upload staging give the reference's values."""
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden" / "step_glue.pt"


def test_step_glue_device_path_matches_reference(cuda_dev):
    from navillm_b200.step_glue import EmbedSlab, GraphMap, StepPacker
    gold = torch.load(GOLD, weights_only=False)
    B = len(gold["start"])
    slab = EmbedSlab(B, gold["D"], cuda_dev, cap=32)
    gmaps = [GraphMap(v, slab, i) for i, v in enumerate(gold["start"])]
    packer = StepPacker(gold["F"], cuda_dev)
    for t, st in enumerate(gold["steps"]):
        obs = st["obs"]
        for i, gm in enumerate(gmaps):
            gm.update_graph(obs[i])
            gm.node_step_ids[obs[i]["viewpoint"]] = t + 1
        pano = packer.panorama_feature_variable_object(obs)
        assert pano["view_img_fts"].is_cuda and not pano["nav_types"].is_cuda
        assert torch.allclose(pano["view_img_fts"].cpu(), st["pano_in"]["view_img_fts"], atol=1e-6, rtol=0)
        pe, pm = st["pano_embeds"].to(cuda_dev), st["pano_masks"].to(cuda_dev)
        avg = (pe * pm.unsqueeze(2)).sum(1) / pm.sum(1, keepdim=True)
        for i, gm in enumerate(gmaps):
            gm.update_node_embed(obs[i]["viewpoint"], avg[i], rewrite=True)
            for j, cv in enumerate(pano["cand_vpids"][i]):
                if not gm.graph.visited(cv):
                    gm.update_node_embed(cv, pe[i, j])
        nav = packer.nav_variables(obs, gmaps, pe, pm, pano["cand_vpids"], pano["nav_types"].to(cuda_dev))
        for k in ("gmap_img_embeds", "gmap_pos_fts", "vp_pos_fts", "vp_img_embeds"):
            assert nav[k].is_cuda
            assert torch.allclose(nav[k].cpu(), st["nav"][k].float(), atol=1e-5, rtol=0), (t, k)
        for k in ("gmap_step_ids", "gmap_visited_masks", "gmap_masks"):
            assert not nav[k].is_cuda and torch.equal(nav[k], st["nav"][k])
    assert packer.h2d_copies == 2 * len(gold["steps"])
