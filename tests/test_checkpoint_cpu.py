"""Checkpoint / feature wire formats (navillm_b200/checkpoint.py) - host logic, no GPU.

The reference checkpoint is the golden state dict of the unmodified (stubbed) reference NavModel
(tests/golden/nav_fp32.pt, written by make_golden.py): its key set and shapes are what released checkpoints have."""
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
GOLD = Path(__file__).resolve().parent / "golden"


def _model(g):
    from navillm_b200.nav_model import NavModel
    from navillm_b200.tokenizer import SyntheticTokenizer
    d = g["meta"]["dims"]
    tok = SyntheticTokenizer(base_vocab=d["base_vocab"])
    args = types.SimpleNamespace(precision="amp_bf16", pretrained_model_name_or_path="vicuna-tiny", image_feat_size=d["image_feat_size"],
                                 angle_feat_size=4, obj_feat_size=d["obj_feat_size"], enable_og=True, fuse_obj=False, feat_dropout=0.4,
                                 resume_from_checkpoint=None, from_scratch=True)
    mc = types.SimpleNamespace(num_pano_layers=2, tokenizer=tok,
                               llama_config=dict(hidden_size=d["hidden"], intermediate_size=d["inter"], num_hidden_layers=d["n_layers"],
                                                 num_attention_heads=d["n_heads"], vocab_size=d["base_vocab"]),
                               vis_config=dict(hidden_size=d["pano_hidden"], num_attention_heads=d["pano_heads"],
                                               intermediate_size=d["pano_inter"]))
    return NavModel(args, None, mc), args


def test_reference_layout_checkpoint_round_trip(tmp_path):
    from navillm_b200.checkpoint import check_checkpoint, save_checkpoint
    g = torch.load(GOLD / "nav_fp32.pt", weights_only=False)
    ref_sd = g["state_dict"]
    # a checkpoint as the reference's save_checkpoint writes it from a DDP-wrapped model, plus one stale tensor
    disk = {"model_state_dict": {"module." + k: v for k, v in ref_sd.items()}, "epoch": 3, "optimizer": {"state": {}, "param_groups": [{}]}}
    disk["model_state_dict"]["module.lang_model.lm_head.weight"] = torch.zeros(5, 7)        # wrong shape -> ignored
    disk["model_state_dict"]["module.not_a_parameter"] = torch.zeros(1)
    p = tmp_path / "ref.pt"
    torch.save(disk, p)
    model, args = _model(g)
    before = model.state_dict()["lang_model.lm_head.weight"].clone()
    args.resume_from_checkpoint = str(p)
    logs = []
    logger = types.SimpleNamespace(info=logs.append)
    opt = types.SimpleNamespace(load_state_dict=lambda sd: logs.append("opt loaded"))
    assert check_checkpoint(args, model, opt, None, logger) == 4
    sd = model.state_dict()
    assert set(sd) == set(ref_sd)                                                           # same HF parameter names
    for k, v in ref_sd.items():
        if k == "lang_model.lm_head.weight":
            assert torch.equal(sd[k], before)                                               # shape mismatch: kept
        else:
            assert torch.equal(sd[k].float(), v.float().to(sd[k].dtype).float()), k
    assert any("Ignore weight lang_model.lm_head.weight" in str(m) for m in logs)
    assert any("Ignore weight not_a_parameter" in str(m) for m in logs) and "opt loaded" in logs
    # and back: what we write is what the reference's check_checkpoint reads (key names, shapes, layout)
    q = tmp_path / "ours.pt"
    save_checkpoint(types.SimpleNamespace(module=model), q, optimizer=types.SimpleNamespace(state_dict=lambda: {"state": {}}), epoch=7,
                    save_states=True)
    back = torch.load(q, weights_only=False)
    assert set(back) == {"model_state_dict", "optimizer", "epoch"} and back["epoch"] == 7
    assert set(back["model_state_dict"]) == set(ref_sd)
    for k, v in ref_sd.items():
        assert back["model_state_dict"][k].shape == v.shape, k


def test_feature_store_matches_hdf5_reader_semantics(tmp_path):
    from navillm_b200.checkpoint import FeatureStore
    rng = np.random.RandomState(0)
    blocks = {"scanA_vp1": rng.randn(36, 24), "scanA_vp2": rng.randn(36, 24), "scanB": rng.randn(24)}   # fp64 like the HDF5 files
    FeatureStore.build(tmp_path / "fts", blocks.items(), dtype="float32")
    db = FeatureStore(str(tmp_path / "fts"), image_feat_size=16)
    for (scan, vp), key in ((("scanA", "vp1"), "scanA_vp1"), (("scanA", "vp2"), "scanA_vp2")):
        ft = db.get_image_feature(scan, vp)
        assert ft.dtype == np.float32 and ft.shape == (36, 16)
        assert np.array_equal(ft, blocks[key][:, :16].astype(np.float32))                  # tasks/feature_db.py:26-29
    one = db.get_image_feature("scanB")
    assert one.shape == (16,) and np.array_equal(one, blocks["scanB"][:16].astype(np.float32))
    cached = db.get_image_feature("scanA", "vp1", load_in_memory=True)
    assert db.get_image_feature("scanA", "vp1") is cached
    # the default storage dtype is float32: exactly the reference reader's cast
    FeatureStore.build(tmp_path / "d", blocks.items())
    d = FeatureStore(str(tmp_path / "d"), image_feat_size=16).get_image_feature("scanA", "vp2")
    assert np.array_equal(d, blocks["scanA_vp2"][:, :16].astype(np.float32))
    # opt-in fp16 storage (36 x 1408 blocks halve to 99 KB): values within half precision of the fp32 cast
    FeatureStore.build(tmp_path / "h", blocks.items(), dtype="float16")
    h = FeatureStore(str(tmp_path / "h"), image_feat_size=16).get_image_feature("scanA", "vp2")
    assert np.allclose(h, blocks["scanA_vp2"][:, :16], atol=2e-3, rtol=1e-3)
