"""Checkpoint and feature wire formats around the hot path.

* ``save_checkpoint`` / ``check_checkpoint``: the reference's ``torch.save`` layout and resume rules
  (tools/optims.py:7-33, 65-78): ``{"model_state_dict": ..., ["optimizer": ..., "epoch": ...]}`` with HF parameter names,
  ``module.`` prefixes stripped on load, shape-mismatched / unknown keys skipped and logged, ``strict=False`` - so released
  zd11024/NaviLLM checkpoints load into ``navillm_b200.NavModel`` and checkpoints written here load into the reference.
  The optimizer entry uses torch.optim.AdamW's state-dict layout (``FlatAdamW.state_dict`` emits per-parameter views of
  its flat moment buffers in ``model.parameters()`` order).
* ``FeatureStore``: a memory-mapped replacement for the per-access ``h5py.File`` open + fp64->fp32 cast of
  ``ImageFeaturesDB.get_image_feature`` (tasks/feature_db.py:18-31), which becomes the input bottleneck once a navigation
  step takes milliseconds: one contiguous fp32 (opt-in fp16: values rounded to half precision) block file + a JSON index; ``get_image_feature(scan, viewpoint)``
  returns the same ``float32 [n_views, image_feat_size]`` slice.  ``FeatureStore.convert_hdf5`` builds it from the
  reference's HDF5 files when h5py is available (it is not in the build image; ``FeatureStore.build`` takes any
  ``(key, array)`` iterable)."""
from __future__ import annotations

import json
import os
from pathlib import Path
from typing import Dict, Iterable, Optional, Tuple

import numpy as np
import torch


def save_checkpoint(model, model_path, optimizer=None, epoch: int = 0, save_states: bool = False) -> None:
    """tools/optims.py:65-78."""
    if hasattr(model, "module"):
        model = model.module
    state = {"model_state_dict": model.state_dict()}
    if save_states:
        state.update({"optimizer": optimizer.state_dict(), "epoch": epoch})
    torch.save(state, model_path)


def check_checkpoint(args, model, optimizer=None, lr_scheduler=None, logger=None) -> int:
    """tools/optims.py:7-33: returns the epoch to resume from (0 when nothing is loaded)."""
    resume_from_epoch = 0
    path = getattr(args, "resume_from_checkpoint", None)
    if path is None:
        return 0
    log = (lambda m: logger.info(m)) if logger is not None else (lambda m: None)
    if getattr(args, "rank", 0) == 0:
        log(f"Loading checkpoint from {path}")
    checkpoint = torch.load(path, map_location="cpu", weights_only=False)
    target = model.module if hasattr(model, "module") else model
    own = target.state_dict()
    on_disk = {k.replace("module.", ""): v for k, v in checkpoint["model_state_dict"].items()}
    update = {}
    for key, val in on_disk.items():
        if key in own and own[key].shape == val.shape:
            update[key] = val
        else:
            log("Ignore weight %s: %s" % (key, str(tuple(val.shape))))
    log(str(target.load_state_dict(update, strict=False)))
    if "epoch" in checkpoint:
        resume_from_epoch = checkpoint["epoch"] + 1
        log("Resume from Epoch {}".format(resume_from_epoch))
        if optimizer is not None:
            optimizer.load_state_dict(checkpoint["optimizer"])
    return resume_from_epoch


class FeatureStore:
    """Memory-mapped ``(n_views, dim)`` feature blocks keyed like the reference's HDF5 datasets (``scan_viewpoint`` or
    ``scan``).  File layout: ``<path>.bin`` = all blocks back to back in ``dtype``; ``<path>.json`` = {"dtype", "dim",
    "index": {key: [row_offset, n_rows]}}."""

    def __init__(self, path: str, image_feat_size: int):
        self.path, self.image_feat_size = str(path), image_feat_size
        meta = json.loads(Path(self.path + ".json").read_text())
        self.dim, self.dtype, self.index = int(meta["dim"]), np.dtype(meta["dtype"]), meta["index"]
        if image_feat_size > self.dim:
            raise ValueError(f"image_feat_size {image_feat_size} exceeds the stored feature width {self.dim}")
        rows = os.path.getsize(self.path + ".bin") // (self.dim * self.dtype.itemsize)
        self._mm = np.memmap(self.path + ".bin", dtype=self.dtype, mode="r", shape=(rows, self.dim))
        self._feature_store: Dict[str, np.ndarray] = {}

    def get_image_feature(self, scan: str, viewpoint: Optional[str] = None, load_in_memory: bool = False) -> np.ndarray:
        """tasks/feature_db.py:18-31: float32 [n_views, image_feat_size] (or [image_feat_size] for 1-row entries stored
        from 1-D datasets)."""
        key = "%s_%s" % (scan, viewpoint) if viewpoint is not None else scan
        if key in self._feature_store:
            return self._feature_store[key]
        off, n, one_d = self.index[key]
        ft = np.asarray(self._mm[off:off + n, :self.image_feat_size], dtype=np.float32)
        if one_d:
            ft = ft[0]
        if load_in_memory:
            self._feature_store[key] = ft
        return ft

    @staticmethod
    def build(path: str, items: Iterable[Tuple[str, np.ndarray]], dtype: str = "float32") -> "None":
        index, off, dim = {}, 0, None
        with open(str(path) + ".bin", "wb") as f:
            for key, arr in items:
                a = np.asarray(arr)
                one_d = a.ndim == 1
                a = a.reshape(1, -1) if one_d else a
                if dim is None:
                    dim = a.shape[1]
                if a.shape[1] != dim:
                    raise ValueError(f"{key}: feature width {a.shape[1]} differs from {dim}")
                f.write(np.ascontiguousarray(a, dtype=dtype).tobytes())
                index[key] = [off, int(a.shape[0]), bool(one_d)]
                off += int(a.shape[0])
        Path(str(path) + ".json").write_text(json.dumps({"dtype": dtype, "dim": dim, "index": index}))

    @staticmethod
    def convert_hdf5(h5_path: str, out_path: str, dtype: str = "float32") -> None:
        try:
            import h5py
        except ImportError as e:  # pragma: no cover - h5py is not part of the build image
            raise RuntimeError("FeatureStore.convert_hdf5 needs h5py (pip install h5py) to read the reference's feature files") from e
        with h5py.File(h5_path, "r") as f:
            FeatureStore.build(out_path, ((k, f[k][...]) for k in f.keys()), dtype=dtype)
