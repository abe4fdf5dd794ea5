"""Host-side step glue of a navigation rollout — the callers' side of the hot path.

Mirrors, with the same names, argument meaning and results, what the reference does per step between the simulator
and ``model('panorama' | 'navigation', ...)``:

    models/graph_utils.py:46-165      FloydGraph, GraphMap (update_graph, update_node_embed, get_node_embed, get_pos_fts)
    tasks/agents/mp3d_agent.py:143-212 panorama_feature_variable_object
    tasks/agents/mp3d_agent.py:264-371 nav_vp_variable, nav_gmap_variable

What is different (the reason this module exists once the GPU step takes a few milliseconds per episode):
  * shortest-path relaxation is one vectorised numpy min-plus update per visited node instead of an O(N^2) Python loop
    over nested defaultdicts; node ids are small integers per graph;
  * node embeddings (running sums + counts) of ALL graphs of a batch live in one device slab; ``gmap_img_embeds`` is one
    gather + divide instead of a Python list of per-node tensors stacked and padded per sample;
  * ``gmap_pair_dists`` - an O(G^2) Python double loop whose result the model never reads (models/nav_model.py:141-143) -
    is not computed (returned as None);
  * every float feature block of a step (view features, location features, position features) is written into ONE
    pinned staging buffer and uploaded with ONE host-to-device copy; masks / step ids / lengths, which NavModel needs on
    the host to build its index maps, stay host tensors (the reference uploads them and NavModel would read them back).
The results are value-identical to the reference's (tests/test_step_glue_cpu.py pins them against golden outputs of the
unmodified reference functions).
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence

import numpy as np
import torch

MAX_DIST = 30          # models/graph_utils.py:5-6
MAX_STEP = 10
_INF = 95959595.0      # the reference's "no edge" sentinel (models/graph_utils.py:48)


def calc_position_distance(a, b) -> float:
    return float(np.sqrt((b[0] - a[0]) ** 2 + (b[1] - a[1]) ** 2 + (b[2] - a[2]) ** 2))


def calculate_vp_rel_pos_fts(a, b, base_heading=0.0, base_elevation=0.0):
    """models/graph_utils.py:18-35 (the simulator's x/y axes are transposed)."""
    dx, dy, dz = b[0] - a[0], b[1] - a[1], b[2] - a[2]
    xy_dist = max(np.sqrt(dx ** 2 + dy ** 2), 1e-8)
    xyz_dist = max(np.sqrt(dx ** 2 + dy ** 2 + dz ** 2), 1e-8)
    heading = np.arcsin(dx / xy_dist)
    if b[1] < a[1]:
        heading = np.pi - heading
    heading -= base_heading
    elevation = np.arcsin(dz / xyz_dist) - base_elevation
    return heading, elevation, xyz_dist


def get_angle_fts(headings, elevations, angle_feat_size: int) -> np.ndarray:
    ang = np.vstack([np.sin(headings), np.cos(headings), np.sin(elevations), np.cos(elevations)]).transpose().astype(np.float32)
    rep = angle_feat_size // 4
    return np.concatenate([ang] * rep, 1) if rep > 1 else ang


class FloydGraph:
    """models/graph_utils.py:46-95 on integer node ids and dense numpy matrices (capacity grows by doubling)."""

    def __init__(self, capacity: int = 64):
        self._ids: Dict[Any, int] = {}
        self._dis = np.full((capacity, capacity), _INF, dtype=np.float64)
        self._point = np.full((capacity, capacity), -1, dtype=np.int64)      # -1 = direct edge / unknown ("" upstream)
        self._has_edge = np.zeros(capacity, dtype=bool)                       # nodes that appear as keys of upstream's _dis
        self._visited = set()

    def _id(self, x) -> int:
        i = self._ids.get(x)
        if i is None:
            i = len(self._ids)
            if i >= self._dis.shape[0]:
                n = self._dis.shape[0] * 2
                d = np.full((n, n), _INF, dtype=np.float64); d[:i, :i] = self._dis[:i, :i]
                p = np.full((n, n), -1, dtype=np.int64); p[:i, :i] = self._point[:i, :i]
                h = np.zeros(n, dtype=bool); h[:i] = self._has_edge[:i]
                self._dis, self._point, self._has_edge = d, p, h
            self._ids[x] = i
        return i

    def distance(self, x, y) -> float:
        if x == y:
            return 0
        i, j = self._ids.get(x), self._ids.get(y)
        return _INF if i is None or j is None else self._dis[i, j]

    def add_edge(self, x, y, dis: float) -> None:
        i, j = self._id(x), self._id(y)
        self._has_edge[i] = self._has_edge[j] = True
        if dis < self._dis[i, j]:
            self._dis[i, j] = self._dis[j, i] = dis
            self._point[i, j] = self._point[j, i] = -1

    def update(self, k) -> None:
        """Relax every pair through k.  Row/column k never change during upstream's sweep (dis[k][k] is the sentinel), so
        one min-plus outer sum is the same computation; ties keep the old entry (strict '<' upstream)."""
        kk = self._id(k)
        n = len(self._ids)
        act = self._has_edge[:n]
        d = self._dis[:n, :n]
        via = d[:, kk][:, None] + d[kk, :][None, :]
        better = via < d
        better &= act[:, None] & act[None, :]
        np.fill_diagonal(better, False)
        d[better] = via[better]
        self._point[:n, :n][better] = kk
        self._visited.add(k)

    def visited(self, k) -> bool:
        return k in self._visited

    def _hops(self, i: int, j: int) -> int:
        if i == j:
            return 0
        k = self._point[i, j]
        return 1 if k < 0 else self._hops(i, k) + self._hops(k, j)

    def path_len(self, x, y) -> int:
        """len(FloydGraph.path(x, y)) of the reference (lazy recursion over the CURRENT split points)."""
        if x == y:
            return 0
        return self._hops(self._id(x), self._id(y))


class EmbedSlab:
    """Running sums + counts of the node embeddings of every graph of a batch: [B * cap, D] on the device."""

    def __init__(self, batch_size: int, dim: int, device, cap: int = 128):
        self.B, self.cap, self.dim = batch_size, cap, dim
        self.sum = torch.zeros((batch_size * cap, dim), dtype=torch.float32, device=device)
        self.cnt = torch.zeros((batch_size * cap,), dtype=torch.float32, device=device)


class GraphMap:
    """models/graph_utils.py:99-165.  Node embeddings are rows of a shared EmbedSlab when one is given (batched device
    path), else per-node tensors like the reference (CPU tests)."""

    def __init__(self, start_vp, slab: Optional[EmbedSlab] = None, row: int = 0):
        self.start_vp = start_vp
        self.node_positions: Dict[Any, Any] = {}
        self.graph = FloydGraph()
        self.node_step_ids: Dict[Any, int] = {}
        self.node_stop_scores: Dict[Any, Any] = {}
        self.node_nav_scores: Dict[Any, Any] = {}
        self.pooling_mode = "mean"
        self._slab, self._row = slab, row
        self._slot: Dict[Any, int] = {}
        self.node_embeds: Dict[Any, list] = {}

    def update_graph(self, ob) -> None:
        self.node_positions[ob["viewpoint"]] = ob["position"]
        for cc in ob["candidate"]:
            self.node_positions[cc["viewpointId"]] = cc["position"]
            self.graph.add_edge(ob["viewpoint"], cc["viewpointId"], calc_position_distance(ob["position"], cc["position"]))
        self.graph.update(ob["viewpoint"])

    # ---- embeddings ----
    def slot(self, vp) -> int:
        s = self._slot.get(vp)
        if s is None:
            s = len(self._slot)
            if self._slab is not None and s >= self._slab.cap:
                raise RuntimeError(f"graph has more than {self._slab.cap} nodes: enlarge EmbedSlab(cap=...)")
            self._slot[vp] = s
        return s

    def update_node_embed(self, vp, embed: torch.Tensor, rewrite: bool = False) -> None:
        if self.pooling_mode != "mean":
            raise NotImplementedError('`pooling_mode` only supports "mean" here (the reference default)')
        if self._slab is not None:
            r = self._row * self._slab.cap + self.slot(vp)
            if rewrite:
                self._slab.sum[r] = embed
                self._slab.cnt[r] = 1
            else:
                self._slab.sum[r] += embed
                self._slab.cnt[r] += 1
            return
        if rewrite or vp not in self.node_embeds:
            self.node_embeds[vp] = [embed.clone() if not rewrite else embed, 1]
        else:
            self.node_embeds[vp][0] = self.node_embeds[vp][0] + embed
            self.node_embeds[vp][1] += 1

    def get_node_embed(self, vp) -> torch.Tensor:
        if self._slab is not None:
            r = self._row * self._slab.cap + self._slot[vp]
            return self._slab.sum[r] / self._slab.cnt[r]
        return self.node_embeds[vp][0] / self.node_embeds[vp][1]

    # ---- position features ----
    def get_pos_fts(self, cur_vp, gmap_vpids: Sequence, cur_heading, cur_elevation, angle_feat_size: int = 4) -> np.ndarray:
        n = len(gmap_vpids)
        rel_angles = np.zeros((n, 2), dtype=np.float64)
        rel_dists = np.zeros((n, 3), dtype=np.float64)
        a = self.node_positions[cur_vp]
        for j, vp in enumerate(gmap_vpids):
            if vp is None:
                continue
            h, e, dist = calculate_vp_rel_pos_fts(a, self.node_positions[vp], cur_heading, cur_elevation)
            rel_angles[j] = (h, e)
            rel_dists[j] = (dist / MAX_DIST, self.graph.distance(cur_vp, vp) / MAX_DIST, self.graph.path_len(cur_vp, vp) / MAX_STEP)
        rel_angles = rel_angles.astype(np.float32)
        rel_dists = rel_dists.astype(np.float32)
        return np.concatenate([get_angle_fts(rel_angles[:, 0], rel_angles[:, 1], angle_feat_size), rel_dists], 1)


class StepPacker:
    """Builds the 'panorama' and 'navigation' batches of one rollout step.  All float blocks go through ONE pinned
    staging buffer and ONE host-to-device copy per call (``h2d_copies`` counts them)."""

    def __init__(self, image_feat_size: int, device, enc_full_graph: bool = True):
        self.image_feat_size, self.device, self.enc_full_graph = image_feat_size, torch.device(device), enc_full_graph
        self._stage: Optional[torch.Tensor] = None
        self.h2d_copies = 0

    def _upload(self, blocks: List[np.ndarray]) -> List[torch.Tensor]:
        total = sum(b.size for b in blocks)
        if self._stage is None or self._stage.numel() < total:
            st = torch.empty(max(total, 1 << 16), dtype=torch.float32)
            self._stage = st.pin_memory() if torch.cuda.is_available() else st
        host, o = self._stage[:total], 0
        hn = host.numpy()
        for b in blocks:
            hn[o:o + b.size] = b.reshape(-1)
            o += b.size
        dev = host.to(self.device, non_blocking=True) if self.device.type == "cuda" else host.clone()
        self.h2d_copies += 1
        outs, o = [], 0
        for b in blocks:
            outs.append(dev[o:o + b.size].view(*b.shape))
            o += b.size
        return outs

    # ---- tasks/agents/mp3d_agent.py:143-212 ----
    def panorama_feature_variable_object(self, obs) -> Dict[str, Any]:
        F = self.image_feat_size
        has_obj = "obj_img_fts" in obs[0]
        B = len(obs)
        per, cand_vpids, lens = [], [], []
        for ob in obs:
            used, img, ang, nav, cv = set(), [], [], [], []
            for cc in ob["candidate"]:
                img.append(cc["feature"][:F]); ang.append(cc["feature"][F:]); nav.append(1)
                cv.append(cc["viewpointId"]); used.add(cc["pointId"])
            for k, x in enumerate(ob["feature"]):
                if k not in used:
                    img.append(x[:F]); ang.append(x[F:])
            nav.extend([0] * (36 - len(used)))
            per.append((np.stack(img, 0), np.stack(ang, 0), nav))
            cand_vpids.append(cv); lens.append(len(img))
        N = max(lens)
        A = per[0][1].shape[1]
        view = np.zeros((B, N, F), dtype=np.float32)
        loc = np.zeros((B, N, A + 3), dtype=np.float32)
        nav_types = torch.zeros((B, max(len(p[2]) for p in per)), dtype=torch.long)
        for i, (img, ang, nav) in enumerate(per):
            view[i, :lens[i]] = img
            loc[i, :lens[i], :A] = ang
            loc[i, :lens[i], A:] = 1.0
            nav_types[i, :len(nav)] = torch.tensor(nav, dtype=torch.long)
        blocks = [view, loc]
        if has_obj:
            O = max(len(ob["obj_img_fts"]) for ob in obs)
            od = obs[0]["obj_img_fts"].shape[1]
            ol = np.concatenate([obs[0]["obj_ang_fts"], obs[0]["obj_box_fts"]], 1).shape[1]
            obj = np.zeros((B, O, od), dtype=np.float32)
            oloc = np.zeros((B, O, ol), dtype=np.float32)
            for i, ob in enumerate(obs):
                n = len(ob["obj_img_fts"])
                obj[i, :n] = ob["obj_img_fts"]
                oloc[i, :n] = np.concatenate([ob["obj_ang_fts"], ob["obj_box_fts"]], 1)
            blocks += [obj, oloc]
        dev = self._upload(blocks)
        ret = {"view_img_fts": dev[0], "loc_fts": dev[1], "nav_types": nav_types, "view_lens": torch.tensor(lens, dtype=torch.long),
               "cand_vpids": cand_vpids}
        if has_obj:
            ret.update({"obj_img_fts": dev[2], "obj_loc_fts": dev[3], "obj_lens": torch.tensor([len(ob["obj_img_fts"]) for ob in obs], dtype=torch.long),
                        "obj_ids": [ob["obj_ids"] for ob in obs]})
        return ret

    # ---- tasks/agents/mp3d_agent.py:264-371 (nav_vp_variable + nav_gmap_variable in one pass, one upload) ----
    def nav_variables(self, obs, gmaps: List[GraphMap], pano_embeds: torch.Tensor, pano_masks: torch.Tensor, cand_vpids,
                      nav_types: torch.Tensor) -> Dict[str, Any]:
        B = len(obs)
        NV1 = pano_embeds.shape[1] + 1
        vp_pos = np.zeros((B, NV1, 14), dtype=np.float32)
        gm_vpids, gm_vis, gm_step, gm_pos, no_left = [], [], [], [], []
        for i, gmap in enumerate(gmaps):
            ob = obs[i]
            cand = gmap.get_pos_fts(ob["viewpoint"], cand_vpids[i], ob["heading"], ob["elevation"])
            start = gmap.get_pos_fts(ob["viewpoint"], [gmap.start_vp], ob["heading"], ob["elevation"])
            vp_pos[i, :, :7] = start
            vp_pos[i, 1:len(cand) + 1, 7:] = cand
            visited = [k for k in gmap.node_positions if gmap.graph.visited(k)]
            unvisited = [k for k in gmap.node_positions if not gmap.graph.visited(k)]
            no_left.append(len(unvisited) == 0)
            if self.enc_full_graph:
                vpids = [None] + visited + unvisited
                vis = [0] + [1] * len(visited) + [0] * len(unvisited)
            else:
                vpids = [None] + unvisited
                vis = [0] * len(vpids)
            gm_vpids.append(vpids); gm_vis.append(vis)
            gm_step.append([gmap.node_step_ids.get(vp, 0) for vp in vpids])
            gm_pos.append(gmap.get_pos_fts(ob["viewpoint"], vpids, ob["heading"], ob["elevation"]))
        lens = [len(v) for v in gm_vpids]
        G = max(lens)
        pos = np.zeros((B, G, 7), dtype=np.float32)
        step_ids = torch.zeros((B, G), dtype=torch.long)
        visited_m = torch.zeros((B, G), dtype=torch.bool)
        masks = torch.zeros((B, G), dtype=torch.bool)
        for i in range(B):
            pos[i, :lens[i]] = gm_pos[i]
            step_ids[i, :lens[i]] = torch.tensor(gm_step[i], dtype=torch.long)
            visited_m[i, :lens[i]] = torch.tensor(gm_vis[i], dtype=torch.bool)
            masks[i, :lens[i]] = True
        vp_pos_d, gm_pos_d = self._upload([vp_pos, pos])
        # node embeddings: one gather over the shared slab (row 0 of every graph = the zero [stop] embedding)
        slab = gmaps[0]._slab
        if slab is not None and all(g._slab is slab for g in gmaps):
            idx = np.zeros((B, G), dtype=np.int64)
            valid = np.zeros((B, G), dtype=np.float32)
            for i, gmap in enumerate(gmaps):
                for j, vp in enumerate(gm_vpids[i]):
                    if vp is not None:
                        idx[i, j] = gmap._row * slab.cap + gmap._slot[vp]
                        valid[i, j] = 1.0
            idx_t = torch.from_numpy(idx.reshape(-1)).to(slab.sum.device)
            val_t = torch.from_numpy(valid.reshape(-1)).to(slab.sum.device)
            emb = slab.sum.index_select(0, idx_t) / slab.cnt.index_select(0, idx_t).clamp_min(1.0).unsqueeze(1)
            gmap_img = (emb * val_t.unsqueeze(1)).view(B, G, slab.dim)
        else:
            D = pano_embeds.shape[2]
            gmap_img = torch.zeros((B, G, D), dtype=pano_embeds.dtype, device=pano_embeds.device)
            for i, gmap in enumerate(gmaps):
                for j, vp in enumerate(gm_vpids[i]):
                    if vp is not None:
                        gmap_img[i, j] = gmap.get_node_embed(vp)
        vp_img = torch.cat([torch.zeros_like(pano_embeds[:, :1]), pano_embeds], 1)
        pm = torch.cat([torch.ones_like(pano_masks[:, :1]), pano_masks], 1)
        nav_m = torch.cat([torch.ones((B, 1), dtype=torch.bool, device=nav_types.device), nav_types == 1], 1)
        return {"vp_img_embeds": vp_img, "pano_masks": pm, "vp_pos_fts": vp_pos_d, "vp_nav_masks": nav_m,
                "vp_cand_vpids": [[None] + list(x) for x in cand_vpids],
                "gmap_vpids": gm_vpids, "gmap_img_embeds": gmap_img, "gmap_step_ids": step_ids, "gmap_pos_fts": gm_pos_d,
                "gmap_visited_masks": visited_m, "gmap_pair_dists": None, "gmap_masks": masks, "no_vp_left": no_left}
