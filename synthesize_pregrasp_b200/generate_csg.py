"""Contact-state-graph node scoring and path selection: the batched counterpart of neurals/scripts/generate_csg.py.

The reference script (generate_csg.py:108-130) scores every ordered pair of contact regions ("node") by averaging the
score network over NUM_SAMPLES sampled fingertip placements, one batch-1 network call per (node, sample) -- 37 820 calls
for the plate's 62 regions.  Here all of them go through ONE `pg_score` launch (the cost kernel's network half, clouds
shared by every evaluation), then paths of K nodes are enumerated and ranked on the host exactly as :136-180 does.

Inputs that the reference computes with Open3D (region sampling, de-occlusion of the initial cloud) are arguments here:
this module starts where the arithmetic starts.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch

NUM_SAMPLES = 10      # generate_csg.py:16
K_PATH = 2            # generate_csg.py:17


def all_ordered_nodes(region_ids: Sequence[int]) -> List[Tuple[int, int]]:
    """generate_csg.py:80-88: every ordered pair of two different regions, in double-loop order."""
    return [(i, j) for i in region_ids for j in region_ids if i != j]


def node_scores(engine, cloud: np.ndarray, df: np.ndarray, sample_points: Sequence[np.ndarray],
                nodes: Sequence[Tuple[int, int]], chunk: int = 8192) -> np.ndarray:
    """generate_csg.py:108-130 (`compute_score`) for all nodes at once.

    cloud [P,3], df [P]: the point cloud and its distance field (the "init" or the "all" pair of :104-107);
    sample_points[r] [S,3]: sampled placements of region r; nodes: (i, j) region pairs.
    -> [len(nodes)] float64, mean over the S samples of pred_score(cloud, [p_i(t), p_j(t)], df).
    """
    cloud = np.asarray(cloud, np.float32)
    df = np.asarray(df, np.float32).reshape(-1)
    P = cloud.shape[0]
    S = len(sample_points[nodes[0][1]])
    grasp = np.empty((len(nodes), S, 6), np.float32)
    for n, (i, j) in enumerate(nodes):
        grasp[n, :, :3] = np.asarray(sample_points[i], np.float32)[:S]
        grasp[n, :, 3:] = np.asarray(sample_points[j], np.float32)[:S]
    grasp = grasp.reshape(-1, 6)
    dev = torch.device("cuda", engine.device)
    out = np.empty(len(grasp), np.float64)
    pts_t = torch.as_tensor(cloud, device=dev)
    df_t = torch.as_tensor(df, device=dev)
    for a in range(0, len(grasp), chunk):
        b = min(a + chunk, len(grasp))
        B = b - a
        logit = engine.score(pts_t.unsqueeze(0).expand(B, P, 3).contiguous(), df_t.unsqueeze(0).expand(B, P).contiguous(),
                             torch.as_tensor(grasp[a:b], device=dev))
        out[a:b] = logit.double().cpu().numpy()
    return out.reshape(len(nodes), S).sum(axis=1) / S


def select_paths(init_nodes: Sequence[Tuple[int, int]], init_scores: Sequence[float], all_nodes: Sequence[Tuple[int, int]],
                 all_scores: Sequence[float], k: int = K_PATH, top: int = 10):
    """generate_csg.py:136-180: depth-first enumeration of k-node paths (consecutive nodes keep one finger's region),
    ascending sort by summed score, the first `top` paths and the node table they index.
    -> (csg_nodes [M,2] int, paths_id [top,k] int, sorted_paths, sorted_scores)."""
    paths, scores = [], []

    def grow(path, score, left):
        if left == 0:
            paths.append(list(path))
            scores.append(score)
            return
        last = path[-1]
        for node, s in zip(all_nodes, all_scores):
            if node[0] == last[0] or node[1] == last[1]:
                grow(path + [node], score + s, left - 1)

    for node, s in zip(init_nodes, init_scores):
        grow([tuple(node)], s, k - 1)
    order = np.argsort(scores)                     # ascending, numpy's default (quicksort) as in the reference
    sorted_paths = [paths[o] for o in order]
    chosen = sorted_paths[:top]
    table = list({tuple(n) for p in chosen for n in p})      # the reference keeps Python's set order: reproduced as is
    ids = np.asarray([[table.index(tuple(n)) for n in p] for p in chosen])
    return np.asarray(table), ids, sorted_paths, np.asarray(scores)[order]
