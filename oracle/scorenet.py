"""LargeScoreFunction forward in numpy (test oracle, see oracle/__init__.py).

Restates neurals/network.py:386-412 (PCN.forward) and :434-455 (LargeScoreFunction.pred_score):
  x[P,4]=cat(xyz, df) -> conv1(4->32)+ReLU -> conv2(32->64) -> g=max_P -> cat(x, g)[P,128]
  -> conv3(128->256)+ReLU -> conv4(256->20) -> max_P -> [20]; hstack grasp[6] -> fc1 512 ReLU
  -> fc2 512 ReLU -> fc3 256 ReLU -> out 1 (the LOGIT is the score; 100x logit is the reward,
  envs/plate_contact_graph_env.py:611-612).
Weights are the fp32 tensors of only_score_model_with_df/2980.pth; arithmetic here is fp64 on
those fp32 values, i.e. the exact value an fp32 implementation approximates.
"""
import numpy as np


def score_logit(w, pts, df, grasp, dtype=np.float64):
    """pts [P,3], df [P], grasp [6] -> scalar logit.  P may be any size >= 1."""
    f = lambda k: np.asarray(w[k], dtype)
    x = np.concatenate([np.asarray(pts, np.float32), np.asarray(df, np.float32).reshape(-1, 1)], 1).astype(dtype)
    h1 = np.maximum(x @ f("pcn_conv1_weight").T + f("pcn_conv1_bias"), 0)
    h2 = h1 @ f("pcn_conv2_weight").T + f("pcn_conv2_bias")
    g = h2.max(0)
    cat = np.concatenate([h2, np.broadcast_to(g, h2.shape)], 1)
    h3 = np.maximum(cat @ f("pcn_conv3_weight").T + f("pcn_conv3_bias"), 0)
    h4 = h3 @ f("pcn_conv4_weight").T + f("pcn_conv4_bias")
    lat = np.concatenate([h4.max(0), np.asarray(grasp, np.float32).astype(dtype)])
    z = np.maximum(f("fc1_weight") @ lat + f("fc1_bias"), 0)
    z = np.maximum(f("fc2_weight") @ z + f("fc2_bias"), 0)
    z = np.maximum(f("fc3_weight") @ z + f("fc3_bias"), 0)
    return float((f("out_weight") @ z + f("out_bias"))[0])
