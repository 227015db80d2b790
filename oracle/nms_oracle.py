"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the instance NMS of PointGroup.test.

Allowed importers: tests/ (the product package d3net_b200 never imports this module).

  cross_iou      model/pointgroup.py:577-590: a dense 0/1 mask [nProposal, N] scattered from the rows of
                 proposals_idx (repeated rows collapse), intersection = mask @ mask.T in fp32,
                 npoint = row sums, cross_ious = inter / (npoint_h + npoint_v - inter), every step in fp32,
                 left to right.
  nms_instances  lib/utils/eval.py:75-97 (get_nms_instances): visit proposals by descending score; keep the
                 first unvisited one, drop every later one whose IoU with it exceeds the threshold.
                 The reference orders with np.argsort(-scores) (introsort: ties in unspecified order); this
                 restatement breaks ties by index, the CUDA op does the same, and the golden inputs have
                 distinct scores.

Parity status: pinned by tests/golden/ref_nms.npz -- cross_ious from a torch-CPU execution of the quoted
reference lines, picks from the reference's own get_nms_instances source executed in place
(tests/golden/make_golden.py nms)."""
import numpy as np


def cross_iou(proposals_idx, num_proposals, N):
    pidx = np.asarray(proposals_idx).reshape(-1, 2)
    mask = np.zeros((num_proposals, N), np.float32)
    mask[pidx[:, 0], pidx[:, 1]] = 1.0                                   # :579-580
    inter = (mask.astype(np.float64) @ mask.astype(np.float64).T).astype(np.float32)   # exact integers (:588)
    npoint = mask.sum(1, dtype=np.float32)                                # :589
    h = np.repeat(npoint[:, None], num_proposals, 1)
    v = np.repeat(npoint[None, :], num_proposals, 0)
    with np.errstate(invalid="ignore", divide="ignore"):
        return ((inter / ((h + v).astype(np.float32) - inter).astype(np.float32)).astype(np.float32),
                npoint.astype(np.int32))


def nms_instances(cross_ious, scores, threshold):
    scores = np.asarray(scores, np.float32)
    order = np.argsort(-scores, kind="stable")
    alive = np.ones(len(order), bool)
    pick = []
    for k, i in enumerate(order):
        if not alive[k]:
            continue
        pick.append(i)
        later = order[k + 1:]
        alive[k + 1:] &= ~(cross_ious[i, later] > threshold)
    return np.array(pick, dtype=np.int32)
