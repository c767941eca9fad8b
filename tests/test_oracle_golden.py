"""Pins oracle/oracle_ops.c to the reference's own known-answer tests:
mmdetection3d/tests/test_models/test_common_modules/test_pointnet_ops.py:10-24 (test_fps)
and :27-74 (test_ball_query).  The literal vectors below are those tests' inputs/outputs."""
import torch

from oracle import ops

XYZ_FPS = [[[-0.2748, 1.0020, -1.1674], [0.1015, 1.3952, -1.2681], [-0.8070, 2.4137, -0.5845],
            [-1.0001, 2.1982, -0.5859], [0.3841, 1.8983, -0.7431]],
           [[-1.0696, 3.0758, -0.1899], [-0.2559, 3.5521, -0.1402], [0.8164, 4.0081, -0.1839],
            [-1.1000, 3.0213, -0.8205], [-0.0518, 3.7251, -0.3950]]]

NEW_XYZ = [[[-0.0740, 1.3147, -1.3625], [-2.2769, 2.7817, -0.2334], [-0.4003, 2.4666, -0.5116],
            [-0.0740, 1.3147, -1.3625], [-0.0740, 1.3147, -1.3625]],
           [[-2.0289, 2.4952, -0.1708], [-2.0668, 6.0278, -0.4875], [0.4066, 1.4211, -0.2947],
            [-2.0289, 2.4952, -0.1708], [-2.0289, 2.4952, -0.1708]]]

XYZ = [[[-0.0740, 1.3147, -1.3625], [0.5555, 1.0399, -1.3634], [-0.4003, 2.4666, -0.5116],
        [-0.5251, 2.4379, -0.8466], [-0.9691, 1.1418, -1.3733], [-0.2232, 0.9561, -1.3626],
        [-2.2769, 2.7817, -0.2334], [-0.2822, 1.3192, -1.3645], [0.1533, 1.5024, -1.0432],
        [0.4917, 1.1529, -1.3496]],
       [[-2.0289, 2.4952, -0.1708], [-0.7188, 0.9956, -0.5096], [-2.0668, 6.0278, -0.4875],
        [-1.9304, 3.3092, 0.6610], [0.0949, 1.4332, 0.3140], [-1.2879, 2.0008, -0.7791],
        [-0.7252, 0.9611, -0.6371], [0.4066, 1.4211, -0.2947], [0.3220, 1.4447, 0.3548],
        [-0.9744, 2.3856, -1.2000]]]


def test_fps_known_answer():
    idx = ops.furthest_point_sample(torch.tensor(XYZ_FPS), 3)
    assert idx.tolist() == [[0, 2, 4], [0, 2, 1]]


def test_ball_query_known_answer():
    idx = ops.ball_query(0, 0.2, 5, torch.tensor(XYZ), torch.tensor(NEW_XYZ))
    assert idx.tolist() == [[[0, 0, 0, 0, 0], [6, 6, 6, 6, 6], [2, 2, 2, 2, 2], [0, 0, 0, 0, 0], [0, 0, 0, 0, 0]],
                            [[0, 0, 0, 0, 0], [2, 2, 2, 2, 2], [7, 7, 7, 7, 7], [0, 0, 0, 0, 0], [0, 0, 0, 0, 0]]]


def test_ball_query_dilated_known_answer():
    idx = ops.ball_query(0.2, 0.4, 5, torch.tensor(XYZ), torch.tensor(NEW_XYZ))
    assert idx.tolist() == [[[0, 5, 7, 0, 0], [6, 6, 6, 6, 6], [2, 3, 2, 2, 2], [0, 5, 7, 0, 0], [0, 5, 7, 0, 0]],
                            [[0, 0, 0, 0, 0], [2, 2, 2, 2, 2], [7, 7, 7, 7, 7], [0, 0, 0, 0, 0], [0, 0, 0, 0, 0]]]


def test_fps_tie_rule_on_integer_lattice():
    """Integer voxel coordinates make distance ties ubiquitous.  The kernel's
    winner is: max distance; among equal distances the shared-memory tree
    (furthest_point_sample_cuda.cu:17-23,76-136) keeps the lower slot at every halving step,
    which orders threads by the BIT-REVERSED thread id t = k mod 1024; within a thread the
    strided scan (:56-71) keeps the smaller k.  Checked against an independent statement of
    that rule on a lattice with N > 1024 so the block is 1024 wide."""
    import numpy as np
    g = np.random.default_rng(0)
    pts = np.stack(np.meshgrid(np.arange(14), np.arange(14), np.arange(8), indexing="ij"), -1).reshape(-1, 3)
    pts = pts[g.permutation(len(pts))[:1400]]
    pts = pts[np.lexsort((pts[:, 2], pts[:, 1], pts[:, 0]))].astype(np.int64)
    m = 64
    idx = ops.furthest_point_sample(torch.from_numpy(pts.astype(np.float32))[None], m)[0].numpy()
    n = len(pts)
    temp = np.full(n, 10 ** 10, dtype=np.int64)
    k = np.arange(n)
    old, exp = 0, [0]
    for _ in range(1, m):
        d = ((pts - pts[old]) ** 2).sum(1)
        temp = np.minimum(temp, d)
        cand = np.nonzero(temp == temp.max())[0]
        t = cand % 1024
        brev = np.array([int(format(int(v), "010b")[::-1], 2) for v in t])
        order = np.lexsort((cand, brev))
        old = int(cand[order[0]])
        exp.append(old)
    assert idx.tolist() == exp
