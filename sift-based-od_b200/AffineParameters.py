"""Affine least-squares helpers, drop-in for AffineParameters.py:11-160.

Gen_A / Gen_b / Ext_Params only lay data out (host).  The fit and the residual test run in
csrc/sod_affine.cu: AffineParameters(posebin) is one fit pass, remove_outliers one residual pass,
Main.apply_affine_parameters the whole fixed-point iteration in a single launch.
"""
import numpy as np

from sod_b200 import dropin as _dropin


def Gen_A(x, y, votes):
    """2*votes x 6 design matrix: [x, y, 0, 0, 1, 0] and [0, 0, x, y, 0, 1] per model point."""
    A = []
    for k in range(votes):
        A.append([x[k], y[k], 0, 0, 1, 0])
        A.append([0, 0, x[k], y[k], 0, 1])
    return A


def Gen_b(b_x, b_y, votes):
    """Right-hand side: the image points interleaved u0, v0, u1, v1, ..."""
    b = []
    for k in range(votes):
        b.append(b_x[k])
        b.append(b_y[k])
    return b


class _Pt:
    __slots__ = ("pt",)

    def __init__(self, x, y):
        self.pt = (x, y)


class _Bin:
    def __init__(self, pairs):
        self.keypoint_pairs = pairs
        self.pose = (0, 0, 0, 0)


def Calc_x(A, b):
    """pinv(A^T A) A^T b for a design matrix built by Gen_A (the only shape the path produces)."""
    A = np.asarray(A, dtype=np.float64).reshape(-1, 6)
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    even, odd = A[0::2], A[1::2]
    ok = (len(A) % 2 == 0 and len(b) == len(A) and np.all(even[:, [2, 3, 5]] == 0) and np.all(even[:, 4] == 1)
          and np.all(odd[:, [0, 1, 4]] == 0) and np.all(odd[:, 5] == 1) and np.all(even[:, :2] == odd[:, 2:4]))
    if not ok:
        raise ValueError("Calc_x expects the matrix layout produced by Gen_A")
    pairs = [(_Pt(x, y), _Pt(u, v)) for (x, y), u, v in zip(even[:, :2], b[0::2], b[1::2])]
    (params, _, _, _), = _dropin.affine_run([_Bin(pairs)], (1, 1), 0.0, 0.0, 0, 1)
    return params


def Ext_Params(x):
    return x[0], x[1], x[2], x[3], x[4], x[5]


def AffineParameters(posebin):
    """Fit [m1, m2, m3, m4, tx, ty] to the bin's first `votes` pairs and store it on the bin
    (reference :89-113).  An empty bin keeps its previous parameters."""
    if posebin.votes <= 0:
        return
    view = _Bin(posebin.keypoint_pairs[:posebin.votes])
    (params, _, _, _), = _dropin.affine_run([view], (1, 1), 0.0, 0.0, 0, 1)
    posebin.affine_parameters = [params[0], params[1], params[2], params[3], params[4], params[5]]


def remove_outliers(posebin, image_query_size, x_factor=8, y_factor=8):
    """Drop the pairs whose residual under posebin.affine_parameters exceeds
    W * pose[3] / x_factor or H * pose[3] / y_factor (pose[3] is the sigma bin index, as in the
    reference :116-160).  Returns (posebin, changed)."""
    x_ref = image_query_size[0] * posebin.pose[3] / x_factor
    y_ref = image_query_size[1] * posebin.pose[3] / y_factor
    pairs = posebin.keypoint_pairs
    keep = _dropin.residual_keep([p[0].pt for p in pairs], [p[1].pt for p in pairs],
                                 posebin.affine_parameters, x_ref, y_ref)
    changed = not bool(keep.all())
    posebin.keypoint_pairs = [p for p, k in zip(pairs, keep) if k]
    posebin.votes = len(posebin.keypoint_pairs)
    return posebin, changed
