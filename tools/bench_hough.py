"""Dev tool: time sod_hough_vote + sod_affine_verify on the C5 stress shape:
2M ratio-passing matches, 500 objects, 90 % outliers.  CUDA events, device-resident inputs."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "sift-based-od_b200"))
sys.path.insert(0, str(ROOT / "tests"))
import scenes  # noqa: E402
from sod_b200 import _capi  # noqa: E402
from sod_b200 import engine as E  # noqa: E402


def main(n_objects=500, per_object=4000, iters=10):
    d = scenes.make_match_stress(103, n_objects=n_objects, per_object=per_object)
    m = len(d["match_q"])
    sc = E.SceneArrays(d["q_xy"], d["q_angle"], d["q_octave"], d["m_xy"], d["m_angle"], d["m_octave"], d["m_image"],
                       d["img_centroid"], d["img_size"].astype(np.float64),
                       np.array([[d["width"], d["height"]]], np.int32), img_group=d["img_group"],
                       groups_per_frame=n_objects)
    mq = torch.from_numpy(d["match_q"]).cuda()
    mt = torch.from_numpy(d["match_t"]).cuda()
    voter = E.HoughVoter(sc, 15)
    res = voter.vote(mq, mt)
    aff = E.affine_verify(sc, mq, mt, res, 5, 4)
    torch.cuda.synchronize()
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    th, ta = [], []
    _capi.timing_enable(True)
    for _ in range(iters):
        e0, e1, e2 = ev(), ev(), ev()
        e0.record()
        res = voter.vote(mq, mt)
        e1.record()
        aff = E.affine_verify(sc, mq, mt, res, 5, 4, result=aff)
        e2.record()
        torch.cuda.synchronize()
        th.append(e0.elapsed_time(e1))
        ta.append(e1.elapsed_time(e2))
    stages = {k: _capi.timing_read(k) for k in _capi.STAGES}
    _capi.timing_enable(False)
    print("stage medians (ms): " + "  ".join(f"{k} {np.median(v):.3f}" for k, v in stages.items() if v))
    c = res.counters.cpu().numpy()
    a = aff.host(int(c[1]))
    hm, am = float(np.median(th)), float(np.median(ta))
    alg = 108.0 * m
    print(f"matches={m} bins={c[0]} votes={c[1]} valid={a['n_valid']} live={int(a['live'].sum())}")
    print(f"hough_vote: {hm:.3f} ms  -> {m / hm / 1e3:.1f} M matches/s, algorithmic {alg / 1e6:.0f} MB -> {alg / hm / 1e6:.1f} GB/s")
    print(f"affine_verify: {am:.3f} ms")


if __name__ == "__main__":
    main(*(int(x) for x in sys.argv[1:]))
