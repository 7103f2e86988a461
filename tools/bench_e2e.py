"""Dev tool: where the end-to-end step spends its time beyond the kernels (host<->device legs)."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "sift-based-od_b200"))
import bench  # noqa: E402
from sod_b200.pipeline import DetectionPipeline, ModelDatabase  # noqa: E402


def main():
    args = type("A", (), dict(objects=1000, kp_per_object=1000, frames=256, per_frame=5000, instances=4,
                              inlier_frac=0.10, false_frac=0.01))()
    dev = torch.device("cuda")
    wl = bench.make_workload(args, dev)
    db = bench.make_database(wl)
    nq = args.frames * args.per_frame
    host = tuple(torch.as_tensor(wl[k]).cpu().pin_memory() for k in ("q_des", "q_xy", "q_angle", "q_octave", "q_frame"))
    pipe = DetectionPipeline(db, nq, wl["frame_wh"], device=dev)
    sync = torch.cuda.synchronize

    def timed(f, reps=5):
        ts = []
        for _ in range(reps):
            sync(); t = time.perf_counter(); r = f(); sync(); ts.append((time.perf_counter() - t) * 1e3)
        return float(np.median(ts)), r

    for _ in range(2):
        pipe.detect(*host)
    t_load, n = timed(lambda: pipe.load_queries(*host))
    t_dev, r = timed(lambda: pipe.detect_device(n))
    t_fetch, out = timed(lambda: pipe.fetch(r))
    t_detect, _ = timed(lambda: pipe.detect(*host))
    for _ in pipe.detect_batches(host for _ in range(3)):   # first use allocates the second buffer set
        pass
    sync(); t0 = time.perf_counter(); k = 0
    for _ in pipe.detect_batches(host for _ in range(6)):
        k += 1
    sync(); t_pipe = (time.perf_counter() - t0) * 1e3 / k
    print(f"load_queries {t_load:.2f} ms | detect_device {t_dev:.2f} ms | fetch {t_fetch:.2f} ms | "
          f"detect (serial) {t_detect:.2f} ms | detect_batches {t_pipe:.2f} ms/step")
    for name, v in out.items():
        if isinstance(v, np.ndarray):
            print(f"   fetched {name}: {v.nbytes / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
