"""Dev tool (torchrun, one rank per GPU): cost of the database-sharded top-2 exchange alone.

Every rank holds lists [nq,2] as a shard sweep would leave them; the two forms of the exchange
(DetectionPipeline exchange="gather" / "scatter") are timed with CUDA events, max over ranks.
  torchrun --nproc-per-node 8 tools/bench_exchange.py [nq]
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "sift-based-od_b200"))
from sod_b200 import engine as E  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    nq = int(sys.argv[1]) if len(sys.argv) > 1 else 1_280_000
    g = torch.Generator(device=dev).manual_seed(7 + rank)
    rows = 1_000_000 // world
    idx = (torch.randint(0, rows, (nq, 2), device=dev, generator=g, dtype=torch.int32) + rank * rows).contiguous()
    d2 = torch.randint(100_000, 400_000, (nq, 2), device=dev, generator=g, dtype=torch.int32)
    d2[:, 1] += d2[:, 0]                       # strictly ascending per row, as a sweep leaves them
    gi = torch.empty((world, nq, 2), dtype=torch.int32, device=dev)
    gd = torch.empty((world, nq, 2), dtype=torch.int32, device=dev)

    def gather():
        dist.all_gather_into_tensor(gi, idx)
        dist.all_gather_into_tensor(gd, d2)
        return E.merge_top2(gi, gd)

    def scatter():
        return E.exchange_merge_top2(idx, d2, world, lambda o, i: dist.all_to_all_single(o, i),
                                     lambda o, i: dist.all_gather_into_tensor(o, i))

    a, b = gather(), scatter()
    torch.cuda.synchronize()
    same = all(torch.equal(x, y) for x, y in zip(a, b))
    out = {}
    for name, f in (("gather", gather), ("scatter", scatter)):
        for _ in range(5):
            f()
        torch.cuda.synchronize()
        dist.barrier()
        ts = []
        for _ in range(20):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dist.barrier()
            e0.record()
            f()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = torch.tensor([float(np.median(ts))], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[name] = float(t.item())
    if rank == 0:
        print(f"exchange of {nq} query rows over {world} GPUs: " +
              "  ".join(f"{k} {v:.3f} ms" for k, v in out.items()) + f"  identical={same}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
