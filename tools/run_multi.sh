#!/usr/bin/env bash
# Multi-GPU session on one box: dist_check (sharded == single GPU), the default bench line, and variants of
# the database-sharded step.  Usage (under gpurun --gpus N): bash tools/run_multi.sh N [tag]
set -u
cd "$(dirname "$0")/.."
N="${1:-8}"; TAG="${2:-r02}"; VARIANTS="${3:-yes}"
OUT=gpurun_out; mkdir -p $OUT
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port "$1" "${@:2}"; }
nvidia-smi --query-gpu=index,name,clocks.max.sm,power.limit --format=csv,noheader > $OUT/${TAG}_n${N}_smi.txt
(timeout 300 bash -c "$(declare -f run); N=$N run 29611 tests/dist_check.py" > $OUT/${TAG}_dist_check_n${N}.log 2>&1; echo "exit $?" >> $OUT/${TAG}_dist_check_n${N}.log)
(NCCL_DEBUG=WARN timeout 400 bash -c "$(declare -f run); N=$N run 29612 bench.py --gpus $N --steps 10 --warmup 3" > $OUT/${TAG}_bench_n${N}.json 2> $OUT/${TAG}_bench_n${N}.err; echo "exit $?" >> $OUT/${TAG}_bench_n${N}.err)
i=0
for variant in "--seed-rows 0" "--sweep-stages 1" "--exchange gather"; do
  [ "$VARIANTS" = "yes" ] || break
  i=$((i+1))
  (timeout 300 bash -c "$(declare -f run); N=$N run $((29620+i)) bench.py --gpus $N --steps 10 --warmup 3 --no-alt --no-configs --no-parity-check $variant" > $OUT/${TAG}_bench_n${N}_v$i.json 2> $OUT/${TAG}_bench_n${N}_v$i.err; echo "exit $? ($variant)" >> $OUT/${TAG}_bench_n${N}_v$i.err)
done
tail -n 3 $OUT/${TAG}_dist_check_n${N}.log
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/${TAG}_bench_n${N}*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "value %.3e" % d["value"], "ms %.3f" % d["ms_per_step"], "kernel %.3f" % d["roofline"]["kernel_ms"],
              "seed %.3f" % d["roofline"]["seed_sweep_ms"], "e2e %.3e" % d["e2e"]["value"], d.get("parity_check"),
              {k: round(v.get("ms", 0), 3) for k, v in d.get("configs", {}).items() if isinstance(v, dict)},
              (d.get("alt_partition") or {}).get("value"))
    except Exception as e:
        print(f, "unreadable", e)
PY
