#!/bin/bash
# Standard GPU-side checks with SHORT inner timeouts (a hung multi-rank run once burned a whole round's GPU budget: every
# multi-process command here is capped at 4 minutes and tried at 2 ranks before anything larger).
#   gpurun --timeout 900 -- 'bash tools/gpu_check.sh tests'        # pytest -m gpu + smoke
#   gpurun --timeout 600 -- 'bash tools/gpu_check.sh bench'        # default bench line -> gpurun_out/bench.json
#   gpurun --timeout 600 -- 'bash tools/gpu_check.sh train'        # training step timing + ncu launch list
#   gpurun --gpus 2 --timeout 600 -- 'bash tools/gpu_check.sh scale 2'
set -u
mkdir -p gpurun_out
case "${1:-tests}" in
  tests)
    timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
    timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
    ;;
  bench)
    timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
    python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench.json") if l.startswith("{")][-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "sync", d["e2e"].get("sync_value"),
      "train", (d.get("train") or {}).get("value"), "prep", (d.get("preprocess") or {}).get("value"), "roofline", d["roofline"]["frac"])
PY
    tail -2 gpurun_out/bench.err
    ;;
  train)
    timeout 200 python tools/train_bench.py --dropout 0.0 | tail -1
    timeout 200 python tools/train_bench.py --dropout 0.2 | tail -1
    timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv \
      python tools/train_bench.py --steps 1 --warmup 1 --dropout 0.2 > gpurun_out/ncu_train.log 2>&1
    python tools/summarise_train_launches.py gpurun_out/train_launches.csv | head -25
    ;;
  scale)
    N="${2:-2}"
    timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29541 \
      bench.py --gpus "$N" --steps 20 --warmup 3 > "gpurun_out/bench_${N}gpu.json" 2> "gpurun_out/bench_${N}gpu.err"
    echo "exit $?"; wc -l "gpurun_out/bench_${N}gpu.json"; tail -c 600 "gpurun_out/bench_${N}gpu.json"; tail -3 "gpurun_out/bench_${N}gpu.err"
    ;;
esac
