#!/bin/bash
# Round-2 GPU checks; every step has its own short timeout. usage: bash tools/gpu_r2.sh <stage> [tag]
set -u
mkdir -p gpurun_out
TAG="${2:-r2}"
kern_table() {
python - "$1" <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "launches", d.get("gpu_launches"))
for k in d.get("kernels", []):
    print("  %-22s n=%2d share=%.3f us=%7.1f tf=%6.1f frac_t=%.3f frac_h=%.3f" % (k["label"], k["launches_per_step"], k["share"], k["us_per_launch"], k["tflops"], k["frac_tensor"], k["frac_hbm"]))
print("roofline", json.dumps(d["roofline"])[:400])
if d.get("train"): print("train", d["train"]["value"], d["train"]["ms_per_step"], d["train"]["gpu_launches_per_step"])
PY
}
case "${1:-quick}" in
  quick)   # parity of the forward + variants + a short bench with the per-kernel table
    timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -15
    timeout 600 python -m pytest tests/test_gpu_variants.py -x -q -m gpu -k opt_in -s 2>&1 | tail -25
    timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-train > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
    tail -3 gpurun_out/${TAG}_bench.err
    kern_table gpurun_out/${TAG}_bench.json
    ;;
  tests)
    timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
    timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
    ;;
  c1b_trace)  # timeline trace only (profiling build)
    ISHARA_B200_LIB=$PWD/ishara_b200/lib/libishara_b200_trace.so ISHARA_C1B_TRACE_N=3 timeout 200 python tools/fwd_once.py 256 1 > gpurun_out/${TAG}_trace.log 2>&1
    grep "rank\|K=" gpurun_out/${TAG}_trace.log | sed 's/(cycles.*//' | cut -c1-420
    ;;
  attn_trace)
    ISHARA_B200_LIB=$PWD/ishara_b200/lib/libishara_b200_trace.so ISHARA_C1B_TRACE_N=0 timeout 200 python tools/fwd_once.py 256 1 > gpurun_out/${TAG}_atrace.log 2>&1
    grep -A1 "attn trace" gpurun_out/${TAG}_atrace.log | cut -c1-1500
    ;;
  c1b_prof)  # timeline trace (profiling build) + ncu --set full of the fused Conv1DBlock kernel
    ISHARA_B200_LIB=$PWD/ishara_b200/lib/libishara_b200_trace.so timeout 200 python tools/fwd_once.py 256 2 > gpurun_out/${TAG}_trace.log 2>&1
    grep -A4 "c1b trace" gpurun_out/${TAG}_trace.log | head -40
    CMD="python tools/fwd_once.py 256 2" timeout 600 bash tools/ncu_capture.sh ${TAG}_prof_c1b:conv1d_block:3:3
    ;;
  lanes)   # sub-batch lanes sweep: parity once (default), then a short bench per lane count
    timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5
    for L in 1 2 3 4; do
      ISHARA_LANES=$L timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-train > gpurun_out/${TAG}_lanes$L.json 2> gpurun_out/${TAG}_lanes$L.err
      python - gpurun_out/${TAG}_lanes$L.json $L <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print("lanes", sys.argv[2], "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "cfg5", d.get("cfg5", {}).get("value"), "cfg1", d.get("cfg1", {}).get("ms_per_step"))
PY
    done
    ;;
  launches)  # ncu launch list of the bench command (after the same command ran clean without ncu)
    timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --no-train > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || { tail -5 gpurun_out/${TAG}_plain.err; exit 1; }
    timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${TAG}_launches_raw.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu --no-train > gpurun_out/${TAG}_ncu.log 2>&1
    python tools/summarise_launches.py gpurun_out/${TAG}_launches_raw.csv > gpurun_out/${TAG}_launches_step_summary.csv 2>&1
    head -30 gpurun_out/${TAG}_launches_step_summary.csv
    ;;
  fullprof)  # ncu --set full of the three largest kernels, one launch = 256 sequences (one lane) like the profiled pass of the bench
    export ISHARA_LANES=1
    CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-train" timeout 1400 bash tools/ncu_capture.sh \
      ${TAG}_prof_c1b:conv1d_block:12:3 ${TAG}_prof_attn:attn_tc2:4:2 ${TAG}_prof_ffn:ffn_tc:8:2 ${TAG}_prof_rowgemm:gemm_tc_kernel.*16:20:3
    ;;
  dp)   # data-parallel training step at N ranks under a few exchange settings (short inner timeouts)
    N="${3:-2}"
    run() { env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 tools/train_dp_bench.py --steps 20 2>gpurun_out/${TAG}_dp.err | tail -1 | cut -c1-300; }
    timeout 100 python tools/train_dp_bench.py --steps 20 2>/dev/null | tail -1 | cut -c1-200
    run A=1
    run NCCL_MAX_NCHANNELS=2
    run NCCL_MAX_NCHANNELS=4 ISHARA_COMM_BUCKET=4194304
    run ISHARA_COMM_BUCKET=100000000
    ;;
  bench)
    timeout 500 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
    tail -3 gpurun_out/${TAG}_bench.err
    kern_table gpurun_out/${TAG}_bench.json
    ;;
esac
