#!/bin/bash
# One gpurun call, many artefacts (GPU slots are scarce): usage  tools/gpu_batch.sh "<stages>" <tag>
#   T tests   S dcn sweep   B bench (inference)   R bench --mode train   N ncu captures   C other configs
stages="$1"; tag="${2:-r02}"
mkdir -p gpurun_out
run() { echo "== $1"; shift; "$@"; }
if [[ "$stages" == *T* ]]; then
  timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/${tag}_pytest_gpu.log; tail -15 gpurun_out/${tag}_pytest_gpu.log
fi
if [[ "$stages" == *S* ]]; then
  EDVR_B200_DCN_MC=1 timeout 300 python tools/dcn_sweep.py --modes fused --json gpurun_out/${tag}_dcn_sweep_mc1.json 2>&1 | tail -12
  EDVR_B200_DCN_MC=0 timeout 300 python tools/dcn_sweep.py --modes fused,legacy --json gpurun_out/${tag}_dcn_sweep_mc0.json 2>&1 | tail -24
fi
if [[ "$stages" == *B* ]]; then
  timeout 900 python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; tail -c 400 gpurun_out/${tag}_bench_n1.err
  python - <<PY
import json
try:
    d=json.loads([x for x in open("gpurun_out/${tag}_bench_n1.json") if x.startswith("{")][-1])
    print("bench", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "b1", d.get("latency_b1",{}).get("value"), d["clocks"])
    print("cpu", d.get("cpu_baseline")); print("ref_cuda", d.get("ref_cuda")); print("dcn", d.get("roofline_dcn"))
    print({k:(v["ms"],v["tflops"]) for k,v in d["kernel_shares"].items()})
except Exception as e: print("bench parse error", e)
PY
fi
if [[ "$stages" == *R* ]]; then
  timeout 900 python bench.py --mode train > gpurun_out/${tag}_bench_train_n1.json 2> gpurun_out/${tag}_bench_train_n1.err; tail -c 1500 gpurun_out/${tag}_bench_train_n1.err; cut -c1-1800 gpurun_out/${tag}_bench_train_n1.json
fi
if [[ "$stages" == *N* ]]; then
  EDVR_B200_DCN_SITE=fused timeout 600 ncu --set full --clock-control none --import-source on -k regex:dcn_site -s 3 -c 1 -o gpurun_out/${tag}_ncu_dcn_site python tools/one_site.py 4 180 320 128 8 0.02 2>&1 | tail -2
  EDVR_BENCH_PROFILING=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${tag}_ncu_launches_bench.csv python bench.py --steps 2 --warmup 3 > /dev/null 2>&1; wc -l gpurun_out/${tag}_ncu_launches_bench.csv
fi
if [[ "$stages" == *C* ]]; then
  timeout 900 python tools/bench_configs.py 2>&1 | tail -25; cp gpurun_out/bench_configs.json gpurun_out/${tag}_bench_configs.json 2>/dev/null
fi
