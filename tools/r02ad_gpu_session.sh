# the other workloads of DESIGN's round-1 table on the current kernels (decode-only, 65 536 frames, 5 steps)
set -x
for k in "oms2 dvbs2" "oms2 qc" "wrcq1 dvbs2" "n2d1 dvbs2" "nnms dvbs2" "n2d2 qc" "n2d2 dv12" "rcq dv12"; do
  set -- $k
  python bench.py --decoder $1 --code $2 --steps 5 --warmup 3 --decode-only --no-e2e --no-cpu --no-configs >> gpurun_out/r02ad_all_workloads.jsonl 2>> gpurun_out/r02ad_all_workloads.err
done
python bench.py --decoder basic --code dvbs2 --frames 32768 --steps 5 --warmup 3 --no-e2e --no-cpu --no-configs >> gpurun_out/r02ad_all_workloads.jsonl 2>> gpurun_out/r02ad_all_workloads.err
python - <<'PY'
import json
for l in open("gpurun_out/r02ad_all_workloads.jsonl"):
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]
        print(d["config"]["workload"][:70], round(d["frames_per_s"]), round(d["ms_per_step"],2), "cn", round(r["cn_kernel"]["frac"],3), "vn", round(r["vn_kernel"]["frac"],3), "step", round(r["whole_step"]["frac"],3))
PY
