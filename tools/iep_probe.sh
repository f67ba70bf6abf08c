run() { python bench.py --steps $2 --warmup 3 --skip $1 $3 > gpurun_out/s14.log 2>&1; python - <<PY
import json
for l in open("gpurun_out/s14.log"):
    if l.startswith("{"):
        d=json.loads(l)["ie_pipeline"]; print("skip=$1 steps=$2 $3 ->", round(d["ms_per_batch_node_ie"],2), "ms/batch node-IE;", round(d["ms_per_batch_average"],2), "ms/batch average")
PY
}
run e2e,gated,ie,gpu_eager,cpu,dp_parity,other_format 5 ""
run gpu_eager,cpu,dp_parity,other_format,sustained 5 "--sustain-s 0"
run gpu_eager,cpu,dp_parity,other_format,sustained 50 "--sustain-s 0"
run gpu_eager,cpu,dp_parity 50 ""
