# last check of a round: every -m gpu test, then bench.py with the driver's flags
mkdir -p gpurun_out
TAG=${1:-final}
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 )
timeout 400 python bench.py --gpus 1 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench.json"))
print("default run: value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "steps", d["steps"], "warmup", d["warmup"], "e2e", round(d["e2e"]["value"], 1),
      "launches", d["gpu_launches_per_step"], "single", d.get("single_stream", {}).get("ms_per_step"), "fused_all", d.get("fused_cluster_and_glue", {}).get("ms_per_step"))
print("roofline", d["roofline"]["kernel"], d["roofline"]["frac"], "cpu", d["cpu_baseline"]["value"], "clocks", d["clocks"])
PY
