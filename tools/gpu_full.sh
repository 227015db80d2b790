# full check: every -m gpu test, then the bench lines of all configs
set -x
mkdir -p gpurun_out
TAG=${1:-full}
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_tests.log 2>&1; tail -6 gpurun_out/${TAG}_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err || tail -20 gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --config 3 --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err || tail -20 gpurun_out/${TAG}_bench_c3.err
timeout 300 python bench.py --config 4 --steps 2 > gpurun_out/${TAG}_bench_c4.json 2> gpurun_out/${TAG}_bench_c4.err || tail -20 gpurun_out/${TAG}_bench_c4.err
python - <<PY
import json
for f in ("bench", "bench_c3", "bench_c4"):
    try:
        d = json.load(open("gpurun_out/${TAG}_%s.json" % f))
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, "value", round(d.get("value", 0), 1), "ms", round(d.get("ms_per_step", 0), 3), "e2e", d.get("e2e", {}).get("value"),
          "single_stream", d.get("single_stream", {}).get("ms_per_step"), "cpu", (d.get("cpu_baseline") or {}).get("value"),
          "mixed", (d.get("reference_mixed") or {}).get("value"), "unchanged", (d.get("unchanged_caller") or {}).get("value"))
    if f == "bench_c4":
        print("  speaker", {k: d["speaker_forward"].get(k) for k in ("value", "ms_per_step", "ops_share", "ops_ms", "captions")})
        print("  detector", {k: d["detector_only"].get(k) for k in ("value", "ms_per_step", "ops_share", "ops_ms")})
PY
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
