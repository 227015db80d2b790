# A/B of an environment switch: tools/ab_env.sh VAR v1 v2 ... (per-kernel totals of one single-stream step each)
VAR=$1; shift
for v in "$@"; do
  echo "== $VAR=$v"
  env $VAR=$v timeout 200 python tools/timeline.py --out gpurun_out/tl_ab.json 2>&1 | grep -E "${AB_GREP:-span}"
done
