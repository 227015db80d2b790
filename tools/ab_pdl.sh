# A/B of programmatic dependent launch: the same build with and without the launch attribute
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pdl_tests.log 2>&1; tail -4 gpurun_out/pdl_tests.log
for v in 0 1 0 1; do
  PG_B200_NO_PDL=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --overlap-variant --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('NO_PDL=$v value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'single', round(d['single_stream']['ms_per_step'],3), 'fused_all', round(d['fused_cluster_and_glue']['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))
"
done
