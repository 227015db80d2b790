# every -m gpu test, then a short bench line with the schedule variants
set -x
mkdir -p gpurun_out
TAG=${1:-chk}
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_tests.log 2>&1; tail -5 gpurun_out/${TAG}_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --overlap-variant > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err || tail -20 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print('single_stream', d.get('single_stream',{}).get('ms_per_step'), 'fused_cluster', d.get('fused_cluster',{}).get('ms_per_step'), 'fused_all', d.get('fused_cluster_and_glue',{}).get('ms_per_step')); print('value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),'launches',d['gpu_launches_per_step'],d['all_kernel_launches_per_step'])
for k,v in d['per_op'].items(): print('  %-32s %s'%(k,v['ms']))
for k,v in list(d['per_kernel'].items())[:16]: print('  %-28s %s'%(k,v['ms_per_step']))
PY
