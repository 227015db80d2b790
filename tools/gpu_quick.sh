# quick check after a kernel change: parity tests of the touched ops, then a short bench line
set -x
mkdir -p gpurun_out
TAG=${1:-q}
KEXPR=${2:-"ballquery or bfs or chain or smoke"}
timeout 600 python -m pytest tests -m gpu -x -q -k "$KEXPR" 2>&1 | tail -4
timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --overlap-variant > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err || tail -20 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print('single_stream', d.get('single_stream',{}).get('ms_per_step'), 'fused_cluster', d.get('fused_cluster',{}).get('ms_per_step'), 'fused_all', d.get('fused_cluster_and_glue',{}).get('ms_per_step')); print('value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),'launches',d['gpu_launches_per_step'],d['all_kernel_launches_per_step'])
for k,v in d['per_op'].items(): print('  %-32s %s'%(k,v['ms']))
for k,v in list(d['per_kernel'].items())[:16]: print('  %-28s %s'%(k,v['ms_per_step']))
PY
