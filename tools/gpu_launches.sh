set -x
TAG=${1:-l}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --profile-mode --steps 1 --warmup 1 > gpurun_out/${TAG}_launches.log 2>&1
wc -l gpurun_out/${TAG}_launches.csv
