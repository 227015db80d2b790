# ncu --set full of chosen kernels (one launch each, the shifted-set launch of a warm step)
set -x
mkdir -p gpurun_out
TAG=${1:-p}
shift
for K in "$@"; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -f -o gpurun_out/${TAG}_$K python bench.py --profile-mode --steps 1 --warmup 1 > gpurun_out/${TAG}_$K.log 2>&1
  ls -la gpurun_out/${TAG}_$K.ncu-rep
done
