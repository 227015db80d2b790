# ncu --set full of ONE launch: gpu_ncu1.sh TAG KERNEL_REGEX SKIP
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o gpurun_out/$1 python bench.py --profile-mode --steps 1 --warmup 1 > gpurun_out/$1.log 2>&1
ls -la gpurun_out/$1.ncu-rep
