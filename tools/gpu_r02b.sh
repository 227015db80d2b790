set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; tail -5 gpurun_out/r02b_bench.err; cut -c1-300 gpurun_out/r02b_bench.json
timeout 300 python bench.py --config 3 --steps 5 --warmup 3 > gpurun_out/r02b_bench_c3.json 2> gpurun_out/r02b_bench_c3.err; tail -5 gpurun_out/r02b_bench_c3.err; cut -c1-300 gpurun_out/r02b_bench_c3.json
timeout 300 python bench.py --config 4 --steps 2 > gpurun_out/r02b_bench_c4.json 2> gpurun_out/r02b_bench_c4.err; tail -5 gpurun_out/r02b_bench_c4.err; cut -c1-1500 gpurun_out/r02b_bench_c4.json
timeout 400 python bench.py --scaling strong --steps 3 --warmup 3 --no-extras > gpurun_out/r02b_bench_strong1.json 2> gpurun_out/r02b_bench_strong1.err; tail -5 gpurun_out/r02b_bench_strong1.err; cut -c1-300 gpurun_out/r02b_bench_strong1.json
timeout 300 python -m pytest tests -m gpu -x -q -k "harness or boundary or malformed or two_threads" 2>&1 | tail -5
