set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv | tail -1
free -g | head -2; nproc
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02a_tests.log 2>&1; tail -5 gpurun_out/r02a_tests.log
timeout 400 compute-sanitizer --tool memcheck --log-file gpurun_out/r02a_memcheck_smoke.log python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02a_memcheck_smoke.out 2>&1; tail -3 gpurun_out/r02a_memcheck_smoke.log
timeout 400 compute-sanitizer --tool racecheck --log-file gpurun_out/r02a_racecheck_smoke.log python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02a_racecheck_smoke.out 2>&1; tail -3 gpurun_out/r02a_racecheck_smoke.log
timeout 500 compute-sanitizer --tool memcheck --log-file gpurun_out/r02a_memcheck_bfs.log python -m pytest tests/test_gpu_parity.py -q -x -k "truncated or one_way or malformed or cap" > gpurun_out/r02a_memcheck_bfs.out 2>&1; tail -3 gpurun_out/r02a_memcheck_bfs.log; tail -3 gpurun_out/r02a_memcheck_bfs.out
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; cut -c1-600 gpurun_out/r02a_bench.json
