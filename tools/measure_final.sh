set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/r1_bench_n1.json 2> gpurun_out/r1_bench_n1.err; echo bench rc=$?
timeout 200 python tools/step_profile.py --steps 3 --top 70 > gpurun_out/r1_step_profile.txt 2>&1
