set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 400 python bench.py > gpurun_out/r1_bench_n1.json 2> gpurun_out/r1_bench_n1.err; echo bench rc=$?
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1_bench_ref.json 2> gpurun_out/r1_bench_ref.err; echo ref rc=$?
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_elapsed
timeout 300 ncu --metrics $M --clock-control none -k regex:attn_ --csv --log-file gpurun_out/r1_metrics_attn.csv python tools/kernel_bench.py attn --B 64 --N 4096 --C 16 --bwd --iters 1 > /dev/null 2>&1
timeout 300 ncu --metrics $M --clock-control none -k regex:"attn_|gemm_" --csv --log-file gpurun_out/r1_metrics_big.csv python tools/kernel_bench.py attn --B 16 --N 4096 --C 512 --iters 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_tc_kernel --launch-skip 2 -c 1 -o gpurun_out/r1_attn_bwd_full python tools/kernel_bench.py attn --B 64 --N 4096 --C 16 --bwd --iters 1 > /dev/null 2>&1
timeout 200 python tools/step_profile.py --steps 3 --top 70 > gpurun_out/r1_step_profile.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file gpurun_out/r1_launches_bf16_tc_step.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out
