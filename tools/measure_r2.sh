# Round-2 measurement pass (run on the GPU box through gpurun; every ncu command runs AFTER the same program has exited 0
# without ncu; numbers printed under ncu are never bench values):
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/measure_r2.sh'
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err; echo bench rc=$?
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > $O/r2_bench_ref.json 2> $O/r2_bench_ref.err; echo ref rc=$?
timeout 300 python bench.py --config cond128 --steps 20 --warmup 3 --no-cpu-baseline > $O/r2_bench_cond128.json 2> $O/r2_bench_cond128.err
timeout 300 python tools/sweep.py --out $O/r2_sweep.json > /dev/null 2> $O/r2_sweep.err
timeout 200 python tools/step_profile.py --steps 3 --top 70 > $O/r2_step_profile.txt 2>&1
timeout 200 python tools/op_profile.py 16 4096 512 > $O/r2_op_profile_C512.txt 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_elapsed
timeout 300 python tools/kernel_bench.py attn --B 64 --N 4096 --C 16 --bwd --iters 1 > /dev/null 2>&1 && \
timeout 400 ncu --metrics $M --clock-control none -k regex:attn_ --csv --log-file $O/r2_metrics_attn.csv python tools/kernel_bench.py attn --B 64 --N 4096 --C 16 --bwd --iters 1 > /dev/null 2>&1
timeout 300 python tools/kernel_bench.py attn --B 16 --N 4096 --C 512 --bwd --iters 1 > /dev/null 2>&1 && \
timeout 400 ncu --metrics $M --clock-control none -k regex:"attn_|gemm_|gram_|conv_tc" --csv --log-file $O/r2_metrics_big.csv python tools/kernel_bench.py attn --B 16 --N 4096 --C 512 --bwd --iters 1 > /dev/null 2>&1
timeout 300 python tools/kernel_bench.py sn --rows 4096 --cols 4096 --iters 1 > /dev/null 2>&1 && \
timeout 300 ncu --metrics $M --clock-control none -k regex:sn_power --csv --log-file $O/r2_metrics_sn_4096x4096.csv python tools/kernel_bench.py sn --rows 4096 --cols 4096 --iters 1 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_tc_kernel --launch-skip 2 -c 1 -o $O/r2_attn_bwd_full python tools/kernel_bench.py attn --B 64 --N 4096 --C 16 --bwd --iters 1 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tc4_kernel --launch-skip 2 -c 1 -o $O/r2_attn_fwd_full python tools/kernel_bench.py attn --B 64 --N 4096 --C 16 --iters 1 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_big_kernel --launch-skip 2 -c 1 -o $O/r2_attn_bwd_big_full python tools/kernel_bench.py attn --B 16 --N 4096 --C 512 --bwd --iters 1 > /dev/null 2>&1
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-variant > /dev/null 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r2_launches_bf16_tc_step.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-variant > /dev/null 2>&1
ls -la $O | tail -30
