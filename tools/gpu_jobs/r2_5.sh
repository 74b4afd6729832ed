set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"opt_tail_kernel|mb_begin_kernel|tc_wgrad_conv1_kernel" -s 6 -c 6 -o gpurun_out/r2_tail python tools/profile_step.py --minibatches 3 > gpurun_out/r2_ncu_tail.log 2>&1
tail -3 gpurun_out/r2_ncu_tail.log
