set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"opt_tail_kernel|pf_loss_kernel" -s 4 -c 3 -o gpurun_out/r2_tail2 python tools/profile_step.py --minibatches 3 > gpurun_out/r2_ncu_tail2.log 2>&1
tail -2 gpurun_out/r2_ncu_tail2.log
