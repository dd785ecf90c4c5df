set -x
G="python bench.py --workload gravity_256k --steps 3 --warmup 3 --no-cpu --no-e2e"
K="python bench.py --no-sub --no-e2e --no-cpu --steps 3 --warmup 3"
$G > gpurun_out/plain_g.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:pcl_k_gravity_x2 -s 4 -c 1 -o gpurun_out/r2_gravity_x2_v2 $G > gpurun_out/ncu_g.log 2>&1
$K > gpurun_out/plain_k.log 2>&1 && ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:pcl_k_kinematics -s 4 -c 2 --csv --log-file gpurun_out/r2_kin_1b_dram.csv $K > gpurun_out/ncu_k.log 2>&1
tail -n 3 gpurun_out/ncu_g.log; tail -n 3 gpurun_out/ncu_k.log; cat gpurun_out/r2_kin_1b_dram.csv | tail -8
