set -x
G="python bench.py --workload gravity_256k --steps 3 --warmup 3 --no-cpu --no-e2e"
W="python bench.py --workload wavelength_64m --steps 8 --warmup 8 --no-cpu --no-e2e"
$G > gpurun_out/plain_g.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:pcl_k_gravity_x2 -s 4 -c 1 -o gpurun_out/r2_gravity_x2 $G > gpurun_out/ncu_g.log 2>&1
$W > gpurun_out/plain_w.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:pcl_k_photon_multi -s 1 -c 1 -o gpurun_out/r2_photon_wave $W > gpurun_out/ncu_w.log 2>&1
tail -2 gpurun_out/ncu_g.log gpurun_out/ncu_w.log
ls -la gpurun_out/*.ncu-rep
