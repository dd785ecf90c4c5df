set -x
P="python bench.py --workload photon_sphere_16m --steps 20 --warmup 5 --no-cpu --no-e2e"
W="python bench.py --workload wavelength_64m --steps 8 --warmup 8 --no-cpu --no-e2e"
$P > gpurun_out/plain_p.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:pcl_k_photon_multi -s 6 -c 4 -o gpurun_out/r2_photon_multi_v3 $P > gpurun_out/ncu_p.log 2>&1
$W > gpurun_out/plain_w.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:pcl_k_photon_multi -s 1 -c 1 -o gpurun_out/r2_photon_wave_v3 $W > gpurun_out/ncu_w.log 2>&1
PCL_PHOTON_FUSE=1 python scripts/tune_photon.py
ls -la gpurun_out/*v3.ncu-rep
