set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_row_kernel|ddpm_update" -c 8 -o gpurun_out/r2n_row python bench.py --steps 1 --warmup 3 --no-cpu --no-aggregation > gpurun_out/r2n_ncu.log 2>&1
ncu -i gpurun_out/r2n_row.ncu-rep --page raw --csv > gpurun_out/r2n_row_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none -k regex:blend_gather4 -c 1 -o gpurun_out/r2n_blend python scripts/diag_blend.py 2 > gpurun_out/r2n_blend_ncu.log 2>&1
ncu -i gpurun_out/r2n_blend.ncu-rep --page raw --csv > gpurun_out/r2n_blend_raw.csv 2>/dev/null
rm -f gpurun_out/r2n_row.ncu-rep gpurun_out/r2n_blend.ncu-rep
