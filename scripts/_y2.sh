#!/bin/bash
# clean lib: layer table without the TMA store path
DRS_V2_NO_TMA_STORE=1 python bench.py --steps 20 --warmup 5 --no-cpu --no-aggregation --layers gpurun_out/y2_layers_nostage.json > gpurun_out/y2_bench_nostage.json 2>/dev/null
cp _trace/libdrs_b200.so diffusionremotesensing_b200/libdrs_b200.so
for L in ups.2.transform ups.1.transform ups.0.transform; do
  DRS_V2_TIMELINE=1 DRS_V2_TIMELINE_LAYER=$L DRS_TL_PAIRS=8 python scripts/diag_layer_timeline.py > gpurun_out/y2_tl_$L.log 2>&1
done
DRS_V2_NO_TMA_STORE=1 DRS_V2_TIMELINE=1 DRS_V2_TIMELINE_LAYER=ups.2.transform DRS_TL_PAIRS=8 python scripts/diag_layer_timeline.py > gpurun_out/y2_tl_ups.2.transform_nostage.log 2>&1
tail -30 gpurun_out/y2_tl_ups.2.transform.log
