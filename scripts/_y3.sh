#!/bin/bash
cp _trace/libdrs_b200.so diffusionremotesensing_b200/libdrs_b200.so
for L in ups.2.transform ups.1.transform ups.0.transform; do
  DRS_V2_TIMELINE=1 DRS_V2_TIMELINE_LAYER=$L DRS_TL_PAIRS=8 python scripts/diag_layer_timeline.py > gpurun_out/y5_tl_$L.log 2>&1
done
tail -12 gpurun_out/y5_tl_ups.2.transform.log
