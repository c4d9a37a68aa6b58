#!/bin/bash
cp _trace/libdrs_b200.so diffusionremotesensing_b200/libdrs_b200.so
for L in ups.0.transform conv_blocks.2.conv2 ups.1.conv; do
  DRS_CG2=none DRS_V2_TIMELINE=1 DRS_V2_TIMELINE_LAYER=$L DRS_TL_PAIRS=8 python scripts/diag_layer_timeline.py > gpurun_out/y8_tl_$L.log 2>&1
done
