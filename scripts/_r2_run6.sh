set -x
mkdir -p gpurun_out
for l in up_convs.2 conv_blocks.0.conv1 conv_blocks.0.conv2; do
  DRS_V2_TIMELINE=1 DRS_V2_TIMELINE_LAYER=$l DRS_TL_PAIRS=32 timeout 300 python scripts/diag_layer_timeline.py > gpurun_out/r2f_tl_$l.log 2>&1
done
timeout 300 python scripts/diag_blend.py 3 > gpurun_out/r2f_blend.log 2>&1
