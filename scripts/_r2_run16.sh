set -x
for l in up_convs.2 conv_blocks.0.conv1; do
  DRS_V2_TIMELINE=1 DRS_V2_TIMELINE_LAYER=$l timeout 300 python scripts/diag_row_timeline.py > gpurun_out/r2p_tl_$l.log 2>&1
  DRS_V2_TIMELINE=8 DRS_V2_TIMELINE_LAYER=$l timeout 300 python scripts/diag_row_cta_times.py > gpurun_out/r2p_cta_$l.log 2>&1
done
