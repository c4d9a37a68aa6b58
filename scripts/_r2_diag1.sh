set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.log 2>&1
timeout 300 python scripts/diag_mma_rate2.py > gpurun_out/r2a_mma_rate2.log 2>&1
DRS_V2_TIMELINE=4 timeout 300 python scripts/diag_graph_spans.py > gpurun_out/r2a_spans.log 2>&1
for l in up_convs.2 conv_blocks.0.conv1 conv_blocks.0.conv2 ups.2.transform attention_blocks.2.result gating_signals.2 attention_blocks.2.psi ups.2.conv; do
  DRS_V2_TIMELINE=1 DRS_V2_TIMELINE_LAYER=$l DRS_TL_PAIRS=20 timeout 300 python scripts/diag_layer_timeline.py > gpurun_out/r2a_tl_$l.log 2>&1
done
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu --layers gpurun_out/r2a_layers.json > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
