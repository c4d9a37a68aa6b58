set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
timeout 900 python bench.py --steps 100 --warmup 5 --layers gpurun_out/r2k_layers.json > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
echo "bench rc=$?" >> gpurun_out/r2k_bench.err
DRS_V2_TIMELINE=4 timeout 300 python scripts/diag_graph_spans.py > gpurun_out/r2k_spans.log 2>&1
