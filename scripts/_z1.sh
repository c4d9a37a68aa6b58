#!/bin/bash
python -m pytest tests -m gpu -q > gpurun_out/z1_tests.log 2>&1; tail -3 gpurun_out/z1_tests.log
python bench.py --steps 300 --warmup 10 --layers gpurun_out/z1_layers.json > gpurun_out/z1_bench.json 2> gpurun_out/z1_bench.err; head -c 300 gpurun_out/z1_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/z1_ref.json 2> gpurun_out/z1_ref.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/z1_smoke.log 2>&1; tail -1 gpurun_out/z1_smoke.log
bash scripts/profile_ncu.sh gpurun_out
ls -la gpurun_out/ncu_*
