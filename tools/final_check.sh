mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/final_gputests.txt; cat gpurun_out/final_gputests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -c 3000 gpurun_out/final_bench.json
timeout 600 python tools/bench_configs.py > gpurun_out/final_configs.txt 2>&1; tail -5 gpurun_out/final_configs.txt
