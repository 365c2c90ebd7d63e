timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r02_final_tests4.log 2>&1; tail -2 gpurun_out/r02_final_tests4.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_c4_v3.json 2> gpurun_out/r02_bench_c4_v3.err; cut -c1-220 gpurun_out/r02_bench_c4_v3.json
python bench.py --leaves 16 > gpurun_out/r02_bench_k16_v3.json 2>/dev/null; cut -c1-160 gpurun_out/r02_bench_k16_v3.json
python bench.py --game ttt > gpurun_out/r02_bench_ttt_v3.json 2>/dev/null; cut -c1-160 gpurun_out/r02_bench_ttt_v3.json
python bench.py --game chess > gpurun_out/r02_bench_chess_v3.json 2>/dev/null; cut -c1-160 gpurun_out/r02_bench_chess_v3.json
ncu --set full --clock-control none --import-source on -k regex:k_eval_umma -s 1 -c 1 -f -o gpurun_out/r02_ncu_eval_final python tools/prof_async.py async > gpurun_out/r02_ncu_eval_final.log 2>&1; tail -1 gpurun_out/r02_ncu_eval_final.log
