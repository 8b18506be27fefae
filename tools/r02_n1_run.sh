set -x
timeout 900 python bench.py --steps 100 --warmup 5 > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r02_bench_n1_final.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "ref rc=$?"
cat gpurun_out/r02_bench_ref.json | cut -c1-1200
K='regex:gemm_filter|prep_queries|scan_topk|refine_topk|exchange_merge|merge_topk|radix_pass|collect_kernel|sort_emit|select_init|shadow_rows|append_rows'
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/r02_bench_n1_i8_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02_ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_filter_small -s 8 -c 2 -o gpurun_out/r02_filter_small_i8_c3 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | tail -8
