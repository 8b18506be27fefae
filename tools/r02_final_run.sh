# Final round-2 validation on one B200 (gpurun): GPU tests, smoke, bench (both arms), ncu of the small-batch kernel on
# the slice a rank holds at N = 8 and on C2, every BASELINE config with the final kernels.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_final_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_final_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_final_smoke.log
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/r02_final_bench_n1.json 2> gpurun_out/r02_final_bench_n1.err; echo "bench rc=$?"
cut -c1-600 gpurun_out/r02_final_bench_n1.json; tail -c 600 gpurun_out/r02_final_bench_n1.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02_final_bench_ref.json 2> gpurun_out/r02_final_bench_ref.err; echo "ref rc=$?"
cut -c1-300 gpurun_out/r02_final_bench_ref.json
timeout 300 python tools/profile_gemm.py 1250000 768 1 > gpurun_out/r02_slice_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_filter_small -s 2 -c 1 -o gpurun_out/r02_filter_small_slice python tools/profile_gemm.py 1250000 768 1 > gpurun_out/r02_slice_ncu.log 2>&1; echo "ncu slice rc=$?"
timeout 300 python tools/profile_gemm.py 1000000 384 1 > gpurun_out/r02_c2_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_filter_small -s 2 -c 1 -o gpurun_out/r02_filter_small_c2 python tools/profile_gemm.py 1000000 384 1 > gpurun_out/r02_c2_ncu.log 2>&1; echo "ncu c2 rc=$?"
timeout 900 python tools/config_sweep.py > gpurun_out/r02_config_sweep.log 2>&1; echo "sweep rc=$?"; tail -3 gpurun_out/r02_config_sweep.log | cut -c1-400
ls -la gpurun_out | tail -12
