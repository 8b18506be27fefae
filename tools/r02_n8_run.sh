set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 200 --warmup 10 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; echo "bench8 rc=$?"
tail -c 600 gpurun_out/r02_bench_n8.err
cut -c1-400 gpurun_out/r02_bench_n8.json
timeout 400 python tools/inproc_multi.py 8 > gpurun_out/r02_inproc_n8.json 2> gpurun_out/r02_inproc_n8.err; echo "inproc8 rc=$?"
cat gpurun_out/r02_inproc_n8.json; tail -3 gpurun_out/r02_inproc_n8.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 tools/c4_multi.py > gpurun_out/r02_c4_n8.json 2> gpurun_out/r02_c4_n8.err; echo "c4 rc=$?"
cat gpurun_out/r02_c4_n8.json; tail -3 gpurun_out/r02_c4_n8.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 200 --warmup 10 --no-extra > gpurun_out/r02_bench_n4.json 2> gpurun_out/r02_bench_n4.err; echo "bench4 rc=$?"
cut -c1-300 gpurun_out/r02_bench_n4.json
