set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1; nproc >> gpurun_out/topo.txt; lscpu | grep -i "numa\|socket\|model name" >> gpurun_out/topo.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_n8.log 2> gpurun_out/bench_n8.err; echo "rc=$?" >> gpurun_out/bench_n8.err
tail -5 gpurun_out/bench_n8.err; cat gpurun_out/bench_n8.log | cut -c1-400
