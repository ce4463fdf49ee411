set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1; nproc >> gpurun_out/topo.txt; lscpu | grep -i "numa\|socket\|model name" >> gpurun_out/topo.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err; echo "rc=$?" >> gpurun_out/bench_n2.err
tail -5 gpurun_out/bench_n2.err; cat gpurun_out/bench_n2.log | cut -c1-400
