python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 100 --warmup 5 --allgather > gpurun_out/s22_bench2.json 2> gpurun_out/s22_bench2.err
wc -l gpurun_out/s22_bench2.json; tail -3 gpurun_out/s22_bench2.err
python -m pytest tests/test_gpu_trie.py -m gpu -x -q -k multi_device 2>&1 | tail -2
