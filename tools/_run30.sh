python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/s30_bench2.json 2> gpurun_out/s30_bench2.err
echo "stdout lines: $(wc -l < gpurun_out/s30_bench2.json)"; head -c 150 gpurun_out/s30_bench2.json; echo
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/s30_ref2.json 2> gpurun_out/s30_ref2.err
echo "ref stdout lines: $(wc -l < gpurun_out/s30_ref2.json)"
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/async_bench.py 1024 2 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
