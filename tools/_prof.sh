set -x
python bench.py > gpurun_out/r01b_bench.json 2> gpurun_out/r01b_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01b_bench_ref.json 2>> gpurun_out/r01b_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01b_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r01b_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"tile_kernel|permute_bulk|span_kernel" -s 9 -c 3 -o gpurun_out/r01b_step python tools/profile_step.py --steps 6 > gpurun_out/r01b_prof_step.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lse_sample -s 4 -c 4 -o gpurun_out/r01b_sampler python tools/profile_sampler.py > gpurun_out/r01b_prof_sampler.log 2>&1
python tools/bench_sampler.py > gpurun_out/r01b_sampler.txt 2>&1
ls -la gpurun_out | tail -12
