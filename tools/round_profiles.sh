# End-of-round measurements on one B200 (run through gpurun from the repo root); outputs under gpurun_out/.
set -x
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_final.json 2>> gpurun_out/bench_final.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --skip-cpu-baseline --main-py-steps 0 --frames 0 > gpurun_out/b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"field_backward_mma" -s 3 -c 1 -o gpurun_out/final_bwd python bench.py --steps 2 --warmup 1 --skip-cpu-baseline --main-py-steps 0 --frames 0 --no-graph > gpurun_out/b2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"field_forward_umma" -s 7 -c 1 -o gpurun_out/final_fwd python bench.py --steps 2 --warmup 1 --skip-cpu-baseline --main-py-steps 0 --frames 0 --no-graph > gpurun_out/b3.log 2>&1
bash tools/cull_sweep.sh > gpurun_out/cull_final.log 2>&1
python tools/time_forward_culling.py 2>&1 | grep frac > gpurun_out/fwdcull_final.log
python tools/ab_backward.py --cfg cfg3 2>&1 | tail -1 > gpurun_out/bwd_cfg3.log
