python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py > gpurun_out/bench_r1d.log 2> gpurun_out/bench_r1d.err; tail -c 1500 gpurun_out/bench_r1d.log; tail -3 gpurun_out/bench_r1d.err
python bench.py --impl reference > gpurun_out/bench_ref_r1d.log 2>&1; tail -c 600 gpurun_out/bench_ref_r1d.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_train_core.csv python scripts/train_step_probe.py 26 > gpurun_out/ncu_train_core.log 2>&1; tail -2 gpurun_out/ncu_train_core.log
ncu --set full --clock-control none --import-source on -k regex:resize -c 4 -o gpurun_out/prof_resize python scripts/resize_bench.py > gpurun_out/ncu_resize.log 2>&1; tail -2 gpurun_out/ncu_resize.log
