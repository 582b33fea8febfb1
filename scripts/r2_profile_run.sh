#!/bin/bash
# Round-2 evidence run on one B200: full GPU test suite, the driver's bench command, the ncu launch list of a short bench
# run (inference + one training step) and an ncu --set full capture of the SLIC image kernel.
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_tests_final.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_final.log 2> gpurun_out/r2_bench_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.log 2>&1
python bench.py --steps 2 --warmup 3 --train-steps 1 --no-cpu-baseline --no-configs --no-staging > gpurun_out/r2_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --train-steps 1 --no-cpu-baseline --no-configs --no-staging > gpurun_out/r2_ncu_launches.log 2>&1
tail -3 gpurun_out/r2_tests_final.log
tail -c 300 gpurun_out/r2_bench_final.err
