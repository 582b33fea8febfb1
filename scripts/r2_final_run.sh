#!/bin/bash
# Final evidence run of round 2 on one B200: GPU test suite, the driver's bench commands, ncu launch list of a short bench,
# ncu --set full of the edge launch (roofline.traffic) and of the node launch with the folded aggregation.
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/f_tests.log
python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_bench_ref.json 2>&1
python bench.py --steps 2 --warmup 3 --train-steps 1 --no-cpu-baseline --no-configs --no-staging > gpurun_out/f_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/f_launches.csv \
    python bench.py --steps 2 --warmup 3 --train-steps 1 --no-cpu-baseline --no-configs --no-staging > gpurun_out/f_ncu_launches.log 2>&1
python scripts/chain_ncu_case.py 512 node > gpurun_out/f_case.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tc_chain2_kernel -s 2 -c 1 -f -o gpurun_out/f_chain2_edge python scripts/chain_ncu_case.py 512 > gpurun_out/f_ncu_edge.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_chain2n_kernel -s 2 -c 1 -f -o gpurun_out/f_chain2n_node_agg python scripts/chain_ncu_case.py 512 node > gpurun_out/f_ncu_node.log 2>&1
tail -3 gpurun_out/f_tests.log
tail -c 300 gpurun_out/f_bench.err
