python -m pytest tests/test_gpu_tc_engine.py tests/test_gpu_model.py -x -q -m gpu 2>&1 | tail -15
for p in core opwise core opwise; do GNC_TRAIN_PATH=$p python scripts/train_step_probe.py 104 2>&1 | grep -E "^step|sum of" ; done
