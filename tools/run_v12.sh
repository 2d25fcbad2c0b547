FWB_SWEEP=400 python -m pytest tests -m gpu -q --tb=line -k "random_option_sweep" 2>&1 | tail -12 > gpurun_out/pytest_v12.log; tail -12 gpurun_out/pytest_v12.log
