set -x
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -x -q > gpurun_out/t_train.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_train.log
tail -30 gpurun_out/t_train.log
