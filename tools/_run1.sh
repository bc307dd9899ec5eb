set -x
export DH_NO_BUILD=1
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 0 python -m pytest tests/test_gpu_configs.py tests/test_gpu_biwi.py -m gpu -x -q -k "box_image_shapes or decode or predict_from_compressed or variants" > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/memcheck.log
tail -15 gpurun_out/memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_biwi.py tests/test_gpu_parity.py -m gpu -x -q -k "decode_edge or small_forest_every_stage or hand_built" > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/racecheck.log
tail -15 gpurun_out/racecheck.log
