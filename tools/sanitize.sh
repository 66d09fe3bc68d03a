# compute-sanitizer passes over the GPU parity tests (run on a GPU box, e.g. through gpurun):
#   bash tools/sanitize.sh > gpurun_out/sanitize.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q \
    -k "full_steps_strict or push_deposit_fast" 2>&1 | tail -6
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q \
    -k "test_full_steps_strict and 2-cdims1" 2>&1 | tail -6
