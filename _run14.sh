timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_population.py tests/test_gpu_parity.py -m gpu -x -q -k "expansion or matches_oracle or random_mixtures or two_patterns or raw_pattern or generic_rank or none_phase or fixed_matrices or invalid_prob" > gpurun_out/memcheck.log 2>&1
echo "exit $?"; tail -15 gpurun_out/memcheck.log
