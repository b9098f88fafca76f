out=gpurun_out/r01y; mkdir -p $out
timeout 900 python -m pytest tests/test_widen_gpu.py tests/test_gpu_parity.py tests/test_golden.py tests/test_capsule_filter.py -m gpu -q -p no:cacheprovider > $out/tests.log 2>&1; echo "pytest exit $?" >> $out/tests.log; tail -15 $out/tests.log
bash profiles/variant_sweep.sh base R128 S8 S12 2>&1 | tee $out/sweep.log | cut -c1-600
timeout 600 python bench.py --no-cpu-baseline > $out/bench.json 2> $out/bench.err; tail -c 1500 $out/bench.json; tail -3 $out/bench.err
