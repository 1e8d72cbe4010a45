#!/bin/bash
# GPU call 21: new tests of the generic-sweep machinery + full suite
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_generic_sweeps.py -m gpu -q -x > gpurun_out/r2v_generic_tests.log 2>&1; echo "rc $?" >> gpurun_out/r2v_generic_tests.log
timeout 1500 python -m pytest tests -m gpu -q -rs > gpurun_out/r2v_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r2v_pytest.log
tail -n 25 gpurun_out/r2v_generic_tests.log | cut -c1-250
grep -v "^$" gpurun_out/r2v_pytest.log | tail -n 8 | cut -c1-300
