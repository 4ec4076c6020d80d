# last sanity check of a round: the GPU suite (minus the one 55 s CPU-oracle-bound case) + smoke() on the final tree
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q --maxfail=10 -rf --deselect "tests/test_gpu_parity.py::test_baseline_configs_vs_oracle[car-3-cmamppi-375-50-10-kw3]" ) > gpurun_out/final_check_pytest.log 2>&1
tail -5 gpurun_out/final_check_pytest.log | cut -c1-200
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
