mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_golden.py -q -m gpu > gpurun_out/r1_golden.log 2>&1; echo "golden rc=$?"
tail -15 gpurun_out/r1_golden.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
