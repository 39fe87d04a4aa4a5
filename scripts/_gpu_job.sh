python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
(python -m pytest tests/test_gpu_engine.py -x -q -m gpu) 2>&1 | tail -3
