#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/final_tests.log 2>&1; tail -4 gpurun_out/final_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-160
