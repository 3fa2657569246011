#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3f_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r3f_tests.log
python tools/check_fwd_tiles.py 512; echo "tiles rc=$?"
cd tools && python bench_corrupt.py > ../gpurun_out/r3f_corrupt.json 2> ../gpurun_out/r3f_corrupt.err; cd ..; cat gpurun_out/r3f_corrupt.json; tail -2 gpurun_out/r3f_corrupt.err
