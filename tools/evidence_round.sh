#!/bin/bash
# round-2 evidence on the committed tree: GPU tests, the bench with its CPU leg, the reference arm, launch list, full captures
python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/${1:-r2d}_tests.log; tail -3 gpurun_out/${1:-r2d}_tests.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${1:-r2d}_ref.json 2> gpurun_out/${1:-r2d}_ref.err; tail -c 600 gpurun_out/${1:-r2d}_ref.json
bash tools/profile_round.sh ${1:-r2d}
