#!/bin/bash
# One measurement pass on a B200 box (run through gpurun): GPU tests, the plain bench, the ncu launch list of the same
# command, and one `ncu --set full` capture of the dominant kernels.  Outputs go to gpurun_out/ (scratch); the summaries
# that are kept are copied into profiles/ by tools/profile_collect.py.
set -u
tag=${1:-r01}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -x -q -m gpu > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest_gpu.log
timeout 600 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 > $out/${tag}_ncu_bench.log 2>&1; echo "ncu launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:star_fused -c 1 -s 3 -f -o $out/${tag}_star_fused \
    python tools/prof_star.py 1 2368 8 > $out/${tag}_ncu_star.log 2>&1; echo "ncu star rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vocab_argmax_tc -c 1 -s 3 -f -o $out/${tag}_vocab \
    python tools/prof_vocab.py > $out/${tag}_ncu_vocab.log 2>&1; echo "ncu vocab rc=$?"
tail -3 $out/${tag}_pytest_gpu.log; cat $out/${tag}_bench.json
