#!/bin/bash
# One measurement pass on a B200 box (run through gpurun): GPU tests, smoke(), the plain bench, the ncu launch list of the
# same command, and one `ncu --set full` capture of the dominant kernels.  Outputs go to gpurun_out/ (scratch); the
# summaries that are kept are copied into profiles/ by tools/profile_collect.py.
#   tools/profile_round.sh <tag> [steps: tests smoke bench launches star vocab trace]
set -u
tag=${1:-r02}; shift
steps=${*:-tests smoke bench launches star}
out=gpurun_out
mkdir -p $out
for s in $steps; do
case $s in
tests) timeout 1200 python -m pytest tests -x -q -m gpu > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest_gpu.log; tail -15 $out/${tag}_pytest_gpu.log;;
smoke) timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -8 $out/${tag}_smoke.log;;
bench) timeout 900 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"; cat $out/${tag}_bench.json; tail -3 $out/${tag}_bench.err;;
launches) timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --train-steps 0 > $out/${tag}_bench_short.json 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --train-steps 0 > $out/${tag}_ncu_bench.log 2>&1; echo "ncu launch list rc=$?";;
star) timeout 300 python tools/prof_star.py 1 2368 8 > $out/${tag}_prof_star_plain.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:star_fused -c 1 -s 3 -f -o $out/${tag}_star_fused \
    python tools/prof_star.py 1 2368 8 > $out/${tag}_ncu_star.log 2>&1; echo "ncu star rc=$?"; cat $out/${tag}_prof_star_plain.log;;
vocab) timeout 300 python tools/prof_vocab.py > $out/${tag}_prof_vocab_plain.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:vocab_argmax_tc -c 1 -s 3 -f -o $out/${tag}_vocab \
    python tools/prof_vocab.py > $out/${tag}_ncu_vocab.log 2>&1; echo "ncu vocab rc=$?";;
kernels) # one `ncu --set full` capture each of the other kernels of a bench step (graph nodes are profiled individually)
  for k in vocab_argmax_tc tar_tail mha_decode_attention gemm_k128 "gemm_tc_kernel<128, 3, 2>" add_layernorm; do
    f=$(echo $k | tr -c 'a-zA-Z0-9_' '_' | cut -c1-24)
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$k" -c 1 -s 40 -f -o $out/${tag}_k_$f \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --train-steps 0 > $out/${tag}_ncu_k_$f.log 2>&1; echo "ncu $k rc=$?"
  done;;
trace) timeout 300 python tools/star_trace.py 1 17 4 > $out/${tag}_star_trace.txt 2>&1; echo "trace rc=$?"; tail -5 $out/${tag}_star_trace.txt;;
esac
done
