set -u
out=gpurun_out; mkdir -p $out
timeout 600 python bench.py > $out/r02b_bench.json 2> $out/r02b_bench.err; echo "bench rc=$?"; cat $out/r02b_bench.json; tail -5 $out/r02b_bench.err
for ch in AWGN Rayleigh; do
timeout 300 python tools/snr_sweep.py --data tests/golden/europarl_test.npz --channel $ch --out $out/r02_sweep_real_$ch.pkl > $out/r02_sweep_real_$ch.log 2>&1; echo "sweep $ch rc=$?"; tail -2 $out/r02_sweep_real_$ch.log
done
