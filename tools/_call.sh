mkdir -p gpurun_out/r2
O=gpurun_out/r2
python -m pytest tests -m gpu -q -x 2>&1 | tail -2 | tee $O/t42_all.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > $O/bench42_default_with_baselines.log 2>&1; tail -1 $O/bench42_default_with_baselines.log | cut -c1-200
for w in hqavitv2_c100 qavitv2_c100 hqavit_stl96 hqavit_tinyin; do
  python bench.py --workload $w --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > $O/bench42_$w.log 2>&1; tail -1 $O/bench42_$w.log | cut -c1-160
done
for w in hqavit_c100 qavitv2_c100; do
  python bench.py --workload $w --mode infer --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > $O/bench42_infer_$w.log 2>&1; tail -1 $O/bench42_infer_$w.log | cut -c1-160
done
python bench.py --workload qavitv2_c100 --batch 256 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > $O/bench42_qavitv2_b256_n1.log 2>&1; tail -1 $O/bench42_qavitv2_b256_n1.log | cut -c1-160
