mkdir -p gpurun_out/r2
O=gpurun_out/r2
python bench.py > $O/bench41_default_with_baselines.log 2>&1; tail -1 $O/bench41_default_with_baselines.log | cut -c1-300
python bench.py --impl reference --steps 8 --warmup 3 > $O/bench41_reference_arm.log 2>&1; tail -1 $O/bench41_reference_arm.log | cut -c1-200
for w in hqavitv2_c100 qavitv2_c100 hqavit_stl96 hqavit_tinyin; do
  python bench.py --workload $w --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > $O/bench41_$w.log 2>&1; tail -1 $O/bench41_$w.log | cut -c1-160
done
for w in hqavit_c100 qavitv2_c100; do
  python bench.py --workload $w --mode infer --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > $O/bench41_infer_$w.log 2>&1; tail -1 $O/bench41_infer_$w.log | cut -c1-160
done
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > $O/launch_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches_v2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > $O/launch_ncu.log 2>&1
tail -2 $O/launch_ncu.log | cut -c1-200
