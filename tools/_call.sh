mkdir -p gpurun_out/r2
python -m pytest tests/test_gpu_tokens_fused.py -q -s > gpurun_out/r2/t05_fused.log 2>&1; tail -12 gpurun_out/r2/t05_fused.log
python -m pytest tests -q -m gpu -s > gpurun_out/r2/t05_all.log 2>&1; tail -6 gpurun_out/r2/t05_all.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/r2/bench05.log 2>&1; tail -c 900 gpurun_out/r2/bench05.log
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2/launches05.csv python tools/profile_step.py --batch 4736 --dropout 0.1 --drop-path 0.1 > gpurun_out/r2/prof05_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/r2/launches05.csv 70 > gpurun_out/r2/launches05.summary.txt; head -24 gpurun_out/r2/launches05.summary.txt
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'tlf_bwd|upf_bwd|cmpf_bwd|cga_mma_bwd|attn_mma_bwd' -c 7 -o gpurun_out/r2/ncu05_bwd python tools/profile_step.py --batch 4736 --dropout 0.1 --drop-path 0.1 > gpurun_out/r2/prof05_ncufull.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'tlf_fwd|cmpf_fwd|upf_fwd' -c 3 -o gpurun_out/r2/ncu05_fwd python tools/profile_step.py --batch 4736 --dropout 0.1 --drop-path 0.1 >> gpurun_out/r2/prof05_ncufull.log 2>&1
ls -la gpurun_out/r2/ | tail -5
