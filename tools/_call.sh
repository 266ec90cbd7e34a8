# scratch command file for `gpurun -- 'bash tools/_call.sh'` (edited per experiment)
mkdir -p gpurun_out/r2
python tools/kineto_step.py --graph --top 200 2>&1 | grep -v "Warn\|_warn_once" > gpurun_out/r2/kineto_final.txt; head -3 gpurun_out/r2/kineto_final.txt; grep "stream\|gaps" gpurun_out/r2/kineto_final.txt | head -5
