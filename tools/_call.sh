mkdir -p gpurun_out/r2 /tmp/ncu
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline 2>&1 | tail -1 | grep -o '"ms_per_step": [0-9.]*' | head -1
SEC="--section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section SchedulerStats --section WarpStateStats --section InstructionStats"
timeout 600 ncu $SEC --clock-control none --profile-from-start off -c 450 -o /tmp/ncu/parts python tools/profile_parts.py --batch 1184 > gpurun_out/r2/ncu_parts.log 2>&1
tail -2 gpurun_out/r2/ncu_parts.log
python tools/ncu_brief.py /tmp/ncu/parts.ncu-rep > gpurun_out/r2/ncu_brief_parts_b1184.txt 2>&1; wc -l gpurun_out/r2/ncu_brief_parts_b1184.txt
# compute-sanitizer on a small configuration (B = 8): memcheck over one fp32 + one bf16 step, racecheck over the bf16 step
cat > /tmp/san.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch, qavit_b200 as Q
prec = sys.argv[1]
torch.manual_seed(0)
m = Q.HQAViT(Q.HQAViTConfig()).cuda().train().set_precision(prec)
opt = Q.FusedAdamW(m.named_parameters(), lr=6e-4, betas=(0.95, 0.999), weight_decay=0.06, max_grad_norm=0.5)
x = torch.randn(8, 3, 32, 32, device="cuda"); y = torch.randint(0, 100, (8,), device="cuda")
for _ in range(2):
    opt.zero_grad()
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=prec == "bf16"):
        lo = m(x)
    loss = Q.cross_entropy(lo.float(), y, label_smoothing=0.12); loss.backward(); opt.clip(); opt.step()
torch.cuda.synchronize(); print("loss", loss.item())
PY
for prec in bf16 fp32; do
  timeout 420 compute-sanitizer --tool memcheck --print-limit 20 python /tmp/san.py $prec > gpurun_out/r2/sanitizer_memcheck_$prec.log 2>&1; echo "memcheck $prec rc=$?"; tail -3 gpurun_out/r2/sanitizer_memcheck_$prec.log
done
timeout 420 compute-sanitizer --tool racecheck --print-limit 20 python /tmp/san.py bf16 > gpurun_out/r2/sanitizer_racecheck_bf16.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/r2/sanitizer_racecheck_bf16.log
