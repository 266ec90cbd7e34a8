mkdir -p gpurun_out/r2
O=gpurun_out/r2
python tools/fmid_probe.py > $O/fmid_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ffn_mid -s 6 -c 4 -o $O/fmid_final -f python tools/fmid_probe.py > $O/fmid_ncu.log 2>&1
python tools/gemm_one.py 75776 576 192 > $O/gemm_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:tc_gemm_nt -s 2 -c 1 -o $O/gemm_qkv_final -f python tools/gemm_one.py 75776 576 192 > $O/gemm_ncu.log 2>&1
cat $O/fmid_plain.log $O/gemm_plain.log; tail -1 $O/fmid_ncu.log; tail -1 $O/gemm_ncu.log
