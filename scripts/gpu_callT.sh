# ncu --set full of the giant tier's kernels (dense library, 3*10^6 reads)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:huge_exec_kernel --launch-skip 10 -c 2 -f -o gpurun_out/r02ae_huge_exec python tests/scale_check.py dense 0.3 > gpurun_out/ncu_ae1.log 2>&1; echo "ncu exec rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"huge_reserve_kernel|huge_seed_kernel|huge_head_kernel" --launch-skip 1 -c 3 -f -o gpurun_out/r02ae_huge_other python tests/scale_check.py dense 0.3 > gpurun_out/ncu_ae2.log 2>&1; echo "ncu other rc=$?"
