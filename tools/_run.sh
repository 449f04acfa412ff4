mkdir -p gpurun_out/s4
( timeout 120 python tools/stage_probe.py 4096 3
for v in wsreg wsreg7 wsd1 wsd2; do MSCKF_B200_LIB=$PWD/gpurun_ab_$v.so timeout 120 python tools/stage_probe.py 4096 3; done
MSCKF_B200_LIB=$PWD/gpurun_ab_wsreg.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "update_vs_golden or compress_invariants or perturbed or edge_cases" 2>&1 | tail -3
) > gpurun_out/s4/probe.txt 2>&1
cat gpurun_out/s4/probe.txt
