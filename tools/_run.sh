mkdir -p gpurun_out/s8
( timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "propagate" 2>&1 | tail -3
timeout 120 python tools/prop_probe.py 4096 9 5
timeout 120 python tools/prop_probe.py 4096 20 3
timeout 120 python tools/prop_probe.py 256 9 5
) > gpurun_out/s8/probe.txt 2>&1
cat gpurun_out/s8/probe.txt
