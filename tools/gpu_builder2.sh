set -x
mkdir -p gpurun_out
export BLUES_B200_BUILDER=2
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "neighbour or tiled or full_size or bitwise" 2>&1 | tail -3
python -m tests.gpu_perf_probe 1 300 2>&1 | grep -E "R=1|neighbor|pair |sum" 
python -m tests.gpu_perf_probe 8 200 2>&1 | grep -E "R=8|neighbor|pair |sum"
unset BLUES_B200_BUILDER
python -m tests.gpu_perf_probe 1 300 2>&1 | grep -E "R=1|neighbor|pair |sum" 
python -m tests.gpu_perf_probe 8 200 2>&1 | grep -E "R=8|neighbor|pair |sum"
