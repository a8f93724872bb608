mkdir -p gpurun_out
python -m pytest tests/test_gpu_fused.py -x -q 2>&1 | tail -3
for regs in 64 80 96 128; do export PYSTL_B200_FUSED_REGS=$regs; echo "== regs $regs"; NTS="${NTS:-96 128 192 256 512}" bash gpu_sweep.sh; done
