set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -q --timeout 600 -k "beam_kernels or finalize" 2>&1 | tail -25
python -m pytest tests/test_gpu_parity.py -q --timeout 900 -k "golden" 2>&1 | tail -40
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
python bench.py --steps 3 --warmup 3 --precision fp32 --no-cpu-baseline 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 --precision tf32x3 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 --precision bf16x3 --no-cpu-baseline --parity-queries 0 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 --precision bf16 --no-cpu-baseline --parity-queries 0 2>&1 | tail -3
