mkdir -p gpurun_out
timeout 600 python tools/ab_variants.py run 3840x2160x128 2>&1 | cut -c1-260
(timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4)
