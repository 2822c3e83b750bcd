for c in 4 5; do (timeout 600 python bench.py --config $c --steps 3 --no-cpu-baseline 2>/dev/null | tail -1 | cut -c1-120); done
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3)
