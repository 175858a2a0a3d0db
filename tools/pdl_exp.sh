timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for o in pdl=1 pdl=0; do echo "== $o"; DHG_OPTS=$o python bench.py --no-cpu-baseline --steps 3 2>&1 | grep -o '"us_per_denoiser_step": [0-9.]*'; done
