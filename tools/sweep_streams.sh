for s in 1 2 3 4; do
  timeout 300 python bench.py --steps 200 --warmup 20 --skip-configs --no-cpu-baseline --streams $s 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('streams', d['config']['launch'], d['ms_per_step'], d['value'])"
done
