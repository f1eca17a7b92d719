# headline step time for every experiment build speech_enhancement_by_s3prl_b200/libse_exp_*.so: bash tools/sweep_lib.sh
for f in speech_enhancement_by_s3prl_b200/libse_exp_*.so ""; do
  if [ -n "$f" ]; then export SE_B200_LIB=$PWD/$f; else unset SE_B200_LIB; fi
  echo "== ${f:-default}"
  timeout 300 python bench.py --steps 200 --warmup 20 --skip-configs --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline']['kernel_ms'])"
done
