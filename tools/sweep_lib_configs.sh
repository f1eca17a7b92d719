# evaluation-step time at the three geometries for every experiment build libse_exp_*.so: bash tools/sweep_lib_configs.sh
for f in speech_enhancement_by_s3prl_b200/libse_exp_*.so ""; do
  if [ -n "$f" ]; then export SE_B200_LIB=$PWD/$f; else unset SE_B200_LIB; fi
  echo "== ${f:-default}"
  for i in 1 2; do timeout 300 python tools/time_configs.py 2>&1 | tail -2 | awk '{printf "%s %s | ", $1, $5} END {print ""}'; done
done
