# evaluation-step time against the forced K3 run length (emit windows per run): bash tools/sweep_runlen.sh
for r in 0 6 8 10 12 14 16 20 24 28 32; do
  echo "== SE_B200_RUN_LEN=$r"; SE_B200_RUN_LEN=$r timeout 300 python tools/time_configs.py 2>&1 | tail -3 | awk '{print $1, $2, $3, $5}'
done
