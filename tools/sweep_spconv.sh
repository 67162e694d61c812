for cfg in "1 48 48" "1 96 48 3 3" "2 96 96" "2 192 96 3 3" "3 192 192" "3 384 192 3 3" "4 384 384" "4 768 384 3 3"; do
  python tools/run_spconv.py $cfg
done
