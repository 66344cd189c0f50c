for t in "" "8,1" "4,2" "8,2" "16,1"; do
  echo "=== PAIG_TILE=$t"
  if [ -z "$t" ]; then PAIG_DEBUG=1 python bench.py --steps 3 --warmup 3 2>&1 >/dev/null | grep "cycles per op" | tail -2
  else PAIG_TILE=$t PAIG_DEBUG=1 python bench.py --steps 3 --warmup 3 2>&1 >/dev/null | grep "cycles per op" | tail -2; fi
done
