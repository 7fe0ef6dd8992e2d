for m in 256 64 16 2; do
echo "== DEVICE_SCALAR_MIN=$m"
python tools/bench_ac20.py --log2n 7 10 16 --repeat 3 --scalar-min $m 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['N'], round(d['prove_s']*1e3, 2), round(d['verify_s']*1e3, 2))
"
done
