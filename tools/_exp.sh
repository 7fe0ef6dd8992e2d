python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for v in _base "" _base ""; do
export VMSM_LIB=/root/repo/verifiable_mpc_b200/libvmsm$v.so
python bench.py --steps 60 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$v', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), d['config']['result_checked_vs_known_dlog'])
"
done
unset VMSM_LIB
python tools/sweep.py --logn 14 16 18 22 --steps 30 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['log2n'], round(d['ms'], 3))
"
python tools/bench_bn256.py --log2n 14 2>&1 | cut -c1-120
