for s in 4; do echo "=== streams=$s"; timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --streams $s 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'path frac', round(d['whole_path_tensor_frac'],3), d['clocks'])
"; done
