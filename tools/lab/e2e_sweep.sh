# sweep of the host pipeline knobs (chunk size in MiB, compute streams) on the default bench
for cfg in "32 4" "24 4" "48 4" "16 4" "32 3"; do set -- $cfg
  echo "chunk_mb=$1 kstreams=$2"; SVB_CHUNK_MB=$1 SVB_KSTREAMS=$2 timeout 300 python bench.py --steps 5 --no-cpu-baseline 2>&1 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('  value',round(d['value']),'e2e',round(d['e2e']['value']))"; done
