for cfg in "8 3" "16 3" "16 4" "12 4"; do set -- $cfg; echo "chunk=$1 lanes=$2"; timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --chunk $1 --lanes $2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']))"; done
