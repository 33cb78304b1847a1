for ch in 8 16 31; do echo "chunk=$ch"; timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --chunk $ch 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), [round(s['ms'],2) for s in d['roofline_stages']], d['config'].get('refined_per_gpu_per_step'))"; done
