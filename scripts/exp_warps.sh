for r in 24; do echo "rows_per_warp=$r"; VSP_FUSED_ROWS_PER_WARP=$r timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), [round(s['ms'],2) for s in d['roofline_stages']])"; done
