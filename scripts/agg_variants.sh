for v in 1 0 2 3 4 5; do GRAPES_AGG_VARIANT=$v python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('variant', $v, d['ms_per_step'], d['breakdown_ms_per_step']['grapes_aggregate'], d['rooflines']['grapes_aggregate']['frac'])"; done
