# sustained (power-capped) encoder / scorer timing through bench.py's per-stage breakdown, one line per option set
run() { PN_OPTIONS="$1" python bench.py --sequences 2048 --labels 1024 --steps 2 --warmup 1 --no-cpu-baseline --no-fast --no-parity --no-configs 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); b=d['breakdown_rank0']; print('OPT $1 encoder_ms %.1f tflops %.1f scorer_ms %.1f pairs/s %.3e step_ms %.1f'%(b['encoder_ms'],b['encoder_algorithmic_tflops'],b['pair_scorer_ms'],b['scorer_pairs_per_s'],d['ms_per_step']))
    elif 'Error' in l or 'error' in l: print(l.rstrip())
"; }
for o in "$@"; do run "$o"; done
