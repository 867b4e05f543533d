set -x
run() { PN_OPTIONS="$1" python bench.py --sequences 2048 --labels 1024 --steps 2 --warmup 1 --no-cpu-baseline --no-fast 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); b=d['breakdown_rank0']; print('OPT $1 encoder_ms %.1f tflops %.1f scorer_ms %.1f step_ms %.1f'%(b['encoder_ms'],b['encoder_algorithmic_tflops'],b['pair_scorer_ms'],d['ms_per_step']))
"; }
run "promote_k_encoder=32,promote_k_pointwise=32"
run "promote_k_encoder=64,promote_k_pointwise=64"
run "promote_k_encoder=64,promote_k_pointwise=576"
run "promote_k_encoder=64,promote_k_pointwise=288"
run "promote_k_encoder=128,promote_k_pointwise=576"
run "promote_k_encoder=64,promote_k_pointwise=576,cta2=1"
run "promote_k_encoder=128,promote_k_pointwise=576,cta2=1"
