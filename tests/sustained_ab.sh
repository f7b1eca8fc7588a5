# Developer tool: same-box A/B of kernel switches on the SUSTAINED bench loop (the loop is power-capped: per-op timings of short
# runs do not predict it).  usage: bash tests/sustained_ab.sh "tag ENV=val [ENV=val]" ...   (default: the attention switches)
run() { tag=$1; shift; env "$@" python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-modes --no-eager-baseline 2>/dev/null | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('$tag', round(d['value'],2), 'ms/unet', round(d['ms_per_unet_step'],3), 'MHz', d['clocks']['sm_mhz'], 'attnTF', round(d['roofline']['attention_TFLOPs'],1))"; }
if [ $# -eq 0 ]; then set -- "default X=1" "submax1 LIDM_ATTN_SUBMAX=1" "poly2 LIDM_ATTN_POLY=2" "default X=1" "submax1 LIDM_ATTN_SUBMAX=1"; fi
for spec in "$@"; do run $spec; done
