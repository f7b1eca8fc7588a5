# Developer tool: same-box A/B of kernel switches on the SUSTAINED bench loop (the loop is power-capped: per-op timings of short runs do not predict it).
run() { tag=$1; shift; env "$@" python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-modes --no-eager-baseline 2>/dev/null | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('$tag', round(d['value'],2), 'ms/unet', round(d['ms_per_unet_step'],3), 'MHz', d['clocks']['sm_mhz'], 'attnTF', round(d['roofline']['attention_TFLOPs'],1))"; }
run default X=1
run submax1 LIDM_ATTN_SUBMAX=1
run poly2 LIDM_ATTN_POLY=2
run poly0 LIDM_ATTN_POLY=0
run noidres LIDM_NO_IDRES=1
run default X=1
run submax1 LIDM_ATTN_SUBMAX=1
run poly2 LIDM_ATTN_POLY=2
