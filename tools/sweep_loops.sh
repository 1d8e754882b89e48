#!/bin/bash
# Diagnostic: loop-kernel tuning hooks against stage times / repair counts (GPU box). Usage: tools/sweep_loops.sh [log2 samples]
LG=${1:-29}
run() { echo "== $*"; env "$@" python tools/profile_step.py c3 $LG 2 2>&1 | tail -1 | python -c "
import sys,ast,re
l=sys.stdin.read()
m=re.search(r'timing (\{.*?\}) stats (\{.*\})',l)
t=ast.literal_eval(m.group(1)); s=ast.literal_eval(m.group(2))
print('   costas %.3f mm %.3f fir %.3f vit %.3f sum %.3f | repairs %s unconv %s/%s' % (t['costas'],t['mm'],t['agc_fir'],t['viterbi'],t['stages_sum'], s['demod']['repairs'], s['demod']['costas_unconverged'], s['demod']['mm_unconverged']))"; }
run X=0
run B200_SEG_CTAS=2
run B200_SEG_CTAS=4
run B200_MM_WARMUP_SCALE=0.8
run B200_COSTAS_WARMUP_SCALE=0.8
