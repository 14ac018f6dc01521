# A/B of two builds of the library on ONE box: bash scripts/dev/ab.sh <libA> <libB> [bench args]
A=$1; B=$2; shift 2
for rep in 1 2; do
for L in $A $B; do
  NPDRB_LIB_PATH=$PWD/$L python bench.py --no-target --no-cpu-baseline --no-variants --no-e2e "$@" 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$L', round(d['ms_per_step'],2), 'first', round(d['stages']['pass_first_ms'],2), 'irls', round(d['roofline']['avg_launch_ms'],2), 'clk', d['clocks']['sm_mhz'])"
done
done
