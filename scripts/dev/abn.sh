# one bench line per library given: bash scripts/dev/abn.sh lib1 lib2 ... (same box)
for L in "$@"; do
  NPDRB_LIB_PATH=$PWD/$L python bench.py --no-target --no-cpu-baseline --no-variants --no-e2e --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$L', round(d['ms_per_step'],2), 'first', round(d['stages']['pass_first_ms'],2), 'irls', round(d['roofline']['avg_launch_ms'],2), 'clk', d['clocks']['sm_mhz'])"
done
