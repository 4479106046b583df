# A/B builds of the library on the same box: tools/ab.sh "<variant.so> [variant2.so ...]" [bench args]
VS=$1; shift
for i in 1 2; do
for lib in "" $VS; do
HALVA_DPA_LIB=$lib python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-variants "$@" 2>&1 | tail -1 | python -c "
import json,sys,os; d=json.loads(sys.stdin.read()); print('${lib:-default}'.split('/')[-1], round(d['ms_per_step'],3), {k:v['ms_per_step'] for k,v in d['kernels'].items() if v['ms_per_step']>0.5})"
done; done
