for r in 1 2; do for lib in libfaster_b200.so libfq_n13.so; do
  FQ_LIB=$PWD/faster_b200/lib/$lib python bench.py --config cfg5 --no-cpu-baseline --steps 10 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib', round(d['value']/1e6,2),'M cand/s', 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'parity', d.get('parity',{}).get('flag_mismatches'))"
done; done
