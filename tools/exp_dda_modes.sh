for mode in 0; do
  VDBM_DDA_MODE=$mode python bench.py --steps 20 --warmup 3 --no-cpu-baseline | python -c "
import sys,json; d=json.loads(sys.stdin.read()); k=d['roofline']['by_kernel']; print('mode $mode', 'dda_ms', round(k['raycast_dda_kernel']['ms'],3), 'upd_ms', round(k['apply_update_kernel']['ms'],3), 'ms/step', round(d['ms_per_step'],3))"
done
