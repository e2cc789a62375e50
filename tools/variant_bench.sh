for v in "$@"; do
  RCM_B200_LIB=$PWD/our_first_climate_model_b200/$v/librcm_b200.so timeout 300 python bench.py --steps 20 --warmup 3 --no-lbl --no-cpu-baseline --no-strong 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', 'step ms %.4f kernel ms %.4f e2e %.4f' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step']), 'parity flux %.2e dE %.2e dT %.2e' % (d['parity']['max_rel_flux'], d['parity']['max_rel_dE'], d['parity']['max_abs_dT']))"
done
