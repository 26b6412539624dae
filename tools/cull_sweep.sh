for f in 0.34 0.5 0.6 0.7 0.8 0.9 0.97; do for c in 0 1; do
  VSRD_CULL=$c timeout 200 python bench.py --steps 12 --warmup 3 --skip-cpu-baseline --main-py-steps 0 --frames 0 --schedule-frac $f 2>/dev/null | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.readline())
k=d['kernel_ms']
print('frac $f cull $c: step %.4f ms  fwd_coarse %.4f fwd_fine %.4f bwd %.4f  bwd-skipped %.3f fwd-skipped %s T=%.3f frac_bwd %.3f frac_fwd %.3f'%(d['ms_per_step'],k['field_forward_coarse'],k['field_forward_fine'],k['field_backward'],d['culling']['skipped_fraction'],json.dumps(d['culling']['forward_pairs_skipped_fraction']),d['config']['schedule']['temperature'],d['roofline']['frac'],d['roofline']['forward_fine']['frac']))"
done; done
