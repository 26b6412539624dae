for f in 0.5 0.97; do for c in 1; do
  VSRD_CULL=$c timeout 200 python bench.py --steps 12 --warmup 3 --skip-cpu-baseline --main-py-steps 0 --frames 0 --schedule-frac $f 2>/dev/null | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.readline())
k=d['kernel_ms']
print('frac $f cull $c: step %.4f ms e2e %.4f fwd_coarse %.4f fwd_fine %.4f bwd %.4f  bwd-skipped %.3f T=%.3f frac_bwd %.3f frac_fwd %.3f'%(d['ms_per_step'],d['e2e']['ms_per_step'],k['field_forward_coarse'],k['field_forward_fine'],k['field_backward'],d['culling']['skipped_fraction'],d['config']['schedule']['temperature'],d['roofline']['frac'],d['roofline']['forward_fine']['frac']))"
done; done
