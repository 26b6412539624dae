for c in 1 0 1 0; do
  VSRD_CULL=$c timeout 300 python bench.py --steps 5 --warmup 3 --skip-cpu-baseline --main-py-steps 0 2>/dev/null | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.readline())
f=d['frames']
print('cull $c: sampler %.2f s balanced %.2f s  frames/hour %.0f / %.0f'%(f['sampler']['seconds'],f['balanced']['seconds'],f['sampler']['frames_per_hour'],f['balanced']['frames_per_hour']))"
done
