"""Print the interesting parts of a bench.py JSON line."""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d['config']['name'], 'value', round(d['value'], 1), d['unit'], 'ms/step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'], 1),
      'e2e_fp32', round(d['e2e'].get('fp32_frames', {}).get('value', 0), 1), 'launches', d['gpu_launches'], 'clocks', d.get('clocks'))
if 'sustained' in d: print('sustained', d['sustained'])
r = d['roofline']
print('top', r.get('kernel'), 'frac', r.get('frac'), 'whole', r['whole_step'])
for e in r.get('per_kernel', [])[:30]: print('  ', e)
print('breakdown', r.get('kernel_breakdown_ms_per_step'))
for n, c in d.get('configs', {}).items():
    print(n, round(c['value'], 1), 'ms', round(c['ms_per_step'], 3), 'e2e', round(c['e2e']['value'], 1), 'fp32', round(c['e2e'].get('fp32_frames', {}).get('value', 0), 1), 'frac', round(c['roofline']['whole_step']['frac'], 4), 'launches', c['gpu_launches'])
if d.get('cpu_baseline'): print('cpu', d['cpu_baseline'])
if d.get('dp_selfcheck'): print('dp', d['dp_selfcheck'])
