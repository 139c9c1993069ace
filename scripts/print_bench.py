import json, sys
lines = [l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')]
d = json.loads(lines[-1])
r = d['roofline']
print('ms/step %.3f  value %.2f img/s  e2e %.2f  conv frac %.3f (%.0f TF)  unet fwd frac %.3f  clocks %s' % (
    d['ms_per_step'], d['value'], d['e2e']['value'] if d.get('e2e') else float('nan'), r['frac'], r['achieved'],
    r['unet_forward']['frac'], d['clocks']))
print('other kernels ms:', {k: round(v, 3) for k, v in r['other_kernels_ms'].items()}, 'conv ms', round(r['ms_per_step_in_kernel'], 3))
