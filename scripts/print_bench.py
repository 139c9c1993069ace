import json, sys
lines = [l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')]
d = json.loads(lines[-1])
r = d['roofline']
print('ms/step %.3f  value %.2f img/s  e2e %.2f  conv frac %.3f (%.0f TF)  step frac %.3f (%.0f TF)  gn %.0f GB/s  attn %.0f TF  clocks %s' % (
    d['ms_per_step'], d['value'], d['e2e']['value'] if d.get('e2e') else float('nan'), r['frac'], r['achieved'],
    r.get('step_frac', float('nan')), r.get('step_tflops', float('nan')), r.get('gn_apply_gbs') or float('nan'),
    r.get('attention_tflops') or float('nan'), d['clocks']))
