import json,sys
for f in sys.argv[1:]:
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l)
            r=d.get('roofline',{})
            print(f, 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value']/1e6,2),'M', 'dom', r.get('kernel'), 'frac', r.get('frac'), 'step_frac_burst', r.get('step_frac_burst'))
            print('  phases', {k:round(v,4) for k,v in r.get('phases_ms',{}).items()})
            if 'sustained' in d and d['sustained']: print('  sustained', d['sustained'].get('ms_per_step'), d['sustained'].get('step_frac_sustained'))
            if 'other_activation_format' in d: print('  other', d['other_activation_format'].get('format'), d['other_activation_format'].get('ms_per_step'))
            print('  stats', d.get('final_step_stats'))
