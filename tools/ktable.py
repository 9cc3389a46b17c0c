import json,sys
d=json.load(open(sys.argv[1] if len(sys.argv)>1 else 'gpurun_out/kernels_b1024.json'))
print('graph ms/step %.3f  img/s %.0f'%(d['ms_per_step_graph'], d['img_per_s']))
tot=sum(r['ms_per_step'] for r in d['kernels'])
print('sum eager ms/step %.3f'%tot)
for r in d['kernels']:
    rf=r.get('roofline') or {}
    if r['ms_per_step']<0.02: continue
    print(f"{r['op']:16s} {str(r['shape']):26s} {' '.join(r['flags'])[:18]:18s} n={r['calls_per_step']:2d} {r['ms_per_call']*1e3:7.1f} us {r['ms_per_step']:6.3f} ms {100*r['ms_per_step']/tot:5.1f}%  {rf.get('achieved',0):7.1f} {rf.get('unit','')} {rf.get('frac',0):.3f}")
