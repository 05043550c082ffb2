# per-kernel time of the eye pass on the shipped scene (ncu launch list restricted to the render kernels; 2-3 frames)
R="host/_build/spcbpt_render --cache data/_ref/house.spcscene --dim=1920x1080 --no-images --quiet --no-pipeline --frames 3"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_eye|k_trace_persist|k_accumulate" -c 420 --csv --log-file gpurun_out/kt.csv $R > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/kt.csv')))
hdr=None;data=[]
for r in rows:
    if len(r)>5 and r[0]=='ID': hdr=r; continue
    if hdr and len(r)==len(hdr) and r[0].isdigit(): data.append(dict(zip(hdr,r)))
names=[d['Kernel Name'].split('(')[0].replace('spc::','').replace('void ','') for d in data]
vals=[float(d['Metric Value'].replace(',','')) for d in data]
idx=[i for i,n in enumerate(names) if n=='k_eye_init']
for f in range(len(idx)-1):
    s,e=idx[f],idx[f+1]
    agg=collections.Counter()
    for i in range(s,e): agg[names[i]]+=vals[i]
    print('frame',f,'eye pass ms %.2f'%(sum(agg.values())/1e6), {k:round(v/1e6,2) for k,v in agg.most_common(9)})
    if f==1: print('  bounce0:', [(names[i][:16], round(vals[i]/1e3)) for i in range(s,s+8)])
PY
