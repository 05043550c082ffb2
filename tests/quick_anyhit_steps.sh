# node steps per triangle phase for occlusion rays: shadow-ray kernel time per house frame (ncu launch list) and microbench set C
R="host/_build/spcbpt_render --cache data/_ref/house.spcscene --dim=1920x1080 --no-images --quiet --no-pipeline --frames 2"
for v in any1 any2 any3; do
  mkdir -p /tmp/l_$v; cp alt_lib/$v.so /tmp/l_$v/libspcbpt_b200.so
  LD_LIBRARY_PATH=/tmp/l_$v ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_trace_persist" -c 120 --csv --log-file gpurun_out/any_$v.csv $R > /dev/null 2>&1
  python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/any_$v.csv')))
hdr=None;tot={'<1':0.0,'<0':0.0}
for r in rows:
    if len(r)>5 and r[0]=='ID': hdr=r; continue
    if hdr and len(r)==len(hdr) and r[0].isdigit():
        d=dict(zip(hdr,r)); n=d['Kernel Name']; v=float(d['Metric Value'].replace(',',''))
        if 'k_trace_persist<(bool)1' in n or 'k_trace_persist<1' in n: tot['<1']+=v
        else: tot['<0']+=v
print('$v', 'shadow ms (120 launches)', round(tot['<1']/1e6,3), 'closest ms', round(tot['<0']/1e6,3))
PY
done
