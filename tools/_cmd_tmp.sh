for m in "stereo 1184 88200" "sr 1184 44100" "denoiser 1184 44100"; do
set -- $m
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2N_$1.csv python tools/probe_forward.py $1 $2 $3 1 > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.DictReader(l for l in open('gpurun_out/r2N_$1.csv') if l.startswith('"')))
per={}; order=[]
for r in rows:
    k=r['ID']
    if k not in per: per[k]={'name':r['Kernel Name']}; order.append(k)
    v=float(r['Metric Value']); u=r['Metric Unit']
    if r['Metric Name'].startswith('dram'): v*={'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9}.get(u,1)
    else: v/= {'ns':1e6,'us':1e3,'ms':1}.get(u,1)
    per[k][r['Metric Name']]=v
order=order[len(order)//2:]
print('$1', 'total %.2f ms' % sum(per[k]['gpu__time_duration.sum'] for k in order))
for k in order:
    x=per[k]; gb=(x.get('dram__bytes_read.sum',0)+x.get('dram__bytes_write.sum',0))/1e9
    print('  %8.3f ms %7.2f GB %5.2f TB/s  %s' % (x['gpu__time_duration.sum'], gb, gb/x['gpu__time_duration.sum'], x['name'][:60]))
PY
done
