# GPU job: ncu launch list (durations only) of three sequential frames.  usage: bash tools/job_ncu_launches.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches_frame.csv python profiles/frame_profile.py 3 > gpurun_out/fp.log 2>&1; tail -2 gpurun_out/fp.log
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/${tag}_launches_frame.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); mi=hdr.index('Metric Name'); ii=hdr.index('ID')
data=rows[1:]
ids=sorted(set(int(r[ii]) for r in data))
per={}
for r in data: per.setdefault(int(r[ii]),{})[r[mi]]=(r[ki],r[vi])
n=len(ids)//3
for i in ids[2*n:]:
    d=per[i]; k=list(d.values())[0][0]
    print(k[:70].ljust(70), d.get('gpu__time_duration.sum',('',''))[1], d.get('smsp__inst_executed.sum',('',''))[1])
PY
