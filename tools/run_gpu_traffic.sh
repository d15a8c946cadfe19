cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_tc -s 564 -c 188 --csv --log-file gpurun_out/gemm_traffic_v2.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/gemm_traffic_v2.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/gemm_traffic_v2.csv')))
hi=[i for i,r in enumerate(rows) if "Kernel Name" in r][0]; h=rows[hi]
idi,mn,mv,mu=h.index("ID"),h.index("Metric Name"),h.index("Metric Value"),h.index("Metric Unit")
d={}
for r in rows[hi+1:]:
    if len(r)>mv: d.setdefault(r[idi],{})[r[mn]]=(float(r[mv].replace(',','')), r[mu])
def tot(k):
    s=0
    for v in d.values():
        x,u=v[k]; s+= x*{'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9,'ns':1e-6,'us':1e-3,'ms':1,'usecond':1e-3,'nsecond':1e-6,'msecond':1}.get(u,1)
    return s
print("launches",len(d),"read GB %.2f"%(tot('dram__bytes_read.sum')/1e9),"write GB %.2f"%(tot('dram__bytes_write.sum')/1e9),"time ms %.2f"%tot('gpu__time_duration.sum'))
PY
