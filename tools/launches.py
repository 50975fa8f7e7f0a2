import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: h=r; start=i; break
ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value'); idi=h.index('ID')
from collections import OrderedDict
d=OrderedDict()
for r in rows[start+1:]:
    if len(r)<=vi: continue
    d.setdefault((r[idi],r[ki].split('(')[0][-28:]),{})[r[mi]]=r[vi]
agg={}
for (i,k),m in d.items():
    t=float(m.get('gpu__time_duration.sum',0)); rd=float(m.get('dram__bytes_read.sum',0) or 0); wr=float(m.get('dram__bytes_write.sum',0) or 0)
    a=agg.setdefault(k,[0,0,0,0,[]]); a[0]+=1; a[1]+=t; a[2]+=rd; a[3]+=wr; a[4].append(round(t/1000,1))
tot=sum(a[1] for a in agg.values())
print('total us',tot/1000)
for k,a in sorted(agg.items(), key=lambda x:-x[1][1]):
    print('%-30s n=%3d  sum %8.1f us (%4.1f%%)  rd %7.1f MB wr %7.1f MB  per-launch %s'%(k,a[0],a[1]/1000,100*a[1]/tot,a[2]/1e6,a[3]/1e6,a[4][:9]))
