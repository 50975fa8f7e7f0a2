import sys,re
rows=[]
for l in open(sys.argv[1]):
    m=re.match(r'\s+(\d+)\s+(\w+)\s+([\d.]+)\s+([\d.]+)\s+([\d.]+)',l)
    if m: rows.append((int(m.group(1)),m.group(2),float(m.group(3)),float(m.group(4)),float(m.group(5))))
tile={d:e for d,st,r,s,e in rows if st=='tile'}
out=[]
for d,st,r,s,e in rows:
    if st=='geometry_back' and d>0: out.append(f"d{d}: res-prev_end {r-tile[d-1]:+.1f} start-prev_end {s-tile[d-1]:+.1f}")
print(sys.argv[1], ' | '.join(out), '| total', max(e for *_,e in rows))
