import subprocess,csv,sys
rep=sys.argv[1]; kern=sys.argv[2]; n=int(sys.argv[3]) if len(sys.argv)>3 else 22
out=subprocess.run(['ncu','-i',rep,'--page','source','--print-source','cuda,sass','--csv','--kernel-name','regex:'+kern],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=None;cur=None;res=[]
for r in rows:
    if not r: continue
    if r[0]=='File Path': cur=r[1].split('/')[-1]; continue
    if r[0]=='Line No': hdr=r; continue
    if hdr is None or len(r)<len(hdr) or r[0]=='': continue
    try: samp=float(r[hdr.index('# Samples')]); inst=float(r[hdr.index('Instructions Executed')])
    except: continue
    res.append((samp,inst,cur,r[0],r[1].strip()[:110]))
ts=sum(x[0] for x in res) or 1; ti=sum(x[1] for x in res) or 1
print('samples',ts,'inst',ti)
for x in sorted(res,reverse=True)[:n]: print('%5.1f%% samp %5.1f%% inst %s:%s %s'%(100*x[0]/ts,100*x[1]/ti,x[2],x[3],x[4]))
