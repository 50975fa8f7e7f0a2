import csv,sys,subprocess
rep,kern=sys.argv[1],sys.argv[2]
topn=int(sys.argv[3]) if len(sys.argv)>3 else 40
out=subprocess.run(['ncu','-i',rep,'--page','source','--print-source','cuda,sass','--csv','--kernel-name','regex:'+kern]+sys.argv[4:],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=None; cur=None; res=[]
for r in rows:
    if not r: continue
    if r[0]=='File Path': cur=r[1].split('/')[-1]; continue
    if r[0]=='Line No': hdr=r; continue
    if hdr is None or len(r)<len(hdr): continue
    if r[0]=='' : continue  # sass rows
    try:
        inst=float(r[hdr.index('Instructions Executed')]); samp=float(r[hdr.index('# Samples')])
    except: continue
    res.append((inst,samp,cur,r[0],r[1].strip()[:120]))
tot=sum(x[0] for x in res) or 1; ts=sum(x[1] for x in res) or 1
print('total warp inst %.3g samples %d'%(tot,ts))
for x in sorted(res,reverse=True)[:topn]:
    print('%5.1f%% inst %5.1f%% samp %s:%s  %s'%(100*x[0]/tot,100*x[1]/ts,x[2],x[3],x[4]))
