import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hi=[i for i,r in enumerate(rows) if "Instructions Executed" in r][0]
h=rows[hi]; ie=h.index("Instructions Executed"); ss=h.index("# Samples"); src=h.index("Source")
data=[(i,r[src],int(r[ie]),int(r[ss])) for i,r in enumerate(rows[hi+1:]) if len(r)>ie and r[ie].isdigit()]
tot=sum(d[2] for d in data); tots=sum(d[3] for d in data)
print("total inst",tot,"samples",tots, "n sass",len(data))
cur=None; start=0; acc=0; accs=0
regions=[]
for i,s,n,sm in data:
    if cur is None or abs(n-cur)>0.05*max(cur,1):
        if cur is not None: regions.append((start,i-1,cur,acc,accs))
        cur=n; start=i; acc=0; accs=0
    acc+=n; accs+=sm
regions.append((start,len(data)-1,cur,acc,accs))
thr=float(sys.argv[2]) if len(sys.argv)>2 else 0.01
for r in regions:
    if r[3]>thr*tot or r[4]>thr*tots:
        ops=" ".join(d[1].strip().split()[0 if not d[1].strip().startswith('@') else 1] for d in data[r[0]:min(r[1]+1,r[0]+12)])
        print(f"sass {r[0]:4d}-{r[1]:4d} exec/inst {r[2]:9d} inst {100*r[3]/tot:5.1f}% samples {100*r[4]/tots:5.1f}%  {ops[:150]}")
