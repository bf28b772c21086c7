import sys, ast, collections, re
rows=[]
for l in open(sys.argv[1]):
    p=l.rstrip("\n").split("|"); a,loc=p[0],p[1]; st=p[-1]; s=p[-2]; n=p[-3]; op="|".join(p[2:-3])
    rows.append((int(a,16),loc,op,int(n),int(s),ast.literal_eval(st)))
# find execution-count regimes: consumers execute with count ~ tiles*7 ; io with loop counts
tot_i=sum(r[3] for r in rows); tot_s=sum(r[4] for r in rows)
print("total inst",tot_i,"samples",tot_s)
# classify by source line for kernel7.cu lines 179-385 => IO ; need address ranges. print regions of contiguous addresses by role guess
def role(loc):
    f,ln=loc.rsplit(":",1); ln=int(ln)
    if f=="scan3d_fused_kernel7.cu":
        if 179<=ln<=385: return "io"
        if ln>=388: return "cons"
        return "pre"
    return None
# propagate role by address neighborhood: assign inlined code (math headers) to role of nearest kernel7.cu line before it
cur="pre"; roles=[]
for r in rows:
    ro=role(r[1])
    if ro: cur=ro
    roles.append(cur)
agg=collections.defaultdict(lambda:[0,0,collections.Counter()])
for r,ro in zip(rows,roles):
    a=agg[ro]; a[0]+=r[3]; a[1]+=r[4]; a[2].update(r[5])
for k,(i,s,c) in agg.items():
    print(k,"inst %.1f%%"%(100*i/tot_i),"samples %.1f%%"%(100*s/tot_s), [(x,round(100*y/s,1)) for x,y in c.most_common(9)])
# opcode class mix for consumers
mix=collections.Counter()
for r,ro in zip(rows,roles):
    if ro=="cons":
        op=r[2].split()[0]
        if op.startswith("@"): op=r[2].split()[1]
        mix[op.split(".")[0]]+=r[3]
ci=agg["cons"][0]
print("consumer opcode mix:")
for op,n in mix.most_common(40): print("  %-10s %6.2f%%"%(op,100*n/ci))
