import csv, sys, subprocess
sys.path.insert(0,'scripts')
import ncu_lines
rep, obj, kern = sys.argv[1:4]
out = subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[1]; ci={h:i for i,h in enumerate(hdr)}
data=rows[2:]
lines=ncu_lines.sass_lines(obj,kern)
S=lambda r:int(r[ci['# Samples']] or 0); I=lambda r:int(r[ci['Instructions Executed']] or 0)
tot_s=sum(S(r) for r in data); tot_i=sum(I(r) for r in data)
bars=[k for k,r in enumerate(data) if 'BAR.SYNC' in r[ci['Source']]]
prev=0
print('total samples',tot_s,'instr',tot_i)
for b in bars+[len(data)-1]:
    seg=data[prev:b+1]
    s=sum(S(r) for r in seg); i=sum(I(r) for r in seg)
    # barrier wait = samples on the instruction right after previous BAR
    print('SASS %5d-%5d (lines %s-%s): samples %5.1f%% instr %5.1f%% = %.0fk/SM ; first-instr(wait) %4.1f%%'%(prev,b,lines[prev],lines[b],100*s/tot_s,100*i/tot_i,i/148/1e3,100*S(data[prev])/tot_s))
    prev=b+1
import collections,re
def hist(lo,hi,top=18):
    h=collections.Counter()
    for r in data[lo:hi+1]:
        src=r[ci['Source']].strip()
        op=src.split()[0] if not src.startswith('@') else src.split()[1]
        op=op.split('.')[0]+('.'+op.split('.')[1] if '.' in op and op.split('.')[0] in ('LDS','LDG','STS','STG','LDL','STL','SHFL') else '')
        h[op]+=I(r)
    tot=sum(h.values())
    print('range',lo,hi,'total %.0fk/SM'%(tot/148/1e3))
    for k,v in h.most_common(top): print('   %-10s %6.1fk/SM %5.1f%%'%(k,v/148/1e3,100*v/tot))
if len(sys.argv)>4:
    for a in sys.argv[4:]:
        lo,hi=map(int,a.split('-')); hist(lo,hi)
