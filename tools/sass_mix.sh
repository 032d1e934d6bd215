#!/bin/bash
# usage: mix.sh name  -> instruction count summary of the largest loop in build/<name>.sass
f=/root/repo/sde-sim-rs_b200/build/$1.sass
python3 - "$f" <<'PY'
import re,sys,collections
lines=open(sys.argv[1]).read().splitlines()
ins=[]
for l in lines:
    m=re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);',l)
    if m: ins.append((int(m.group(1),16),m.group(2).strip()))
addr={a:i for i,(a,_) in enumerate(ins)}
loops=[]
for i,(a,t) in enumerate(ins):
    m=re.search(r'BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)',t)
    if m:
        tgt=int(m.group(1),16)
        if tgt<a and tgt in addr: loops.append((addr[tgt],i))
def nf64(l): return sum(1 for a,t in ins[l[0]:l[1]+1] if re.sub(r'^@!?U?P\d+\s+','',t).startswith(('DFMA','DMUL','DADD')))
big=[l for l in loops if nf64(l)>=60]
lo,hi=min(big,key=lambda l:l[1]-l[0]) if big else max(loops,key=lambda l:l[1]-l[0])
c=collections.Counter(); f3=0; f64=0
for a,t in ins[lo:hi+1]:
    t2=re.sub(r'^@!?U?P\d+\s+','',t)
    op=t2.split()[0]
    c[op.split('.')[0]]+=1
    if op.split('.')[0] in('DFMA','DMUL','DADD','DSETP'):
        f64+=1
        args=[x.strip() for x in t2[len(op):].split(',')]
        regs=set(re.sub(r'[-|]|\.reuse','',x) for x in args[1:] if re.match(r'^[-|]*R\d+',x))
        if len(regs)>=3: f3+=1
tot=hi-lo+1
print(f"loop instrs {tot}  fp64 {f64} (3-reg {f3})  other {tot-f64}  dispatch-cycles {tot+f64+f3}")
print(sorted(c.items(), key=lambda x:-x[1])[:24])
PY
