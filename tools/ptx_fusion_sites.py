#!/usr/bin/env python3
"""List every PTX add/sub.f32 whose operand is produced by a non-.rn mul.f32 — the only
sites where ptxas (--fmad=true) may still contract after NVVM.  Used to derive the exact
FMA association of the reference kernels (SURVEY.md §7 hard part 1)."""
import re, sys
from collections import defaultdict
def kernels(path):
    cur=None; body=[]
    for l in open(path):
        m=re.match(r'\.visible \.entry (\S+?)\(',l)
        if m:
            if cur: yield cur,body
            cur=m.group(1); body=[]
        elif cur: body.append(l.rstrip())
    if cur: yield cur,body
def analyse(name, body):
    defs={}; uses=defaultdict(int)
    ins=[]
    for i,l in enumerate(body):
        m=re.match(r'\s+(@%p\d+\s+)?([a-z0-9.]+)\s+(.*);',l)
        if not m: continue
        op=m.group(2); args=[a.strip() for a in m.group(3).split(',')]
        ins.append((i,op,args))
        if op.endswith('f32') and args and args[0].startswith('%f'):
            defs[args[0]]=(i,op,args[1:])
        for a in args[1:]:
            for r in re.findall(r'%f\d+',a): uses[r]+=1
        if op.startswith('st.'):
            for r in re.findall(r'%f\d+',args[-1]): uses[r]+=1
    sites=[]
    for i,op,args in ins:
        if op in('add.f32','sub.f32'):
            srcs=args[1:]
            muls=[(k,s) for k,s in enumerate(srcs) if s in defs and defs[s][1]=='mul.f32']
            if muls:
                sites.append((i,op,args,[(k,s,defs[s][2],uses[s]) for k,s in muls]))
    print(f'== {name[:40]}: {len(sites)} candidate sites')
    for i,op,args,muls in sites:
        print(f'  L{i}: {op} {args}  muls={muls}')
for n,b in kernels(sys.argv[1]):
    if len(sys.argv)>2 and sys.argv[2] not in n: continue
    analyse(n,b)
