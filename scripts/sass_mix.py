"""Instruction mix of the loops of a kernel in npdr_b200/libnpdr_b200.so (cuobjdump -sass): how the "18 FP64 + 10.6 other
instructions per evaluation" figures in DESIGN.md / bench.py were counted.
usage: python scripts/sass_mix.py <substring of the mangled kernel name> [lo hi]   e.g. pass2_kernelILi2ELi0ELi0 (IRLS, q = 0);
with lo/hi (hex addresses of a loop) the instructions of that range are listed."""
import re,collections,sys,subprocess
import os
so=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'npdr_b200', 'libnpdr_b200.so')
pat=sys.argv[1]
out=subprocess.run(['cuobjdump','-sass',so],capture_output=True,text=True).stdout
# split functions
funcs=re.split(r'\n\s*Function : ',out)
for f in funcs[1:]:
    name=f.split('\n',1)[0]
    if pat not in name: continue
    ins=[]
    for l in f.split('\n'):
        m=re.search(r'/\*([0-9a-f]{4,5})\*/\s+(.*?);',l)
        if m: ins.append((int(m.group(1),16),m.group(2)))
    # backward branches = loops
    loops=[]
    for a,t in ins:
        m=re.search(r'BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)',t)
        if m and int(m.group(1),16)<a: loops.append((int(m.group(1),16),a))
    print(name, len(ins),'instr')
    for lo,hi in loops:
        body=[(a,t) for a,t in ins if lo<=a<=hi]
        cnt=collections.Counter()
        for a,t in body:
            toks=t.split()
            op=toks[1] if toks[0].startswith('@') else toks[0]
            cnt[op.split('.')[0]]+=1
        fp=cnt['DFMA']+cnt['DADD']+cnt['DMUL']
        if len(body)>60: print(hex(lo),hex(hi),len(body),'fp64',fp,dict(cnt.most_common()))
    if len(sys.argv)>2:
        lo,hi=int(sys.argv[2],16),int(sys.argv[3],16)
        for a,t in ins:
            if lo<=a<=hi: print(hex(a),t)
