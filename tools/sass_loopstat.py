import re,sys
from collections import Counter
def analyze(path, key=(sys.argv[2] if len(sys.argv)>2 else 'ILi0E')):
    txt=open(path).read().split("Function :")
    for blk in txt[1:]:
        name=blk.split('\n')[0].strip()
        if key not in name: continue
        ins=[]
        for l in blk.split('\n'):
            m=re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*?);', l)
            if m: ins.append((int(m.group(1),16), m.group(2).strip()))
        idx={a:i for i,(a,_) in enumerate(ins)}
        print(name, len(ins), 'instr;', sum('STL' in x[1] for x in ins),'STL', sum('LDL' in x[1] for x in ins),'LDL')
        for i,(a,t) in enumerate(ins):
            m=re.search(r'BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)', t)
            if m:
                tgt=int(m.group(1),16)
                if tgt<a and tgt in idx:
                    body=[x[1] for x in ins[idx[tgt]:i+1]]
                    nidp=sum('IDP' in x for x in body)
                    if nidp>=32 and len(body)<330:
                        ds=[]
                        for k,t2 in enumerate(body):
                            mm=re.match(r'LDS\.64 (R\d+),',t2)
                            if mm:
                                r=int(mm.group(1)[1:]); regs={f'R{r}',f'R{r+1}'}
                                for j in range(k+1,len(body)):
                                    ops=body[j].split(None,1)[1] if ' ' in body[j] else ''
                                    srcs=ops.split(',')[1:]
                                    if any(re.sub(r'\.reuse','',x.strip()).strip('[]') in regs for x in srcs):
                                        ds.append(j-k); break
                        c=Counter(re.sub(r'^@!?U?P\d+\s+','',x).split()[0].split('.')[0] for x in body)
                        print('  loop',hex(tgt),len(body),'instr; LDS->use dist avg %.1f min %d'%(sum(ds)/max(1,len(ds)),min(ds) if ds else 0), 'S2R',c.get('S2R',0),'IMAD',c.get('IMAD',0),'IADD3',c.get('IADD3',0), 'LDL', c.get('LDL',0))
analyze(sys.argv[1])
