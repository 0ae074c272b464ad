"""Static loop report of the built library: instructions, spills, fp64 ops and loads per SASS loop.

    python scripts/sass_loops.py <regex on the function name> [--fresh]
"""
import re,sys,subprocess,os
os.makedirs('/tmp/sass',exist_ok=True)
os.chdir('/tmp/sass')
for f in os.listdir('.'):
    if f.endswith('.cubin') and '--fresh' in sys.argv: os.remove(f)
if '--fresh' in sys.argv: subprocess.run('cuobjdump -xelf all /root/repo/coveragecontrol_b200/libccb200.so >/dev/null 2>&1',shell=True)
if '--fresh' in sys.argv: subprocess.run('nvdisasm -g -c *.cubin > all.sass 2>/dev/null',shell=True)
pat=sys.argv[1]
lines=open('/tmp/sass/all.sass').read().split('\n')
i=0
while i<len(lines):
    m=re.match(r'\s*\.text\.(\S+):',lines[i]) or re.match(r'(\$\S+):',lines[i])
    if m and re.search(pat,m.group(1)):
        name=m.group(1); j=i+1
        while j<len(lines) and not re.match(r'\s*\.text\.',lines[j]) and not re.match(r'\$\S+:',lines[j]): j+=1
        fn=lines[i:j]
        labels={}; ins=[]; curline=None
        for l in fn:
            mm=re.match(r'\s*(\.L_x_\d+):',l)
            if mm: labels[mm.group(1)]=len(ins); continue
            mm=re.search(r'//## File "([^"]+)", line (\d+)',l)
            if mm: curline=(mm.group(1).split('/')[-1][:8],int(mm.group(2))); continue
            mm=re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*?);',l)
            if mm: ins.append((mm.group(2),curline))
        print('==',name,len(ins),'instructions')
        for idx,(t,ln) in enumerate(ins):
            mm=re.search(r'BRA\S*\s+.*?(\.L_x_\d+)',t)
            if mm and mm.group(1) in labels and labels[mm.group(1)]<=idx:
                s=labels[mm.group(1)]; body=ins[s:idx+1]
                if len(body)<12: continue
                cnt=lambda rx: sum(1 for b in body if re.search(rx,b[0]))
                print('  loop %d-%d len %d  spills %d dp %d f2f %d ldg %d lds %d  @%s'%(s,idx,len(body),cnt(r'\b(STL|LDL)'),cnt(r'\bD(ADD|FMA|MUL|SETP)'),cnt(r'\bF2F'),cnt(r'\bLDG'),cnt(r'\bLDS'),ln))
        i=j
    else: i+=1
