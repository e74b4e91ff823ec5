"""Malformed text / truncated .bed inputs through the rgb200 driver linked against the mock ABI (tests/mock/mock_abi.cpp) and
built with -fsanitize=address,undefined:

    g++ -O1 -g -fsanitize=address,undefined -std=c++17 -I include -o /tmp/rgb200_mock_asan regenie_b200/host/*.cpp \
        tests/mock/mock_abi.cpp -lz -lpthread -ldl
    python tools/fuzz_text_inputs.py [seed]

Phenotype, covariate, .bim, .fam and .loco files get random token edits, dropped / duplicated / emptied lines and
truncations; the .bed gets truncated.  Every run of Step 1 and Step 2 must end normally or with `ERROR: ...`, without a
sanitizer report (round 1: seeds 1-2, 240 runs each; the only finding was the handle not being released when an error
unwinds, fixed with a scope guard).
"""
import os, random, shutil, subprocess, sys
d=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'example')
T='/tmp/fzt'; shutil.rmtree(T, ignore_errors=True); os.makedirs(T)
random.seed(int(sys.argv[1]) if len(sys.argv)>1 else 1)
M='/tmp/rgb200_mock_asan'
def run(args):
    r=subprocess.run([M]+args,capture_output=True,timeout=120)
    return r.returncode, r.stderr.decode('utf-8','replace'), r.stdout.decode('utf-8','replace')
# a valid step-1 output to corrupt
rc,err,out=run(['--step','1','--bed',d+'/example_3chr','--phenoFile',d+'/phenotype.txt','--covarFile',d+'/covariates.txt','--bsize','100','--out',T+'/fit'])
assert rc==0, out[-500:]
def corrupt_text(src, dst):
    lines=open(src,'rb').read().split(b'\n')
    for k in range(random.randint(1,3)):
        i=random.randrange(len(lines)); mode=random.randrange(6)
        toks=lines[i].split()
        if mode==0 and toks: toks[random.randrange(len(toks))]=random.choice([b'NA',b'x',b'-',b'1e999',b'',b'\xff\xfe',b'9'*40])
        elif mode==1 and toks: del toks[random.randrange(len(toks))]
        elif mode==2: toks.append(b'extra')
        elif mode==3: lines[i]=b''; continue
        elif mode==4: lines.insert(i, lines[i]); continue
        elif mode==5: lines=lines[:max(1,i)]; break
        lines[i]=b' '.join(toks)
    open(dst,'wb').write(b'\n'.join(lines))
bad=0
for it in range(120):
    shutil.copy(d+'/example_3chr.bed',T+'/g.bed')
    for ext in ('.bim','.fam'): shutil.copy(d+'/example_3chr'+ext,T+'/g'+ext)
    ph,cv,pl=d+'/phenotype.txt',d+'/covariates.txt',T+'/fit_pred.list'
    which=it%6
    if which==0: corrupt_text(ph,T+'/ph.txt'); ph=T+'/ph.txt'
    elif which==1: corrupt_text(cv,T+'/cv.txt'); cv=T+'/cv.txt'
    elif which==2: corrupt_text(d+'/example_3chr.bim',T+'/g.bim')
    elif which==3: corrupt_text(d+'/example_3chr.fam',T+'/g.fam')
    elif which==4:
        corrupt_text(T+'/fit_1.loco',T+'/c_1.loco'); open(T+'/c.list','w').write('Y1 %s/c_1.loco\nY2 %s/fit_2.loco\n'%(T,T)); pl=T+'/c.list'
    else:
        b=bytearray(open(d+'/example_3chr.bed','rb').read()); b=b[:random.randrange(1,len(b))]; open(T+'/g.bed','wb').write(b)
    for step in (['--step','2','--pred',pl],['--step','1']):
        rc,err,out=run(step+['--bed',T+'/g','--phenoFile',ph,'--covarFile',cv,'--bsize','100','--out',T+'/o'])
        if 'Sanitizer' in err or 'runtime error' in err or rc<0:
            bad+=1; print(it,which,step[1],rc,err[:400].replace('\n',' | '))
    if bad>4: break
print('text fuzz done, bad',bad)
