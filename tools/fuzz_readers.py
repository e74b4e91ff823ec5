"""Corruption fuzz of the host .pgen and .bgen readers (host/pgen.cpp, host/bgen.cpp) through the host probe built with
-fsanitize=address,undefined:

    g++ -O1 -g -fsanitize=address,undefined -std=c++17 -o /tmp/probe_asan regenie_b200/host/probe/probe_main.cpp \
        regenie_b200/host/{bgen,bt_null,data,output,pgen}.cpp -lz -lpthread -ldl
    python tools/fuzz_readers.py [seed]

Every damaged file (1-4 bit flips, every tenth one also truncated) must end in a clean `ERROR: ...` or a normal exit,
never in a sanitizer report or a signal.  Round 1: seeds 3-7, 200 files per format and seed, no report after three fixes
(over-long variable-length integer, genotype value out of range in .pgen; unchecked block lengths in the .bgen header).
"""
import os, random, shutil, subprocess, sys
d=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'example')
T='/tmp/fzp'; shutil.rmtree(T, ignore_errors=True); os.makedirs(T)
random.seed(int(sys.argv[1]) if len(sys.argv)>1 else 3)
def run(args):
    r=subprocess.run(['/tmp/probe_asan']+args,capture_output=True,timeout=60)
    return r.returncode, r.stderr.decode('utf-8','replace')
def fuzz(name, raw, make, n=200):
    bad=0
    for it in range(n):
        b=bytearray(raw)
        for k in range(random.randint(1,4)):
            p=random.randrange(len(b)); b[p]^=1<<random.randrange(8)
        if it%10==0: b=b[:random.randrange(24,len(b))]
        rc,err=make(b)
        if 'AddressSanitizer' in err or 'runtime error' in err or rc<0:
            bad+=1; print(name, it, rc, err[:400].replace('\n',' | '))
            if bad>3: break
    print(name,'done bad',bad)
for ext in ('.pvar','.psam'): shutil.copy(d+'/example'+ext, T+'/x'+ext)
def mk_pgen(b):
    open(T+'/x.pgen','wb').write(b); return run(['rows','--pgen',T+'/x',T+'/o.bin'])
def mk_bgen(b):
    open(T+'/x.bgen','wb').write(b); return run(['bgen-probs',T+'/x.bgen','0','1000',T+'/o.bin'])
fuzz('PGEN', open(d+'/example.pgen','rb').read(), mk_pgen)
fuzz('BGEN', open(d+'/example.bgen','rb').read(), mk_bgen)
