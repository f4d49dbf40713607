"""Sanitizer fuzz campaign over every golden stream (host half only): python tools/fuzz_goldens.py SEED MUTANTS_PER_STREAM
Needs build/asan/parse_fuzz (make build/asan/parse_fuzz); prints ok / rejected counts per 100 files and stops at the first report."""
import os, random, subprocess, sys, glob, tempfile, shutil
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); GOLD=os.path.join(ROOT,"tests/golden/streams")
names=sorted(os.path.basename(p)[:-4] for p in glob.glob(GOLD+"/*.ivf"))
seed=int(sys.argv[1]) if len(sys.argv)>1 else 1
per=int(sys.argv[2]) if len(sys.argv)>2 else 30
rng=random.Random(seed)
tmp=os.path.join(ROOT,"build/fuzz"); shutil.rmtree(tmp,ignore_errors=True); os.makedirs(tmp)
files=[]
for name in names:
    base=open(os.path.join(GOLD,name+".ivf"),"rb").read()
    for k in range(per):
        data=bytearray(base)
        kind=rng.choice(["flip","flip","flip","flip","trunc","zero","dup","ff"])
        if kind=="flip":
            for _ in range(rng.randint(1,6)): data[rng.randrange(32,len(data))]^=1<<rng.randrange(8)
        elif kind=="trunc": data=data[:rng.randrange(40,len(data))]
        elif kind=="zero":
            p=rng.randrange(44,max(45,len(data)-16)); data[p:p+16]=bytes(16)
        elif kind=="ff":
            p=rng.randrange(44,max(45,len(data)-8)); data[p:p+8]=b"\xff"*8
        else:
            p=rng.randrange(44,max(45,len(data)-64)); data[p:p+32]=data[p+32:p+64]
        f=os.path.join(tmp,f"{name}_{k}.ivf"); open(f,"wb").write(bytes(data)); files.append(f)
env=dict(os.environ, ASAN_OPTIONS="detect_leaks=0:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1")
exe=os.path.join(ROOT,"build/asan/parse_fuzz")
bad=0
for i in range(0,len(files),100):
    chunk=files[i:i+100]
    r=subprocess.run([exe]+chunk,capture_output=True,text=True,env=env,timeout=1200)
    if r.returncode!=0:
        bad+=1
        print("FAIL chunk",i,r.returncode); print(r.stderr[-3000:])
        # bisect to single file
        for f in chunk:
            r1=subprocess.run([exe,f],capture_output=True,text=True,env=env,timeout=300)
            if r1.returncode!=0:
                print("  culprit",f); shutil.copy(f, os.path.join(ROOT,"build",os.path.basename(f)+".crash")); break
        break
    else:
        print(i, r.stdout.strip())
print("done bad chunks",bad)
