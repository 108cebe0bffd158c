"""Per kernel of the built objects: occurrences of the SASS mnemonics that identify the hardware path (cuobjdump -sass)."""
import collections, glob, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ['UTCHMMA', 'UTCBAR', 'LDTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'SYNCS', 'HMMA', 'LDSM', 'LDGSTS', 'FFMA2', 'FMUL2', 'FADD2', 'REDG', 'RED', 'ATOMG', 'MUFU']
print('SASS evidence (cuobjdump -sass on the sm_100a objects of this commit): per kernel, occurrences of the mnemonics that identify the\n'
      'hardware path (B200_PROFILING.md): UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM = tcgen05.ld, UTMALDG = TMA tensor load,\n'
      'UTMASTG = TMA tensor store, SYNCS = mbarrier, HMMA = mma.sync, LDSM = ldmatrix, LDGSTS = cp.async, FFMA2/FMUL2/FADD2 = packed fp32x2,\n'
      'REDG/ATOMG = global reductions / atomics.\n')
for obj in sorted(glob.glob(os.path.join(ROOT, 'leod_b200', 'csrc', '.obj', '*.o'))):
    out = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True).stdout
    cur, counts = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
        if m and cur:
            op = m.group(1)
            if op in KEYS:
                counts[cur][op] += 1
    print(os.path.basename(obj))
    for fn, c in counts.items():
        if not c:
            continue
        name = subprocess.run(['c++filt', fn], capture_output=True, text=True).stdout.strip()
        name = re.sub(r'\(anonymous namespace\)::', '', name).replace('void ', '')[:88]
        print(f'  {name:88s} ' + ' '.join(f'{k}={c[k]}' for k in KEYS if c[k]))
