"""Opcode counts per kernel from `cuobjdump -sass libmonocon_b200.so`: the mnemonics that prove a Blackwell-native kernel
(UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = cp.async.bulk.tensor, UTCBAR = tcgen05.commit, SYNCS = mbarrier)
next to the legacy ones that must be absent from the convolution path (HMMA = mma.sync)."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else 'monocon_pytorch_b200/libmonocon_b200.so'
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
OPS = ['UTCHMMA', 'UTCQMMA', 'UTMALDG', 'UTMASTG', 'LDTM', 'STTM', 'UTCBAR', 'SYNCS', 'HMMA', 'FFMA', 'LDG', 'STG', 'ATOM', 'RED', 'SHFL']
cur, counts, order = None, collections.defaultdict(collections.Counter), []
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1)
        order.append(cur)
        continue
    if cur is None:
        continue
    m = re.search(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m:
        op = m.group(1).split('.')[0]
        if op in OPS:
            counts[cur][op] += 1
        counts[cur]['_total'] += 1
demangle = subprocess.run(['c++filt'], input='\n'.join(order), capture_output=True, text=True).stdout.splitlines()
print(f'# {lib}: SASS opcode counts per kernel (cuobjdump -sass), sm_100a')
print(f'{"kernel":92s} {"instr":>7s} ' + ' '.join(f'{o:>7s}' for o in OPS))
tot = collections.Counter()
for name, dm in zip(order, demangle):
    c = counts[name]
    short = re.sub(r'\(anonymous namespace\)::|mc::', '', dm)
    short = re.sub(r'\(.*', '', short)[:92]
    print(f'{short:92s} {c["_total"]:7d} ' + ' '.join(f'{c[o]:7d}' for o in OPS))
    tot.update(c)
print(f'{"TOTAL":92s} {tot["_total"]:7d} ' + ' '.join(f'{tot[o]:7d}' for o in OPS))
