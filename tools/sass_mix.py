"""Static instruction mix of the library's kernels from `cuobjdump -sass` (no GPU needed):
    python tools/sass_mix.py [sca_b200/libscab200.so] > profiles/rNN_sass_mix.md
Per kernel: SASS instructions, code bytes (16 B each), and how many of them are fp64 arithmetic (DADD/DMUL/DFMA/DSETP/
MUFU.*64), fp32 arithmetic, integer / logic, constant-bank loads (kernel
parameters), global memory, local memory (spills / thread-local arrays), shared memory, warp primitives (SHFL/VOTE/REDUX/MATCH),
branches.  It is a picture of code size and of what a kernel is made of, not of what it executes."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else 'sca_b200/libscab200.so'
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
CLASSES = [('fp64', r'^(DADD|DMUL|DFMA|DSETP|DMNMX|MUFU\.\w*64|F2F\.F64|F2F\.F32\.F64|I2F\.F64|F2I\.\w*F64)'),
           ('fp32', r'^(FADD|FMUL|FFMA|FSETP|FMNMX|FSEL|MUFU|FCHK|F2F|I2F|F2I|FRND)'),
           ('int', r'^(IADD|IMAD|ISETP|LOP|SHF|LEA|VIMNMX|IMNMX|SEL|PRMT|POPC|FLO|BREV|MOV|IABS|PLOP|P2R|R2P|CS2R|S2R|UMOV|ULDC|UIADD|ULOP|USHF|UISETP|UIMAD|ULEA|UFLO|UPOPC|USEL|UPRMT|R2UR|S2UR|VOTEU|UP2UR)'),
           ('const', r'^(LDC|ULDC|LDCU)'),
           ('gmem', r'^(LDG|STG|LD\.|ST\.|ATOMG|RED|ATOM|CCTL)'),
           ('local', r'^(LDL|STL)'),
           ('smem', r'^(LDS|STS|ATOMS|LDSM)'),
           ('warp', r'^(SHFL|VOTE|REDUX|MATCH|WARPSYNC|BAR|ELECT)'),
           ('branch', r'^(BRA|BRX|JMP|CALL|RET|EXIT|BSSY|BSYNC|BREAK|YIELD|NANOSLEEP|BPT|KILL|BMOV)')]
mix = collections.OrderedDict()
name = None
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        name = m.group(1)
        mix[name] = collections.Counter()
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m and name:
        op = m.group(1)
        mix[name]['total'] += 1
        for cls, rx in CLASSES:
            if re.match(rx, op):
                mix[name][cls] += 1
                break
        else:
            mix[name]['other'] += 1
demangle = subprocess.run(['c++filt'] + list(mix), capture_output=True, text=True).stdout.splitlines()
print('| kernel | SASS instructions | code KB | fp64 | fp32 | int / logic | const | global mem | local mem | shared mem | warp prims | branch | other |')
print('|---|---|---|---|---|---|---|---|---|---|---|---|---|')
for (k, c), d in sorted(zip(mix.items(), demangle), key=lambda x: -x[0][1]['total']):
    d = re.sub(r'\(.*', '', d).replace('void ', '')
    print('| `%s` | %d | %.1f | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d |' % (
        d, c['total'], c['total'] * 16 / 1024.0, c['fp64'], c['fp32'], c['int'], c['const'], c['gmem'], c['local'], c['smem'], c['warp'],
        c['branch'], c['other']))
