#!/bin/bash
# One SASS listing per __global__ of the built liblps_b200.so into profiles/sass/ (no GPU needed): tools/dump_sass.sh r02
set -u
cd "$(dirname "$0")/.."
TAG=${1:-r02}
SO=longphase-s_b200/liblps_b200.so
OUT=profiles/sass
mkdir -p $OUT
rm -f $OUT/${TAG}_*.sass
cuobjdump -sass $SO > /tmp/all_sass.txt
python - "$TAG" "$OUT" <<'PY'
import re, subprocess, sys, hashlib
tag, out = sys.argv[1], sys.argv[2]
text = open('/tmp/all_sass.txt').read()
parts = re.split(r'\n\s*Function : ', text)
index = []
for p in parts[1:]:
    mangled = p.split('\n', 1)[0].strip()
    dem = subprocess.run(['c++filt', mangled], capture_output=True, text=True).stdout.strip()
    m = re.search(r'(k_[a-z0-9_]+)(<[^>]*>)?', dem)
    if not m:
        continue      # library kernels (CUB)
    name = m.group(1) + (m.group(2) or '')
    name = re.sub(r'\(int\)', '', name)
    fn = re.sub(r'[^A-Za-z0-9_]+', '_', name).strip('_')
    body = 'Function : ' + p
    n_inst = len(re.findall(r'^\s+/\*[0-9a-f]{4,}\*/', body, re.M))
    marks = {k: len(re.findall(k, body)) for k in ('UBLKCP', 'SYNCS', 'LDGSTS', 'IDP', 'ATOM', 'RED', 'SHFL', 'LDS', 'STS', 'LDG', 'STG')}
    open(f'{out}/{tag}_{fn}.sass', 'w').write(f'// {dem}\n' + body)
    index.append((fn, n_inst, marks))
so = open('longphase-s_b200/liblps_b200.so', 'rb').read()
with open(f'{out}/{tag}_INDEX.md', 'w') as f:
    f.write(f'# SASS listings of liblps_b200.so (sha256 {hashlib.sha256(so).hexdigest()[:16]}), one file per __global__\n\n')
    f.write('| kernel | instructions | UBLKCP | SYNCS | LDGSTS | IDP | ATOM+RED | SHFL | LDS | STS | LDG | STG |\n|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n')
    for fn, n, m in sorted(index):
        f.write(f"| `{fn}` | {n} | {m['UBLKCP']} | {m['SYNCS']} | {m['LDGSTS']} | {m['IDP']} | {m['ATOM'] + m['RED']} | {m['SHFL']} | {m['LDS']} | {m['STS']} | {m['LDG']} | {m['STG']} |\n")
print(len(index), 'kernels')
PY
