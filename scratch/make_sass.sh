#!/bin/bash
# SASS evidence for profiles/: per kernel, the instruction classes that matter (FP64 tensor core,
# bulk async copies + mbarrier, cp.async, warp reductions) with counts, and a short excerpt.
# usage: scratch/make_sass.sh r02
cd /root/repo
R=${1:-r02}
SO=pywfo_b200/libwfo_b200.so
cuobjdump -sass $SO > /tmp/wfo_all.sass
python - "$R" <<'PY'
import re, sys, collections
R = sys.argv[1]
txt = open('/tmp/wfo_all.sass').read()
funcs = re.split(r'\n\s*Function : ', txt)[1:]
groups = {"t1_fused": "t1_fused_kernel", "gemm": "gemm_f64_kernel", "t0": "ll_kernel|wpm_kernel|warp_kernel", "base": "base_inverse|pivoted_inverse"}
pat = re.compile(r'\b(DMMA\.\w+|UBLKCP\.\S+|SYNCS\.\S+|LDGSTS\.\S+|CREDUX\.\S+|REDUX\.\S+|DFMA|LDS\.128|STS\.128|BAR\.SYNC\S*|UTMALDG\S*|UTCHMMA\S*)')
for g, rx in groups.items():
    out = []
    for f in funcs:
        name = f.split('\n', 1)[0].strip()
        if not re.search(rx, name):
            continue
        c = collections.Counter(m.group(1).split('.')[0] + ('.' + m.group(1).split('.')[1] if '.' in m.group(1) else '') for m in pat.finditer(f))
        n_inst = len(re.findall(r'/\*[0-9a-f]{4,5}\*/', f))
        out.append(f"{name}\n    instructions: {n_inst}   " + "  ".join(f"{k} x{v}" for k, v in sorted(c.items())))
        shown = set()
        for line in f.split('\n'):
            m = pat.search(line)
            if m and m.group(1).split('.')[0] in ("DMMA", "UBLKCP", "SYNCS", "LDGSTS", "CREDUX") and m.group(1) not in shown and len(shown) < 8:
                shown.add(m.group(1))
                out.append("        " + line.strip()[:110])
    open(f'profiles/{R}_sass_{g}.txt', 'w').write(
        f"# cuobjdump -sass pywfo_b200/libwfo_b200.so, kernels matching /{rx}/ : instruction-class counts and one line per class\n" + "\n".join(out) + "\n")
    print(g, len(out))
PY
