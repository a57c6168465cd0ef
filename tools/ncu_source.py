"""Map the SASS-level samples of an ncu report back to source lines with the line table of the in-tree cubin.
usage: python tools/ncu_source.py <report.ncu-rep> <mangled-substring> [kernel index in report] [top]"""
import collections, csv, io, os, re, subprocess, sys, tempfile

def line_map(so, want):
    d = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(so)], cwd=d, capture_output=True)
    maps = {}
    for f in os.listdir(d):
        if not f.endswith('.cubin'): continue
        sass = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(d, f)], capture_output=True, text=True).stdout.split('\n')
        cur = fname = line = None
        for l in sass:
            m = re.match(r'\s*\.section\s+\.text\.(\S+?),', l)
            if m: cur = m.group(1); maps[cur] = {}; continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', l)
            if m: fname = m.group(1); line = int(m.group(2)); continue
            m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*)', l)
            if m and cur: maps[cur][int(m.group(1), 16)] = (fname, line, m.group(2))
    return {k: v for k, v in maps.items() if want in k}

def main():
    rep, want = sys.argv[1], sys.argv[2]
    which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
    maps = line_map(os.environ.get('SAPIO_SO', 'sapiogenesis_b200/libsapio_b200.so'), want)
    assert len(maps) == 1, list(maps)
    mp = next(iter(maps.values()))
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    secs = out.split('"Kernel Name",')[1:]
    tag = re.search(r'ILi(\d+)E', want)
    if tag: secs = [x for x in secs if '(int)%s>' % tag.group(1) in x.split('\n')[0]]
    else: secs = [x for x in secs if (want + '(') in x.split('\n')[0]] or secs
    sec = secs[which]
    lines = sec.split('\n'); print(lines[0][:100])
    rdr = csv.reader(io.StringIO('\n'.join(lines[1:]))); hdr = next(rdr); ix = {h: i for i, h in enumerate(hdr)}
    rows = [r for r in rdr if len(r) == len(hdr)]
    base = int(rows[0][ix['Address']], 16)
    agg = collections.defaultdict(lambda: [0, 0, 0]); tot = toti = 0
    cache = {}
    for r in rows:
        f, ln, _ = mp.get(int(r[ix['Address']], 16) - base, ('?', 0, ''))
        s, ie, te = int(r[ix['# Samples']]), int(r[ix['Instructions Executed']]), int(r[ix['Thread Instructions Executed']])
        a = agg[(f, ln)]; a[0] += s; a[1] += ie; a[2] += te; tot += s; toti += ie
    print('samples', tot, 'warp instructions', toti)
    for (f, ln), a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
        text = ''
        if f and os.path.exists(f):
            if f not in cache: cache[f] = open(f).read().split('\n')
            text = cache[f][ln - 1].strip()[:105] if 0 < ln <= len(cache[f]) else ''
        print(f"{100 * a[0] / tot:5.1f}% smp {100 * a[1] / toti:5.1f}% inst thr/inst={a[2] / max(a[1], 1):5.1f} {os.path.basename(f or '?')}:{ln} {text}")

main()
