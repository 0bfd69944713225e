"""Join an `ncu --page source --csv` SASS table with `nvdisasm --print-line-info` of the same cubin: executed warp
instructions and stall samples per SOURCE LINE of one kernel.
  python tools/sass_lines.py <ncu_source.csv> <nvdisasm_all.txt> <mangled-name-substring> [source file to quote]"""
import csv, re, sys
from collections import defaultdict

def main(src_csv, dis_txt, fn, quote=None):
    rows = list(csv.reader(open(src_csv)))
    t = [i for i, r in enumerate(rows) if 'Source' in r][0]
    h = rows[t]; si, col, sc = h.index('Source'), h.index('Instructions Executed'), h.index('# Samples')
    sass = []
    for r in rows[t + 1:]:
        try: sass.append((r[si].strip(), int(r[col]), int(r[sc])))
        except Exception: pass
    cur, dis, on = None, [], False
    for line in open(dis_txt):
        if line.startswith('//--------------------- .text.'):
            on = fn in line
            continue
        if not on: continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', line)
        if m: dis.append((cur, m.group(2).strip()))
    print(f'{len(sass)} SASS rows in the ncu table, {len(dis)} in the disassembly')
    agg = defaultdict(lambda: [0, 0])
    for i in range(min(len(sass), len(dis))):
        agg[dis[i][0]][0] += sass[i][1]; agg[dis[i][0]][1] += sass[i][2]
    tot = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
    text = open(quote).read().splitlines() if quote else []
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:28]:
        q = text[k[1] - 1].strip()[:95] if k and quote and quote.endswith(k[0]) and k[1] <= len(text) else ''
        print(f'{100 * v[0] / max(tot, 1):5.1f}% instr {100 * v[1] / max(ts, 1):5.1f}% samples  {k}  {q}')

if __name__ == '__main__':
    main(*sys.argv[1:5])
