"""Rank the CUDA source lines of an .ncu-rep (compiled with -lineinfo, captured with --import-source on) by stall
samples: python tools/ncu_lines.py rep [top]"""
import csv, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
lines = []
for r in rows[3:]:
    if r and r[0] != '':
        try:
            lines.append((int(r[6]), r[0], r[1].strip()[:110], r[7]))
        except (ValueError, IndexError):
            pass
tot = sum(s for s, _, _, _ in lines) or 1
print('total samples', tot)
for s, ln, src, ex in sorted(lines, reverse=True)[:top]:
    print(f'{100 * s / tot:5.1f}% L{ln:>4} ex={ex:>9} {src}')
