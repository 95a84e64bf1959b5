"""Executed instructions per source line of one kernel: joins an ncu SASS page (per-address execution counts)
with `nvdisasm -g` of the same cubin (address -> file:line).  Development tool, not part of the product.

  ncu -i cap.ncu-rep --page source --csv --print-source sass > sass.csv
  cuobjdump -xelf all lib.so && nvdisasm -g -c *.cubin > lib.dis
  python tools/sass_lines.py sass.csv lib.dis 'msim_step_kernelIN4msim7WL_RaftELb0' 59.4e6 [top]
"""
import collections
import csv
import re
import sys


def main():
    sass_csv, dis, kernel_substr, events = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 60
    rows = list(csv.reader(open(sass_csv)))
    hdr = rows[1]
    ix, ax, sx = hdr.index("Instructions Executed"), hdr.index("Address"), hdr.index("Source")
    counts, text = {}, {}
    base = None
    for r in rows[2:]:
        if len(r) <= ix:
            continue
        a = int(r[ax], 16)
        base = a if base is None else base
        counts[a - base] = int(r[ix] or 0)
        text[a - base] = r[sx].strip()
    line_of, cur, inside = {}, None, False
    for ln in open(dis):
        if ln.startswith(".text."):
            inside = kernel_substr in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", ln)
        if m:
            line_of[int(m.group(1), 16)] = cur
    per_line, per_file = collections.Counter(), collections.Counter()
    total = 0
    for a, n in counts.items():
        k = line_of.get(a, ("?", 0))
        per_line[k] += n
        per_file[k[0]] += n
        total += n
    print(f"total {total}  = {total / events:.1f} instr/event; mapped {len(line_of)} of {len(counts)} addresses")
    for f, n in per_file.most_common():
        print(f"  {f:28s} {n / events:8.1f}/event {100 * n / total:5.1f}%")
    print("-- top lines")
    for (f, l), n in per_line.most_common(top):
        print(f"  {f:24s}:{l:<5d} {n / events:7.2f}/event {100 * n / total:5.2f}%")


if __name__ == "__main__":
    main()
